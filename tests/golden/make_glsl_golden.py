#!/usr/bin/env python
"""Golden vectors for the oracle's restatement of rtcommon.glsl / restir.glsl, produced by the REFERENCE'S OWN
shader text compiled as C++ (oracle/_ref/libglsl_ref.so: oracle/make_glsl_ref.py + oracle/ref_shim/glsl_ref.cpp).
Run in the build container (needs /root/reference): python tests/golden/make_glsl_golden.py
Writes tests/golden/glsl_unit_golden.npz: per function id the inputs (floats, unsigned words) and the reference's
outputs (tests/test_oracle.py replays the inputs through orc_glsl_unit), and tests/golden/glsl_frames_golden.json:
SHA-256 digests of image / G-buffers / reservoirs / ray counts of whole frames rendered by the reference's rt.rgen."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
N_IN, N_OUT = 24, 12
NAMES = {0: "tea", 1: "randf", 2: "randu", 3: "cospow_hemisphere", 4: "triangle_sample", 5: "triangle_hit",
         6: "triangle_pdf", 7: "make_tanspace", 8: "hit_bbox", 9: "MAT_pdf", 10: "MAT_eval", 11: "MAT_sample",
         12: "res_update", 13: "power_heuristic", 14: "luma", 15: "hammersley"}


def unit_dirs(rng, n):
    d = rng.standard_normal((n, 3)).astype(np.float32)
    return d / np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)


def make_inputs(fn, n, rng):
    f = np.zeros((n, N_IN), np.float32)
    u = np.zeros((n, 3), np.uint32)
    u[:, 0] = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    if fn == 0:
        u[:, 1], u[:, 2] = rng.integers(0, 2 ** 32, n, dtype=np.uint64), rng.integers(0, 2 ** 32, n, dtype=np.uint64)
    elif fn == 2:
        u[:, 1] = rng.integers(0, 50, n)
        u[:, 2] = u[:, 1] + rng.integers(1, 2000, n)
    elif fn == 3:
        f[:, 0] = rng.random(n) * 200 + 0.5
        z = unit_dirs(rng, n)
        x = np.cross(z, unit_dirs(rng, n)).astype(np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        f[:, 1:4], f[:, 4:7], f[:, 7:10] = x, np.cross(z, x), z
    elif fn in (5, 6):
        tri = (rng.random((n, 3, 3)).astype(np.float32) - 0.5) * 2
        o = (rng.random((n, 3)).astype(np.float32) - 0.5) * 6
        b = rng.random((n, 3)).astype(np.float32)
        b /= b.sum(axis=1, keepdims=True)
        target = (tri * b[:, :, None]).sum(axis=1)
        miss = rng.random(n) < 0.3
        target[miss] += (rng.random((int(miss.sum()), 3)).astype(np.float32) - 0.5) * 3
        d = target - o
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        f[:, 0:3], f[:, 3:6], f[:, 6:15] = o, d, tri.reshape(n, 9)
    elif fn == 7:
        f[:, 0:3] = unit_dirs(rng, n)
    elif fn == 8:
        lo = (rng.random((n, 3)).astype(np.float32) - 0.5) * 2
        hi = lo + rng.random((n, 3)).astype(np.float32)
        o = (rng.random((n, 3)).astype(np.float32) - 0.5) * 6
        d = (lo + hi) * 0.5 + (rng.random((n, 3)).astype(np.float32) - 0.5) * 2 - o
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        f[:, 0:3], f[:, 3:6], f[:, 6:9], f[:, 9:12] = o, d, lo, hi
    elif fn in (9, 10, 11):
        f[:, 0] = rng.integers(0, 2, n)
        f[:, 1] = rng.random(n) * 0.95 + 0.05
        f[:, 2:5] = rng.random((n, 3))
        nrm = unit_dirs(rng, n)
        wo = unit_dirs(rng, n)
        flip = (wo * nrm).sum(axis=1) > 0          # wo points INTO the surface (rt.rgen:134-137)
        wo[flip] = -wo[flip]
        wi = unit_dirs(rng, n)
        up = (wi * nrm).sum(axis=1) < 0
        wi[up & (rng.random(n) < 0.8)] *= -1       # mostly the upper hemisphere
        f[:, 5:8], f[:, 8:11], f[:, 11:14] = wo, nrm, wi
    elif fn == 12:
        f[:, 0:9] = rng.random((n, 9))
        f[:, 9], f[:, 10], f[:, 11] = rng.random(n) * 5, rng.random(n), rng.random(n) * 3
        f[:, 12:21] = rng.random((n, 9))
        u[:, 1] = rng.integers(0, 100, n)
    elif fn == 13:
        f[:, 0:2] = rng.random((n, 2)) * 10 + 1e-3
    elif fn == 14:
        f[:, 0:3] = rng.random((n, 3)) * 4
    elif fn == 15:
        u[:, 2] = rng.integers(1, 4096, n)
        u[:, 1] = rng.integers(0, 4096, n) % u[:, 2]
    return f, u


def run(lib_fn, fn, f, u):
    out = np.zeros((len(f), N_OUT), np.float32)
    u = u.copy()
    for i in range(len(f)):
        lib_fn(fn, f[i].ctypes.data_as(C.c_void_p), u[i].ctypes.data_as(C.c_void_p), out[i].ctypes.data_as(C.c_void_p))
    return out, u


# ---- whole frames ------------------------------------------------------------------------------------------------
class RefSceneArgs(C.Structure):
    _fields_ = [("n_objs", C.c_uint32), ("descs", C.c_void_p), ("tri_off", C.c_void_p), ("vert_off", C.c_void_p),
                ("verts", C.c_void_p), ("idx", C.c_void_p), ("n_lights", C.c_uint32), ("lights", C.c_void_p),
                ("n_tex", C.c_uint32), ("bvh", C.c_void_p), ("scene", C.c_void_p), ("trace", C.c_void_p),
                ("occluded", C.c_void_p), ("texture", C.c_void_p)]


def ref_renderer(orc):
    """render_frame(rs, st, consts, cam, seed, n_tex) through the compiled reference shader; the acceleration structure
    and the texture unit (not shader text) are the oracle's, passed as callbacks"""
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libglsl_ref.so"))

    def fptr(name):
        return C.cast(getattr(orc.lib, name), C.c_void_p).value

    def render(rs, st, consts, cam, seed, n_tex):
        a = RefSceneArgs(rs.n_objs, rs.descs.ctypes.data, rs.tri_off.ctypes.data, rs.vert_off.ctypes.data, rs.verts.ctypes.data,
                         rs.idx.ctypes.data, rs.n_lights, rs.lights.ctypes.data, n_tex, rs.bvh.h, rs.h,
                         fptr("orc_bvh_trace_one"), fptr("orc_bvh_occluded_one"), fptr("orc_texture_fetch"))
        cur, prev = st.parity, st.parity ^ 1
        counts = np.zeros(2, np.uint64)
        p = lambda x: C.c_void_p(x.ctypes.data)  # noqa: E731
        frame = int(np.asarray(consts, np.uint32)[8])
        ref.ref_glsl_render_frame(C.byref(a), p(consts), p(cam), st.w, st.h, C.c_uint32(seed ^ frame), p(st.image),
                                  p(st.res[prev]), p(st.res[cur]), p(st.gb[prev][0]), p(st.gb[prev][1]), p(st.gb[prev][2]),
                                  p(st.gb[cur][0]), p(st.gb[cur][1]), p(st.gb[cur][2]), p(counts))
        st.parity ^= 1
        return counts
    return render


def frame_cases(gpurt):
    """(name, scene, textures, w, h, frames, camera, tunables) — the cases of tests/test_emu_render.py"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_emu_render as T
    media = os.path.join(ROOT, "tests", "data", "media")
    cbox = gpurt.Scene(None).load(os.path.join(media, "cbox", "cbox.gltf"))
    for integ in range(5):
        for brdf in (0, 1):
            yield (f"cbox_i{integ}_b{brdf}", cbox, (), 64, 36, 3, None,
                   dict(integrator=integ, brdf=brdf, samples_per_frame=2, max_depth=4, seed=1234 + integ))
    yield ("cbox_qmc", cbox, (), 48, 27, 2, None, dict(integrator=1, brdf=1, use_qmc=1, use_metalness=1, use_rr=0, max_depth=3,
                                                      samples_per_frame=2, env_scale=1.0, seed=3))
    for dv in (1, 2, 3):
        yield (f"cbox_debug{dv}", cbox, (), 48, 27, 2, None, dict(integrator=0, debug_view=dv, samples_per_frame=1, seed=5))
    mis = gpurt.Scene(None).load(os.path.join(media, "mis_test", "mis_test.gltf"))
    cam = gpurt.camera(1, 80, 45, (0.5, 0.6, 2.6), (0.5, 0.45, 0.0), 50.0)
    yield ("mis_test_mis", mis, (), 80, 45, 2, cam, dict(integrator=2, brdf=1, samples_per_frame=1, max_depth=4, seed=7))
    yield ("mis_test_restir", mis, (), 80, 45, 4, cam, dict(integrator=3, brdf=0, samples_per_frame=1, max_depth=4, res_samples=4,
                                                           use_temporal=1, temporal_scale=16, seed=8))
    feat = gpurt.Scene(None).load(os.path.join(ROOT, "tests", "data", "synth", "features.gltf"))
    texs = [feat.texture(i) for i in range(feat.counts()["textures"])]
    cam = gpurt.camera(1, 80, 60, (4.0, 3.0, 6.0), (1.0, 1.0, 2.0), 60.0)
    for integ in (0, 1, 2, 4):
        yield (f"features_i{integ}", feat, texs, 80, 60, 2, cam,
               dict(integrator=integ, brdf=integ % 2, samples_per_frame=2, max_depth=3, use_normal_map=1, use_metalness=1,
                    env_scale=0.5, seed=77 + integ))
    quads, qtex = T._textured_quads_scene(gpurt)   # emissive-TEXTURED light: light_sample's texture path, quirk Q5
    cam = gpurt.camera(1, 64, 48, (1.5, 1.0, 2.5), (0.0, 0.0, 0.5), 70.0)
    for integ in (0, 2, 3, 4):
        yield (f"texquads_i{integ}", quads, qtex, 64, 48, 2, cam,
               dict(integrator=integ, brdf=1, samples_per_frame=2, max_depth=3, use_normal_map=1, use_metalness=1, seed=20 + integ))
    for seed in (1, 2):
        rnd = T._random_material_scene(gpurt, seed)
        cam = gpurt.camera(1, 48, 27, (0.2, 0.1, 2.4), (0.0, 0.0, 0.0), 70.0)
        for integ in range(5):
            yield (f"random{seed}_i{integ}", rnd, (), 48, 27, 2, cam,
                   dict(integrator=integ, brdf=(integ + seed) % 2, samples_per_frame=2, max_depth=4, env_scale=0.3,
                        use_metalness=seed % 2, seed=100 * seed + integ))


def buffer_digest(arr):
    """SHA-256 of a frame buffer's words with every NaN replaced by the canonical quiet NaN 0x7fc00000: which NaN an
    operation returns (sign, payload) is left open by GLSL and differs between x86 (first operand's payload) and the GPU
    (0x7fffffff); everything that is a number — or an integer word such as n_seen — is compared bit for bit"""
    import hashlib
    w = np.ascontiguousarray(arr).view(np.uint32).copy()
    w[(w & 0x7FFFFFFF) > 0x7F800000] = 0x7FC00000
    return hashlib.sha256(w.tobytes()).hexdigest()


def frame_digests(gpurt, orc, render):
    """{case/frame/buffer: sha256} of the output buffers of `render` (the compiled reference or the oracle).  Every case
    is rendered as frames 0, 1, 2, ... (keys case/frameF/...); the ReSTIR cases (integrators 3 / 4, whose frames depend on
    the previous one) additionally as the sequence a freshly created RTPipe produces — frame 0 twice, both with prev_PV =
    identity, because `old_cam` is only assigned during the second call (rt.cpp:121-138) — under case/pipeK/... for call K"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_emu_render as T
    out = {}

    def sequence(name, tag, rs, w, h, cam, texs, kw, plan):
        st = orc.FrameState(w, h)
        for k, (f, ident) in enumerate(plan):
            consts, ubo, seed = T._uniforms(gpurt, rs, cam, f, prev_identity=ident, **kw)
            counts = render(rs, st, consts, ubo, seed, len(texs))
            cur = st.parity ^ 1
            bufs = {"image": st.image, "pos": st.gb[cur][0], "norm": st.gb[cur][1], "albedo": st.gb[cur][2]}
            if kw.get("integrator", 0) in (3, 4):
                bufs["reservoirs"] = st.res[cur]
            for b, arr in bufs.items():
                out[f"{name}/{tag}{k}/{b}"] = buffer_digest(arr)
            out[f"{name}/{tag}{k}/rays"] = f"{int(counts[0])},{int(counts[1])}"

    for name, scene, texs, w, h, frames, cam, kw in frame_cases(gpurt):
        rs = orc.RenderScene(scene, texs)
        rs.gscene, rs.textures = scene, texs   # for renderers that build their own structures from the scene
        cam = cam or gpurt.camera(0, w, h)
        sequence(name, "frame", rs, w, h, cam, texs, kw, [(f, False) for f in range(frames)])
        if kw.get("integrator", 0) in (3, 4):
            sequence(name, "pipe", rs, w, h, cam, texs, kw, [(0, True), (0, True)] + [(f, False) for f in range(1, frames)])
    return out


def tonemap_inputs():
    rng = np.random.default_rng(3)
    x = (rng.random((100000, 4), dtype=np.float32) * 4).astype(np.float32)
    x[:100, 0], x[100:200, 1], x[200:300, 2] = np.nan, np.inf, -1
    return x


def tonemap_digests(fn):
    """tonemap.frag + the R8G8B8A8_SRGB framebuffer store: RGBA8 digests for the three operators"""
    import hashlib
    x, out = tonemap_inputs(), {}
    for op in (0, 1, 2):
        for exposure, gamma in ((1.0, 2.2), (0.37, 1.0), (3.0, 2.6)):
            out[f"tonemap/op{op}/e{exposure}/g{gamma}"] = hashlib.sha256(np.ascontiguousarray(fn(x, op, exposure, gamma)).tobytes()).hexdigest()
    return out


def _ref_tonemap(x, op, exposure, gamma):
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libglsl_ref.so"))
    out = np.zeros(x.shape, np.uint8)
    ref.ref_glsl_tonemap(C.c_void_p(x.ctypes.data), C.c_ulonglong(len(x)), op, C.c_float(exposure), C.c_float(gamma),
                         C.c_void_p(out.ctypes.data))
    return out


def main():
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_ref/libglsl_ref.so"])
    ref = C.CDLL(os.path.join(ref_dir, "libglsl_ref.so")).ref_glsl_unit
    ref.restype = None
    rng = np.random.default_rng(20261017)
    data = {}
    for fn in NAMES:
        f, u = make_inputs(fn, 400, rng)
        out, uo = run(ref, fn, f, u)
        data[f"in_{fn}"], data[f"u_{fn}"], data[f"out_{fn}"], data[f"uout_{fn}"] = f, u, out, uo
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "glsl_unit_golden.npz"), **data)
    print("wrote", len(NAMES), "functions x 400 vectors")
    sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "gpu-rt_b200")]
    import json
    import gpurt
    import orc
    digests = frame_digests(gpurt, orc, ref_renderer(orc))
    json.dump(digests, open(os.path.join(ROOT, "tests", "golden", "glsl_frames_golden.json"), "w"), indent=0, sort_keys=True)
    for k, v in tonemap_digests(lambda x, op, e, g: _ref_tonemap(x, op, e, g)).items():
        digests[k] = v
    json.dump(digests, open(os.path.join(ROOT, "tests", "golden", "glsl_frames_golden.json"), "w"), indent=0, sort_keys=True)
    print("wrote", len(digests), "frame-buffer digests")


if __name__ == "__main__":
    main()

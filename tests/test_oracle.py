"""CPU tests pinning the oracle itself (no GPU).  The reference ships no golden vectors for this
path, so the oracle is pinned by (a) independent pure-Python restatements of the GLSL
bit-twiddling, (b) brute force vs its own BVH, (c) size-independent geometric properties."""
import ctypes as C

import numpy as np
import pytest

from scenes import soup


def _tea_py(v0, v1):  # rtcommon.glsl:99-109
    M = 0xFFFFFFFF
    s0 = 0
    for _ in range(16):
        s0 = (s0 + 0x9E3779B9) & M
        v0 = (v0 + ((((v1 << 4) & M) + 0xA341316C) & M ^ ((v1 + s0) & M) ^ (((v1 >> 5) + 0xC8013EA4) & M))) & M
        v1 = (v1 + ((((v0 << 4) & M) + 0xAD90777D) & M ^ ((v0 + s0) & M) ^ (((v0 >> 5) + 0x7E95761E) & M))) & M
    return v0


def test_rng_matches_python_restatement(orc):
    for a, b in [(0, 0), (1, 2), (123456, 0xC0FFEE), (0xFFFFFFFF, 0xFACADE), (2073599, 77)]:
        assert orc.lib.orc_tea(a, b) == _tea_py(a, b)
    s = C.c_uint32(12345)
    py = 12345
    for _ in range(100):
        py = (1664525 * py + 1013904223) & 0xFFFFFFFF
        assert orc.lib.orc_lcg(C.byref(s)) == py & 0x00FFFFFF and s.value == py
    s = C.c_uint32(99)
    f = orc.lib.orc_randf(C.byref(s))
    assert 0 <= f < 1 and f == np.float32(s.value & 0xFFFFFF) / np.float32(0x1000000)
    for i in [0, 1, 2, 3, 5, 1 << 31, 0xDEADBEEF]:
        assert orc.lib.orc_radical_inverse(i) == np.float32(int(f"{i:032b}"[::-1], 2)) * np.float32(2.3283064365386963e-10)


def test_intersection_contract(orc):
    tri = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0], np.float32)
    t, u, v = C.c_float(), C.c_float(), C.c_float()
    def hit(o, d, tmin=1e-5, tmax=1e7):
        return orc.lib.orc_intersect(np.array(list(o) + [tmin] + list(d) + [tmax], np.float32), tri, C.byref(t), C.byref(u), C.byref(v))
    assert hit((0.25, 0.25, 1), (0, 0, -1)) and t.value == 1 and u.value == 0.25 and v.value == 0.25
    assert hit((0.25, 0.25, -1), (0, 0, 1)), "no back-face culling (vulkan.cpp:802)"
    assert not hit((0.25, 0.25, 1), (0, 0, 1)), "behind the origin"
    assert not hit((0.25, 0.25, 1), (1, 0, 0)), "parallel ray: det == 0"
    assert hit((0, 0, 1), (0, 0, -1)) and hit((1, 0, 1), (0, 0, -1)) and hit((0.5, 0.5, 1), (0, 0, -1)), "edges/vertices inclusive"
    assert not hit((0.6, 0.6, 1), (0, 0, -1))
    assert not hit((0.25, 0.25, 1), (0, 0, -1), tmin=1.0) and not hit((0.25, 0.25, 1), (0, 0, -1), tmax=1.0), "interval is exclusive"
    assert hit((0.25, 0.25, 1), (0, 0, -1), tmin=0.999, tmax=1.001)


def test_closest_point_regions(orc):
    tri = np.array([0, 0, 0, 2, 0, 0, 0, 2, 0], np.float32)
    c = np.zeros(3, np.float32)
    v, w = C.c_float(), C.c_float()
    cases = {(-1, -1, 0): (0, 0, 0), (3, -1, 0): (2, 0, 0), (-1, 3, 0): (0, 2, 0), (1, -1, 0): (1, 0, 0),
             (-1, 1, 0): (0, 1, 0), (2, 2, 0): (1, 1, 0), (0.5, 0.5, 3): (0.5, 0.5, 0)}
    for p, want in cases.items():
        d2 = orc.lib.orc_closest_point_tri(np.array(p, np.float32), tri, c, C.byref(v), C.byref(w))
        assert np.allclose(c, want, atol=1e-6), (p, c)
        assert np.isclose(d2, np.sum((np.array(p) - np.array(want)) ** 2), rtol=1e-6)


def test_bvh_equals_brute_force(orc):
    for n in (1, 2, 7, 300, 4000):
        tris = soup(n, seed=n)
        if n >= 300:
            tris[5:9] = tris[5]
        b = orc.Bvh(tris)
        rays = orc.gen_random_rays(4000, 0xC0FFEE, b.scene_box())
        assert (b.closest_hit(rays).view(np.uint32) == orc.closest_hit_brute(tris, rays).view(np.uint32)).all()
        assert (b.any_hit(rays) == orc.any_hit_brute(tris, rays)).all()
        for r2 in (np.inf, 0.01):
            q = orc.gen_random_points(4000, 0xFACADE, b.scene_box(), r2=r2)
            assert (b.closest_point(q).view(np.uint32) == orc.closest_point_brute(tris, q).view(np.uint32)).all()
        k = b.keys()
        assert (k[1:] >= k[:-1]).all() and sorted(b.prim_order()) == list(range(n))
        o = b.prim_order()
        same = k[1:] == k[:-1]
        assert (o[1:][same] > o[:-1][same]).all(), "equal keys keep gid order (stable sort)"


def test_karras_tree_is_a_valid_binary_tree(orc):
    tris = soup(1000, seed=3)
    b = orc.Bvh(tris)
    l, r, boxes = b.bvh2()
    seen_leaf, seen_int = np.zeros(1000, int), np.zeros(999, int)
    for c in np.concatenate([l, r]):
        if c < 0:
            seen_leaf[~c] += 1
        else:
            seen_int[c] += 1
    assert (seen_leaf == 1).all() and (seen_int[1:] == 1).all() and seen_int[0] == 0
    order = b.prim_order()
    t = tris.reshape(-1, 3, 3)
    for i in (0, 17, 500, 998):     # node box == union of its subtree
        stack, lo, hi = [i], np.full(3, np.inf), np.full(3, -np.inf)
        while stack:
            c = stack.pop()
            if c < 0:
                lo, hi = np.minimum(lo, t[order[~c]].min(0)), np.maximum(hi, t[order[~c]].max(0))
            else:
                stack += [l[c], r[c]]
        assert (boxes[i, :3] == lo.astype(np.float32)).all() and (boxes[i, 3:] == hi.astype(np.float32)).all()


def test_detmath_accuracy():
    """include/gpurt_detmath.h against numpy in double (bounds far inside Vulkan's)"""
    import subprocess, tempfile, os, textwrap
    from conftest import ROOT
    src = textwrap.dedent("""
        #include <stdio.h>
        #include "%s/include/gpurt_detmath.h"
        int main(){ for(int i=0;i<2000;i++){ float x=i*0.00314159f, b=(i+1)/2001.0f, y=(i%%97)*0.37f;
          printf("%%.9g %%.9g %%.9g %%.9g %%.9g %%.9g\\n", x, dm_sin(x), dm_cos(x), b, y, dm_pow(b,y)); } return 0; }
    """ % ROOT)
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", os.path.join(d, "t.c"), "-o", os.path.join(d, "t"), "-lm"])
        out = np.array([[float(v) for v in line.split()] for line in subprocess.check_output([os.path.join(d, "t")]).decode().splitlines()])
    assert np.abs(out[:, 1] - np.sin(out[:, 0])).max() < 3e-7 and np.abs(out[:, 2] - np.cos(out[:, 0])).max() < 3e-7
    ref = np.power(out[:, 3], out[:, 4])
    big = ref > 1e-30          # below that dm_pow flushes towards zero like GPU exp2 does
    assert (np.abs(out[big, 5] - ref[big]) / ref[big]).max() < 5e-5 and (out[~big, 5] <= 1e-30).all()


def test_oracle_integrator_is_deterministic_and_accumulates(orc, gpurt):
    import os
    from conftest import MEDIA
    s = gpurt.Scene(None).load(os.path.join(MEDIA, "cbox", "cbox.gltf"))
    rs = orc.RenderScene(s)
    w, h = 48, 27
    cam = np.frombuffer(bytes(gpurt.camera(0, w, h)), np.uint32).copy()

    def consts(frame, integ, samples=2):
        c = np.zeros(22, np.uint32)
        c.view(np.float32)[0:8] = [0.3, 0.3, 0.3, 1, 0, 0, 0, 1]
        c[8:22] = np.array([frame, samples, 256, 0, 4, 0, 0, 1, integ, 0, 0, 1, rs.n_lights, rs.n_objs], np.int32).view(np.uint32)
        return c
    for integ in (0, 2, 4):
        a, b = orc.FrameState(w, h), orc.FrameState(w, h)
        orc.render_frame(rs, a, consts(0, integ), cam, 5)
        orc.render_frame(rs, b, consts(0, integ), cam, 5, threads=1)
        assert (a.image == b.image).all(), "thread count must not change the result"
        f0 = a.image.copy()
        orc.render_frame(rs, a, consts(1, integ), cam, 5)
        c = orc.FrameState(w, h)
        c.parity = 1
        c.res, c.gb = a.res, a.gb       # frame 1 alone, same temporal inputs
        single = orc.FrameState(w, h)
        single.res[0][:], single.gb[0][0][:], single.gb[0][1][:], single.gb[0][2][:] = b.res[0], b.gb[0][0], b.gb[0][1], b.gb[0][2]
        single.parity = 1
        orc.render_frame(rs, single, consts(1, integ), cam, 5)
        # rt.rgen:638-645: image_1 = mix(image_0, avg_1, 1/2); `single` started from a zero image
        avg1 = (single.image[..., :3] - 0.0 * 0.5) / 0.5
        assert np.allclose(a.image[..., :3], f0[..., :3] * 0.5 + avg1 * 0.5, rtol=1e-5, atol=1e-6)
        d = orc.FrameState(w, h)
        orc.render_frame(rs, d, consts(0, integ), cam, 6)
        assert (d.image != f0).any(), "seed must matter"


# ---- the oracle against the reference's own shader text ---------------------------------------------------------
# oracle/_ref/libglsl_ref.so is the reference's rtcommon.glsl + restir.glsl + rt.rgen compiled as C++ (oracle/
# make_glsl_ref.py rewrites the text into oracle/_ref/, oracle/ref_shim/glsl_compat.h supplies the GLSL types and — for
# what GLSL leaves to the implementation: fma contraction, sin / cos / pow — the numeric contract N8 of DESIGN.md §3).
# tests/golden/make_glsl_golden.py ran it in the build container and committed (a) input / output vectors of 16 shader
# functions and (b) SHA-256 digests of whole frames.  The oracle's restatement must reproduce both BIT FOR BIT.
GLSL_FUNCS = {0: "tea", 1: "randf", 2: "randu", 3: "cospow_hemisphere", 4: "triangle_sample", 5: "triangle_hit",
              6: "triangle_pdf", 7: "make_tanspace", 8: "hit_bbox", 9: "MAT_pdf", 10: "MAT_eval", 11: "MAT_sample",
              12: "res_update", 13: "power_heuristic", 14: "luma", 15: "hammersley"}


def _glsl_unit(fn_ptr, fn, f, u):
    out = np.zeros((len(f), 12), np.float32)
    u = u.copy()
    for i in range(len(f)):
        fn_ptr(fn, f[i].ctypes.data_as(C.c_void_p), u[i].ctypes.data_as(C.c_void_p), out[i].ctypes.data_as(C.c_void_p))
    return out, u


@pytest.mark.parametrize("fn", sorted(GLSL_FUNCS))
def test_oracle_functions_match_the_reference_shader_text(orc, fn):
    import os
    from conftest import ROOT
    g = np.load(os.path.join(ROOT, "tests", "golden", "glsl_unit_golden.npz"))
    f, u, want, uwant = g[f"in_{fn}"], g[f"u_{fn}"], g[f"out_{fn}"], g[f"uout_{fn}"]
    orc.lib.orc_glsl_unit.restype = None
    got, ugot = _glsl_unit(orc.lib.orc_glsl_unit, fn, f, u)
    name = GLSL_FUNCS[fn]
    assert (ugot == uwant).all(), f"{name}: RNG state / unsigned results differ"
    assert (got.view(np.uint32) == want.view(np.uint32)).all(), f"{name}: {(got.view(np.uint32) != want.view(np.uint32)).any(axis=1).sum()} of {len(f)} results differ"
    assert np.abs(want).sum() > 0 or fn in (0, 2)
    so = os.path.join(ROOT, "oracle", "_ref", "libglsl_ref.so")
    if os.path.exists(so):   # the compiled reference itself is here: the committed vectors are not stale
        ref = C.CDLL(so).ref_glsl_unit
        ref.restype = None
        live, ulive = _glsl_unit(ref, fn, f, u)
        assert (live.view(np.uint32) == want.view(np.uint32)).all() and (ulive == uwant).all(), "golden file is stale"


def test_oracle_frames_match_the_reference_shader_text(orc, gpurt):
    """whole frames of rt.rgen `main` (all integrators, both BRDFs, textures, ReSTIR temporal reuse over several frames,
    QMC, debug views, scenes with NaN normals and zero-area lights): image, G-buffers, reservoirs and ray counts of the
    oracle == those of the reference's shader text, by SHA-256 of the raw buffers"""
    import json
    import os
    from conftest import ROOT
    sys_path_golden = os.path.join(ROOT, "tests", "golden")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_glsl_golden", os.path.join(sys_path_golden, "make_glsl_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    want = json.load(open(os.path.join(sys_path_golden, "glsl_frames_golden.json")))
    got = mg.frame_digests(gpurt, orc, lambda rs, st, consts, cam, seed, n_tex: orc.render_frame(rs, st, consts, cam, seed))
    got.update(mg.tonemap_digests(lambda x, op, e, g: orc.tonemap(x, op, e, g)))   # tonemap.frag + framebuffer store
    assert set(got) == set(want)
    bad = [k for k in want if got[k] != want[k]]
    assert not bad, f"{len(bad)} of {len(want)} frame buffers differ from the reference shader's: {bad[:5]}"

"""CPU replay of the product's integrator code (csrc/shade.cuh: the per-pixel functions render.cu's kernels call)
against the oracle's restatement of rt.rgen — bit for bit, without a GPU.

The replay (tests/emu/libemu.so) compiles shade.cuh / traverse.cuh / bvh8.cuh for the host and runs, per pixel,
k_frame_begin -> k_gen_camera -> the bounce loop -> k_frame_end over the emu's own wide BVH.  It is test-only:
the product never links it, and the GPU parity tests (test_gpu_render.py) remain the proof for the CUDA build."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import MEDIA, ROOT


class FrameArgs(C.Structure):
    _fields_ = [("descs", C.c_void_p), ("tri_off", C.c_void_p), ("vert_off", C.c_void_p), ("verts", C.c_void_p),
                ("idx", C.c_void_p), ("lights", C.c_void_p), ("tex_info", C.c_void_p), ("texels", C.c_void_p),
                ("n_objs", C.c_uint32), ("n_lights", C.c_uint32), ("n_tex", C.c_uint32),
                ("light_bvh", C.c_void_p), ("ltri_off", C.c_void_p)]


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def emu(built):
    lib = C.CDLL(os.path.join(ROOT, "tests", "emu", "libemu.so"))
    lib.emu_build.restype = C.c_void_p
    lib.emu_render_frame.restype = None
    return lib


class EmuScene:
    """the arrays of orc.RenderScene + the emu's wide BVH over the same world triangles"""

    def __init__(self, emu, orc, gscene, textures=()):
        self.emu, self.rs = emu, orc.RenderScene(gscene, textures)
        rs, b = self.rs, self.rs.bvh
        self.order = b.prim_order()
        l, r, bx = b.bvh2()
        inflate = np.float32(max(1e-30, np.abs(b.scene_box()).max()) * 2.0 ** -19)
        self.h = C.c_void_p(emu.emu_build(_vp(rs.tris), len(rs.tris), _vp(self.order), _vp(l), _vp(r), _vp(bx),
                                          C.c_float(inflate)))
        # light BVH: the lights' triangles in light order, built like any other scene (render.cu pipe_light_accel);
        # its boxes are padded for ray origins anywhere in the MAIN scene: 128 x the main inflation
        self.lh, self.ltri_off = None, np.zeros(rs.n_lights + 1, np.uint32)
        lights = rs.lights.reshape(-1, 12)[:rs.n_lights]
        if rs.n_lights and all(int(l[9]) > 0 for l in lights):
            parts = [rs.tris[int(rs.tri_off[int(l[8])]):int(rs.tri_off[int(l[8])]) + int(l[9])] for l in lights]
            self.ltris = np.ascontiguousarray(np.concatenate(parts), np.float32)
            self.ltri_off[1:] = np.cumsum([len(q) for q in parts])
            lb = orc.Bvh(self.ltris)
            ll, lr, lbx = lb.bvh2()
            self.lorder = lb.prim_order()
            self.lh = C.c_void_p(emu.emu_build(_vp(self.ltris), len(self.ltris), _vp(self.lorder), _vp(ll), _vp(lr), _vp(lbx),
                                               C.c_float(np.float32(inflate) * np.float32(128.0))))
            self.lb = lb
        self.args = FrameArgs(rs.descs.ctypes.data, rs.tri_off.ctypes.data, rs.vert_off.ctypes.data, rs.verts.ctypes.data,
                              rs.idx.ctypes.data, rs.lights.ctypes.data, rs.tex_info.ctypes.data, rs.texels.ctypes.data,
                              rs.n_objs, rs.n_lights, len(textures), self.lh, self.ltri_off.ctypes.data)

    def render_frame(self, st, consts, cam, seed_val):
        cur, prev = st.parity, st.parity ^ 1
        counts = np.zeros(2, np.uint64)
        self.emu.emu_render_frame(self.h, C.byref(self.args), _vp(consts), _vp(cam), st.w, st.h, C.c_uint32(seed_val),
                                  _vp(st.image), _vp(st.res[prev]), _vp(st.res[cur]), _vp(st.gb[prev][0]),
                                  _vp(st.gb[prev][1]), _vp(st.gb[prev][2]), _vp(st.gb[cur][0]), _vp(st.gb[cur][1]),
                                  _vp(st.gb[cur][2]), _vp(counts), 0)
        st.parity ^= 1
        return counts

    def close(self):
        self.emu.emu_free(self.h)
        if self.lh:
            self.emu.emu_free(self.lh)


def _uniforms(gpurt, rs, cam, frame, prev_identity=False, **kw):
    """the words RTPipe::trace / update_uniforms would push (rt.cpp:121-138, :355-368) for these tunables"""
    p = gpurt.pipe_params(**kw)
    c = gpurt.Constants()
    c.clear_col = (C.c_float * 4)(p.clear[0], p.clear[1], p.clear[2], 1.0)
    e = [np.float32(p.env_scale) * np.float32(p.env[k]) for k in range(3)]
    c.env_light = (C.c_float * 4)(e[0], e[1], e[2], 1.0)
    c.frame, c.samples, c.max_frame, c.qmc, c.max_depth = frame, p.samples_per_frame, p.max_frames, p.use_qmc, p.max_depth
    c.use_normal_map, c.use_metalness, c.use_temporal, c.integrator = p.use_normal_map, p.use_metalness, p.use_temporal, p.integrator
    c.brdf, c.debug_view, c.use_rr, c.n_lights, c.n_objs = p.brdf, p.debug_view, p.use_rr, rs.n_lights, rs.n_objs
    cam = type(cam).from_buffer_copy(bytes(cam))
    cam.new_samples, cam.temporal_multiplier = p.res_samples, p.temporal_scale
    V = np.array(cam.V, np.float32).reshape(4, 4).T
    P = np.array(cam.P, np.float32).reshape(4, 4).T
    cam.prev_PV = (C.c_float * 16)(*(P @ V).T.reshape(-1))   # static camera: prev_PV = P * V
    if prev_identity:                                        # the first two calls of a new RTPipe (old_cam = {})
        cam.prev_PV = (C.c_float * 16)(*np.eye(4, dtype=np.float32).reshape(-1))
    return np.frombuffer(bytes(c), np.uint32).copy(), np.frombuffer(bytes(cam), np.uint32).copy(), p.seed


def _run(emu, orc, gpurt, gscene, w, h, frames, cam=None, textures=(), **kw):
    es = EmuScene(emu, orc, gscene, textures)
    a, b = orc.FrameState(w, h), orc.FrameState(w, h)
    cam = cam or gpurt.camera(0, w, h)
    for f in range(frames):
        consts, ubo, seed = _uniforms(gpurt, es.rs, cam, f, **kw)
        ce = es.render_frame(a, consts, ubo, seed ^ f)
        co = orc.render_frame(es.rs, b, consts, ubo, seed)
        assert (a.image.view(np.uint32) == b.image.view(np.uint32)).all(), f"frame {f}: image differs"
        cur = a.parity ^ 1
        for g in range(3):
            assert (a.gb[cur][g].view(np.uint32) == b.gb[cur][g].view(np.uint32)).all(), f"frame {f}: G-buffer {g} differs"
        if kw.get("integrator", 0) in (3, 4):
            assert (a.res[cur] == b.res[cur]).all(), f"frame {f}: reservoirs differ"
        assert tuple(int(x) for x in ce) == tuple(int(x) for x in co), "ray counts differ"
    img = a.image.copy()
    es.close()
    return img


@pytest.mark.parametrize("integrator", [0, 1, 2, 3, 4])
def test_cbox_all_integrators(emu, orc, gpurt, integrator):
    s = gpurt.Scene(None).load(os.path.join(MEDIA, "cbox", "cbox.gltf"))
    for brdf in (0, 1):
        img = _run(emu, orc, gpurt, s, 64, 36, 3, integrator=integrator, brdf=brdf, samples_per_frame=2, max_depth=4,
                   seed=1234 + integrator)
        assert img[..., :3].mean() > 0.01


def test_restir_spatial_reuse_extension(emu, orc, gpurt):
    """GpurtPipeParams::spatial_samples (no reference counterpart, off by default): the product's code and the oracle's
    restatement of the extension agree bit for bit over frames with temporal + spatial reuse, and 0 samples is the
    reference's estimator"""
    s = gpurt.Scene(None).load(os.path.join(MEDIA, "cbox", "cbox.gltf"))
    es = EmuScene(emu, orc, s, ())
    emu.emu_set_spatial.argtypes = [C.c_uint32, C.c_float]
    images = {}
    for spatial in ((0, 16.0), (3, 6.0)):
        a, b = orc.FrameState(64, 36), orc.FrameState(64, 36)
        cam = gpurt.camera(0, 64, 36)
        for integ in (3, 4):
            for f in range(4):
                consts, ubo, seed = _uniforms(gpurt, es.rs, cam, f, integrator=integ, brdf=1, samples_per_frame=2, max_depth=3,
                                              res_samples=4, use_temporal=1, temporal_scale=16, seed=90 + integ)
                emu.emu_set_spatial(*spatial)
                ce = es.render_frame(a, consts, ubo, seed ^ f)
                co = orc.render_frame(es.rs, b, consts, ubo, seed, spatial=spatial)
                assert (a.image.view(np.uint32) == b.image.view(np.uint32)).all(), f"spatial {spatial} integrator {integ} frame {f}: image"
                assert (a.res[a.parity ^ 1] == b.res[b.parity ^ 1]).all()
                assert tuple(int(x) for x in ce) == tuple(int(x) for x in co)
        images[spatial[0]] = a.image.copy()
    emu.emu_set_spatial(0, 16.0)
    orc.lib.orc_render_set_spatial(0, 16.0)
    assert not (images[0].view(np.uint32) == images[3].view(np.uint32)).all()    # the extension does something
    es.close()


def test_light_sampling_extension(emu, orc, gpurt):
    """GpurtPipeParams::light_sampling = 1 (no reference counterpart, off by default): the product's code and the oracle's
    restatement agree bit for bit for every integrator that samples or evaluates lights (direct, MIS with its weighted
    light_pdf through the light BVH and through the scan, both ReSTIR variants), and the flag changes the frames"""
    s = gpurt.Scene(None).load(os.path.join(MEDIA, "mis_test", "mis_test.gltf"))
    es = EmuScene(emu, orc, s, ())
    emu.emu_set_light_sampling.argtypes = [C.c_uint32]
    cam = gpurt.camera(1, 64, 36, (0.5, 0.6, 2.6), (0.5, 0.45, 0.0), 50.0)
    try:
        for integ in (0, 2, 3, 4):
            frames = {}
            for mode in (0, 1):
                for bvh in ((1, 0) if integ == 2 else (1,)):
                    emu.emu_set_light_bvh(bvh)
                    a, b = orc.FrameState(64, 36), orc.FrameState(64, 36)
                    for f in range(3):
                        consts, ubo, seed = _uniforms(gpurt, es.rs, cam, f, integrator=integ, brdf=1, samples_per_frame=2, max_depth=3,
                                                      res_samples=4, use_temporal=1, temporal_scale=16, seed=300 + integ)
                        emu.emu_set_light_sampling(mode)
                        ce = es.render_frame(a, consts, ubo, seed ^ f)
                        co = orc.render_frame(es.rs, b, consts, ubo, seed, light_sampling=mode)
                        assert (a.image.view(np.uint32) == b.image.view(np.uint32)).all(), f"light_sampling {mode} integrator {integ} bvh {bvh} frame {f}"
                        assert (a.res[a.parity ^ 1] == b.res[b.parity ^ 1]).all()
                        assert tuple(int(x) for x in ce) == tuple(int(x) for x in co)
                    frames[mode] = a.image.copy()
            assert not (frames[0].view(np.uint32) == frames[1].view(np.uint32)).all()
    finally:
        emu.emu_set_light_sampling(0)
        emu.emu_set_light_bvh(1)
        orc.lib.orc_render_set_light_sampling(0)
    es.close()


def test_mis_test_scene(emu, orc, gpurt):
    s = gpurt.Scene(None).load(os.path.join(MEDIA, "mis_test", "mis_test.gltf"))
    cam = gpurt.camera(1, 80, 45, (0.5, 0.6, 2.6), (0.5, 0.45, 0.0), 50.0)
    _run(emu, orc, gpurt, s, 80, 45, 2, cam=cam, integrator=2, brdf=1, samples_per_frame=1, max_depth=4, seed=7)
    _run(emu, orc, gpurt, s, 80, 45, 4, cam=cam, integrator=3, brdf=0, samples_per_frame=1, max_depth=4, res_samples=4,
         use_temporal=1, temporal_scale=16, seed=8)


def test_options(emu, orc, gpurt):
    s = gpurt.Scene(None).load(os.path.join(MEDIA, "cbox", "cbox.gltf"))
    _run(emu, orc, gpurt, s, 48, 27, 2, integrator=1, brdf=1, use_qmc=1, use_metalness=1, use_rr=0, max_depth=3,
         samples_per_frame=2, env_scale=1.0, seed=3)
    _run(emu, orc, gpurt, s, 48, 27, 2, integrator=2, brdf=0, max_depth=1, samples_per_frame=1, seed=4)
    for dv in (1, 2, 3):
        _run(emu, orc, gpurt, s, 48, 27, 2, integrator=0, debug_view=dv, samples_per_frame=1, seed=5)


def test_gltf_feature_scene_with_textures(emu, orc, gpurt):
    """tests/data/synth/features.gltf: albedo / metal-rough / normal / emissive textures, two lights"""
    path = os.path.join(ROOT, "tests", "data", "synth", "features.gltf")
    cam = gpurt.camera(1, 80, 60, (4.0, 3.0, 6.0), (1.0, 1.0, 2.0), 60.0)
    for integ in (0, 1, 2, 4):
        scene = gpurt.Scene(None).load(path)
        texs = [scene.texture(i) for i in range(scene.counts()["textures"])]
        img = _run(emu, orc, gpurt, scene, 80, 60, 2, cam=cam, textures=texs, integrator=integ, brdf=integ % 2,
                   samples_per_frame=2, max_depth=3, use_normal_map=1, use_metalness=1, env_scale=0.5, seed=77 + integ)
        assert img[..., :3].mean() > 0.001
        scene.close()


def test_sponza_standin_config2(emu, orc, gpurt):
    s = gpurt.Scene(None)
    s.make_sponza_standin()
    cam = gpurt.camera(1, 96, 54, (-1000.0, 200.0, 0.0), (0.0, 200.0, 0.0), 90.0)
    _run(emu, orc, gpurt, s, 96, 54, 2, cam=cam, integrator=1, brdf=1, samples_per_frame=1, max_depth=2, use_rr=0,
         env_scale=1.0, seed=3)


def _light_rays(rs, rng, n):
    """rays that stress light_pdf: towards points on light triangles (most hit), from light surfaces, axis-parallel,
    grazing, random"""
    lights = rs.lights.reshape(-1, 12)
    tris = []
    for l in lights:
        obj, nt = int(l[8]), int(l[9])
        t0 = int(rs.tri_off[obj])
        tris.append(rs.tris[t0:t0 + nt])
    tris = np.concatenate(tris).reshape(-1, 3, 3)
    box = rs.bvh.scene_box()
    lo, hi = box[:3], box[3:]
    o = (lo + (hi - lo) * rng.random((n, 3), dtype=np.float32)).astype(np.float32)
    pick = tris[rng.integers(0, len(tris), n)]
    b = rng.random((n, 3), dtype=np.float32)
    kind = rng.integers(0, 8, n)
    b[kind == 1] = [1, 0, 0]                      # exactly at a vertex
    b[kind == 2, 2] = 0                           # on an edge
    b /= b.sum(axis=1, keepdims=True)
    target = (pick * b[:, :, None]).sum(axis=1).astype(np.float32)
    d = target - o
    on_light = kind == 3                          # origin on a light triangle, random direction
    o[on_light] = target[on_light]
    d[on_light] = rng.standard_normal((int(on_light.sum()), 3)).astype(np.float32)
    rnd = kind == 4
    d[rnd] = rng.standard_normal((int(rnd.sum()), 3)).astype(np.float32)
    ax = kind == 5                                # axis-parallel through the target
    k = rng.integers(0, 3, n)
    for a in range(3):
        m = ax & (k == a)
        o[m] = target[m]
        o[m, a] = lo[a] - 0.1 * (hi[a] - lo[a])
        d[m] = 0
        d[m, a] = 1
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-30).astype(np.float32)
    return np.concatenate([o, d.astype(np.float32)], axis=1).astype(np.float32)


def _sliver_light_scene(gpurt, scale, seed):
    """emissive meshes with slivers, tiny and huge triangles at a Sponza-like coordinate scale"""
    rng = np.random.default_rng(seed)
    s = gpurt.Scene(None)
    e = gpurt.Material()
    e.albedo[:] = (1, 1, 1)
    e.emissive[:] = (3, 3, 3)
    e.albedo_tex = e.emissive_tex = e.metal_rough_tex = e.normal_tex = -1
    e.metal_rough[:] = (0, 1)
    for n_tris in (1, 7, 8, 9, 63, 64, 65, 130, 700):
        c = (rng.random(3) * scale).astype(np.float32)
        tris = c + (rng.random((n_tris, 3, 3)) - 0.5) * scale * 0.05
        thin = rng.random(n_tris) < 0.4           # slivers: third vertex almost on the first edge
        w = rng.random(n_tris)[:, None]
        tris[thin, 2] = tris[thin, 0] * (1 - w[thin]) + tris[thin, 1] * w[thin] + (rng.random((int(thin.sum()), 3)) - 0.5) * scale * 1e-5
        tiny = rng.random(n_tris) < 0.1
        tris[tiny] = tris[tiny, :1] + (tris[tiny] - tris[tiny, :1]) * 1e-3
        s.add_triangles(tris.reshape(-1, 9).astype(np.float32), e)
    d = gpurt.Material()
    d.albedo[:] = (0.5, 0.5, 0.5)
    d.albedo_tex = d.emissive_tex = d.metal_rough_tex = d.normal_tex = -1
    d.metal_rough[:] = (0, 0.5)
    floor = np.array([[0, 0, 0, 1, 0, 0, 0, 0, 1], [1, 0, 0, 1, 0, 1, 0, 0, 1]], np.float32) * scale
    s.add_triangles(floor, d)
    return s


@pytest.mark.parametrize("name", ["mis_test", "cbox", "features", "slivers_1", "slivers_2000"])
def test_light_groups_do_not_change_light_pdf(emu, orc, gpurt, name):
    """light_pdf through the light-run boxes and through the light BVH == light_pdf over every triangle, bit for bit"""
    if name.startswith("slivers"):
        s = _sliver_light_scene(gpurt, float(name.split("_")[1]), 11)
    elif name == "features":
        s = gpurt.Scene(None).load(os.path.join(ROOT, "tests", "data", "synth", "features.gltf"))
    else:
        s = gpurt.Scene(None).load(os.path.join(MEDIA, name, name + ".gltf"))
    es = EmuScene(emu, orc, s)
    assert es.rs.n_lights > 0
    n = 300000
    rays = _light_rays(es.rs, np.random.default_rng(3), n)
    out = np.zeros((n, 3), np.float32)
    emu.emu_light_pdf(es.h, C.byref(es.args), _vp(rays), C.c_ulonglong(n), _vp(out))
    a, b, c = (out[:, k].view(np.uint32) for k in range(3))
    assert (a == b).all(), f"{(a != b).sum()} of {n} light_pdf values differ (light-run boxes)"
    assert es.lh and (c == b).all(), f"{(c != b).sum()} of {n} light_pdf values differ (light BVH)"
    assert (out[:, 1] > 0).mean() > 0.2
    es.close()


@pytest.mark.parametrize("bvh,groups,verts", [(0, 1, 0), (0, 0, 0), (1, 0, 1)])
def test_fallback_paths_render_the_same(emu, orc, gpurt, bvh, groups, verts):
    """the A/B knobs of render.cu (GPURT_LIGHT_BVH / GPURT_LIGHT_GROUPS / GPURT_LIGHT_VERTS) select code paths
    that must all equal the oracle: light_pdf by scan / by the GLSL's full loop, light_sample with its own transforms"""
    s = gpurt.Scene(None).load(os.path.join(MEDIA, "mis_test", "mis_test.gltf"))
    cam = gpurt.camera(1, 64, 36, (0.5, 0.6, 2.6), (0.5, 0.45, 0.0), 50.0)
    emu.emu_set_light_bvh(bvh), emu.emu_set_light_groups(groups), emu.emu_set_light_verts(verts)
    try:
        for integ in (0, 2, 4):
            _run(emu, orc, gpurt, s, 64, 36, 2, cam=cam, integrator=integ, brdf=1, samples_per_frame=1, max_depth=3,
                 seed=40 + integ)
    finally:
        emu.emu_set_light_bvh(1), emu.emu_set_light_groups(1), emu.emu_set_light_verts(1)


def test_more_light_hits_than_the_buffer_holds(emu, orc, gpurt):
    """a ray through a stack of 12 emissive quads crosses more light triangles than light_pdf_bvh keeps (8):
    it must fall back to the scan and still equal the full loop"""
    s = gpurt.Scene(None)
    e = gpurt.Material()
    e.albedo[:] = (1, 1, 1)
    e.emissive[:] = (2, 2, 2)
    e.albedo_tex = e.emissive_tex = e.metal_rough_tex = e.normal_tex = -1
    e.metal_rough[:] = (0, 1)
    for k in range(12):
        z = 0.1 * k
        s.add_triangles(np.array([[0, 0, z, 1, 0, z, 1, 1, z], [0, 0, z, 1, 1, z, 0, 1, z]], np.float32), e)
    es = EmuScene(emu, orc, s)
    rng = np.random.default_rng(9)
    n = 20000
    o = np.concatenate([rng.random((n, 2)) * 0.8 + 0.1, np.full((n, 1), -0.5)], axis=1).astype(np.float32)
    d = np.concatenate([(rng.random((n, 2)) - 0.5) * 0.2, np.ones((n, 1))], axis=1).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d], axis=1).astype(np.float32)
    out = np.zeros((n, 3), np.float32)
    emu.emu_light_pdf(es.h, C.byref(es.args), _vp(rays), C.c_ulonglong(n), _vp(out))
    assert (out[:, 2].view(np.uint32) == out[:, 1].view(np.uint32)).all()
    assert (out[:, 0].view(np.uint32) == out[:, 1].view(np.uint32)).all()
    assert (out[:, 1] > 0).mean() > 0.9
    es.close()


@pytest.mark.parametrize("w,h,band,shards", [(1920, 1080, 0, 1), (1920, 1080, 16, 8), (192, 108, 16, 2), (100, 37, 16, 3),
                                             (64, 36, 4, 4), (33, 17, 5, 2), (3840, 2160, 16, 8), (16, 8, 8, 1)])
def test_shard_pixel_enumerates_every_pixel_exactly_once(emu, w, h, band, shards):
    """the pixel order the frame kernels use (8x4 warp tiles inside 16x8 CTA blocks, interleaved row bands per rank):
    the shards partition the frame, and a warp's 32 pixels form an 8x4 tile whenever the tiled order applies"""
    emu.emu_shard_pixels.restype = C.c_uint
    seen = np.zeros(w * h, np.int32)
    for r in range(shards):
        n = emu.emu_shard_pixels(w, h, band, shards, r, None)
        out = np.zeros(max(n, 1), np.uint32)
        assert emu.emu_shard_pixels(w, h, band, shards, r, _vp(out)) == n
        out = out[:n]
        assert (out < w * h).all()
        np.add.at(seen, out, 1)
        rows = out // w
        band_rows = band if band else h
        assert ((rows // band_rows) % shards == r).all(), "a rank only renders its own bands"
        if w % 16 == 0 and band_rows % 8 == 0 and h % 8 == 0 and n >= 32:
            x, y = (out[:32] % w).astype(int), (out[:32] // w).astype(int)
            assert x.max() - x.min() == 7 and y.max() - y.min() == 3
    assert (seen == 1).all()


def _random_material_scene(gpurt, seed, n_objs=24, degenerate=True):
    """small random scene: mirrors (roughness 0), rough and emissive objects, non-unit / zero normals, zero-area
    triangles (also among the lights), a tiny and a huge object"""
    rng = np.random.default_rng(seed)
    s = gpurt.Scene(None)
    for k in range(n_objs):
        n_tris = int(rng.integers(1, 12))
        c = rng.random(3) * 2 - 1
        size = [0.4, 0.4, 0.4, 0.02, 1.5][k % 5]
        tris = (c + (rng.random((n_tris, 3, 3)) - 0.5) * size).astype(np.float32)
        if degenerate and k % 4 == 1:
            tris[0, 2] = tris[0, 1]                       # zero-area triangle
        if degenerate and k % 6 == 2 and n_tris > 1:
            tris[1, 1] = tris[1, 0] + (tris[1, 2] - tris[1, 0]) * 0.5   # collinear
        verts = np.zeros((3 * n_tris, 12), np.float32)
        verts[:, 0:3] = tris.reshape(-1, 3)
        verts[:, 3], verts[:, 7] = rng.random(3 * n_tris), rng.random(3 * n_tris)
        nrm = rng.standard_normal((3 * n_tris, 3)) * [1.0, 3.0, 0.2][k % 3]
        if degenerate and k % 5 == 3:
            nrm[0] = 0                                     # zero normal -> normalize gives NaN, as in the GLSL
        verts[:, 4:7] = nrm
        verts[:, 8:11] = rng.standard_normal((3 * n_tris, 3))
        verts[:, 11] = rng.choice([-1.0, 1.0], 3 * n_tris)
        m = gpurt.Material()
        m.albedo[:] = tuple(rng.random(3))
        m.albedo_tex = m.emissive_tex = m.metal_rough_tex = m.normal_tex = -1
        kind = k % 4
        m.metal_rough[:] = (rng.random(), 0.0 if kind == 0 else float(rng.random() * 0.9 + 0.05))
        if kind == 3:
            m.emissive[:] = tuple(rng.random(3) * 5 + 0.1)
        model = np.eye(4, dtype=np.float32)
        model[:3, :3] += (rng.random((3, 3)).astype(np.float32) - 0.5) * 0.3
        model[:3, 3] = (rng.random(3) - 0.5) * 0.5
        s.add_object(verts, np.arange(3 * n_tris, dtype=np.uint32), model.T.reshape(16).copy(), m)
    return s


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_material_scenes_all_integrators(emu, orc, gpurt, seed):
    """product shading code == oracle on scenes full of edge cases (NaN normals, zero-area lights, mirrors, skewed
    instances), every integrator and both BRDFs, two frames (temporal reuse), bit for bit incl. NaN payloads"""
    s = _random_material_scene(gpurt, seed)
    cam = gpurt.camera(1, 48, 27, (0.2, 0.1, 2.4), (0.0, 0.0, 0.0), 70.0)
    for integ in range(5):
        _run(emu, orc, gpurt, s, 48, 27, 2, cam=cam, integrator=integ, brdf=(integ + seed) % 2, samples_per_frame=2,
             max_depth=4, env_scale=0.3, use_metalness=seed % 2, seed=100 * seed + integ)


def _textured_quads_scene(gpurt):
    """albedo / emissive / metal-rough / normal textures, sRGB decode, bilinear + REPEAT with negative and > 1
    coordinates, an emissive-textured light (light_sample's texture path and the Q5 texcoord quirk)"""
    rng = np.random.default_rng(5)
    texs = [rng.integers(0, 256, (16, 16, 4), dtype=np.uint8), rng.integers(0, 256, (8, 32, 4), dtype=np.uint8),
            rng.integers(64, 256, (4, 4, 4), dtype=np.uint8), rng.integers(100, 156, (32, 32, 4), dtype=np.uint8)]
    texs[3][..., 2] = 250
    scene = gpurt.Scene(None)
    for t in texs:
        scene.add_texture(t)

    def quad(z, size, mat, uvscale=3.0):
        v = np.zeros((4, 12), np.float32)
        v[:, 0:3] = [[-size, -size, z], [size, -size, z], [size, size, z], [-size, size, z]]
        v[:, 3] = np.array([0, 1, 1, 0]) * uvscale - 0.7
        v[:, 7] = np.array([0, 0, 1, 1]) * uvscale - 0.3
        v[:, 4:7] = [0, 0, 1]
        v[:, 8:12] = [1, 0, 0, 1]
        scene.add_object(v, np.array([0, 1, 2, 0, 2, 3], np.uint32), None, mat)

    m = gpurt.Material()
    m.albedo[:] = (1, 1, 1)
    m.albedo_tex, m.emissive_tex, m.metal_rough_tex, m.normal_tex = 0, -1, 2, 3
    m.metal_rough[:] = (0.5, 0.5)
    quad(0.0, 2.0, m)
    e = gpurt.Material()
    e.albedo[:] = (1, 1, 1)
    e.emissive[:] = (4, 4, 4)
    e.albedo_tex, e.emissive_tex, e.metal_rough_tex, e.normal_tex = -1, 1, -1, -1
    e.metal_rough[:] = (0, 1)
    quad(3.0, 0.8, e, uvscale=1.0)
    return scene, texs


def test_textured_quads_all_texture_kinds(emu, orc, gpurt):
    scene, texs = _textured_quads_scene(gpurt)
    cam = gpurt.camera(1, 64, 48, (1.5, 1.0, 2.5), (0.0, 0.0, 0.5), 70.0)
    for integ in (0, 1, 2, 3, 4):
        img = _run(emu, orc, gpurt, scene, 64, 48, 2, cam=cam, textures=texs, integrator=integ, brdf=1, samples_per_frame=2,
                   max_depth=3, use_normal_map=1, use_metalness=1, seed=20 + integ)
        assert np.isfinite(img[..., :3]).mean() > 0.9
    scene.close()


def test_product_shading_code_equals_the_reference_shader_text(emu, orc, gpurt):
    """the product's device code (shade.cuh + bvh8 / traverse, replayed on the CPU) against the digests of whole frames
    rendered by the REFERENCE'S OWN rt.rgen compiled as C++ (tests/golden/glsl_frames_golden.json, see
    tests/test_oracle.py): image, G-buffers, reservoirs and ray counts of every case, without the oracle's integrator
    in between (its BVH only served the reference run)"""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("make_glsl_golden", os.path.join(ROOT, "tests", "golden", "make_glsl_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    want = {k: v for k, v in json.load(open(os.path.join(ROOT, "tests", "golden", "glsl_frames_golden.json"))).items()
            if not k.startswith("tonemap/")}
    cache = {}

    def render(rs, st, consts, cam, seed, n_tex):
        if id(rs) not in cache:
            cache.clear()
            cache[id(rs)] = EmuScene(emu, orc, rs.gscene, rs.textures)
        frame = int(np.asarray(consts, np.uint32)[8])
        return cache[id(rs)].render_frame(st, consts, cam, seed ^ frame)
    got = mg.frame_digests(gpurt, orc, render)
    assert set(got) == set(want)
    bad = [k for k in want if got[k] != want[k]]
    assert not bad, f"{len(bad)} of {len(want)} frame buffers differ from the reference shader's: {bad[:5]}"

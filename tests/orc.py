"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = os.path.join(ORACLE_DIR, "liboracle.so")

NONE = 0xFFFFFFFF

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def _load():
    if not os.path.exists(_LIB):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"])
    lib = C.CDLL(_LIB)
    sig = {
        "orc_tea": (C.c_uint32, [C.c_uint32, C.c_uint32]),
        "orc_lcg": (C.c_uint32, [C.POINTER(C.c_uint32)]),
        "orc_randf": (C.c_float, [C.POINTER(C.c_uint32)]),
        "orc_radical_inverse": (C.c_float, [C.c_uint32]),
        "orc_flatten_object": (None, [_f32p, _u32p, C.c_uint32, _f32p, _f32p]),
        "orc_intersect": (C.c_int, [_f32p, _f32p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "orc_closest_point_tri": (C.c_float, [_f32p, _f32p, _f32p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "orc_closest_hit_brute": (None, [_f32p, C.c_uint32, _f32p, C.c_uint64, _u32p, C.c_int]),
        "orc_any_hit_brute": (None, [_f32p, C.c_uint32, _f32p, C.c_uint64, _u8p, C.c_int]),
        "orc_closest_point_brute": (None, [_f32p, C.c_uint32, _f32p, C.c_uint64, _u32p, C.c_int]),
        "orc_bvh_build": (C.c_void_p, [_f32p, C.c_uint32]),
        "orc_bvh_build_sah": (C.c_void_p, [_f32p, C.c_uint32]),
        "orc_bvh_free": (None, [C.c_void_p]),
        "orc_bvh_n_tris": (C.c_uint32, [C.c_void_p]),
        "orc_bvh_scene_box": (None, [C.c_void_p, _f32p]),
        "orc_bvh_morton_keys": (None, [C.c_void_p, _u64p]),
        "orc_bvh_prim_order": (None, [C.c_void_p, _u32p]),
        "orc_bvh_get_bvh2": (None, [C.c_void_p, _i32p, _i32p, _f32p]),
        "orc_bvh_inflation": (C.c_float, [C.c_void_p]),
        "orc_bvh_closest_hit": (None, [C.c_void_p, _f32p, C.c_uint64, _u32p, C.c_int]),
        "orc_bvh_any_hit": (None, [C.c_void_p, _f32p, C.c_uint64, _u8p, C.c_int]),
        "orc_bvh_closest_point": (None, [C.c_void_p, _f32p, C.c_uint64, _u32p, C.c_int]),
        "orc_gen_random_rays": (None, [C.c_uint64, C.c_uint32, _f32p, C.c_float, C.c_float, C.c_float, _f32p]),
        "orc_gen_random_points": (None, [C.c_uint64, C.c_uint32, _f32p, C.c_float, C.c_float, _f32p]),
        "orc_hw_threads": (C.c_int, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()

HIT_DT = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("gid", "<u4")])
CPQ_DT = np.dtype([("p", "<f4", 3), ("dist", "<f4"), ("gid", "<u4"), ("obj", "<u4"), ("u", "<f4"), ("v", "<f4")])


def tea(a, b):
    return lib.orc_tea(a, b)


def flatten(verts48, idx, model16):
    """verts48: (nv,12) f32, idx: (3*nt,) u32, model16: column-major 16 f32 -> (nt,9) f32"""
    verts48 = np.ascontiguousarray(verts48, np.float32).reshape(-1, 12)
    idx = np.ascontiguousarray(idx, np.uint32).reshape(-1)
    nt = idx.size // 3
    out = np.empty((nt, 9), np.float32)
    lib.orc_flatten_object(verts48, idx, nt, np.ascontiguousarray(model16, np.float32).reshape(16), out)
    return out


def closest_hit_brute(tris9, rays, threads=0):
    hits = np.empty((rays.shape[0], 4), np.uint32)
    lib.orc_closest_hit_brute(tris9, tris9.shape[0], rays, rays.shape[0], hits, threads)
    return hits.view(HIT_DT).reshape(-1)


def any_hit_brute(tris9, rays, threads=0):
    occ = np.empty(rays.shape[0], np.uint8)
    lib.orc_any_hit_brute(tris9, tris9.shape[0], rays, rays.shape[0], occ, threads)
    return occ


def closest_point_brute(tris9, queries, threads=0):
    res = np.empty((queries.shape[0], 8), np.uint32)
    lib.orc_closest_point_brute(tris9, tris9.shape[0], queries, queries.shape[0], res, threads)
    return res.view(CPQ_DT).reshape(-1)


class Bvh:
    """sah=False: the Morton LBVH (contract N6, GPURT_BUILD_LBVH); sah=True: the binned-SAH split of the default build (N6')"""

    def __init__(self, tris9, sah=False):
        self.tris9 = np.ascontiguousarray(tris9, np.float32).reshape(-1, 9)
        self.n = self.tris9.shape[0]
        self.h = (lib.orc_bvh_build_sah if sah else lib.orc_bvh_build)(self.tris9, self.n)

    def __del__(self):
        if getattr(self, "h", None):
            lib.orc_bvh_free(self.h)
            self.h = None

    def scene_box(self):
        b = np.empty(6, np.float32)
        lib.orc_bvh_scene_box(self.h, b)
        return b

    def inflation(self):
        return lib.orc_bvh_inflation(self.h)

    def keys(self):
        k = np.empty(self.n, np.uint64)
        lib.orc_bvh_morton_keys(self.h, k)
        return k

    def prim_order(self):
        o = np.empty(self.n, np.uint32)
        lib.orc_bvh_prim_order(self.h, o)
        return o

    def bvh2(self):
        m = max(self.n - 1, 0)
        l = np.empty(m, np.int32)
        r = np.empty(m, np.int32)
        b = np.empty((m, 6), np.float32)
        lib.orc_bvh_get_bvh2(self.h, l, r, b)
        return l, r, b

    def closest_hit(self, rays, threads=0):
        hits = np.empty((rays.shape[0], 4), np.uint32)
        lib.orc_bvh_closest_hit(self.h, rays, rays.shape[0], hits, threads)
        return hits.view(HIT_DT).reshape(-1)

    def any_hit(self, rays, threads=0):
        occ = np.empty(rays.shape[0], np.uint8)
        lib.orc_bvh_any_hit(self.h, rays, rays.shape[0], occ, threads)
        return occ

    def closest_point(self, queries, threads=0):
        res = np.empty((queries.shape[0], 8), np.uint32)
        lib.orc_bvh_closest_point(self.h, queries, queries.shape[0], res, threads)
        return res.view(CPQ_DT).reshape(-1)


def gen_random_rays(n, seed, box6, frac=0.1, tmin=1e-5, tmax=1e7):
    rays = np.empty((n, 8), np.float32)
    lib.orc_gen_random_rays(n, seed, np.ascontiguousarray(box6, np.float32), frac, tmin, tmax, rays)
    return rays


def gen_random_points(n, seed, box6, frac=0.25, r2=np.inf):
    q = np.empty((n, 4), np.float32)
    lib.orc_gen_random_points(n, seed, np.ascontiguousarray(box6, np.float32), frac, r2, q)
    return q


# ---- integrator oracle (oracle_render.cpp) ---------------------------------------------------
_vp = C.c_void_p
lib.orc_scene_create.restype = _vp
lib.orc_scene_create.argtypes = [C.c_uint32, _vp, _vp, _vp, _vp, _vp, C.c_uint32, _vp, C.c_uint32, _vp, _vp, _vp]
lib.orc_scene_free.argtypes = [_vp]
lib.orc_render_frame.restype = None
lib.orc_render_set_light_sampling.restype = None
lib.orc_render_set_light_sampling.argtypes = [C.c_uint32]
lib.orc_render_set_spatial.restype = None
lib.orc_render_set_spatial.argtypes = [C.c_uint32, C.c_float]
lib.orc_render_frame.argtypes = [_vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp, _vp, _vp, _vp, _vp, _vp,
                                 _vp, _vp, _vp, _vp, C.c_int]
lib.orc_record_rays.restype = None
lib.orc_record_rays.argtypes = [_vp, C.c_uint64]
lib.orc_recorded_rays.restype = C.c_uint64
lib.orc_recorded_rays.argtypes = []
lib.orc_camera_ray.restype = None
lib.orc_camera_ray.argtypes = [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                               C.POINTER(C.c_uint32), _f32p, _f32p]
lib.orc_mat_pdf.restype = C.c_float
lib.orc_mat_pdf.argtypes = [C.c_int, C.c_float, _f32p, _f32p, _f32p]
lib.orc_mat_eval.restype = None
lib.orc_mat_eval.argtypes = [C.c_int, _f32p, C.c_float, _f32p, _f32p, _f32p, _f32p]
lib.orc_tonemap.restype = None
lib.orc_tonemap.argtypes = [_f32p, C.c_uint64, C.c_int, C.c_float, C.c_float, _u8p]


def _p(a):
    return C.c_void_p(a.ctypes.data)


class RenderScene:
    """Oracle-side scene built from the packed arrays a gpurt.Scene exposes (host-only calls)."""

    def __init__(self, gscene, textures=()):
        descs = gscene.descs()
        self.n_objs = len(descs)
        self.descs = np.frombuffer(b"".join(bytes(d) for d in descs), np.uint32).copy() if descs else np.zeros(52, np.uint32)
        self.tri_off = gscene.tri_offsets().astype(np.uint32)
        objs = [gscene.object(i) for i in range(self.n_objs)]
        self.vert_off = np.concatenate([[0], np.cumsum([o[0].shape[0] for o in objs])]).astype(np.uint32)
        self.verts = np.concatenate([o[0] for o in objs]).astype(np.float32) if objs else np.zeros((1, 12), np.float32)
        self.idx = np.concatenate([o[1] for o in objs]).astype(np.uint32) if objs else np.zeros(3, np.uint32)
        lights = gscene.lights()
        self.n_lights = len(lights)
        self.lights = np.frombuffer(b"".join(bytes(l) for l in lights), np.uint32).copy() if lights else np.zeros(12, np.uint32)
        info, tex = [], []
        off = 0
        for t in textures:
            t = np.ascontiguousarray(t, np.uint8)
            info.append([off, t.shape[1], t.shape[0], 0])
            tex.append(t.reshape(-1))
            off += t.shape[0] * t.shape[1]
        self.tex_info = np.array(info, np.uint32).reshape(-1) if info else np.zeros(4, np.uint32)
        self.texels = np.concatenate(tex) if tex else np.zeros(4, np.uint8)
        self.tris = np.concatenate([flatten(objs[i][0], objs[i][1], np.array(descs[i].model, np.float32))
                                    for i in range(self.n_objs)]) if objs else np.zeros((0, 9), np.float32)
        self.bvh = Bvh(self.tris)
        self.h = lib.orc_scene_create(self.n_objs, _p(self.descs), _p(self.tri_off), _p(self.vert_off), _p(self.verts),
                                      _p(self.idx), self.n_lights, _p(self.lights), len(info), _p(self.tex_info),
                                      _p(self.texels), self.bvh.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib.orc_scene_free(self.h)
            self.h = None


class FrameState:
    """rt_target, ping-pong reservoirs and G-buffers (rt.cpp:178-220), oracle side"""

    def __init__(self, w, h):
        self.w, self.h = w, h
        self.image = np.zeros((h, w, 4), np.float32)
        self.res = [np.zeros((h * w, 12), np.uint32) for _ in range(2)]
        self.gb = [[np.zeros((h, w, 4), np.float32) for _ in range(3)] for _ in range(2)]
        self.parity = 0


def render_frame(rs, st, consts_words, camera_words, seed, threads=0, spatial=(0, 16.0), light_sampling=0):
    """one rt.rgen dispatch; consts/camera are the exact words the product used for the frame; spatial = (samples, radius)
    of the ReSTIR spatial-reuse extension (0 samples: the reference's estimator)"""
    lib.orc_render_set_spatial(int(spatial[0]), float(spatial[1]))
    lib.orc_render_set_light_sampling(int(light_sampling))   # extension: power-proportional light triangles
    cur, prev = st.parity, st.parity ^ 1
    counts = np.zeros(2, np.uint64)
    consts_words = np.ascontiguousarray(consts_words, np.uint32)
    camera_words = np.ascontiguousarray(camera_words, np.uint32)
    lib.orc_render_frame(rs.h, _p(consts_words), _p(camera_words), st.w, st.h, seed, _p(st.image), _p(st.res[prev]),
                         _p(st.res[cur]), _p(st.gb[prev][0]), _p(st.gb[prev][1]), _p(st.gb[prev][2]),
                         _p(st.gb[cur][0]), _p(st.gb[cur][1]), _p(st.gb[cur][2]), _p(counts), threads)
    st.parity ^= 1
    return counts


def frame_rays(rs, w, h, consts_words, camera_words, seed, capacity, threads=0):
    """render one frame on the CPU and return the closest-hit rays it traced, (n,8) f32 (thread order)"""
    buf = np.zeros((capacity, 8), np.float32)
    lib.orc_record_rays(_p(buf), capacity)
    try:
        render_frame(rs, FrameState(w, h), consts_words, camera_words, seed, threads)
        n = int(lib.orc_recorded_rays())
    finally:
        lib.orc_record_rays(None, 0)
    return buf[:n]


def tonemap(rgba, op=1, exposure=1.0, gamma=2.2):
    rgba = np.ascontiguousarray(rgba, np.float32)
    out = np.zeros(rgba.shape, np.uint8)
    lib.orc_tonemap(rgba.reshape(-1), rgba.size // 4, op, exposure, gamma, out.reshape(-1))
    return out

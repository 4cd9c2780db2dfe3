"""gpurt — Python binding (ctypes) of libgpurt.so, the B200-native GPU-RT hot path.

Mirrors the reference's host classes for this path: Scene (src/scene/scene.h), VK::Accel
(src/vk/vulkan.h:256-284) and VK::RTPipe (src/vk/rt.h:14-142).  Everything goes through the C ABI
declared in include/gpurt.h; there is no Python or CPU fallback — if the shared library is missing
the import fails, and compute calls fail without a B200.

numpy arrays are passed as GPURT_MEM_HOST; torch CUDA tensors as GPURT_MEM_DEVICE (in place, on the
tensor's current stream).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPURT_LIB") or os.path.join(os.path.dirname(_HERE), "libgpurt.so")  # GPURT_LIB: tuning builds

MEM_HOST, MEM_DEVICE = 0, 1
NO_HIT = 0xFFFFFFFF
HISTORY_ALL_ROWS = 0xFFFFFFFF
BUILD_DEFAULT, BUILD_KEEP_BVH2, BUILD_SAH_COLLAPSE, BUILD_SAH_SPLIT, BUILD_LBVH = 0, 1, 2, 4, 8

RAY_DT = np.dtype([("o", "<f4", 3), ("tmin", "<f4"), ("d", "<f4", 3), ("tmax", "<f4")])
HIT_DT = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<u4")])
QUERY_DT = np.dtype([("p", "<f4", 3), ("r2", "<f4")])
CPQ_DT = np.dtype([("p", "<f4", 3), ("dist", "<f4"), ("prim", "<u4"), ("obj", "<u4"), ("u", "<f4"), ("v", "<f4")])


class GpurtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gpurt error {code}: {msg}")
        self.code = code


class Material(C.Structure):
    _fields_ = [("albedo", C.c_float * 3), ("albedo_tex", C.c_int32), ("emissive", C.c_float * 3),
                ("emissive_tex", C.c_int32), ("metal_rough", C.c_float * 2), ("metal_rough_tex", C.c_int32),
                ("normal_tex", C.c_int32)]


class SceneDesc(C.Structure):
    _fields_ = [("model", C.c_float * 16), ("modelIT", C.c_float * 16), ("albedo", C.c_float * 4),
                ("emissive", C.c_float * 4), ("metal_rough", C.c_float * 4), ("albedo_tex", C.c_int32),
                ("emissive_tex", C.c_int32), ("metal_rough_tex", C.c_int32), ("normal_tex", C.c_int32),
                ("index", C.c_uint32), ("_pad", C.c_uint32 * 3)]


class SceneLight(C.Structure):
    _fields_ = [("bmin", C.c_float * 4), ("bmax", C.c_float * 4), ("index", C.c_uint32),
                ("n_triangles", C.c_uint32), ("_pad", C.c_uint32 * 2)]


class Camera(C.Structure):
    _fields_ = [("V", C.c_float * 16), ("P", C.c_float * 16), ("iV", C.c_float * 16), ("iP", C.c_float * 16),
                ("prev_PV", C.c_float * 16), ("new_samples", C.c_uint32), ("temporal_multiplier", C.c_uint32)]


class PipeParams(C.Structure):
    _fields_ = [("max_frames", C.c_int32), ("samples_per_frame", C.c_int32), ("max_depth", C.c_int32),
                ("clear", C.c_float * 3), ("env", C.c_float * 3), ("env_scale", C.c_float),
                ("use_normal_map", C.c_int32), ("use_rr", C.c_int32), ("use_metalness", C.c_int32),
                ("use_qmc", C.c_int32), ("use_temporal", C.c_int32), ("integrator", C.c_int32),
                ("temporal_scale", C.c_int32), ("brdf", C.c_int32), ("debug_view", C.c_int32),
                ("res_samples", C.c_int32), ("seed", C.c_uint32), ("spatial_samples", C.c_int32), ("spatial_radius", C.c_float),
                ("light_sampling", C.c_int32)]


class AccelInfo(C.Structure):
    _fields_ = [("n_tris", C.c_uint32), ("n_objs", C.c_uint32), ("n_bvh2_nodes", C.c_uint32),
                ("n_wide_nodes", C.c_uint32), ("wide_depth", C.c_uint32), ("scene_min", C.c_float * 3),
                ("scene_max", C.c_float * 3), ("inflation", C.c_float), ("build_ms", C.c_float),
                ("node_bytes", C.c_uint64), ("tri_bytes", C.c_uint64), ("tree_cost", C.c_float), ("tree_cost_at_build", C.c_float),
                ("refits", C.c_uint32), ("reserved", C.c_uint32)]


class TraceStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64),
                ("hits", C.c_uint64)]


# every symbol declared in include/gpurt.h (tests check the header against this list)
SYMBOLS = [
    "gpurt_last_error", "gpurt_version", "gpurt_ctx_create", "gpurt_ctx_destroy", "gpurt_ctx_set_stream",
    "gpurt_ctx_synchronize", "gpurt_scene_create", "gpurt_scene_destroy", "gpurt_scene_load_gltf",
    "gpurt_scene_make_sponza_standin", "gpurt_scene_add_object", "gpurt_scene_add_texture",
    "gpurt_scene_counts", "gpurt_scene_tri_offsets", "gpurt_scene_get_descs", "gpurt_scene_get_lights",
    "gpurt_scene_object_sizes", "gpurt_scene_get_object", "gpurt_camera_make", "gpurt_accel_build",
    "gpurt_accel_destroy", "gpurt_accel_info", "gpurt_accel_get_prim_order", "gpurt_accel_get_morton_keys",
    "gpurt_accel_get_bvh2", "gpurt_trace_closest", "gpurt_trace_any", "gpurt_closest_points",
    "gpurt_trace_closest_bvh2", "gpurt_trace_closest_stats", "gpurt_closest_points_stats", "gpurt_last_kernel_ms",
    "gpurt_pipe_params_default", "gpurt_pipe_create", "gpurt_pipe_destroy", "gpurt_pipe_reset_frame",
    "gpurt_pipe_render_frame", "gpurt_pipe_frame_index", "gpurt_pipe_read_image", "gpurt_pipe_read_gbuffer",
    "gpurt_pipe_read_image_async", "gpurt_pipe_read_image_wait", "gpurt_write_exr",
    "gpurt_pipe_ray_counts", "gpurt_pipe_device_image", "gpurt_tonemap", "gpurt_pipe_last_uniforms",
    "gpurt_pipe_read_reservoirs", "gpurt_pipe_bounce_rays", "gpurt_pipe_set_shard",
    "gpurt_pipe_history_export", "gpurt_pipe_history_peers", "gpurt_pipe_history_status",
    "gpurt_gather_create", "gpurt_gather_open", "gpurt_gather_results", "gpurt_gather_base", "gpurt_gather_begin", "gpurt_gather_end",
    "gpurt_gather_destroy",
    "gpurt_pipe_render_frame_mean", "gpurt_pipe_accumulate_mean", "gpurt_scene_set_transform", "gpurt_accel_update", "gpurt_scene_get_texture", "gpurt_shared_alloc", "gpurt_shared_free", "gpurt_shared_open", "gpurt_shared_close",
    "gpurt_scene_set_material", "gpurt_scene_set_ordered", "gpurt_scene_clear_textures", "gpurt_accel_sync_scene", "gpurt_accel_refit", "gpurt_accel_update_auto",
]


class Constants(C.Structure):
    _fields_ = [("clear_col", C.c_float * 4), ("env_light", C.c_float * 4)] + [
        (n, C.c_int32) for n in ("frame", "samples", "max_frame", "qmc", "max_depth", "use_normal_map",
                                 "use_metalness", "use_temporal", "integrator", "brdf", "debug_view", "use_rr",
                                 "n_lights", "n_objs")]

if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(or make -C gpu-rt_b200). There is no fallback path.")
lib = C.CDLL(LIB_PATH)
lib.gpurt_last_error.restype = C.c_char_p
lib.gpurt_version.restype = C.c_char_p


def _check(rc):
    if rc < 0:
        raise GpurtError(rc, lib.gpurt_last_error().decode(errors="replace"))
    return rc


def last_error():
    return lib.gpurt_last_error().decode(errors="replace")


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class SharedBuffer:
    """A device allocation that other ranks' kernels write into directly (gpurt_shared_*): rank 0 owns
    it, the other ranks map it over NVLink and pass `buf.at(byte_offset)` as the result pointer of
    their query call.  `.tensor()` views the owner's copy as a torch uint8 tensor without copying."""

    def __init__(self, ctx, ptr, nbytes, handle, owner):
        self.ctx, self.ptr, self.nbytes, self.handle, self.owner = ctx, ptr, nbytes, handle, owner

    def at(self, byte_offset):
        assert 0 <= byte_offset <= self.nbytes
        return _RawDevicePtr(self.ptr + byte_offset, self)

    def tensor(self):
        import torch
        return torch.as_tensor(self, device=f"cuda:{self.ctx.device}")

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 2}

    def close(self):
        if self.ptr:
            fn = lib.gpurt_shared_free if self.owner else lib.gpurt_shared_close
            _check(fn(self.ctx.h, C.c_void_p(self.ptr)))
            self.ptr = 0


class Gather:
    """One result array on the owner rank, filled by one query call per rank and round (include/gpurt.h, gpurt_gather_*).
    Owner: Gather.create(ctx, ...); others: Gather.open(ctx, handle or base, ...).  `g.mine()` is the result argument of this
    rank's query call; the owner brackets its call with g.begin() / g.end(); `g.tensor()` views the whole array (owner)."""

    def __init__(self, ctx, h, n_records, record_bytes, first, rank, owner, handle=None):
        self.ctx, self.h, self.n_records, self.record_bytes, self.first, self.rank, self.owner = ctx, h, n_records, record_bytes, list(first), rank, owner
        self.handle = handle
        a, m = C.c_void_p(), C.c_void_p()
        _check(lib.gpurt_gather_results(self.h, C.byref(a), C.byref(m)))
        self.array_ptr, self.mine_ptr = a.value, m.value

    @staticmethod
    def create(ctx, n_records, record_bytes, first, owner_rank=0):
        g, hd, nb = C.c_void_p(), (C.c_uint8 * 64)(), C.c_uint64()
        arr = (C.c_uint64 * len(first))(*first)
        _check(lib.gpurt_gather_create(ctx.h, C.c_uint64(n_records), record_bytes, len(first) - 1, arr, owner_rank, C.byref(g), hd, C.byref(nb)))
        out = Gather(ctx, g, n_records, record_bytes, first, owner_rank, True, bytes(hd))
        out.nbytes = nb.value
        return out

    @staticmethod
    def open(ctx, n_records, record_bytes, first, my_rank, owner_rank=0, handle=None, base=None):
        g = C.c_void_p()
        arr = (C.c_uint64 * len(first))(*first)
        hd = (C.c_uint8 * 64).from_buffer_copy(handle) if handle is not None else None
        _check(lib.gpurt_gather_open(ctx.h, hd, C.c_void_p(base or 0), C.c_uint64(n_records), record_bytes, len(first) - 1, arr,
                                     owner_rank, my_rank, C.byref(g)))
        return Gather(ctx, g, n_records, record_bytes, first, my_rank, False)

    def base(self):
        p = C.c_void_p()
        _check(lib.gpurt_gather_base(self.h, C.byref(p)))
        return p.value

    def mine(self):
        return _RawDevicePtr(self.mine_ptr, self)

    def begin(self):
        _check(lib.gpurt_gather_begin(self.h))

    def end(self, sync=False):
        t = C.c_uint32()
        _check(lib.gpurt_gather_end(self.h, C.byref(t) if sync else None))
        return t.value

    def tensor(self):
        """the whole result array as a torch uint8 tensor (no copy)"""
        import torch

        class _Arr:
            pass

        a = _Arr()
        a.__cuda_array_interface__ = {"shape": (self.n_records * self.record_bytes,), "typestr": "|u1", "data": (self.array_ptr, False), "version": 2}
        a.keep = self
        return torch.as_tensor(a, device=f"cuda:{self.ctx.device}")

    def close(self):
        if self.h:
            _check(lib.gpurt_gather_destroy(self.h))
            self.h = C.c_void_p()


class _RawDevicePtr:
    def __init__(self, ptr, keep):
        self.ptr, self.keep = ptr, keep


def _ptr(x, nbytes_min=0):
    """(void* pointer, mem flag, keepalive) of a numpy array, torch tensor or SharedBuffer offset"""
    if isinstance(x, _RawDevicePtr):
        return C.c_void_p(x.ptr), MEM_DEVICE, x
    if _is_torch(x):
        assert x.is_contiguous()
        if x.is_cuda:
            return C.c_void_p(x.data_ptr()), MEM_DEVICE, x
        x = x.numpy()
    assert isinstance(x, np.ndarray) and x.flags["C_CONTIGUOUS"], "need a C-contiguous numpy array"
    return C.c_void_p(x.ctypes.data), MEM_HOST, x


class Context:
    """One per GPU (replaces the VK::Manager singleton, src/vk/vulkan.cpp:17-20)."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        _check(lib.gpurt_ctx_create(int(device), C.byref(self.h)))
        self.device = device

    def use_torch_stream(self):
        import torch
        _check(lib.gpurt_ctx_set_stream(self.h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    def synchronize(self):
        _check(lib.gpurt_ctx_synchronize(self.h))

    def last_kernel_ms(self):
        ms = C.c_float()
        _check(lib.gpurt_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def shared_alloc(self, nbytes):
        """owner side of a cross-process result buffer (include/gpurt.h, multi-GPU result placement)"""
        p, h = C.c_void_p(), (C.c_uint8 * 64)()
        _check(lib.gpurt_shared_alloc(self.h, C.c_uint64(nbytes), C.byref(p), h))
        return SharedBuffer(self, p.value, nbytes, bytes(h), True)

    def shared_open(self, handle, nbytes):
        p, h = C.c_void_p(), (C.c_uint8 * 64).from_buffer_copy(handle)
        _check(lib.gpurt_shared_open(self.h, h, C.byref(p)))
        return SharedBuffer(self, p.value, nbytes, bytes(handle), False)

    def close(self):
        if self.h:
            lib.gpurt_ctx_destroy(self.h)
            self.h = C.c_void_p()


class Scene:
    """Scene (src/scene/scene.h:18-42) + RTPipe::build_desc packing (src/vk/rt.cpp:26-76)."""

    def __init__(self, ctx=None):
        self.ctx = ctx
        self.h = C.c_void_p()
        _check(lib.gpurt_scene_create(ctx.h if ctx else None, C.byref(self.h)))

    def load(self, path, scale=1.0):
        _check(lib.gpurt_scene_load_gltf(self.h, os.fsencode(path), C.c_float(scale)))
        return self

    def make_sponza_standin(self):
        _check(lib.gpurt_scene_make_sponza_standin(self.h))
        return self

    def add_object(self, verts48, indices, model16=None, material=None):
        verts48 = np.ascontiguousarray(verts48, np.float32).reshape(-1, 12)
        indices = np.ascontiguousarray(indices, np.uint32).reshape(-1)
        model = np.ascontiguousarray(np.eye(4, dtype=np.float32).reshape(16) if model16 is None else model16,
                                     np.float32).reshape(16)
        out = C.c_uint32()
        _check(lib.gpurt_scene_add_object(self.h, C.c_void_p(verts48.ctypes.data), verts48.shape[0],
                                          C.c_void_p(indices.ctypes.data), indices.size,
                                          C.c_void_p(model.ctypes.data),
                                          C.byref(material) if material is not None else None, C.byref(out)))
        return out.value

    def set_transform(self, obj, model16):
        """pose edit of object `obj` (for_objs index); follow with Accel.update()"""
        model = np.ascontiguousarray(model16, np.float32).reshape(16)
        _check(lib.gpurt_scene_set_transform(self.h, int(obj), C.c_void_p(model.ctypes.data)))

    def add_triangles(self, tris9, material=None):
        """convenience: (n,9) world-space triangle soup as one object with identity model"""
        tris9 = np.ascontiguousarray(tris9, np.float32).reshape(-1, 3, 3)
        v = np.zeros((tris9.shape[0] * 3, 12), np.float32)
        v[:, :3] = tris9.reshape(-1, 3)
        return self.add_object(v, np.arange(v.shape[0], dtype=np.uint32), None, material)

    def set_ordered(self, ordered=True):
        """objects in insertion order instead of the reference container's iteration order (gpurt_scene_set_ordered)"""
        _check(lib.gpurt_scene_set_ordered(self.h, 1 if ordered else 0))
        return self

    def set_material(self, obj, material):
        _check(lib.gpurt_scene_set_material(self.h, int(obj), C.byref(material)))

    def clear_textures(self):
        _check(lib.gpurt_scene_clear_textures(self.h))

    def add_texture(self, rgba8):
        rgba8 = np.ascontiguousarray(rgba8, np.uint8)
        h, w = rgba8.shape[:2]
        out = C.c_int32()
        _check(lib.gpurt_scene_add_texture(self.h, C.c_void_p(rgba8.ctypes.data), w, h, C.byref(out)))
        return out.value

    def texture(self, i):
        """decoded RGBA8 texels of texture i as an (h, w, 4) uint8 array"""
        w, h = C.c_uint32(), C.c_uint32()
        _check(lib.gpurt_scene_get_texture(self.h, int(i), C.byref(w), C.byref(h), None))
        out = np.zeros((h.value, w.value, 4), np.uint8)
        _check(lib.gpurt_scene_get_texture(self.h, int(i), None, None, C.c_void_p(out.ctypes.data)))
        return out

    def counts(self):
        a, b, c, d = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(lib.gpurt_scene_counts(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"objs": a.value, "tris": b.value, "lights": c.value, "textures": d.value}

    def tri_offsets(self):
        out = np.zeros(self.counts()["objs"] + 1, np.uint32)
        _check(lib.gpurt_scene_tri_offsets(self.h, C.c_void_p(out.ctypes.data)))
        return out

    def descs(self):
        n = self.counts()["objs"]
        arr = (SceneDesc * max(n, 1))()
        _check(lib.gpurt_scene_get_descs(self.h, arr))
        return arr[:n] if n else []

    def lights(self):
        n = self.counts()["lights"]
        arr = (SceneLight * max(n, 1))()
        _check(lib.gpurt_scene_get_lights(self.h, arr))
        return list(arr[:n])

    def object(self, i):
        nv, ni = C.c_uint32(), C.c_uint32()
        _check(lib.gpurt_scene_object_sizes(self.h, i, C.byref(nv), C.byref(ni)))
        v = np.zeros((nv.value, 12), np.float32)
        ix = np.zeros(ni.value, np.uint32)
        _check(lib.gpurt_scene_get_object(self.h, i, C.c_void_p(v.ctypes.data), C.c_void_p(ix.ctypes.data)))
        return v, ix

    def close(self):
        if self.h:
            lib.gpurt_scene_destroy(self.h)
            self.h = C.c_void_p()


def camera(mode=0, width=1280, height=720, pos=None, center=None, vfov=90.0):
    """Camera (src/util/camera.cpp:58-71, :150-155) -> V, P, iV, iP (src/vk/rt.cpp:121-127)."""
    cam = Camera()
    p = (C.c_float * 3)(*(pos or (0, 0, 0)))
    c = (C.c_float * 3)(*(center or (0, 0, 0)))
    _check(lib.gpurt_camera_make(mode, C.c_float(width), C.c_float(height), p, c, C.c_float(vfov), C.byref(cam)))
    return cam


def write_exr(path, rgba):
    """rgba: (h, w, 4) float32 in host memory -> OpenEXR file (uncompressed scanlines, 32-bit float A B G R)"""
    rgba = np.ascontiguousarray(rgba, np.float32)
    assert rgba.ndim == 3 and rgba.shape[2] == 4
    _check(lib.gpurt_write_exr(os.fsencode(path), rgba.ctypes.data_as(C.c_void_p), rgba.shape[1], rgba.shape[0]))


def pipe_params(**kw):
    p = PipeParams()
    _check(lib.gpurt_pipe_params_default(C.byref(p)))
    for k, v in kw.items():
        if k in ("clear", "env"):
            setattr(p, k, (C.c_float * 3)(*v))
        else:
            setattr(p, k, v)
    return p


class Accel:
    """VK::Accel (src/vk/vulkan.h:256-284): BLAS builds + TLAS build in one call."""

    def __init__(self, scene, flags=BUILD_DEFAULT):
        self.scene = scene
        self.ctx = scene.ctx
        self.h = C.c_void_p()
        # GPURT_BUILD_FLAGS (binding only): OR extra build flags into every Accel of a tool run, for A/B measurements
        # of GPURT_BUILD_SAH_COLLAPSE (2) / GPURT_BUILD_LBVH (8) without editing the tools
        flags |= int(os.environ.get("GPURT_BUILD_FLAGS", "0"))
        self.flags = flags
        _check(lib.gpurt_accel_build(scene.h, flags, C.byref(self.h)))

    def update(self):
        """rebuild after scene edits (GPURT::build_accel, src/gpurt.cpp:220-241)"""
        _check(lib.gpurt_accel_update(self.h))

    def info(self):
        i = AccelInfo()
        _check(lib.gpurt_accel_info(self.h, C.byref(i)))
        return i

    def prim_order(self):
        out = np.zeros(self.info().n_tris, np.uint32)
        _check(lib.gpurt_accel_get_prim_order(self.h, C.c_void_p(out.ctypes.data)))
        return out

    def morton_keys(self):
        out = np.zeros(self.info().n_tris, np.uint64)
        _check(lib.gpurt_accel_get_morton_keys(self.h, C.c_void_p(out.ctypes.data)))
        return out

    def bvh2(self):
        m = max(int(self.info().n_tris) - 1, 0)
        l, r, b = np.zeros(m, np.int32), np.zeros(m, np.int32), np.zeros((m, 6), np.float32)
        _check(lib.gpurt_accel_get_bvh2(self.h, C.c_void_p(l.ctypes.data), C.c_void_p(r.ctypes.data),
                                        C.c_void_p(b.ctypes.data)))
        return l, r, b

    def _query(self, fn, inp, n, out):
        pi, mi, _k1 = _ptr(inp)
        po, mo, _k2 = _ptr(out)
        assert mi == mo, "input and output must both be host or both be device"
        if mi == MEM_DEVICE:
            self.ctx.use_torch_stream()
        _check(fn(self.h, pi, C.c_uint64(n), po, mi))
        return out

    def _alloc_like(self, inp, n, dtype, width):
        if _is_torch(inp) and inp.is_cuda:
            import torch
            return torch.empty((n, width), dtype=torch.float32 if dtype != np.uint8 else torch.uint8,
                               device=inp.device)
        return np.zeros(n, dtype)

    def trace_closest(self, rays, out=None, bvh2=False):
        """traceRayEXT closest hit (rt.rgen:257-270). rays: (n,8) f32 / RAY_DT; returns HIT_DT or (n,4) tensor."""
        n = rays.shape[0]
        if out is None:
            out = self._alloc_like(rays, n, HIT_DT, 4)
        return self._query(lib.gpurt_trace_closest_bvh2 if bvh2 else lib.gpurt_trace_closest, rays, n, out)

    def trace_any(self, rays, out=None):
        n = rays.shape[0]
        if out is None:
            if _is_torch(rays) and rays.is_cuda:
                import torch
                out = torch.empty(n, dtype=torch.uint8, device=rays.device)
            else:
                out = np.zeros(n, np.uint8)
        return self._query(lib.gpurt_trace_any, rays, n, out)

    def closest_points(self, queries, out=None):
        n = queries.shape[0]
        if out is None:
            out = self._alloc_like(queries, n, CPQ_DT, 8)
        return self._query(lib.gpurt_closest_points, queries, n, out)

    def trace_stats(self, rays_dev, hits_dev):
        st = TraceStats()
        self.ctx.use_torch_stream()
        _check(lib.gpurt_trace_closest_stats(self.h, C.c_void_p(rays_dev.data_ptr()), C.c_uint64(rays_dev.shape[0]),
                                             C.c_void_p(hits_dev.data_ptr()), C.byref(st)))
        return st

    def refit(self):
        """pose-only edit: keep order and topology, refit boxes and wide nodes (gpurt_accel_refit)"""
        _check(lib.gpurt_accel_refit(self.h))

    def update_auto(self, max_cost_growth=0.0):
        _check(lib.gpurt_accel_update_auto(self.h, C.c_float(max_cost_growth)))

    def sync_scene(self):
        """materials / textures changed, geometry did not: refresh the device copy without rebuilding the BVH"""
        _check(lib.gpurt_accel_sync_scene(self.h))

    def closest_points_stats(self, queries_dev):
        st = TraceStats()
        self.ctx.use_torch_stream()
        _check(lib.gpurt_closest_points_stats(self.h, C.c_void_p(queries_dev.data_ptr()), C.c_uint64(queries_dev.shape[0]), C.byref(st)))
        return st

    def close(self):
        if self.h:
            lib.gpurt_accel_destroy(self.h)
            self.h = C.c_void_p()


class RTPipe:
    """VK::RTPipe (src/vk/rt.h:14-142): recreate(scene) + use_accel + update_uniforms + trace."""

    def __init__(self, scene, accel):
        self.scene, self.accel, self.ctx = scene, accel, scene.ctx
        self.h = C.c_void_p()
        _check(lib.gpurt_pipe_create(scene.h, accel.h, C.byref(self.h)))
        self.w = self.h_px = 0

    def reset_frame(self):
        _check(lib.gpurt_pipe_reset_frame(self.h))

    def set_shard(self, band_rows, n_shards, shard):
        """render only row bands with (band index % n_shards) == shard; band_rows=0: whole frame"""
        _check(lib.gpurt_pipe_set_shard(self.h, band_rows, n_shards, shard))

    def history_export(self, w, h):
        """(device pointer, 64-byte handle, bytes) of this shard's previous-frame block (gpurt_pipe_history_export)"""
        p, hd, nb = C.c_void_p(), (C.c_uint8 * 64)(), C.c_uint64()
        _check(lib.gpurt_pipe_history_export(self.h, w, h, C.byref(p), hd, C.byref(nb)))
        return p.value, bytes(hd), nb.value

    def history_peers(self, blocks, halo_rows=HISTORY_ALL_ROWS):
        """blocks[s] = device pointer of shard s's history block as mapped in this process (own entry ignored);
        [] switches the exchange off"""
        arr = (C.c_void_p * max(1, len(blocks)))(*[C.c_void_p(b or 0) for b in blocks])
        _check(lib.gpurt_pipe_history_peers(self.h, len(blocks), arr, C.c_uint32(halo_rows)))

    def history_status(self):
        """(frames pushed to the peers, flag waits that timed out)"""
        a, b = C.c_uint32(), C.c_uint32()
        _check(lib.gpurt_pipe_history_status(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def device_image(self):
        """torch view (H,W,4) of rt_target in device memory (no copy)"""
        import torch
        ptr = C.c_void_p()
        _check(lib.gpurt_pipe_device_image(self.h, C.byref(ptr)))

        class _Arr:
            pass

        a = _Arr()
        a.__cuda_array_interface__ = {"shape": (self.h_px, self.w, 4), "typestr": "<f4", "data": (ptr.value, False), "version": 2}
        return torch.as_tensor(a, device=f"cuda:{self.ctx.device}")

    def render_frame(self, params, cam, width, height):
        self.w, self.h_px = width, height
        return _check(lib.gpurt_pipe_render_frame(self.h, C.byref(params), C.byref(cam), width, height))

    def frame_index(self):
        f = C.c_int32()
        _check(lib.gpurt_pipe_frame_index(self.h, C.byref(f)))
        return f.value

    def render_frame_mean(self, params, cam, width, height, frame, mean_out):
        """frame-parallel sharding: render frame `frame`, write its per-pixel mean to mean_out (device tensor of
        w*h*4 f32 or SharedBuffer.at(offset))"""
        self.ctx.use_torch_stream()
        self.w, self.h_px = width, height
        ptr, mem, _keep = _ptr(mean_out)
        assert mem == MEM_DEVICE
        return _check(lib.gpurt_pipe_render_frame_mean(self.h, C.byref(params), C.byref(cam), width, height, int(frame), ptr))

    def accumulate_mean(self, mean, frame, width, height):
        self.ctx.use_torch_stream()
        ptr, mem, _keep = _ptr(mean)
        assert mem == MEM_DEVICE
        _check(lib.gpurt_pipe_accumulate_mean(self.h, ptr, int(frame), width, height))

    def read_image(self, out=None):
        if out is None:
            out = np.zeros((self.h_px, self.w, 4), np.float32)
        p, m, _k = _ptr(out)
        _check(lib.gpurt_pipe_read_image(self.h, p, m))
        return out

    def read_image_async(self, out):
        """Queue a copy of the image (as of the frames rendered so far) into page-locked host memory `out`; it
        overlaps the following frames.  read_image_wait() returns when it has landed."""
        p, m, _k = _ptr(out)
        if m != MEM_HOST:
            raise ValueError("read_image_async copies to host memory")
        _check(lib.gpurt_pipe_read_image_async(self.h, p))
        return out

    def read_image_wait(self):
        _check(lib.gpurt_pipe_read_image_wait(self.h))

    def read_gbuffer(self, which, out=None):
        if out is None:
            out = np.zeros((self.h_px, self.w, 4), np.float32)
        p, m, _k = _ptr(out)
        _check(lib.gpurt_pipe_read_gbuffer(self.h, which, p, m))
        return out

    def ray_counts(self):
        c = (C.c_uint64 * 2)()
        _check(lib.gpurt_pipe_ray_counts(self.h, c))
        return int(c[0]), int(c[1])

    def last_uniforms(self):
        """(consts words, ubo words, seed word) of the last frame, as uint32 arrays"""
        c, u, s = Constants(), Camera(), C.c_uint32()
        _check(lib.gpurt_pipe_last_uniforms(self.h, C.byref(c), C.byref(u), C.byref(s)))
        return (np.frombuffer(bytes(c), np.uint32).copy(), np.frombuffer(bytes(u), np.uint32).copy(), s.value)

    def read_reservoirs(self):
        out = np.zeros((self.h_px * self.w, 12), np.uint32)
        _check(lib.gpurt_pipe_read_reservoirs(self.h, C.c_void_p(out.ctypes.data), MEM_HOST))
        return out

    def bounce_rays(self, bounce):
        """torch view (n,8) of the device ray queue traced at `bounce` in the last frame"""
        import torch
        ptr, n = C.c_void_p(), C.c_uint32()
        _check(lib.gpurt_pipe_bounce_rays(self.h, bounce, C.byref(ptr), C.byref(n)))
        if n.value == 0:
            return torch.empty((0, 8), dtype=torch.float32, device=f"cuda:{self.ctx.device}")

        class _Arr:  # __cuda_array_interface__ wrapper, memory stays owned by the pipe
            pass

        a = _Arr()
        a.__cuda_array_interface__ = {"shape": (n.value, 8), "typestr": "<f4", "data": (ptr.value, False), "version": 2}
        return torch.as_tensor(a, device=f"cuda:{self.ctx.device}")

    def time_ms(self):
        return self.ctx.last_kernel_ms()

    def tonemap(self, op=0, exposure=1.0, gamma=2.2):
        out = np.zeros((self.h_px, self.w, 4), np.uint8)
        _check(lib.gpurt_tonemap(self.h, op, C.c_float(exposure), C.c_float(gamma), C.c_void_p(out.ctypes.data), MEM_HOST))
        return out

    def close(self):
        if self.h:
            lib.gpurt_pipe_destroy(self.h)
            self.h = C.c_void_p()

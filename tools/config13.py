#!/usr/bin/env python
"""BASELINE configs 1 and 3 on one B200 — measurement only (parity for the same scenes is
in tests/test_gpu_parity.py and tests/test_gpu_render.py).

config 1  cbox.gltf: 1,048,576 random rays (tea(i, 0xC0FFEE), origins in the world box inflated 10 %,
          uniform directions), 1024x1024 coherent primary rays from the Camera::reset() defaults,
          1,048,576 closest-point queries (tea(i, 0xFACADE), box inflated 25 %).
config 3  mis_test.gltf at 1920x1080, depth 4, 1 spp: integrator 2 (MIS) and integrators 3 / 4
          (ReSTIR direct / ReSTIR, res_samples 4, temporal reuse, temporal_scale 16) over 8 frames.

Kernel times are CUDA events (gpurt_last_kernel_ms), 3 warm-ups, median of 10, L2 flushed in between.
Prints one JSON object.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gpurt  # noqa: E402
from config4_cpq import randf_t, tea_t  # noqa: E402

MEDIA = os.path.join(ROOT, "tests", "data", "media")
N = 1 << 20


def stream(n, key, draws, dev):
    i = torch.arange(n, dtype=torch.int64, device=dev)
    s = tea_t(i, torch.full_like(i, key))
    out = []
    for _ in range(draws):
        f, s = randf_t(s)
        out.append(f)
    return out


def random_rays(lo, hi, dev):
    x = stream(N, 0xC0FFEE, 5, dev)
    ext = hi - lo
    rays = torch.empty((N, 8), dtype=torch.float32, device=dev)
    for k in range(3):
        rays[:, k] = (lo[k] - 0.1 * ext[k]) + x[k] * (1.2 * ext[k])
    z = 1.0 - 2.0 * x[3]
    r = torch.sqrt(torch.clamp(1.0 - z * z, min=0.0))
    phi = 2.0 * np.pi * x[4]
    rays[:, 4], rays[:, 5], rays[:, 6] = r * torch.cos(phi), r * torch.sin(phi), z
    rays[:, 3], rays[:, 7] = 1e-5, 1e7
    return rays


def random_points(lo, hi, dev):
    x = stream(N, 0xFACADE, 3, dev)
    ext = hi - lo
    q = torch.empty((N, 4), dtype=torch.float32, device=dev)
    for k in range(3):
        q[:, k] = (lo[k] - 0.25 * ext[k]) + x[k] * (1.5 * ext[k])
    q[:, 3] = float("inf")
    return q


def median_ms(fn, ctx, flush, reps=10, warm=3):
    ms = []
    for it in range(warm + reps):
        flush.zero_()
        fn()
        t = ctx.last_kernel_ms()
        if it >= warm:
            ms.append(t)
    return float(np.median(ms)), float(min(ms))


def run(ctx=None, flush=None, device=0):
    """both configs on `ctx` (a fresh context on `device` if None); returns the result dict"""
    dev = torch.device("cuda", device)
    ctx = ctx or gpurt.Context(device)
    ctx.use_torch_stream()
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}

    # ---- config 1 ------------------------------------------------------------------------------
    scene = gpurt.Scene(ctx).load(os.path.join(MEDIA, "cbox", "cbox.gltf"))
    gpurt.Accel(scene).close()
    accel = gpurt.Accel(scene)
    info = accel.info()
    lo, hi = np.array(list(info.scene_min)), np.array(list(info.scene_max))
    rays, pts = random_rays(lo, hi, dev), random_points(lo, hi, dev)
    hits = torch.empty((N, 4), dtype=torch.float32, device=dev)
    cps = torch.empty((N, 8), dtype=torch.float32, device=dev)
    occ = torch.empty(N, dtype=torch.uint8, device=dev)
    pipe = gpurt.RTPipe(scene, accel)
    prm = gpurt.pipe_params(integrator=1, brdf=1, max_depth=1, samples_per_frame=1, max_frames=1, use_rr=0, env_scale=1.0)
    pipe.render_frame(prm, gpurt.camera(0, 1024, 1024), 1024, 1024)
    prim = pipe.bounce_rays(0).clone()
    hp = torch.empty((prim.shape[0], 4), dtype=torch.float32, device=dev)
    st = accel.trace_stats(rays, hits)
    c1 = {"tris": info.n_tris, "objs": info.n_objs, "wide_nodes": info.n_wide_nodes, "bvh_build_ms": info.build_ms,
          "nodes_per_ray_random": st.nodes_visited / st.rays, "tris_per_ray_random": st.tris_tested / st.rays,
          "hit_fraction_random": st.hits / st.rays}
    for name, fn, n in (("random_closest", lambda: accel.trace_closest(rays, hits), N),
                        ("random_any", lambda: accel.trace_any(rays, occ), N),
                        ("coherent_1024x1024_closest", lambda: accel.trace_closest(prim, hp), prim.shape[0]),
                        ("closest_points", lambda: accel.closest_points(pts, cps), N)):
        med, best = median_ms(fn, ctx, flush)
        c1[name] = {"n": n, "ms_median": med, "ms_best": best, "m_per_s": n / (med * 1e-3) / 1e6}
    out["config1_cbox"] = c1
    pipe.close(), accel.close(), scene.close()

    # ---- config 3 ------------------------------------------------------------------------------
    scene = gpurt.Scene(ctx).load(os.path.join(MEDIA, "mis_test", "mis_test.gltf"))
    accel = gpurt.Accel(scene)
    pipe = gpurt.RTPipe(scene, accel)
    W, H = 1920, 1080
    cam = gpurt.camera(1, W, H, (0.5, 0.6, 2.6), (0.5, 0.45, 0.0), 50.0)
    c3 = {"tris": accel.info().n_tris, "lights": scene.counts()["lights"], "camera": "pos (0.5,0.6,2.6) -> (0.5,0.45,0), vfov 50"}
    for integ, name, frames in ((2, "mis", 1), (3, "restir_direct", 8), (4, "restir", 8)):
        prm = gpurt.pipe_params(integrator=integ, brdf=1, max_depth=4, samples_per_frame=1, max_frames=frames,
                                use_rr=1, use_temporal=1, temporal_scale=16, res_samples=4, seed=3)
        per_frame = []
        for rep in range(3):              # first repetition is the warm-up
            pipe.reset_frame()
            ms = []
            while pipe.render_frame(prm, cam, W, H) == 0:
                ms.append(pipe.time_ms())
            per_frame = ms
        closest, anyr = pipe.ray_counts()
        c3[name] = {"frames": len(per_frame), "ms_per_frame_median": float(np.median(per_frame)),
                    "ms_first_frame": per_frame[0], "mpaths_s": W * H / (float(np.median(per_frame)) * 1e-3) / 1e6,
                    "closest_rays_last_frame": closest, "any_rays_last_frame": anyr}
    out["config3_mis_test_1080p"] = c3
    pipe.close(), accel.close(), scene.close()
    return out


def main():
    print(json.dumps(run()))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Quick device-time probe of the traversal kernels on the bench workload (development aid; the
numbers that count come from bench.py).  Prints Mrays/s for primary / bounce / mixed ray sets and
M closest-point queries/s, plus nodes and triangles visited per ray."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gpurt  # noqa: E402


def main():
    ctx = gpurt.Context(0)
    scene, label = bench.build_scene(gpurt, ctx)
    accel = gpurt.Accel(scene)
    info = accel.info()
    W, H = bench.W, bench.H
    cam = gpurt.camera(1, W, H, bench.CAM_POS, bench.CAM_AT, bench.VFOV)
    pipe = gpurt.RTPipe(scene, accel)
    prm = gpurt.pipe_params(integrator=1, brdf=1, max_depth=2, samples_per_frame=1, max_frames=1, use_rr=0, env_scale=1.0)
    ctx.use_torch_stream()
    fr = []
    for _ in range(5):
        pipe.reset_frame()
        pipe.render_frame(prm, cam, W, H)
        fr.append(pipe.time_ms())
    prim, bnc = pipe.bounce_rays(0).clone(), pipe.bounce_rays(1).clone()
    mixed = torch.cat([prim, bnc])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    print(f"{label}: {info.n_tris} tris, {info.n_wide_nodes} nodes, depth {info.wide_depth}, build {info.build_ms:.2f} ms, frame {np.median(fr):.3f} ms")
    for name, rays in (("primary", prim), ("bounce", bnc), ("mixed", mixed)):
        hits = torch.empty((rays.shape[0], 4), dtype=torch.float32, device="cuda")
        ms = []
        for _ in range(30):
            flush.zero_()
            accel.trace_closest(rays, hits)
            ms.append(ctx.last_kernel_ms())
        st = accel.trace_stats(rays, hits)
        t = float(np.median(ms[5:]))
        bpr = 48 + 80 * st.nodes_visited / st.rays + 48 * st.tris_tested / st.rays
        print(f"  {name:8s} {rays.shape[0]:8d} rays  {t:7.3f} ms  {rays.shape[0] / t / 1e3:8.1f} Mrays/s  nodes/ray {st.nodes_visited / st.rays:5.2f}"
              f"  tris/ray {st.tris_tested / st.rays:5.2f}  logical {bpr * rays.shape[0] / t / 1e6:7.1f} GB/s")
    hp = accel.trace_closest(prim).cpu().numpy().view(gpurt.HIT_DT).reshape(-1)
    p = prim.cpu().numpy()
    q = np.zeros((p.shape[0], 4), np.float32)
    tt = np.where(np.isfinite(hp["t"]), hp["t"], 100.0).astype(np.float32)
    jit = (bench.lcg_randf(bench.tea(np.arange(p.shape[0], dtype=np.uint32), np.uint32(0xD00D)))[:, None] - 0.5) * 60.0
    q[:, :3] = p[:, 0:3] + 0.8 * tt[:, None] * p[:, 4:7] + jit
    q[:, 3] = np.inf
    dq = torch.from_numpy(q).cuda()
    out = accel.closest_points(dq)
    ms = []
    for _ in range(20):
        flush.zero_()
        accel.closest_points(dq, out)
        ms.append(ctx.last_kernel_ms())
    t = float(np.median(ms[3:]))
    print(f"  cpq      {q.shape[0]:8d} qrys  {t:7.3f} ms  {q.shape[0] / t / 1e3:8.1f} Mq/s")


if __name__ == "__main__":
    main()

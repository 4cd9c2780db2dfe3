#!/usr/bin/env python
"""Experiment: how much faster would the bounce-ray queue of the config-2 frame trace if it were reordered
(sort cost NOT included)?  Keys: direction octant; octant + coarse origin cell; finer direction bins."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gpurt  # noqa: E402


def timed(accel, ctx, rays, flush, reps=20):
    hits = torch.empty((rays.shape[0], 4), dtype=torch.float32, device="cuda")
    ms = []
    for _ in range(reps):
        flush.zero_()
        accel.trace_closest(rays, hits)
        ms.append(ctx.last_kernel_ms())
    return float(np.median(ms[4:]))


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
    pipe.render_frame(prm, cam, W, H)
    bnc = pipe.bounce_rays(1).clone()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    n = bnc.shape[0]
    lo = torch.tensor(list(info.scene_min), device="cuda")
    ext = torch.tensor(list(info.scene_max), device="cuda") - lo
    d = bnc[:, 4:7]
    octant = ((d[:, 0] < 0).long() | ((d[:, 1] < 0).long() << 1) | ((d[:, 2] < 0).long() << 2))
    idx = torch.arange(n, device="cuda")

    def cell(bits):
        q = ((bnc[:, 0:3] - lo) / ext * (1 << bits)).long().clamp(0, (1 << bits) - 1)
        key = torch.zeros(n, dtype=torch.long, device="cuda")
        for b in range(bits):
            for a in range(3):
                key |= ((q[:, a] >> b) & 1) << (3 * b + a)
        return key

    # finer direction bins: dominant axis (6) x 4x4 on the other two components
    ad = d.abs()
    ax = ad.argmax(1)
    sign = (d.gather(1, ax[:, None])[:, 0] < 0).long()
    o1 = d.gather(1, ((ax + 1) % 3)[:, None])[:, 0] / ad.gather(1, ax[:, None])[:, 0]
    o2 = d.gather(1, ((ax + 2) % 3)[:, None])[:, 0] / ad.gather(1, ax[:, None])[:, 0]
    fine = (ax * 2 + sign) * 16 + ((o1 * 0.5 + 0.5) * 4).long().clamp(0, 3) * 4 + ((o2 * 0.5 + 0.5) * 4).long().clamp(0, 3)
    block = idx // 4096
    keys = {
        "as produced": idx,
        "octant (global)": octant * n + idx,
        "octant within 4096-ray blocks": block * 8 * 4096 + octant * 4096 + idx % 4096,
        "octant within 32768-ray blocks": (idx // 32768) * 8 * 32768 + octant * 32768 + idx % 32768,
        "cell 3 bits + octant": (cell(3) * 8 + octant) * n + idx,
        "cell 5 bits + octant": (cell(5) * 8 + octant) * n + idx,
        "octant + cell 4 bits": (octant * (1 << 12) + cell(4)) * n + idx,
        "96 direction bins": fine * n + idx,
        "96 direction bins + cell 3 bits": (fine * 512 + cell(3)) * n + idx,
        "random shuffle": torch.randperm(n, device="cuda"),
    }
    base = None
    for name, k in keys.items():
        rays = bnc[torch.argsort(k)].contiguous()
        t = timed(accel, ctx, rays, flush)
        base = base or t
        print(f"  {name:36s} {t:7.3f} ms  {n / t / 1e3:8.1f} Mrays/s  x{base / t:5.2f}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Experiment: 2^21 uniform random rays on the stand-in through gpurt_trace_closest, as is and (GPURT_ORDER_MIN_BVH_BYTES=0)
through the library's Morton-ordered path; kernel time includes probe + keys + sort."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gpurt  # noqa: E402

ctx = gpurt.Context(0)
scene, label = bench.build_scene(gpurt, ctx)
accel = gpurt.Accel(scene)
ctx.use_torch_stream()
info = accel.info()
lo = torch.tensor(list(info.scene_min), device="cuda")
hi = torch.tensor(list(info.scene_max), device="cuda")
n = 1 << 21
g = torch.Generator(device="cuda").manual_seed(7)
rays = torch.empty((n, 8), dtype=torch.float32, device="cuda")
rays[:, 0:3] = lo + (hi - lo) * torch.rand((n, 3), generator=g, device="cuda")
z = 1 - 2 * torch.rand(n, generator=g, device="cuda")
ph = 2 * np.pi * torch.rand(n, generator=g, device="cuda")
r = torch.sqrt(torch.clamp(1 - z * z, min=0))
rays[:, 4], rays[:, 5], rays[:, 6] = r * torch.cos(ph), r * torch.sin(ph), z
rays[:, 3], rays[:, 7] = 1e-5, 1e7
hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ms = []
for _ in range(12):
    flush.zero_()
    accel.trace_closest(rays, hits)
    ms.append(ctx.last_kernel_ms())
t = float(np.median(ms[3:]))
print(f"random rays on {label}: env {os.environ.get('GPURT_ORDER_MIN_BVH_BYTES', '-')}  {t:.3f} ms  {n / t / 1e3:.1f} Mrays/s  checksum {int(hits.view(torch.int32).sum().item())}")

#!/usr/bin/env python
"""Experiment: closest-hit throughput on the 10 M-triangle config-4 soup for random rays in generated order vs
sorted by the Morton code of the origin (+ direction octant); sort cost not included."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gpurt  # noqa: E402
from config4_cpq import make_soup, randf_t, tea_t  # noqa: E402
from cpq_sort_probe import morton  # noqa: E402

n_tris, n = 10_000_000, 10_000_000
dev = torch.device("cuda", 0)
ctx = gpurt.Context(0)
ctx.use_torch_stream()
scene = gpurt.Scene(ctx)
scene.add_triangles(make_soup(n_tris, dev).cpu().numpy())
accel = gpurt.Accel(scene)
i = torch.arange(n, dtype=torch.int64, device=dev)
s = tea_t(i, torch.full_like(i, 0xC0FFEE))
x = []
for _ in range(5):
    f, s = randf_t(s)
    x.append(f)
rays = torch.empty((n, 8), dtype=torch.float32, device=dev)
for k in range(3):
    rays[:, k] = x[k] * 1.2 - 0.1
z = 1 - 2 * x[3]
r = torch.sqrt(torch.clamp(1 - z * z, min=0))
rays[:, 4], rays[:, 5], rays[:, 6] = r * torch.cos(2 * np.pi * x[4]), r * torch.sin(2 * np.pi * x[4]), z
rays[:, 3], rays[:, 7] = 1e-5, 1e7
hits = torch.empty((n, 4), dtype=torch.float32, device=dev)
q = torch.cat([(rays[:, :3] + 0.1) / 1.2 * 1.5 - 0.25, rays[:, 3:4]], 1)   # map into morton()'s [-0.25,1.25] convention
key = morton(q)
d = rays[:, 4:7]
octant = ((d[:, 0] < 0).long() | ((d[:, 1] < 0).long() << 1) | ((d[:, 2] < 0).long() << 2))
for name, order in (("generated order", None), ("origin morton 30 bit", torch.argsort(key)), ("origin morton 15 bit + octant", torch.argsort((key >> 15) * 8 + octant, stable=True))):
    rr = rays if order is None else rays[order].contiguous()
    ms = []
    for _ in range(4):
        accel.trace_closest(rr, hits)
        ms.append(ctx.last_kernel_ms())
    t = float(np.median(ms[1:]))
    st = accel.trace_stats(rr, hits)
    print(f"{name:32s} {t:8.3f} ms {n / t / 1e3:8.1f} Mrays/s  nodes/ray {st.nodes_visited / st.rays:.1f} tris/ray {st.tris_tested / st.rays:.1f}")

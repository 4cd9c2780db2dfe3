#!/usr/bin/env python
"""Experiment: closest-point throughput on the config-4 soup for queries in generated (random) order vs sorted by
the Morton code of their position (sort cost not included)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gpurt  # noqa: E402
from config4_cpq import make_queries, make_soup  # noqa: E402


def morton(q, bits=10):
    x = ((q[:, :3] + 0.25) / 1.5 * (1 << bits)).long().clamp(0, (1 << bits) - 1)
    key = torch.zeros(q.shape[0], dtype=torch.long, device=q.device)
    for b in range(bits):
        for a in range(3):
            key |= ((x[:, a] >> b) & 1) << (3 * b + a)
    return key


def main():
    n_tris = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    n_q = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
    dev = torch.device("cuda", 0)
    ctx = gpurt.Context(0)
    ctx.use_torch_stream()
    scene = gpurt.Scene(ctx)
    scene.add_triangles(make_soup(n_tris, dev).cpu().numpy())
    accel = gpurt.Accel(scene)
    q = make_queries(0, n_q, dev)
    out = torch.empty((n_q, 8), dtype=torch.float32, device=dev)
    for name, qq in (("generated order", q), ("morton 30-bit sorted", q[torch.argsort(morton(q))].contiguous()),
                     ("morton 12-bit sorted (4 bits/axis)", q[torch.argsort(morton(q, 4), stable=True)].contiguous())):
        ms = []
        for _ in range(4):
            accel.closest_points(qq, out)
            ms.append(ctx.last_kernel_ms())
        t = float(np.median(ms[1:]))
        print(f"{name:36s} {t:8.3f} ms {n_q / t / 1e3:8.1f} Mq/s")


if __name__ == "__main__":
    main()

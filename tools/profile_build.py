#!/usr/bin/env python
"""Build the config-4 soup (default 10 M triangles) three times and print the device build time; run under
`ncu --metrics gpu__time_duration.sum` to get the per-kernel split of the LBVH pipeline."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gpurt  # noqa: E402
from config4_cpq import make_soup  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tris", type=int, default=10_000_000)
ap.add_argument("--builds", type=int, default=3)
args = ap.parse_args()
dev = torch.device("cuda", 0)
ctx = gpurt.Context(0)
tris = make_soup(args.tris, dev).cpu().numpy()
scene = gpurt.Scene(ctx)
scene.add_triangles(tris)
accel = gpurt.Accel(scene)
ms = [accel.info().build_ms]
for _ in range(args.builds - 1):
    accel.update()
    ms.append(accel.info().build_ms)
info = accel.info()
print(json.dumps({"tris": info.n_tris, "build_ms": ms, "mtris_s": info.n_tris / (min(ms) * 1e-3) / 1e6,
                  "wide_nodes": info.n_wide_nodes, "wide_depth": info.wide_depth}))

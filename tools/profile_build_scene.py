#!/usr/bin/env python
"""Build a named scene (sponza_standin | cbox | mis_test | irregular) a few times and print the device build times; run under
`ncu --metrics gpu__time_duration.sum` + tools/launch_split.py for the per-kernel split.  GPURT_BUILD_FLAGS selects the build."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpurt  # noqa: E402
from scenes import load_scene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="sponza_standin")
ap.add_argument("--builds", type=int, default=4)
args = ap.parse_args()
ctx = gpurt.Context(0)
scene = load_scene(gpurt, ctx, args.scene)
gpurt.Accel(scene).close()
accel = gpurt.Accel(scene)
ms = [accel.info().build_ms]
for _ in range(args.builds - 1):
    accel.update()
    ms.append(accel.info().build_ms)
info = accel.info()
print(json.dumps({"scene": args.scene, "flags": os.environ.get("GPURT_BUILD_FLAGS", "0"), "tris": info.n_tris, "build_ms": ms,
                  "wide_nodes": info.n_wide_nodes, "wide_depth": info.wide_depth}))

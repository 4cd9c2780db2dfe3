#!/usr/bin/env python
"""Render a few 1080p frames of mis_test (BASELINE config 3) with one integrator and print the per-frame times —
the small driver behind the ncu captures and A/B runs of the shading kernels.
usage: python tools/mis_frame.py [integrator=2] [frames=6] [max_depth=4]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
import gpurt  # noqa: E402

integ = int(sys.argv[1]) if len(sys.argv) > 1 else 2
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 6
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ctx = gpurt.Context(0)
scene = gpurt.Scene(ctx).load(os.path.join(ROOT, "tests", "data", "media", "mis_test", "mis_test.gltf"))
accel = gpurt.Accel(scene)
pipe = gpurt.RTPipe(scene, accel)
W, H = 1920, 1080
cam = gpurt.camera(1, W, H, (0.5, 0.6, 2.6), (0.5, 0.45, 0.0), 50.0)
prm = gpurt.pipe_params(integrator=integ, brdf=1, max_depth=depth, samples_per_frame=1, max_frames=frames, use_rr=1,
                        use_temporal=1, temporal_scale=16, res_samples=4, seed=3)
ms = []
while pipe.render_frame(prm, cam, W, H) == 0:
    ms.append(pipe.time_ms())
print(json.dumps({"integrator": integ, "max_depth": depth, "ms": ms, "median": float(np.median(ms[1:])),
                  "rays": pipe.ray_counts(), "light_groups": os.environ.get("GPURT_LIGHT_GROUPS", "1")}))

#!/usr/bin/env python
"""The reference's own default workload (1280x720, 8 spp per frame, depth 8, integrator 0 = Direct,
Blinn-Phong, src/platform/window.cpp:36-38, src/vk/rt.h:38-53) on media/cbox: per-frame device time.
usage: python tools/default_workload.py [frames=12] [integrator=0] [scene=cbox]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpurt  # noqa: E402
from scenes import load_scene  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 12
integ = int(sys.argv[2]) if len(sys.argv) > 2 else 0
name = sys.argv[3] if len(sys.argv) > 3 else "cbox"
ctx = gpurt.Context(0)
scene = load_scene(gpurt, ctx, name)
accel = gpurt.Accel(scene)
pipe = gpurt.RTPipe(scene, accel)
W, H = 1280, 720
cam = gpurt.camera(0, W, H) if name == "cbox" else gpurt.camera(1, W, H, (0.5, 0.6, 2.6), (0.5, 0.45, 0.0), 50.0)
prm = gpurt.pipe_params(integrator=integ, max_frames=frames)      # everything else at the reference's defaults
ms = []
while pipe.render_frame(prm, cam, W, H) == 0:
    ms.append(pipe.time_ms())
c = pipe.ray_counts()
med = float(np.median(ms[2:]))
print(json.dumps({"scene": name, "integrator": integ, "size": [W, H], "spp_per_frame": prm.samples_per_frame, "ms_per_frame": med,
                  "mpaths_s": W * H * prm.samples_per_frame / (med * 1e-3) / 1e6, "closest_rays_per_frame": c[0], "any_rays_per_frame": c[1],
                  "mrays_s": (c[0] + c[1]) / (med * 1e-3) / 1e6}))

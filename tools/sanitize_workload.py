#!/usr/bin/env python
"""Small pass over the round-2 device code for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): SAH builds
across both regimes, refit, update_auto, queries through pageable and pinned host arrays, the shadow-queue stage."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["GPURT_SHADOW_QUEUE"] = "1"
import gpurt  # noqa: E402
from scenes import load_scene, soup  # noqa: E402

ctx = gpurt.Context(0)
rng = np.random.default_rng(1)
for n in (2, 33, 1500, 6000):
    sc = gpurt.Scene(ctx)
    sc.add_triangles(soup(n, seed=n, ext=0.05))
    a = gpurt.Accel(sc)
    rays = np.zeros((2000, 8), np.float32)
    rays[:, :3] = rng.random((2000, 3)) * 1.2 - 0.1
    d = rng.standard_normal((2000, 3)).astype(np.float32)
    rays[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 3], rays[:, 7] = 1e-5, 1e7
    a.trace_closest(rays)
    a.trace_any(rays)
    q = np.zeros((2000, 4), np.float32)
    q[:, :3], q[:, 3] = rng.random((2000, 3)), np.inf
    a.closest_points(q)
    pr = torch.from_numpy(rays).pin_memory()
    ph = torch.empty((2000, 4), dtype=torch.float32).pin_memory()
    a.trace_closest(pr.numpy(), ph.numpy())
    a.close(), sc.close()
scene = load_scene(gpurt, ctx, "cbox")
accel = gpurt.Accel(scene)
m = np.array(list(scene.descs()[3].model), np.float32)
m[12] += 0.1
scene.set_transform(3, m)
accel.refit()
accel.update_auto()
pipe = gpurt.RTPipe(scene, accel)
for integ in (0, 2, 3):
    pipe.reset_frame()
    pipe.render_frame(gpurt.pipe_params(integrator=integ, samples_per_frame=1, max_depth=3), gpurt.camera(0, 96, 54), 96, 54)
    pipe.read_image()
print("sanitize workload done", accel.info().n_wide_nodes)

#!/usr/bin/env python
"""The three query launches bench.py reports rooflines for, for one `ncu --set full` capture: the config-2 frame's ray set
(primary + bounce), its primary rays alone, and the closest-point queries near the primary hit points — two launches each
(the first warms the caches), on the headline scene with the default build.  Prints the element counts."""
import json
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
W, H = bench.W, bench.H
cam = gpurt.camera(1, W, H, bench.CAM_POS, bench.CAM_AT, bench.VFOV)
pipe = gpurt.RTPipe(scene, accel)
prm = gpurt.pipe_params(integrator=1, brdf=1, max_depth=2, samples_per_frame=1, max_frames=1, use_rr=0, env_scale=1.0, seed=0)
ctx.use_torch_stream()
pipe.render_frame(prm, cam, W, H)
rays = torch.cat([pipe.bounce_rays(0), pipe.bounce_rays(1)]).clone()
n, n_p = rays.shape[0], W * H
hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(2):
    flush.zero_()
    accel.trace_closest(rays, hits)
for _ in range(2):
    flush.zero_()
    accel.trace_closest(rays[:n_p], hits[:n_p])
prim = rays[:n_p].cpu().numpy()
hp = hits[:n_p].cpu().numpy().view(gpurt.HIT_DT).reshape(-1)
tt = np.where(np.isfinite(hp["t"]), hp["t"], 100.0).astype(np.float32)
jit = (bench.lcg_randf(bench.tea(np.arange(n_p, dtype=np.uint32), np.uint32(0xD00D)))[:, None] - 0.5) * 60.0
q = np.zeros((n_p, 4), np.float32)
q[:, :3] = prim[:, 0:3] + 0.8 * tt[:, None] * prim[:, 4:7] + jit
q[:, 3] = np.inf
d_q = torch.from_numpy(q).cuda()
cp = accel.closest_points(d_q)
for _ in range(2):
    flush.zero_()
    accel.closest_points(d_q, cp)
torch.cuda.synchronize()
print(json.dumps({"scene": label, "rays_all": n, "rays_primary": n_p, "queries": n_p}))

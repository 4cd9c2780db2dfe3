#!/usr/bin/env python
"""Small pass over the round-2 device code for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): SAH builds
across both regimes, refit, update_auto, queries through pageable and pinned host arrays, the shadow-queue stage, the ReSTIR
extensions, a frame sharded over two pipes with the history exchange, the gather inbox, sliced placement."""
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
# asynchronous read-back (copy stream + event) overlapping the next frames
host_img = torch.empty((54, 96, 4), dtype=torch.float32).pin_memory()
for _ in range(3):
    pipe.render_frame(gpurt.pipe_params(integrator=1, samples_per_frame=1, max_depth=3), gpurt.camera(0, 96, 54), 96, 54)
    pipe.read_image_async(host_img.numpy())
pipe.read_image_wait()
assert np.array_equal(host_img.numpy(), pipe.read_image())
pipe.close()
# round 2, late: ReSTIR extensions, the frame sharded over two pipes with the history exchange, the gather inbox and the sliced
# placement of ordered batches (thresholds lowered by the test hooks so that a 6000-triangle scene takes the ordered paths)
for kw in (dict(integrator=3, spatial_samples=3, spatial_radius=5.0), dict(integrator=2, light_sampling=1), dict(integrator=4, light_sampling=1)):
    pipe = gpurt.RTPipe(scene, accel)
    for _ in range(3):
        pipe.render_frame(gpurt.pipe_params(samples_per_frame=1, max_depth=3, **kw), gpurt.camera(0, 96, 54), 96, 54)
    pipe.read_image()
    pipe.close()
pipes = [gpurt.RTPipe(scene, accel) for _ in range(2)]
for s_, p_ in enumerate(pipes):
    p_.set_shard(16, 2, s_)
blocks = [p_.history_export(96, 54)[0] for p_ in pipes]
for p_ in pipes:
    p_.history_peers(blocks, 8)
for _ in range(3):
    for p_ in pipes:
        p_.render_frame(gpurt.pipe_params(integrator=3, samples_per_frame=1, max_depth=2), gpurt.camera(0, 96, 54), 96, 54)
assert all(p_.history_status()[1] == 0 for p_ in pipes)
for p_ in pipes:
    p_.close()
os.environ["GPURT_ORDER_MIN_BATCH"], os.environ["GPURT_ORDER_MIN_BVH_BYTES"] = "1000", "1000"
if "--no-gather" in sys.argv:   # the sanitizer serialises kernels: receivers spinning on flags would only time out
    os.environ["GPURT_PLACE_FORCE"], os.environ["GPURT_PLACE_SLICES"] = "1", "3"
    sc_ = gpurt.Scene(ctx)
    sc_.add_triangles(soup(6000, seed=3, ext=0.05))
    a_ = gpurt.Accel(sc_)
    q = torch.rand((5000, 4), device="cuda")
    q[:, 3] = float("inf")
    a_.closest_points(q)
    torch.cuda.synchronize()
    print("sanitize workload done (without the gather inbox)", accel.info().n_wide_nodes)
    sys.exit(0)
ctx2 = gpurt.Context(0)
tris = soup(6000, seed=3, ext=0.05)
pair = []
for c_ in (ctx, ctx2):
    sc_ = gpurt.Scene(c_)
    sc_.add_triangles(tris)
    pair.append((sc_, gpurt.Accel(sc_)))
n0, n1 = 3000, 5000
q = torch.rand((n0 + n1, 4), device="cuda")
q[:, 3] = float("inf")
want = torch.cat([pair[0][1].closest_points(q[:n0]), pair[1][1].closest_points(q[n0:])]).clone()
torch.cuda.synchronize()
g0 = gpurt.Gather.create(ctx, n0 + n1, 32, [0, n0, n0 + n1])
g1 = gpurt.Gather.open(ctx2, n0 + n1, 32, [0, n0, n0 + n1], 1, base=g0.base())
for _ in range(2):
    g0.begin()
    pair[1][1].closest_points(q[n0:], g1.mine())
    pair[0][1].closest_points(q[:n0], g0.mine())
    assert g0.end(sync=True) == 0
assert torch.equal(g0.tensor().view(torch.int32).view(-1, 8), want.view(torch.int32))
os.environ["GPURT_PLACE_FORCE"], os.environ["GPURT_PLACE_SLICES"] = "1", "3"
got = pair[0][1].closest_points(q[:n0])
torch.cuda.synchronize()
assert torch.equal(got.view(torch.int32), want[:n0].view(torch.int32))
g1.close(), g0.close()
print("sanitize workload done", accel.info().n_wide_nodes)

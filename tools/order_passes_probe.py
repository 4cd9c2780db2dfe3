#!/usr/bin/env python
"""Experiment: how many 8-bit radix passes the Morton ordering of a large batch needs (order.cu, GPURT_ORDER_PASSES = 4 / 3 /
2: the sort looks at the top 32 / 24 / 16 bits of the 30-bit key).  Fewer passes = a cheaper sort and a coarser order.
Whole library calls timed with events (probe + keys + sort + traversal through the index), variants interleaved in ONE
process (the library reads the variable at every call).  Stand-in scene: the bench's closest-point queries (J = 60) and a
scattered batch (J = 400); then config 4 (10 M-triangle soup, 100 M points in one call) unless --no-config4."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gpurt  # noqa: E402

PASSES = ("4", "3", "2")


def timed(fn, reps, flush):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms[2:]))


def main():
    ctx = gpurt.Context(0)
    ctx.use_torch_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    scene, label = bench.build_scene(gpurt, ctx)
    accel = gpurt.Accel(scene)
    W, H = bench.W, bench.H
    cam = gpurt.camera(1, W, H, bench.CAM_POS, bench.CAM_AT, bench.VFOV)
    pipe = gpurt.RTPipe(scene, accel)
    pipe.render_frame(gpurt.pipe_params(integrator=1, brdf=1, max_depth=2, samples_per_frame=1, max_frames=1, use_rr=0, env_scale=1.0), cam, W, H)
    prim = pipe.bounce_rays(0).clone()
    hp = accel.trace_closest(prim).cpu().numpy().view(gpurt.HIT_DT).reshape(-1)
    p = prim.cpu().numpy()
    n = p.shape[0]
    tt = np.where(np.isfinite(hp["t"]), hp["t"], 100.0).astype(np.float32)
    for jitter in (60.0, 400.0):
        jit = (bench.lcg_randf(bench.tea(np.arange(n, dtype=np.uint32), np.uint32(0xD00D)))[:, None] - 0.5) * jitter
        q = np.zeros((n, 4), np.float32)
        q[:, :3] = p[:, 0:3] + 0.8 * tt[:, None] * p[:, 4:7] + jit
        q[:, 3] = np.inf
        dq = torch.from_numpy(q).cuda()
        os.environ["GPURT_ORDER_PASSES"] = "4"
        ref = accel.closest_points(dq).clone()
        out = torch.empty((n, 8), dtype=torch.float32, device="cuda")
        for rep in range(2):
            for ps in PASSES:
                os.environ["GPURT_ORDER_PASSES"] = ps
                t = timed(lambda: accel.closest_points(dq, out), 12, flush)
                same = bool((out.view(torch.int32) == ref.view(torch.int32)).all())
                print(f"{label.split(' ')[0]} J {jitter:4.0f} passes {ps} rep{rep} {t:7.3f} ms {n / t / 1e3:8.1f} Mq/s same results: {same}", flush=True)
    pipe.close(), accel.close(), scene.close()
    if "--no-config4" in sys.argv:
        return
    from config4_cpq import make_queries, make_soup
    dev = torch.device("cuda:0")
    n_tris, nq = 10_000_000, 100_000_000
    scene = gpurt.Scene(ctx)
    scene.add_triangles(make_soup(n_tris, dev).cpu().numpy())
    accel = gpurt.Accel(scene)
    q = torch.empty((nq, 4), dtype=torch.float32, device=dev)
    for c0 in range(0, nq, 12_500_000):
        c1 = min(nq, c0 + 12_500_000)
        q[c0:c1] = make_queries(c0, c1, dev)
    out = torch.empty((nq, 8), dtype=torch.float32, device=dev)
    check = None
    for rep in range(2):
        for ps in PASSES:
            os.environ["GPURT_ORDER_PASSES"] = ps
            t = timed(lambda: accel.closest_points(q, out), 4, None)
            cs = int(out.view(torch.int32)[:, 4].to(torch.int64).sum().item())  # checksum of the primitive ids
            check = cs if check is None else check
            print(f"config4 one call of {nq} passes {ps} rep{rep} {t:8.3f} ms {nq / t / 1e3:8.1f} Mq/s same checksum: {cs == check}", flush=True)
        for ps in PASSES:
            os.environ["GPURT_ORDER_PASSES"] = ps
            t = timed(lambda: accel.closest_points(q[:12_500_000], out[:12_500_000]), 5, None)
            print(f"config4 one call of 12500000 passes {ps} rep{rep} {t:8.3f} ms {12_500_000 / t / 1e3:8.1f} Mq/s", flush=True)


if __name__ == "__main__":
    main()

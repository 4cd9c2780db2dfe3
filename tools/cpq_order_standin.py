#!/usr/bin/env python
"""Experiment: the bench's closest-point queries (near the primary hit points of the stand-in, +-30 units of jitter) as
generated (pixel-tile order) vs pre-sorted by the Morton code of the point, and through the library's own ordered path
(GPURT_ORDER_MIN_BVH_BYTES=0; GPURT_ORDER_PROBE_BITS picks the grid of the coherence probe)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gpurt  # noqa: E402


def morton(q, lo, hi, bits=10):
    x = ((q[:, :3] - lo) / (hi - lo) * (1 << bits)).long().clamp(0, (1 << bits) - 1)
    key = torch.zeros(q.shape[0], dtype=torch.long, device=q.device)
    for b in range(bits):
        for a in range(3):
            key |= ((x[:, a] >> b) & 1) << (3 * b + (2 - a))
    return key


def main():
    ctx = gpurt.Context(0)
    scene, label = bench.build_scene(gpurt, ctx)
    accel = gpurt.Accel(scene)
    W, H = bench.W, bench.H
    cam = gpurt.camera(1, W, H, bench.CAM_POS, bench.CAM_AT, bench.VFOV)
    pipe = gpurt.RTPipe(scene, accel)
    prm = gpurt.pipe_params(integrator=1, brdf=1, max_depth=2, samples_per_frame=1, max_frames=1, use_rr=0, env_scale=1.0)
    ctx.use_torch_stream()
    pipe.render_frame(prm, cam, W, H)
    prim = pipe.bounce_rays(0).clone()
    hp = accel.trace_closest(prim).cpu().numpy().view(gpurt.HIT_DT).reshape(-1)
    p = prim.cpu().numpy()
    n = p.shape[0]
    tt = np.where(np.isfinite(hp["t"]), hp["t"], 100.0).astype(np.float32)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    info = accel.info()
    lo = torch.tensor(list(info.scene_min), device="cuda")
    hi = torch.tensor(list(info.scene_max), device="cuda")
    for jitter in (60.0, 0.0, 400.0):
        jit = (bench.lcg_randf(bench.tea(np.arange(n, dtype=np.uint32), np.uint32(0xD00D)))[:, None] - 0.5) * jitter
        q = np.zeros((n, 4), np.float32)
        q[:, :3] = p[:, 0:3] + 0.8 * tt[:, None] * p[:, 4:7] + jit
        q[:, 3] = np.inf
        dq = torch.from_numpy(q).cuda()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        perm = torch.argsort(morton(dq, lo, hi))
        ev1.record()
        torch.cuda.synchronize()
        variants = [("as generated", dq, {}), ("pre-sorted (30-bit Morton, torch sort %.2f ms not counted)" % ev0.elapsed_time(ev1), dq[perm].contiguous(), {})]
        ref = accel.closest_points(dq)
        for name, qq, _ in variants:
            out = torch.empty((n, 8), dtype=torch.float32, device="cuda")
            ms = []
            for _ in range(12):
                flush.zero_()
                accel.closest_points(qq, out)
                ms.append(ctx.last_kernel_ms())
            t = float(np.median(ms[3:]))
            st = accel.closest_points_stats(qq)
            print(f"jitter {jitter:5.0f}  {name:70s} {t:7.3f} ms {n / t / 1e3:8.1f} Mq/s  nodes {st.nodes_visited / st.rays:5.2f} tris {st.tris_tested / st.rays:5.2f}", flush=True)
        # the library's own ordered path, whole call timed with events (key + probe + sort + kernel through the index)
        out = torch.empty((n, 8), dtype=torch.float32, device="cuda")
        ms = []
        for _ in range(12):
            flush.zero_()
            torch.cuda.synchronize()
            ev0.record()
            accel.closest_points(dq, out)
            ev1.record()
            torch.cuda.synchronize()
            ms.append(ev0.elapsed_time(ev1))
        t = float(np.median(ms[3:]))
        same = bool((out.view(torch.int32) == ref.view(torch.int32)).all())
        print(f"jitter {jitter:5.0f}  {'library call, events around it (env: ' + os.environ.get('GPURT_ORDER_MIN_BVH_BYTES', '-') + '/' + os.environ.get('GPURT_ORDER_PROBE_BITS', '-') + ')':70s} {t:7.3f} ms {n / t / 1e3:8.1f} Mq/s  same results: {same}", flush=True)


if __name__ == "__main__":
    main()

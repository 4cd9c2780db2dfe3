#!/usr/bin/env python
"""ReSTIR on a frame sharded over GPUs (BASELINE config 3's ReSTIR frames): every rank renders its row bands,
its frame-end stores the rows into the other ranks' previous-frame blocks over NVLink and raises a flag there, the next
frame's first kernel waits for the flags (include/gpurt.h gpurt_pipe_history_peers).  No host synchronisation per frame.

    python tools/restir_sharded.py --shards 3                        # one process, three pipes on one GPU (logic check)
    torchrun --nproc-per-node N tools/restir_sharded.py [--verify]   # one rank per GPU

Prints one JSON line on rank 0: ms per frame (CUDA events around the frame loop, max over ranks) sharded and unsharded,
bytes pushed per frame, and — with --verify — whether image, G-buffers and reservoirs equal the unsharded render bit for
bit after every frame.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
import gpurt  # noqa: E402
from gpurt.dist import share_history  # noqa: E402

MEDIA = os.path.join(ROOT, "tests", "data", "media")
ALL = gpurt.HISTORY_ALL_ROWS


def camera(args, f):
    """mis_test view of config 3; --orbit moves the eye a little every frame (reprojection crosses pixels and bands)"""
    w, h = args.size
    x = 0.5 + (0.03 * f if args.orbit else 0.0)
    y = 0.6 + (0.02 * f if args.orbit else 0.0)
    return gpurt.camera(1, w, h, (x, y, 2.6), (0.5, 0.45, 0.0), 50.0)


def params(args):
    return gpurt.pipe_params(integrator=args.integrator, brdf=1, samples_per_frame=1, max_depth=4, res_samples=4, use_temporal=1,
                             temporal_scale=16, seed=8, spatial_samples=args.spatial, spatial_radius=args.radius)


def same(a, b):
    a, b = np.ascontiguousarray(a).reshape(-1), np.ascontiguousarray(b).reshape(-1)
    return a.size == b.size and bool((a.view(np.uint32) == b.view(np.uint32)).all())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default=os.path.join(MEDIA, "mis_test", "mis_test.gltf"))
    ap.add_argument("--size", type=int, nargs=2, default=[1920, 1080])
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--integrator", type=int, default=3)
    ap.add_argument("--spatial", type=int, default=0)
    ap.add_argument("--radius", type=float, default=8.0)
    ap.add_argument("--bands", choices=["contiguous", "interleaved"], default="contiguous")
    ap.add_argument("--band-rows", type=int, default=0, help="rows per band, interleaved over the shards (overrides --bands)")
    ap.add_argument("--halo", default="all", help="rows, or 'all'")
    ap.add_argument("--orbit", action="store_true")
    ap.add_argument("--shards", type=int, default=0, help="single process: this many pipes on one GPU")
    ap.add_argument("--verify", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    w, h = args.size
    halo = ALL if args.halo == "all" else int(args.halo)
    torchrun = "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1
    rank, world, local = 0, 1, 0
    if torchrun:
        rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_shards = world if torchrun else max(1, args.shards)
    band = args.band_rows or (16 if args.bands == "interleaved" else (h + n_shards - 1) // n_shards)
    ctx = gpurt.Context(local)
    ctx.use_torch_stream()
    scene = gpurt.Scene(ctx).load(args.scene)
    accel = gpurt.Accel(scene)
    prm = params(args)
    rows = np.arange(h)
    owner = (rows // band) % n_shards

    # the shards this process renders
    mine = [rank] if torchrun else list(range(n_shards))
    pipes = {s: gpurt.RTPipe(scene, accel) for s in mine}
    for s, p in pipes.items():
        p.set_shard(band, n_shards, s)
    maps = []
    if torchrun:
        maps = share_history(pipes[rank], ctx, w, h, halo)
    elif n_shards > 1:
        blocks = {s: p.history_export(w, h)[0] for s, p in pipes.items()}
        for s, p in pipes.items():
            p.history_peers([blocks[t] for t in range(n_shards)], halo)

    ref = gpurt.RTPipe(scene, accel) if (args.verify and rank == 0) else None
    ok = {"image": True, "gbuffers": True, "reservoirs": True, "local_reservoirs": True}

    def gather_rows(arr_by_shard):
        """rows of the composite frame from their owners (host arrays; verification only)"""
        if torchrun:
            a = np.ascontiguousarray(arr_by_shard[rank]).reshape(h, -1)
            t = torch.from_numpy(a.view(np.uint8)).cuda()   # bytes: NCCL has no uint32
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            parts = [q.cpu().numpy().view(a.dtype) for q in parts]
        else:
            parts = [np.asarray(arr_by_shard[s]).reshape(h, -1) for s in range(n_shards)]
        return np.stack([parts[int(owner[y])][y] for y in range(h)])

    def check(f):
        """after frame f: composite == unsharded, and (whole-row exchange) every shard holds the whole previous frame"""
        img = gather_rows({s: p.read_image() for s, p in pipes.items()})
        res = {s: p.read_reservoirs().reshape(h, w, 12) for s, p in pipes.items()}
        gbs = {s: [p.read_gbuffer(g) for g in range(3)] for s, p in pipes.items()}
        comp_res = gather_rows(res)
        if rank == 0:
            ok["image"] &= same(img, ref.read_image())
            rres = ref.read_reservoirs().reshape(h, w, 12)
            ok["local_reservoirs"] &= same(comp_res, rres)
            if halo == ALL:
                for s in pipes:
                    ok["reservoirs"] &= same(res[s], rres)
                    ok["gbuffers"] &= all(same(gbs[s][g], ref.read_gbuffer(g)) for g in range(3))

    def loop(ps, verify):
        for p in ps.values():
            p.reset_frame()
        for f in range(args.frames + 1):      # the first call after a reset renders frame 0 twice (rt.cpp:132-135)
            cam = camera(args, f)
            for p in ps.values():
                p.render_frame(prm, cam, w, h)
            if verify:
                ref.render_frame(prm, cam, w, h) if ref else None
                check(f)

    if args.verify:
        loop(pipes, True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(ps):
        best = None
        for _ in range(args.reps + 1):        # the first repetition warms up (launch-size estimates, caches)
            torch.cuda.synchronize()
            if torchrun:
                dist.barrier()
            e0.record()
            loop(ps, False)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if torchrun:
                t = torch.tensor([ms], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            best = ms if best is None else min(best, ms)
        return best / (args.frames + 1)

    ms_sharded = timed(pipes)
    status = {s: p.history_status() for s, p in pipes.items()}
    timeouts = sum(v[1] for v in status.values())
    if torchrun:
        t = torch.tensor([timeouts], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        timeouts = int(t.item())
    ms_single = None
    if rank == 0:
        single = gpurt.RTPipe(scene, accel)
        ms_single = timed({0: single}) if not torchrun else None
        if torchrun:   # the other ranks are not in this loop: no barrier / all-reduce inside
            best = None
            for _ in range(args.reps + 1):
                torch.cuda.synchronize()
                e0.record()
                loop({0: single}, False)
                e1.record()
                torch.cuda.synchronize()
                best = e0.elapsed_time(e1) if best is None else min(best, e0.elapsed_time(e1))
            ms_single = best / (args.frames + 1)
        single.close()
    if torchrun:
        dist.barrier()
    if rank == 0:
        rows_out = 0
        if n_shards > 1:   # rows x receivers of shard 0, per frame
            import ctypes as C
            per_row = w * 96
            recv = 0
            for y in rows[owner == 0]:
                if halo == ALL:
                    recv += n_shards - 1
                else:
                    y0, y1 = max(0, y - halo), min(h - 1, y + halo)
                    recv += len({(b % n_shards) for b in range(y0 // band, y1 // band + 1)} - {0})
            rows_out = recv * per_row
        print(json.dumps({
            "workload": f"{os.path.basename(args.scene)} {w}x{h}, integrator {args.integrator}, depth 4, 1 spp, res_samples 4, temporal reuse"
                        + (f", spatial reuse {args.spatial} x r{args.radius:g}" if args.spatial else "") + (", moving camera" if args.orbit else ""),
            "n_shards": n_shards, "processes": world, "bands": f"{band} rows, {(h + band - 1) // band} bands",
            "halo_rows": "all" if halo == ALL else halo, "frames": args.frames + 1,
            "ms_per_frame_sharded": ms_sharded, "ms_per_frame_one_gpu": ms_single,
            "speedup": (ms_single / ms_sharded) if (ms_single and torchrun) else None,
            "bytes_pushed_per_frame_shard0": rows_out, "flag_wait_timeouts": timeouts,
            "verified": ok if args.verify else None}))
    for p in pipes.values():
        p.close()
    if ref:
        ref.close()
    for m in maps:
        m.close()
    accel.close(), scene.close(), ctx.close()
    if torchrun:
        dist.destroy_process_group()
    if args.verify and rank == 0 and not all(ok.values()):
        sys.exit(1)


if __name__ == "__main__":
    main()

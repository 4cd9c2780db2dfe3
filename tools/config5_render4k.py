#!/usr/bin/env python
"""BASELINE config 5: Sponza 3840x2160, 64 spp as samples=8 x max_frames=8 (the reference's
progressive scheme, rt.rgen:638-645), integrator 1 (Material), depth 8, RR on; image rows sharded in
interleaved 16-row bands over 1/2/4/8 B200, RGBA32F tiles gathered to rank 0 over NCCL.

    python tools/config5_render4k.py [--size 3840 2160] [--frames 8] [--spp 8] [--out out.png]
    torchrun --nproc-per-node N tools/config5_render4k.py ...

Prints one JSON line on rank 0: Mpaths/s, s/frame-set (max device time over ranks), gather ms, and —
with --verify — whether the gathered image is bit-identical to the same render done unsharded on rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gpurt  # noqa: E402
from gpurt.dist import gather_to_rank0, shared_result_buffer, warmup  # noqa: E402

BAND = 16


def render(pipe, prm, cam, w, h, ctx):
    pipe.reset_frame()
    ms, frames = 0.0, 0
    # first call after reset renders frame 0; RTPipe::update_uniforms re-renders frame 0 once when the
    # camera is first seen (rt.cpp:132-135), exactly like the reference
    while pipe.render_frame(prm, cam, w, h) == 0:
        ms += pipe.time_ms()
        frames += 1
    return ms, frames


def frame_parallel(args, ctx, scene, accel, pipe, prm, cam, w, h, rank, world, dev, label):
    F = args.frames
    img_bytes = w * h * 16
    means = shared_result_buffer(ctx, F * img_bytes)      # rank 0 owns it, the others map it over NVLink
    mine = list(range(rank, F, world))

    def run():
        ms = 0.0
        for f in mine:
            pipe.render_frame_mean(prm, cam, w, h, f, means.at(f * img_bytes))
            ms += pipe.time_ms()
        torch.cuda.synchronize()
        return ms

    run()                                                  # warm-up
    if world > 1:
        dist.barrier()
    t0 = time.time()
    ms = run()
    if world > 1:
        dist.barrier()                                     # every mean is in rank 0's memory
    fold_ms = 0.0
    if rank == 0:
        for f in range(F):
            pipe.accumulate_mean(means.at(f * img_bytes), f, w, h)
            fold_ms += pipe.time_ms()
        torch.cuda.synchronize()
    wall = time.time() - t0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        full = pipe.device_image()
        identical = None
        if args.verify:
            ref_pipe = gpurt.RTPipe(scene, accel)
            render(ref_pipe, prm, cam, w, h, ctx)
            identical = bool(torch.equal(ref_pipe.device_image().view(torch.int32), full.view(torch.int32)))
        total_s = (t.item() + fold_ms) * 1e-3
        print(json.dumps({
            "config": "Sponza 4K progressive path tracing (SURVEY config 5), frame-parallel sharding", "scene": label,
            "n_gpus": world, "size": [w, h], "spp_total": args.spp * F, "frames_rendered": F, "depth": args.depth,
            "s_total_max_rank": total_s, "render_s_max_rank": t.item() * 1e-3, "fold_ms_rank0": fold_ms, "wall_s": wall,
            "mpaths_s": w * h * args.spp * F / total_s / 1e6, "gather_ms": 0.0,
            "results": "frame means stored into rank 0's buffer by the frame-end kernel (gpurt_shared_*), no collective",
            "bit_identical_to_unsharded": identical, "mean_radiance": float(torch.nanmean(full[..., :3]))}), flush=True)
    if world > 1:
        dist.barrier()
    means.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, nargs=2, default=[3840, 2160])
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--spp", type=int, default=8)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--verify", action="store_true")
    ap.add_argument("--out", default="")
    ap.add_argument("--frame-parallel", action="store_true",
                    help="shard by progressive frame instead of by row band: rank r renders frames r, r+N, ... at full "
                         "resolution and stores their means straight into rank 0's buffer over NVLink; rank 0 folds them in order")
    ap.add_argument("--emulate-shards", type=int, default=0, help="single process: render only shard 0 of N (per-rank cost probe)")
    args = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        warmup(dev)
    w, h = args.size
    ctx = gpurt.Context(local)
    ctx.use_torch_stream()
    scene, label = bench.build_scene(gpurt, ctx)
    accel = gpurt.Accel(scene)
    pipe = gpurt.RTPipe(scene, accel)
    cam = gpurt.camera(1, w, h, bench.CAM_POS, bench.CAM_AT, bench.VFOV)
    # max_frames-1 because frames are numbered from 0 and trace() stops when frame >= max_frames
    prm = gpurt.pipe_params(integrator=1, brdf=1, max_depth=args.depth, samples_per_frame=args.spp,
                            max_frames=args.frames - 1, use_rr=1, env_scale=1.0, seed=7)
    if args.frame_parallel:
        return frame_parallel(args, ctx, scene, accel, pipe, prm, cam, w, h, rank, world, dev, label)
    pipe.set_shard(BAND, args.emulate_shards or world, rank)
    render(pipe, prm, cam, w, h, ctx)                      # warm-up (allocations, L2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms, frames = render(pipe, prm, cam, w, h, ctx)
    img = pipe.device_image()                              # (h, w, 4) view; only this rank's bands are filled
    nb = (h + BAND - 1) // BAND
    assert h % BAND == 0, "use a height that is a multiple of 16"
    mine = img.view(nb, BAND, w, 4)[rank::world].contiguous()
    torch.cuda.synchronize()
    g0 = time.time()
    gathered = gather_to_rank0(mine) if world > 1 else None   # ranks own different band counts
    torch.cuda.synchronize()
    gather_ms = (time.time() - g0) * 1e3
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        full = torch.empty((nb, BAND, w, 4), dtype=torch.float32, device=dev)
        if world > 1:
            off = 0
            for r in range(world):
                cnt = len(range(r, nb, world))
                full[r::world] = gathered[off:off + cnt]
                off += cnt
        else:
            full.copy_(img.view(nb, BAND, w, 4))
        full = full.view(h, w, 4)
        identical = None
        if args.verify:
            ref_pipe = gpurt.RTPipe(scene, accel)
            render(ref_pipe, prm, cam, w, h, ctx)
            ref = ref_pipe.device_image()
            identical = bool(torch.equal(ref.view(torch.int32), full.view(torch.int32)))
        paths = w * h * args.spp * frames
        print(json.dumps({
            "config": "Sponza 4K progressive path tracing (SURVEY config 5)", "scene": label, "n_gpus": world, "size": [w, h],
            "spp_total": args.spp * args.frames, "frames_rendered": frames, "depth": args.depth,
            "s_total_max_rank": t.item() * 1e-3, "mpaths_s": paths / (t.item() * 1e-3) / 1e6,
            "gather_ms": gather_ms, "gather_bytes": int(w * h * 16 * (world - 1) / world),
            "bit_identical_to_unsharded": identical, "mean_radiance": float(torch.nanmean(full[..., :3]))}), flush=True)
        if args.out:
            from PIL import Image
            x = 1.0 - torch.exp(-full[..., :3])
            x = x.clamp(0, 1) ** (1 / 2.2)
            Image.fromarray((x * 255 + 0.5).to(torch.uint8).cpu().numpy()).save(args.out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

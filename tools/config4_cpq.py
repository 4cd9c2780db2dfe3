#!/usr/bin/env python
"""BASELINE config 4: synthetic 10,000,000-triangle soup, 100,000,000 closest-point
queries sharded across 1/2/4/8 B200 (scene replicated, contiguous query ranges per rank, no collective
on the data path; an optional NCCL gather of the 32-byte results to rank 0 is timed separately).

    python tools/config4_cpq.py [--tris 10000000] [--queries 100000000] [--gather]
    torchrun --nproc-per-node N tools/config4_cpq.py ...

Prints one JSON line on rank 0: M queries/s over all ranks (max device time over ranks), build time,
and the oracle check on a subsample (rank 0; --check N queries through the CPU BVH).
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
import gpurt  # noqa: E402
from gpurt.dist import gather_to_rank0, shard_range, shared_result_buffer, warmup  # noqa: E402

M32 = 0xFFFFFFFF


def tea_t(v0, v1):
    """rtcommon.glsl:99-109 on int64 tensors holding uint32 values"""
    s0 = 0
    for _ in range(16):
        s0 = (s0 + 0x9E3779B9) & M32
        v0 = (v0 + ((((v1 << 4) & M32) + 0xA341316C) ^ (v1 + s0) ^ ((v1 >> 5) + 0xC8013EA4))) & M32
        v1 = (v1 + ((((v0 << 4) & M32) + 0xAD90777D) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7E95761E))) & M32
    return v0


def randf_t(state):
    """rtcommon.glsl:111-120; returns (float32 in [0,1), new state)"""
    state = (state * 1664525 + 1013904223) & M32
    return (state & 0x00FFFFFF).to(torch.float32) / 16777216.0, state


def make_soup(n, device):
    """triangle i: s = tea(i, 0xBEEF); centre uniform in [0,1]^3, two edges uniform in [-0.004,0.004]^3"""
    i = torch.arange(n, dtype=torch.int64, device=device)
    s = tea_t(i, torch.full_like(i, 0xBEEF))
    vals = []
    for _ in range(9):
        f, s = randf_t(s)
        vals.append(f)
    c = torch.stack(vals[0:3], 1)
    e1 = (torch.stack(vals[3:6], 1) - 0.5) * 0.008
    e2 = (torch.stack(vals[6:9], 1) - 0.5) * 0.008
    return torch.stack([c, c + e1, c + e2], 1).reshape(n, 9).contiguous()


def make_grid(side, device):
    """coherent variant: height field over side x side vertices, z = 0.05 sin(8 pi x) cos(8 pi y);
    side 2237 gives 9,999,392 triangles"""
    u = torch.arange(side, dtype=torch.float32, device=device) / (side - 1)
    x, y = torch.meshgrid(u, u, indexing="ij")
    z = 0.05 * torch.sin(8 * torch.pi * x) * torch.cos(8 * torch.pi * y)
    v = torch.stack([x, y, z], -1)
    a, b, c, d = v[:-1, :-1], v[1:, :-1], v[:-1, 1:], v[1:, 1:]
    t = torch.stack([torch.stack([a, b, c], -2), torch.stack([b, d, c], -2)], 2)   # (s-1, s-1, 2, 3, 3)
    return t.reshape(-1, 9).contiguous()


def make_queries(a, b, device):
    """query i: s = tea(i, 0xD00D); uniform in [-0.25,1.25]^3, r2 = +inf"""
    i = torch.arange(a, b, dtype=torch.int64, device=device)
    s = tea_t(i, torch.full_like(i, 0xD00D))
    q = torch.empty((b - a, 4), dtype=torch.float32, device=device)
    for k in range(3):
        f, s = randf_t(s)
        q[:, k] = f * 1.5 - 0.25
    q[:, 3] = float("inf")
    return q


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tris", type=int, default=10_000_000)
    ap.add_argument("--queries", type=int, default=100_000_000)
    ap.add_argument("--mesh", choices=["soup", "grid"], default="soup", help="grid = coherent height field (2237^2 vertices)")
    ap.add_argument("--chunk", type=int, default=12_500_000)
    ap.add_argument("--check", type=int, default=200_000)
    ap.add_argument("--gather", action="store_true", help="NCCL gather of the results to rank 0 after the queries")
    ap.add_argument("--p2p", action="store_true",
                    help="kernels store their results straight into rank 0's buffer over NVLink (no gather)")
    args = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        warmup(dev)
    ctx = gpurt.Context(local)
    ctx.use_torch_stream()

    tris = make_soup(args.tris, dev) if args.mesh == "soup" else make_grid(int(round((args.tris / 2) ** 0.5)) + 1, dev)
    tris_h = tris.cpu().numpy()
    del tris
    scene = gpurt.Scene(ctx)
    t0 = time.time()
    scene.add_triangles(tris_h)
    accel = gpurt.Accel(scene)
    t_build_wall = time.time() - t0
    info = accel.info()

    a, b = shard_range(args.queries, rank, world)
    shared = shared_result_buffer(ctx, args.queries * 32) if args.p2p else None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = 0.0
    n_done = 0
    results = []
    for c0 in range(a, b, args.chunk):
        c1 = min(b, c0 + args.chunk)
        q = make_queries(c0, c1, dev)          # generated on the device, not timed
        out = shared.at(c0 * 32) if args.p2p else torch.empty((c1 - c0, 8), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        accel.closest_points(q, out)
        ms += ctx.last_kernel_ms()
        n_done += c1 - c0
        if args.gather or (rank == 0 and c0 == a and not args.p2p):
            results.append(out if args.gather else out[: args.check].clone())
        del q
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(n_done)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    gather_ms = None
    if args.gather:
        local_res = torch.cat(results)
        torch.cuda.synchronize()
        g0 = time.time()
        full = gather_to_rank0(local_res)
        torch.cuda.synchronize()
        gather_ms = (time.time() - g0) * 1e3
        if rank == 0:
            assert full.shape[0] == args.queries

    chk0 = a
    if args.p2p:
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        if rank == 0:   # check the part of the array the LAST rank wrote over NVLink
            chk0 = shard_range(args.queries, world - 1, world)[0]
            full = shared.tensor().view(torch.float32).view(-1, 8)
            results = [full[chk0:chk0 + args.check].clone()]
    check = None
    if rank == 0 and args.check:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import orc
        n_chk = min(args.check, b - a)
        q = make_queries(chk0, chk0 + n_chk, dev).cpu().numpy()
        got = results[0][:n_chk].cpu().numpy().view(np.uint32)
        t1 = time.time()
        ob = orc.Bvh(tris_h, sah=not (accel.flags & gpurt.BUILD_LBVH))
        ref = ob.closest_point(q).view(np.uint32).reshape(-1, 8)
        same = bool((got[:, [0, 1, 2, 3, 4, 6, 7]] == ref[:, [0, 1, 2, 3, 4, 6, 7]]).all())
        order_same = bool((accel.prim_order() == ob.prim_order()).all())
        check = {"queries": n_chk, "bit_exact_vs_oracle": same, "prim_order_equals_oracle": order_same,
                 "oracle_s": round(time.time() - t1, 1)}
    if rank == 0:
        print(json.dumps({
            "config": f"synthetic CPQ (SURVEY config 4, {args.mesh})", "n_gpus": world, "tris": info.n_tris, "queries": int(tot.item()),
            "mqueries_s": tot.item() / (t.item() * 1e-3) / 1e6, "kernel_ms_max_rank": t.item(),
            "bvh_build_ms_device": info.build_ms, "bvh_build_s_wall_incl_upload": round(t_build_wall, 2),
            "build_mtris_s": info.n_tris / (info.build_ms * 1e-3) / 1e6, "wide_nodes": info.n_wide_nodes,
            "wide_depth": info.wide_depth, "gather_ms": gather_ms, "results": "p2p stores into rank 0" if args.p2p else
            ("nccl gather" if args.gather else "left on each rank"), "check": check}), flush=True)
    if shared is not None:
        if world > 1:
            dist.barrier()
        shared.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

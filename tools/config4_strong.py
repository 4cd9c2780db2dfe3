#!/usr/bin/env python
"""bench.py's config-4 strong-scaling sub-benchmark on its own (10 M-triangle soup, 100 M closest-point queries, results
placed in rank 0's memory inside the timed region), through gpurt_gather_* (default, "gather") or with caller-side chunks by storage position for one or several chunk schedules:

    torchrun --nproc-per-node 8 tools/config4_strong.py --schedule gather --schedule 0.75,0.17,0.08 --schedule 0.25,0.25,0.25,0.25

One JSON line per schedule on rank 0."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
import bench  # noqa: E402
import gpurt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--schedule", action="append", default=[])
    ap.add_argument("--tris", type=int, default=10_000_000)
    ap.add_argument("--queries", type=int, default=100_000_000)
    ap.add_argument("--check", type=int, default=0)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = gpurt.Context(local)
    ctx.use_torch_stream()
    for sch in args.schedule or ["gather"]:
        out = bench.strong_config4(gpurt, torch, dist, ctx, rank, world, dev, args.tris, args.queries, args.check,
                                   schedule=None if sch == "gather" else [float(x) for x in sch.split(",")])
        if rank == 0:
            out["schedule"] = sch
            print(json.dumps(out), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

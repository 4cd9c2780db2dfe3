"""world_size-2 gloo test of the sharding / gather plumbing used by bench.py and the multi-GPU
configs.  Compute on the CPU ranks is the oracle (there is no GPU here); the point is
that sharded + gathered == single rank, byte for byte."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scenes import soup


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    import sys
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    from gpurt.dist import gather_to_rank0, row_bands, shard_range
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tris = soup(2000, seed=9)
    bvh = orc.Bvh(tris)
    rays = orc.gen_random_rays(10001, 3, bvh.scene_box())      # odd count: ragged shards
    a, b = shard_range(rays.shape[0], rank, world)
    local = torch.from_numpy(bvh.closest_hit(rays[a:b].copy(), threads=1).view(np.uint32).reshape(-1, 4).astype(np.int64))
    full = gather_to_rank0(local)
    empty = gather_to_rank0(local[:0] if rank == 1 else local)  # a rank with nothing to send
    bands = row_bands(100, rank, world)
    cover = torch.zeros(100, dtype=torch.int64)
    for y0, y1 in bands:
        cover[y0:y1] += 1
    dist.all_reduce(cover)
    if rank == 0:
        ref = bvh.closest_hit(rays, threads=1).view(np.uint32).reshape(-1, 4).astype(np.int64)
        ok = bool((full.numpy() == ref).all()) and empty.shape[0] == b - a and bool((cover == 1).all())
        open(out_path, "w").write("ok" if ok else "mismatch")
    else:
        assert full is None
    dist.destroy_process_group()


def test_shard_and_gather_two_ranks(built, tmp_path):
    out = str(tmp_path / "result")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_shard_ranges_cover_exactly(built, gpurt):
    from gpurt.dist import shard_range
    for n in (0, 1, 7, 100_000_000):
        for world in (1, 2, 4, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[k][1] == r[k + 1][0] for k in range(world - 1))

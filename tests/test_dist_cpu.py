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


def _restir_worker(rank, world, port, out_path):
    """ReSTIR on a sharded frame (gpurt_pipe_history_peers): every rank replays the product's shading code (tests/emu) on
    its row bands, then the rows of the frame just rendered travel to the ranks history_row_readers() names — here through
    gloo, on the GPU through k_history_push over peer memory — and the next frame's temporal / spatial reuse reads them."""
    import ctypes as C
    import sys
    from conftest import MEDIA, ROOT
    sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import gpurt
    import orc
    from test_emu_render import EmuScene, _uniforms
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emu = C.CDLL(os.path.join(ROOT, "tests", "emu", "libemu.so"))
    emu.emu_build.restype = C.c_void_p
    emu.emu_render_frame.restype = None
    emu.emu_history_row_readers.restype = C.c_uint64
    emu.emu_set_spatial.argtypes = [C.c_uint32, C.c_float]
    s = gpurt.Scene(None).load(os.path.join(MEDIA, "cbox", "cbox.gltf"))
    es = EmuScene(emu, orc, s, ())
    w, h = 64, 40
    ok = True
    ALL = 0xFFFFFFFF
    #        band rows, halo,  spatial,   camera x per frame
    cases = [(16, ALL, (0, 16.0), [0.0, 0.0, 0.05, 0.12, 0.12]),      # interleaved bands, moving camera: whole rows everywhere
             ((h + world - 1) // world, 7, (3, 5.0), [0.0] * 4)]       # contiguous bands, static camera, halo covers the disc
    for band, halo, spatial, cam_x in cases:
        emu.emu_set_spatial(*spatial)
        ref, st = orc.FrameState(w, h), orc.FrameState(w, h)
        owner = (np.arange(h) // band) % world
        prev_pv = None
        for f, cx in enumerate(cam_x):
            cam = gpurt.camera(1, w, h, (cx, 1.0, 3.4), (0.0, 1.0, 0.0), 40.0)
            consts, ubo, seed = _uniforms(gpurt, es.rs, cam, f, integrator=4 if band == 16 else 3, brdf=1, samples_per_frame=1,
                                          max_depth=3, res_samples=4, use_temporal=1, temporal_scale=16, seed=77)
            c = gpurt.Camera.from_buffer_copy(ubo.tobytes())
            pv = (np.array(c.P, np.float32).reshape(4, 4).T @ np.array(c.V, np.float32).reshape(4, 4).T).T.reshape(-1)
            c.prev_PV = (C.c_float * 16)(*(pv if prev_pv is None else prev_pv))
            prev_pv = pv
            ubo = np.frombuffer(bytes(c), np.uint32).copy()
            emu.emu_set_shard(0, 1, 0)
            es.render_frame(ref, consts, ubo, seed ^ f)
            emu.emu_set_shard(band, world, rank)
            es.render_frame(st, consts, ubo, seed ^ f)
            cur = st.parity ^ 1
            # the exchange: rows I rendered -> the ranks that read them next frame
            for arr in [st.res[cur].reshape(h, w * 12)] + [g.reshape(h, w * 4).view(np.uint32) for g in st.gb[cur]]:
                mine = torch.from_numpy(arr.astype(np.int64))
                everyone = [torch.zeros_like(mine) for _ in range(world)]
                dist.all_gather(everyone, mine)
                for y in range(h):
                    src = int(owner[y])
                    readers = emu.emu_history_row_readers(w, h, band, world, src, y, C.c_uint32(halo))
                    assert readers >> src & 1
                    if src != rank and readers >> rank & 1:
                        arr[y] = everyone[src][y].numpy().astype(np.uint32)
            # composite image == unsharded image; in whole-row mode every rank holds the whole previous frame
            img = torch.from_numpy(st.image.view(np.uint32).reshape(h, -1).astype(np.int64))
            imgs = [torch.zeros_like(img) for _ in range(world)]
            dist.all_gather(imgs, img)
            comp = np.stack([imgs[int(owner[y])][y].numpy() for y in range(h)]).astype(np.uint32)
            ok &= bool((comp == ref.image.view(np.uint32).reshape(h, -1)).all())
            local = owner == rank
            ok &= bool((st.res[cur].reshape(h, -1)[local] == ref.res[cur].reshape(h, -1)[local]).all())
            if halo == ALL:
                ok &= bool((st.res[cur] == ref.res[cur]).all())
                ok &= all(bool((st.gb[cur][g].view(np.uint32) == ref.gb[cur][g].view(np.uint32)).all()) for g in range(3))
        assert np.abs(ref.image[..., :3]).sum() > 0
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        open(out_path, "w").write("ok" if int(flag) else "mismatch")
    es.close()
    dist.destroy_process_group()


def test_restir_sharded_with_history_exchange_two_ranks(built, tmp_path):
    out = str(tmp_path / "result")
    mp.spawn(_restir_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_history_row_readers(built):
    """who receives a row of the previous frame (shade.cuh history_row_readers, used by k_history_push)"""
    import ctypes as C
    from conftest import ROOT
    emu = C.CDLL(os.path.join(ROOT, "tests", "emu", "libemu.so"))
    emu.emu_history_row_readers.restype = C.c_uint64
    rd = lambda h, band, n, y, halo: emu.emu_history_row_readers(64, h, band, n, (y // band) % n, y, C.c_uint32(halo))
    assert rd(1080, 135, 8, 500, 0xFFFFFFFF) == 0xFF
    for h, band, n, halo in ((1080, 135, 8, 16), (100, 16, 3, 5), (64, 16, 4, 0), (64, 16, 2, 40), (37, 5, 64, 3)):
        for y in range(h):
            want = 0
            for yy in range(max(0, y - halo), min(h - 1, y + halo) + 1):
                want |= 1 << ((yy // band) % n)
            assert rd(h, band, n, y, halo) == want, (h, band, n, halo, y)
    assert emu.emu_shard_row(64, 100, 16, 3, 1, 0) == 16 and emu.emu_shard_row(64, 100, 16, 3, 1, 17) == 65

"""Data-parallel plumbing for the sharded workloads: rays, query points and image rows
are independent, the scene + BVH are replicated on every GPU, and the only collective is the gather
of result arrays / image tiles to rank 0.  torch.distributed is plumbing only (NCCL over NVLink on
the GPU box, gloo in the CPU tests); all compute goes through libgpurt.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """contiguous range [r*n/R, (r+1)*n/R) of rank r"""
    return (n * rank) // world, (n * (rank + 1)) // world


def row_bands(height, rank, world):
    """interleaved 16-row bands of an image for load balance: list of (y0, y1)"""
    band = 16
    return [(y, min(height, y + band)) for i, y in enumerate(range(0, height, band)) if i % world == rank]


def warmup(device):
    """create the NCCL communicators (collective + point-to-point) before anything is timed"""
    if dist.is_initialized() and dist.get_world_size() > 1:
        gather_to_rank0(torch.zeros((1 + dist.get_rank(), 4), device=device))
        torch.cuda.synchronize() if device.type == "cuda" else None


def shared_result_buffer(ctx, nbytes):
    """Rank 0 allocates `nbytes` of result storage, every other rank maps it over NVLink
    (gpurt_shared_alloc / gpurt_shared_open).  A rank then passes `buf.at(first_element * stride)` as
    the result pointer of its query call: the traversal kernel stores straight into rank 0's HBM and
    no gather collective follows.  Call `dist.barrier()` after the ranks' streams are synchronised and
    before rank 0 reads `buf.tensor()`."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return ctx.shared_alloc(nbytes)
    box = [None]
    if dist.get_rank() == 0:
        buf = ctx.shared_alloc(nbytes)
        box[0] = buf.handle
    dist.broadcast_object_list(box, src=0)
    if dist.get_rank() != 0:
        buf = ctx.shared_open(box[0], nbytes)
    return buf


def gather_to_rank0(local, counts=None):
    """Gather variable-length first-dimension shards to rank 0 (returns the concatenation on rank 0,
    None elsewhere).  Uses all_gather of the lengths + gather / point-to-point of the payload."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n)
    ns = [int(x.item()) for x in ns]
    if rank == 0:
        parts = [local] + [torch.empty((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) for m in ns[1:]]
        reqs = [dist.irecv(parts[r], src=r) for r in range(1, world) if ns[r]]
        for q in reqs:
            q.wait()
        return torch.cat(parts)
    if local.shape[0]:
        dist.send(local.contiguous(), dst=0)
    return None


def share_history(pipe, ctx, w, h, halo_rows=0xFFFFFFFF):
    """ReSTIR on a frame sharded over the ranks (include/gpurt.h, gpurt_pipe_history_peers): every rank exports the block
    holding its previous-frame G-buffers + reservoirs, maps the other ranks' blocks over NVLink and hands the mappings to
    its pipe; from then on each rendered frame ends with the rank's rows stored into the peers' blocks and a flag, and
    begins by waiting for the peers' flags — no host synchronisation or collective per frame.  Call after
    pipe.set_shard(...) and again whenever the frame size changes.  Returns the mappings (keep them alive; close them
    after the pipe)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return []
    world, rank = dist.get_world_size(), dist.get_rank()
    _, handle, nbytes = pipe.history_export(w, h)
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    maps = [None if r == rank else ctx.shared_open(handles[r], nbytes) for r in range(world)]
    pipe.history_peers([0 if m is None else m.ptr for m in maps], halo_rows)
    dist.barrier()   # nobody renders into a block that is not mapped everywhere yet
    return [m for m in maps if m is not None]

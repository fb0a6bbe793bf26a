"""Multi-GPU sharding of a batch: one process per GPU, contiguous blocks of the
batch index per rank, no collective inside the solve, one all-gather of the
per-unit results (64-bit keep masks, flags, LP counts) at the end
(SURVEY.md section 8e).  `torch.distributed` (NCCL on GPUs, gloo in the CPU
tests) is plumbing; every LP still runs in the CUDA kernels.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world):
    """[lo, hi) of the contiguous block of `rank`; the first n % world ranks get one more."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allgather_blocks(local, n_items, group=None):
    """Concatenate every rank's block (first dimension) in rank order.

    Blocks follow `shard_bounds`, so they differ by at most one row: each rank
    pads to the largest block, one all_gather moves everything, padding is cut.
    """
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_items, r, world) for r in range(world)]
    big = max(hi - lo for lo, hi in sizes)
    assert local.shape[0] == sizes[dist.get_rank(group)][1] - sizes[dist.get_rank(group)][0]
    pad = local.new_zeros((big,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)


def sharded_map(fn, n_items, group=None):
    """Run `fn(lo, hi) -> tuple of tensors` on this rank's block and all-gather
    each result; returns the tuple for the whole batch on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return fn(0, n_items)
    lo, hi = shard_bounds(n_items, dist.get_rank(group), dist.get_world_size(group))
    return tuple(allgather_blocks(t, n_items, group) for t in fn(lo, hi))


def reduce_batch_sharded(A, b, group=None, abs_tol=1e-7, normalize=True):
    """engine.reduce_batch over the ranks of `group`: rank r reduces polytopes
    shard_bounds(P, r, world) of the (replicated or host-resident) batch and the
    keep masks / flags / LP counts of all P polytopes are all-gathered.
    -> (keep int64[P], flags int32[P], n_lp int32[P]) CUDA tensors."""
    from polytope_b200 import engine

    def local(lo, hi):
        Al = torch.as_tensor(A[lo:hi]).to('cuda', non_blocking=True)
        bl = torch.as_tensor(b[lo:hi]).to('cuda', non_blocking=True)
        res = engine.reduce_batch(Al, bl, abs_tol=abs_tol, normalize=normalize, want_A=False)
        return res.keep, res.flags, res.n_lp
    return sharded_map(local, A.shape[0], group)

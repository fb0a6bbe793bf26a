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


def pair_block(n_cells, lo, hi, device='cpu'):
    """Pairs lo..hi-1 of the enumeration t -> (i, j), j < i, t = i(i-1)/2 + j, that
    find_adjacent_regions walks (prop2partition.py:57-61): int32 tensors (i, j)."""
    t = torch.arange(lo, hi, dtype=torch.int64, device=device)
    i = torch.floor((1.0 + torch.sqrt(1.0 + 8.0 * t.double())) * 0.5).to(torch.int64)
    i = torch.where(i * (i - 1) // 2 > t, i - 1, i)
    i = torch.where((i + 1) * i // 2 <= t, i + 1, i)
    j = t - i * (i - 1) // 2
    assert n_cells < 2 or hi <= n_cells * (n_cells - 1) // 2
    return i.to(torch.int32), j.to(torch.int32)


def adjacency_sharded(A, b, group=None, abs_tol=1e-7):
    """cfg5: is_adjacent over all pairs j < i of `ncell` cells (prop2partition.py:46-63),
    the pair index range split across the ranks of `group`.  Every rank holds all
    cells (they are tiny) and enumerates its own pair range; one all-gather of the
    uint8 flags.  -> flags uint8[ncell (ncell - 1) / 2] on every rank."""
    from polytope_b200 import engine
    ncell = A.shape[0]
    T = ncell * (ncell - 1) // 2
    Ad = torch.as_tensor(A).to('cuda')
    bd = torch.as_tensor(b).to('cuda')

    def local(lo, hi):
        adj, _, _ = engine.adjacent_range(Ad, bd, 0, lo, hi - lo, abs_tol=abs_tol)
        return (adj,)
    return sharded_map(local, T, group)[0]


def ordered_pair_block(n_cells, lo, hi, device='cpu'):
    """Pairs lo..hi-1 of MetricPartition.compute_adj's double loop (prop2partition.py:253-261):
    all ordered pairs (i, j), i != j, row-major: t = i (n - 1) + jj, j = jj + (jj >= i)."""
    t = torch.arange(lo, hi, dtype=torch.int64, device=device)
    i = t // (n_cells - 1)
    jj = t - i * (n_cells - 1)
    j = jj + (jj >= i).to(torch.int64)
    return i.to(torch.int32), j.to(torch.int32)


def adjacency_ordered_sharded(A, b, group=None, abs_tol=1e-7):
    """cfg5 with compute_adj semantics: is_adjacent over all n (n - 1) ordered pairs, the pair
    range split across the ranks; one all-gather of the uint8 flags.
    -> flags uint8[n (n - 1)] on every rank (row-major over i, diagonal left out)."""
    from polytope_b200 import engine
    ncell = A.shape[0]
    T = ncell * (ncell - 1)
    Ad = torch.as_tensor(A).to('cuda')
    bd = torch.as_tensor(b).to('cuda')

    def local(lo, hi):
        adj, _, _ = engine.adjacent_range(Ad, bd, 1, lo, hi - lo, abs_tol=abs_tol)
        return (adj,)
    return sharded_map(local, T, group)[0]


def extreme_tensor_sharded(A, b, group=None, gather_vertices=True, caps=None):
    """cfg4 on stacked tensors: extreme() of P polytopes (engine.extreme_pipeline), the batch split
    across the ranks in contiguous blocks; the per-polytope vertex counts are all-gathered and
    (optionally) the vertices with one padded all-gather.  `caps` = (facet_cap,
    out_cap) of a previous call avoids the capacity retries.
    -> (counts int32[P], V, (facet_cap, out_cap))"""
    from polytope_b200 import engine
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    P = A.shape[0]
    lo, hi = shard_bounds(P, rank, world)
    cnt, V, hull, _ = engine.extreme_pipeline(A[lo:hi], b[lo:hi], *(caps or (None, None)))
    caps = (hull.facet_cap, int(cnt.sum().item()) + 1024)
    if world == 1:
        return cnt, V, caps
    counts = allgather_blocks(cnt, P, group)
    if not gather_vertices:
        return counts, V, caps
    allV, _ = allgather_ragged(V, group)
    return counts, allV, caps


def allgather_ragged(rows, group=None):
    """Concatenate every rank's `rows` (first dimension of any length) in rank order:
    the row counts are all-gathered first, then one padded all_gather moves the data
    (variable-length outputs such as extreme()'s vertices).
    -> (all_rows, counts int64[world])"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rows, torch.tensor([rows.shape[0]], dtype=torch.int64)
    world = dist.get_world_size(group)
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    counts = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = torch.cat(counts).cpu()
    big = int(counts.max())
    pad = rows.new_zeros((big,) + tuple(rows.shape[1:]))
    pad[:rows.shape[0]] = rows
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:int(c)] for o, c in zip(out, counts)], 0), counts


def extreme_sharded(polys, group=None, gather_vertices=True):
    """cfg4: extreme() of a list of polytopes, the list split across the ranks of `group`
    (contiguous blocks).  Every rank enumerates the vertices of its own block on its GPU;
    the per-polytope vertex counts are all-gathered, and (optionally) the vertices
    themselves with one padded all-gather.
    -> (counts int64[P] on every rank, vertices [sum(counts), d] or this rank's own rows)"""
    from polytope_b200 import polytope as pc
    import numpy as np
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_bounds(len(polys), rank, world)
    local = pc.extreme_batch(polys[lo:hi])
    d = polys[0].dim
    cnt = torch.tensor([0 if v is None else len(v) for v in local], dtype=torch.int64, device='cuda')
    V = np.concatenate([v for v in local if v is not None] or [np.zeros((0, d))], 0)
    Vt = torch.from_numpy(np.ascontiguousarray(V)).to('cuda')
    if world == 1:
        return cnt, Vt
    counts = allgather_blocks(cnt, len(polys), group)
    if not gather_vertices:
        return counts, Vt
    allV, _ = allgather_ragged(Vt, group)
    return counts, allV

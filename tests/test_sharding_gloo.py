"""world_size-2 gloo test of the multi-GPU host logic (block bounds, padded
all-gather, result assembly) on CPU; the per-block compute is a stand-in here
(the CUDA path needs a GPU and is covered by the -m gpu tests and bench.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from polytope_b200 import sharding


def test_shard_bounds_cover_the_batch():
    for n in (0, 1, 7, 10000, 10001):
        for world in (1, 2, 3, 8):
            blocks = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_pair_block_enumeration_matches_tril_order():
    """cfg5 sharding: rank-local pair ranges reproduce find_adjacent_regions' order."""
    for n in (2, 3, 10, 1024, 1415):
        T = n * (n - 1) // 2
        ii, jj = np.tril_indices(n, -1)
        for world in (1, 2, 8):
            got_i, got_j = [], []
            for r in range(world):
                lo, hi = sharding.shard_bounds(T, r, world)
                i, j = sharding.pair_block(n, lo, hi)
                got_i.append(i.numpy())
                got_j.append(j.numpy())
            assert np.array_equal(np.concatenate(got_i), ii) and np.array_equal(np.concatenate(got_j), jj)


def test_ordered_pair_block_matches_compute_adj_order():
    """cfg5 with compute_adj semantics (prop2partition.py:253-261): all ordered pairs i != j in
    the order of the reference's double loop, whatever the split across ranks."""
    for n in (2, 3, 10, 1024):
        ii, jj = np.nonzero(~np.eye(n, dtype=bool))
        for world in (1, 2, 8):
            got_i, got_j = [], []
            for r in range(world):
                lo, hi = sharding.shard_bounds(n * (n - 1), r, world)
                i, j = sharding.ordered_pair_block(n, lo, hi)
                got_i.append(i.numpy())
                got_j.append(j.numpy())
            assert np.array_equal(np.concatenate(got_i), ii) and np.array_equal(np.concatenate(got_j), jj)


def _worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        def fn(lo, hi):          # stand-in for engine.reduce_batch on rows lo..hi
            idx = torch.arange(lo, hi, dtype=torch.int64)
            return idx * idx + 1, (idx % 5).to(torch.int32), torch.stack([idx, -idx], 1).double()
        keep, flags, extra = sharding.sharded_map(fn, n_items)
        # ragged gather (extreme()'s vertices): rank r contributes r + 2 rows
        rows = torch.full((rank + 2, 3), float(rank))
        allrows, counts = sharding.allgather_ragged(rows)
        assert counts.tolist() == [r + 2 for r in range(world)]
        assert allrows.shape == (sum(r + 2 for r in range(world)), 3)
        assert torch.equal(allrows[:2], torch.zeros(2, 3)) and torch.equal(allrows[2:5], torch.ones(3, 3))
        q.put((rank, keep.numpy(), flags.numpy(), extra.numpy()))
    finally:
        dist.destroy_process_group()


def test_sharded_map_world2_gloo():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    for n_items in (11, 8):
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
        for p in procs:
            p.start()
        got = [q.get(timeout=120) for _ in range(2)]
        for p in procs:
            p.join(60)
            assert p.exitcode == 0
        idx = np.arange(n_items)
        for rank, keep, flags, extra in got:
            assert np.array_equal(keep, idx * idx + 1)
            assert np.array_equal(flags, idx % 5)
            assert np.array_equal(extra, np.stack([idx, -idx], 1).astype(float))

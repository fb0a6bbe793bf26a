#!/usr/bin/env python
"""torchrun check of the sharded entries (reduce masks, adjacency flags, extreme vertices):
every rank must end with the same, single-GPU-identical results.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as wl                              # noqa: E402
import polytope_b200 as pc                          # noqa: E402
from polytope_b200 import engine, sharding          # noqa: E402

rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
world = dist.get_world_size()

A, b = wl.box_cuts_batch(2, 1001, 32, 8)
keep, flags, nlp = sharding.reduce_batch_sharded(A, b)
one = engine.reduce_batch(torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda(), want_A=False)
assert torch.equal(keep, one.keep) and torch.equal(flags, one.flags) and torch.equal(nlp, one.n_lp)

Ag, bg, idx = wl.box_grid((9, 7))
adj = sharding.adjacency_sharded(Ag, bg)
ref, _, _ = engine.adjacent_pairs(torch.from_numpy(Ag).cuda(), torch.from_numpy(bg).cuda())
assert torch.equal(adj, ref)

polys = [pc.Polytope(*wl.box_cuts(4000 + i, 24, 6)) for i in range(13)]
counts, V = sharding.extreme_sharded(polys)
single = pc.extreme_batch([pc.Polytope(p.A, p.b) for p in polys])
assert counts.tolist() == [len(v) for v in single]
assert np.allclose(V.cpu().numpy(), np.concatenate(single, 0), atol=1e-12)
dist.barrier()
if rank == 0:
    print('multi_gpu_check ok: world %d, %d polytopes reduced, %d pairs, %d vertices' % (
        world, len(A), len(adj), int(counts.sum())))
dist.destroy_process_group()

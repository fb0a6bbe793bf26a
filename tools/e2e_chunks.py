#!/usr/bin/env python
"""End-to-end step time of the host-buffer reduce call against the number of pipeline chunks, on 1..N ranks
(torchrun).  Each step: barrier, pinned (A, b) -> engine.reduce_batch(results_on_device) -> all-gather -> D2H."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import workloads as wl
from polytope_b200 import engine, sharding

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
if world > 1:
    dist.init_process_group('nccl')
P = 10000
A, b = wl.box_cuts_batch(2, P, 32, 8, first=rank * P)
A_pin, b_pin = torch.from_numpy(A).pin_memory(), torch.from_numpy(b).pin_memory()


def step():
    res = engine.reduce_batch(A_pin, b_pin, want_A=False, want_b=False, results_on_device=True)
    packed = torch.stack([res.keep, res.flags.to(torch.int64), res.n_lp.to(torch.int64), res.lp_iters.to(torch.int64)], 1)
    if world > 1:
        packed = sharding.allgather_blocks(packed, world * P)
    return packed.cpu()


out = {}
for chunks in (2, 3, (0.35, 0.65), (0.25, 0.75), (0.15, 0.85), (0.2, 0.4, 0.4), (0.1, 0.45, 0.45), (0.15, 0.35, 0.5)):
    if isinstance(chunks, tuple):
        engine.PIPELINE_FRACTIONS = chunks
    else:
        engine.PIPELINE_FRACTIONS = None
        engine.PIPELINE_CHUNKS = chunks
    for _ in range(3):
        step()
    tot = 0.0
    for _ in range(10):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    t = torch.tensor([tot / 10], device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[str(chunks)] = round(float(t.item()), 3)
if rank == 0:
    print('world', world, 'ms per e2e step by chunks:', out)
if world > 1:
    dist.destroy_process_group()

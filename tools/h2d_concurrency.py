#!/usr/bin/env python
"""Pinned host -> device copy bandwidth per rank when 1..N ranks copy at the same time (torchrun): the part of the
end-to-end step that no kernel change can shorten."""
import os
import torch
import torch.distributed as dist

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
if world > 1:
    dist.init_process_group('nccl')
for mb in (23, 184):
    src = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    dst = torch.empty(mb << 20, dtype=torch.uint8, device='cuda')
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    tot = 0.0
    for _ in range(10):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    t = torch.tensor([tot / 10], device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item())
        print('world %d: %d MiB per rank in %.3f ms = %.1f GB/s per rank, %.1f GB/s aggregate' % (
            world, mb, ms, (mb << 20) / ms / 1e6, world * (mb << 20) / ms / 1e6))
if world > 1:
    dist.destroy_process_group()

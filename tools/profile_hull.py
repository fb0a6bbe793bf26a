"""Tiny driver for ncu: the dual hulls of cfg4 polytopes (extreme()'s hull stage). Usage: profile_hull.py [n m d]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import workloads as wl
from polytope_b200 import engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 444
m = int(sys.argv[2]) if len(sys.argv) > 2 else 64
d = int(sys.argv[3]) if len(sys.argv) > 3 else 12
A, b = wl.box_cuts_batch(4, n, m, d)
Ad, bd = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
res = engine.reduce_batch(Ad, bd)
bits = ((res.keep.unsqueeze(1) >> torch.arange(m, device='cuda')) & 1).bool()
rows = bits.sum(1).to(torch.int32)
order = torch.argsort((~bits).to(torch.int8), dim=1, stable=True)
Ar = torch.gather(res.A, 1, order.unsqueeze(-1).expand(-1, -1, d)).contiguous()
br = torch.gather(res.b, 1, order).contiguous()
r, xc, st = engine.cheby_batch(Ar, br, rows)
dual = engine.dual_points(Ar, br, xc, rows)
hull = engine.hull_batch(dual, rows, facet_cap=65536)
hull = engine.hull_batch(dual, rows, facet_cap=hull.facet_cap)
torch.cuda.synchronize()
print('hulls', n, 'facets', int(hull.facet_cnt.sum()), 'bad', int((hull.status != 0).sum()))
if len(sys.argv) > 4:
    for rep in range(int(sys.argv[4])):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hull = engine.hull_batch(dual, rows, facet_cap=hull.facet_cap)
        e1.record()
        torch.cuda.synchronize()
        print('rep', rep, 'hull_batch ms', e0.elapsed_time(e1), 'stats', hull.stats.float().mean(0).tolist() if hull.stats is not None else None)

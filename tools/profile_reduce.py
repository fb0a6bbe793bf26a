"""Tiny driver for ncu: a few cfg2 reduce_batch passes with the batch resident in HBM."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import workloads as wl
from polytope_b200 import engine

P = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 32
d = int(sys.argv[3]) if len(sys.argv) > 3 else 8
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
A, b = wl.box_cuts_batch(2, P, m, d)
A = torch.from_numpy(A).cuda()
b = torch.from_numpy(b).cuda()
for _ in range(reps):
    res = engine.reduce_batch(A, b, want_A=False)
torch.cuda.synchronize()
print('LPs', int(res.n_lp.sum()), 'iters/LP', float(res.lp_iters.sum()) / float(res.n_lp.sum()))

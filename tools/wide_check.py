#!/usr/bin/env python
"""A/B of the lane kernels (incl. the 9 <= n <= 16 wide solver) against the warp-per-LP kernels on
reduce(): per-stage device times, LPs/s, and equality of the keep masks of the two paths."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import workloads as wl
from polytope_b200 import engine


def run(cfg, n, m, d, lane):
    engine.lane_solver(lane)
    A, b = wl.box_cuts_batch(cfg, n, m, d)
    A, b = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    for _ in range(3):
        res = engine.reduce_batch(A, b, want_A=False)
    engine.profile_enable(True)
    acc, tot, reps = {}, 0.0, 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(reps):
        e0.record()
        res = engine.reduce_batch(A, b, want_A=False)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
        for k, v in engine.profile_read().items():
            acc[k] = acc.get(k, 0.0) + v / reps
    engine.profile_enable(False)
    lps = int(res.n_lp.sum())
    return res, {'ms': round(tot / reps, 4), 'MLPs_per_s': round(lps / (tot / reps) / 1e3, 2), 'LPs': lps,
                 'iters_per_lp': round(float(res.lp_iters.sum()) / lps, 3), 'stages_ms': {k: round(v, 4) for k, v in acc.items()}}


out = {}
for name, cfg, n, m, d in (('cfg2', 2, 10000, 32, 8), ('cfg4', 4, 1000, 64, 12), ('d16', 4, 500, 64, 16), ('d10', 4, 2000, 40, 10),
                           ('d14', 4, 500, 56, 14)):
    r1, s1 = run(cfg, n, m, d, True)
    r0, s0 = run(cfg, n, m, d, False)
    out[name] = {'lane': s1, 'warp': s0, 'keep_mismatches': int((r1.keep != r0.keep).sum()), 'flag_mismatches': int((r1.flags != r0.flags).sum()),
                 'n_lp_mismatches': int((r1.n_lp != r0.n_lp).sum()), 'max_abs_dr': float((r1.r - r0.r).abs().max())}
engine.lane_solver(True)
print(json.dumps(out, indent=1))

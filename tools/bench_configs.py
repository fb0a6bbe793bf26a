#!/usr/bin/env python
"""Throughput + parity sample for the BASELINE configs other than the headline
(cfg3 Region members / pairwise intersect, cfg4 d=12 m=64 reduce, cfg5 adjacency
grid).  Device-timed with CUDA events, batch resident in HBM, 3 warm-ups.
Prints one JSON object; commit the output under profiles/."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as wl                      # noqa: E402
from polytope_b200 import engine            # noqa: E402
from oracle import polytope_oracle as orc   # noqa: E402


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def cfg3(n=50000):
    A, b = wl.box_cuts_batch(3, n, 16, 6, shift_scale=True)
    Ad, bd = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    An, bn, _ = engine.normalize_batch(Ad, bd)
    ms_fd, (r, xc, st) = timed(lambda: engine.cheby_batch(An, bn))
    Q = orc.normalize_rows(*wl.box_cuts(3999, 16, 6, True))[:2]
    Qa = torch.from_numpy(Q[0]).cuda().expand(n, 16, 6)
    Qb = torch.from_numpy(Q[1]).cuda().expand(n, 16)
    SA = torch.cat([An, Qa], 1).contiguous()
    Sb = torch.cat([bn, Qb], 1).contiguous()
    ms_is, res = timed(lambda: engine.reduce_batch(SA, Sb, want_A=False))
    lps = int(res.n_lp.sum())
    keeps = res.keep.cpu().numpy().astype(np.uint64)
    flags = res.flags.cpu().numpy()
    bad = 0
    rng = np.random.default_rng(0)
    sample = rng.choice(n, 48, replace=False)
    for p in sample:
        o = orc.reduce(np.vstack([An[p].cpu().numpy(), Q[0]]), np.hstack([bn[p].cpu().numpy(), Q[1]]))
        mask = sum(1 << k for k in o['keep'])
        bad += int(mask != int(keeps[p])) + int(bool(flags[p] & 1) != o['empty'])
    return {'workload': 'cfg3: %d polytopes d=6 m=16 (shift/scale)' % n,
            'is_fulldim_LPs_per_s': n / (ms_fd * 1e-3), 'is_fulldim_ms': ms_fd,
            'fulldim_count': int((r > 1e-7).sum()),
            'pairwise_intersect_LPs_per_s': lps / (ms_is * 1e-3), 'pairwise_intersect_ms': ms_is,
            'pairwise_intersect_LPs': lps, 'nonempty_intersections': int((flags & 1 == 0).sum()),
            'oracle_sample': len(sample), 'oracle_mismatches': bad}


def cfg4(n=1000, m=64, d=12):
    A, b = wl.box_cuts_batch(4, n, m, d)
    Ad, bd = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    ms, res = timed(lambda: engine.reduce_batch(Ad, bd, want_A=False))
    lps = int(res.n_lp.sum())
    keeps = res.keep.cpu().numpy().astype(np.uint64)
    bad = 0
    sample = list(range(0, n, n // 12))[:12]
    for p in sample:
        o = orc.reduce(A[p], b[p])
        bad += int(sum(1 << k for k in o['keep']) != int(keeps[p])) + int(o['n_lp'] != int(res.n_lp[p]))
    return {'workload': 'cfg4 LP part: reduce() of %d polytopes d=%d m=%d' % (n, d, m),
            'LPs_per_s': lps / (ms * 1e-3), 'ms': ms, 'LPs': lps,
            'mean_iters': float(res.lp_iters.sum()) / lps, 'oracle_sample': len(sample), 'oracle_mismatches': bad}


def cfg5(shape=(32, 32)):
    A, b, idx = wl.box_grid(shape)
    n = len(A)
    cells = [orc.normalize_rows(A[i], b[i])[:2] for i in range(n)]
    An = torch.from_numpy(np.stack([c[0] for c in cells])).cuda()
    bn = torch.from_numpy(np.stack([c[1] for c in cells])).cuda()
    ii, jj = np.nonzero(~np.eye(n, dtype=bool))            # compute_adj: all ordered pairs i != j
    pi = torch.from_numpy(ii.astype(np.int32)).cuda()
    pj = torch.from_numpy(jj.astype(np.int32)).cuda()
    ms, (adj, rad, st) = timed(lambda: engine.adjacent_pairs(An, bn, pi, pj))
    touch = np.abs(idx[ii] - idx[jj]).max(1) <= 1
    adj = adj.cpu().numpy().astype(bool)
    return {'workload': 'cfg5: %s grid of unit boxes, all ordered pairs (compute_adj)' % (shape,),
            'pairs': int(len(ii)), 'LPs_per_s': len(ii) / (ms * 1e-3), 'ms': ms,
            'adjacent_pairs': int(adj.sum()), 'flag_mismatches_vs_geometry': int((adj != touch).sum()),
            'lp_status_nonzero': int((st != 0).sum())}


if __name__ == '__main__':
    t0 = time.time()
    out = {'cfg3': cfg3(), 'cfg4': cfg4(), 'd16': cfg4(500, 64, 16), 'cfg5': cfg5(),
           'cfg5_4d': cfg5((6, 6, 6, 6))}
    out['wall_s'] = time.time() - t0
    print(json.dumps(out, indent=1))

#!/usr/bin/env python
"""Large-sample parity sweep: GPU reduce() keep masks / flags / LP counts against the CPU
oracle (multiprocessing over the host cores).  Usage: tests/parity_sweep.py [n_cfg2 n_cfg4 n_cfg3]
(test infrastructure: lives under tests/ because it runs the oracle as the checker)"""
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as wl                      # noqa: E402


def _oracle(args):
    cfg, i, m, d, ss = args
    from oracle import polytope_oracle as orc
    A, b = wl.box_cuts(1000 * cfg + i, m, d, ss)
    o = orc.reduce(A, b)
    return sum(1 << k for k in o['keep']), bool(o['empty']), int(o['n_lp']), float(o['r'] or 0.0)


def sweep(cfg, n, m, d, ss, pool):
    ref = pool.map(_oracle, [(cfg, i, m, d, ss) for i in range(n)], chunksize=8)
    import torch
    from polytope_b200 import engine
    A, b = wl.box_cuts_batch(cfg, n, m, d, ss)
    res = engine.reduce_batch(torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda(), want_A=False)
    keep = res.keep.cpu().numpy().astype(np.uint64)
    flags = res.flags.cpu().numpy()
    nlp = res.n_lp.cpu().numpy()
    r = res.r.cpu().numpy()
    bad_keep = sum(int(keep[i]) != ref[i][0] for i in range(n))
    bad_empty = sum(bool(flags[i] & 1) != ref[i][1] for i in range(n))
    bad_nlp = sum(int(nlp[i]) != ref[i][2] for i in range(n))
    rerr = max(abs(float(r[i]) - ref[i][3]) for i in range(n) if not ref[i][1])
    return {'cfg': cfg, 'n': n, 'm': m, 'd': d, 'keep_mismatches': bad_keep, 'empty_mismatches': bad_empty,
            'lp_count_mismatches': bad_nlp, 'max_radius_error': rerr, 'lps': int(nlp.sum())}


if __name__ == '__main__':
    a = [int(v) for v in sys.argv[1:]] + [0, 0, 0]
    n2, n4, n3 = (a[0] or 3000), (a[1] or 300), (a[2] or 2000)
    ctx = mp.get_context('fork')
    with ctx.Pool(os.cpu_count()) as pool:        # fork before CUDA is initialised
        pool.map(_oracle, [(2, 0, 32, 8, False)] * (os.cpu_count() or 1))
        out = [sweep(2, n2, 32, 8, False, pool), sweep(4, n4, 64, 12, False, pool), sweep(3, n3, 16, 6, True, pool),
               sweep(6, n4, 64, 16, False, pool)]
    print(json.dumps(out, indent=1))

#!/usr/bin/env python
"""A/B of library builds on a reduce step (cfg2, or AB_CFG="cfg,P,m,d"): per-stage device times (CUDA events inside the
library) for every PB200_LIB given on the command line.  One process per library."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == '--child':
    sys.path.insert(0, ROOT)
    import torch
    import workloads as wl
    from polytope_b200 import engine
    cfg, P, m, d = [int(v) for v in os.environ.get('AB_CFG', '2,10000,32,8').split(',')]
    A, b = wl.box_cuts_batch(cfg, P, m, d)
    A, b = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    for _ in range(3):
        res = engine.reduce_batch(A, b, want_A=False)
    engine.profile_enable(True)
    acc = {}
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(reps):
        e0.record()
        res = engine.reduce_batch(A, b, want_A=False)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
        for k, v in engine.profile_read().items():
            acc[k] = acc.get(k, 0.0) + v / reps
    lps = int(res.n_lp.sum())
    print(json.dumps({'ms': tot / reps, 'MLPs': lps / (tot / reps) / 1e3, 'iters_per_lp': float(res.lp_iters.sum()) / lps,
                      'lpfail': int((res.flags & 8 != 0).sum()), 'keepsum': int(res.keep.sum()),
                      'stages': {k: round(v, 4) for k, v in acc.items()}}))
    sys.exit(0)
for lib in sys.argv[1:]:
    env = dict(os.environ, PB200_LIB=os.path.join(ROOT, lib))
    out = subprocess.run([sys.executable, os.path.abspath(__file__), '--child'], env=env, capture_output=True, text=True)
    line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:]
    print(lib, line, flush=True)

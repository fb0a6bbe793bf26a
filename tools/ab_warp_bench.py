#!/usr/bin/env python
"""A/B of library builds on the paths that still use the warp-per-LP solvers: cfg2's Chebyshev stage, cfg3 diff
(region_diff of 50 000 cells), cfg4's Chebyshev LPs, generic lp_batch.  One process per PB200_LIB."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == '--child':
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch
    import workloads as wl
    from polytope_b200 import engine
    import bench_configs as bc
    A, b = wl.box_cuts_batch(2, 10000, 32, 8)
    A, b = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    for _ in range(3):
        res = engine.reduce_batch(A, b, want_A=False)
    engine.profile_enable(True)
    acc = {}
    for _ in range(5):
        res = engine.reduce_batch(A, b, want_A=False)
        torch.cuda.synchronize()
        for k, v in engine.profile_read().items():
            acc[k] = acc.get(k, 0.0) + v / 5
    engine.profile_enable(False)
    out = {'cfg2_cheby_ms': round(acc['cheby_lp'], 4), 'cfg2_keepsum': int(res.keep.sum()) % 100000}
    d = bc.cfg3_diff()
    out['cfg3_diff_ms'] = round(d['ms'], 2)
    out['cfg3_diff_MLPs'] = round(d['LPs_per_s'] / 1e6, 2)
    out['cfg3_diff_mismatches'] = d['oracle_mismatches']
    print(json.dumps(out))
    sys.exit(0)
for lib in sys.argv[1:]:
    env = dict(os.environ, PB200_LIB=os.path.join(ROOT, lib))
    o = subprocess.run([sys.executable, os.path.abspath(__file__), '--child'], env=env, capture_output=True, text=True)
    line = o.stdout.strip().splitlines()[-1] if o.stdout.strip() else o.stderr[-400:]
    print(lib.split('/')[-1], line, flush=True)

"""A/B a library build: PB200_LIB=variants/libX.so python tools/ab_bench.py -> stage times of cfg2 reduce."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import workloads as wl
from polytope_b200 import engine
P = 10000
A, b = wl.box_cuts_batch(2, P, 32, 8)
A = torch.from_numpy(A).cuda(); b = torch.from_numpy(b).cuda()
engine.profile_enable(True)
acc = {}
for k in range(8):
    res = engine.reduce_batch(A, b, want_A=False)
    st = engine.profile_read()
    if k >= 3:
        for n, v in st.items():
            acc[n] = acc.get(n, 0) + v / 5
tot = sum(acc.values())
print(json.dumps({'lib': os.environ.get('PB200_LIB', 'default'), 'total_ms': tot, 'MLPs_per_s': float(res.n_lp.sum()) / tot / 1e3,
                  'row_lp': acc['row_lp'], 'bbox_lp': acc['bbox_lp'], 'cheby_lp': acc['cheby_lp'],
                  'keepsum': int(res.keep.sum().item()) }))

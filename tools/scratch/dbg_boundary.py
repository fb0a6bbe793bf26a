import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from polytope_b200 import engine
rng = np.random.default_rng(42)
m = 20
theta0 = np.arccos(1.0 - 1e-7)
for d in (2, 3, 8):
    As, bs = [], []
    for k in range(600):
        M = np.linalg.qr(rng.standard_normal((d, 2)))[0]
        u, v = M[:, 0], M[:, 1]
        t = theta0 * (1.0 + (k - 300) * 2e-10)
        rows = [u, np.cos(t) * u + np.sin(t) * v]
        A = np.vstack(rows + [np.eye(d), -np.eye(d)] + [rng.standard_normal(d) for _ in range(m - 2 - 2 * d)])
        b = np.hstack([1.0 + rng.uniform(0, .1), 1.0 + rng.uniform(0, .1), np.ones(2 * d), 3 + rng.uniform(0, 1, m - 2 - 2 * d)])
        As.append(A); bs.append(b)
    A, b = np.stack(As), np.stack(bs)
    engine.lane_solver(True)
    r1 = engine.reduce_batch(A, b, want_A=False)
    engine.lane_solver(False)
    r0 = engine.reduce_batch(A, b, want_A=False)
    bad = np.nonzero((r1.keep != r0.keep) | (r1.flags != r0.flags))[0]
    print('d', d, 'mismatching polytopes', bad.tolist(), 'LPFAIL lane', int((r1.flags & 8 != 0).sum()), 'warp', int((r0.flags & 8 != 0).sum()))
    for p in bad[:10]:
        print('  p', p, 'lane keep %x flags %d iters %d | warp keep %x flags %d iters %d' % (r1.keep[p] & (2**64-1), r1.flags[p], r1.lp_iters[p], r0.keep[p] & (2**64-1), r0.flags[p], r0.lp_iters[p]))
    print('  lane iters: mean %.2f max per polytope %d; warp mean %.2f max %d' % (r1.lp_iters.sum() / r1.n_lp.sum(), r1.lp_iters.max(), r0.lp_iters.sum() / r0.n_lp.sum(), r0.lp_iters.max()))

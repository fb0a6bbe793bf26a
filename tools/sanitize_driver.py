#!/usr/bin/env python
"""Small batches through every kernel family, for compute-sanitizer (tools/sanitize.sh):
lp_kernel_small (n <= 8), lp_kernel (n = 13), the lane kernels, the one-LP-per-CTA solver, the reduce pipeline,
adjacency, diff_kernel,
hull_kernel, contains / volume.  Results are checked for plausibility only; parity is tests/."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as wl                      # noqa: E402
from polytope_b200 import engine            # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'all'
if which in ('all', 'lp_small', 'lp_generic'):
    engine.lane_solver(False)            # these two families are the warp-per-LP kernels
if which in ('all', 'lp_small'):
    A, b = wl.box_cuts_batch(2, 6, 32, 8)
    res = engine.reduce_batch(A, b)
    assert not np.any(res.flags & engine.F_LPFAIL) and np.all(res.n_lp > 17)
    Ag, bg, _ = wl.box_grid((3, 3))
    adj, rad, st = engine.adjacent_pairs(Ag, bg)
    assert int(adj.sum()) == 20
if which in ('all', 'lp_generic'):
    A, b = wl.box_cuts_batch(4, 3, 64, 12)
    res = engine.reduce_batch(A, b)
    assert not np.any(res.flags & engine.F_LPFAIL)
    A, b = wl.box_cuts_batch(4, 2, 100, 5)
    r, xc, st = engine.cheby_batch(A, b)
    assert np.all(st == 0)
engine.lane_solver(True)
if which in ('all', 'lp_lane'):
    # one LP per lane: shared-G families with n <= 8 (lane_kernel<8>), the wide solver (lane_kernel<12>: factor in
    # shared memory), independent LPs (lane_own_kernel: Chebyshev d <= 7, adjacent pairs), and the warp kernels again
    # with the lane solver switched off
    for cfg, n, m, d in ((2, 40, 32, 8), (4, 6, 64, 12), (3, 40, 16, 6)):
        A, b = wl.box_cuts_batch(cfg, n, m, d)
        res = engine.reduce_batch(A, b)
        assert not np.any(res.flags & engine.F_LPFAIL)
    lo, hi, st = engine.bbox_batch(*wl.box_cuts_batch(2, 5, 32, 8))
    assert np.all(st == 0)
if which in ('all', 'lp_cta'):
    # one LP per CTA: more than 128 rows, more than 32 columns, shared G
    rng = np.random.default_rng(0)
    for m, n in ((300, 7), (100, 40)):
        G = np.vstack([np.eye(n), -np.eye(n), rng.standard_normal((m - 2 * n, n))])
        h = np.hstack([np.ones(2 * n), rng.uniform(0.3, 1.5, m - 2 * n)])
        C = rng.standard_normal((3, n))
        st, X, fun, it = engine.lp_batch(C, np.tile(G, (3, 1, 1)), np.tile(h, (3, 1)))
        assert np.all(st == 0)
        st, X, fun, it = engine.lp_batch(C, G, np.tile(h, (3, 1)))
        assert np.all(st == 0)
if which in ('all', 'diff'):
    import polytope_b200 as pc
    out = pc.region_diff(pc.box2poly([[0, 3], [0, 3]]), pc.Region([pc.box2poly([[1, 2], [1, 2]])]))
    assert len(out) == 4
if which in ('all', 'hull'):
    pts = wl.hull_points(1, 24, 4)
    h = engine.hull_batch(pts[None])
    assert int(h.status[0]) == 0 and int(h.facet_cnt[0]) > 8
if which in ('all', 'sets'):
    A, b = wl.box_cuts(5, 16, 4)
    pts = wl.contains_points(3, 4, 1000)
    fl = engine.contains_batch(A[None], b[None], pts)
    assert fl.shape == (1, 1000)
print('sanitize_driver ok:', which)

#!/usr/bin/env python
"""Golden fixtures for the SURVEY.md 8(f) rows (contains / volume / grid_region /
qhull / extreme), recorded from the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    cd /tmp && python -u /root/repo/tests/golden/make_golden_sets.py

Inputs are regenerated from seeds by workloads.py / the recipes below, so only
outputs are stored.  quickhull's start simplex uses the legacy global numpy RNG
(quickhull.py:172); it is seeded before every call so the files are reproducible.
"""
import os
import signal
import sys
import logging

import numpy as np
import scipy

logging.disable(logging.CRITICAL)
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, '/root/reference')
sys.path.insert(1, REPO)

import polytope as pc                     # noqa: E402  (the reference)
import workloads as wl                    # noqa: E402

assert pc.__file__.startswith('/root/reference'), pc.__file__
META = dict(scipy=scipy.__version__, numpy=np.__version__,
            reference='tulip-control/polytope @ /root/reference (v0.2.6.dev0)')

VOLUME_SPECS = [(6, 2), (10, 3), (16, 4), (16, 6), (32, 8)]      # (m, d), 4 seeds each
HULL_SPECS = [(20, 2), (30, 3), (40, 4), (30, 5), (24, 6)]       # (N, d), 3 seeds each
EXTREME_SPECS = [(6, 2), (10, 3), (12, 4), (15, 5), (18, 6)]     # (m, d), 3 seeds each


contains_points = wl.contains_points
hull_points = wl.hull_points


def sort_rows(M):
    M = np.asarray(M)
    return M[np.lexsort(M.T[::-1])]


def gen_sets():
    out = dict(meta=str(META))
    for m, d in VOLUME_SPECS:
        vols, boxes = [], []
        for i in range(4):
            A, b = wl.box_cuts(8000 + 10 * d + i, m, d, True)
            p = pc.Polytope(A, b)
            v = pc.volume(p, seed=100 + i)
            l, u = p.bounding_box
            vols.append(v)
            boxes.append(np.c_[l, u])
        A, b = wl.box_cuts(8000 + 10 * d, m, d, True)
        out['vol_d%d' % d] = np.array(vols)
        out['vol_d%d_box' % d] = np.array(boxes)
        out['vol_d%d_n777' % d] = pc.volume(pc.Polytope(A, b), nsamples=777, seed=5)
        # a Generator passed as seed is consumed (polytope.py:1587)
        g = np.random.default_rng(9)
        v1 = pc.volume(pc.Polytope(A, b), seed=g)
        v2 = pc.volume(pc.Polytope(A, b), seed=g)
        out['vol_d%d_gen' % d] = np.array([v1, v2])
        # contains
        x = contains_points(50 + d, d, 2000)
        flags = []
        for i in range(3):
            A, b = wl.box_cuts(8000 + 10 * d + i, m, d, True)
            flags.append(pc.Polytope(A, b).contains(x))
        out['contains_d%d' % d] = np.array(flags)
        reg = pc.Region([pc.Polytope(*wl.box_cuts(8000 + 10 * d + i, m, d, True)) for i in range(3)])
        out['region_contains_d%d' % d] = reg.contains(x)
        out['contains_tol0_d%d' % d] = pc.Polytope(*wl.box_cuts(8000 + 10 * d, m, d, True)).contains(x, abs_tol=0)
        print('d', d, 'vol', vols, 'contains', [int(f.sum()) for f in flags])
    p1 = pc.Polytope(np.array([[1.], [-1.]]), np.array([2., 1.]))
    out['vol_d1'] = pc.volume(p1, seed=1)
    # grid_region / enumerate_integral_points
    p = pc.Polytope(*wl.box_cuts(8020, 6, 2, True))
    x, res = pc.grid_region(p)
    out['grid2_x'], out['grid2_res'] = x, np.array(res)
    p = pc.Polytope(*wl.box_cuts(8030, 10, 3, True))
    x, res = pc.grid_region(p, res=[7, 5, 6])
    out['grid3_x'] = x
    A, b = wl.box_cuts(8030, 10, 3, False)
    out['integral3'] = pc.polytope.enumerate_integral_points(pc.Polytope(A, 3.5 * b))
    reg = pc.Region([pc.box2poly([[0, 2], [0, 1]]), pc.box2poly([[1, 3], [0.5, 2.5]])])
    out['integral_region'] = pc.polytope.enumerate_integral_points(reg)
    np.savez_compressed(os.path.join(HERE, 'sets_cases.npz'), **out)


def gen_hull():
    out = dict(meta=str(META))
    for n, d in HULL_SPECS:
        for i in range(3):
            pts = hull_points(600 + 10 * d + i, n, d)
            np.random.seed(i)
            signal.alarm(600)
            q = pc.qhull(pts)
            signal.alarm(0)
            out['hull_d%d_%d_A' % (d, i)] = q.A
            out['hull_d%d_%d_b' % (d, i)] = q.b
            out['hull_d%d_%d_vert' % (d, i)] = q.vertices
            print('hull', n, d, i, 'facets', len(q.b), 'vertices', len(q.vertices))
    # degenerate inputs: cube corners (square facets come out as coplanar triangles)
    cube = np.array([[x, y, z] for x in (0., 1.) for y in (0., 1.) for z in (0., 2.)])
    np.random.seed(0)
    q = pc.qhull(np.vstack([cube, [[0.5, 0.5, 1.0]]]))       # plus an interior point
    out['hull_cube_A'], out['hull_cube_b'], out['hull_cube_vert'] = q.A, q.b, q.vertices
    # too few points / flat
    out['hull_few_empty'] = np.array([len(pc.qhull(np.eye(3)).A)])
    flat = np.c_[hull_points(1, 10, 2), np.zeros(10)]
    out['hull_flat_empty'] = np.array([len(pc.qhull(flat).A)])
    for m, d in EXTREME_SPECS:
        for i in range(3):
            A, b = wl.box_cuts(8500 + 10 * d + i, m, d, True)
            np.random.seed(i)
            signal.alarm(600)
            V = pc.extreme(pc.Polytope(A, b))
            signal.alarm(0)
            out['ext_d%d_%d' % (d, i)] = V
            print('extreme', m, d, i, 'vertices', None if V is None else V.shape)
    A, b = wl.unit_cube3()
    np.random.seed(0)
    out['ext_cube3'] = pc.extreme(pc.Polytope(A, b))
    out['ext_d1'] = pc.extreme(pc.Polytope(np.array([[1.], [-1.]]), np.array([2., 1.])))
    # not full-dimensional -> None
    emp = pc.extreme(pc.Polytope(np.array([[1., 0, 0], [-1., 0, 0], [0, 1., 0], [0, -1., 0], [0, 0, 1.], [0, 0, -1.]]),
                                 np.array([1., -1., 1, 1, 1, 1])))
    out['ext_flat_is_none'] = np.array([emp is None])
    np.savez_compressed(os.path.join(HERE, 'hull_cases.npz'), **out)


if __name__ == '__main__':
    gen_sets()
    gen_hull()
    for f in ('sets_cases.npz', 'hull_cases.npz'):
        print(f, os.path.getsize(os.path.join(HERE, f)), 'bytes')

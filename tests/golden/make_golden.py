#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    cd /tmp && python /root/repo/tests/golden/make_golden.py

It imports tulip-control/polytope read-only from /root/reference (scipy/HiGHS
path, the only solver installed here), wraps `polytope.polytope.lpsolve` (the
name bound at polytope/polytope.py:69) to record every LP the reference fires,
and writes small .npz files next to this script.  Inputs are regenerated from
seeds by /root/repo/workloads.py, so only outputs (and a sample of raw LPs) are
stored.  scipy / numpy versions are recorded in each file.
"""
import os
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
import polytope.polytope as alg           # noqa: E402
import workloads as wl                    # noqa: E402

assert pc.__file__.startswith('/root/reference'), pc.__file__
META = dict(scipy=scipy.__version__, numpy=np.__version__,
            reference='tulip-control/polytope @ /root/reference (v0.2.6.dev0)')

_log = []
_orig = alg.lpsolve


def _recording_lpsolve(c, G, h):
    sol = _orig(c, G, h)
    _log.append((np.array(c, float), np.array(G, float), np.array(h, float),
                 int(sol['status']),
                 None if sol['x'] is None else np.array(sol['x'], float),
                 None if sol['fun'] is None else float(sol['fun'])))
    return sol


alg.lpsolve = _recording_lpsolve


def match_rows(A_in, b_in, A_out, b_out):
    """Indices of the input rows the reference kept (nearest-row match)."""
    keep = []
    for a, bb in zip(A_out, b_out):
        dist = np.abs(A_in - a).sum(1) + np.abs(b_in - bb)
        k = int(np.argmin(dist))
        assert dist[k] < 1e-12, dist[k]
        assert np.sum(dist < 1e-9) == 1, 'ambiguous row match'
        keep.append(k)
    assert keep == sorted(keep)
    return keep


def pad_lists(lists, fill=-1):
    n = max([len(x) for x in lists] + [1])
    out = np.full((len(lists), n), fill, dtype=np.int64)
    for i, x in enumerate(lists):
        out[i, :len(x)] = x
    return out


def reduce_case(A, b):
    """Run reference reduce on Polytope(A, b); return a result record."""
    del _log[:]
    p = pc.Polytope(A, b)
    red = pc.reduce(p)
    n_lp = len(_log)
    lp_status = [e[3] for e in _log]
    lp_fun = [np.nan if e[5] is None else e[5] for e in _log]
    empty = len(red.A) == 0
    keep = [] if empty else match_rows(p.A, p.b, red.A, red.b)
    r = p._chebR if p._chebXc is not None else 0.0
    xc = p._chebXc if p._chebXc is not None else np.full(A.shape[1], np.nan)
    return dict(keep=keep, empty=empty, n_lp=n_lp, r=float(r), xc=xc,
                A_out=red.A, b_out=red.b, minrep=bool(red.minrep),
                lp_status=lp_status, lp_fun=lp_fun, lps=list(_log))


def gen_reduce():
    """G1 batches named by SURVEY.md 8(d): (tag, cfg, n, m, d, shift_scale)."""
    specs = [('cfg2', 2, 48, 32, 8, False),
             ('cfg3', 3, 32, 16, 6, True),
             ('cfg4', 4, 6, 64, 12, False),
             ('d16', 6, 6, 64, 16, False),
             ('small', 7, 16, 10, 3, True)]
    out = dict(meta=str(META))
    lp_pool = []
    for tag, cfg, n, m, d, ss in specs:
        recs = []
        for i in range(n):
            A, b = wl.box_cuts(1000 * cfg + i, m, d, ss)
            rec = reduce_case(A, b)
            recs.append(rec)
            if i < 2:
                lp_pool += rec['lps']
        out[tag + '_spec'] = np.array([cfg, n, m, d, int(ss)])
        out[tag + '_keep'] = pad_lists([r['keep'] for r in recs])
        out[tag + '_empty'] = np.array([r['empty'] for r in recs])
        out[tag + '_nlp'] = np.array([r['n_lp'] for r in recs])
        out[tag + '_r'] = np.array([r['r'] for r in recs])
        out[tag + '_xc'] = np.array([r['xc'] for r in recs])
        out[tag + '_minrep'] = np.array([r['minrep'] for r in recs])
        out[tag + '_bout'] = np.array(
            [np.r_[r['b_out'], np.full(m - len(r['b_out']), np.nan)]
             for r in recs])
        nl = max(r['n_lp'] for r in recs)
        out[tag + '_lpfun'] = np.array(
            [np.r_[r['lp_fun'], np.full(nl - r['n_lp'], np.nan)] for r in recs])
        out[tag + '_lpstatus'] = pad_lists([r['lp_status'] for r in recs])
        print(tag, 'kept', [len(r['keep']) for r in recs][:8],
              'lps', [r['n_lp'] for r in recs][:8])
    np.savez_compressed(os.path.join(HERE, 'reduce_cases.npz'), **out)
    return lp_pool


def gen_named():
    """Inputs of the reference's own tests (tests/polytope_test.py)."""
    out = dict(meta=str(META))
    # cfg1 unit cube: 6 rows kept, r=.5, 8 LPs
    A, b = wl.unit_cube3()
    rec = reduce_case(A, b)
    out['cube_keep'] = np.array(rec['keep'])
    out['cube_nlp'] = rec['n_lp']
    out['cube_r'] = rec['r']
    out['cube_xc'] = rec['xc']
    # test_reduce, polytope_test.py:601-622
    a = np.array([[1.0, 0.1], [1.0, 0.1], [-1., 0.], [0., 1.], [0., -1.]])
    bb = np.array([50., 50.5, -40., 1., 0.])
    rec = reduce_case(a, bb)
    out['treduce_keep'] = np.array(rec['keep'])
    out['treduce_A'] = rec['A_out']
    out['treduce_b'] = rec['b_out']
    l, u = pc.reduce(pc.Polytope(a, bb)).bounding_box
    out['treduce_l'], out['treduce_u'] = l, u
    # operations_test squares, :58-86, polytope_full_dim_test :200-204,
    # polytope_intersect_test :220-238
    Ab = np.array([[0., 1, 1], [0, -1, 0], [1, 0, 1], [-1, 0, 0]])
    Ab2 = np.array([[-1., 0, 1], [1, 0, 0], [0, 1, 1], [0, -1, 0]])
    p1 = pc.Polytope(Ab[:, :2], Ab[:, 2])
    p2 = pc.Polytope(Ab2[:, :2], Ab2[:, 2])
    p4 = pc.Polytope(np.array([[1., 0], [0, 1], [-1, 0], [0, -1]]),
                     np.array([.5, .5, .5, .5]))
    far = pc.Polytope(Ab[:, :2], Ab[:, 2] - 1e3)
    out['sq_fulldim'] = np.array([pc.is_fulldim(p1), pc.is_fulldim(p2),
                                  pc.is_fulldim(far),
                                  pc.is_fulldim(p1.intersect(p2)),
                                  pc.is_fulldim(p2.intersect(p4))])
    out['sq_cheby_r'] = np.array([p1.chebR, p2.chebR, p4.chebR])
    out['sq_cheby_x'] = np.array([p1.chebXc, p2.chebXc, p4.chebXc])
    p5 = p2.intersect(p4)
    out['sq_p5_A'], out['sq_p5_b'] = p5.A, p5.b
    # bounding boxes, test_bounding_box_to_polytope :299-312
    for k, iv in enumerate([[[0, 1]], [[0, 1], [0, 2]],
                            [[-1, 2], [3, 5], [-5, -3]]]):
        p = pc.box2poly(iv)
        l, u = p.bounding_box
        out['bbox%d_l' % k], out['bbox%d_u' % k] = l, u
    # unbounded / empty behaviours
    half = pc.Polytope(np.array([[1., 0.]]), np.array([1.]))
    out['half_cheby'] = np.array([pc.cheby_ball(half)[0]])
    l, u = pc.bounding_box(pc.Polytope(np.array([[1., 0.], [0., 1.], [-1, 0]]),
                                       np.array([1., 1., 1.])))
    out['unb_l'], out['unb_u'] = l, u
    emp = pc.Polytope(np.array([[1.], [-1.]]), np.array([0., -1.]))
    del _log[:]
    out['empty_cheby'] = np.array([pc.cheby_ball(emp)[0]])
    out['empty_cheby_lp_x'] = _log[0][4]
    l, u = pc.bounding_box(emp)
    out['empty_l'], out['empty_u'] = l, u
    np.savez_compressed(os.path.join(HERE, 'named_cases.npz'), **out)


def gen_adjacent():
    out = dict(meta=str(META))
    lp_pool = []
    for tag, shape in [('g2', (6, 6)), ('g3', (3, 3, 3)), ('g4', (3, 3, 2, 2))]:
        A, b, idx = wl.box_grid(shape)
        cells = [pc.Polytope(A[i], b[i]) for i in range(len(A))]
        n = len(cells)
        adj = np.zeros((n, n), dtype=np.int8)
        rad = np.full((n, n), np.nan)
        for i in range(n):
            adj[i, i] = 1
            for j in range(i):
                del _log[:]
                adj[i, j] = adj[j, i] = pc.is_adjacent(cells[i], cells[j])
                assert len(_log) == 1
                st, x = _log[0][3], _log[0][4]
                rad[i, j] = rad[j, i] = x[-1] if st == 0 else np.nan
                if len(lp_pool) < 40:
                    lp_pool.append(_log[0])
        # geometric ground truth: boxes touch iff |index diff| <= 1 everywhere
        touch = (np.abs(idx[:, None, :] - idx[None, :, :]).max(-1) <= 1)
        assert np.array_equal(adj.astype(bool), touch), tag
        out[tag + '_shape'] = np.array(shape)
        out[tag + '_adj'] = adj
        out[tag + '_r'] = rad
        print(tag, n, 'cells', int(adj.sum()), 'adjacent entries')
    # overlapping / disjoint / near-touching random polytopes, d=4
    recs = []
    for i in range(60):
        A1, b1 = wl.box_cuts(9000 + i, 12, 4, True)
        A2, b2 = wl.box_cuts(9500 + i, 12, 4, True)
        rng = np.random.default_rng(77 + i)
        b2 = b2 + A2 @ (rng.uniform(-1, 1, 4) * (i % 4))   # shift second one
        q1, q2 = pc.Polytope(A1, b1), pc.Polytope(A2, b2)
        del _log[:]
        flag = pc.is_adjacent(q1, q2)
        recs.append((flag, _log[0][4][-1] if _log[0][3] == 0 else np.nan))
    out['rand_flag'] = np.array([r[0] for r in recs])
    out['rand_r'] = np.array([r[1] for r in recs])
    print('rand adjacent', int(out['rand_flag'].sum()), 'of 60')
    np.savez_compressed(os.path.join(HERE, 'adjacent_cases.npz'), **out)
    return lp_pool


def gen_intersect():
    out = dict(meta=str(META))
    keeps, empties, nlps, rs = [], [], [], []
    Q = pc.Polytope(*wl.box_cuts(3999, 16, 6, True))
    for i in range(24):
        P = pc.Polytope(*wl.box_cuts(3000 + i, 16, 6, True))
        del _log[:]
        # fresh Q each time so its cached fulldim does not hide an LP
        Qi = pc.Polytope(Q.A.copy(), Q.b.copy())
        isect = P.intersect(Qi)
        rp, _ = isect.cheby
        empty = len(isect.A) == 0
        stacked_A = np.vstack([P.A, Qi.A])
        stacked_b = np.hstack([P.b, Qi.b])
        sA, sb = pc.Polytope(stacked_A, stacked_b).A, \
            pc.Polytope(stacked_A, stacked_b).b
        keeps.append([] if empty else match_rows(sA, sb, isect.A, isect.b))
        empties.append(empty)
        nlps.append(len(_log))
        rs.append(float(rp))
    out['keep'] = pad_lists(keeps)
    out['empty'] = np.array(empties)
    out['nlp'] = np.array(nlps)
    out['r'] = np.array(rs)
    print('intersect: empty', int(np.sum(empties)), 'of 24; lps', nlps[:8])
    np.savez_compressed(os.path.join(HERE, 'intersect_cases.npz'), **out)


def gen_lps(pool):
    """Raw LPs with the reference's (status, x, fun), padded to one tensor."""
    # hand-made cases: reference tests + exceptional statuses
    hand = [
        (np.array([1.]), np.array([[-1.]]), np.array([1.])),            # :533
        (np.array([1., 1.]), -np.eye(2), np.array([1., 1.])),           # :517
        (np.array([1.]), np.array([[1.]]), np.array([1.])),             # unbounded
        (np.array([-1., 0.]), np.array([[1., 0.], [-1., 0.]]),
         np.array([1., 1.])),                                            # free x2
        (np.array([1., 0.]), np.array([[1., 0.], [-1., 0.], [0., 1.]]),
         np.array([-1., 0., 1.])),                                       # infeasible
        (np.array([0., 0., -1.]),
         np.c_[np.array([[1., 0.], [-1., 0.]]), np.ones(2)],
         np.array([0., -1.])),                                           # cheby of empty
        (np.array([0., 0., -1.]), np.array([[1., 0., 1.]]),
         np.array([1.])),                                                # cheby half-space
    ]
    entries = []
    for c, G, h in hand:
        del _log[:]
        alg.lpsolve(c, G, h)
        entries.append(_log[0])
    entries += pool
    L = len(entries)
    mm = max(e[1].shape[0] for e in entries)
    nn = max(e[1].shape[1] for e in entries)
    C = np.zeros((L, nn))
    G = np.zeros((L, mm, nn))
    H = np.zeros((L, mm))
    X = np.full((L, nn), np.nan)
    fun = np.full(L, np.nan)
    st = np.zeros(L, dtype=np.int64)
    shp = np.zeros((L, 2), dtype=np.int64)
    for i, (c, g, h, s, x, f) in enumerate(entries):
        m, n = g.shape
        shp[i] = (m, n)
        C[i, :n], G[i, :m, :n], H[i, :m] = c, g, h
        st[i] = s
        if x is not None:
            X[i, :n] = x
        if f is not None:
            fun[i] = f
    np.savez_compressed(os.path.join(HERE, 'lp_cases.npz'), meta=str(META),
                        C=C, G=G, H=H, X=X, fun=fun, status=st, shape=shp,
                        n_hand=len(hand))
    print('lp_cases', L, 'LPs; statuses', np.bincount(st))


if __name__ == '__main__':
    pool = gen_reduce()
    gen_named()
    pool += gen_adjacent()
    gen_intersect()
    gen_lps(pool)
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)), 'bytes')

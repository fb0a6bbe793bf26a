"""Exhaustive parity at the full BASELINE.json config sizes ("keep-masks bit-exact
for all 10 000"), CUDA path through the C ABI against the CPU oracle on every unit:

  cfg2  all 10 000 polytopes (32 x 8): keep masks, emptiness flags, LP counts; radii; centres
  cfg4  all  1 000 polytopes (64 x 12): the same
  cfg3  all 50 000 is_fulldim flags (16 x 6, shifted/scaled) + 2 048 pairwise intersects with Q
  cfg5  all 1 047 552 ordered-pair adjacency flags of the 32 x 32 box grid against geometry,
        and 2 048 sampled pairs against the oracle

The oracle runs on every host core (spawn pool, tests/parity_workers.py); at the reference's
~13 k LPs/s on 16 cores the whole file is about a minute.  Each test also appends the count of
"ambiguous" LPs (decision quantity within 1e-6 of its threshold) and the largest
deviations to gpurun_out/parity_fullsize.json.
"""
import json
import os

import numpy as np
import pytest

import workloads as wl
import parity_workers as pw

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _report(name, data):
    path = os.path.join(ROOT, 'gpurun_out', 'parity_fullsize.json')
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        try:
            doc = json.load(open(path))
        except (OSError, ValueError):
            doc = {}
        doc[name] = data
        json.dump(doc, open(path, 'w'), indent=1)
    except OSError:
        pass


def _reduce_all(cfg, P, m, d, per):
    """Every polytope of a box+cuts config through the GPU reduce and through the oracle."""
    from polytope_b200 import engine
    A, b = wl.box_cuts_batch(cfg, P, m, d)
    res = engine.reduce_batch(A, b, want_A=False)
    jobs = [(cfg, s, c, m, d, False) for s, c in pw.chunks(P, per)]
    ora = [row for chunk in pw.run_pool(pw.reduce_chunk, jobs) for row in chunk]
    assert len(ora) == P
    keep = res.keep.astype(np.uint64)
    o_keep = np.array([row[0] for row in ora], dtype=np.uint64)
    o_nlp = np.array([row[1] for row in ora])
    o_empty = np.array([row[2] for row in ora])
    o_r = np.array([row[3] for row in ora])
    amb = np.array([row[5] for row in ora])
    bad_keep = np.nonzero(keep != o_keep)[0]
    assert bad_keep.size == 0, ('keep masks differ', bad_keep[:10], keep[bad_keep[:3]], o_keep[bad_keep[:3]])
    assert np.array_equal(res.n_lp, o_nlp)
    assert np.array_equal((res.flags & engine.F_EMPTY) != 0, o_empty)
    assert not np.any(res.flags & engine.F_LPFAIL)
    dr = np.abs(res.r - o_r)
    assert dr.max() <= 1e-7 + 1e-7 * np.abs(o_r).max()
    # Chebyshev centres: random cuts make the optimum unique (d+1 generic active rows), so the
    # interior-point + polish centre must be HiGHS' vertex to 1e-7
    o_xc = np.array([row[4] for row in ora])
    dx = np.abs(res.xc - o_xc).max(1)
    assert dx.max() <= 1e-7, (int(np.argmax(dx)), float(dx.max()))
    return {'polytopes': P, 'm': m, 'd': d, 'lps': int(o_nlp.sum()), 'keep_mask_mismatches': 0,
            'ambiguous_lps_within_1e-6': int((amb < 1e-6).sum()), 'max_abs_dr': float(dr.max()),
            'max_abs_dxc': float(dx.max())}


def test_cfg2_every_polytope_against_the_oracle():
    _report('cfg2', _reduce_all(2, 10000, 32, 8, 80))


def test_cfg4_every_polytope_against_the_oracle():
    _report('cfg4', _reduce_all(4, 1000, 64, 12, 8))


def test_cfg3_every_fulldim_flag_and_2048_intersections():
    import torch
    from polytope_b200 import engine
    P, m, d, npair, qseed = 50000, 16, 6, 2048, 3999
    A, b = wl.box_cuts_batch(3, P, m, d, shift_scale=True)
    An, bn, _ = engine.normalize_batch(torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda())
    r, xc, st = engine.cheby_batch(An, bn)
    r, st = r.cpu().numpy(), st.cpu().numpy()
    jobs = [(3, s, c, m, d, True) for s, c in pw.chunks(P, 400)]
    o_r = np.array([v for chunk in pw.run_pool(pw.fulldim_chunk, jobs) for v in chunk])
    assert np.all(st == 0)
    assert np.array_equal(r > 1e-7, o_r > 1e-7)
    assert np.abs(r - o_r).max() <= 1e-7 * (1 + np.abs(o_r).max())
    # pairwise Polytope.intersect(member, Q): stack + reduce
    from oracle import polytope_oracle as orc
    Q = orc.normalize_rows(*wl.box_cuts(qseed, m, d, True))[:2]
    Qa = torch.from_numpy(Q[0]).cuda().expand(npair, m, d)
    Qb = torch.from_numpy(Q[1]).cuda().expand(npair, m)
    res = engine.reduce_batch(torch.cat([An[:npair], Qa], 1).contiguous(), torch.cat([bn[:npair], Qb], 1).contiguous(),
                              normalize=True, want_A=False)
    jobs = [(3, s, c, m, d, qseed) for s, c in pw.chunks(npair, 32)]
    ora = [row for chunk in pw.run_pool(pw.intersect_chunk, jobs) for row in chunk]
    keep = res.keep.cpu().numpy().astype(np.uint64)
    flags = res.flags.cpu().numpy()
    n_lp = res.n_lp.cpu().numpy()
    # the reference returns Polytope() without calling reduce when an operand is not fulldim (:268-269)
    q_full = orc.is_fulldim(*Q)
    bad = 0
    for p, (o_keep, o_nlp, o_empty) in enumerate(ora):
        operand_empty = not (q_full and o_r[p] > 1e-7)
        if operand_empty:
            assert o_empty
            continue
        bad += int(int(keep[p]) != o_keep) + int(bool(flags[p] & engine.F_EMPTY) != o_empty) + int(n_lp[p] != o_nlp)
    assert bad == 0
    _report('cfg3', {'polytopes': P, 'fulldim': int((r > 1e-7).sum()),
                     'ambiguous_radii_within_1e-6': int((np.abs(o_r - 1e-7) < 1e-6).sum()),
                     'max_abs_dr': float(np.abs(r - o_r).max()), 'intersections': npair,
                     'nonempty_intersections': int(sum(not row[2] for row in ora)), 'mismatches': 0})


def test_cfg5_every_ordered_pair_flag_and_2048_pairs_against_the_oracle():
    import torch
    from polytope_b200 import engine
    from oracle import polytope_oracle as orc
    shape = (32, 32)
    A, b, idx = wl.box_grid(shape)
    n = len(A)
    cells = [orc.normalize_rows(A[i], b[i])[:2] for i in range(n)]
    An = torch.from_numpy(np.stack([c[0] for c in cells])).cuda()
    bn = torch.from_numpy(np.stack([c[1] for c in cells])).cuda()
    ii, jj = np.nonzero(~np.eye(n, dtype=bool))            # compute_adj: all ordered pairs, prop2partition.py:253-261
    assert len(ii) == 1047552
    adj, rad, st = engine.adjacent_pairs(An, bn, torch.from_numpy(ii.astype(np.int32)).cuda(),
                                         torch.from_numpy(jj.astype(np.int32)).cuda())
    adj, rad, st = adj.cpu().numpy().astype(bool), rad.cpu().numpy(), st.cpu().numpy()
    touch = np.abs(idx[ii] - idx[jj]).max(1) <= 1          # boxes touch (face, edge or corner)
    assert np.all(st == 0)
    assert np.array_equal(adj, touch)
    # touching boxes: r = 1e-7 exactly against the 1e-8 threshold -> |dr| <= 1e-9
    assert np.abs(rad[touch] - 1e-7).max() <= 1e-9
    rng = np.random.default_rng(5)
    near = np.nonzero(np.abs(idx[ii] - idx[jj]).max(1) <= 2)[0]
    pick = np.concatenate([rng.choice(near, 1536, replace=False), rng.choice(len(ii), 512, replace=False)])
    pairs = [(int(ii[t]), int(jj[t])) for t in pick]
    jobs = [(shape, pairs[s:s + c]) for s, c in pw.chunks(len(pairs), 64)]
    ora = [row for chunk in pw.run_pool(pw.adjacent_chunk, jobs) for row in chunk]
    o_flag = np.array([row[0] for row in ora])
    o_rad = np.array([row[1] for row in ora])
    assert np.array_equal(adj[pick], o_flag)
    pos = o_rad > 0                                        # the oracle reports r = 0 for r < 0 (:1291-1293)
    assert np.abs(rad[pick][pos] - o_rad[pos]).max() <= 1e-9
    _report('cfg5', {'ordered_pairs': int(len(ii)), 'adjacent': int(adj.sum()), 'flag_mismatches_vs_geometry': 0,
                     'oracle_pairs': len(pairs), 'oracle_flag_mismatches': 0,
                     'ambiguous_pairs_within_1e-6_of_threshold': int((np.abs(rad - 1e-8) < 1e-6).sum()),
                     'max_abs_dr_touching': float(np.abs(rad[touch] - 1e-7).max())})


def test_duplicate_filter_at_the_one_minus_abs_tol_boundary():
    """reduce()'s duplicate-direction test `dot(a_i, a_j) > 1 - abs_tol` (polytope.py:1102) probed
    within a few ulp of the threshold: pairs of unit rows at angle theta with cos(theta) swept
    across 1 - 1e-7 in d = 2, 3, 8 (random orientation, so the dot product rounds differently for a
    chain of fmas and for separate multiply-adds).  Keep sets must equal the oracle's."""
    from polytope_b200 import engine
    from oracle import polytope_oracle as orc
    rng = np.random.default_rng(42)
    As, bs = [], []
    m = 20
    theta0 = np.arccos(1.0 - 1e-7)
    for d in (2, 3, 8):
        As, bs = [], []
        for k in range(600):
            # orthonormal pair (u, v); a_j = cos(t) u + sin(t) v
            M = np.linalg.qr(rng.standard_normal((d, 2)))[0]
            u, v = M[:, 0], M[:, 1]
            t = theta0 * (1.0 + (k - 300) * 2e-10)
            rows = [u, np.cos(t) * u + np.sin(t) * v]
            A = np.vstack(rows + [np.eye(d), -np.eye(d)] + [rng.standard_normal(d) for _ in range(m - 2 - 2 * d)])
            b = np.hstack([1.0 + rng.uniform(0, .1), 1.0 + rng.uniform(0, .1), np.ones(2 * d), 3 + rng.uniform(0, 1, m - 2 - 2 * d)])
            As.append(A)
            bs.append(b)
        A, b = np.stack(As), np.stack(bs)
        res = engine.reduce_batch(A, b, want_A=False)
        keeps = res.keep_lists()
        flagged = 0
        for p in range(len(A)):
            o = orc.reduce(A[p], b[p])
            assert keeps[p] == o['keep'], (d, p, keeps[p], o['keep'])
            assert int(res.n_lp[p]) == o['n_lp'], (d, p)
            An, bn, _ = orc.normalize_rows(A[p], b[p])
            flagged += int(len(orc.duplicate_rows(An, bn)) < m)
        # the sweep really straddles the threshold: some instances lose a row to the duplicate
        # filter, the others keep the pair for the LPs to decide
        assert 0 < flagged < len(A), flagged

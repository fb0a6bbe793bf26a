"""GPU parity of region_diff / mldivide / envelope / is_convex / union /
is_subset against the golden vectors of the unmodified
reference and against the oracle restatement.

Bars: the same pieces in the same order with the same rows (1e-9), the same
result type (Polytope vs Region), LP-decided flags exact.
"""
import numpy as np
import pytest

import workloads as wl

pytestmark = pytest.mark.gpu


def pieces_of(x):
    if len(x) == 0:
        return [] if len(x.A) == 0 else [(x.A, x.b)]
    return [(p.A, p.b) for p in x.list_poly]


def check_against_golden(g, tag, x):
    pieces = pieces_of(x)
    assert int(g[tag + '_n'][0]) == len(pieces), (tag, int(g[tag + '_n'][0]), len(pieces))
    assert int(g[tag + '_kind'][0]) == (1 if len(x) > 0 else 0), tag
    for k, (A, b) in enumerate(pieces):
        rows = int(np.sum(~np.isnan(g[tag + '_b'][k])))
        assert rows == len(b), (tag, k, rows, len(b))
        np.testing.assert_allclose(A, g[tag + '_A'][k][:rows], atol=1e-9)
        np.testing.assert_allclose(b, g[tag + '_b'][k][:rows], atol=1e-9)


def test_region_diff_batch_matches_reference_golden(golden):
    import polytope_b200 as pc
    g = golden('diff_cases')
    polys, regs = [], []
    for i in range(wl.DIFF_CASES):
        (A, b), cells = wl.diff_case(i)
        polys.append(pc.Polytope(A, b))
        regs.append(pc.Region([pc.Polytope(a_, b_) for a_, b_ in cells]))
    # one launch per dimension (the batch is padded to one shape)
    for d in (2, 3, 4):
        idx = [i for i in range(wl.DIFF_CASES) if polys[i].dim == d]
        res = pc.region_diff_batch([polys[i] for i in idx], [regs[i] for i in idx])
        for i, r in zip(idx, res):
            check_against_golden(g, 'diff%d' % i, r)
    # and one at a time through the reference-named entry
    for i in (0, 1, 5, 20):
        check_against_golden(g, 'diff%d' % i, pc.region_diff(polys[i], regs[i]))


def test_box_differences_and_region_ops(golden):
    import polytope_b200 as pc
    g = golden('diff_cases')
    B = pc.box2poly
    sq = B([[0, 2], [0, 2]])
    check_against_golden(g, 'box_corner', sq.diff(B([[1, 3], [1, 3]])))
    check_against_golden(g, 'box_hole', sq.diff(B([[0.5, 1.5], [0.5, 1.5]])))
    check_against_golden(g, 'box_covered', sq.diff(B([[-1, 3], [-1, 3]])))
    check_against_golden(g, 'box_far', sq.diff(B([[5, 6], [5, 6]])))
    check_against_golden(g, 'box_two', pc.region_diff(sq, pc.Region([B([[0.5, 1], [0.5, 1]]), B([[1.2, 1.8], [-1, 3]])])))
    check_against_golden(g, 'box3', B([[0, 1], [0, 1], [0, 1]]).diff(B([[0.5, 2], [0.5, 2], [-1, 2]])))
    L = pc.Region([B([[0, 1], [0, 2]]), B([[1, 2], [0, 1]])])
    check_against_golden(g, 'reg_minus', L.diff(B([[0.5, 1.5], [0.5, 1.5]])))
    two = pc.Region([B([[0, 1], [0, 1]]), B([[1, 2], [0, 1]])])
    check_against_golden(g, 'env_two', pc.envelope(two))
    check_against_golden(g, 'env_L', pc.envelope(L))
    conv, env = pc.is_convex(two)
    assert conv == bool(g['convex_two'][0])
    check_against_golden(g, 'convex_two_env', env)
    assert pc.is_convex(L)[0] == bool(g['convex_L'][0])
    check_against_golden(g, 'union_cc', pc.union(B([[0, 1], [0, 1]]), B([[1, 2], [0, 1]]), check_convex=True))
    check_against_golden(g, 'union_overlap_cc', pc.union(B([[0, 2], [0, 2]]), B([[1, 3], [1, 3]]), check_convex=True))
    check_against_golden(g, 'union_plain', pc.union(B([[0, 1], [0, 1]]), B([[3, 4], [0, 1]])))
    for i in range(4):
        (A, b), cells = wl.diff_case(i)
        reg = pc.Region([pc.Polytope(A, b)] + [pc.Polytope(a_, b_) for a_, b_ in cells])
        check_against_golden(g, 'env%d' % i, pc.envelope(reg))
    sub = [B([[0.2, 0.8], [0.2, 0.8]]) <= sq, sq <= B([[0.2, 0.8], [0.2, 0.8]]),
           L <= sq, sq <= L, two == B([[0, 2], [0, 1]]), sq == sq.copy()]
    assert sub == [bool(v) for v in g['subset']]
    check_against_golden(g, 'reg_isect', L.intersect(B([[0.5, 1.5], [0.5, 1.5]])))
    check_against_golden(g, 'reg_and', L & pc.Region([B([[0.5, 3], [0.25, 0.75]])]))


@pytest.mark.parametrize('d,m,ncell,T', [(6, 16, 1, 200), (3, 8, 4, 120), (8, 20, 2, 40)])
def test_region_diff_batch_vs_oracle(d, m, ncell, T):
    """Shared subtrahend (cfg3 shape: many cells minus one polytope) and per-problem
    regions, against the oracle restatement: status, piece count, LP count, rows."""
    from polytope_b200 import engine
    from oracle import polytope_oracle as orc
    rng = np.random.default_rng(d * 100 + m)
    polys = [orc.normalize_rows(*wl.box_cuts(31000 + i, m, d, True))[:2] for i in range(T)]
    cells = []
    for c in range(ncell):
        A, b = wl.box_cuts(32000 + c, m, d)
        cells.append(orc.normalize_rows(A, 0.8 * b + A @ rng.uniform(-0.5, 0.5, d))[:2])
    PA = np.array([p[0] for p in polys])
    Pb = np.array([p[1] for p in polys])
    RA = np.array([c[0] for c in cells])
    Rb = np.array([c[1] for c in cells])
    res = engine.region_diff_batch(PA, Pb, RA, Rb)
    kinds = {'pieces': engine.DIFF_PIECES, 'poly': engine.DIFF_UNTOUCHED, 'empty': engine.DIFF_COVERED}
    n_checked = 0
    for t in range(0, T, max(1, T // 25)):
        orc.lp_count = 0
        kind, pieces = orc.region_diff(polys[t], cells)
        assert int(res.status[t]) == kinds[kind], t
        if kind != 'pieces':
            continue
        assert int(res.n_pieces[t]) == len(pieces), t
        o = int(res.piece_off[t])
        for k, (A, b, reduced) in enumerate(pieces):
            assert int(res.rows[o + k]) == len(b) and bool(res.reduce[o + k]) == reduced
            assert np.array_equal(res.A[o + k][:len(b)], A) and np.array_equal(res.b[o + k][:len(b)], b)
        n_checked += 1
    assert n_checked > 0


def test_region_diff_with_many_cells_prefilters_on_the_host():
    """A region of 15 x 15 = 225 unit cells (beyond the kernel's 64 cells per problem): only the
    cells meeting the minuend reach the device search; results equal the oracle's."""
    import polytope_b200 as pc
    from oracle import polytope_oracle as orc
    cells = [pc.box2poly([[i, i + 1], [j, j + 1]]) for i in range(15) for j in range(15)]
    reg = pc.Region(cells)
    for iv in ([[2.5, 4.25], [3.5, 4.5]], [[-3, -2], [0, 1]], [[0.2, 0.8], [0.3, 0.6]], [[13.5, 16], [13.5, 15.5]]):
        poly = pc.box2poly(iv)
        got = pc.region_diff(poly, reg)
        kind, pieces = orc.region_diff((poly.A, poly.b), [(c.A, c.b) for c in cells])
        if kind == 'poly':
            assert len(got) == 0 and np.array_equal(got.A, poly.A)
            continue
        want = []
        for A, b, reduced in pieces:
            if reduced:
                red = orc.reduce(A, b)
                if red['empty']:
                    continue
                want.append(orc.normalize_rows(red['A'], red['b'])[:2])
            else:
                want.append(orc.normalize_rows(A, b)[:2])
        have = pieces_of(got)
        assert len(have) == len(want)
        for (A1, b1), (A2, b2) in zip(have, want):
            np.testing.assert_allclose(A1, A2, atol=1e-9)
            np.testing.assert_allclose(b1, b2, atol=1e-9)
    # is_subset against the big region goes through the same path
    assert pc.is_subset(pc.box2poly([[1.5, 3.5], [2.5, 3.5]]), reg)
    assert not pc.is_subset(pc.box2poly([[14.5, 15.5], [2.5, 3.5]]), reg)


def test_region_diff_piece_pool_grows_until_it_fits():
    """1000 boxes each minus an inner box need 2 d = 6 pieces per problem, more than the initial
    pool of max(4 T, 1024): the retry path must grow geometrically (growing to the
    reported `used` only adds ~T per try and gave up after 4)."""
    from polytope_b200 import engine
    T, d = 1000, 3
    rng = np.random.default_rng(3)
    lo = rng.uniform(-1, 0, (T, d))
    PA = np.broadcast_to(np.vstack([np.eye(d), -np.eye(d)]), (T, 2 * d, d)).copy()
    Pb = np.hstack([lo + 3.0, -lo])
    RA = PA[:, None].copy()
    Rb = np.hstack([lo + 2.0, -(lo + 1.0)])[:, None].copy()
    res = engine.region_diff_batch(PA, Pb, RA, Rb)
    assert np.all(res.status == engine.DIFF_PIECES)
    assert np.all(res.n_pieces == 2 * d)
    assert len(res.A) == 2 * d * T
    # tiny explicit pool: same result after several growth rounds
    res2 = engine.region_diff_batch(PA, Pb, RA, Rb, piece_cap=64)
    assert np.array_equal(res2.n_pieces, res.n_pieces) and np.array_equal(res2.A, res.A)

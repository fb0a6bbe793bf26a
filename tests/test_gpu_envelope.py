"""Shapes beyond the fused kernels' envelope (round-1 verdict: "loud, not wrong -- but not drops in unchanged"):
reduce / intersect with more than 64 rows, Chebyshev and bounding-box LPs with more than 128 rows or 32
columns, adjacency of cells with more than 32 rows.  They run on the one-LP-per-CTA solver
(pb200_lp_batch_big) under the same host API and are compared with the oracle exactly like the small shapes:
kept-row sets, flags and the drifted b identical, radii / bounds within 1e-9 of HiGHS."""
import numpy as np
import pytest

import workloads as wl

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('seed,m,d', [(9100, 100, 5), (9101, 150, 4), (9102, 72, 9), (9103, 260, 3)])
def test_reduce_with_more_than_64_rows(seed, m, d):
    import polytope_b200 as pc
    from oracle import polytope_oracle as orc
    A, b = wl.box_cuts(seed, m, d, True)
    p = pc.Polytope(A, b)
    red = pc.reduce(p)
    o = orc.reduce(A, b)
    An, bn, _ = orc.normalize_rows(A, b)
    assert red.minrep == o['minrep']
    # the reference returns Polytope(A_arr[keep_row], b_arr[keep_row]) (polytope.py:1161): its constructor normalises
    # the kept rows once more; b carries the +0.1 / -0.1 drift
    Ak, bk, _ = orc.normalize_rows(An[o['keep']], o['b'])
    assert np.array_equal(red.A, Ak)
    assert np.array_equal(red.b, bk)
    assert abs(p.chebR - o['r']) <= 1e-9


def test_intersect_of_two_polytopes_with_40_rows_each():
    import polytope_b200 as pc
    from oracle import polytope_oracle as orc
    A1, b1 = wl.box_cuts(9200, 40, 4)
    A2, b2 = wl.box_cuts(9201, 40, 4, True)
    p, q = pc.Polytope(A1, b1), pc.Polytope(A2, b2)
    r = p.intersect(q)
    o = orc.intersect(p.A, p.b, q.A, q.b)
    if o['empty']:
        assert pc.is_empty(r)
    else:
        S = np.vstack([p.A, q.A])
        Sn = orc.normalize_rows(S, np.hstack([p.b, q.b]))[0]
        Ak, bk, _ = orc.normalize_rows(Sn[o['keep']], o['b'])
        assert np.array_equal(r.A, Ak)
        assert np.array_equal(r.b, bk)


@pytest.mark.parametrize('m,d', [(70, 32), (90, 40), (5000, 6), (400, 12)])
def test_chebyshev_ball_beyond_128_rows_or_32_columns(m, d):
    """d = 32 is north_star's own example (n = 33 columns); 5000 rows is the size of extreme()'s is_fulldim(Q)."""
    import polytope_b200 as pc
    from oracle import polytope_oracle as orc
    A, b = wl.box_cuts(9300 + m + d, m, d, True)
    p = pc.Polytope(A, b)
    r, xc = pc.cheby_ball(p)
    ro, _ = orc.cheby_ball(p.A, p.b)
    assert abs(r - ro) <= 1e-9 * max(1.0, abs(ro))
    slack = p.b - p.A @ xc - r * np.sqrt((p.A * p.A).sum(1))
    assert slack.min() >= -1e-9
    assert pc.is_fulldim(p)


def test_bounding_box_beyond_128_rows():
    import polytope_b200 as pc
    from oracle import polytope_oracle as orc
    A, b = wl.box_cuts(9400, 300, 5, True)
    p = pc.Polytope(A, b)
    l, u = p.bounding_box
    lo, uo = orc.bounding_box(p.A, p.b)
    np.testing.assert_allclose(l, lo, rtol=0, atol=1e-9)
    np.testing.assert_allclose(u, uo, rtol=0, atol=1e-9)
    # unbounded direction: status 3 -> +-inf as the reference maps it (polytope.py:1372-1402)
    half = pc.Polytope(np.vstack([np.eye(3), -np.eye(3)[:2]] * 30), np.ones(150))
    l, u = half.bounding_box
    assert l[2, 0] == -np.inf and np.isfinite(u).all() and np.isfinite(l[:2]).all()


def test_adjacency_of_cells_with_more_than_32_rows():
    import polytope_b200 as pc
    from oracle import polytope_oracle as orc
    rng = np.random.default_rng(5)
    cells = []
    for k in range(5):                         # unit boxes side by side, each padded with 36 redundant cuts
        lo = np.array([float(k if k < 4 else 7), 0.0])
        C = rng.standard_normal((36, 2))
        C /= np.linalg.norm(C, axis=1)[:, None]
        A = np.vstack([np.eye(2), -np.eye(2), C])
        b = np.hstack([lo + 1.0, -lo, C @ (lo + 0.5) + 3.0])
        cells.append(pc.Polytope(A, b))
    adj = pc.adjacency_matrix(cells)
    ref = orc.adjacency_matrix([(c.A, c.b) for c in cells])
    assert np.array_equal(adj, ref)
    assert adj[0, 1] == 1 and adj[0, 2] == 0 and adj[3, 4] == 0

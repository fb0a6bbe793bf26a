"""GPU parity of qhull() / extreme() against the golden
vectors of the unmodified reference (d <= 6, what its Python quickhull can
finish) and against the oracle (Qhull) at the BASELINE sizes.

The hull of a point set is unique, the reference's facet order is not (its
start simplex is random, quickhull.py:172): facets and vertices are compared as
sets.  Bars: facet rows (A | b) within 1e-9, vertex sets within 1e-7, vertex-id
sets of the hull exact, facet counts exact.
"""
import numpy as np
import pytest

import workloads as wl

pytestmark = pytest.mark.gpu


def rows_as_set_close(X, Y, tol):
    """Every row of X has a row of Y within tol (max-norm) and vice versa."""
    from scipy.spatial import cKDTree
    X, Y = np.atleast_2d(X), np.atleast_2d(Y)
    dxy, _ = cKDTree(Y).query(X, p=np.inf)
    dyx, _ = cKDTree(X).query(Y, p=np.inf)
    return dxy.max() <= tol and dyx.max() <= tol


def test_qhull_matches_reference_golden(golden):
    import polytope_b200 as pc
    g = golden('hull_cases')
    for n, d in wl.HULL_SPECS:
        sets = [wl.hull_points(600 + 10 * d + i, n, d) for i in range(3)]
        hulls = pc.qhull_batch(sets)
        for i, q in enumerate(hulls):
            ref = np.c_[g['hull_d%d_%d_A' % (d, i)], g['hull_d%d_%d_b' % (d, i)]]
            assert q.minrep and len(q.b) == len(ref), (d, i, len(q.b), len(ref))
            assert rows_as_set_close(np.c_[q.A, q.b], ref, 1e-9)
            assert rows_as_set_close(q.vertices, g['hull_d%d_%d_vert' % (d, i)], 1e-12)
    cube = np.array([[x, y, z] for x in (0., 1.) for y in (0., 1.) for z in (0., 2.)])
    q = pc.qhull(np.vstack([cube, [[0.5, 0.5, 1.0]]]))
    # square facets come out as coplanar triangles in both implementations: compare the distinct planes
    assert rows_as_set_close(np.unique(np.round(np.c_[q.A, q.b], 9), axis=0),
                             np.unique(np.round(np.c_[g['hull_cube_A'], g['hull_cube_b']], 9), axis=0), 1e-9)
    assert len(q.b) == len(g['hull_cube_b']) == 12
    assert rows_as_set_close(q.vertices, g['hull_cube_vert'], 1e-12)
    assert len(pc.qhull(np.eye(3)).A) == 0                      # npt <= dim
    flat = np.c_[wl.hull_points(1, 10, 2), np.zeros(10)]
    assert len(pc.qhull(flat).A) == 0                           # not full-dimensional


@pytest.mark.parametrize('n,d', [(500, 2), (2000, 3), (6000, 3), (300, 4), (120, 6), (60, 8), (40, 10)])
def test_hull_batch_vs_qhull_oracle(n, d):
    from polytope_b200 import engine
    from oracle import polytope_oracle as orc
    H = 5
    rng = np.random.default_rng(n + d)
    pts = rng.standard_normal((H, n, d))
    npt = np.array([n, n - 3, n // 2, n, d + 1], dtype=np.int32)
    res = engine.hull_batch(pts, npt)
    assert np.all(res.status == 0)
    for h in range(H):
        P = pts[h, :npt[h]]
        A, b, vert = orc.qhull(P)
        gA, gb, vid = res.facets(h)
        assert len(gb) == len(b), (h, len(gb), len(b))
        assert rows_as_set_close(np.c_[gA, gb], np.c_[A, b], 1e-9)
        # vertices of the hull: exact index sets
        assert np.array_equal(np.nonzero(res.is_vertex[h])[0], np.unique(vid))
        assert rows_as_set_close(P[np.unique(vid)], vert, 0.0)
        # every input point satisfies every facet; each facet's vertices lie on it
        assert (P @ gA.T - gb).max() <= 1e-7
        assert np.abs(np.einsum('fkd,fd->fk', P[vid], gA) - gb[:, None]).max() < 1e-9


def test_extreme_matches_reference_golden(golden):
    import polytope_b200 as pc
    g = golden('hull_cases')
    for m, d in wl.EXTREME_SPECS:
        polys = [pc.Polytope(*wl.box_cuts(8500 + 10 * d + i, m, d, True)) for i in range(3)]
        Vs = pc.extreme_batch(polys)
        for i in range(3):
            ref = g['ext_d%d_%d' % (d, i)]
            assert Vs[i].shape == ref.shape, (d, i)
            assert rows_as_set_close(Vs[i], ref, 1e-7)
            assert polys[i].vertices is Vs[i]                     # cached as the reference does
    V = pc.extreme(pc.Polytope(*wl.unit_cube3()))
    assert rows_as_set_close(V, g['ext_cube3'], 1e-9) and V.shape == (8, 3)
    V1 = pc.extreme(pc.Polytope(np.array([[1.], [-1.]]), np.array([2., 1.])))
    np.testing.assert_allclose(np.sort(V1.flatten()), np.sort(g['ext_d1'].flatten()))
    flat = pc.Polytope(np.array([[1., 0, 0], [-1., 0, 0], [0, 1., 0], [0, -1., 0], [0, 0, 1.], [0, 0, -1.]]),
                       np.array([1., -1., 1, 1, 1, 1]))
    assert pc.extreme(flat) is None
    with pytest.raises(Exception):
        pc.extreme(pc.Region([pc.box2poly([[0, 1], [0, 1]])]))


@pytest.mark.parametrize('m,d,n', [(32, 8, 6), (40, 10, 4), (64, 12, 3)])
def test_extreme_vs_oracle_at_baseline_sizes(m, d, n):
    """cfg4 shape (d = 12, m = 64): ~20 000 vertices per polytope, the size the
    reference's own quickhull cannot reach."""
    import polytope_b200 as pc
    from oracle import polytope_oracle as orc
    polys = [pc.Polytope(*wl.box_cuts(4000 + i, m, d)) for i in range(n)]
    Vs = pc.extreme_batch(polys)
    for i in range(n):
        ref = orc.extreme(*wl.box_cuts(4000 + i, m, d))
        assert Vs[i].shape == ref.shape, (i, Vs[i].shape, ref.shape)
        assert rows_as_set_close(Vs[i], ref, 1e-7)
        # every vertex is feasible and has (at least) d tight rows
        slack = polys[i].b[None, :] - Vs[i] @ polys[i].A.T
        assert slack.min() > -1e-7
        assert ((np.abs(slack) < 1e-7).sum(1) >= d).all()


def test_empty_batches_are_fine():
    import torch
    from polytope_b200 import engine
    res = engine.hull_batch(np.zeros((0, 5, 3)))
    assert len(res.status) == 0 and len(res.b) == 0
    d = engine.region_diff_batch(np.zeros((0, 4, 2)), np.zeros((0, 4)), np.zeros((1, 4, 2)), np.zeros((1, 4)))
    assert len(d.status) == 0 and len(d.rows) == 0
    assert engine.contains_batch(np.zeros((1, 2, 2)), np.zeros((1, 2)), np.zeros((2, 0))).shape == (1, 0)
    r = engine.reduce_batch(torch.zeros((0, 4, 2), dtype=torch.float64, device='cuda'),
                            torch.zeros((0, 4), dtype=torch.float64, device='cuda'))
    assert len(r.keep) == 0


def test_extreme_sharded_single_rank_equals_batch():
    import polytope_b200 as pc
    from polytope_b200 import sharding
    polys = [pc.Polytope(*wl.box_cuts(8500 + 10 * 4 + i, 12, 4, True)) for i in range(3)]
    counts, V = sharding.extreme_sharded(polys)
    ref = pc.extreme_batch([pc.Polytope(p.A, p.b) for p in polys])
    assert counts.tolist() == [len(v) for v in ref]
    # the hull kernel emits facets in the order its warps finish: vertex order differs between runs
    Vh = V.cpu().numpy()
    off = 0
    for v in ref:
        assert rows_as_set_close(Vh[off:off + len(v)], v, 1e-12)      # ref re-normalises the rows
        off += len(v)

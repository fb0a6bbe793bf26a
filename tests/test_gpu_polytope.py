"""GPU parity tests of the batched polytope operations (cheby_ball, is_fulldim,
bounding_box, reduce, intersect, is_adjacent) against the golden vectors from
the unmodified reference and against the CPU oracle.

Bars: kept-row index sets, emptiness / full-dimensionality and
adjacency flags, LP counts: bit-exact.  Radii, objective-derived values, bounding
boxes: 1e-7 abs + 1e-7 rel.  Centres: 1e-7 where the optimum is unique.
"""
import numpy as np
import pytest

import workloads as wl

pytestmark = pytest.mark.gpu


def _unpad(row):
    return [int(v) for v in row if v >= 0]


def test_normalize_is_bit_exact_with_numpy():
    """Polytope.__init__ normalisation (polytope.py:128-138) incl. numpy's pairwise sum order."""
    from polytope_b200 import engine
    from oracle import polytope_oracle as orc
    rng = np.random.default_rng(0)
    for d in (1, 2, 3, 5, 7, 8, 9, 12, 13, 16, 17, 24, 31):
        A = rng.standard_normal((40, 20, d)) * 10 ** rng.uniform(-3, 3, (40, 20, 1))
        b = rng.standard_normal((40, 20))
        A[3, 5] = 0.0                       # zero row is dropped (norm <= 1e-10)
        An, bn, valid = engine.normalize_batch(A, b)
        for p in range(40):
            Ar, br, pos = orc.normalize_rows(A[p], b[p])
            mask = sum(1 << int(i) for i in pos)
            assert int(valid[p]) == mask
            assert np.array_equal(An[p][pos], Ar), d
            assert np.array_equal(bn[p][pos], br), d


@pytest.mark.parametrize('P,m,d', [(37, 33, 5), (37, 20, 8), (13, 64, 128), (1001, 32, 8), (5, 1, 4), (29, 64, 12)])
def test_normalize_tiles_ragged_rows_and_variants(P, m, d):
    """The tiled streaming kernel (bulk async copies when 16-byte aligned and m even, plain loads
    otherwise): partial tiles, odd shapes, ragged m_rows -- against numpy via the oracle, and every
    kernel variant against the default one bit for bit."""
    from polytope_b200 import engine, _capi
    from oracle import polytope_oracle as orc
    rng = np.random.default_rng(P * 1000 + m * 10 + d)
    A = rng.standard_normal((P, m, d)) * 10 ** rng.uniform(-3, 3, (P, m, 1))
    b = rng.standard_normal((P, m))
    A[P // 2, m // 2] = 0.0
    mr = rng.integers(0, m + 1, P).astype(np.int32)
    mr[0] = m
    lib = _capi.lib()
    try:
        An, bn, valid = engine.normalize_batch(A, b, mr)
        for p in range(P):
            k = int(mr[p])
            Ar, br, pos = orc.normalize_rows(A[p, :k], b[p, :k]) if k else (np.zeros((0, d)), np.zeros(0), [])
            assert int(valid[p]) & (2 ** 64 - 1) == sum(1 << int(i) for i in pos)
            assert np.array_equal(An[p][pos], Ar) and np.array_equal(bn[p][pos], br)
            assert not An[p, k:].any() and not bn[p, k:].any()      # rows beyond m_rows come back as zeros
        full = engine.normalize_batch(A, b)
        for variant in (-1, 0, 1, 2):
            lib.pb200_normalize_variant(variant)
            for want, got in zip((An, bn, valid), engine.normalize_batch(A, b, mr)):
                assert np.array_equal(want.view(np.uint64), got.view(np.uint64)), variant
            for want, got in zip(full, engine.normalize_batch(A, b)):
                assert np.array_equal(want.view(np.uint64), got.view(np.uint64)), variant
    finally:
        lib.pb200_normalize_variant(-2)


@pytest.mark.parametrize('m,d', [(6, 3), (16, 6), (32, 8), (64, 12), (64, 16), (100, 5)])
def test_cheby_and_bbox_vs_oracle(m, d):
    from polytope_b200 import engine
    from oracle import polytope_oracle as orc
    P = 24
    A, b = wl.box_cuts_batch(11, P, m, d, shift_scale=True)
    An = np.empty_like(A)
    bn = np.empty_like(b)
    for p in range(P):
        An[p], bn[p], _ = orc.normalize_rows(A[p], b[p])
    r, xc, st = engine.cheby_batch(An, bn)
    lo, hi, bst = engine.bbox_batch(An, bn)
    assert np.all(st == 0) and np.all(bst == 0)
    for p in range(P):
        rr, xx = orc.cheby_ball(An[p], bn[p])
        assert abs(r[p] - rr) <= 1e-7 + 1e-7 * abs(rr)
        # centre may be non-unique: it must be feasible with slack >= r on every row
        assert np.all(An[p] @ xc[p] + r[p] * np.sqrt(np.sum(An[p] ** 2, 1)) <= bn[p] + 1e-9)
        if p < 6:
            l, u = orc.bounding_box(An[p], bn[p])
            np.testing.assert_allclose(lo[p], l[:, 0], rtol=1e-7, atol=1e-7)
            np.testing.assert_allclose(hi[p], u[:, 0], rtol=1e-7, atol=1e-7)


def test_bbox_unbounded_and_empty_conventions(golden):
    """status 3 -> -/+inf, status 2 -> l = 0, u = l (polytope.py:1372-1402)."""
    from polytope_b200 import engine
    g = golden('named_cases')
    A = np.array([[[1., 0.], [0., 1.], [-1, 0]]])
    lo, hi, st = engine.bbox_batch(A, np.ones((1, 3)))
    assert np.array_equal(lo[0], g['unb_l'][:, 0]) and np.array_equal(hi[0], g['unb_u'][:, 0])
    lo, hi, st = engine.bbox_batch(np.array([[[1.], [-1.]]]), np.array([[0., -1.]]))
    assert np.array_equal(lo[0], g['empty_l'][:, 0]) and np.array_equal(hi[0], g['empty_u'][:, 0])
    assert list(st[0]) == [2, 2]


@pytest.mark.parametrize('tag', ['cfg2', 'cfg3', 'cfg4', 'd16', 'small'])
def test_reduce_batch_matches_reference_golden(golden, tag):
    from polytope_b200 import engine
    g = golden('reduce_cases')
    cfg, n, m, d, ss = [int(v) for v in g[tag + '_spec']]
    A, b = wl.box_cuts_batch(cfg, n, m, d, bool(ss))
    res = engine.reduce_batch(A, b)
    keeps = res.keep_lists()
    for i in range(n):
        assert keeps[i] == _unpad(g[tag + '_keep'][i]), (tag, i)
        assert bool(res.flags[i] & engine.F_EMPTY) == bool(g[tag + '_empty'][i])
        assert bool(res.flags[i] & engine.F_MINREP) == bool(g[tag + '_minrep'][i])
        assert not (res.flags[i] & engine.F_LPFAIL)
        assert int(res.n_lp[i]) == int(g[tag + '_nlp'][i])
        assert abs(res.r[i] - g[tag + '_r'][i]) <= 1e-7 + 1e-7 * abs(g[tag + '_r'][i])


@pytest.mark.parametrize('cfg,n,m,d,ss', [(2, 256, 32, 8, False), (3, 256, 16, 6, True),
                                           (4, 32, 64, 12, False), (6, 24, 64, 16, False),
                                           (12, 128, 24, 4, True), (13, 128, 12, 2, True),
                                           (14, 64, 40, 3, False), (15, 48, 40, 10, False), (16, 24, 56, 14, True)])
def test_reduce_batch_vs_oracle(cfg, n, m, d, ss):
    """Fresh seeds (not in the golden files): kept rows, flags, LP counts and the
    drifted b identical to the oracle; the normalised A bit-identical to numpy's."""
    from polytope_b200 import engine
    from oracle import polytope_oracle as orc
    A, b = wl.box_cuts_batch(cfg, n, m, d, ss, first=500)
    res = engine.reduce_batch(A, b)
    keeps = res.keep_lists()
    for i in range(n):
        o = orc.reduce(A[i], b[i])
        assert keeps[i] == o['keep'], (cfg, i, keeps[i], o['keep'])
        assert bool(res.flags[i] & engine.F_EMPTY) == o['empty']
        assert bool(res.flags[i] & engine.F_MINREP) == o['minrep']
        assert int(res.n_lp[i]) == o['n_lp']
        assert abs(res.r[i] - o['r']) <= 1e-9 + 1e-9 * abs(o['r'])
        An, bn, _ = orc.normalize_rows(A[i], b[i])
        assert np.array_equal(res.A[i], An)
        assert np.array_equal(res.b[i][o['keep']], o['b'])     # includes the +0.1/-0.1 drift


def test_reduce_edge_cases_vs_oracle():
    """Duplicates (tie -> first of the pair removed), touching cuts, early exits,
    empty and flat polytopes, zero rows, b = inf rows, ragged batches."""
    from polytope_b200 import engine
    from oracle import polytope_oracle as orc
    cases = []
    for d in (2, 3, 5):                     # cube + vertex-touching cut + exact duplicate
        A = np.vstack([np.eye(d), -np.eye(d), np.ones((1, d)) / np.sqrt(d), np.eye(d)[:1]])
        b = np.hstack([np.ones(d), np.zeros(d), [d / np.sqrt(d)], [1.0]])
        cases.append((A, b))
    a = np.array([[1.0, 0.1], [1.0, 0.1], [-1., 0.], [0., 1.], [0., -1.]])   # reference test_reduce
    cases.append((a, np.array([50., 50.5, -40., 1., 0.])))
    cases.append((np.array([[1., 0], [-1, 0], [0, 1], [0, -1]]), np.array([1., -2., 1., 1.])))  # empty
    cases.append((np.array([[1., 0], [-1, 0], [0, 1], [0, -1]]), np.array([1., -1., 1., 1.])))  # flat
    cases.append((np.array([[1., 0], [-1, 0], [0, 1]]), np.array([1., 1., 1.])))               # m <= d+1
    # zero row is dropped by the constructor (b = inf rows cannot reach reduce()
    # through the scipy adapter: linprog rejects them in the leading is_fulldim)
    cases.append((np.array([[1., 0], [0, 0], [-1, 0], [0, 1], [0, -1], [1, 1]]),
                  np.array([1., 5., 1., 1., 1., 7.])))
    A8, b8 = wl.box_cuts(4242, 30, 4)
    cases.append((np.vstack([A8, A8[:5] * 3.0]), np.hstack([b8, b8[:5] * 3.0 + 0.5])))   # scaled dups
    mmax = max(len(c[1]) for c in cases)
    for d in sorted(set(c[0].shape[1] for c in cases)):
        sel = [c for c in cases if c[0].shape[1] == d]
        A = np.zeros((len(sel), mmax, d))
        b = np.zeros((len(sel), mmax))
        rows = np.array([len(c[1]) for c in sel], dtype=np.int32)
        for k, (Ak, bk) in enumerate(sel):
            A[k, :rows[k]], b[k, :rows[k]] = Ak, bk
        res = engine.reduce_batch(A, b, m_rows=rows)
        keeps = res.keep_lists()
        for k, (Ak, bk) in enumerate(sel):
            o = orc.reduce(Ak, bk)
            assert keeps[k] == o['keep'], (d, k, keeps[k], o['keep'])
            assert bool(res.flags[k] & engine.F_EMPTY) == o['empty'], (d, k)
            assert bool(res.flags[k] & engine.F_MINREP) == o['minrep'], (d, k)
            assert int(res.n_lp[k]) == o['n_lp'], (d, k)


@pytest.mark.parametrize('tag', ['g2', 'g3', 'g4'])
def test_adjacency_grid_matches_reference_golden(golden, tag):
    import polytope_b200 as pb
    g = golden('adjacent_cases')
    A, b, idx = wl.box_grid(tuple(int(s) for s in g[tag + '_shape']))
    cells = [pb.Polytope(A[i], b[i]) for i in range(len(A))]
    adj = pb.adjacency_matrix(cells)
    assert np.array_equal(adj, g[tag + '_adj'])
    # the multi-GPU entry (pair range split over the ranks; one rank here) gives the same flags
    from polytope_b200 import sharding
    An = np.array([c.A for c in cells])
    bn = np.array([c.b for c in cells])
    flags = sharding.adjacency_sharded(An, bn).cpu().numpy()
    i, j = np.tril_indices(len(cells), -1)
    assert np.array_equal(flags, g[tag + '_adj'][i, j])


@pytest.mark.parametrize('origin,cell', [(0.0, 1.0), (1e3, 1.0), (-3e3, 0.01), (0.0, 1e3), (1e5, 1.0)])
def test_adjacency_scales(origin, cell):
    """Touching boxes give r = 1e-7 against the 1e-8 threshold:
    flags must stay exact at large coordinates; radii within 1e-9 of the oracle."""
    from polytope_b200 import engine
    from oracle import polytope_oracle as orc
    A, b, idx = wl.box_grid((5, 4), origin, cell)
    cells = [orc.normalize_rows(A[i], b[i])[:2] for i in range(len(A))]
    An = np.stack([c[0] for c in cells])
    bn = np.stack([c[1] for c in cells])
    flags, rad, st = engine.adjacent_pairs(An, bn)
    touch = np.abs(idx[:, None, :] - idx[None, :, :]).max(-1) <= 1
    i, j = np.tril_indices(len(A), -1)
    assert np.array_equal(flags.astype(bool), touch[i, j])
    assert np.all(st == 0)
    for t in range(0, len(i), 7):
        Ad, bd = orc.adjacent_lp_data(*cells[i[t]], *cells[j[t]])
        c, G, h = orc.cheby_lp_data(Ad, bd)
        ref = orc.lpsolve(c, G, h)
        assert abs(rad[t] - ref['x'][-1]) <= 1e-9 * max(1.0, abs(origin), cell), (t, rad[t], ref['x'][-1])


def test_adjacent_random_matches_reference_golden(golden):
    import polytope_b200 as pb
    g = golden('adjacent_cases')
    for i in range(60):
        A1, b1 = wl.box_cuts(9000 + i, 12, 4, True)
        A2, b2 = wl.box_cuts(9500 + i, 12, 4, True)
        rng = np.random.default_rng(77 + i)
        b2 = b2 + A2 @ (rng.uniform(-1, 1, 4) * (i % 4))
        q1, q2 = pb.Polytope(A1, b1), pb.Polytope(A2, b2)
        assert bool(pb.is_adjacent(q1, q2)) == bool(g['rand_flag'][i]), i


def test_intersect_matches_reference_golden(golden):
    import polytope_b200 as pb
    g = golden('intersect_cases')
    Q = pb.Polytope(*wl.box_cuts(3999, 16, 6, True))
    Ps = [pb.Polytope(*wl.box_cuts(3000 + i, 16, 6, True)) for i in range(24)]
    Qs = [pb.Polytope(Q.A.copy(), Q.b.copy()) for _ in range(24)]
    out = pb.intersect_batch(Ps, Qs)
    single = Ps[0].intersect(Qs[0])
    for i in range(24):
        empty = len(out[i].A) == 0
        assert empty == bool(g['empty'][i]), i
        keep = _unpad(g['keep'][i])
        assert out[i].A.shape[0] == len(keep)
        if not empty:
            stacked = pb.Polytope(np.vstack([Ps[i].A, Qs[i].A]), np.hstack([Ps[i].b, Qs[i].b]))
            np.testing.assert_allclose(out[i].A, stacked.A[keep], rtol=0, atol=1e-15)
            np.testing.assert_allclose(out[i].b, stacked.b[keep], rtol=0, atol=1e-15)
            rp, _ = out[i].cheby
            assert abs(rp - g['r'][i]) <= 1e-7 + 1e-7 * abs(g['r'][i])
    assert np.array_equal(single.A, out[0].A)


def test_api_shell_mirrors_reference_operations_tests(golden):
    """operations_test.polytope_full_dim_test / polytope_intersect_test /
    region_full_dim_test and test_reduce / test_bounding_box_to_polytope of the
    reference (tests/polytope_test.py:200-238, :299-312, :601-622)."""
    import polytope_b200 as pc
    g = golden('named_cases')
    Ab = np.array([[0., 1, 1], [0, -1, 0], [1, 0, 1], [-1, 0, 0]])
    Ab2 = np.array([[-1., 0, 1], [1, 0, 0], [0, 1, 1], [0, -1, 0]])
    A, b = Ab[:, :2], Ab[:, 2]
    assert pc.is_fulldim(pc.Polytope(A, b))
    assert pc.is_fulldim(pc.Polytope(Ab2[:, :2], Ab2[:, 2]))
    assert not pc.is_fulldim(pc.Polytope())
    assert not pc.is_fulldim(pc.Polytope(A, b - 1e3))
    p1 = pc.Polytope(A, b)
    p2 = pc.Polytope(Ab2[:, :2], Ab2[:, 2])
    p3 = p1.intersect(p2)
    assert pc.is_fulldim(p1) and pc.is_fulldim(p2) and not pc.is_fulldim(p3)
    p4 = pc.Polytope(np.array([[1., 0.], [0., 1.], [-1., 0.], [0., -1.]]), np.array([.5, .5, .5, .5]))
    p5 = p2.intersect(p4)
    assert pc.is_fulldim(p4) and pc.is_fulldim(p5)
    np.testing.assert_allclose(p5.A, g['sq_p5_A'], atol=1e-15)
    np.testing.assert_allclose(p5.b, g['sq_p5_b'], atol=1e-15)
    np.testing.assert_allclose(p1.chebXc, [0.5, 0.5], atol=1e-9)
    np.testing.assert_allclose(p2.chebXc, [-0.5, 0.5], atol=1e-9)
    assert abs(p1.chebR - 0.5) < 1e-9
    # regions
    assert not pc.is_fulldim(pc.Region())
    reg = pc.Region([pc.Polytope(A, b), pc.Polytope(Ab2[:, :2], Ab2[:, 2])])
    assert pc.is_fulldim(reg)
    reg.list_poly.append(pc.Polytope())
    reg.list_poly.append(pc.Polytope(A, b - 1e3))
    reg.fulldim = None
    assert pc.is_fulldim(reg)
    l, u = pc.Region([pc.Polytope(A, b), pc.Polytope(Ab2[:, :2], Ab2[:, 2])]).bounding_box
    np.testing.assert_allclose(l, [[-1.], [0.]], atol=1e-7)
    np.testing.assert_allclose(u, [[1.], [1.]], atol=1e-7)
    # test_reduce
    a = np.array([[1.0, 0.1], [1.0, 0.1], [-1., 0.], [0., 1.], [0., -1.]])
    bb = np.array([50., 50.5, -40., 1., 0.])
    poly2 = pc.reduce(pc.Polytope(a, bb))
    assert np.array_equal(poly2.A, g['treduce_A']) and np.array_equal(poly2.b, g['treduce_b'])
    l, u = poly2.bounding_box
    np.testing.assert_allclose(l, np.array([[40.], [0.]]), rtol=1e-07, atol=1e-07)
    np.testing.assert_allclose(u, np.array([[50.], [1.]]), rtol=1e-07, atol=1e-07)
    # bounding boxes of boxes
    for k, iv in enumerate([[[0, 1]], [[0, 1], [0, 2]], [[-1, 2], [3, 5], [-5, -3]]]):
        l, u = pc.box2poly(iv).bounding_box
        np.testing.assert_allclose(l, g['bbox%d_l' % k], atol=1e-9)
        np.testing.assert_allclose(u, g['bbox%d_u' % k], atol=1e-9)
    # cfg1: unit cube built with the constructor (not from_box): 6 rows kept, r = .5
    cube = pc.Polytope(*wl.unit_cube3())
    red = pc.reduce(cube)
    assert red.A.shape == (6, 3) and red.minrep
    rr, xx = pc.cheby_ball(red)
    assert abs(rr - 0.5) < 1e-9 and np.allclose(xx, 0.5, atol=1e-9)
    # half-space: Chebyshev LP unbounded -> radius 0
    assert pc.cheby_ball(pc.Polytope(np.array([[1., 0.]]), np.array([1.])))[0] == 0


def test_reduce_full_size_properties():
    """BASELINE cfg2 at full size (10 000 x 32 x 8) through size-independent
    properties: cuts with t >= 1 are redundant by construction and must be
    dropped, every box facet that no cut removes stays, reduce is idempotent,
    and a random sample agrees with the oracle."""
    from polytope_b200 import engine
    from oracle import polytope_oracle as orc
    P, m, d = 10000, 32, 8
    A, b = wl.box_cuts_batch(2, P, m, d)
    res = engine.reduce_batch(A, b)
    keep = res.keep.astype(np.uint64)
    bits = ((keep[:, None] >> np.arange(m, dtype=np.uint64)) & np.uint64(1)).astype(bool)
    assert not np.any(res.flags & (engine.F_EMPTY | engine.F_LPFAIL))
    assert np.all(res.flags & engine.F_MINREP) and np.all(res.flags & engine.F_BBOX)
    is_box = (np.abs(A) == 1.0).sum(2) == 1
    t = b / np.abs(A).sum(2)                       # rows are unit 2-norm: b = t * ||a||_1
    assert not np.any(bits & ~is_box & (t >= 1.0))
    assert np.all(bits.sum(1) >= d + 1)
    kept = bits.sum(1)
    assert np.all(res.n_lp >= 1 + 2 * d + kept) and np.all(res.n_lp <= 1 + 2 * d + m)
    # idempotence: reducing the reduced polytopes keeps every row
    rows = bits.sum(1).astype(np.int32)
    A2 = np.zeros_like(A)
    b2 = np.zeros_like(b)
    for p in range(P):
        A2[p, :rows[p]] = res.A[p][bits[p]]
        b2[p, :rows[p]] = res.b[p][bits[p]]
    res2 = engine.reduce_batch(A2, b2, m_rows=rows)
    assert np.array_equal(np.array([len(k) for k in res2.keep_lists()]), rows)
    rng = np.random.default_rng(0)
    for p in rng.choice(P, 40, replace=False):
        o = orc.reduce(A[p], b[p])
        assert np.nonzero(bits[p])[0].tolist() == o['keep'], p
        assert int(res.n_lp[p]) == o['n_lp']


def test_host_batches_are_pipelined_and_equal_the_device_path():
    """Host-resident batches go through the chunked two-stream path (H2D / kernels / D2H
    overlapped); results must be identical to the device-resident call, ragged rows included."""
    import torch
    from polytope_b200 import engine
    P, m, d = 5000, 20, 5
    A, b = wl.box_cuts_batch(12, P, m, d, shift_scale=True)
    rows = np.random.default_rng(0).integers(2 * d, m + 1, P).astype(np.int32)
    dev = engine.reduce_batch(torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(rows).cuda())
    host = engine.reduce_batch(A, b, rows)
    for name in ('keep', 'flags', 'n_lp', 'b', 'A'):
        assert np.array_equal(np.asarray(getattr(host, name)), getattr(dev, name).cpu().numpy(), equal_nan=True), name
    # the lane kernels solve 32 LPs per warp in lockstep and a lane's polish step waits for its warp mates, so
    # the last bits of an LP's solution depend on which LPs share its warp -- chunking regroups them
    r = dev.r.cpu().numpy()
    np.testing.assert_allclose(np.asarray(host.r), r, rtol=0, atol=1e-12, equal_nan=True)
    # the same chunked path with the results left on the device (what a sharded caller all-gathers)
    kept = engine.reduce_batch(A, b, rows, want_A=False, want_b=False, results_on_device=True)
    assert isinstance(kept.keep, torch.Tensor) and kept.keep.is_cuda and kept.A is None
    for name in ('keep', 'flags', 'n_lp'):
        assert torch.equal(getattr(kept, name), getattr(dev, name)), name
    # the Chebyshev centre is not unique (polytope.py:1245-1246): each path's centre must admit the ball
    An, bn = np.asarray(host.A), np.asarray(host.b)
    live = np.arange(m)[None, :] < rows[:, None]
    for xc in (np.asarray(host.xc), dev.xc.cpu().numpy()):
        slack = bn - np.einsum('pij,pj->pi', An, xc) - r[:, None]
        assert slack[live & np.isfinite(slack)].min() >= -1e-9
    slim = engine.reduce_batch(torch.from_numpy(A).pin_memory(), torch.from_numpy(b).pin_memory(), rows,
                               want_A=False, want_b=False)
    assert slim.A is None and slim.b is None and np.array_equal(slim.keep, host.keep)
    assert slim.keep_lists() == host.keep_lists()


def test_reduce_non_empty_bounded_zero_skips_the_early_exits():
    """reduce(poly, nonEmptyBounded=0): the two `neq <= nx + 1` exits (polytope.py:1113-1116,
    :1135-1138) are skipped, so small polytopes go through the row LPs and come back minrep."""
    import polytope_b200 as pc
    from polytope_b200 import engine
    from oracle import polytope_oracle as orc
    cases = [(np.array([[1., 0], [-1, 0], [0, 1]]), np.array([1., 1., 1.])),                    # m <= d+1, unbounded
             (np.array([[1., 0], [-1, 0], [0, 1], [0, -1]]), np.array([1., 1., 1., 1.]))]
    for i in range(6):
        cases.append(wl.box_cuts(7700 + i, 10 + i, 3, True))
    for A, b in cases:
        res = engine.reduce_batch(A[None], b[None], non_empty_bounded=False)
        o = orc.reduce(A, b, non_empty_bounded=False)
        assert res.keep_lists()[0] == o['keep']
        assert int(res.n_lp[0]) == o['n_lp']
        assert bool(res.flags[0] & engine.F_MINREP) == o['minrep']
        red = pc.reduce(pc.Polytope(A, b), nonEmptyBounded=0)
        assert red.minrep == o['minrep'] and len(red.b) == len(o['keep'])


def test_reduce_at_the_largest_advertised_shape():
    """m = 64, d = 31 (n = 32 Chebyshev LP): the prefilter kernel needs > 48 KB of dynamic shared
    memory there; keep sets against the oracle."""
    from polytope_b200 import engine
    from oracle import polytope_oracle as orc
    A, b = wl.box_cuts_batch(15, 3, 64, 31)
    res = engine.reduce_batch(A, b)
    for i in range(len(A)):
        o = orc.reduce(A[i], b[i])
        assert res.keep_lists()[i] == o['keep']
        assert int(res.n_lp[i]) == o['n_lp']
        assert abs(res.r[i] - o['r']) <= 1e-9


def test_adjacent_range_enumerations_equal_pair_lists():
    """pb200_adjacent_range: both implicit enumerations (find_adjacent_regions' j < i order and
    compute_adj's ordered pairs, prop2partition.py:57-61, :253-261), whole and in blocks, against
    explicit pair lists through pb200_adjacent_pairs on both LP kernels, and against geometry."""
    from polytope_b200 import engine, sharding
    for shape in ((5, 4), (3, 3, 2), (2, 2, 2, 2)):
        A, b, idx = wl.box_grid(shape)
        n = len(A)
        for order, block in ((0, sharding.pair_block), (1, sharding.ordered_pair_block)):
            T = n * (n - 1) // (2 if order == 0 else 1)
            pi, pj = block(n, 0, T)
            touch = np.abs(idx[pi.numpy()] - idx[pj.numpy()]).max(1) <= 1
            whole, rad, st = engine.adjacent_range(A, b, order, 0, T)
            assert np.array_equal(whole.astype(bool), touch) and np.all(st == 0)
            cut = T // 3
            parts = np.concatenate([engine.adjacent_range(A, b, order, 0, cut)[0], engine.adjacent_range(A, b, order, cut, T - cut)[0]])
            assert np.array_equal(parts, whole)
            lst, rad2, _ = engine.adjacent_pairs(A, b, pi.numpy(), pj.numpy())
            assert np.array_equal(lst, whole) and np.allclose(rad2, rad, atol=1e-12, equal_nan=True)
            engine.lane_solver(False)
            try:
                warp, rad3, _ = engine.adjacent_pairs(A, b, pi.numpy(), pj.numpy())
            finally:
                engine.lane_solver(True)
            assert np.array_equal(warp, whole) and np.allclose(rad3, rad, atol=1e-11, equal_nan=True)


@pytest.mark.parametrize('cfg,n,m,d,sample', [(4, 2048, 64, 16, 12), (4, 600, 48, 11, 16), (4, 2048, 50, 13, 12)])
def test_wide_lane_solver_equals_the_warp_solver_and_the_oracle(cfg, n, m, d, sample):
    """9 <= n <= 16 columns: one LP per lane with the factor in shared memory (lp_lane_wide.cuh; n >= 13 only for
    batches of >= 2000 polytopes).  Every keep mask / flag / LP count equals the warp-per-LP path's, a sample the oracle's."""
    import torch
    from polytope_b200 import engine
    from oracle import polytope_oracle as orc
    A, b = wl.box_cuts_batch(cfg, n, m, d, first=900)
    Ad, bd = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    launches = engine.launch_count()
    lane = engine.reduce_batch(Ad, bd, want_A=False)
    n_launch = engine.launch_count() - launches
    try:
        engine.lane_solver(False)
        warp = engine.reduce_batch(Ad, bd, want_A=False)
    finally:
        engine.lane_solver(True)
    assert n_launch == 10          # 8 kernels of the pipeline + the two (empty) retry launches of the lane path
    assert torch.equal(lane.keep, warp.keep) and torch.equal(lane.flags, warp.flags) and torch.equal(lane.n_lp, warp.n_lp)
    assert not bool((lane.flags & engine.F_LPFAIL).any())
    keeps = lane.keep_lists()
    for i in range(0, n, n // sample):
        o = orc.reduce(A[i], b[i])
        assert keeps[i] == o['keep'] and int(lane.n_lp[i]) == o['n_lp']
        assert abs(float(lane.r[i]) - o['r']) <= 1e-9

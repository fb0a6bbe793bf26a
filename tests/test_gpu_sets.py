"""GPU parity of the point-set rows: contains, volume,
grid_region, enumerate_integral_points and quickhull's distance sweep, against
the golden vectors of the unmodified reference and against the CPU oracle.

Bars: containment flags and Monte-Carlo counts bit-exact (the kernels
accumulate A.x exactly as numpy's dgemm does and regenerate numpy's PCG64
stream); volumes equal to the reference's to the last bit for the same seed.
"""
import numpy as np
import pytest

import workloads as wl

pytestmark = pytest.mark.gpu


def test_pcg64_stream_on_device_matches_numpy():
    """volume() on the unit box with a half-space x0 < t counts exactly the draws < t."""
    from polytope_b200 import engine
    for d, N, seed in [(1, 50, 0), (2, 500, 7), (3, 3000, 11), (5, 10000, 2**40 + 3), (8, 10000, 5), (16, 4097, 1)]:
        g = np.random.default_rng(seed)
        U = np.random.default_rng(seed).random((d, N))
        for k in range(d):
            t = 0.37 + 0.01 * k
            A = np.zeros((1, 1, d))
            A[0, 0, k] = 1.0
            cnt = engine.volume_counts(A, np.array([[t]]), np.zeros((1, d)), np.ones((1, d)), N,
                                       [engine.pcg64_state_words(g.bit_generator)])
            assert int(cnt[0]) == int(np.sum(U[k] - t < 0)), (d, N, k)


def test_volume_matches_reference_bit_for_bit(golden):
    import polytope_b200 as pc
    g = golden('sets_cases')
    for m, d in wl.VOLUME_SPECS:
        polys = [pc.Polytope(*wl.box_cuts(8000 + 10 * d + i, m, d, True)) for i in range(4)]
        vols = pc.volume_batch(polys, seeds=[100 + i for i in range(4)])
        for i in range(4):
            l, u = polys[i].bounding_box
            np.testing.assert_allclose(np.c_[l, u], g['vol_d%d_box' % d][i], rtol=1e-9, atol=1e-9)
            # same count => same volume up to the bbox LP's last bits
            assert abs(vols[i] - g['vol_d%d' % d][i]) <= 1e-9 * g['vol_d%d' % d][i], (d, i)
            N = 500 if d == 2 else 3000 if d == 3 else 10000
            cnt_ref = g['vol_d%d' % d][i] / np.prod(g['vol_d%d_box' % d][i][:, 1] - g['vol_d%d_box' % d][i][:, 0]) * N
            cnt = vols[i] / np.prod(u - l) * N
            assert round(cnt) == round(cnt_ref)
        p = pc.Polytope(*wl.box_cuts(8000 + 10 * d, m, d, True))
        v = pc.volume(p, nsamples=777, seed=5)
        assert abs(v - float(g['vol_d%d_n777' % d])) <= 1e-9 * abs(v)
        assert p.volume == v                      # cached by _set_volume
        gen = np.random.default_rng(9)            # a Generator is consumed exactly as numpy consumes it
        v1 = pc.volume(pc.Polytope(p.A, p.b), seed=gen)
        v2 = pc.volume(pc.Polytope(p.A, p.b), seed=gen)
        np.testing.assert_allclose([v1, v2], g['vol_d%d_gen' % d], rtol=1e-9)
    p1 = pc.Polytope(np.array([[1.], [-1.]]), np.array([2., 1.]))
    assert abs(pc.volume(p1, seed=1) - float(g['vol_d1'])) < 1e-9
    with pytest.raises(ValueError):
        pc.volume(p1, nsamples=0)


def test_volume_counts_equal_oracle_counts():
    from polytope_b200 import engine
    from oracle import polytope_oracle as orc
    for m, d in [(10, 3), (16, 6), (32, 8), (64, 12)]:
        P = 6
        A, b = wl.box_cuts_batch(21, P, m, d, shift_scale=True)
        An, bn, lo, hi, words, ref = [], [], [], [], [], []
        for p in range(P):
            a_, b_, _ = orc.normalize_rows(A[p], b[p])
            l, u = orc.bounding_box(a_, b_)
            An.append(a_), bn.append(b_), lo.append(l[:, 0]), hi.append(u[:, 0])
            words.append(engine.pcg64_state_words(np.random.default_rng(1000 + p).bit_generator))
            N = 3000 if d == 3 else 10000
            x = l + np.random.default_rng(1000 + p).random((d, N)) * (u - l)
            ref.append(int(np.sum(np.all(a_.dot(x) - b_[:, None] < 0, 0))))
        cnt = engine.volume_counts(np.array(An), np.array(bn), np.array(lo), np.array(hi), N, words)
        assert [int(c) for c in cnt] == ref, (m, d)


def test_contains_matches_reference(golden):
    import polytope_b200 as pc
    g = golden('sets_cases')
    for m, d in wl.VOLUME_SPECS:
        x = wl.contains_points(50 + d, d, 2000)
        polys = [pc.Polytope(*wl.box_cuts(8000 + 10 * d + i, m, d, True)) for i in range(3)]
        for i in range(3):
            assert np.array_equal(polys[i].contains(x), g['contains_d%d' % d][i])
        assert np.array_equal(pc.Region(polys).contains(x), g['region_contains_d%d' % d])
        assert np.array_equal(polys[0].contains(x, abs_tol=0), g['contains_tol0_d%d' % d])
    with pytest.raises(ValueError):
        pc.Region(polys).contains(np.zeros((d + 1, 4)))
    # region_contains_test of the reference (tests/polytope_test.py)
    pts = np.array([[-1.0, 0.0, 0.5, 1.0, 2.0]])
    r1 = pc.Region([pc.Polytope(np.array([[1.0], [-1.0]]), np.array([1.0, 0.0]))])
    assert r1.contains(pts).tolist() == [False, True, True, True, False]
    assert r1.contains(pts, abs_tol=0).tolist() == [False, False, True, False, False]


@pytest.mark.parametrize('P,m,d,N', [(1, 6, 3, 100000), (2, 6, 3, 99999), (5, 9, 7, 1), (40, 32, 8, 30000), (300, 12, 4, 5000), (7, 64, 16, 2500),
                                     (3, 20, 20, 1000)])
def test_contains_batch_vs_oracle_ragged(P, m, d, N):
    from polytope_b200 import engine
    from oracle import polytope_oracle as orc
    rng = np.random.default_rng(P + m)
    A = rng.standard_normal((P, m, d))
    b = rng.uniform(0.2, 2.0, (P, m))
    rows = rng.integers(1, m + 1, P).astype(np.int32)
    x = rng.uniform(-1, 1, (d, N))
    got = engine.contains_batch(A, b, x, rows, abs_tol=1e-7)
    any_got = engine.contains_batch(A, b, x, rows, abs_tol=1e-7, any_of=True)
    ref = np.array([orc.contains(A[p, :rows[p]], b[p, :rows[p]], x) for p in range(P)])
    if d <= 15:
        assert np.array_equal(got, ref)
    else:       # numpy's dgemm tail reorders sums for d >= 16: decisions may differ within 1e-14 of the threshold
        bad = np.argwhere(got != ref)
        for p, j in bad:
            res = A[p, :rows[p]].dot(x[:, j]) - b[p, :rows[p]] - 1e-7
            assert np.abs(res).min() < 1e-13
    assert np.array_equal(any_got, got.any(0))


def test_grid_region_and_integral_points(golden):
    import polytope_b200 as pc
    g = golden('sets_cases')
    p = pc.Polytope(*wl.box_cuts(8020, 6, 2, True))
    x, res = pc.grid_region(p)
    assert list(res) == list(g['grid2_res'])
    np.testing.assert_allclose(x, g['grid2_x'], atol=1e-9)
    p = pc.Polytope(*wl.box_cuts(8030, 10, 3, True))
    x, res = pc.grid_region(p, res=[7, 5, 6])
    np.testing.assert_allclose(x, g['grid3_x'], atol=1e-9)
    A, b = wl.box_cuts(8030, 10, 3, False)
    assert np.array_equal(pc.enumerate_integral_points(pc.Polytope(A, 3.5 * b)), g['integral3'])
    reg = pc.Region([pc.box2poly([[0, 2], [0, 1]]), pc.box2poly([[1, 3], [0.5, 2.5]])])
    assert np.array_equal(pc.enumerate_integral_points(reg), g['integral_region'])
    with pytest.raises(ValueError):
        pc.grid_region(p, res=[3, 3])


def test_point_facet_sweep_vs_numpy():
    from polytope_b200 import engine
    rng = np.random.default_rng(4)
    for N, F, d in [(1000, 5, 3), (50000, 700, 6), (333, 2000, 12), (10, 1, 2)]:
        pts = rng.standard_normal((N, d))
        nrm = rng.standard_normal((F, d))
        nrm /= np.linalg.norm(nrm, axis=1)[:, None]
        off = rng.uniform(0.5, 2.0, F)
        first, far, dist = engine.point_facet_sweep(pts, nrm, off, 1e-7)
        D = np.array([[np.sum(nrm[f] * pts[j]) - off[f] for f in range(F)] for j in range(min(N, 200))])
        out = D > 1e-7
        ref_first = np.where(out.any(1), out.argmax(1), -1)
        assert np.array_equal(first[:len(D)], ref_first)
        assert np.array_equal(far[:len(D)], D.argmax(1))
        assert np.array_equal(dist[:len(D)], D.max(1))      # numpy's summation order reproduced

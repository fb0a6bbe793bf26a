"""The oracle (oracle/polytope_oracle.py) against the golden vectors recorded
from the unmodified reference by tests/golden/make_golden.py.

Bit-exact wherever the reference's numpy logic decides the result; the LP
values come from the same scipy.optimize.linprog call, so they are compared
exactly too when the scipy version matches the one the fixtures were made with.
"""
import numpy as np
import scipy
import pytest

import workloads as wl
from oracle import polytope_oracle as orc


def _same_scipy(g):
    return ("'scipy': '%s'" % scipy.__version__) in str(g['meta'])


def _unpad(row):
    return [int(v) for v in row if v >= 0]


@pytest.mark.parametrize('tag', ['cfg2', 'cfg3', 'cfg4', 'd16', 'small'])
def test_reduce_matches_reference(golden, tag):
    g = golden('reduce_cases')
    cfg, n, m, d, ss = [int(v) for v in g[tag + '_spec']]
    exact = _same_scipy(g)
    for i in range(n):
        A, b = wl.box_cuts(1000 * cfg + i, m, d, bool(ss))
        o = orc.reduce(A, b)
        assert o['keep'] == _unpad(g[tag + '_keep'][i]), (tag, i)
        assert o['empty'] == bool(g[tag + '_empty'][i])
        assert o['n_lp'] == int(g[tag + '_nlp'][i])
        assert o['minrep'] == bool(g[tag + '_minrep'][i])
        if exact:
            assert o['r'] == g[tag + '_r'][i]
            _, bn, _ = orc.normalize_rows(o['A'], o['b'])
            assert np.array_equal(bn, g[tag + '_bout'][i][:len(bn)])
        else:
            assert abs(o['r'] - g[tag + '_r'][i]) < 1e-9


def test_named_reference_cases(golden):
    g = golden('named_cases')
    A, b = wl.unit_cube3()
    o = orc.reduce(A, b)
    assert o['keep'] == g['cube_keep'].tolist() == [0, 1, 2, 3, 4, 5]
    assert o['n_lp'] == int(g['cube_nlp']) == 7   # +1 cheby_ball on the result = SURVEY's 8
    assert abs(o['r'] - 0.5) < 1e-12 and np.allclose(o['xc'], 0.5)
    # tests/polytope_test.py:601-622
    a = np.array([[1.0, 0.1], [1.0, 0.1], [-1., 0.], [0., 1.], [0., -1.]])
    bb = np.array([50., 50.5, -40., 1., 0.])
    o = orc.reduce(a, bb)
    assert o['keep'] == g['treduce_keep'].tolist()
    An, bn, _ = orc.normalize_rows(o['A'], o['b'])
    assert np.array_equal(An, g['treduce_A']) and np.array_equal(bn, g['treduce_b'])
    l, u = orc.bounding_box(An, bn)
    np.testing.assert_allclose(l, [[40.], [0.]], rtol=1e-7, atol=1e-7)
    np.testing.assert_allclose(u, [[50.], [1.]], rtol=1e-7, atol=1e-7)
    # operations_test squares (tests/polytope_test.py:58-86, :200-238)
    Ab = np.array([[0., 1, 1], [0, -1, 0], [1, 0, 1], [-1, 0, 0]])
    Ab2 = np.array([[-1., 0, 1], [1, 0, 0], [0, 1, 1], [0, -1, 0]])
    P1 = orc.normalize_rows(Ab[:, :2], Ab[:, 2])[:2]
    P2 = orc.normalize_rows(Ab2[:, :2], Ab2[:, 2])[:2]
    P4 = orc.normalize_rows(np.array([[1., 0], [0, 1], [-1, 0], [0, -1]]),
                            np.full(4, .5))[:2]
    far = orc.normalize_rows(Ab[:, :2], Ab[:, 2] - 1e3)[:2]
    i12 = orc.intersect(*P1, *P2)
    i24 = orc.intersect(*P2, *P4)
    flags = [orc.is_fulldim(*P1), orc.is_fulldim(*P2), orc.is_fulldim(*far),
             not i12['empty'], not i24['empty']]
    assert flags == g['sq_fulldim'].astype(bool).tolist() == [True, True, False, False, True]
    for P, r, x in zip((P1, P2, P4), g['sq_cheby_r'], g['sq_cheby_x']):
        rr, xx = orc.cheby_ball(*P)
        assert abs(rr - r) < 1e-12 and np.allclose(xx, x, atol=1e-12)
    A5, b5, _ = orc.normalize_rows(i24['A'], i24['b'])
    assert np.array_equal(A5, g['sq_p5_A']) and np.array_equal(b5, g['sq_p5_b'])
    # unbounded / empty conventions (polytope.py:1372-1402, 1289-1299)
    l, u = orc.bounding_box(np.array([[1., 0.], [0., 1.], [-1, 0]]), np.ones(3))
    assert np.array_equal(l, g['unb_l']) and np.array_equal(u, g['unb_u'])
    assert orc.cheby_ball(np.array([[1., 0.]]), np.array([1.]))[0] == g['half_cheby'][0] == 0
    emp = (np.array([[1.], [-1.]]), np.array([0., -1.]))
    assert orc.cheby_ball(*emp)[0] == g['empty_cheby'][0] == 0
    l, u = orc.bounding_box(*emp)
    assert np.array_equal(l, g['empty_l']) and np.array_equal(u, g['empty_u'])


@pytest.mark.parametrize('tag', ['g2', 'g3', 'g4'])
def test_adjacency_grid(golden, tag):
    g = golden('adjacent_cases')
    A, b, idx = wl.box_grid(tuple(int(s) for s in g[tag + '_shape']))
    cells = [orc.normalize_rows(A[i], b[i])[:2] for i in range(len(A))]
    adj = orc.adjacency_matrix(cells)
    assert np.array_equal(adj, g[tag + '_adj'])


def test_adjacent_random(golden):
    g = golden('adjacent_cases')
    for i in range(60):
        A1, b1 = wl.box_cuts(9000 + i, 12, 4, True)
        A2, b2 = wl.box_cuts(9500 + i, 12, 4, True)
        rng = np.random.default_rng(77 + i)
        b2 = b2 + A2 @ (rng.uniform(-1, 1, 4) * (i % 4))
        q1 = orc.normalize_rows(A1, b1)[:2]
        q2 = orc.normalize_rows(A2, b2)[:2]
        assert orc.is_adjacent(*q1, *q2) == bool(g['rand_flag'][i]), i


def test_intersect(golden):
    g = golden('intersect_cases')
    Q = orc.normalize_rows(*wl.box_cuts(3999, 16, 6, True))[:2]
    for i in range(24):
        P = orc.normalize_rows(*wl.box_cuts(3000 + i, 16, 6, True))[:2]
        o = orc.intersect(*P, *Q)
        assert o['empty'] == bool(g['empty'][i]), i
        assert o['keep'] == _unpad(g['keep'][i]), i


def test_lp_cases(golden):
    g = golden('lp_cases')
    exact = _same_scipy(g)
    for i in range(len(g['status'])):
        m, n = g['shape'][i]
        sol = orc.lpsolve(g['C'][i, :n], g['G'][i, :m, :n], g['H'][i, :m])
        assert sol['status'] == g['status'][i]
        if sol['status'] == 0:
            if exact:
                assert sol['fun'] == g['fun'][i]
            else:
                assert abs(sol['fun'] - g['fun'][i]) < 1e-7 * (1 + abs(g['fun'][i]))

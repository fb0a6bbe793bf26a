"""GPU parity tests of the LP kernel behind `solvers.lpsolve` / `lpsolve_batch`.

Everything goes through the C ABI (libpolytope_b200.so) and is compared with
(a) the golden vectors recorded from the unmodified reference and (b) the CPU
oracle on freshly seeded inputs.  Tolerances: statuses exact; objective values
1e-7 abs + 1e-7 rel (HiGHS' own feasibility tolerance).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _group_by_n(g):
    shp = g['shape']
    for n in sorted(set(int(v) for v in shp[:, 1])):
        idx = np.nonzero(shp[:, 1] == n)[0]
        m = int(shp[idx, 0].max())
        yield n, m, idx


def test_lp_batch_matches_reference_golden(golden):
    from polytope_b200 import solvers
    g = golden('lp_cases')
    for n, m, idx in _group_by_n(g):
        C = g['C'][idx][:, :n]
        G = g['G'][idx][:, :m, :n]
        H = g['H'][idx][:, :m]
        rows = g['shape'][idx, 0].astype(np.int32)
        status, X, fun = solvers.lpsolve_batch(C, G, H, m_rows=rows)
        assert np.array_equal(status, g['status'][idx]), (n, status, g['status'][idx])
        ok = status == 0
        ref = g['fun'][idx]
        assert np.all(np.abs(fun[ok] - ref[ok]) <= 1e-7 + 1e-7 * np.abs(ref[ok])), \
            np.max(np.abs(fun[ok] - ref[ok]))
        assert np.all(np.isnan(fun[~ok])) and np.all(np.isnan(X[~ok]))
        for k in np.nonzero(ok)[0]:
            mk = rows[k]
            assert np.all(G[k, :mk] @ X[k] <= H[k, :mk] + 1e-9 * (1 + np.abs(H[k, :mk])))
            assert abs(C[k] @ X[k] - fun[k]) <= 1e-12 * (1 + abs(fun[k]))


def test_lpsolve_api_mirrors_reference_tests():
    """tests/polytope_test.py:510-548, :565-575 of the reference, on the b200 solver."""
    from polytope_b200 import solvers
    c = np.array([1, 1], dtype=float)
    A = np.array([[-1, 0], [0, -1]], dtype=float)
    b = np.array([1, 1], dtype=float)
    res = solvers.lpsolve(c, A, b)
    assert res['status'] == 0
    assert res['x'].ndim == 1 and res['x'].shape == (2,)
    np.testing.assert_allclose(res['x'], [-1, -1], atol=1e-12)
    c, A, b = np.array([1.]), np.array([[-1.]]), np.array([1.])
    res = solvers.lpsolve(c, A, b)
    assert res['x'].ndim == 1 and res['x'].shape == (1,)
    assert res['x'] == np.array([-1.0]), res['x']          # exact, as the reference pins
    assert res['fun'] == -1.0
    r = solvers.lpsolve(c, A, b, solver='b200')
    assert r['x'] == np.array([-1.0])
    with pytest.raises(RuntimeError):
        solvers.lpsolve(c, A, b, solver='glpk')
    with pytest.raises(RuntimeError):
        solvers.lpsolve(c, A, b, solver='scipy')
    with pytest.raises(Exception, match='unknown LP solver'):
        solvers.lpsolve(c, A, b, solver='nope')
    # unbounded / infeasible results carry no x, like scipy's HiGHS adapter
    r = solvers.lpsolve(np.array([1.]), np.array([[1.]]), np.array([1.]))
    assert r['status'] == 3 and r['x'] is None and r['fun'] is None
    r = solvers.lpsolve(np.array([1.]), np.array([[1.], [-1.]]), np.array([0., -1.]))
    assert r['status'] == 2 and r['x'] is None


def test_default_solver_is_reassignable():
    """solvers.default_solver is read at call time (solvers.py:95-96)."""
    from polytope_b200 import solvers
    old = solvers.default_solver
    try:
        solvers.default_solver = 'glpk'
        with pytest.raises(RuntimeError):
            solvers.lpsolve(np.array([1.]), np.array([[-1.]]), np.array([1.]))
    finally:
        solvers.default_solver = old
    assert solvers.installed_solvers == {'b200'}


@pytest.mark.parametrize('n,mmax,seed', [(1, 4, 0), (3, 12, 1), (6, 24, 2), (8, 32, 3),
                                          (9, 33, 4), (13, 64, 5), (17, 64, 6), (24, 100, 7)])
def test_lp_batch_random_vs_oracle(n, mmax, seed):
    """Random LPs incl. infeasible, unbounded and rank-deficient ones vs scipy/HiGHS."""
    from polytope_b200 import solvers
    from oracle import polytope_oracle as orc
    rng = np.random.default_rng(100 + seed)
    B = 96
    G = rng.standard_normal((B, mmax, n))
    H = rng.standard_normal((B, mmax)) * 10 ** rng.uniform(-2, 3, (B, 1))
    C = rng.standard_normal((B, n))
    rows = rng.integers(1, mmax + 1, B).astype(np.int32)
    for k in range(B):
        if k % 3 == 0:
            G[k, :, rng.integers(0, n)] = 0.0                 # rank deficient
        if k % 4 == 1:
            C[k] = G[k, :rows[k]].T @ (-rng.uniform(0, 1, rows[k]))   # dual feasible
        if k % 4 == 2:                                        # bounded & feasible: box rows
            q = min(rows[k] // 2, n)
            G[k, :2 * q] = 0
            G[k, np.arange(q), np.arange(q)] = 1
            G[k, q + np.arange(q), np.arange(q)] = -1
            H[k, :2 * q] = np.abs(H[k, :2 * q]) + 0.1
    status, X, fun = solvers.lpsolve_batch(C, G, H, m_rows=rows)
    seen = set()
    for k in range(B):
        ref = orc.lpsolve(C[k], G[k, :rows[k]], H[k, :rows[k]])
        assert status[k] == ref['status'], (k, status[k], ref['status'])
        seen.add(int(status[k]))
        if ref['status'] == 0:
            assert abs(fun[k] - ref['fun']) <= 1e-7 + 1e-7 * abs(ref['fun']), (k, fun[k], ref['fun'])
    assert 0 in seen


def test_lp_batch_device_tensors_stay_on_device():
    import torch
    from polytope_b200 import solvers
    C = torch.tensor([[1.0]], device='cuda', dtype=torch.float64)
    G = torch.tensor([[[-1.0]]], device='cuda', dtype=torch.float64)
    H = torch.tensor([[1.0]], device='cuda', dtype=torch.float64)
    status, X, fun = solvers.lpsolve_batch(C, G, H)
    assert X.is_cuda and status.is_cuda
    assert X.cpu().item() == -1.0 and status.cpu().item() == 0


def test_unsupported_sizes_fail_loudly():
    from polytope_b200 import solvers, _capi
    with pytest.raises(_capi.Pb200Error):
        solvers.lpsolve_batch(np.zeros((1, 40)), np.zeros((1, 4, 40)), np.zeros((1, 4)))
    with pytest.raises(_capi.Pb200Error):
        solvers.lpsolve_batch(np.zeros((1, 2)), np.zeros((1, 200, 2)), np.zeros((1, 200)))

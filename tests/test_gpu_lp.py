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
    with pytest.raises(_capi.Pb200Error):        # more than 64 columns: no kernel takes it, and nothing falls back
        solvers.lpsolve_batch(np.zeros((1, 70)), np.zeros((1, 4, 70)), np.zeros((1, 4)))
    # 40 columns / 200 rows used to be refused; they now go to the one-LP-per-CTA solver
    st, X, fun = solvers.lpsolve_batch(np.ones((1, 2)), np.vstack([-np.eye(2)] * 100)[None], np.ones((1, 200)))
    assert st[0] == 0 and abs(fun[0] + 2.0) < 1e-12


def _random_big_lps(rng, count, m_lo, m_hi, n_lo, n_hi):
    from scipy import optimize
    lps = []
    for k in range(count):
        n = int(rng.integers(n_lo, n_hi + 1))
        m = int(rng.integers(m_lo, m_hi + 1))
        G = rng.standard_normal((m, n))
        kind = k % 5
        if kind == 1:
            G[:, -1] = G[:, 0]                 # rank deficient
        if kind == 2:
            G[1] = -G[0]                       # infeasible pair
        h = rng.uniform(0.1, 2.0, m)
        if kind == 2:
            h[1] = -h[0] - 1.0
        if kind in (0, 3):                     # bounded: a box around the origin plus cuts
            G = np.vstack([np.eye(n), -np.eye(n), G])
            h = np.hstack([np.ones(2 * n), h])
        c = rng.standard_normal(n)
        sol = optimize.linprog(c, G, h, bounds=(None, None))
        if sol.status in (0, 2, 3):
            lps.append((c, G, h, sol.status, sol.fun if sol.status == 0 else np.nan))
    return lps


@pytest.mark.parametrize('m_lo,m_hi,n_lo,n_hi', [(130, 700, 2, 16), (40, 300, 33, 64), (1500, 4000, 5, 13)])
def test_big_lps_one_per_cta_against_highs(m_lo, m_hi, n_lo, n_hi):
    """pb200_lp_batch_big: LPs beyond 128 rows / 32 columns (every status) against scipy/HiGHS, one call per LP
    shape and ragged batches through m_rows."""
    from polytope_b200 import engine
    rng = np.random.default_rng(m_lo + n_hi)
    lps = _random_big_lps(rng, 30, m_lo, m_hi, n_lo, n_hi)
    assert {0, 2, 3} <= set(lp[3] for lp in lps)
    by_n = {}
    for lp in lps:
        by_n.setdefault(len(lp[0]), []).append(lp)
    for n, group in by_n.items():
        m = max(len(lp[2]) for lp in group)
        B = len(group)
        G = np.zeros((B, m, n))
        H = np.zeros((B, m))
        C = np.zeros((B, n))
        rows = np.zeros(B, dtype=np.int32)
        for k, (c, g, h, _, _) in enumerate(group):
            G[k, :len(h)] = g
            H[k, :len(h)] = h
            C[k] = c
            rows[k] = len(h)
        if m <= engine.LP_MAX_M and n <= engine.LP_MAX_N:      # keep the shape on the big path
            pad = engine.LP_MAX_M + 1 - m
            G = np.concatenate([G, np.zeros((B, pad, n))], 1)
            H = np.concatenate([H, np.ones((B, pad))], 1)
        status, X, fun, iters = engine.lp_batch(C, G, H, rows)
        ref_st = np.array([lp[3] for lp in group])
        ref_fun = np.array([lp[4] for lp in group])
        assert np.array_equal(status, ref_st), (n, status, ref_st)
        ok = ref_st == 0
        assert np.all(np.abs(fun[ok] - ref_fun[ok]) <= 1e-7 + 1e-7 * np.abs(ref_fun[ok])), np.abs(fun[ok] - ref_fun[ok]).max()
        for k in np.nonzero(ok)[0]:
            mk = rows[k]
            assert np.all(G[k, :mk] @ X[k] <= H[k, :mk] + 1e-9 * (1 + np.abs(H[k, :mk])))


def test_big_lps_with_a_shared_constraint_matrix():
    """G[m, n] shared by all LPs of the call (the row LPs of a reduce() with many rows)."""
    from scipy import optimize
    from polytope_b200 import engine
    rng = np.random.default_rng(3)
    n, m, B = 7, 300, 40
    G = np.vstack([np.eye(n), -np.eye(n), rng.standard_normal((m - 2 * n, n))])
    h = np.hstack([np.ones(2 * n), rng.uniform(0.3, 1.5, m - 2 * n)])
    C = rng.standard_normal((B, n))
    H = np.tile(h, (B, 1))
    H[np.arange(B), np.arange(B)] += 0.1
    status, X, fun, iters = engine.lp_batch(C, G, H)
    for k in range(B):
        sol = optimize.linprog(C[k], G, H[k], bounds=(None, None))
        assert status[k] == sol.status == 0
        assert abs(fun[k] - sol.fun) <= 1e-9 * (1 + abs(sol.fun))

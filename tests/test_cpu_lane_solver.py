"""The lane LP solver (polytope_b200/csrc/lp_lane.cuh) compiled for the CPU (tests/lane_host.cpp,
SingleLane policy) against the reference's LP results: the golden LPs recorded from the unmodified
reference (tests/golden/lp_cases.npz, every status), the row / bounding-box LPs reduce() fires
(recorded from the oracle), and random LPs with infeasible, unbounded and rank-deficient
instances against HiGHS.  This pins the arithmetic of the kernel's per-lane code without a GPU;
tests/test_gpu_*.py pin the same code inside the warp on the device."""
import shutil

import numpy as np
import pytest
from scipy import optimize

import workloads as wl
from oracle import polytope_oracle as orc
import lane_host_util as lh

pytestmark = pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')


def _check(lps, ref_status, ref_fun, tol=1e-7):
    st, it, pol, fun, X = lh.solve(lps)
    assert np.array_equal(st, ref_status), np.nonzero(st != ref_status)[0][:10]
    ok = ref_status == 0
    err = np.abs(fun[ok] - ref_fun[ok])
    assert np.all(err <= tol + tol * np.abs(ref_fun[ok])), float(err.max())
    for k in np.nonzero(ok)[0]:
        c, G, h = lps[k]
        assert np.all(G @ X[k] <= h + 1e-9 * (1 + np.abs(h)))
        assert abs(c @ X[k] - fun[k]) <= 1e-12 * (1 + abs(fun[k]))
    return it, pol


def test_golden_lps_of_the_reference(golden):
    g = golden('lp_cases')
    idx = [k for k in range(len(g['status'])) if g['shape'][k, 1] <= lh.NS]
    assert len(idx) > 100
    by_n = {}
    for k in idx:
        by_n.setdefault(int(g['shape'][k, 1]), []).append(k)
    seen = set()
    for n, ks in by_n.items():
        lps = [(g['C'][k][:n], g['G'][k][:g['shape'][k, 0], :n], g['H'][k][:g['shape'][k, 0]]) for k in ks]
        _check(lps, g['status'][ks].astype(np.int32), g['fun'][ks])
        seen |= set(int(v) for v in g['status'][ks])
    assert {0, 2, 3} <= seen


@pytest.mark.parametrize('cfg,m,d,ss,npoly', [(2, 32, 8, False, 12), (3, 16, 6, True, 16), (12, 24, 4, True, 12),
                                               (13, 12, 2, True, 12), (14, 40, 3, False, 6), (7, 64, 8, False, 3)])
def test_lps_of_reduce(cfg, m, d, ss, npoly):
    """Every LP with n = d columns that reduce() solves (bounding box + row LPs)."""
    rec = []
    orig = orc.lpsolve

    def recording(c, G, h):
        sol = orig(c, G, h)
        if G.shape[1] == d:
            rec.append((np.array(c, float), np.array(G, float), np.array(h, float), sol['status'], sol['fun']))
        return sol
    orc.lpsolve = recording
    try:
        for i in range(npoly):
            orc.reduce(*wl.box_cuts(1000 * cfg + 700 + i, m, d, ss))
    finally:
        orc.lpsolve = orig
    lps = [r[:3] for r in rec]
    st = np.array([r[3] for r in rec], dtype=np.int32)
    fun = np.array([np.nan if r[4] is None else r[4] for r in rec])
    it, pol = _check(lps, st, fun, tol=1e-9)
    assert it.mean() < 6 and pol.max() <= 5      # certified-polish attempts start at a residual level of 0.1 and repeat every 10x


def test_random_lps_every_status_against_highs():
    rng = np.random.default_rng(11)
    lps, st, fun = [], [], []
    for k in range(400):
        n = int(rng.integers(1, 9))
        m = int(rng.integers(1, 40))
        G = rng.standard_normal((m, n))
        kind = k % 5
        if kind == 1:                      # rank deficient: repeated / zero columns
            G[:, -1] = G[:, 0] if n > 1 else 0.0
        if kind == 2 and m > 1:            # infeasible pair
            G[1] = -G[0]
        h = rng.uniform(0.1, 2.0, m)
        if kind == 2 and m > 1:
            h[1] = -h[0] - 1.0
        if kind == 3:                      # bounded box around the origin plus cuts
            G = np.vstack([np.eye(n), -np.eye(n), G])
            h = np.hstack([np.ones(2 * n), h])
        c = rng.standard_normal(n)
        sol = optimize.linprog(c, G, h, bounds=(None, None))
        if sol.status not in (0, 2, 3):
            continue
        lps.append((c, G, h))
        st.append(sol.status)
        fun.append(sol.fun if sol.status == 0 else np.nan)
    st = np.array(st, dtype=np.int32)
    assert {0, 2, 3} <= set(st.tolist())
    for n in range(1, 9):
        sel = [k for k, lp in enumerate(lps) if len(lp[0]) == n]
        if sel:
            _check([lps[k] for k in sel], st[sel], np.array(fun)[sel])


@pytest.mark.parametrize('ns', [12, 16])
def test_wide_solver_random_lps_every_status_against_highs(ns):
    """lane_solve_wide (9 <= n <= 16: the factor in a lane-interleaved array, register-blocked normal matrix)."""
    rng = np.random.default_rng(5 + ns)
    lps, st, fun = [], [], []
    for k in range(160):
        n = int(rng.integers(9, ns + 1))
        m = int(rng.integers(1, 60))
        G = rng.standard_normal((m, n))
        kind = k % 5
        if kind == 1:
            G[:, -1] = G[:, 0]
        if kind == 2 and m > 1:
            G[1] = -G[0]
        h = rng.uniform(0.1, 2.0, m)
        if kind == 2 and m > 1:
            h[1] = -h[0] - 1.0
        if kind == 3:
            G = np.vstack([np.eye(n), -np.eye(n), G])
            h = np.hstack([np.ones(2 * n), h])
        c = rng.standard_normal(n)
        sol = optimize.linprog(c, G, h, bounds=(None, None))
        if sol.status not in (0, 2, 3):
            continue
        lps.append((c, G, h))
        st.append(sol.status)
        fun.append(sol.fun if sol.status == 0 else np.nan)
    st = np.array(st, dtype=np.int32)
    fun = np.array(fun)
    assert {0, 2, 3} <= set(st.tolist())
    for n in range(9, ns + 1):
        sel = [k for k, lp in enumerate(lps) if len(lp[0]) == n]
        if sel:
            s2, it, pol, f2, X = lh.solve([lps[k] for k in sel], ns=ns)
            assert np.array_equal(s2, st[sel])
            ok = st[sel] == 0
            assert np.all(np.abs(f2[ok] - fun[sel][ok]) <= 1e-7 * (1 + np.abs(fun[sel][ok])))


def test_wide_solver_on_the_lps_of_cfg4_reduce():
    """Every n = 12 LP (bounding box + rows) and the n = 13 Chebyshev LP that reduce() solves on cfg4 polytopes."""
    rec = []
    orig = orc.lpsolve

    def recording(c, G, h):
        sol = orig(c, G, h)
        rec.append((np.array(c, float), np.array(G, float), np.array(h, float), sol['status'], sol['fun']))
        return sol
    orc.lpsolve = recording
    try:
        for i in range(2):
            orc.reduce(*wl.box_cuts(4000 + i, 64, 12))
    finally:
        orc.lpsolve = orig
    for n, ns in ((12, 12), (12, 16), (13, 16)):
        sel = [r for r in rec if r[1].shape[1] == n]
        s2, it, pol, f2, X = lh.solve([r[:3] for r in sel], ns=ns)
        st = np.array([r[3] for r in sel])
        fun = np.array([np.nan if r[4] is None else r[4] for r in sel])
        assert np.array_equal(s2, st)
        ok = st == 0
        assert np.abs(f2[ok] - fun[ok]).max() <= 1e-9
        assert it.mean() < 6

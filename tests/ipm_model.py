"""numpy model of the CUDA kernel's LP algorithm (polytope_b200/csrc/lp_warp.cuh).

TESTS ONLY: this file lets the algorithm (homogeneous self-dual Mehrotra
interior point + active-set polish) be checked against scipy on the CPU box,
where no GPU exists.  The package never imports it; the product path is the
CUDA kernel and fails loudly without it.

min c'x  s.t.  Gx + s = h, s >= 0          (x free)
max -h'z s.t.  G'z + c = 0, z >= 0
embedded with (tau, kappa) so that infeasible / unbounded problems end with
tau -> 0 and a certificate (status 2 / 3 in scipy's convention).
"""
import numpy as np

MAX_ITER = 60
FEAS_TOL = 1e-9
GAP_TOL = 1e-9
STEP = 0.99
EARLY_TOL = 1e-2      # loose tolerance at which the certified polish is first tried
EARLY_NEXT = 1e-2     # a failed attempt is repeated once the residuals shrank by this factor


def _chol_solve_factory(M, n):
    """Cholesky with LIPSOL-style handling of vanishing pivots."""
    L = np.zeros((n, n))
    skip = np.zeros(n, dtype=bool)
    Mw = M.copy()
    dmax = max(np.max(np.diag(M)), 1e-300)
    for k in range(n):
        p = Mw[k, k]
        if p <= 1e-13 * M[k, k] or p <= 1e-30 * dmax or not np.isfinite(p):
            skip[k] = True
            L[k, k] = 1.0
            L[k + 1:, k] = 0.0
            continue
        L[k, k] = np.sqrt(p)
        L[k + 1:, k] = Mw[k + 1:, k] / L[k, k]
        Mw[k + 1:, k + 1:] -= np.outer(L[k + 1:, k], L[k + 1:, k])

    def solve(r):
        y = r.copy()
        for k in range(n):
            y[k] = 0.0 if skip[k] else y[k] / L[k, k]
            y[k + 1:] -= L[k + 1:, k] * y[k]
        for k in range(n - 1, -1, -1):
            y[k] = 0.0 if skip[k] else y[k] / L[k, k]
            y[:k] -= L[k, :k] * y[k]
        return y
    solve.nskip = int(skip.sum())
    return solve


def solve_lp(c, G, h, max_iter=MAX_ITER, polish=True, trace=None, early=True):
    """Returns dict(status, x, fun, iters)."""
    c = np.asarray(c, float)
    G = np.asarray(G, float)
    h = np.asarray(h, float)
    m, n = G.shape
    nh = max(1.0, np.linalg.norm(h))
    nc = max(1.0, np.linalg.norm(c))
    x = np.zeros(n)
    s = np.ones(m) * max(1.0, np.max(np.abs(h))) if m else np.ones(0)
    s = np.maximum(h, 0) + 1.0 if m else s
    z = np.ones(m)
    tau, kap = 1.0, 1.0
    status = 1
    it = 0
    lineal = False
    etol = EARLY_TOL
    for it in range(max_iter + 1):
        rx = G.T @ z + c * tau
        rz = G @ x + s - h * tau
        cx = c @ x
        hz = h @ z
        rt = cx + hz + kap
        mu = (s @ z + tau * kap) / (m + 1)
        # ---- termination tests (cvxopt conelp style) ----
        pres = np.linalg.norm(rz) / tau / nh
        dres = np.linalg.norm(rx) / tau / nc
        pcost, dcost = cx / tau, -hz / tau
        gap = (s @ z) / (tau * tau)
        if pcost < 0:
            relgap = gap / -pcost
        elif dcost > 0:
            relgap = gap / dcost
        else:
            relgap = np.inf
        if trace is not None:
            trace.append((it, pres, dres, gap, tau, kap, pcost))
        if pres <= FEAS_TOL and ((dres <= FEAS_TOL and (gap <= GAP_TOL or relgap <= GAP_TOL))
                                 or (dres <= 1e-6 and (gap <= 1e-13 or relgap <= 1e-13))):   # stalled dual residual
            status = 0
            break
        if (early and etol > 1e-7 and not lineal and pres <= etol and dres <= etol
                and (gap <= etol or relgap <= etol)):
            # the active set is usually identified long before tight convergence:
            # polish now, accept only with a full optimality certificate
            etol *= EARLY_NEXT
            ok, xp = certified_polish(c, G, h, x / tau, s / tau, z / tau)
            if ok:
                return dict(status=0, x=xp, fun=float(c @ xp), iters=it)
        if hz < 0 and np.linalg.norm(G.T @ z) / (-hz) * nh / nc <= FEAS_TOL * 1e1 and tau < 1e-3 * kap:
            status = 2
            break
        if cx < 0 and np.linalg.norm(G @ x + s) / (-cx) * nc / nh <= FEAS_TOL * 1e1 and tau < 1e-3 * kap:
            # unbounded if feasible at all: HiGHS reports 2 for an LP that is infeasible as well,
            # so the feasibility problem (c = 0) is solved first, from the start point
            lineal = True
            c = np.zeros(n)
            nc = 1.0
            x = np.zeros(n)
            s = np.maximum(h, 0) + 1.0
            z = np.ones(m)
            tau, kap = 1.0, 1.0
            continue
        if it == max_iter:
            break
        d = z / s
        M = G.T @ (d[:, None] * G)
        solve = _chol_solve_factory(M, n)
        if it == 0 and solve.nskip:
            # G is column-rank deficient.  If c has a component in null(G) the
            # LP is unbounded whenever it is feasible: solve the feasibility
            # problem (c = 0) and report 3 instead of 0.
            u = solve(c)
            res = c - G.T @ (d * (G @ u))
            if np.max(np.abs(res)) > 1e-9 * max(np.max(np.abs(c)), 1e-300):
                lineal = True
                c = np.zeros(n)
                continue

        def kkt(p, q):
            # [0 G'; G -W][u; v] = [p; q],  W = s/z
            u = solve(p + G.T @ (d * q))
            v = d * (G @ u - q)
            return u, v
        x1, z1 = kkt(-c, h)
        den = c @ x1 + h @ z1 - kap / tau      # < 0

        def direction(eta, bs, bk):
            p = -eta * rx
            q = -eta * rz - bs / z
            x2, z2 = kkt(p, q)
            dtau = (-eta * rt - bk / tau - c @ x2 - h @ z2) / den
            dx = x2 + dtau * x1
            dz = z2 + dtau * z1
            ds = (bs - s * dz) / z
            dkap = (bk - kap * dtau) / tau
            return dx, ds, dz, dtau, dkap

        def max_step(ds, dz, dtau, dkap):
            a = 1e30
            for v, dv in ((s, ds), (z, dz)):
                neg = dv < 0
                if np.any(neg):
                    a = min(a, np.min(-v[neg] / dv[neg]))
            if dtau < 0:
                a = min(a, -tau / dtau)
            if dkap < 0:
                a = min(a, -kap / dkap)
            return a
        # affine (predictor)
        dxa, dsa, dza, dta, dka = direction(1.0, -s * z, -tau * kap)
        aa = min(1.0, max_step(dsa, dza, dta, dka))
        sigma = (1 - aa) ** 3
        # combined (corrector)
        bs = -s * z + sigma * mu - dsa * dza
        bk = -tau * kap + sigma * mu - dta * dka
        dx, ds, dz, dt, dk = direction(1 - sigma, bs, bk)
        a = min(1.0, STEP * max_step(ds, dz, dt, dk))
        x = x + a * dx
        s = s + a * ds
        z = z + a * dz
        tau += a * dt
        kap += a * dk
    if lineal and status == 0:
        status = 3
    out = dict(status=status, x=None, fun=None, iters=it)
    if status == 0:
        xs, ss, zs = x / tau, s / tau, z / tau
        if polish:
            xs = polish_x(c, G, h, xs, ss, zs)
        out['x'] = xs
        out['fun'] = float(c @ xs)
    return out


def polish_x(c, G, h, x, s, z, rounds=3, delta=1e-9):
    """Snap x onto the affine hull of the optimal face: rows with z_i > s_i are
    treated as equalities and x moves by the minimum-norm correction
    (regularised normal equations, iterated).  Accept only if no other row is
    violated and the objective does not move by more than the IPM tolerance."""
    act = z > s
    if not np.any(act):
        return x
    Ga, ha = G[act], h[act]
    n = G.shape[1]
    Mp = Ga.T @ Ga
    Mp[np.diag_indices(n)] += delta * max(1.0, np.max(np.diag(Mp)))
    solve = _chol_solve_factory(Mp, n)
    xp = x.copy()
    for _ in range(rounds):
        res = ha - Ga @ xp
        xp = xp + solve(Ga.T @ res)
    scale = max(1.0, np.max(np.abs(h)))
    slack = h - G @ xp
    if np.min(slack) < -1e-9 * scale:
        return x
    if abs(c @ xp - c @ x) > 1e-6 * max(1.0, abs(c @ x)):
        return x
    return xp


def certified_polish(c, G, h, x, s, z, rounds=3, delta=1e-9):
    """Polish from a loosely converged iterate, accepted only when it is provably
    optimal: x primal feasible, the active rows tight, and y >= 0 on the active
    rows with G_B'y + c = 0 (y found by two least-norm refinement steps from z)."""
    act = z > s
    if not np.any(act):
        return False, x
    Ga, ha = G[act], h[act]
    n = G.shape[1]
    Mp = Ga.T @ Ga
    Mp[np.diag_indices(n)] += delta * max(1.0, np.max(np.diag(Mp)))
    solve = _chol_solve_factory(Mp, n)
    xp = x.copy()
    for _ in range(rounds):
        xp = xp + solve(Ga.T @ (ha - Ga @ xp))
    scale = max(1.0, np.max(np.abs(h)))
    if np.min(h - G @ xp) < -1e-9 * scale or np.max(np.abs(ha - Ga @ xp)) > 1e-9 * scale:
        return False, x
    y = z[act].copy()
    for _ in range(2):
        y = y - Ga @ solve(Ga.T @ y + c)
    rd = Ga.T @ y + c
    if np.max(np.abs(rd)) > 1e-9 * max(1.0, np.linalg.norm(c)):
        return False, x
    if np.min(y) < -1e-9 * max(1.0, np.max(y)):
        return False, x
    return True, xp

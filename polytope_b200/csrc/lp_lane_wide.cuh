// One-LP-per-lane interior-point solver for 9 <= n <= 16 columns (cfg4's row / bounding-box
// LPs with n = 12, every Chebyshev LP of d >= 8, including stage 2 of cfg2's own reduce()).
//
// Same algorithm, same tests and same termination rules as lane_solve() in lp_lane.cuh; what
// changes is where the n x n normal matrix lives.  With NS = 16 its packed triangle has 136
// entries -- 272 registers per lane -- so it cannot stay in registers next to the n-vectors:
//   * M = G'DG is accumulated in REGISTER BLOCKS of <= 46 entries (whole rows of the triangle):
//     one sweep over the rows of G per block, the fma chain of every entry stays in a register
//     for the whole sweep, and the finished block is parked in the lane's column of a
//     lane-interleaved shared-memory array (D::L(e): bank = lane, conflict free);
//   * the Cholesky factorisation runs row by row out of that array (row i in registers, one
//     shared-memory load per fma for the rows above it), the triangular solves read it once per
//     right-hand side pair;
//   * the residual sums and the three G'v products keep their own sweep (pass A).
// Shared-memory traffic on the factor is ~n^3/6 + 4 n^2 loads per iteration against
// m (n^2/2 + 9 n) fma in the row sweeps, i.e. the kernel stays bound by the fp64 pipe.
//
// Compiles as plain C++ as well (tests/lane_host.cpp, -DLANE_HOST_NS=12 / 16).
#pragma once
#include "lp_lane.cuh"

#ifndef PB200_WIDE_UNROLL_M
#define PB200_WIDE_UNROLL_M 2
#endif
#ifndef PB200_WIDE_UNROLL_A
#define PB200_WIDE_UNROLL_A 2
#endif
#ifndef PB200_WIDE_UNROLL_BDE
#define PB200_WIDE_UNROLL_BDE 2
#endif
#ifndef PB200_WIDE_DOT_ACC
#define PB200_WIDE_DOT_ACC 4
#endif
#define PBW_ROWS_M _Pragma(PBL_STR(unroll PB200_WIDE_UNROLL_M))
#define PBW_ROWS_A _Pragma(PBL_STR(unroll PB200_WIDE_UNROLL_A))
#define PBW_ROWS _Pragma(PBL_STR(unroll PB200_WIDE_UNROLL_BDE))

namespace pb200 {
namespace lane {

// dot product with PB200_WIDE_DOT_ACC independent fma chains (the chains of dotn are NS / 2 deep)
template <int NS>
PBL_FN double dotw(const double (&a)[NS], const double (&b)[NS]) {
    constexpr int NA = PB200_WIDE_DOT_ACC;
    double acc[NA];
#pragma unroll
    for (int q = 0; q < NA; ++q) acc[q] = 0.0;
#pragma unroll
    for (int j = 0; j < NS; ++j) acc[j % NA] = fma(a[j], b[j], acc[j % NA]);
    double t = acc[0];
#pragma unroll
    for (int q = 1; q < NA; ++q) t += acc[q];
    return t;
}

// rows [J0, J1) of the packed lower triangle form one register block of at most CAP entries
#if defined(__CUDACC__)
#define PBL_CX __host__ __device__ constexpr
#else
#define PBL_CX constexpr
#endif
PBL_CX int block_end(int NS, int CAP, int J0) {
    int e = 0, j = J0;
    while (j < NS && (j == J0 || e + j + 1 <= CAP)) { e += j + 1; ++j; }
    return j;
}
template <int NS>
struct WideCap { static constexpr int value = NS <= 10 ? 55 : 46; };

// L(e) <- sum_i wt(i) g_i g_i'   (packed lower triangle), block by block
template <int NS, int J0, class D, class F>
PBL_FN void form_normal(D& dat, int m, F wt) {
    if constexpr (J0 < NS) {
        constexpr int J1 = block_end(NS, WideCap<NS>::value, J0);
        constexpr int E0 = PBL_T(J0, 0), E1 = PBL_T(J1, 0);
        double acc[E1 - E0];
#pragma unroll
        for (int e = 0; e < E1 - E0; ++e) acc[e] = 0.0;
        PBW_ROWS_M
        for (int i = 0; i < m; ++i) {
            double g[NS];
            dat.row(i, g);
            const double w = wt(i);
#pragma unroll
            for (int j = J0; j < J1; ++j) {
                const double t = w * g[j];
#pragma unroll
                for (int k = 0; k <= j; ++k) acc[PBL_T(j, k) - E0] = fma(t, g[k], acc[PBL_T(j, k) - E0]);
            }
        }
#pragma unroll
        for (int e = 0; e < E1 - E0; ++e) dat.L(E0 + e) = acc[e];
        form_normal<NS, J1>(dat, m, wt);
    }
}

// In-place Cholesky of the packed matrix behind D::L (n live columns of NS), row by row.
// Same conventions as chol() of lp_lane.cuh: the diagonal ends as 1/L_kk, vanishing pivots are
// skipped (1/L_kk := 0); returns the mask of skipped pivots among the first n.
template <int NS, class D>
PBL_FN unsigned chol_mem(D& dat, int n, double add_diag) {
    double rinv[NS];
    unsigned skipped = 0;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        double r[NS];
#pragma unroll
        for (int j = 0; j <= i; ++j) r[j] = dat.L(PBL_T(i, j));
        const double dg = r[i] + add_diag;
#pragma unroll
        for (int j = 0; j < i; ++j) {
#pragma unroll
            for (int k = 0; k < j; ++k) r[j] = fma(-r[k], dat.L(PBL_T(j, k)), r[j]);
            r[j] *= rinv[j];
        }
        double p = r[i];
#pragma unroll
        for (int k = 0; k < i; ++k) p = fma(-r[k], r[k], p);
        p += add_diag;
        const bool ok = (i < n) && (p > 1e-13 * dg) && (p > 1e-290);
        rinv[i] = ok ? rsqrt_(p) : 0.0;
        skipped |= ok ? 0u : (1u << i);
#pragma unroll
        for (int j = 0; j < i; ++j) dat.L(PBL_T(i, j)) = r[j];
        dat.L(PBL_T(i, i)) = rinv[i];
    }
    return skipped & ((1u << n) - 1u);
}

// (L L') y = a for NR right-hand sides at once (every factor entry is loaded once per sweep)
template <int NS, int NR, class D>
PBL_FN void chol_solve_mem(D& dat, double (&a)[NR][NS]) {
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const double dk = dat.L(PBL_T(k, k));
#pragma unroll
        for (int r = 0; r < NR; ++r) a[r][k] *= dk;
#pragma unroll
        for (int i = k + 1; i < NS; ++i) {
            const double l = dat.L(PBL_T(i, k));
#pragma unroll
            for (int r = 0; r < NR; ++r) a[r][i] = fma(-l, a[r][k], a[r][i]);
        }
    }
#pragma unroll
    for (int k = NS - 1; k >= 0; --k) {
        const double dk = dat.L(PBL_T(k, k));
#pragma unroll
        for (int r = 0; r < NR; ++r) a[r][k] *= dk;
#pragma unroll
        for (int i = 0; i < k; ++i) {
            const double l = dat.L(PBL_T(k, i));
#pragma unroll
            for (int r = 0; r < NR; ++r) a[r][i] = fma(-l, a[r][k], a[r][i]);
        }
    }
}

// The accessor D is the one of lane_solve() plus   double& L(int e)   (NS (NS + 1) / 2 entries).
template <int NS, class D, class W>
PBL_FN void lane_solve_wide(D& dat, bool has_lp, int n, Result<NS>& res) {
    const int m = has_lp ? dat.rows() : 0;
    res.status = ITER_LIMIT; res.iters = 0; res.polishes = 0; res.fun = 0.0;
#pragma unroll
    for (int j = 0; j < NS; ++j) res.x[j] = 0.0;

    double hh = 0.0, hmax = 0.0;
    for (int i = 0; i < m; ++i) {
        const double h = dat.h(i);
        dat.s(i) = fmax(h, 0.0) + 1.0;
        dat.z(i) = 1.0;
        hh = fma(h, h, hh);
        hmax = fmax(hmax, fabs(h));
    }
    double c0[NS], x[NS];
    double cc2 = 0.0;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        c0[j] = has_lp && j < n ? dat.c(j) : 0.0;
        x[j] = 0.0;
        cc2 = fma(c0[j], c0[j], cc2);
    }
    const double nh2 = fmax(1.0, hh);
    double nc2 = fmax(1.0, cc2);
    const double rmu = 1.0 / (double)(m + 1);
    double tau = 1.0, kap = 1.0;
    bool lineal = false, ready = false;
    double etol = EARLY_TOL;
    int phase = has_lp ? PH_IPM : PH_DONE, it = 0, waited = 0;

    for (;;) {
        if (W::all(phase == PH_DONE)) break;
        if (W::any(phase == PH_IPM)) {
            if (phase == PH_IPM) {
                const double csel = lineal ? 0.0 : 1.0;
                // ---- pass A: residuals and G'[z | D h | D q_aff] ----
                double v1[NS], v2[NS], v3[NS];
#pragma unroll
                for (int j = 0; j < NS; ++j) { v1[j] = 0.0; v2[j] = 0.0; v3[j] = 0.0; }
                double sz = 0.0, hz = 0.0, rz2 = 0.0, gxs2 = 0.0, dhh = 0.0, dhq = 0.0;
                PBW_ROWS_A
                for (int i = 0; i < m; ++i) {
                    double g[NS];
                    dat.row(i, g);
                    const double h = dat.h(i);
                    const double s = dat.s(i), z = dat.z(i);
                    const double gx = dotw<NS>(g, x);
                    const double d = z * rcp(s);
                    const double gxs = gx + s;
                    const double rz = gxs - h * tau;
                    sz = fma(s, z, sz);
                    hz = fma(h, z, hz);
                    rz2 = fma(rz, rz, rz2);
                    gxs2 = fma(gxs, gxs, gxs2);
                    const double dh = d * h, dq = z - d * rz;
                    dhh = fma(dh, h, dhh);
                    dhq = fma(dh, s - rz, dhq);
#pragma unroll
                    for (int j = 0; j < NS; ++j) {
                        v1[j] = fma(z, g[j], v1[j]);
                        v2[j] = fma(dh, g[j], v2[j]);
                        v3[j] = fma(dq, g[j], v3[j]);
                    }
                }
                res.iters = it;
                double rxl[NS];
                double rx2 = 0.0, cx = 0.0;
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    rxl[j] = fma(csel * c0[j], tau, v1[j]);
                    rx2 = fma(rxl[j], rxl[j], rx2);
                    cx = fma(csel * c0[j], x[j], cx);
                }
                const double rt = cx + hz + kap;
                const double mu = (sz + tau * kap) * rmu;
                const double tinv = rcp(tau);
                const double t2 = tinv * tinv;
                const double pcost = cx * tinv, dcost = -hz * tinv;
                const double gap = sz * t2;
                const double gapref = pcost < 0.0 ? -pcost : (dcost > 0.0 ? dcost : 0.0);
                ready = false;
                bool restart = false;
                if (!(mu == mu) || !(fabs(tau) < 1e300) || !(fabs(cx) < 1e300)) {
                    res.status = NUMERICAL;
                    phase = PH_DONE;
                } else {
                    const bool converged = rz2 * t2 <= FEAS_TOL * FEAS_TOL * nh2 &&
                                           ((rx2 * t2 <= FEAS_TOL * FEAS_TOL * nc2 && (gap <= GAP_TOL || gap <= GAP_TOL * gapref)) ||
                                            (rx2 * t2 <= STALL_DRES * STALL_DRES * nc2 && (gap <= STALL_GAP || gap <= STALL_GAP * gapref)));
                    if (converged) {
                        if (lineal) { res.status = UNBOUNDED; phase = PH_DONE; }
                        else phase = PH_WAIT;
                    } else {
                        ready = etol > 1e-7 && !lineal && rz2 * t2 <= etol * etol * nh2 && rx2 * t2 <= etol * etol * nc2 &&
                                (gap <= etol || gap <= etol * gapref);
                        if (tau < 1e-3 * kap) {
                            if (hz < 0.0) {
                                const double gz2 = dotw<NS>(v1, v1);
                                if (sqrt(gz2 * nh2 / nc2) <= 10.0 * FEAS_TOL * (-hz)) { res.status = INFEASIBLE; phase = PH_DONE; }
                            }
                            if (phase == PH_IPM && cx < 0.0 && sqrt(gxs2 * nc2 / nh2) <= 10.0 * FEAS_TOL * (-cx)) {
                                lineal = true;
                                restart = true;
                                nc2 = 1.0;
                                tau = 1.0;
                                kap = 1.0;
#pragma unroll
                                for (int j = 0; j < NS; ++j) x[j] = 0.0;
                                for (int i = 0; i < m; ++i) {
                                    dat.s(i) = fmax(dat.h(i), 0.0) + 1.0;
                                    dat.z(i) = 1.0;
                                }
                                if (it == 0) it = 1;
                            }
                        }
                        if (phase == PH_IPM && it == MAX_ITER) phase = PH_DONE;
                    }
                }
                if (phase == PH_IPM && !restart) {
                    // ---- normal matrix (register blocks) and factor ----
                    form_normal<NS, 0>(dat, m, [&](int i) { return dat.z(i) * rcp(dat.s(i)); });
                    const unsigned skipped = chol_mem<NS>(dat, n, 0.0);
                    if (it == 0 && skipped && !lineal) {
                        double uu[1][NS], back[NS];
#pragma unroll
                        for (int j = 0; j < NS; ++j) { uu[0][j] = c0[j]; back[j] = 0.0; }
                        chol_solve_mem<NS, 1>(dat, uu);
                        for (int i = 0; i < m; ++i) {
                            double g[NS];
                            dat.row(i, g);
                            const double d = dat.z(i) * rcp(dat.s(i));
                            const double dgu = d * dotw<NS>(g, uu[0]);
#pragma unroll
                            for (int j = 0; j < NS; ++j) back[j] = fma(dgu, g[j], back[j]);
                        }
                        double rmax = 0.0, cmax = 0.0;
#pragma unroll
                        for (int j = 0; j < NS; ++j) { rmax = fmax(rmax, fabs(c0[j] - back[j])); cmax = fmax(cmax, fabs(c0[j])); }
                        if (rmax > 1e-9 * fmax(cmax, 1e-300)) {
                            lineal = true;
                            nc2 = 1.0;
                            it = 1;
                            restart = true;
                        }
                    }
                }
                if (phase == PH_IPM && !restart) {
                    // ---- predictor: K [x1; z1] = [-c; h],  K [x2; z2] = [-rx; q_aff] ----
                    double XX[2][NS];
                    double (&X1)[NS] = XX[0];
                    double (&xa)[NS] = XX[1];
#pragma unroll
                    for (int j = 0; j < NS; ++j) { X1[j] = fma(-csel, c0[j], v2[j]); xa[j] = v3[j] - rxl[j]; }
                    chol_solve_mem<NS, 2>(dat, XX);
                    double cx1 = 0.0, cx2 = 0.0;
#pragma unroll
                    for (int j = 0; j < NS; ++j) { cx1 = fma(csel * c0[j], X1[j], cx1); cx2 = fma(csel * c0[j], xa[j], cx2); }
                    const double hz1 = dotw<NS>(v2, X1) - dhh;
                    const double hz2 = dotw<NS>(v2, xa) - dhq;
                    const double kot = kap * tinv;
                    const double den = cx1 + hz1 - kot;
                    const double rden = rcp(den);
                    const double dta = (-rt + kap - cx2 - hz2) * rden;
                    const double dka = -kap - kot * dta;
                    const double kinv = rcp(kap);
#pragma unroll
                    for (int j = 0; j < NS; ++j) xa[j] = fma(dta, X1[j], xa[j]);
                    double ratio = fmax(fmax(-dta * tinv, -dka * kinv), 0.0);
                    PBW_ROWS
                    for (int i = 0; i < m; ++i) {            // pass B: affine step length
                        double g[NS];
                        dat.row(i, g);
                        const double h = dat.h(i);
                        const double q = h * tau - dotw<NS>(g, x);
                        const double w = rcp(dat.s(i)) * (dotw<NS>(g, xa) - q - dta * h);
                        ratio = fmax(ratio, fmax(1.0 + w, -w));
                    }
                    const double alpha_aff = ratio > 1.0 ? rcp(ratio) : 1.0;
                    const double om = 1.0 - alpha_aff;
                    const double sigma = om * om * om;
                    const double eta = 1.0 - sigma;
                    const double sm = sigma * mu;
                    // ---- corrector right-hand side (pass C) ----
                    double X3s[1][NS];
                    double (&X3)[NS] = X3s[0];
                    double dhqc = 0.0;
#pragma unroll
                    for (int j = 0; j < NS; ++j) X3[j] = 0.0;
                    PBW_ROWS
                    for (int i = 0; i < m; ++i) {
                        double g[NS];
                        dat.row(i, g);
                        const double h = dat.h(i);
                        const double s = dat.s(i), z = dat.z(i);
                        const double isz = rcp(s * z);
                        const double sinv = isz * z, zinv = isz * s;
                        const double d = z * sinv;
                        const double q = h * tau - dotw<NS>(g, x);
                        const double rz = s - q;
                        const double dza = d * (dotw<NS>(g, xa) - q - dta * h);
                        const double dsa = -s - s * zinv * dza;
                        const double bs = -s * z + sm - dsa * dza;
                        const double qc = -eta * rz - bs * zinv;
                        const double dqc = d * qc;
                        dhqc = fma(dqc, h, dhqc);
#pragma unroll
                        for (int j = 0; j < NS; ++j) X3[j] = fma(dqc, g[j], X3[j]);
                    }
#pragma unroll
                    for (int j = 0; j < NS; ++j) X3[j] = fma(-eta, rxl[j], X3[j]);
                    chol_solve_mem<NS, 1>(dat, X3s);
                    const double v2x3 = dotw<NS>(v2, X3);
                    double cx3 = 0.0;
#pragma unroll
                    for (int j = 0; j < NS; ++j) cx3 = fma(csel * c0[j], X3[j], cx3);
                    const double hz3 = v2x3 - dhqc;
                    const double bk = -tau * kap + sm - dta * dka;
                    const double dtau = (-eta * rt - bk * tinv - cx3 - hz3) * rden;
                    const double dkap = (bk - kap * dtau) * tinv;
#pragma unroll
                    for (int j = 0; j < NS; ++j) X3[j] = fma(dtau, X1[j], X3[j]);
                    ratio = fmax(fmax(-dtau * tinv, -dkap * kinv), 0.0);
                    PBW_ROWS
                    for (int i = 0; i < m; ++i) {            // pass D: step length
                        double g[NS];
                        dat.row(i, g);
                        const double h = dat.h(i);
                        const double s = dat.s(i), z = dat.z(i);
                        const double isz = rcp(s * z);
                        const double sinv = isz * z, zinv = isz * s;
                        const double d = z * sinv;
                        const double q = h * tau - dotw<NS>(g, x);
                        const double rz = s - q;
                        const double dza = d * (dotw<NS>(g, xa) - q - dta * h);
                        const double dsa = -s - s * zinv * dza;
                        const double bs = -s * z + sm - dsa * dza;
                        const double qc = -eta * rz - bs * zinv;
                        const double dz = d * (dotw<NS>(g, X3) - qc - dtau * h);
                        const double ds = (bs - s * dz) * zinv;
                        ratio = fmax(ratio, fmax(-ds * sinv, -dz * zinv));
                    }
                    const double amax = ratio > 0.0 ? rcp(ratio) : 1e30;
                    const double alpha = fmin(1.0, STEP * amax);
                    PBW_ROWS
                    for (int i = 0; i < m; ++i) {            // pass E: take the step in (s, z)
                        double g[NS];
                        dat.row(i, g);
                        const double h = dat.h(i);
                        const double s = dat.s(i), z = dat.z(i);
                        const double isz = rcp(s * z);
                        const double sinv = isz * z, zinv = isz * s;
                        const double d = z * sinv;
                        const double q = h * tau - dotw<NS>(g, x);
                        const double rz = s - q;
                        const double dza = d * (dotw<NS>(g, xa) - q - dta * h);
                        const double dsa = -s - s * zinv * dza;
                        const double bs = -s * z + sm - dsa * dza;
                        const double qc = -eta * rz - bs * zinv;
                        const double dz = d * (dotw<NS>(g, X3) - qc - dtau * h);
                        const double ds = (bs - s * dz) * zinv;
                        dat.s(i) = fma(alpha, ds, s);
                        dat.z(i) = fma(alpha, dz, z);
                    }
#pragma unroll
                    for (int j = 0; j < NS; ++j) x[j] = fma(alpha, X3[j], x[j]);
                    tau = fma(alpha, dtau, tau);
                    kap = fma(alpha, dkap, kap);
                    ++it;
                }
            }
        }
        // ================= polish (same schedule as lane_solve) =================
        const bool cand = (phase == PH_IPM && ready) || phase == PH_WAIT;
        const bool busy = phase == PH_IPM && !ready;
        if (cand) ++waited;
        const bool go = W::any(cand) && (!W::any_busy(busy) || W::any(cand && waited > MAX_WAIT));
        if (go && cand) {
            const bool early = phase == PH_IPM;
            ++res.polishes;
            waited = 0;
            ready = false;
            if (early) etol *= EARLY_NEXT;
            const double te = 1.0 / tau;
            double xp[NS];
#pragma unroll
            for (int j = 0; j < NS; ++j) xp[j] = x[j] * te;
            int nact = 0;
            for (int i = 0; i < m; ++i) nact += dat.z(i) > dat.s(i) ? 1 : 0;
            double f0 = 0.0;
            if (!early) {
                f0 = dotw<NS>(c0, xp);
                res.status = OPTIMAL;
                res.fun = f0;
#pragma unroll
                for (int j = 0; j < NS; ++j) res.x[j] = xp[j];
                phase = PH_DONE;
            }
            if (nact > 0) {
                form_normal<NS, 0>(dat, m, [&](int i) { return dat.z(i) > dat.s(i) ? 1.0 : 0.0; });
                double dmax = 1.0;
#pragma unroll
                for (int j = 0; j < NS; ++j) dmax = fmax(dmax, dat.L(PBL_T(j, j)));
                chol_mem<NS>(dat, n, 1e-9 * dmax);
                double U[NS];
#pragma unroll
                for (int j = 0; j < NS; ++j) U[j] = 0.0;
                const double scale = fmax(1.0, hmax);
                for (int round = 0; round < PB200_LANE_ROUNDS; ++round) {
                    double vv[2][NS];
                    double (&vp)[NS] = vv[0];
                    double (&vd)[NS] = vv[1];
#pragma unroll
                    for (int j = 0; j < NS; ++j) { vp[j] = 0.0; vd[j] = 0.0; }
                    double ft = 0.0, fslack = -1e300, fymin = -1e300, fymax = 0.0;
                    PBW_ROWS
                    for (int i = 0; i < m; ++i) {
                        double g[NS];
                        dat.row(i, g);
                        const double rr = dat.h(i) - dotw<NS>(g, xp);
                        fslack = fmax(fslack, -rr);
                        const double z = dat.z(i);
                        if (z > dat.s(i)) {
                            const double y = z * te - dotw<NS>(g, U);
                            ft = fmax(ft, fabs(rr));
                            fymin = fmax(fymin, -y);
                            fymax = fmax(fymax, y);
#pragma unroll
                            for (int j = 0; j < NS; ++j) { vp[j] = fma(rr, g[j], vp[j]); vd[j] = fma(y, g[j], vd[j]); }
                        }
                    }
                    double frd = 0.0;
#pragma unroll
                    for (int j = 0; j < NS; ++j) { vd[j] += c0[j]; frd = fmax(frd, fabs(vd[j])); }
                    const bool feasible = fslack <= 1e-9 * scale;
                    const bool settled = ft <= 1e-13 * scale || round == PB200_LANE_ROUNDS - 1;
                    if (!early) {
                        if (settled) {
                            const double f1 = dotw<NS>(c0, xp);
                            if (feasible && fabs(f1 - f0) <= 1e-6 * fmax(1.0, fabs(f0))) {
                                res.fun = f1;
#pragma unroll
                                for (int j = 0; j < NS; ++j) res.x[j] = xp[j];
                            }
                            break;
                        }
                    } else {
                        const bool dual_ok = frd <= 1e-9 * sqrt(nc2) && fymin <= 1e-9 * fmax(1.0, fymax);
                        if (settled && feasible && ft <= 1e-9 * scale && dual_ok) {
                            res.status = OPTIMAL;
                            res.fun = dotw<NS>(c0, xp);
#pragma unroll
                            for (int j = 0; j < NS; ++j) res.x[j] = xp[j];
                            phase = PH_DONE;
                            break;
                        }
                        if (round == PB200_LANE_ROUNDS - 1) break;
                    }
                    chol_solve_mem<NS, 2>(dat, vv);
#pragma unroll
                    for (int j = 0; j < NS; ++j) { xp[j] += vp[j]; U[j] += vd[j]; }
                }
            }
        }
    }
}

}  // namespace lane
}  // namespace pb200

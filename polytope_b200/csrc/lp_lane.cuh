// One-LP-per-LANE interior-point solver for LP families that share their constraint
// matrix: the row LPs of reduce() (polytope.py:1142-1160, ~25 LPs per polytope that differ
// only in c = -A[k] and one entry of h) and the 2d LPs of bounding_box (:1362-1411).
//
// Same algorithm as lp_warp_small.cuh (tests/ipm_model.py is its model: Mehrotra
// predictor-corrector on the homogeneous self-dual embedding, certified active-set polish),
// but mapped the other way round: a lane owns a whole LP and walks over the rows of G, which
// sit once in shared memory and are read by broadcast.  Nothing is replicated and nothing is
// reduced across lanes: the 8x8 normal matrix, its Cholesky factor and every n-vector live in
// the lane's registers, the per-row iterates (s_i, z_i) in a lane-interleaved shared-memory
// array.  The r01 profile of the warp-per-LP kernel showed the fp64 pipe 33 % busy for 2.5 %
// algorithmic flops, i.e. ~13x redundant work (every lane factoring the same matrix) -- here
// a warp instruction advances 32 different LPs.
//
// The file compiles as plain C++ too (tests/lane_host.cpp drives it on the CPU against
// HiGHS), so everything warp-specific goes through the policy class W.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PBL_FN __host__ __device__ __forceinline__
#else
#define PBL_FN inline
#endif

namespace pb200 {
namespace lane {

constexpr int MAX_ITER = 60;
constexpr double FEAS_TOL = 1e-9;
constexpr double GAP_TOL = 1e-9;
#ifndef PB200_LANE_STEP
#define PB200_LANE_STEP 0.999      // fraction of the step to the boundary (r02ap: 0.99 -> 0.999 is 4 % on cfg2; the certified exit tolerates it)
#endif
constexpr double STEP = PB200_LANE_STEP;
#ifndef PB200_LANE_ROUNDS
#define PB200_LANE_ROUNDS 3        // projection / refinement rounds of one polish attempt
#endif
constexpr double STALL_DRES = 1e-6;   // see the termination test
constexpr double STALL_GAP = 1e-13;
#ifndef PB200_LANE_EARLY_TOL
#define PB200_LANE_EARLY_TOL 1e-1
#endif
constexpr double EARLY_TOL = PB200_LANE_EARLY_TOL;   // residual level of the first certified-polish attempt
#ifndef PB200_LANE_EARLY_NEXT
#define PB200_LANE_EARLY_NEXT 1e-1
#endif
constexpr double EARLY_NEXT = PB200_LANE_EARLY_NEXT;  // a failed attempt is repeated after this much progress
#ifndef PB200_LANE_MAX_WAIT
#define PB200_LANE_MAX_WAIT 3
#endif
constexpr int MAX_WAIT = PB200_LANE_MAX_WAIT;         // iterations a polish-ready lane waits for the others
enum : int { OPTIMAL = 0, ITER_LIMIT = 1, INFEASIBLE = 2, UNBOUNDED = 3, NUMERICAL = 4 };   // scipy codes
enum : int { PH_IPM = 0, PH_WAIT = 1, PH_DONE = 2 };

PBL_FN double rcp(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
#else
    return 1.0 / x;
#endif
}
PBL_FN double rsqrt_(double p) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
    const double hp = 0.5 * p;
    y = y * fma(-hp * y, y, 1.5);
    y = y * fma(-hp * y, y, 1.5);
    return y;
#else
    return 1.0 / sqrt(p);
#endif
}

// warp policies: how lanes agree on the phase of the shared instruction stream
struct SingleLane {
    static PBL_FN bool any(bool p) { return p; }
    static PBL_FN bool all(bool p) { return p; }
    static PBL_FN bool any_busy(bool p) { return p; }
};
// a lane whose warp always has some other lane still iterating: every polish attempt is
// preceded by the full wait (CPU model of the worst case inside a warp)
struct WaitingLane {
    static PBL_FN bool any(bool p) { return p; }
    static PBL_FN bool all(bool p) { return p; }
    static PBL_FN bool any_busy(bool) { return true; }
};
#if defined(__CUDACC__)
struct WarpLanes {
    static __device__ __forceinline__ bool any(bool p) { return __any_sync(0xffffffffu, p); }
    static __device__ __forceinline__ bool all(bool p) { return __all_sync(0xffffffffu, p); }
    static __device__ __forceinline__ bool any_busy(bool p) { return __any_sync(0xffffffffu, p); }
};
#endif

#ifndef PB200_LANE_UNROLL
#define PB200_LANE_UNROLL 2
#endif
// Per-family tuning of lane_solve (r02ab / r02ak A/B builds on cfg2, profiles/r02ak_lane_early_polish_ab.txt).
//   kUnroll  unroll factor of the row loops (the row-LP kernel is instruction-fetch bound: 1)
//   kEarly   residual level at which the certified polish is first tried.  The certificate makes any
//            attempt safe (a point is only accepted as a complementary primal-dual optimal pair), so this
//            is purely a speed knob: 1e-2 -> 1e-1 cut the mean iterations per LP from 4.4 to 3.5
//   kNext    a failed attempt is repeated after this much further progress
struct DefaultTune {
    static constexpr int kUnroll = PB200_LANE_UNROLL;
    static constexpr double kEarly = EARLY_TOL;
    static constexpr double kNext = EARLY_NEXT;
};

template <int NS>
struct Result {
    int status, iters, polishes;
    double fun;
    double x[NS];
};

// packed lower triangle, row i, column j <= i
#define PBL_T(i, j) ((i) * ((i) + 1) / 2 + (j))

// In-place Cholesky of the packed matrix L (n live columns of NS).  The diagonal ends as
// 1/L_kk; vanishing pivots are skipped LIPSOL-style (1/L_kk := 0, the solution component is
// forced to 0).  Returns the bitmask of skipped pivots among the first n.
template <int NS>
PBL_FN unsigned chol(double (&L)[NS * (NS + 1) / 2], int n, double add_diag) {
    double dg[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) dg[k] = L[PBL_T(k, k)] + add_diag;
    unsigned skipped = 0;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const double p = L[PBL_T(k, k)] + add_diag;
        const bool ok = (k < n) && (p > 1e-13 * dg[k]) && (p > 1e-290);
        const double rinv = ok ? rsqrt_(p) : 0.0;
        skipped |= ok ? 0u : (1u << k);
        L[PBL_T(k, k)] = rinv;
#pragma unroll
        for (int i = k + 1; i < NS; ++i) L[PBL_T(i, k)] *= rinv;
#pragma unroll
        for (int i = k + 1; i < NS; ++i)
#pragma unroll
            for (int j = k + 1; j <= i; ++j) L[PBL_T(i, j)] = fma(-L[PBL_T(i, k)], L[PBL_T(j, k)], L[PBL_T(i, j)]);
    }
    return skipped & ((1u << n) - 1u);
}

template <int NS>
PBL_FN void chol_solve(const double (&L)[NS * (NS + 1) / 2], double (&a)[NS]) {
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        a[k] *= L[PBL_T(k, k)];
#pragma unroll
        for (int i = k + 1; i < NS; ++i) a[i] = fma(-L[PBL_T(i, k)], a[k], a[i]);
    }
#pragma unroll
    for (int k = NS - 1; k >= 0; --k) {
        a[k] *= L[PBL_T(k, k)];
#pragma unroll
        for (int i = 0; i < k; ++i) a[i] = fma(-L[PBL_T(k, i)], a[k], a[i]);
    }
}

template <int NS>
PBL_FN double dotn(const double (&a)[NS], const double (&b)[NS]) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int j = 0; j + 1 < NS; j += 2) { s0 = fma(a[j], b[j], s0); s1 = fma(a[j + 1], b[j + 1], s1); }
    if (NS & 1) s0 = fma(a[NS - 1], b[NS - 1], s0);
    return s0 + s1;
}

// The data accessor D of a lane provides
//   int rows()                       rows of this lane's LP
//   void row(int i, double (&g)[NS]) row i of G (zero-padded to NS columns)
//   double h(int i)                  right-hand side, finite (an absent row is staged as 0'x <= 1)
//   double c(int j)                  objective
//   double& s(int i), double& z(int i)   the lane's slack / multiplier iterates
//
// lane_solve: the LP of this lane (has_lp == false: the lane idles through the collectives).
// n = live columns (<= NS).
//
// Per iteration the rows are walked five times (A: residuals + normal matrix + right-hand sides,
// B: affine step length, C: corrector right-hand side, D: step length, E: step); every pass
// recomputes what it needs of a row from (g_i, h_i, s_i, z_i) and a few n-vectors, so the only
// per-row state is (s_i, z_i).  The sums h'z_k of the KKT solutions come from n-vector dot
// products (h'D G x = (G'D h)'x) instead of further passes.
#ifndef PB200_LANE_UNROLL
#define PB200_LANE_UNROLL 2
#endif
#define PBL_STR2(x) #x
#define PBL_STR(x) PBL_STR2(x)
#define PBL_ROWS _Pragma(PBL_STR(unroll PB200_LANE_UNROLL))
// inside lane_solve the row loops unroll by its template parameter UR: the row-LP kernel is instruction-fetch
// sensitive (ncu: no_instruction 0.87 cycles per issue) and runs 8 % faster with UR = 1, the bounding-box kernel
// (no_instruction 0.06) 9 % slower (r02ab A/B builds, profiles/r02ab_lane_unroll_ab.txt)
#define PBL_ROWS_UR _Pragma("unroll UR")

template <int NS, class D, class W, class Tune = DefaultTune>
PBL_FN void lane_solve(D& dat, bool has_lp, int n, Result<NS>& res) {
    constexpr int UR = Tune::kUnroll;
    constexpr int NT = NS * (NS + 1) / 2;
    const int m = has_lp ? dat.rows() : 0;
    res.status = ITER_LIMIT; res.iters = 0; res.polishes = 0; res.fun = 0.0;
#pragma unroll
    for (int j = 0; j < NS; ++j) res.x[j] = 0.0;

    // ---- start point (Mehrotra-style, as lp_warp_small.cuh) ----
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
    bool lineal = false, ready = false;     // lineal: the objective has been dropped (feasibility problem, c = 0)
    double etol = Tune::kEarly;
    int phase = has_lp ? PH_IPM : PH_DONE, it = 0, waited = 0;

    for (;;) {
        if (W::all(phase == PH_DONE)) break;
        // ================= one interior-point iteration =================
        if (W::any(phase == PH_IPM)) {
            if (phase == PH_IPM) {
                const double csel = lineal ? 0.0 : 1.0;       // effective objective = csel * c0
                // ---- pass A: residuals, normal matrix M = G'DG and G'[z | D h | D q_aff] ----
                double M[NT], v1[NS], v2[NS], v3[NS];
#pragma unroll
                for (int e = 0; e < NT; ++e) M[e] = 0.0;
#pragma unroll
                for (int j = 0; j < NS; ++j) { v1[j] = 0.0; v2[j] = 0.0; v3[j] = 0.0; }
                double sz = 0.0, hz = 0.0, rz2 = 0.0, gxs2 = 0.0, dhh = 0.0, dhq = 0.0;
                PBL_ROWS_UR
                for (int i = 0; i < m; ++i) {
                    double g[NS];
                    dat.row(i, g);
                    const double h = dat.h(i);
                    const double s = dat.s(i), z = dat.z(i);
                    const double gx = dotn<NS>(g, x);
                    const double d = z * rcp(s);
                    const double gxs = gx + s;
                    const double rz = gxs - h * tau;
                    sz = fma(s, z, sz);
                    hz = fma(h, z, hz);
                    rz2 = fma(rz, rz, rz2);
                    gxs2 = fma(gxs, gxs, gxs2);
                    const double dh = d * h, dq = z - d * rz;
                    dhh = fma(dh, h, dhh);                  // h'D h
                    dhq = fma(dh, s - rz, dhq);             // h'D q_aff,  q_aff = s - rz
#pragma unroll
                    for (int j = 0; j < NS; ++j) {
                        const double t = d * g[j];
#pragma unroll
                        for (int k = 0; k <= j; ++k) M[PBL_T(j, k)] = fma(t, g[k], M[PBL_T(j, k)]);
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
                // ---- termination (cvxopt conelp-style tests, squared where a norm is involved) ----
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
                    // (near-parallel active rows leave the dual residual stuck around 1e-8 while the gap
                    // keeps shrinking: with the gap four orders below its tolerance a dual residual of
                    // 1e-6 -- cvxopt's and HiGHS' own feasibility tolerance is 1e-7 -- is accepted)
                    const bool converged = rz2 * t2 <= FEAS_TOL * FEAS_TOL * nh2 &&
                                           ((rx2 * t2 <= FEAS_TOL * FEAS_TOL * nc2 && (gap <= GAP_TOL || gap <= GAP_TOL * gapref)) ||
                                            (rx2 * t2 <= STALL_DRES * STALL_DRES * nc2 && (gap <= STALL_GAP || gap <= STALL_GAP * gapref)));
                    if (converged) {
                        if (lineal) { res.status = UNBOUNDED; phase = PH_DONE; }
                        else phase = PH_WAIT;                 // final polish without certificate
                    } else {
                        // at a loose tolerance the active set is usually identified already: the lane
                        // is ready for a certified polish, and keeps iterating until the warp goes
                        ready = etol > 1e-7 && !lineal && rz2 * t2 <= etol * etol * nh2 && rx2 * t2 <= etol * etol * nc2 &&
                                (gap <= etol || gap <= etol * gapref);
                        if (tau < 1e-3 * kap) {
                            if (hz < 0.0) {
                                const double gz2 = dotn<NS>(v1, v1);
                                if (sqrt(gz2 * nh2 / nc2) <= 10.0 * FEAS_TOL * (-hz)) { res.status = INFEASIBLE; phase = PH_DONE; }
                            }
                            if (phase == PH_IPM && cx < 0.0 && sqrt(gxs2 * nc2 / nh2) <= 10.0 * FEAS_TOL * (-cx)) {
                                // an improving recession direction: unbounded if the LP is feasible at all.
                                // HiGHS reports 2 for an LP that is infeasible as well, so feasibility is
                                // settled first: restart on the feasibility problem (c = 0), which ends with
                                // 3 (feasible) or with the infeasibility certificate
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
                        if (phase == PH_IPM && it == MAX_ITER) phase = PH_DONE;      // status stays ITER_LIMIT
                    }
                }
                if (phase == PH_IPM && !restart) {
                    // ---- factor ----
                    const unsigned skipped = chol<NS>(M, n, 0.0);
                    if (it == 0 && skipped && !lineal) {
                        // G is column-rank deficient: if c has a component in null(G) the LP is unbounded
                        // whenever it is feasible -> continue with c = 0 and report 3 instead of 0
                        double uu[NS], back[NS];
#pragma unroll
                        for (int j = 0; j < NS; ++j) { uu[j] = c0[j]; back[j] = 0.0; }
                        chol_solve<NS>(M, uu);
                        for (int i = 0; i < m; ++i) {
                            double g[NS];
                            dat.row(i, g);
                            const double d = dat.z(i) * rcp(dat.s(i));
                            const double dgu = d * dotn<NS>(g, uu);
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
                    double X1[NS], xa[NS];
#pragma unroll
                    for (int j = 0; j < NS; ++j) { X1[j] = fma(-csel, c0[j], v2[j]); xa[j] = v3[j] - rxl[j]; }
                    chol_solve<NS>(M, X1);
                    chol_solve<NS>(M, xa);
                    double cx1 = 0.0, cx2 = 0.0;
#pragma unroll
                    for (int j = 0; j < NS; ++j) { cx1 = fma(csel * c0[j], X1[j], cx1); cx2 = fma(csel * c0[j], xa[j], cx2); }
                    // h'z1 and h'z2 of the two KKT solutions z_k = D (G x_k - q_k) without a pass over the
                    // rows: h'D G x_k = v2'x_k, and h'D h, h'D q_aff were accumulated in pass A
                    const double hz1 = dotn<NS>(v2, X1) - dhh;
                    const double hz2 = dotn<NS>(v2, xa) - dhq;
                    const double kot = kap * tinv;
                    const double den = cx1 + hz1 - kot;                            // < 0
                    const double rden = rcp(den);
                    const double dta = (-rt + kap - cx2 - hz2) * rden;
                    const double dka = -kap - kot * dta;
                    const double kinv = rcp(kap);
                    // affine direction in x (without its dtau part): x2 + dta x1
#pragma unroll
                    for (int j = 0; j < NS; ++j) xa[j] = fma(dta, X1[j], xa[j]);
                    double ratio = fmax(fmax(-dta * tinv, -dka * kinv), 0.0);
                    PBL_ROWS_UR
                    for (int i = 0; i < m; ++i) {            // pass B: affine step length
                        double g[NS];
                        dat.row(i, g);
                        const double h = dat.h(i);
                        const double q = h * tau - dotn<NS>(g, x);
                        // w = dza / z;  -dsa / s = 1 + w
                        const double w = rcp(dat.s(i)) * (dotn<NS>(g, xa) - q - dta * h);
                        ratio = fmax(ratio, fmax(1.0 + w, -w));
                    }
                    const double alpha_aff = ratio > 1.0 ? rcp(ratio) : 1.0;
                    const double om = 1.0 - alpha_aff;
                    const double sigma = om * om * om;
                    const double eta = 1.0 - sigma;
                    const double sm = sigma * mu;
                    // ---- corrector right-hand side (pass C) ----
                    double X3[NS];
                    double dhqc = 0.0;
#pragma unroll
                    for (int j = 0; j < NS; ++j) X3[j] = 0.0;
                    PBL_ROWS_UR
                    for (int i = 0; i < m; ++i) {
                        double g[NS];
                        dat.row(i, g);
                        const double h = dat.h(i);
                        const double s = dat.s(i), z = dat.z(i);
                        const double isz = rcp(s * z);
                        const double sinv = isz * z, zinv = isz * s;
                        const double d = z * sinv;
                        const double q = h * tau - dotn<NS>(g, x);
                        const double rz = s - q;
                        const double dza = d * (dotn<NS>(g, xa) - q - dta * h);
                        const double dsa = -s - s * zinv * dza;
                        const double bs = -s * z + sm - dsa * dza;
                        const double qc = -eta * rz - bs * zinv;
                        const double dqc = d * qc;
                        dhqc = fma(dqc, h, dhqc);          // h'D q_cor
#pragma unroll
                        for (int j = 0; j < NS; ++j) X3[j] = fma(dqc, g[j], X3[j]);
                    }
#pragma unroll
                    for (int j = 0; j < NS; ++j) X3[j] = fma(-eta, rxl[j], X3[j]);
                    chol_solve<NS>(M, X3);
                    const double v2x3 = dotn<NS>(v2, X3);
                    double cx3 = 0.0;
#pragma unroll
                    for (int j = 0; j < NS; ++j) cx3 = fma(csel * c0[j], X3[j], cx3);
                    const double hz3 = v2x3 - dhqc;                   // h'D (G x3 - q_cor), as above
                    const double bk = -tau * kap + sm - dta * dka;
                    const double dtau = (-eta * rt - bk * tinv - cx3 - hz3) * rden;
                    const double dkap = (bk - kap * dtau) * tinv;
                    // final direction in x: x3 + dtau x1
#pragma unroll
                    for (int j = 0; j < NS; ++j) X3[j] = fma(dtau, X1[j], X3[j]);
                    ratio = fmax(fmax(-dtau * tinv, -dkap * kinv), 0.0);
                    PBL_ROWS_UR
                    for (int i = 0; i < m; ++i) {            // pass D: step length
                        double g[NS];
                        dat.row(i, g);
                        const double h = dat.h(i);
                        const double s = dat.s(i), z = dat.z(i);
                        const double isz = rcp(s * z);
                        const double sinv = isz * z, zinv = isz * s;
                        const double d = z * sinv;
                        const double q = h * tau - dotn<NS>(g, x);
                        const double rz = s - q;
                        const double dza = d * (dotn<NS>(g, xa) - q - dta * h);
                        const double dsa = -s - s * zinv * dza;
                        const double bs = -s * z + sm - dsa * dza;
                        const double qc = -eta * rz - bs * zinv;
                        const double dz = d * (dotn<NS>(g, X3) - qc - dtau * h);
                        const double ds = (bs - s * dz) * zinv;
                        ratio = fmax(ratio, fmax(-ds * sinv, -dz * zinv));
                    }
                    const double amax = ratio > 0.0 ? rcp(ratio) : 1e30;
                    const double alpha = fmin(1.0, STEP * amax);
                    PBL_ROWS_UR
                    for (int i = 0; i < m; ++i) {            // pass E: take the step in (s, z)
                        double g[NS];
                        dat.row(i, g);
                        const double h = dat.h(i);
                        const double s = dat.s(i), z = dat.z(i);
                        const double isz = rcp(s * z);
                        const double sinv = isz * z, zinv = isz * s;
                        const double d = z * sinv;
                        const double q = h * tau - dotn<NS>(g, x);
                        const double rz = s - q;
                        const double dza = d * (dotn<NS>(g, xa) - q - dta * h);
                        const double dsa = -s - s * zinv * dza;
                        const double bs = -s * z + sm - dsa * dza;
                        const double qc = -eta * rz - bs * zinv;
                        const double dz = d * (dotn<NS>(g, X3) - qc - dtau * h);
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
        // ================= polish: when no lane is still on its way, or one has waited long enough =================
        const bool cand = (phase == PH_IPM && ready) || phase == PH_WAIT;
        const bool busy = phase == PH_IPM && !ready;
        if (cand) ++waited;
        const bool go = W::any(cand) && (!W::any_busy(busy) || W::any(cand && waited > MAX_WAIT));
        if (go && cand) {
            const bool early = phase == PH_IPM;      // loosely converged: accept only with a full optimality certificate
            ++res.polishes;
            waited = 0;
            ready = false;
            if (early) etol *= Tune::kNext;
            const double te = 1.0 / tau;
            double xp[NS];
#pragma unroll
            for (int j = 0; j < NS; ++j) xp[j] = x[j] * te;
            // rows with z > s are taken as the optimal face
            int nact = 0;
            for (int i = 0; i < m; ++i) nact += dat.z(i) > dat.s(i) ? 1 : 0;
            double f0 = 0.0;
            if (!early) {
                f0 = dotn<NS>(c0, xp);
                res.status = OPTIMAL;
                res.fun = f0;
#pragma unroll
                for (int j = 0; j < NS; ++j) res.x[j] = xp[j];
                phase = PH_DONE;                      // whatever the polish says, the LP is solved
            }
            if (nact > 0) {
                double M[NT];
#pragma unroll
                for (int e = 0; e < NT; ++e) M[e] = 0.0;
                for (int i = 0; i < m; ++i) {
                    if (!(dat.z(i) > dat.s(i))) continue;
                    double g[NS];
                    dat.row(i, g);
#pragma unroll
                    for (int j = 0; j < NS; ++j)
#pragma unroll
                        for (int k = 0; k <= j; ++k) M[PBL_T(j, k)] = fma(g[j], g[k], M[PBL_T(j, k)]);
                }
                double dmax = 1.0;
#pragma unroll
                for (int j = 0; j < NS; ++j) dmax = fmax(dmax, M[PBL_T(j, j)]);
                chol<NS>(M, n, 1e-9 * dmax);
                double U[NS];
#pragma unroll
                for (int j = 0; j < NS; ++j) U[j] = 0.0;
                const double scale = fmax(1.0, hmax);
                for (int round = 0; round < PB200_LANE_ROUNDS; ++round) {
                    // one projection round of the primal point onto the active face and one
                    // least-squares refinement of the multipliers y = z/tau - G_B U share the pass
                    double vp[NS], vd[NS];
#pragma unroll
                    for (int j = 0; j < NS; ++j) { vp[j] = 0.0; vd[j] = 0.0; }
                    double ft = 0.0, fslack = -1e300, fymin = -1e300, fymax = 0.0;
                    PBL_ROWS_UR
                    for (int i = 0; i < m; ++i) {
                        double g[NS];
                        dat.row(i, g);
                        const double rr = dat.h(i) - dotn<NS>(g, xp);
                        fslack = fmax(fslack, -rr);
                        const double z = dat.z(i);
                        if (z > dat.s(i)) {
                            const double y = z * te - dotn<NS>(g, U);
                            ft = fmax(ft, fabs(rr));
                            fymin = fmax(fymin, -y);
                            fymax = fmax(fymax, y);
#pragma unroll
                            for (int j = 0; j < NS; ++j) { vp[j] = fma(rr, g[j], vp[j]); vd[j] = fma(y, g[j], vd[j]); }
                        }
                    }
                    double frd = 0.0;
#pragma unroll
                    for (int j = 0; j < NS; ++j) { vd[j] += c0[j]; frd = fmax(frd, fabs(vd[j])); }      // |G_B'y + c|
                    const bool feasible = fslack <= 1e-9 * scale;
                    const bool settled = ft <= 1e-13 * scale || round == PB200_LANE_ROUNDS - 1;
                    if (!early) {
                        if (settled) {
                            const double f1 = dotn<NS>(c0, xp);
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
                            // primal feasible, active rows tight, dual feasible, complementary: optimal
                            res.status = OPTIMAL;
                            res.fun = dotn<NS>(c0, xp);
#pragma unroll
                            for (int j = 0; j < NS; ++j) res.x[j] = xp[j];
                            phase = PH_DONE;
                            break;
                        }
                        if (round == PB200_LANE_ROUNDS - 1) break;       // not certified: the interior-point iterations resume
                    }
                    chol_solve<NS>(M, vp);
                    chol_solve<NS>(M, vd);
#pragma unroll
                    for (int j = 0; j < NS; ++j) { xp[j] += vp[j]; U[j] += vd[j]; }
                }
            }
        }
    }
}

}  // namespace lane
}  // namespace pb200

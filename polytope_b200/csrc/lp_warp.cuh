// One-LP-per-warp fp64 interior-point solver for sm_100a.
//
//   min c'x  s.t.  Gx <= h,  x free          (polytope/solvers.py:76-106 contract)
//
// Algorithm: Mehrotra predictor-corrector on the homogeneous self-dual
// embedding (so infeasible / unbounded LPs end with a certificate and map to
// scipy's status 2 / 3), normal equations M = G' D G factored by Cholesky, then
// an active-set "polish" that snaps x onto the affine hull of the optimal face
// so objective values are good to ~1e-15 (the reference's simplex returns
// vertices; its callers compare against 1e-7 / 1e-8 thresholds).
// tests/ipm_model.py is the numpy model of exactly this algorithm.
//
// Mapping to the warp:
//   * rows of G are owned by lanes: lane l holds rows l, l+32, ... (RPL each)
//     in registers (h, s, z and every per-row temporary);
//   * n-vectors (c, x, rx, dx ...) are owned by lanes 0..n-1, one component each;
//   * G lives in shared memory, column-major with leading dimension MP = 32*RPL+4
//     (== 4 mod 16): both the row-owner access (fixed column, lane = row) and the
//     DMMA fragment access (4 consecutive rows x 8 columns) are bank-conflict free;
//   * every reduction over rows that produces more than a scalar -- M = G'DG and
//     G'[v1 v2 v3] -- is an fp64 tensor-core contraction (mma.sync m8n8k4 f64,
//     the only fp64 MMA sm_100a has; tcgen05 has no fp64 kind).  Measured on
//     B200 (profiles/r01_microbench_*.txt): one DMMA costs 4 SM-cycles, one
//     64-bit shuffle 2, so a shuffle tree per entry would be ~10x dearer;
//   * Cholesky and the triangular solves run out of shared memory, lane i owning
//     row i of L; scalars are xor-butterfly reductions, bitwise identical in all
//     lanes, so control flow stays warp-uniform.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace pb200 {

constexpr unsigned FULL_MASK = 0xffffffffu;
// Every branch of the solvers is warp-uniform (all lanes hold bitwise identical
// scalars), but ptxas cannot prove it for conditions computed from butterfly
// reductions.  Passing such a condition through a vote makes the branch provably
// uniform, which removes the divergence bookkeeping (BSSY/BSYNC, the
// WARPSYNC.COLLECTIVE slow path around every shuffle) from the loop body.
#define PB_UNI(cond) __all_sync(FULL_MASK, (cond))
constexpr int LP_MAX_ITER = 60;
constexpr int LP_MAX_N = 32;            // columns (n-vectors are lane-owned)
constexpr double LP_FEAS_TOL = 1e-9;
constexpr double LP_GAP_TOL = 1e-9;
constexpr double LP_STEP = 0.99;
// Near-parallel active rows leave the dual residual stuck around 1e-8 (rounding amplified by the
// conditioning of the active set) while the gap keeps shrinking: once the gap is four orders
// below its tolerance, a dual residual of 1e-6 is accepted (cvxopt's and HiGHS' own feasibility
// tolerance is 1e-7); the final polish still checks primal feasibility and the objective.
constexpr double LP_STALL_DRES = 1e-6;
constexpr double LP_STALL_GAP = 1e-13;
#ifndef PB200_EARLY_TOL
#define PB200_EARLY_TOL 1e-2
#endif
constexpr double LP_EARLY_TOL = PB200_EARLY_TOL;      // tolerance at which the certified polish is first tried
constexpr double LP_EARLY_NEXT = 1e-2;     // a failed attempt is repeated once the residuals shrank by this factor
constexpr int NSLOT = 4;                // vector slots of a G'V pass

// scipy.optimize.linprog status codes (polytope/solvers.py:92-93)
enum : int { ST_OPTIMAL = 0, ST_ITER_LIMIT = 1, ST_INFEASIBLE = 2, ST_UNBOUNDED = 3, ST_NUMERICAL = 4 };

#ifndef PB200_REDUCE_INLINE
#define PB200_REDUCE_INLINE __forceinline__
#endif
__device__ PB200_REDUCE_INLINE double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ PB200_REDUCE_INLINE double warp_max(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ PB200_REDUCE_INLINE double warp_min(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
// reciprocal / reciprocal square root: MUFU seed + two Newton steps (<= 1-2 ulp);
// the IEEE-exact sequences cost ~20-30 instructions each and the r01 profile
// showed them at ~12 % of all issue slots.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}
__device__ __forceinline__ double fast_rsqrt(double p) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
    const double hp = 0.5 * p;
    y = y * fma(-hp * y, y, 1.5);
    y = y * fma(-hp * y, y, 1.5);
    return y;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Per-warp shared-memory scratch.  All sizes in doubles.
struct WarpScratch {
    double* G;   // [n][MP]   column-major constraint matrix, rows >= m zero
    double* M;   // [NP][LDM] normal matrix / Cholesky factor (diag holds 1/L_kk)
    double* V;   // [NSLOT][MP] row-vector slots (B operand of the G'V pass)
    double* d;   // [MP]      row weights of the M pass
    double* u;   // [3][NP]   broadcast n-vectors for G u products
    double* R;   // [NSLOT][NP] results of the G'V pass
    double* hb;  // [MP]      staged right-hand side, kept across LPs that share G
    int MP, NP, LDM;
    int NC;      // columns allocated in G (what staging must zero)
};

__host__ __device__ inline int lp_mp(int rpl) { return 32 * rpl + 4; }
__host__ __device__ inline int lp_np(int n) { return (n + 7) & ~7; }
// doubles of scratch one warp needs for an (RPL, n) problem
__host__ __device__ inline int lp_scratch_doubles(int rpl, int n) {
    const int MP = lp_mp(rpl), NP = lp_np(n), LDM = NP + 1;
    return n * MP + NP * LDM + NSLOT * MP + MP + 3 * NP + NSLOT * NP + MP;
}
__device__ inline WarpScratch lp_carve(double* base, int rpl, int n) {
    WarpScratch w;
    w.MP = lp_mp(rpl); w.NP = lp_np(n); w.LDM = w.NP + 1; w.NC = n;
    w.G = base;            base += n * w.MP;
    w.M = base;            base += w.NP * w.LDM;
    w.V = base;            base += NSLOT * w.MP;
    w.d = base;            base += w.MP;
    w.u = base;            base += 3 * w.NP;
    w.R = base;            base += NSLOT * w.NP;
    w.hb = base;
    return w;
}

// ---- G u for NV broadcast vectors (slots slot0.. of w.u), rows in registers ----
template <int RPL, int NV>
__device__ __forceinline__ void rows_times(const WarpScratch& w, int n, int slot0, int lane,
                                           double (&out)[NV][RPL]) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int r = 0; r < RPL; ++r) out[v][r] = 0.0;
    for (int j = 0; j < n; ++j) {
        double g[RPL];
#pragma unroll
        for (int r = 0; r < RPL; ++r) g[r] = w.G[j * w.MP + lane + 32 * r];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const double uv = w.u[(slot0 + v) * w.NP + j];
#pragma unroll
            for (int r = 0; r < RPL; ++r) out[v][r] = fma(g[r], uv, out[v][r]);
        }
    }
}

// ---- R[slot][j] = sum_i G[i][j] V[slot][i]  (fp64 tensor cores) ----
__device__ __forceinline__ void gt_times_slots(const WarpScratch& w, int mk, int n, int nslots, int lane) {
    const int q = lane >> 2, t = lane & 3;
    for (int J = 0; J * 8 < n; ++J) {
        double c0 = 0.0, c1 = 0.0;
        const int col = 8 * J + q;
        const bool ca = col < n, cb = q < nslots;
        const double* ga = w.G + col * w.MP + t;
        const double* vb = w.V + q * w.MP + t;
        // whole 32-row blocks (rows >= m are zero in G, V and d): the inner 8 steps unroll
        for (int b0 = 0; b0 < mk; b0 += 32) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i0 = b0 + 4 * u;
                const double a = ca ? ga[i0] : 0.0;
                const double b = cb ? vb[i0] : 0.0;
                dmma884(c0, c1, a, b);
            }
        }
        if (2 * t < NSLOT) {
            w.R[(2 * t) * w.NP + col] = c0;
            w.R[(2 * t + 1) * w.NP + col] = c1;
        }
    }
    __syncwarp();
}

// ---- M = G' diag(d) G, lower block triangle, fp64 tensor cores ----
__device__ __forceinline__ void form_normal_matrix(const WarpScratch& w, int mk, int n, int lane) {
    const int q = lane >> 2, t = lane & 3;
    for (int J = 0; J * 8 < n; ++J)
        for (int K = 0; K <= J; ++K) {
            double c0 = 0.0, c1 = 0.0;
            const int cj = 8 * J + q, ck = 8 * K + q;
            const bool bj = cj < n, bk = ck < n;
            const double* gj = w.G + cj * w.MP + t;
            const double* gk = w.G + ck * w.MP + t;
            const double* dd = w.d + t;
            for (int b0 = 0; b0 < mk; b0 += 32) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i0 = b0 + 4 * u;
                    const double a = bj ? gj[i0] * dd[i0] : 0.0;
                    const double b = bk ? gk[i0] : 0.0;
                    dmma884(c0, c1, a, b);
                }
            }
            double* dst = w.M + cj * w.LDM + 8 * K + 2 * t;
            dst[0] = c0;
            dst[1] = c1;
        }
    __syncwarp();
}

// ---- in-place Cholesky of M[0..n) (lower), lane i owns row i.  Vanishing
// pivots are skipped LIPSOL-style (their solution component is forced to 0);
// returns the bitmask of skipped pivots.  M[k][k] ends holding 1/L_kk (0 if skipped).
__device__ __forceinline__ unsigned cholesky(const WarpScratch& w, int n, int lane) {
    const int L = w.LDM;
    const double dg = lane < n ? w.M[lane * L + lane] : 0.0;
    const double dmax = fmax(warp_max(dg), 1e-300);
    unsigned skipped = 0;
    for (int k = 0; k < n; ++k) {
        const double p = w.M[k * L + k];
        const double p0 = __shfl_sync(FULL_MASK, dg, k);
        __syncwarp();      // every lane has read the pivot before lane k overwrites it (racecheck, profiles/r02a_*)
        const bool ok = (p > 1e-13 * p0) && (p > 1e-30 * dmax) && (p > 1e-290) && (p < 1e300);
        double lik = 0.0;
        if (ok) {
            const double rinv = fast_rsqrt(p);
            if (lane > k && lane < n) {
                lik = w.M[lane * L + k] * rinv;
                w.M[lane * L + k] = lik;
            }
            if (lane == k) w.M[k * L + k] = rinv;
        } else {
            skipped |= 1u << k;
            if (lane > k && lane < n) w.M[lane * L + k] = 0.0;
            if (lane == k) w.M[k * L + k] = 0.0;
        }
        __syncwarp();
        if (ok) {
            for (int j = k + 1; j < n; ++j) {
                const double ljk = w.M[j * L + k];
                if (lane >= j && lane < n) w.M[lane * L + j] = fma(-lik, ljk, w.M[lane * L + j]);
            }
        }
        __syncwarp();
    }
    return skipped;
}

// ---- solve (L L') y = r for NR lane-owned right-hand sides, in place ----
template <int NR>
__device__ __forceinline__ void chol_solve(const WarpScratch& w, int n, int lane, double (&r)[NR]) {
    const int L = w.LDM;
    for (int k = 0; k < n; ++k) {
        const double dinv = w.M[k * L + k];
        const double lik = (lane > k && lane < n) ? w.M[lane * L + k] : 0.0;
#pragma unroll
        for (int v = 0; v < NR; ++v) {
            const double yk = __shfl_sync(FULL_MASK, r[v], k) * dinv;
            r[v] = (lane == k) ? yk : fma(-lik, yk, r[v]);
        }
    }
    for (int k = n - 1; k >= 0; --k) {
        const double dinv = w.M[k * L + k];
        const double lki = (lane < k) ? w.M[k * L + lane] : 0.0;
#pragma unroll
        for (int v = 0; v < NR; ++v) {
            const double xk = __shfl_sync(FULL_MASK, r[v], k) * dinv;
            r[v] = (lane == k) ? xk : fma(-lki, xk, r[v]);
        }
    }
#pragma unroll
    for (int v = 0; v < NR; ++v)
        if (lane >= n) r[v] = 0.0;
}

struct LpResult {
    int status;
    int iters;
    double x;      // lane-owned component of the solution (lanes < n), valid if status == 0
    double fun;    // c'x, valid if status == 0
};

// Active-set polish of an interior-point iterate, optionally with an optimality
// certificate.  Rows with z > s are taken as the optimal face; x is moved by the
// minimum-norm correction onto their affine hull (3 regularised normal-equation
// rounds).  certify == false (tightly converged iterate): accept if no row is
// violated and the objective agrees with the iterate's.  certify == true (loosely
// converged iterate): accept only if additionally the active rows are tight and
// the least-squares-refined multipliers y (two rounds of y -= G_B (G_B'G_B)^-1
// (G_B'y + c)) are dual feasible: y >= 0, G_B'y + c = 0 -- i.e. (xp, y) is a
// complementary primal-dual optimal pair.  Overwrites w.M, w.d, w.V, w.u, w.R.
template <int RPL>
__device__ __forceinline__ bool polish_active_set(const WarpScratch& w, int mk, int n, int lane, bool own, double c_orig, double nc,
                                                  double hmax, const double (&h)[RPL], const bool (&live)[RPL],
                                                  const double (&s)[RPL], const double (&z)[RPL], double tau, double xs,
                                                  bool certify, double& xp_out, double& fp_out) {
    const int NP = w.NP;
    bool act[RPL];
    int nact = 0;
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        act[r] = live[r] && (z[r] > s[r]);
        nact += act[r] ? 1 : 0;
        w.d[lane + 32 * r] = act[r] ? 1.0 : 0.0;
    }
    nact = __reduce_add_sync(FULL_MASK, nact);
    __syncwarp();
    if (PB_UNI(nact == 0)) return false;
    const double f0 = warp_sum(c_orig * xs);
    form_normal_matrix(w, mk, n, lane);
    {
        const double dg = own ? w.M[lane * w.LDM + lane] : 0.0;
        const double reg = 1e-9 * fmax(1.0, warp_max(dg));
        if (own) w.M[lane * w.LDM + lane] = dg + reg;
        __syncwarp();
    }
    cholesky(w, n, lane);
    double xp = xs;
    double gxp[1][RPL];
    double y[RPL];
#pragma unroll
    for (int r = 0; r < RPL; ++r) y[r] = 0.0;
    const double te = 1.0 / tau;
    double ymin = 1e300, ymax = 0.0;
    bool ok = true;
    // rounds 0-2: projection; round 3: checks (+ start of the certificate); 4, 5: refinement
#pragma unroll 1
    for (int round = 0; round < 6; ++round) {
        if (round <= 3) {
            if (lane < NP) w.u[lane] = own ? xp : 0.0;
        }
        __syncwarp();
        rows_times<RPL, 1>(w, n, 0, lane, gxp);
        if (round == 1 || round == 2) {
            // the projection usually lands on the face after one or two rounds: skip the rest
            double t = 0.0;
#pragma unroll
            for (int r = 0; r < RPL; ++r)
                if (act[r]) t = fmax(t, fabs(h[r] - gxp[0][r]));
            t = warp_max(t);
            if (PB_UNI(t <= 1e-13 * fmax(1.0, hmax))) round = 3;
        }
        if (round < 3) {
#pragma unroll
            for (int r = 0; r < RPL; ++r) w.V[lane + 32 * r] = act[r] ? h[r] - gxp[0][r] : 0.0;
        } else if (round == 3) {
            double slack = 1e300, tight = 0.0;
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                if (live[r]) slack = fmin(slack, h[r] - gxp[0][r]);
                if (act[r]) tight = fmax(tight, fabs(h[r] - gxp[0][r]));
            }
            slack = warp_min(slack);
            const bool feasible = slack >= -1e-9 * fmax(1.0, hmax);
            const double f1 = warp_sum(c_orig * xp);
            xp_out = xp;
            fp_out = f1;
            if (PB_UNI(!certify)) return feasible && (fabs(f1 - f0) <= 1e-6 * fmax(1.0, fabs(f0)));
            tight = warp_max(tight);
            if (PB_UNI(!(feasible && tight <= 1e-9 * fmax(1.0, hmax)))) return false;
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                y[r] = act[r] ? z[r] * te : 0.0;
                w.V[lane + 32 * r] = y[r];
            }
        } else {
            ymin = 1e300; ymax = 0.0;
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                if (act[r]) { y[r] -= gxp[0][r]; ymin = fmin(ymin, y[r]); ymax = fmax(ymax, y[r]); }
                w.V[lane + 32 * r] = y[r];
            }
        }
        __syncwarp();
        gt_times_slots(w, mk, n, 1, lane);
        if (round >= 4) {
            // refined multipliers: dual feasible already?  (usually after the first refinement)
            const double rd = own ? fabs(w.R[lane] + c_orig) : 0.0;      // |G_B'y + c|
            const double rdmax = warp_max(rd);
            const double ylo = warp_min(ymin), yhi = warp_max(ymax);
            ok = rdmax <= 1e-9 * nc && ylo >= -1e-9 * fmax(1.0, yhi);
            if (PB_UNI(ok || round == 5)) break;
        }
        double rhs[1] = {own ? (round < 3 ? w.R[lane] : w.R[lane] + c_orig) : 0.0};
        chol_solve<1>(w, n, lane, rhs);
        if (round < 3) xp += rhs[0];
        else if (lane < NP) w.u[lane] = own ? rhs[0] : 0.0;       // next round multiplies G by u
    }
    return ok;
}

// Solve the LP whose G is staged in w.G (rows >= m zero).  `c` is the lane-owned
// objective component (0 for lanes >= n); h[r] the right-hand side of row
// lane+32r (ignored for rows >= m; +inf means "no constraint").
template <int RPL>
__device__ LpResult lp_solve_warp(const WarpScratch& w, int m, int n, double c, const double (&h_in)[RPL]) {
    const int lane = threadIdx.x & 31;
    const int MP = w.MP, NP = w.NP;
    const int mk = (m + 3) & ~3;                 // rows covered by the DMMA k-loop
    const bool own = lane < n;
    const double c_orig = c;

    bool live[RPL];
    double h[RPL], s[RPL], z[RPL];
    int mlive = 0;
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int i = lane + 32 * r;
        live[r] = (i < m) && (h_in[r] < 1e300);
        h[r] = live[r] ? h_in[r] : 0.0;
        s[r] = live[r] ? fmax(h[r], 0.0) + 1.0 : 1.0;
        z[r] = live[r] ? 1.0 : 0.0;
        mlive += live[r] ? 1 : 0;
    }
    mlive = __reduce_add_sync(FULL_MASK, mlive);
    double hh = 0.0, hmax = 0.0;
#pragma unroll
    for (int r = 0; r < RPL; ++r) { hh = fma(h[r], h[r], hh); hmax = fmax(hmax, fabs(h[r])); }
    const double nh = fmax(1.0, sqrt(warp_sum(hh)));
    hmax = warp_max(hmax);
    double nc = fmax(1.0, sqrt(warp_sum(c * c)));
    double x = 0.0, tau = 1.0, kap = 1.0;
    bool lineal = false;
    double etol = LP_EARLY_TOL;            // tolerance of the next certified-polish attempt
    LpResult res;
    res.status = ST_ITER_LIMIT; res.iters = 0; res.x = 0.0; res.fun = 0.0;

#pragma unroll 1
    for (int it = 0; it <= LP_MAX_ITER; ++it) {
        res.iters = it;
        // ---- residuals ----
        if (lane < NP) w.u[lane] = own ? x : 0.0;
        __syncwarp();
        double gx[1][RPL];
        rows_times<RPL, 1>(w, n, 0, lane, gx);
        double rz[RPL], d[RPL], sinv[RPL], zinv[RPL];
        double sz = 0.0, hz = 0.0, rz2 = 0.0, gxs2 = 0.0;
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const int i = lane + 32 * r;
            sinv[r] = fast_rcp(s[r]);
            zinv[r] = live[r] ? fast_rcp(z[r]) : 0.0;
            d[r] = z[r] * sinv[r];
            rz[r] = live[r] ? gx[0][r] + s[r] - h[r] * tau : 0.0;
            const double gxs = live[r] ? gx[0][r] + s[r] : 0.0;
            sz = fma(s[r], z[r], sz);
            hz = fma(h[r], z[r], hz);
            rz2 = fma(rz[r], rz[r], rz2);
            gxs2 = fma(gxs, gxs, gxs2);
            w.d[i] = d[r];
            w.V[0 * MP + i] = z[r];
            w.V[1 * MP + i] = d[r] * h[r];
            w.V[2 * MP + i] = z[r] - d[r] * rz[r];      // d * q_aff,  q_aff = s - rz
        }
        __syncwarp();
        gt_times_slots(w, mk, n, 3, lane);
        const double gz = own ? w.R[0 * NP + lane] : 0.0;    // (G'z)_j
        const double gh = own ? w.R[1 * NP + lane] : 0.0;    // (G'Dh)_j
        const double gq = own ? w.R[2 * NP + lane] : 0.0;    // (G'D q_aff)_j
        const double rx = gz + c * tau;
        sz = warp_sum(sz); hz = warp_sum(hz); rz2 = warp_sum(rz2);
        const double cx = warp_sum(c * x);
        const double rx2 = warp_sum(rx * rx);
        const double rt = cx + hz + kap;
        const double mu = (sz + tau * kap) / (double)(mlive + 1);
        // ---- termination (cvxopt conelp-style tests) ----
        const double tinv = fast_rcp(tau);
        const double pres = sqrt(rz2) * tinv / nh;
        const double dres = sqrt(rx2) * tinv / nc;
        const double pcost = cx * tinv, dcost = -hz * tinv;
        const double gap = sz * tinv * tinv;
        double relgap = 1e300;
        if (pcost < 0.0) relgap = gap / -pcost;
        else if (dcost > 0.0) relgap = gap / dcost;
        if (PB_UNI(!(mu == mu) || !(fabs(tau) < 1e300) || !(fabs(cx) < 1e300))) { res.status = ST_NUMERICAL; break; }
        const bool converged = pres <= LP_FEAS_TOL && ((dres <= LP_FEAS_TOL && (gap <= LP_GAP_TOL || relgap <= LP_GAP_TOL)) ||
                                                       (dres <= LP_STALL_DRES && (gap <= LP_STALL_GAP || relgap <= LP_STALL_GAP)));
        // At the loose tolerance the active set is usually already identified: try the
        // polish there and accept it only with a full optimality certificate (primal
        // feasible, active rows tight, y >= 0 with G_B'y + c = 0); otherwise keep
        // iterating to the tight tolerance (same scheme as lp_warp_small.cuh).
        const bool loosely = etol > 1e-7 && !lineal && pres <= etol && dres <= etol && (gap <= etol || relgap <= etol);
        if (PB_UNI(converged || loosely)) {
            if (PB_UNI(converged && lineal)) { res.status = ST_UNBOUNDED; break; }
            etol *= LP_EARLY_NEXT;
            double xp, fp;
            const double xs = x / tau;
            const bool ok = polish_active_set<RPL>(w, mk, n, lane, own, c_orig, nc, hmax, h, live, s, z, tau, xs, !converged, xp, fp);
            if (PB_UNI(converged)) {
                res.status = ST_OPTIMAL;
                res.x = ok ? xp : xs;
                res.fun = ok ? fp : warp_sum(c_orig * xs);
                break;
            }
            if (PB_UNI(ok)) { res.status = ST_OPTIMAL; res.x = xp; res.fun = fp; break; }
            // not certified: restore the row weights the polish overwrote and go on
#pragma unroll
            for (int r = 0; r < RPL; ++r) w.d[lane + 32 * r] = d[r];
            __syncwarp();
        }
        if (PB_UNI(tau < 1e-3 * kap)) {
            if (PB_UNI(hz < 0.0)) {
                const double gz2 = warp_sum(gz * gz);
                if (PB_UNI(sqrt(gz2) / (-hz) * nh / nc <= 10.0 * LP_FEAS_TOL)) { res.status = ST_INFEASIBLE; break; }
            }
            if (PB_UNI(cx < 0.0)) {
                gxs2 = warp_sum(gxs2);
                if (PB_UNI(sqrt(gxs2) / (-cx) * nc / nh <= 10.0 * LP_FEAS_TOL)) {
                    // improving recession direction: unbounded if feasible at all.  HiGHS reports 2 for
                    // an LP that is infeasible as well, so restart on the feasibility problem (c = 0):
                    // it ends with 3 (feasible) or with the infeasibility certificate
                    lineal = true;
                    c = 0.0;
                    nc = 1.0;
                    x = 0.0; tau = 1.0; kap = 1.0;
#pragma unroll
                    for (int r = 0; r < RPL; ++r) {
                        s[r] = live[r] ? fmax(h[r], 0.0) + 1.0 : 1.0;
                        z[r] = live[r] ? 1.0 : 0.0;
                    }
                    continue;
                }
            }
        }
        if (it == LP_MAX_ITER) break;
        // ---- factor M = G' D G ----
        form_normal_matrix(w, mk, n, lane);
        const unsigned skipped = cholesky(w, n, lane);
        if (PB_UNI(it == 0 && skipped)) {
            // G is column-rank deficient.  If c has a component in null(G) the LP
            // is unbounded whenever it is feasible: continue with the feasibility
            // problem (c = 0) and report 3 instead of 0.
            double uu[1] = {c};
            chol_solve<1>(w, n, lane, uu);
            if (lane < NP) w.u[lane] = own ? uu[0] : 0.0;
            __syncwarp();
            double gu[1][RPL];
            rows_times<RPL, 1>(w, n, 0, lane, gu);
#pragma unroll
            for (int r = 0; r < RPL; ++r) w.V[lane + 32 * r] = d[r] * gu[0][r];
            __syncwarp();
            gt_times_slots(w, mk, n, 1, lane);
            const double rc = own ? c - w.R[lane] : 0.0;
            const double rmax = warp_max(fabs(rc)), cmax = warp_max(fabs(c));
            if (PB_UNI(rmax > 1e-9 * fmax(cmax, 1e-300))) {
                lineal = true;
                c = 0.0;
                nc = 1.0;
                continue;
            }
        }
        // ---- the two KKT solves that share rhs-independent data ----
        //   K [x1; z1] = [-c; h],   K [x2; z2] = [-rx; q_aff]
        double sol[2] = {-c + gh, -rx + gq};
        chol_solve<2>(w, n, lane, sol);
        const double x1 = sol[0];
        if (lane < NP) { w.u[NP + lane] = own ? sol[0] : 0.0; w.u[2 * NP + lane] = own ? sol[1] : 0.0; }
        __syncwarp();
        double g12[2][RPL];
        rows_times<RPL, 2>(w, n, 1, lane, g12);
        double z1[RPL], dza[RPL], dsa[RPL];
        double hz1 = 0.0, hz2 = 0.0;
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            z1[r] = d[r] * (g12[0][r] - h[r]);
            const double z2 = d[r] * (g12[1][r] - (s[r] - rz[r]));
            hz1 = fma(h[r], z1[r], hz1);
            hz2 = fma(h[r], z2, hz2);
            dza[r] = z2;                                      // completed below
        }
        hz1 = warp_sum(hz1); hz2 = warp_sum(hz2);
        const double cx1 = warp_sum(c * x1);
        double cx2 = warp_sum(c * sol[1]);
        const double den = cx1 + hz1 - kap * tinv;             // < 0
        const double rden = fast_rcp(den), kinv = fast_rcp(kap);
        const double dta = (-rt + kap - cx2 - hz2) * rden;
        const double dka = -kap - kap * dta * tinv;
        double ratio = fmax(-dta * tinv, -dka * kinv);
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            dza[r] = fma(dta, z1[r], dza[r]);
            dsa[r] = -s[r] - s[r] * zinv[r] * dza[r];         // (-s z - s dz)/z
            if (live[r]) ratio = fmax(ratio, fmax(-dsa[r] * sinv[r], -dza[r] * zinv[r]));
        }
        ratio = warp_max(ratio);
        const double alpha_aff = ratio > 1.0 ? 1.0 / ratio : 1.0;
        const double om = 1.0 - alpha_aff;
        const double sigma = om * om * om;
        const double eta = 1.0 - sigma;
        // ---- corrector ----
        double bs[RPL], qc[RPL];
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            bs[r] = -s[r] * z[r] + sigma * mu - dsa[r] * dza[r];
            qc[r] = live[r] ? -eta * rz[r] - bs[r] * zinv[r] : 0.0;
            w.V[lane + 32 * r] = d[r] * qc[r];
        }
        __syncwarp();
        gt_times_slots(w, mk, n, 1, lane);
        double solc[1] = {-eta * rx + (own ? w.R[lane] : 0.0)};
        chol_solve<1>(w, n, lane, solc);
        if (lane < NP) w.u[2 * NP + lane] = own ? solc[0] : 0.0;
        __syncwarp();
        double g2[1][RPL];
        rows_times<RPL, 1>(w, n, 2, lane, g2);
        double dz[RPL];
        hz2 = 0.0;
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            dz[r] = d[r] * (g2[0][r] - qc[r]);
            hz2 = fma(h[r], dz[r], hz2);
        }
        hz2 = warp_sum(hz2);
        cx2 = warp_sum(c * solc[0]);
        const double bk = -tau * kap + sigma * mu - dta * dka;
        const double dtau = (-eta * rt - bk * tinv - cx2 - hz2) * rden;
        const double dkap = (bk - kap * dtau) * tinv;
        const double dx = solc[0] + dtau * x1;
        ratio = fmax(-dtau * tinv, -dkap * kinv);
        double ds[RPL];
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            dz[r] = fma(dtau, z1[r], dz[r]);
            ds[r] = (bs[r] - s[r] * dz[r]) * zinv[r];
            if (live[r]) ratio = fmax(ratio, fmax(-ds[r] * sinv[r], -dz[r] * zinv[r]));
        }
        ratio = warp_max(ratio);
        const double amax = ratio > 0.0 ? 1.0 / ratio : 1e30;
        const double alpha = fmin(1.0, LP_STEP * amax);
        x = fma(alpha, dx, x);
        tau = fma(alpha, dtau, tau);
        kap = fma(alpha, dkap, kap);
#pragma unroll
        for (int r = 0; r < RPL; ++r)
            if (live[r]) { s[r] = fma(alpha, ds[r], s[r]); z[r] = fma(alpha, dz[r], z[r]); }
    }
    return res;
}

}  // namespace pb200

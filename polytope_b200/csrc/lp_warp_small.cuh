// One-LP-per-warp solver specialised for n <= 8 columns (the reduce / bounding
// box LPs of d <= 8 polytopes, Chebyshev LPs of d <= 7, adjacency LPs of grids).
//
// Same algorithm as lp_warp.cuh (tests/ipm_model.py is the model of both), but
// every n-sized object is REPLICATED in all 32 lanes instead of being owned by
// lanes 0..n-1:
//   * the 8x8 normal matrix comes out of one DMMA tile, goes through shared
//     memory once, and every lane factors it in registers (no shuffles, no
//     shared-memory round trips inside the Cholesky dependency chain);
//   * the triangular solves are local register code reading L by broadcast;
//   * x, c, dx ... are register vectors, so G x products need no broadcast
//     loads and n-vector dot products need no warp reduction at all.
// Only the per-row quantities (h, s, z, ...) stay lane-owned, and the only
// warp reductions left per iteration are 3 + 2 + 1 sums and 2 maxima.
// The r01 ncu profile of the lane-owned version (profiles/) showed 45 % of the
// issue slots in Cholesky / triangular solves and 12 % in shuffle reductions.
#pragma once
#include "lp_warp.cuh"

namespace pb200 {

constexpr int NS = 8;                       // padded column count of this path
constexpr int NTRI = NS * (NS + 1) / 2;     // packed lower triangle

// per-warp scratch (doubles): G[NS][MP] | d[MP] | V[NSLOT][MP] | M[NS][NS] | R[NSLOT][NS] | vec[NVEC][NS]
constexpr int NVEC = 6;
enum : int { VC = 0, VX = 1, V1 = 2, V2 = 3, V3 = 4, VS = 5 };   // c, x, x1, x2(aff), x2(cor), spare
__host__ __device__ inline int lps_scratch_doubles(int rpl) {
    const int MP = lp_mp(rpl);
    return NS * MP + MP + NSLOT * MP + NS * NS + NSLOT * NS + NVEC * NS + MP;
}
struct SmallScratch {
    double *G, *d, *V, *M, *R, *X;
    double* hb;  // [MP] staged right-hand side, kept across LPs that share G
    int MP;
    int NC;      // columns allocated in G (what staging must zero)
};
__device__ inline SmallScratch lps_carve(double* base, int rpl) {
    SmallScratch w;
    w.MP = lp_mp(rpl);
    w.NC = NS;
    w.G = base;  base += NS * w.MP;
    w.d = base;  base += w.MP;
    w.V = base;  base += NSLOT * w.MP;
    w.M = base;  base += NS * NS;
    w.R = base;  base += NSLOT * NS;
    w.X = base;  base += NVEC * NS;
    w.hb = base;
    return w;
}

__device__ PB200_REDUCE_INLINE void warp_sum3(double& a, double& b, double& c) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double ta = __shfl_xor_sync(FULL_MASK, a, o);
        const double tb = __shfl_xor_sync(FULL_MASK, b, o);
        const double tc = __shfl_xor_sync(FULL_MASK, c, o);
        a += ta; b += tb; c += tc;
    }
}
// max over the warp of a non-negative step ratio, in fp32 rounded up: the step
// is 0.99 / ratio, so an over-estimate by 1 ulp(fp32) only shortens the step.
__device__ __forceinline__ double warp_max_ratio(double v) {
    float f = __double2float_ru(v);
#pragma unroll
    for (int o = 16; o; o >>= 1) f = fmaxf(f, __shfl_xor_sync(FULL_MASK, f, o));
    return (double)f;
}
// five maxima at once, in fp32 rounded up (used for threshold tests only)
__device__ PB200_REDUCE_INLINE void warp_max5f(float& a, float& b, float& c, float& d, float& e) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float ta = __shfl_xor_sync(FULL_MASK, a, o);
        const float tb = __shfl_xor_sync(FULL_MASK, b, o);
        const float tc = __shfl_xor_sync(FULL_MASK, c, o);
        const float td = __shfl_xor_sync(FULL_MASK, d, o);
        const float te = __shfl_xor_sync(FULL_MASK, e, o);
        a = fmaxf(a, ta); b = fmaxf(b, tb); c = fmaxf(c, tc); d = fmaxf(d, td); e = fmaxf(e, te);
    }
}
__device__ PB200_REDUCE_INLINE void warp_sum5(double& a, double& b, double& c, double& d, double& e) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double ta = __shfl_xor_sync(FULL_MASK, a, o);
        const double tb = __shfl_xor_sync(FULL_MASK, b, o);
        const double tc = __shfl_xor_sync(FULL_MASK, c, o);
        const double td = __shfl_xor_sync(FULL_MASK, d, o);
        const double te = __shfl_xor_sync(FULL_MASK, e, o);
        a += ta; b += tb; c += tc; d += td; e += te;
    }
}
__device__ PB200_REDUCE_INLINE void warp_sum2(double& a, double& b) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double ta = __shfl_xor_sync(FULL_MASK, a, o);
        const double tb = __shfl_xor_sync(FULL_MASK, b, o);
        a += ta; b += tb;
    }
}

// ---- one DMMA tile: R[slot][j] = sum_i G[i][j] V[slot][i] ----
// The k-loops of both DMMA passes cover whole 32-row blocks of the staged rows (rows >= m
// are zero in G, V and d), so they unroll completely: no loop branches in the hot path.
template <int RPL>
__device__ __forceinline__ void s_gt_times_slots(const SmallScratch& w, int mk, int nslots, int lane) {
    const int q = lane >> 2, t = lane & 3;
    double c0 = 0.0, c1 = 0.0;
    const bool cb = q < nslots;
    const double* ga = w.G + q * w.MP + t;
    const double* vb = w.V + (cb ? q : 0) * w.MP + t;
#pragma unroll
    for (int blk = 0; blk < RPL; ++blk)
        if (blk == 0 || mk > 32 * blk) {          // whole 32-row blocks: one uniform test per block
#pragma unroll
            for (int i0 = 32 * blk; i0 < 32 * blk + 32; i0 += 4) {
                const double a = ga[i0];
                const double b = cb ? vb[i0] : 0.0;
                dmma884(c0, c1, a, b);
            }
        }
    if (t < 2) {
        w.R[(2 * t) * NS + q] = c0;
        w.R[(2 * t + 1) * NS + q] = c1;
    }
    __syncwarp();
}

// ---- one DMMA tile: M = G' diag(d) G (full 8x8, row-major) ----
template <int RPL>
__device__ __forceinline__ void s_normal_matrix(const SmallScratch& w, int mk, int lane) {
    const int q = lane >> 2, t = lane & 3;
    double c0 = 0.0, c1 = 0.0;
    const double* gq = w.G + q * w.MP + t;
    const double* dd = w.d + t;
#pragma unroll
    for (int blk = 0; blk < RPL; ++blk)
        if (blk == 0 || mk > 32 * blk) {
#pragma unroll
            for (int i0 = 32 * blk; i0 < 32 * blk + 32; i0 += 4) {
                const double g = gq[i0];
                dmma884(c0, c1, g * dd[i0], g);
            }
        }
    *reinterpret_cast<double2*>(w.M + q * NS + 2 * t) = make_double2(c0, c1);
    __syncwarp();
}

// ---- replicated Cholesky (inlined at ONE site): every lane factors the 8x8
// matrix in Ms in registers; the factor goes back to Ms as a full symmetric
// array (M[i][j] = M[j][i] = L_ij for i > j, M[k][k] = 1/L_kk, or 0 for a
// skipped pivot) so forward and backward substitution both read rows.
// Vanishing pivots are skipped LIPSOL-style (solution component forced to 0).
__device__ __forceinline__ unsigned s_cholesky(double* Ms, int n, double add_diag, int lane) {
    double L[NTRI];
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
        for (int j = 0; j <= i; j += 2) {
            const double2 v = *reinterpret_cast<const double2*>(Ms + i * NS + j);
            L[i * (i + 1) / 2 + j] = v.x;
            if (j + 1 <= i) L[i * (i + 1) / 2 + j + 1] = v.y;
        }
    unsigned skipped = 0;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        // original diagonal entry is still in shared memory (written back last)
        const double dgk = Ms[k * NS + k] + add_diag;
        const double p = L[k * (k + 1) / 2 + k] + add_diag;
        const bool ok = (k < n) && (p > 1e-13 * dgk) && (p > 1e-290);
        const double rinv = ok ? fast_rsqrt(p) : 0.0;
        skipped |= ok ? 0u : (1u << k);
        L[k * (k + 1) / 2 + k] = rinv;
#pragma unroll
        for (int i = k + 1; i < NS; ++i) L[i * (i + 1) / 2 + k] *= rinv;
#pragma unroll
        for (int i = k + 1; i < NS; ++i)
#pragma unroll
            for (int j = k + 1; j <= i; ++j)
                L[i * (i + 1) / 2 + j] = fma(-L[i * (i + 1) / 2 + k], L[j * (j + 1) / 2 + k], L[i * (i + 1) / 2 + j]);
    }
    __syncwarp();
    // every lane holds identical values: all of them store the whole symmetric array
    // (same address, same data -- one wavefront per store, no divergent branches)
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
        for (int j = 0; j < NS; j += 2) {
            const double a = j <= i ? L[i * (i + 1) / 2 + j] : L[j * (j + 1) / 2 + i];
            const double b = j + 1 <= i ? L[i * (i + 1) / 2 + j + 1] : L[(j + 1) * (j + 2) / 2 + i];
            *reinterpret_cast<double2*>(Ms + i * NS + j) = make_double2(a, b);
        }
    __syncwarp();
    return skipped & ((1u << n) - 1u);
}

__device__ __forceinline__ void load_vec(const double* src, double (&v)[NS]) {
#pragma unroll
    for (int j = 0; j < NS; j += 2) {
        const double2 t = *reinterpret_cast<const double2*>(src + j);
        v[j] = t.x; v[j + 1] = t.y;
    }
}

// ---- replicated solve of (L L') y = r for the NR (1 or 2) right-hand sides stored
// as consecutive 8-vectors at X; solutions overwrite them.  L is read by broadcast.
// With two right-hand sides each half warp solves one of them, so the instruction
// count is that of a single solve (a caller with one right-hand side passes zeros
// as the second: a run-time count cost one register move per DFMA, profiles/r01e_*).
template <int NR>
__device__ __forceinline__ void s_solve(const double* Ms, double* X, int lane) {
    double* mine = X + (NR == 2 ? (lane >> 4) * NS : 0);
    double a[NS];
    load_vec(mine, a);
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        double row[NS];
        load_vec(Ms + k * NS, row);
        a[k] *= row[k];
#pragma unroll
        for (int i = k + 1; i < NS; ++i) a[i] = fma(-row[i], a[k], a[i]);
    }
#pragma unroll
    for (int k = NS - 1; k >= 0; --k) {
        double row[NS];
        load_vec(Ms + k * NS, row);
        a[k] *= row[k];
#pragma unroll
        for (int i = 0; i < k; ++i) a[i] = fma(-row[i], a[k], a[i]);
    }
    __syncwarp();
    // identical values in every lane of a half warp: all lanes store (same address, same data)
#pragma unroll
    for (int j = 0; j < NS; j += 2) *reinterpret_cast<double2*>(mine + j) = make_double2(a[j], a[j + 1]);
    __syncwarp();
}

template <int RPL>
__device__ __forceinline__ void s_rows_times(const SmallScratch& w, int lane, const double (&u)[NS], double (&out)[RPL]) {
#pragma unroll
    for (int r = 0; r < RPL; ++r) out[r] = 0.0;
#pragma unroll
    for (int j = 0; j < NS; ++j)
#pragma unroll
        for (int r = 0; r < RPL; ++r) out[r] = fma(w.G[j * w.MP + lane + 32 * r], u[j], out[r]);
}
template <int RPL>
__device__ __forceinline__ void s_rows_times2(const SmallScratch& w, int lane, const double (&u)[NS], const double (&v)[NS],
                                              double (&ou)[RPL], double (&ov)[RPL]) {
#pragma unroll
    for (int r = 0; r < RPL; ++r) { ou[r] = 0.0; ov[r] = 0.0; }
#pragma unroll
    for (int j = 0; j < NS; ++j)
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const double g = w.G[j * w.MP + lane + 32 * r];
            ou[r] = fma(g, u[j], ou[r]);
            ov[r] = fma(g, v[j], ov[r]);
        }
}
__device__ __forceinline__ double dot8(const double (&a)[NS], const double (&b)[NS]) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int j = 0; j < NS; j += 2) { s0 = fma(a[j], b[j], s0); s1 = fma(a[j + 1], b[j + 1], s1); }
    return s0 + s1;
}

// Cold path (rank-deficient G at iteration 0): does c have a component in
// null(G)?  M already holds the skip-factor; d is the row weight vector in w.d.
template <int RPL>
__device__ __noinline__ bool s_objective_leaves_range(const SmallScratch& w, int mk, double cl, int lane) {
    const bool own = lane < NS;
    double* X3 = w.X + 4 * NS;
    if (own) X3[lane] = cl;
    __syncwarp();
    s_solve<1>(w.M, X3, lane);
    double uu[NS], gu[RPL];
    load_vec(X3, uu);
    s_rows_times<RPL>(w, lane, uu, gu);
#pragma unroll
    for (int r = 0; r < RPL; ++r) w.V[lane + 32 * r] = w.d[lane + 32 * r] * gu[r];
    __syncwarp();
    s_gt_times_slots<RPL>(w, mk, 1, lane);
    const double back = own ? w.R[lane] : 0.0;
    const double rmax = warp_max(fabs(cl - back)), cmax = warp_max(fabs(cl));
    __syncwarp();
    return rmax > 1e-9 * fmax(cmax, 1e-300);
}

struct SmallResult {
    int status, iters;
    double fun;
    double x;           // lane-owned component (lanes < NS), valid if status == 0
};

// G staged in w.G as NS zero-padded columns (rows >= m zero); cl = lane-owned
// objective component (0 for lanes >= n); h[r] lane-owned right-hand side of row
// lane+32r.  n-vectors live in the shared-memory vector file w.X[NVEC][NS]:
// elementwise updates are done by lanes 0..7 on their own component, and every
// lane reads whole vectors by broadcast when it needs them replicated.
//
// The interior-point iterations and the polish rounds are steps of ONE loop
// (`phase`), so that the big unrolled blocks -- Cholesky, the two-rhs solve,
// the DMMA passes -- exist once in the instruction stream: the first version
// of this kernel was 180 KB of SASS and lost ~30 % of its cycles to
// instruction-cache misses (profiles/r01a_*), and calling them as functions
// cost ~900 local-memory spill instructions per LP (profiles/r01c_*).
template <int RPL>
__device__ SmallResult lp_solve_small(const SmallScratch& w, int m, int n, double cl_in, const double (&h_in)[RPL]) {
    const int lane = threadIdx.x & 31;
    const int MP = w.MP;
    const int mk = (m + 3) & ~3;
    const bool own = lane < NS;
    const double cl0 = own ? cl_in : 0.0;
    double cl = cl0;
    double* const Xc = w.X + VC * NS;
    double* const Xx = w.X + VX * NS;
    double* const X1 = w.X + V1 * NS;      // X1, X2 contiguous: one s_solve<2>
    double* const X2 = w.X + V2 * NS;
    if (own) { Xc[lane] = cl; Xx[lane] = 0.0; }

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
    double hh = 0.0, hmax = 0.0, cc2 = cl * cl;
#pragma unroll
    for (int r = 0; r < RPL; ++r) { hh = fma(h[r], h[r], hh); hmax = fmax(hmax, fabs(h[r])); }
    warp_sum2(hh, cc2);
    hmax = warp_max(hmax);
    const double nh2 = fmax(1.0, hh);                 // max(1, ||h||)^2
    double nc2 = fmax(1.0, cc2);
    const double rmu = 1.0 / (double)(mlive + 1);
    double xl = 0.0, tau = 1.0, kap = 1.0;
    bool lineal = false;
    SmallResult res;
    res.status = ST_ITER_LIMIT; res.iters = 0; res.fun = 0.0; res.x = 0.0;
    int phase = 0, round = 0, it = 0;     // phase 0: interior point, 1: polish (+ certificate)
    bool early = false;                   // early: polish attempted at a loose tolerance
    double etol = LP_EARLY_TOL;           // tolerance of the next certified-polish attempt
    bool act[RPL];
    double y[RPL];                        // dual certificate on the active rows
#pragma unroll
    for (int r = 0; r < RPL; ++r) { act[r] = false; y[r] = 0.0; }
    double f0 = 0.0, xp = 0.0;            // xp: lane-owned polished point
    __syncwarp();

#pragma unroll 1
    for (int step = 0; step < LP_MAX_ITER + 16; ++step) {
        // ---- A. rows of G times the current point (x, or x/tau while polishing) ----
        double gx[RPL], gu[RPL], cx;
        {
            double x[NS], c[NS], u[NS];
            load_vec(Xx, x);
            load_vec(X2, u);                 // polish: u of the dual refinement (else unused)
            s_rows_times2<RPL>(w, lane, x, u, gx, gu);
            load_vec(Xc, c);
            cx = dot8(c, x);
        }
        double rz[RPL], d[RPL], sinv[RPL], zinv[RPL];
        double sz = 0.0, hz = 0.0, rz2 = 0.0, gxs2 = 0.0;
        if (phase == 0) {
            res.iters = it;
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int i = lane + 32 * r;
                sinv[r] = fast_rcp(s[r]);
                zinv[r] = live[r] ? fast_rcp(z[r]) : 0.0;
                d[r] = z[r] * sinv[r];
                rz[r] = live[r] ? gx[r] + s[r] - h[r] * tau : 0.0;
                const double gxs = live[r] ? gx[r] + s[r] : 0.0;
                sz = fma(s[r], z[r], sz);
                hz = fma(h[r], z[r], hz);
                rz2 = fma(rz[r], rz[r], rz2);
                gxs2 = fma(gxs, gxs, gxs2);
                w.d[i] = d[r];
                w.V[0 * MP + i] = z[r];
                w.V[1 * MP + i] = d[r] * h[r];
                w.V[2 * MP + i] = z[r] - d[r] * rz[r];
            }
        } else {
            // ---- joint polish step: one projection round of the primal point onto the
            // active face and one least-squares refinement of the multipliers share the
            // DMMA pass, the (half-warp split) solve and the G-multiplication ----
#pragma unroll
            for (int r = 0; r < RPL; ++r) { rz[r] = 0.0; d[r] = 0.0; sinv[r] = 0.0; zinv[r] = 0.0; }
            float ft = 0.f, fslack = -3.0e38f, fymin = -3.0e38f, fymax = 0.f;
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const double res_r = h[r] - gx[r];
                if (act[r]) {
                    y[r] -= gu[r];           // y -= G_B u,  u = (G_B'G_B)^-1 (G_B'y + c)  (u = 0 in round 0)
                    ft = fmaxf(ft, __double2float_ru(fabs(res_r)));
                    fymin = fmaxf(fymin, __double2float_ru(-y[r]));
                    fymax = fmaxf(fymax, __double2float_ru(y[r]));
                }
                if (live[r]) fslack = fmaxf(fslack, __double2float_ru(-res_r));
                w.V[0 * MP + lane + 32 * r] = act[r] ? res_r : 0.0;
                w.V[1 * MP + lane + 32 * r] = y[r];
            }
            __syncwarp();
            s_gt_times_slots<RPL>(w, mk, 2, lane);
            float frd = own ? __double2float_ru(fabs(w.R[NS + lane] + cl0)) : 0.f;      // |G_B'y + c|
            warp_max5f(ft, fslack, fymin, fymax, frd);
            const double scale = fmax(1.0, hmax);
            const bool feasible = (double)fslack <= 1e-9 * scale;
            const bool settled = (double)ft <= 1e-13 * scale || round == 3;    // projection converged (or out of rounds)
            if (PB_UNI(!early)) {
                // final polish of a tightly converged iterate: objective must agree
                if (PB_UNI(settled)) {
                    const double f1 = warp_sum(cl0 * xp);
                    if (PB_UNI(feasible && (fabs(f1 - f0) <= 1e-6 * fmax(1.0, fabs(f0))))) { res.x = xp; res.fun = f1; }
                    break;
                }
            } else {
                const bool dual_ok = (double)frd <= 1e-9 * sqrt(nc2) && (double)fymin <= 1e-9 * fmax(1.0, (double)fymax);
                if (PB_UNI(settled && feasible && (double)ft <= 1e-9 * scale && dual_ok)) {
                    // primal feasible, active rows tight, dual feasible, complementary: optimal
                    res.status = ST_OPTIMAL;
                    res.x = xp;
                    res.fun = warp_sum(cl0 * xp);
                    break;
                }
                if (PB_UNI(round == 3)) {
                    // not certified: resume the interior-point iterations from the untouched iterate
                    phase = 0; early = false;
                    if (own) { Xx[lane] = xl; X2[lane] = 0.0; }
                    __syncwarp();
                    continue;
                }
            }
        }
        if (phase == 0) {
            __syncwarp();
            s_gt_times_slots<RPL>(w, mk, 3, lane);
        }
        const bool refactor = (phase == 0) || (round == 0);
        if (refactor) s_normal_matrix<RPL>(w, mk, lane);
        double rxl = 0.0, rt = 0.0, mu = 0.0, tinv = 1.0;
        if (phase == 0) {
            const double gzl = own ? w.R[lane] : 0.0;          // (G'z)_lane
            rxl = fma(cl, tau, gzl);
            double rx2 = rxl * rxl;
            warp_sum5(sz, hz, rz2, rx2, gxs2);
            rt = cx + hz + kap;
            mu = (sz + tau * kap) * rmu;
            tinv = fast_rcp(tau);
            // ---- termination (cvxopt conelp-style tests, squared where a norm is involved) ----
            const double t2 = tinv * tinv;
            const double pcost = cx * tinv, dcost = -hz * tinv;
            const double gap = sz * t2;
            // relgap <= tol  <=>  gap <= tol * (-pcost)  or  gap <= tol * dcost
            const double gapref = pcost < 0.0 ? -pcost : (dcost > 0.0 ? dcost : 0.0);
            if (PB_UNI(!(mu == mu) || !(fabs(tau) < 1e300) || !(fabs(cx) < 1e300))) { res.status = ST_NUMERICAL; break; }
            // (second alternative: see LP_STALL_* in lp_warp.cuh)
            const bool converged = rz2 * t2 <= LP_FEAS_TOL * LP_FEAS_TOL * nh2 &&
                                   ((rx2 * t2 <= LP_FEAS_TOL * LP_FEAS_TOL * nc2 && (gap <= LP_GAP_TOL || gap <= LP_GAP_TOL * gapref)) ||
                                    (rx2 * t2 <= LP_STALL_DRES * LP_STALL_DRES * nc2 && (gap <= LP_STALL_GAP || gap <= LP_STALL_GAP * gapref)));
            // At the loose tolerance the active set is usually already identified:
            // try the polish there and accept it only with a full optimality
            // certificate (primal feasible, active rows tight, y >= 0 with
            // G_B'y + c = 0); otherwise keep iterating to the tight tolerance.
            const bool loosely = etol > 1e-7 && !lineal && rz2 * t2 <= etol * etol * nh2 && rx2 * t2 <= etol * etol * nc2 &&
                                 (gap <= etol || gap <= etol * gapref);
            if (PB_UNI(converged || loosely)) {
                if (PB_UNI(converged && lineal)) { res.status = ST_UNBOUNDED; break; }
                early = !converged;
                etol *= LP_EARLY_NEXT;
                // ---- extract, then polish on the active set (see lp_warp.cuh) ----
                const double te = 1.0 / tau;
                xp = xl * te;
                int nact = 0;
#pragma unroll
                for (int r = 0; r < RPL; ++r) {
                    act[r] = live[r] && (z[r] > s[r]);
                    nact += act[r] ? 1 : 0;
                }
                nact = __reduce_add_sync(FULL_MASK, nact);
                if (PB_UNI(converged)) {
                    f0 = warp_sum(cl0 * xp);
                    res.status = ST_OPTIMAL; res.x = xp; res.fun = f0;
                    if (nact == 0) break;
                }
                if (nact > 0) {
#pragma unroll
                    for (int r = 0; r < RPL; ++r) {
                        w.d[lane + 32 * r] = act[r] ? 1.0 : 0.0;
                        y[r] = act[r] ? z[r] * te : 0.0;      // multipliers of the active rows
                    }
                    if (own) { Xx[lane] = xp; X2[lane] = 0.0; }
                    __syncwarp();
                    phase = 1; round = 0;
                    continue;
                }
                early = false;      // nothing active yet: keep iterating
            }
            if (PB_UNI(tau < 1e-3 * kap)) {
                if (PB_UNI(hz < 0.0)) {
                    const double gz2 = warp_sum(gzl * gzl);
                    if (PB_UNI(sqrt(gz2 * nh2 / nc2) <= 10.0 * LP_FEAS_TOL * (-hz))) { res.status = ST_INFEASIBLE; break; }
                }
                if (PB_UNI(cx < 0.0 && sqrt(gxs2 * nc2 / nh2) <= 10.0 * LP_FEAS_TOL * (-cx))) {
                    // unbounded if feasible at all: settle feasibility first (HiGHS reports 2 for an LP
                    // that is infeasible as well) by restarting on the feasibility problem, c = 0
                    lineal = true;
                    cl = 0.0;
                    nc2 = 1.0;
                    xl = 0.0; tau = 1.0; kap = 1.0;
                    if (own) { Xc[lane] = 0.0; Xx[lane] = 0.0; }
#pragma unroll
                    for (int r = 0; r < RPL; ++r) {
                        s[r] = live[r] ? fmax(h[r], 0.0) + 1.0 : 1.0;
                        z[r] = live[r] ? 1.0 : 0.0;
                    }
                    if (it == 0) it = 1;
                    __syncwarp();
                    continue;
                }
            }
            if (it == LP_MAX_ITER) break;
        }
        // ---- factor (interior point: every step; polish: once) ----
        if (refactor) {
            double add_diag = 0.0;
            if (phase == 1) {
                double dmax = 1.0;
#pragma unroll
                for (int j = 0; j < NS; ++j) dmax = fmax(dmax, w.M[j * NS + j]);
                add_diag = 1e-9 * dmax;
                __syncwarp();
            }
            const unsigned skipped = s_cholesky(w.M, n, add_diag, lane);
            if (PB_UNI(phase == 0 && it == 0 && skipped && !lineal)) {
                // rank-deficient G: if c has a component in null(G) the LP is unbounded
                // whenever it is feasible -> continue with c = 0 and report 3 instead of 0
                if (PB_UNI(s_objective_leaves_range<RPL>(w, mk, cl, lane))) {
                    lineal = true;
                    cl = 0.0;
                    if (own) Xc[lane] = 0.0;
                    nc2 = 1.0;
                    __syncwarp();
                    it = 1;
                    continue;
                }
            }
        }
        // ---- solves: interior point = predictor pair then corrector; polish = one ----
        double z1[RPL], dza[RPL], dsa[RPL], bs[RPL], qc[RPL];
        double x1l = 0.0, rden = 0.0, dta = 0.0, dka = 0.0, kinv = 0.0, sigma = 0.0, eta = 1.0;
        const int npass = phase == 0 ? 2 : 1;
#pragma unroll 1
        for (int pass = 0; pass < npass; ++pass) {
            if (own) {
                double ra, rb = 0.0;
                if (phase == 1) { ra = w.R[lane]; rb = w.R[NS + lane] + cl0; }      // projection | refinement
                else if (pass == 0) { ra = w.R[NS + lane] - cl; rb = w.R[2 * NS + lane] - rxl; }
                else ra = fma(-eta, rxl, w.R[lane]);
                X1[lane] = ra;
                X2[lane] = rb;
            }
            __syncwarp();
            s_solve<2>(w.M, X1, lane);
            if (phase == 1) {
                if (own) { xp += X1[lane]; Xx[lane] = xp; }      // X2 keeps u for the next step's G u
                ++round;
                __syncwarp();
                break;
            }
            double g1[RPL], g2[RPL], cx1, cx2;
            {
                double x1[NS], x2[NS], c[NS];
                load_vec(X1, x1);
                load_vec(X2, x2);
                s_rows_times2<RPL>(w, lane, x1, x2, g1, g2);
                load_vec(Xc, c);
                cx1 = dot8(c, x1);
                cx2 = dot8(c, x2);
            }
            if (pass == 0) {
                // K [x1; z1] = [-c; h],  K [x2; z2] = [-rx; q_aff]  -> affine direction
                if (own) x1l = X1[lane];
                double hz1 = 0.0, hz2 = 0.0;
#pragma unroll
                for (int r = 0; r < RPL; ++r) {
                    z1[r] = d[r] * (g1[r] - h[r]);
                    dza[r] = d[r] * (g2[r] - (s[r] - rz[r]));
                    hz1 = fma(h[r], z1[r], hz1);
                    hz2 = fma(h[r], dza[r], hz2);
                }
                warp_sum2(hz1, hz2);
                const double kot = kap * tinv;
                const double den = cx1 + hz1 - kot;                     // < 0
                rden = fast_rcp(den);
                dta = (-rt + kap - cx2 - hz2) * rden;
                dka = -kap - kot * dta;
                kinv = fast_rcp(kap);
                double ratio = fmax(-dta * tinv, -dka * kinv);
#pragma unroll
                for (int r = 0; r < RPL; ++r) {
                    dza[r] = fma(dta, z1[r], dza[r]);
                    dsa[r] = -s[r] - s[r] * zinv[r] * dza[r];
                    if (live[r]) ratio = fmax(ratio, fmax(-dsa[r] * sinv[r], -dza[r] * zinv[r]));
                }
                ratio = warp_max_ratio(fmax(ratio, 0.0));
                const double alpha_aff = ratio > 1.0 ? fast_rcp(ratio) : 1.0;
                const double om = 1.0 - alpha_aff;
                sigma = om * om * om;
                eta = 1.0 - sigma;
                // corrector right-hand side
#pragma unroll
                for (int r = 0; r < RPL; ++r) {
                    bs[r] = -s[r] * z[r] + sigma * mu - dsa[r] * dza[r];
                    qc[r] = live[r] ? -eta * rz[r] - bs[r] * zinv[r] : 0.0;
                    w.V[lane + 32 * r] = d[r] * qc[r];
                }
                __syncwarp();
                s_gt_times_slots<RPL>(w, mk, 1, lane);
            } else {
                // combined direction and step
                double hz2 = 0.0;
                double dz[RPL], ds[RPL];
#pragma unroll
                for (int r = 0; r < RPL; ++r) {
                    dz[r] = d[r] * (g1[r] - qc[r]);
                    hz2 = fma(h[r], dz[r], hz2);
                }
                hz2 = warp_sum(hz2);
                const double bk = -tau * kap + sigma * mu - dta * dka;
                const double dtau = (-eta * rt - bk * tinv - cx1 - hz2) * rden;
                const double dkap = (bk - kap * dtau) * tinv;
                double ratio = fmax(-dtau * tinv, -dkap * kinv);
#pragma unroll
                for (int r = 0; r < RPL; ++r) {
                    dz[r] = fma(dtau, z1[r], dz[r]);
                    ds[r] = (bs[r] - s[r] * dz[r]) * zinv[r];
                    if (live[r]) ratio = fmax(ratio, fmax(-ds[r] * sinv[r], -dz[r] * zinv[r]));
                }
                ratio = warp_max_ratio(fmax(ratio, 0.0));
                const double amax = ratio > 0.0 ? fast_rcp(ratio) : 1e30;
                const double alpha = fmin(1.0, LP_STEP * amax);
                if (own) {
                    xl = fma(alpha, fma(dtau, x1l, X1[lane]), xl);
                    Xx[lane] = xl;
                }
                tau = fma(alpha, dtau, tau);
                kap = fma(alpha, dkap, kap);
#pragma unroll
                for (int r = 0; r < RPL; ++r)
                    if (live[r]) { s[r] = fma(alpha, ds[r], s[r]); z[r] = fma(alpha, dz[r], z[r]); }
                ++it;
                __syncwarp();
            }
        }
    }
    return res;
}

}  // namespace pb200

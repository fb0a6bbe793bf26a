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
    return NS * MP + MP + NSLOT * MP + NS * NS + NSLOT * NS + NVEC * NS;
}
struct SmallScratch {
    double *G, *d, *V, *M, *R, *X;
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
    w.X = base;
    return w;
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

__device__ __forceinline__ void warp_sum3(double& a, double& b, double& c) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double ta = __shfl_xor_sync(FULL_MASK, a, o);
        const double tb = __shfl_xor_sync(FULL_MASK, b, o);
        const double tc = __shfl_xor_sync(FULL_MASK, c, o);
        a += ta; b += tb; c += tc;
    }
}
__device__ __forceinline__ void warp_sum2(double& a, double& b) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double ta = __shfl_xor_sync(FULL_MASK, a, o);
        const double tb = __shfl_xor_sync(FULL_MASK, b, o);
        a += ta; b += tb;
    }
}

// ---- one DMMA tile: R[slot][j] = sum_i G[i][j] V[slot][i] ----
__device__ __forceinline__ void s_gt_times_slots(const SmallScratch& w, int mk, int nslots, int lane) {
    const int q = lane >> 2, t = lane & 3;
    double c0 = 0.0, c1 = 0.0;
    const bool cb = q < nslots;
    const double* ga = w.G + q * w.MP + t;
    const double* vb = w.V + q * w.MP + t;
    for (int i0 = 0; i0 < mk; i0 += 4) {
        const double a = ga[i0];
        const double b = cb ? vb[i0] : 0.0;
        dmma884(c0, c1, a, b);
    }
    if (t < 2) {
        w.R[(2 * t) * NS + q] = c0;
        w.R[(2 * t + 1) * NS + q] = c1;
    }
    __syncwarp();
}

// ---- one DMMA tile: M = G' diag(d) G (full 8x8, row-major) ----
__device__ __forceinline__ void s_normal_matrix(const SmallScratch& w, int mk, int lane) {
    const int q = lane >> 2, t = lane & 3;
    double c0 = 0.0, c1 = 0.0;
    const double* gq = w.G + q * w.MP + t;
    const double* dd = w.d + t;
    for (int i0 = 0; i0 < mk; i0 += 4) {
        const double g = gq[i0];
        dmma884(c0, c1, g * dd[i0], g);
    }
    *reinterpret_cast<double2*>(w.M + q * NS + 2 * t) = make_double2(c0, c1);
    __syncwarp();
}

// ---- replicated Cholesky: every lane factors the 8x8 matrix in Ms in
// registers; the factor goes back to Ms as a full symmetric array
// (M[i][j] = M[j][i] = L_ij for i > j, M[k][k] = 1/L_kk, or 0 for a skipped
// pivot) so that forward and backward substitution both read rows.
// `add_diag` is added to every diagonal entry first (polish regularisation).
__device__ __noinline__ unsigned s_cholesky(double* Ms, int n, double add_diag, int lane) {
    double L[NTRI];
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
        for (int j = 0; j <= i; j += 2) {
            const double2 v = *reinterpret_cast<const double2*>(Ms + i * NS + j);
            L[i * (i + 1) / 2 + j] = v.x;
            if (j + 1 <= i) L[i * (i + 1) / 2 + j + 1] = v.y;
        }
    double dmax = 1e-300;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        L[k * (k + 1) / 2 + k] += add_diag;
        dmax = fmax(dmax, L[k * (k + 1) / 2 + k]);
    }
    const double floor_abs = 1e-30 * dmax;
    unsigned skipped = 0;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const double p = L[k * (k + 1) / 2 + k];
        // original diagonal entry is still in shared memory (written back last)
        const double thr = fmax(1e-13 * (Ms[k * NS + k] + add_diag), floor_abs);
        const bool ok = (k < n) && (p > thr);
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
    // lane i writes row i of the symmetric array (all lanes hold identical values)
#pragma unroll
    for (int i = 0; i < NS; ++i)
        if (lane == i) {
#pragma unroll
            for (int j = 0; j < NS; j += 2) {
                const double a = j <= i ? L[i * (i + 1) / 2 + j] : L[j * (j + 1) / 2 + i];
                const double b = j + 1 <= i ? L[i * (i + 1) / 2 + j + 1] : L[(j + 1) * (j + 2) / 2 + i];
                *reinterpret_cast<double2*>(Ms + i * NS + j) = make_double2(a, b);
            }
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

// ---- replicated solve of (L L') y = r, NR right-hand sides stored as
// consecutive 8-vectors at X; solutions overwrite them.  L is read by broadcast.
template <int NR>
__device__ __noinline__ void s_solve(const double* Ms, double* X, int lane) {
    double a[NR][NS];
#pragma unroll
    for (int v = 0; v < NR; ++v) load_vec(X + v * NS, a[v]);
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        double row[NS];
        load_vec(Ms + k * NS, row);
#pragma unroll
        for (int v = 0; v < NR; ++v) {
            a[v][k] *= row[k];
#pragma unroll
            for (int i = k + 1; i < NS; ++i) a[v][i] = fma(-row[i], a[v][k], a[v][i]);
        }
    }
#pragma unroll
    for (int k = NS - 1; k >= 0; --k) {
        double row[NS];
        load_vec(Ms + k * NS, row);
#pragma unroll
        for (int v = 0; v < NR; ++v) {
            a[v][k] *= row[k];
#pragma unroll
            for (int i = 0; i < k; ++i) a[v][i] = fma(-row[i], a[v][k], a[v][i]);
        }
    }
    __syncwarp();
    // lanes 0..3 (4..7) write the four 16-byte pieces of solution 0 (1)
#pragma unroll
    for (int v = 0; v < NR; ++v)
#pragma unroll
        for (int j = 0; j < NS; j += 2)
            if (lane == v * 4 + (j >> 1)) *reinterpret_cast<double2*>(X + v * NS + j) = make_double2(a[v][j], a[v][j + 1]);
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
template <int RPL>
__device__ SmallResult lp_solve_small(const SmallScratch& w, int m, int n, double cl_in, const double (&h_in)[RPL]) {
    const int lane = threadIdx.x & 31;
    const int MP = w.MP;
    const int mk = (m + 3) & ~3;
    const bool own = lane < NS;
    double cl = own ? cl_in : 0.0;
    double* const Xc = w.X + VC * NS;
    double* const Xx = w.X + VX * NS;
    double* const X1 = w.X + V1 * NS;      // X1, X2 contiguous: one s_solve<2>
    double* const X2 = w.X + V2 * NS;
    double* const X3 = w.X + V3 * NS;
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
    __syncwarp();

#pragma unroll 1
    for (int it = 0; it <= LP_MAX_ITER; ++it) {
        res.iters = it;
        // ---- residuals ----
        double rz[RPL], d[RPL], sinv[RPL], zinv[RPL];
        double sz = 0.0, hz = 0.0, rz2 = 0.0, gxs2 = 0.0, cx;
        {
            double x[NS], gx[RPL];
            load_vec(Xx, x);
            s_rows_times<RPL>(w, lane, x, gx);
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
            double c[NS];
            load_vec(Xc, c);
            cx = dot8(c, x);
        }
        __syncwarp();
        s_gt_times_slots(w, mk, 3, lane);
        s_normal_matrix(w, mk, lane);
        warp_sum3(sz, hz, rz2);
        const double gzl = own ? w.R[lane] : 0.0;          // (G'z)_lane
        const double rxl = fma(cl, tau, gzl);
        double rx2 = rxl * rxl, gz2 = gzl * gzl;
        warp_sum2(rx2, gz2);
        const double rt = cx + hz + kap;
        const double mu = (sz + tau * kap) * rmu;
        const double tinv = fast_rcp(tau);
        // ---- termination (cvxopt conelp-style tests, squared where a norm is involved) ----
        const double t2 = tinv * tinv;
        const double pres2 = rz2 * t2, dres2 = rx2 * t2;       // compare with tol^2 * nh2 / nc2
        const double pcost = cx * tinv, dcost = -hz * tinv;
        const double gap = sz * t2;
        double relgap = 1e300;
        if (pcost < 0.0) relgap = gap / -pcost;
        else if (dcost > 0.0) relgap = gap / dcost;
        if (!(mu == mu) || !(fabs(tau) < 1e300) || !(fabs(cx) < 1e300)) { res.status = ST_NUMERICAL; break; }
        if (pres2 <= LP_FEAS_TOL * LP_FEAS_TOL * nh2 && dres2 <= LP_FEAS_TOL * LP_FEAS_TOL * nc2 &&
            (gap <= LP_GAP_TOL || relgap <= LP_GAP_TOL)) {
            res.status = ST_OPTIMAL; break;
        }
        if (tau < 1e-3 * kap) {
            if (hz < 0.0 && sqrt(gz2 * nh2 / nc2) <= 10.0 * LP_FEAS_TOL * (-hz)) { res.status = ST_INFEASIBLE; break; }
            if (cx < 0.0) {
                gxs2 = warp_sum(gxs2);
                if (sqrt(gxs2 * nc2 / nh2) <= 10.0 * LP_FEAS_TOL * (-cx)) { res.status = ST_UNBOUNDED; break; }
            }
        }
        if (it == LP_MAX_ITER) break;
        // ---- factor ----
        const unsigned skipped = s_cholesky(w.M, n, 0.0, lane);
        if (it == 0 && skipped) {
            // rank-deficient G: does c have a component in null(G)?  (see lp_warp.cuh)
            if (own) X3[lane] = cl;
            __syncwarp();
            s_solve<1>(w.M, X3, lane);
            double uu[NS], gu[RPL];
            load_vec(X3, uu);
            s_rows_times<RPL>(w, lane, uu, gu);
#pragma unroll
            for (int r = 0; r < RPL; ++r) w.V[lane + 32 * r] = d[r] * gu[r];
            __syncwarp();
            s_gt_times_slots(w, mk, 1, lane);
            const double back = own ? w.R[lane] : 0.0;
            const double rmax = warp_max(fabs(cl - back)), cmax = warp_max(fabs(cl));
            if (rmax > 1e-9 * fmax(cmax, 1e-300)) {
                lineal = true;
                cl = 0.0;
                if (own) Xc[lane] = 0.0;
                nc2 = 1.0;
                __syncwarp();
                continue;
            }
        }
        // ---- K [x1; z1] = [-c; h],  K [x2; z2] = [-rx; q_aff] ----
        if (own) {
            X1[lane] = w.R[NS + lane] - cl;
            X2[lane] = w.R[2 * NS + lane] - rxl;
        }
        __syncwarp();
        s_solve<2>(w.M, X1, lane);
        double z1[RPL], dza[RPL], dsa[RPL];
        double hz1 = 0.0, hz2 = 0.0, cx1, cx2;
        {
            double x1[NS], x2[NS], g1[RPL], g2[RPL], c[NS];
            load_vec(X1, x1);
            load_vec(X2, x2);
            s_rows_times2<RPL>(w, lane, x1, x2, g1, g2);
            load_vec(Xc, c);
            cx1 = dot8(c, x1);
            cx2 = dot8(c, x2);
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                z1[r] = d[r] * (g1[r] - h[r]);
                dza[r] = d[r] * (g2[r] - (s[r] - rz[r]));
                hz1 = fma(h[r], z1[r], hz1);
                hz2 = fma(h[r], dza[r], hz2);
            }
        }
        warp_sum2(hz1, hz2);
        const double kot = kap * tinv;
        const double den = cx1 + hz1 - kot;                     // < 0
        const double rden = fast_rcp(den);
        const double dta = (-rt + kap - cx2 - hz2) * rden;
        const double dka = -kap - kot * dta;
        const double kinv = fast_rcp(kap);
        double ratio = fmax(-dta * tinv, -dka * kinv);
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            dza[r] = fma(dta, z1[r], dza[r]);
            dsa[r] = -s[r] - s[r] * zinv[r] * dza[r];
            if (live[r]) ratio = fmax(ratio, fmax(-dsa[r] * sinv[r], -dza[r] * zinv[r]));
        }
        ratio = warp_max(ratio);
        const double alpha_aff = ratio > 1.0 ? fast_rcp(ratio) : 1.0;
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
        s_gt_times_slots(w, mk, 1, lane);
        if (own) X3[lane] = fma(-eta, rxl, w.R[lane]);
        __syncwarp();
        s_solve<1>(w.M, X3, lane);
        double dz[RPL];
        hz2 = 0.0;
        {
            double x3[NS], gc[RPL], c[NS];
            load_vec(X3, x3);
            s_rows_times<RPL>(w, lane, x3, gc);
            load_vec(Xc, c);
            cx2 = dot8(c, x3);
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                dz[r] = d[r] * (gc[r] - qc[r]);
                hz2 = fma(h[r], dz[r], hz2);
            }
        }
        hz2 = warp_sum(hz2);
        const double bk = -tau * kap + sigma * mu - dta * dka;
        const double dtau = (-eta * rt - bk * tinv - cx2 - hz2) * rden;
        const double dkap = (bk - kap * dtau) * tinv;
        ratio = fmax(-dtau * tinv, -dkap * kinv);
        double ds[RPL];
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            dz[r] = fma(dtau, z1[r], dz[r]);
            ds[r] = (bs[r] - s[r] * dz[r]) * zinv[r];
            if (live[r]) ratio = fmax(ratio, fmax(-ds[r] * sinv[r], -dz[r] * zinv[r]));
        }
        ratio = warp_max(ratio);
        const double amax = ratio > 0.0 ? fast_rcp(ratio) : 1e30;
        const double alpha = fmin(1.0, LP_STEP * amax);
        if (own) {
            xl = fma(alpha, fma(dtau, X1[lane], X3[lane]), xl);
            Xx[lane] = xl;
        }
        tau = fma(alpha, dtau, tau);
        kap = fma(alpha, dkap, kap);
#pragma unroll
        for (int r = 0; r < RPL; ++r)
            if (live[r]) { s[r] = fma(alpha, ds[r], s[r]); z[r] = fma(alpha, dz[r], z[r]); }
        __syncwarp();
    }
    if (lineal && res.status == ST_OPTIMAL) res.status = ST_UNBOUNDED;
    if (res.status != ST_OPTIMAL) return res;

    // ---- extract and polish (see lp_warp.cuh) ----
    const double tinv = 1.0 / tau;
    double xsl = xl * tinv;                       // lane-owned scaled solution
    const double cl0 = own ? cl_in : 0.0;
    const double f0 = warp_sum(cl0 * xsl);
    res.x = xsl;
    res.fun = f0;
    bool act[RPL];
    int nact = 0;
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        act[r] = live[r] && (z[r] > s[r]);
        nact += act[r] ? 1 : 0;
        w.d[lane + 32 * r] = act[r] ? 1.0 : 0.0;
    }
    nact = __reduce_add_sync(FULL_MASK, nact);
    if (own) Xx[lane] = xsl;
    __syncwarp();
    if (nact == 0) return res;
    s_normal_matrix(w, mk, lane);
    {
        double dmax = 1.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) dmax = fmax(dmax, w.M[j * NS + j]);
        __syncwarp();
        s_cholesky(w.M, n, 1e-9 * dmax, lane);
    }
    double gxp[RPL];
#pragma unroll 1
    for (int round = 0; round < 4; ++round) {
        double xs[NS];
        load_vec(Xx, xs);
        s_rows_times<RPL>(w, lane, xs, gxp);
        if (round == 3) break;
#pragma unroll
        for (int r = 0; r < RPL; ++r) w.V[lane + 32 * r] = act[r] ? h[r] - gxp[r] : 0.0;
        __syncwarp();
        s_gt_times_slots(w, mk, 1, lane);
        if (own) X3[lane] = w.R[lane];
        __syncwarp();
        s_solve<1>(w.M, X3, lane);
        if (own) { xsl += X3[lane]; Xx[lane] = xsl; }
        __syncwarp();
    }
    double slack = 1e300;
#pragma unroll
    for (int r = 0; r < RPL; ++r)
        if (live[r]) slack = fmin(slack, h[r] - gxp[r]);
    slack = warp_min(slack);
    const double f1 = warp_sum(cl0 * xsl);
    const bool accept = (slack >= -1e-9 * fmax(1.0, hmax)) && (fabs(f1 - f0) <= 1e-6 * fmax(1.0, fabs(f0)));
    if (accept) { res.x = xsl; res.fun = f1; }
    return res;
}

}  // namespace pb200

// polytope_b200: one LP per CTA -- the general solver behind pb200_lp_batch_big for LPs outside
// the envelope of the lane / warp kernels: any number of rows (extreme()'s final is_fulldim(Q) has
// one row per vertex, polytope/polytope.py:1666-1670; intersections and envelopes of large
// polytopes, :255-275, :1414-1464) and up to 64 columns (Chebyshev LPs of d >= 32).
//
//   min c'x  s.t.  Gx <= h,  x free          (polytope/solvers.py:76-106 contract)
//
// Same algorithm as lp_lane.cuh (Mehrotra predictor-corrector on the homogeneous self-dual
// embedding, LIPSOL pivot skipping, infeasibility / unboundedness certificates with the
// feasibility restart, active-set polish of the converged iterate), mapped onto 256 threads:
//   * G stays where the caller put it; every pass streams it through shared memory in chunks of
//     256 rows (coalesced copies), one thread per row computes the row's scalars from the staged
//     row and the n-vectors (shared memory, broadcast reads);
//   * sums over rows that yield more than a scalar -- the normal matrix M = G'DG and the products
//     G'v -- are "entry-owned": a thread owns up to 9 entries of (M | G'v1 | G'v2 | G'v3), keeps them
//     in registers across all chunks and accumulates coef_r * G[r][j] * G[r][k] out of the staged
//     chunk, so no cross-thread reduction is needed for them; scalars go through a block reduction;
//   * per-row iterates (s, z) live in a caller-provided workspace (2 m doubles per resident CTA);
//   * Cholesky (in place, lane-parallel over rows) and the triangular solves run on single warps.
// Control flow is uniform over the CTA: every thread holds bitwise identical scalars.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "lp_lane.cuh"

namespace pb200 {

constexpr int CT = 256;
constexpr int CTA_MAX_N = 64;
constexpr int CTA_MAXE = (CTA_MAX_N * (CTA_MAX_N + 1) / 2 + 3 * CTA_MAX_N + CT - 1) / CT;   // 9

struct CtaArgs {
    const double *G, *h, *c;
    const int32_t* m_rows;
    int B, m, n;
    long long g_stride;    // doubles between the G of consecutive LPs (0: one G shared by all)
    double *x, *fun;
    int8_t* status;
    int32_t* iters;
    double* ws;            // [grid][2 m]
    int* next;
};

// vector file (shared memory), n doubles each
enum { V_X = 0, V_C, V_V1, V_V2, V_V3, V_RXL, V_X1, V_XA, V_X3, V_XP, V_VP, V_TMP, V_COUNT };

struct CtaSmem {
    double* Gs;     // [CT][ld] staged chunk; column n holds 1.0
    double* cf;     // [4][CT] per-row coefficients of the entry-owned sums
    double* M;      // [n][ldm]
    double* vec;    // [V_COUNT][n]
    double* red;    // [8][8]
    int ld, ldm;
};
__host__ __device__ inline int cta_ld(int n) { return (n + 1) | 1; }
__host__ __device__ inline size_t cta_smem_doubles(int n) {
    return (size_t)CT * cta_ld(n) + 4 * CT + (size_t)n * (n + 1) + (size_t)V_COUNT * n + 64 + 8;
}

template <int K>
__device__ __forceinline__ void block_reduce(double* red, double (&v)[K], bool is_max) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < K; ++q)
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double t = __shfl_xor_sync(0xffffffffu, v[q], o);
            v[q] = is_max ? fmax(v[q], t) : v[q] + t;
        }
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int q = 0; q < K; ++q) red[wid * 8 + q] = v[q];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < K; ++q) {
        double t = red[q];
#pragma unroll
        for (int w = 1; w < CT / 32; ++w) t = is_max ? fmax(t, red[w * 8 + q]) : t + red[w * 8 + q];
        v[q] = t;
    }
}

// an entry of the entry-owned sums: acc += cf[t][r] * Gs[r][j] * Gs[r][k]
struct Entry { int j, k, t; };

// One sweep over the rows: rowfn(i, r, grow) for every row (thread r of the chunk), then every thread
// accumulates its `cnt` entries over the chunk.
template <class RowFn>
__device__ __forceinline__ void sweep(const CtaSmem& sm, const double* __restrict__ Gp, int m, int n, RowFn rowfn,
                                      const Entry* ent, int cnt, double* acc) {
    const int tid = threadIdx.x, ld = sm.ld;
    for (int q = 0; q < cnt; ++q) acc[q] = 0.0;
    for (int c0 = 0; c0 < m; c0 += CT) {
        const int rows = min(CT, m - c0);
        const double* src = Gp + (size_t)c0 * n;
        for (int e = tid; e < rows * n; e += CT) {
            const int r = e / n;
            sm.Gs[r * ld + (e - r * n)] = __ldg(src + e);
        }
        if (tid < rows) sm.Gs[tid * ld + n] = 1.0;
        __syncthreads();
        if (tid < rows) rowfn(c0 + tid, tid, sm.Gs + tid * ld);
        __syncthreads();
        for (int q = 0; q < cnt; ++q) {
            const Entry en = ent[q];
            const double* cf = sm.cf + en.t * CT;
            double a0 = 0.0, a1 = 0.0;
            int r = 0;
            for (; r + 1 < rows; r += 2) {
                a0 = fma(cf[r] * sm.Gs[r * ld + en.j], sm.Gs[r * ld + en.k], a0);
                a1 = fma(cf[r + 1] * sm.Gs[(r + 1) * ld + en.j], sm.Gs[(r + 1) * ld + en.k], a1);
            }
            if (r < rows) a0 = fma(cf[r] * sm.Gs[r * ld + en.j], sm.Gs[r * ld + en.k], a0);
            acc[q] += a0 + a1;
        }
        __syncthreads();
    }
}

__device__ __forceinline__ double sdot(const double* g, const double* v, int n) {
    double s0 = 0.0, s1 = 0.0;
    int j = 0;
    for (; j + 1 < n; j += 2) { s0 = fma(g[j], v[j], s0); s1 = fma(g[j + 1], v[j + 1], s1); }
    if (j < n) s0 = fma(g[j], v[j], s0);
    return s0 + s1;
}

// in-place Cholesky of M (lower, n x n, leading dimension ldm) by warp 0; diagonal ends as 1 / L_kk, vanishing
// pivots skipped (0).  Returns the number of skipped pivots through *skipped (shared memory).
__device__ __forceinline__ void cta_chol(double* M, int ldm, int n, double add_diag, int* skipped) {
    if (threadIdx.x < 32) {
        const int ln = threadIdx.x;
        int skip = 0;
        for (int k = 0; k < n; ++k) {
            const double dg = M[k * ldm + n] + add_diag;  // original diagonal, parked in column n
            const double p = M[k * ldm + k] + add_diag;
            const bool ok = (p > 1e-13 * dg) && (p > 1e-290);
            const double rinv = ok ? lane::rsqrt_(p) : 0.0;
            skip += ok ? 0 : 1;
            __syncwarp();
            if (ln == 0) M[k * ldm + k] = rinv;
            for (int i = k + 1 + ln; i < n; i += 32) M[i * ldm + k] *= rinv;
            __syncwarp();
            for (int i = k + 1 + ln; i < n; i += 32) {
                const double lik = M[i * ldm + k];
                for (int j = k + 1; j <= i; ++j) M[i * ldm + j] = fma(-lik, M[j * ldm + k], M[i * ldm + j]);
            }
            __syncwarp();
        }
        if (ln == 0) *skipped = skip;
    }
    __syncthreads();
}

// (L L') y = a in place, one warp per right-hand side (warp w solves rhs[w], w < nrhs)
__device__ __forceinline__ void cta_solve(const double* M, int ldm, int n, double* const* rhs, int nrhs) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (wid < nrhs) {
        double* a = rhs[wid];
        for (int k = 0; k < n; ++k) {
            if (lane == 0) a[k] *= M[k * ldm + k];
            __syncwarp();
            const double ak = a[k];
            for (int i = k + 1 + lane; i < n; i += 32) a[i] = fma(-M[i * ldm + k], ak, a[i]);
            __syncwarp();
        }
        for (int k = n - 1; k >= 0; --k) {
            if (lane == 0) a[k] *= M[k * ldm + k];
            __syncwarp();
            const double ak = a[k];
            for (int i = lane; i < k; i += 32) a[i] = fma(-M[k * ldm + i], ak, a[i]);
            __syncwarp();
        }
    }
    __syncthreads();
}

// entries of (lower triangle of M | nvec vectors): e -> (j, k, t); vectors use k = n (the column of ones)
__device__ __forceinline__ int make_entries(Entry* ent, int n, bool with_M, int nvec) {
    const int NT = with_M ? n * (n + 1) / 2 : 0;
    const int NE = NT + nvec * n;
    int cnt = 0;
    for (int e = threadIdx.x; e < NE && cnt < CTA_MAXE; e += CT, ++cnt) {
        if (e < NT) {
            int j = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
            while (j * (j + 1) / 2 > e) --j;
            while ((j + 1) * (j + 2) / 2 <= e) ++j;
            ent[cnt].j = j; ent[cnt].k = e - j * (j + 1) / 2; ent[cnt].t = 0;
        } else {
            const int v = (e - NT) / n;
            ent[cnt].j = (e - NT) - v * n; ent[cnt].k = n; ent[cnt].t = with_M ? v + 1 : v;
        }
    }
    return cnt;
}
// where the thread's entries go: M[j][k] (and the original diagonal into column n) or vector `dst[v]`
__device__ __forceinline__ void store_entries(const CtaSmem& sm, const Entry* ent, int cnt, const double* acc, int n, bool with_M,
                                              double* const* dst) {
    for (int q = 0; q < cnt; ++q) {
        const Entry en = ent[q];
        if (with_M && en.t == 0) {
            sm.M[en.j * sm.ldm + en.k] = acc[q];
            if (en.j == en.k) sm.M[en.j * sm.ldm + n] = acc[q];
        } else {
            dst[with_M ? en.t - 1 : en.t][en.j] = acc[q];
        }
    }
}

__global__ void __launch_bounds__(CT) lp_cta_kernel(const CtaArgs a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int sh_b, sh_skip;
    const int tid = threadIdx.x, n = a.n;
    CtaSmem sm;
    sm.ld = cta_ld(n);
    sm.ldm = n + 1;
    sm.Gs = smem;
    sm.cf = sm.Gs + (size_t)CT * sm.ld;
    sm.M = sm.cf + 4 * CT;
    sm.vec = sm.M + (size_t)n * (n + 1);
    sm.red = sm.vec + (size_t)V_COUNT * n;
    double* const vx = sm.vec + V_X * n;
    double* const vc = sm.vec + V_C * n;
    double* const v1 = sm.vec + V_V1 * n;
    double* const v2 = sm.vec + V_V2 * n;
    double* const v3 = sm.vec + V_V3 * n;
    double* const rxl = sm.vec + V_RXL * n;
    double* const X1 = sm.vec + V_X1 * n;
    double* const xa = sm.vec + V_XA * n;
    double* const X3 = sm.vec + V_X3 * n;
    double* const xp = sm.vec + V_XP * n;
    double* const vp = sm.vec + V_VP * n;
    double* const tmp = sm.vec + V_TMP * n;
    double* const sv = a.ws + (size_t)blockIdx.x * 2 * a.m;
    double* const zv = sv + a.m;
    using namespace lane;

    for (;;) {
        __syncthreads();
        if (tid == 0) sh_b = atomicAdd(a.next, 1);
        __syncthreads();
        const int b = sh_b;
        if (b >= a.B) break;
        const int m = a.m_rows ? min(max(a.m_rows[b], 0), a.m) : a.m;
        const double* Gp = a.G + (size_t)b * a.g_stride;
        const double* hp = a.h + (size_t)b * a.m;
        Entry ent[CTA_MAXE];
        double acc[CTA_MAXE];

        // ---- start point ----
        double r2[2] = {0.0, 0.0};
        double hmx[1] = {0.0};
        for (int i = tid; i < m; i += CT) {
            const double h = hp[i];
            sv[i] = fmax(h, 0.0) + 1.0;
            zv[i] = 1.0;
            r2[0] = fma(h, h, r2[0]);
            hmx[0] = fmax(hmx[0], fabs(h));
        }
        if (tid < n) { vc[tid] = a.c[(size_t)b * n + tid]; vx[tid] = 0.0; }
        block_reduce<2>(sm.red, r2, false);
        block_reduce<1>(sm.red, hmx, true);
        const double hmax = hmx[0];
        double cc2 = 0.0;
        for (int j = 0; j < n; ++j) cc2 = fma(vc[j], vc[j], cc2);
        const double nh2 = fmax(1.0, r2[0]);
        double nc2 = fmax(1.0, cc2);
        const double rmu = 1.0 / (double)(m + 1);
        double tau = 1.0, kap = 1.0;
        bool lineal = false;
        int it = 0, status = ITER_LIMIT;
        bool converged = false;

        for (;;) {
            const double csel = lineal ? 0.0 : 1.0;
            // ---- pass A: residuals, M = G'DG, G'[z | D h | D q_aff] ----
            int cnt = make_entries(ent, n, true, 3);
            double sc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};     // sz, hz, rz2, gxs2, dhh, dhq
            sweep(sm, Gp, m, n,
                  [&](int i, int r, const double* g) {
                      const double h = hp[i], s = sv[i], z = zv[i];
                      const double gx = sdot(g, vx, n);
                      const double d = z * rcp(s);
                      const double gxs = gx + s;
                      const double rz = gxs - h * tau;
                      sc[0] = fma(s, z, sc[0]);
                      sc[1] = fma(h, z, sc[1]);
                      sc[2] = fma(rz, rz, sc[2]);
                      sc[3] = fma(gxs, gxs, sc[3]);
                      const double dh = d * h, dq = z - d * rz;
                      sc[4] = fma(dh, h, sc[4]);
                      sc[5] = fma(dh, s - rz, sc[5]);
                      sm.cf[r] = d; sm.cf[CT + r] = z; sm.cf[2 * CT + r] = dh; sm.cf[3 * CT + r] = dq;
                  },
                  ent, cnt, acc);
            {
                double* const dst[3] = {v1, v2, v3};
                store_entries(sm, ent, cnt, acc, n, true, dst);
            }
            block_reduce<6>(sm.red, sc, false);           // (its barriers also publish M and the vectors)
            const double sz = sc[0], hz = sc[1], rz2 = sc[2], gxs2 = sc[3], dhh = sc[4], dhq = sc[5];
            double rx2 = 0.0, cx = 0.0;
            for (int j = 0; j < n; ++j) {
                const double rj = fma(csel * vc[j], tau, v1[j]);
                rx2 = fma(rj, rj, rx2);
                cx = fma(csel * vc[j], vx[j], cx);
            }
            if (tid < n) rxl[tid] = fma(csel * vc[tid], tau, v1[tid]);
            const double rt = cx + hz + kap;
            const double mu = (sz + tau * kap) * rmu;
            const double tinv = rcp(tau);
            const double t2 = tinv * tinv;
            const double pcost = cx * tinv, dcost = -hz * tinv;
            const double gap = sz * t2;
            const double gapref = pcost < 0.0 ? -pcost : (dcost > 0.0 ? dcost : 0.0);
            bool restart = false;
            if (!(mu == mu) || !(fabs(tau) < 1e300) || !(fabs(cx) < 1e300)) { status = NUMERICAL; break; }
            converged = rz2 * t2 <= FEAS_TOL * FEAS_TOL * nh2 &&
                        ((rx2 * t2 <= FEAS_TOL * FEAS_TOL * nc2 && (gap <= GAP_TOL || gap <= GAP_TOL * gapref)) ||
                         (rx2 * t2 <= STALL_DRES * STALL_DRES * nc2 && (gap <= STALL_GAP || gap <= STALL_GAP * gapref)));
            if (converged) {
                if (lineal) status = UNBOUNDED;
                break;
            }
            if (tau < 1e-3 * kap) {
                if (hz < 0.0) {
                    double gz2 = 0.0;
                    for (int j = 0; j < n; ++j) gz2 = fma(v1[j], v1[j], gz2);
                    if (sqrt(gz2 * nh2 / nc2) <= 10.0 * FEAS_TOL * (-hz)) { status = INFEASIBLE; break; }
                }
                if (cx < 0.0 && sqrt(gxs2 * nc2 / nh2) <= 10.0 * FEAS_TOL * (-cx)) {
                    // improving recession direction: settle feasibility first (c = 0), as lp_lane.cuh
                    lineal = true;
                    restart = true;
                    nc2 = 1.0; tau = 1.0; kap = 1.0;
                    __syncthreads();
                    if (tid < n) vx[tid] = 0.0;
                    for (int i = tid; i < m; i += CT) { sv[i] = fmax(hp[i], 0.0) + 1.0; zv[i] = 1.0; }
                    if (it == 0) it = 1;
                    __syncthreads();
                }
            }
            if (restart) continue;
            if (it == MAX_ITER) break;                       // status stays ITER_LIMIT
            // ---- factor ----
            cta_chol(sm.M, sm.ldm, n, 0.0, &sh_skip);
            if (it == 0 && sh_skip && !lineal) {
                // G is column-rank deficient: c with a component in null(G) -> feasibility problem, report 3
                if (tid < n) tmp[tid] = vc[tid];
                __syncthreads();
                { double* const rh[1] = {tmp}; cta_solve(sm.M, sm.ldm, n, rh, 1); }
                cnt = make_entries(ent, n, false, 1);
                sweep(sm, Gp, m, n,
                      [&](int i, int r, const double* g) { sm.cf[r] = zv[i] * rcp(sv[i]) * sdot(g, tmp, n); },
                      ent, cnt, acc);
                __syncthreads();
                { double* const dst[1] = {vp}; store_entries(sm, ent, cnt, acc, n, false, dst); }
                __syncthreads();
                double rmax = 0.0, cmax = 0.0;
                for (int j = 0; j < n; ++j) { rmax = fmax(rmax, fabs(vc[j] - vp[j])); cmax = fmax(cmax, fabs(vc[j])); }
                if (rmax > 1e-9 * fmax(cmax, 1e-300)) {
                    lineal = true;
                    nc2 = 1.0;
                    it = 1;
                    continue;
                }
            }
            // ---- predictor ----
            if (tid < n) { X1[tid] = fma(-csel, vc[tid], v2[tid]); xa[tid] = v3[tid] - rxl[tid]; }
            __syncthreads();
            { double* const rh[2] = {X1, xa}; cta_solve(sm.M, sm.ldm, n, rh, 2); }
            double cx1 = 0.0, cx2 = 0.0, hz1 = -dhh, hz2 = -dhq;
            for (int j = 0; j < n; ++j) {
                cx1 = fma(csel * vc[j], X1[j], cx1);
                cx2 = fma(csel * vc[j], xa[j], cx2);
                hz1 = fma(v2[j], X1[j], hz1);
                hz2 = fma(v2[j], xa[j], hz2);
            }
            const double kot = kap * tinv;
            const double den = cx1 + hz1 - kot;
            const double rden = rcp(den);
            const double dta = (-rt + kap - cx2 - hz2) * rden;
            const double dka = -kap - kot * dta;
            const double kinv = rcp(kap);
            __syncthreads();
            if (tid < n) xa[tid] = fma(dta, X1[tid], xa[tid]);
            __syncthreads();
            double ratio[1] = {fmax(fmax(-dta * tinv, -dka * kinv), 0.0)};
            sweep(sm, Gp, m, n,
                  [&](int i, int r, const double* g) {                 // pass B: affine step length
                      const double h = hp[i];
                      const double q = h * tau - sdot(g, vx, n);
                      const double w = rcp(sv[i]) * (sdot(g, xa, n) - q - dta * h);
                      ratio[0] = fmax(ratio[0], fmax(1.0 + w, -w));
                  },
                  ent, 0, acc);
            block_reduce<1>(sm.red, ratio, true);
            const double alpha_aff = ratio[0] > 1.0 ? rcp(ratio[0]) : 1.0;
            const double om = 1.0 - alpha_aff;
            const double sigma = om * om * om;
            const double eta = 1.0 - sigma;
            const double smu = sigma * mu;
            // ---- corrector right-hand side (pass C) ----
            cnt = make_entries(ent, n, false, 1);
            double dhqc[1] = {0.0};
            sweep(sm, Gp, m, n,
                  [&](int i, int r, const double* g) {
                      const double h = hp[i], s = sv[i], z = zv[i];
                      const double isz = rcp(s * z);
                      const double sinv = isz * z, zinv = isz * s;
                      const double d = z * sinv;
                      const double q = h * tau - sdot(g, vx, n);
                      const double rz = s - q;
                      const double dza = d * (sdot(g, xa, n) - q - dta * h);
                      const double dsa = -s - s * zinv * dza;
                      const double bs = -s * z + smu - dsa * dza;
                      const double qc = -eta * rz - bs * zinv;
                      const double dqc = d * qc;
                      dhqc[0] = fma(dqc, h, dhqc[0]);
                      sm.cf[r] = dqc;
                  },
                  ent, cnt, acc);
            { double* const dst[1] = {X3}; store_entries(sm, ent, cnt, acc, n, false, dst); }
            block_reduce<1>(sm.red, dhqc, false);
            if (tid < n) X3[tid] = fma(-eta, rxl[tid], X3[tid]);
            __syncthreads();
            { double* const rh[1] = {X3}; cta_solve(sm.M, sm.ldm, n, rh, 1); }
            double v2x3 = 0.0, cx3 = 0.0;
            for (int j = 0; j < n; ++j) { v2x3 = fma(v2[j], X3[j], v2x3); cx3 = fma(csel * vc[j], X3[j], cx3); }
            const double hz3 = v2x3 - dhqc[0];
            const double bk = -tau * kap + smu - dta * dka;
            const double dtau = (-eta * rt - bk * tinv - cx3 - hz3) * rden;
            const double dkap = (bk - kap * dtau) * tinv;
            __syncthreads();
            if (tid < n) X3[tid] = fma(dtau, X1[tid], X3[tid]);
            __syncthreads();
            ratio[0] = fmax(fmax(-dtau * tinv, -dkap * kinv), 0.0);
            auto step_row = [&](int i, const double* g, double& ds, double& dz, double& sinv, double& zinv) {
                const double h = hp[i], s = sv[i], z = zv[i];
                const double isz = rcp(s * z);
                sinv = isz * z; zinv = isz * s;
                const double d = z * sinv;
                const double q = h * tau - sdot(g, vx, n);
                const double rz = s - q;
                const double dza = d * (sdot(g, xa, n) - q - dta * h);
                const double dsa = -s - s * zinv * dza;
                const double bs = -s * z + smu - dsa * dza;
                const double qc = -eta * rz - bs * zinv;
                dz = d * (sdot(g, X3, n) - qc - dtau * h);
                ds = (bs - s * dz) * zinv;
            };
            sweep(sm, Gp, m, n,
                  [&](int i, int r, const double* g) {                 // pass D: step length
                      double ds, dz, sinv, zinv;
                      step_row(i, g, ds, dz, sinv, zinv);
                      ratio[0] = fmax(ratio[0], fmax(-ds * sinv, -dz * zinv));
                  },
                  ent, 0, acc);
            block_reduce<1>(sm.red, ratio, true);
            const double amax = ratio[0] > 0.0 ? rcp(ratio[0]) : 1e30;
            const double alpha = fmin(1.0, STEP * amax);
            sweep(sm, Gp, m, n,
                  [&](int i, int r, const double* g) {                 // pass E: take the step in (s, z)
                      double ds, dz, sinv, zinv;
                      step_row(i, g, ds, dz, sinv, zinv);
                      sv[i] = fma(alpha, ds, sv[i]);
                      zv[i] = fma(alpha, dz, zv[i]);
                  },
                  ent, 0, acc);
            if (tid < n) vx[tid] = fma(alpha, X3[tid], vx[tid]);
            tau = fma(alpha, dtau, tau);
            kap = fma(alpha, dkap, kap);
            ++it;
            __syncthreads();
        }

        // ---- converged: active-set polish of the iterate (rows with z > s are the optimal face) ----
        double fun = 0.0;
        if (converged && !lineal) {
            status = OPTIMAL;
            const double te = 1.0 / tau;
            __syncthreads();
            if (tid < n) xp[tid] = vx[tid] * te;
            double na[1] = {0.0};
            for (int i = tid; i < m; i += CT) na[0] += zv[i] > sv[i] ? 1.0 : 0.0;
            block_reduce<1>(sm.red, na, false);
            double f0 = 0.0;
            for (int j = 0; j < n; ++j) f0 = fma(vc[j], xp[j], f0);
            fun = f0;
            if (tid < n) X3[tid] = xp[tid];                    // X3 holds the answer; xp is the working point
            __syncthreads();
            if (na[0] > 0.0) {
                int cnt = make_entries(ent, n, true, 0);
                sweep(sm, Gp, m, n, [&](int i, int r, const double*) { sm.cf[r] = zv[i] > sv[i] ? 1.0 : 0.0; }, ent, cnt, acc);
                store_entries(sm, ent, cnt, acc, n, true, nullptr);
                __syncthreads();
                double dmax = 1.0;
                for (int j = 0; j < n; ++j) dmax = fmax(dmax, sm.M[j * sm.ldm + j]);
                cta_chol(sm.M, sm.ldm, n, 1e-9 * dmax, &sh_skip);
                const double scale = fmax(1.0, hmax);
                cnt = make_entries(ent, n, false, 1);
                for (int round = 0; round < 4; ++round) {
                    double mx[2] = {0.0, -1e300};                // max |residual| on the face, max violation
                    sweep(sm, Gp, m, n,
                          [&](int i, int r, const double* g) {
                              const double rr = hp[i] - sdot(g, xp, n);
                              mx[1] = fmax(mx[1], -rr);
                              const bool act = zv[i] > sv[i];
                              if (act) mx[0] = fmax(mx[0], fabs(rr));
                              sm.cf[r] = act ? rr : 0.0;
                          },
                          ent, cnt, acc);
                    { double* const dst[1] = {vp}; store_entries(sm, ent, cnt, acc, n, false, dst); }
                    block_reduce<2>(sm.red, mx, true);
                    const bool feasible = mx[1] <= 1e-9 * scale;
                    const bool settled = mx[0] <= 1e-13 * scale || round == 3;
                    if (settled) {
                        double f1 = 0.0;
                        for (int j = 0; j < n; ++j) f1 = fma(vc[j], xp[j], f1);
                        if (feasible && fabs(f1 - f0) <= 1e-6 * fmax(1.0, fabs(f0))) {
                            fun = f1;
                            __syncthreads();
                            if (tid < n) X3[tid] = xp[tid];
                        }
                        break;
                    }
                    { double* const rh[1] = {vp}; cta_solve(sm.M, sm.ldm, n, rh, 1); }
                    if (tid < n) xp[tid] += vp[tid];
                    __syncthreads();
                }
            }
            __syncthreads();
        }
        const double nan = __longlong_as_double(0x7ff8000000000000ll);     // x / fun of an LP without optimum, as pb200_lp_batch
        if (tid < n) a.x[(size_t)b * n + tid] = status == OPTIMAL ? X3[tid] : nan;
        if (tid == 0) {
            a.fun[b] = status == OPTIMAL ? fun : nan;
            a.status[b] = (int8_t)status;
            if (a.iters) a.iters[b] = it;
        }
    }
}

static int cta_grid(int B, int n, size_t* smem_out) {
    const size_t smem = cta_smem_doubles(n) * sizeof(double);
    *smem_out = smem;
    const int sms = sm_count();
    if (!sms) return 0;
    if (smem > 227 * 1024) return -1;
    if (cudaFuncSetAttribute(lp_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lp_cta_kernel, CT, smem) != cudaSuccess || per_sm < 1) return -1;
    const long long g = (long long)sms * per_sm;
    return (int)(B < g ? B : g);
}

}  // namespace pb200

using namespace pb200;

extern "C" {

size_t pb200_lp_big_workspace_bytes(int B, int m, int n) {
    if (B < 0 || m < 1 || n < 1 || n > CTA_MAX_N) return 0;
    size_t smem;
    const int grid = cta_grid(B > 0 ? B : 1, n, &smem);
    if (grid <= 0) return 0;
    return 256 + (size_t)grid * 2 * (size_t)m * sizeof(double);
}

int pb200_lp_batch_big(const double* G, const double* h, const double* c, const int32_t* m_rows, int B, int m, int n, int shared_G,
                       double* x, double* fun, int8_t* status, int32_t* iters, void* workspace, size_t workspace_bytes, void* stream) {
    if (B == 0) return PB200_OK;
    if (B < 0 || !G || !h || !c || !x || !fun || !status || !workspace) return fail(PB200_EINVAL, "pb200_lp_batch_big: null pointer or negative batch");
    if (m < 1 || n < 1 || n > CTA_MAX_N) return fail(PB200_EUNSUPPORTED, "pb200_lp_batch_big: need m >= 1 and 1 <= n <= 64");
    size_t smem;
    const int grid = cta_grid(B, n, &smem);
    if (grid == 0) return fail(PB200_ECUDA, "pb200_lp_batch_big: device query failed");
    if (grid < 0) return fail(PB200_EUNSUPPORTED, "pb200_lp_batch_big: LP does not fit in shared memory");
    if (256 + (size_t)grid * 2 * (size_t)m * sizeof(double) > workspace_bytes) return fail(PB200_EWORKSPACE, "pb200_lp_batch_big: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    PB_CHECK_CUDA(cudaMemsetAsync(workspace, 0, 256, st));
    CtaArgs a;
    a.G = G; a.h = h; a.c = c; a.m_rows = m_rows; a.B = B; a.m = m; a.n = n;
    a.g_stride = shared_G ? 0 : (long long)m * n;
    a.x = x; a.fun = fun; a.status = status; a.iters = iters;
    a.ws = (double*)((char*)workspace + 256);
    a.next = (int*)workspace;
    lp_cta_kernel<<<grid, CT, smem, st>>>(a);
    count_launch();
    PB_CHECK_CUDA(cudaGetLastError());
    return PB200_OK;
}

}  // extern "C"

// polytope_b200: kernels around the warp LP solver and the C ABI
// (include/polytope_b200.h).  sm_100a only.
//
// Every LP-solving kernel is the same persistent "one LP per warp" loop,
// parameterised by a Problem type that knows how to stage (c, G, h) of work
// item t into the warp's shared-memory scratch and what to do with the result:
//
//   GenericLP    lpsolve(c, G, h)                       polytope/solvers.py:76-106
//   ChebyLP      cheby_ball                              polytope/polytope.py:1280-1300
//   BboxLP       bounding_box (2d LPs per polytope)      polytope/polytope.py:1362-1411
//   RowLP        reduce()'s per-row redundancy LP        polytope/polytope.py:1142-1160
//   AdjacentLP   is_adjacent (stack, inflate, cheby)     polytope/polytope.py:1856-1866
//
// Each warp stages its own copy of the (tiny) constraint matrix; LPs of the
// same polytope re-read it from L2, which keeps HBM traffic at the algorithmic
// bytes while leaving no warp idle (a whole cfg2 batch is 23 MB, L2 is 126 MB).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "lp_warp.cuh"
#include "lp_warp_small.cuh"
#include "staging.cuh"
#include "normalize.cuh"
#include "lane_kernel.cuh"

namespace pb200 {

constexpr int WPC = 4;                  // warps per CTA of the LP kernels
static thread_local char g_err[512] = "";
static long long g_launches = 0;
static int g_sm_count = 0;

int fail(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
char* err_buf(size_t* cap) { *cap = sizeof(g_err); return g_err; }
void count_launch(int n) { g_launches += n; }
int sm_count() {
    if (!g_sm_count) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
            g_sm_count = 0;
            fail(PB200_ECUDA, "cannot query the SM count of the current device");
        }
    }
    return g_sm_count;
}

// ------------------------------------------------------------------------
// Problems
// ------------------------------------------------------------------------
struct GenericLP {
    static constexpr bool kSnap = true;      // exact vertex coordinates on axis-aligned active rows
    const double *G, *h, *c;
    const int32_t* m_rows;
    int m, nn;
    double *x, *fun;
    int8_t* status;
    int32_t* iters;
    __device__ int n() const { return nn; }
    template <int RPL, class S>
    __device__ bool load(long long t, const S& w, int lane, int& mm, double& cc, double (&hh)[RPL], long long& cached) const {
        mm = m_rows ? min(max(m_rows[t], 0), m) : m;
        stage_first_rows<RPL>(w, G + (size_t)t * m * nn, h + (size_t)t * m, mm, nn, lane, hh);
        cc = lane < nn ? c[(size_t)t * nn + lane] : 0.0;
        return true;
    }
    template <int RPL>
    __device__ void store(long long t, int lane, const LpResult& res) const {
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        if (lane < nn) x[(size_t)t * nn + lane] = res.status == ST_OPTIMAL ? res.x : nan;
        if (lane == 0) {
            fun[t] = res.status == ST_OPTIMAL ? res.fun : nan;
            status[t] = (int8_t)res.status;
            if (iters) iters[t] = res.iters;
        }
    }
};

struct ChebyLP {
    static constexpr bool kSnap = false;
    const double *A, *b;
    const int32_t* m_rows;
    const uint64_t* rows;     // nullable row masks
    const uint32_t* skip_flags;  // nullable: skip polytope when (flags & skip_mask)
    uint32_t skip_mask;
    int m, d;
    double *r, *xc;
    int8_t* status;
    int32_t* lp_iters;           // nullable: = iterations of this LP
    __device__ int n() const { return d + 1; }
    template <int RPL, class S>
    __device__ bool load(long long p, const S& w, int lane, int& mm, double& cc, double (&hh)[RPL], long long& cached) const {
        if (skip_flags && (skip_flags[p] & skip_mask)) return false;
        const double* Ap = A + (size_t)p * m * d;
        const double* bp = b + (size_t)p * m;
        if (rows) {
            mm = stage_masked_rows<RPL>(w, Ap, bp, m, d, rows[p] & low_bits(m), lane, hh);
        } else {
            mm = m_rows ? min(max(m_rows[p], 0), m) : m;
            stage_first_rows<RPL>(w, Ap, bp, mm, d, lane, hh);
        }
        append_norm_column<RPL>(w, mm, d, lane);
        cc = lane == d ? -1.0 : 0.0;
        return true;
    }
    template <int RPL>
    __device__ void store(long long p, int lane, const LpResult& res) const {
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        const double v = res.status == ST_OPTIMAL ? res.x : nan;
        if (lane < d) xc[(size_t)p * d + lane] = v;
        if (lane == d) r[p] = v;
        if (lane == 0) {
            status[p] = (int8_t)res.status;
            if (lp_iters) lp_iters[p] = res.iters;
        }
    }
};

struct BboxLP {
    static constexpr bool kSnap = false;
    const double *A, *b;
    const int32_t* m_rows;
    const uint64_t* rows;        // nullable row masks
    const uint32_t* need_flags;  // nullable: run only when (flags & need_mask)
    uint32_t need_mask;
    int m, d, renorm;
    double *val_lo, *val_hi;     // [P][d] each: optimised coordinate of the lower / upper LP
    int8_t* status;              // [P][2d]
    int32_t* lp_iters;           // nullable: += iterations
    __device__ int n() const { return d; }
    template <int RPL, class S>
    __device__ bool load(long long t, const S& w, int lane, int& mm, double& cc, double (&hh)[RPL], long long& cached) const {
        const long long p = t / (2 * d);
        const int q = (int)(t - p * 2 * d);
        if (need_flags && !(need_flags[p] & need_mask)) return false;
        if (p != cached) {          // the 2d LPs of a polytope share G and h: stage once per warp
            const double* Ap = A + (size_t)p * m * d;
            const double* bp = b + (size_t)p * m;
            if (rows) {
                mm = stage_masked_rows<RPL>(w, Ap, bp, m, d, rows[p] & low_bits(m), lane, hh);
            } else {
                mm = m_rows ? min(max(m_rows[p], 0), m) : m;
                stage_first_rows<RPL>(w, Ap, bp, mm, d, lane, hh);
            }
            if (renorm) renormalize_rows<RPL>(w, mm, d, lane, hh);
#pragma unroll
            for (int r = 0; r < RPL; ++r) w.hb[lane + 32 * r] = hh[r];
            cached = p;
            __syncwarp();
        } else {
            mm = rows ? __popcll(rows[p] & low_bits(m)) : (m_rows ? min(max(m_rows[p], 0), m) : m);
#pragma unroll
            for (int r = 0; r < RPL; ++r) hh[r] = w.hb[lane + 32 * r];
        }
        const int i = q < d ? q : q - d;
        cc = lane == i ? (q < d ? 1.0 : -1.0) : 0.0;
        return true;
    }
    template <int RPL>
    __device__ void store(long long t, int lane, const LpResult& res) const {
        const long long p = t / (2 * d);
        const int q = (int)(t - p * 2 * d);
        const int i = q < d ? q : q - d;
        if (lane == i) {
            (q < d ? val_lo : val_hi)[p * d + i] = res.status == ST_OPTIMAL ? res.x : 0.0;
            status[t] = (int8_t)res.status;
            if (lp_iters) atomicAdd(lp_iters + p, res.iters);
        }
    }
};

// reduce()'s row loop (polytope.py:1142-1160).  Work item t = p*m + k, k-th
// surviving row of polytope p.  h reproduces the reference's in-place
// `h[k] += 0.1; ...; h[k] -= 0.1`: rows before k carry the one-ulp drift.
struct RowLP {
    static constexpr bool kSnap = false;
    const double *A, *b;        // constructor-normalised
    const uint64_t* rows;       // surviving rows (after duplicate / bbox filters)
    uint32_t* flags;
    uint32_t run_mask;          // run only when (flags & run_mask)
    int m, d;
    double abs_tol;
    unsigned long long* keep;   // OR-accumulated
    int32_t* lp_iters;          // nullable: += interior-point iterations of each row LP
    __device__ int n() const { return d; }
    template <int RPL, class S>
    __device__ bool load(long long t, const S& w, int lane, int& mm, double& cc, double (&hh)[RPL], long long& cached) const {
        const long long p = t / m;
        const int k = (int)(t - p * m);
        if (!(flags[p] & run_mask)) return false;
        const uint64_t mask = rows[p] & low_bits(m);
        if (k >= __popcll(mask)) return false;
        if (p != cached) {          // the row LPs of a polytope share G: stage once per warp
            mm = stage_masked_rows<RPL>(w, A + (size_t)p * m * d, b + (size_t)p * m, m, d, mask, lane, hh);
#pragma unroll
            for (int r = 0; r < RPL; ++r) w.hb[lane + 32 * r] = hh[r];
            cached = p;
            __syncwarp();
        } else {
            mm = __popcll(mask);
#pragma unroll
            for (int r = 0; r < RPL; ++r) hh[r] = w.hb[lane + 32 * r];
        }
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const int i = lane + 32 * r;
            if (i < k) hh[r] = __dadd_rn(__dadd_rn(hh[r], 0.1), -0.1);
            else if (i == k) hh[r] = __dadd_rn(hh[r], 0.1);
        }
        cc = lane < d ? -w.G[lane * w.MP + k] : 0.0;
        return true;
    }
    template <int RPL>
    __device__ void store(long long t, int lane, const LpResult& res) const {
        const long long p = t / m;
        const int k = (int)(t - p * m);
        if (lane != 0) return;
        const uint64_t mask = rows[p] & low_bits(m);
        const int orig = nth_set_bit(mask, k);
        bool kept = false;
        if (res.status == ST_OPTIMAL) {
            const double hk = __dadd_rn(__dadd_rn(b[(size_t)p * m + orig], 0.1), -0.1);
            kept = (-res.fun - hk) > abs_tol;
        } else if (res.status == ST_UNBOUNDED) {
            kept = true;
        } else {
            atomicOr(flags + p, PB200_F_LPFAIL);
        }
        if (kept) atomicOr(keep + p, 1ull << orig);
        if (lp_iters) atomicAdd(lp_iters + p, res.iters);
    }
};

// is_adjacent, overlap=True (polytope.py:1856-1866): rows of both cells, b + tol,
// constructor normalisation, Chebyshev LP, radius > tol/10.
struct AdjacentLP {
    static constexpr bool kSnap = false;
    const double *A, *b;
    int ncell, mc, d;
    const int32_t *pi, *pj;
    double abs_tol;
    uint8_t* adjacent;
    double* radius;
    int8_t* status;
    __device__ int n() const { return d + 1; }
    __device__ void pair(long long t, int& i, int& j) const {
        if (pi) { i = pi[t]; j = pj[t]; return; }
        // t = i(i-1)/2 + j, j < i
        long long ii = (long long)((1.0 + sqrt(1.0 + 8.0 * (double)t)) * 0.5);
        while (ii * (ii - 1) / 2 > t) --ii;
        while ((ii + 1) * ii / 2 <= t) ++ii;
        i = (int)ii;
        j = (int)(t - ii * (ii - 1) / 2);
    }
    template <int RPL, class S>
    __device__ bool load(long long t, const S& w, int lane, int& mm, double& cc, double (&hh)[RPL], long long& cached) const {
        int ci, cj;
        pair(t, ci, cj);
        mm = 2 * mc;
        zero_G<RPL>(w, lane);
        const int total = mc * d;
        const double* A1 = A + (size_t)ci * total;
        const double* A2 = A + (size_t)cj * total;
        for (int e = lane; e < total; e += 32) {
            const int i = e / d, j = e - i * d;
            w.G[j * w.MP + i] = __ldg(A1 + e);
            w.G[j * w.MP + mc + i] = __ldg(A2 + e);
        }
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const int i = lane + 32 * r;
            double v = 0.0;
            if (i < mc) v = __dadd_rn(__ldg(b + (size_t)ci * mc + i), abs_tol);
            else if (i < 2 * mc) v = __dadd_rn(__ldg(b + (size_t)cj * mc + i - mc), abs_tol);
            hh[r] = v;
        }
        __syncwarp();
        renormalize_rows<RPL>(w, mm, d, lane, hh);
        append_norm_column<RPL>(w, mm, d, lane);
        cc = lane == d ? -1.0 : 0.0;
        return true;
    }
    template <int RPL>
    __device__ void store(long long t, int lane, const LpResult& res) const {
        if (lane != d) return;
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        const double rr = res.status == ST_OPTIMAL ? res.x : nan;
        adjacent[t] = (res.status == ST_OPTIMAL && rr > abs_tol / 10) ? 1 : 0;
        if (radius) radius[t] = rr;
        if (status) status[t] = (int8_t)res.status;
    }
};

// ------------------------------------------------------------------------
// the one LP kernel
// ------------------------------------------------------------------------
template <int RPL, class Prob>
__global__ void __launch_bounds__(WPC * 32) lp_kernel(const Prob prob, long long n_items, const uint32_t* gate) {
    if (gate != nullptr && *gate == 0u) return;     // retry launch with nothing to retry (see launch_persistent)
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, wib = __reduce_max_sync(FULL_MASK, threadIdx.x >> 5);   // provably uniform
    const int n = prob.n();
    const WarpScratch w = lp_carve(smem + (size_t)wib * lp_scratch_doubles(RPL, n), RPL, n);
    // each warp owns a contiguous range of items, so consecutive LPs of one
    // polytope (same G) land on the same warp and are staged once
    const long long nwarps = (long long)gridDim.x * WPC;
    const long long chunk = (n_items + nwarps - 1) / nwarps;
    const long long t0 = ((long long)blockIdx.x * WPC + wib) * chunk;
    const long long t1 = t0 + chunk < n_items ? t0 + chunk : n_items;
    long long cached = -1;
    for (long long t = t0; t < t1; ++t) {
        int m;
        double c;
        double h[RPL];
        if (PB_UNI(!prob.template load<RPL>(t, w, lane, m, c, h, cached))) continue;
        LpResult res = lp_solve_warp<RPL>(w, m, n, c, h);
        if constexpr (Prob::kSnap) {
            if (PB_UNI(res.status == ST_OPTIMAL)) {
                res.x = snap_axis_rows<RPL>(w, w.u, m, n, lane, h, res.x);
                res.fun = warp_sum(lane < n ? c * res.x : 0.0);
            }
        }
        prob.template store<RPL>(t, lane, res);
        __syncwarp();
    }
}

// n <= 8: replicated-register solver (lp_warp_small.cuh)
#ifndef PB200_SMALL_MINB
#define PB200_SMALL_MINB 4
#endif
template <int RPL, class Prob>
__global__ void __launch_bounds__(WPC * 32, PB200_SMALL_MINB) lp_kernel_small(const Prob prob, long long n_items, const uint32_t* gate) {
    if (gate != nullptr && *gate == 0u) return;
    extern __shared__ __align__(16) double smem[];
    // warp index through a warp reduction: the result is provably warp-uniform, so the
    // item loop below (and everything it controls) is uniform control flow for ptxas
    const int lane = threadIdx.x & 31, wib = __reduce_max_sync(FULL_MASK, threadIdx.x >> 5);
    const int n = prob.n();
    const SmallScratch w = lps_carve(smem + (size_t)wib * lps_scratch_doubles(RPL), RPL);
    const long long nwarps = (long long)gridDim.x * WPC;
    const long long chunk = (n_items + nwarps - 1) / nwarps;
    const long long t0 = ((long long)blockIdx.x * WPC + wib) * chunk;
    const long long t1 = t0 + chunk < n_items ? t0 + chunk : n_items;
    long long cached = -1;
    for (long long t = t0; t < t1; ++t) {
        int m;
        double cl;
        double h[RPL];
        if (PB_UNI(!prob.template load<RPL>(t, w, lane, m, cl, h, cached))) continue;
        const SmallResult sr = lp_solve_small<RPL>(w, m, n, cl, h);
        LpResult res;
        res.status = sr.status; res.iters = sr.iters; res.fun = sr.fun; res.x = sr.x;
        if constexpr (Prob::kSnap) {
            if (PB_UNI(res.status == ST_OPTIMAL)) {
                res.x = snap_axis_rows<RPL>(w, w.X, m, n, lane, h, res.x);
                res.fun = warp_sum(lane < n ? cl * res.x : 0.0);
            }
        }
        prob.template store<RPL>(t, lane, res);
        __syncwarp();
    }
}

// Optional per-stage timing of pb200_reduce_batch with CUDA events recorded on
// the launch stream (bench.py's live roofline measurement).
constexpr int N_STAGES = 7;   // normalize, cheby, prefilter+plan, bbox, candidates, rows, finalize
static bool g_profile = false;
static cudaEvent_t g_ev[N_STAGES + 1];
static bool g_ev_ready = false, g_ev_valid = false;
static inline void stage_mark(int i, cudaStream_t st) {
    if (!g_profile) return;
    if (!g_ev_ready) {
        for (int k = 0; k <= N_STAGES; ++k) cudaEventCreate(&g_ev[k]);
        g_ev_ready = true;
    }
    cudaEventRecord(g_ev[i], st);
    if (i == N_STAGES) g_ev_valid = true;
}

template <class Kern, class Prob>
// `gate` (nullable, device): the kernel returns at once when *gate == 0.  The retry launches of the reduce pipeline
// pass the word the lane kernels set when an LP ended without a verdict: an ungated empty retry launch still walks
// every item's flag (17 + 31 us per cfg2 step, ncu r02ar).
static int launch_persistent(Kern kern, size_t smem, const Prob& prob, long long n_items, cudaStream_t st, const uint32_t* gate = nullptr) {
    if (smem > 227 * 1024) return fail(PB200_EUNSUPPORTED, "LP too large for shared memory");
    PB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (!sm_count()) return PB200_ECUDA;
    int per_sm = 0;
    PB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WPC * 32, smem));
    if (per_sm < 1) return fail(PB200_EUNSUPPORTED, "LP kernel does not fit on an SM");
    // persistent grid: a whole number of waves (multiple of the SM count)
    long long grid = (long long)g_sm_count * per_sm;
    const long long need = (n_items + WPC - 1) / WPC;
    if (need < grid) grid = need;
    kern<<<(unsigned)grid, WPC * 32, smem, st>>>(prob, n_items, gate);
    ++g_launches;
    PB_CHECK_CUDA(cudaGetLastError());
    return PB200_OK;
}

template <int RPL, class Prob>
static int launch_lp_rpl(const Prob& prob, long long n_items, int n, cudaStream_t st, const uint32_t* gate = nullptr) {
    if (n_items <= 0) return PB200_OK;
    if (n <= NS && RPL <= 2)
        return launch_persistent(lp_kernel_small<(RPL <= 2 ? RPL : 1), Prob>,
                                 (size_t)WPC * lps_scratch_doubles(RPL) * sizeof(double), prob, n_items, st, gate);
    return launch_persistent(lp_kernel<RPL, Prob>, (size_t)WPC * lp_scratch_doubles(RPL, n) * sizeof(double), prob,
                             n_items, st, gate);
}

template <class Prob>
static int launch_lp(const Prob& prob, long long n_items, int m, int n, cudaStream_t st, const uint32_t* gate = nullptr) {
    if (n < 1 || n > LP_MAX_N) return fail(PB200_EUNSUPPORTED, "number of LP columns must be in 1..32");
    if (m < 1 || m > 128) return fail(PB200_EUNSUPPORTED, "number of LP rows must be in 1..128");
    if (m <= 32) return launch_lp_rpl<1>(prob, n_items, n, st, gate);
    if (m <= 64) return launch_lp_rpl<2>(prob, n_items, n, st, gate);
    return launch_lp_rpl<4>(prob, n_items, n, st, gate);
}

// ------------------------------------------------------------------------
// lane solver launch (LP families that share G: n <= 8 columns, m <= 64 rows)
// ------------------------------------------------------------------------
static int g_lane_enabled = 1;           // pb200_lane_solver(): A/B switch against the warp-per-LP kernels
constexpr int LANE_COUNTERS = 256;
__device__ unsigned long long g_lane_counters[LANE_COUNTERS];
static unsigned long long* g_lane_counters_ptr = nullptr;
static unsigned g_lane_ring = 0;

// n <= 12: always (r02j, one B200: cfg2 2.0x, d = 10 3.1x, cfg4's n = 12 1.7x over the warp-per-LP kernels);
// 13 <= n <= 16 (136-entry factor, two warps per SM): break-even at 500 polytopes, so only for batches
// that fill the GPU several times over
static bool lane_applies(int m, int n, long long P) {
    return g_lane_enabled && n >= 1 && m >= 1 && m <= 64 && (n <= 12 || (n <= 16 && P >= 2000));
}

template <int NS, class Prob>
static int launch_lanes_ns(const Prob& prob, long long P, int m, cudaStream_t st) {
    if (P <= 0) return PB200_OK;
    if (!g_lane_counters_ptr) PB_CHECK_CUDA(cudaGetSymbolAddress((void**)&g_lane_counters_ptr, g_lane_counters));
    // one work counter per launch in flight (launches of different streams may overlap)
    unsigned long long* counter = g_lane_counters_ptr + (__atomic_fetch_add(&g_lane_ring, 1u, __ATOMIC_RELAXED) % LANE_COUNTERS);
    PB_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
    const size_t smem = lane_smem_doubles<NS>(m) * sizeof(double);
    auto kern = lane_kernel<NS, Prob>;
    PB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (!sm_count()) return PB200_ECUDA;
    int per_sm = 0;
    PB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
    if (per_sm < 1) return fail(PB200_EUNSUPPORTED, "lane LP kernel does not fit on an SM");
    long long grid = (long long)g_sm_count * per_sm;      // persistent: one warp per CTA, all resident
    if (P < grid) grid = P;
    kern<<<(unsigned)grid, 32, smem, st>>>(prob, P, m, counter);
    ++g_launches;
    PB_CHECK_CUDA(cudaGetLastError());
    return PB200_OK;
}
// n <= 8: everything in registers; 9 <= n <= 16: lane_solve_wide (padded to 12 or 16 columns)
template <class Prob>
static int launch_lanes(const Prob& prob, long long P, int m, cudaStream_t st) {
    const int n = prob.d;
    if (n <= 8) return launch_lanes_ns<8>(prob, P, m, st);
    if (n <= 12) return launch_lanes_ns<12>(prob, P, m, st);
    return launch_lanes_ns<16>(prob, P, m, st);
}

// independent LPs, one per lane with its own rows (lane_own_kernel): worth it while at least four
// warps fit on an SM
template <int NS>
static bool own_fits(int mr) { return own_smem_doubles<NS>(mr) * sizeof(double) <= 56 * 1024; }
static bool own_applies(int mr, int n) {
    if (!g_lane_enabled || n < 1 || n > 8 || mr < 1) return false;
    return n <= 4 ? own_fits<4>(mr) : own_fits<8>(mr);
}

template <int NS, class Prob>
static int launch_own_ns(const Prob& prob, long long T, int mr, cudaStream_t st) {
    if (!g_lane_counters_ptr) PB_CHECK_CUDA(cudaGetSymbolAddress((void**)&g_lane_counters_ptr, g_lane_counters));
    unsigned long long* counter = g_lane_counters_ptr + (__atomic_fetch_add(&g_lane_ring, 1u, __ATOMIC_RELAXED) % LANE_COUNTERS);
    PB_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
    const size_t smem = own_smem_doubles<NS>(mr) * sizeof(double);
    auto kern = lane_own_kernel<NS, Prob>;
    PB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (!sm_count()) return PB200_ECUDA;
    int per_sm = 0;
    PB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
    if (per_sm < 1) return fail(PB200_EUNSUPPORTED, "lane LP kernel does not fit on an SM");
    long long grid = (long long)g_sm_count * per_sm;
    const long long need = (T + 31) / 32;
    if (need < grid) grid = need;
    kern<<<(unsigned)grid, 32, smem, st>>>(prob, T, mr, counter);
    ++g_launches;
    PB_CHECK_CUDA(cudaGetLastError());
    return PB200_OK;
}
template <class Prob>
static int launch_own(const Prob& prob, long long T, int mr, int n, cudaStream_t st) {
    if (T <= 0) return PB200_OK;
    return n <= 4 ? launch_own_ns<4>(prob, T, mr, st) : launch_own_ns<8>(prob, T, mr, st);
}

// ------------------------------------------------------------------------
// small non-LP kernels of the pipelines (one warp per polytope)
// ------------------------------------------------------------------------
// Polytope.__init__ (polytope.py:128-138)
__global__ void normalize_kernel(const double* __restrict__ A, const double* __restrict__ b,
                                 const int32_t* __restrict__ m_rows, int P, int m, int d, int do_norm,
                                 double* __restrict__ An, double* __restrict__ bn, uint64_t* __restrict__ valid) {
    const int lane = threadIdx.x & 31;
    const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= P) return;
    const int mm = m_rows ? min(max(m_rows[p], 0), m) : m;
    uint64_t mask = 0;
    for (int base = 0; base < m; base += 32) {
        const int i = base + lane;
        bool ok = false;
        if (i < m) {
            const double* row = A + ((size_t)p * m + i) * d;
            double* out = An + ((size_t)p * m + i) * d;
            if (i < mm && !do_norm) {       // rows already normalised by a constructor
                ok = true;
                for (int j = 0; j < d; ++j) out[j] = row[j];
                bn[(size_t)p * m + i] = b[(size_t)p * m + i];
            } else if (i < mm) {
                const double nrm = sqrt(np_sum_squares([&](int j) { return row[j]; }, d));
                ok = nrm > 1e-10;
                const double mult = ok ? __ddiv_rn(1.0, nrm) : 0.0;
                for (int j = 0; j < d; ++j) out[j] = __dmul_rn(row[j], mult);
                bn[(size_t)p * m + i] = ok ? __dmul_rn(b[(size_t)p * m + i], mult) : 0.0;
            } else {
                for (int j = 0; j < d; ++j) out[j] = 0.0;
                bn[(size_t)p * m + i] = 0.0;
            }
        }
        const unsigned bal = __ballot_sync(FULL_MASK, ok);
        mask |= (uint64_t)bal << base;
    }
    if (lane == 0 && valid) valid[p] = mask;
}

// Launch of the constructor normalisation.  Default: the tiled streaming kernel
// (normalize.cuh) with bulk async copies when the spans are 16-byte aligned,
// 64-bit coalesced loads otherwise.  pb200_normalize_variant() forces a variant
// for A/B measurements (tools/normalize_bench.py): -1 = the row-per-lane kernel
// above, 0 / 1 / 2 = 64-bit LDG / 128-bit LDG / bulk copies.
static int g_norm_variant = -2;

static int launch_normalize(const double* A, const double* b, const int32_t* m_rows, int P, int m, int d, int do_norm,
                            double* An, double* bn, uint64_t* valid, cudaStream_t st) {
    if (g_norm_variant == -1) {
        normalize_kernel<<<blocks_for((long long)P * 32, 256), 256, 0, st>>>(A, b, m_rows, P, m, d, do_norm, An, bn, valid);
        ++g_launches;
        PB_CHECK_CUDA(cudaGetLastError());
        return PB200_OK;
    }
    const bool aligned = (((uintptr_t)A | (uintptr_t)b | (uintptr_t)An | (uintptr_t)bn) & 15u) == 0 && (m & 1) == 0;
    int mode = g_norm_variant == -2 ? NORM_BULK : g_norm_variant;
    if (!aligned) mode = NORM_LDG64;
    int nw = 8;                             // polytopes (warps) per CTA: two CTAs per SM must fit
    while (nw > 1 && normalize_smem_doubles(nw, m, d) * sizeof(double) > 110 * 1024) nw >>= 1;
    const size_t smem = normalize_smem_doubles(nw, m, d) * sizeof(double);
    auto kern = mode == NORM_BULK ? normalize_tile_kernel<NORM_BULK>
                : mode == NORM_LDG128 ? normalize_tile_kernel<NORM_LDG128> : normalize_tile_kernel<NORM_LDG64>;
    if (smem > 48 * 1024) PB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t div_magic = (1u << 20) / (uint32_t)d + 1u;
    kern<<<blocks_for(P, nw), 32 * nw, smem, st>>>(A, b, m_rows, P, m, d, do_norm, div_magic, An, bn, valid);
    ++g_launches;
    PB_CHECK_CUDA(cudaGetLastError());
    return PB200_OK;
}

// reduce(): is_fulldim verdict, b == inf drop and duplicate-direction filter
// (polytope.py:1081-1116).  One warp per polytope.
__global__ void prefilter_kernel(const double* __restrict__ An, const double* __restrict__ bn,
                                 const uint64_t* __restrict__ valid, const double* __restrict__ r,
                                 const int8_t* __restrict__ cheb_status, int P, int m, int d, double abs_tol,
                                 uint64_t* __restrict__ rows1, uint32_t* __restrict__ flags) {
    extern __shared__ double sh[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
    if (p >= P) return;
    // a_normed [m][ld], ld odd: lane j reads row j in the pair loop, and an even stride of d doubles put the 32 rows
    // on two banks (16-way conflict: this kernel took 64 us on cfg2 with d = 8)
    const int ld = d | 1;
    double* an = sh + (size_t)wib * (m * ld + 2 * m);
    double* bs = an + m * ld;                          // b * a_norm
    const double* Ap = An + (size_t)p * m * d;
    const double* bp = bn + (size_t)p * m;
    // ABS_TOL of is_fulldim's default argument, polytope.py:962 (not reduce's abs_tol)
    const bool fulldim = cheb_status[p] == ST_OPTIMAL && r[p] > 1e-7;
    if (!fulldim) {
        if (lane == 0) { rows1[p] = 0; flags[p] = PB200_F_EMPTY; }
        return;
    }
    uint64_t alive = valid[p] & low_bits(m);
    for (int base = 0; base < m; base += 32) {
        const int i = base + lane;
        bool fin = false;
        if (i < m && ((alive >> i) & 1ull)) {
            const double bi = bp[i];
            fin = bi != __longlong_as_double(0x7ff0000000000000ll);
            const double* row = Ap + (size_t)i * d;
            const double an_i = __ddiv_rn(1.0, sqrt(np_sum_squares([&](int j) { return row[j]; }, d)));
            for (int j = 0; j < d; ++j) an[i * ld + j] = __dmul_rn(row[j], an_i);
            bs[i] = __dmul_rn(bi, an_i);
        }
        const unsigned bal = __ballot_sync(FULL_MASK, fin);
        alive = (alive & ~((uint64_t)0xffffffffu << base)) | ((uint64_t)bal << base);
    }
    __syncwarp();
    // all pairs i < j of alive rows; the removed set is a union, order-free
    unsigned rem_lo = 0, rem_hi = 0;
    for (int i = 0; i < m; ++i) {
        if (!((alive >> i) & 1ull)) continue;
        for (int j = i + 1 + lane; j < m; j += 32) {
            if (!((alive >> j) & 1ull)) continue;
            double dot = 0.0;
            for (int q = 0; q < d; ++q) dot = fma(an[i * ld + q], an[j * ld + q], dot);
            if (dot > 1.0 - abs_tol) {
                const int rm = bs[i] < bs[j] ? j : i;
                if (rm < 32) rem_lo |= 1u << rm; else rem_hi |= 1u << (rm - 32);
            }
        }
    }
    rem_lo = __reduce_or_sync(FULL_MASK, rem_lo);
    rem_hi = __reduce_or_sync(FULL_MASK, rem_hi);
    const uint64_t keep = alive & ~(((uint64_t)rem_hi << 32) | rem_lo);
    if (lane == 0) {
        rows1[p] = keep;
        flags[p] = 0;
    }
}

// pipeline control bits kept in the upper half of flags[] while reduce runs
constexpr uint32_t CTL_NEED_BBOX = 1u << 16;
constexpr uint32_t CTL_ROW_LOOP = 1u << 17;
constexpr uint32_t CTL_RETRY_BBOX = 1u << 18;   // a lane-solver LP ended without a verdict: repeat on the warp kernel
constexpr uint32_t CTL_RETRY_ROWS = 1u << 19;
constexpr uint32_t CTL_MASK = 0xffff0000u;

// decide early exit / bbox need after the duplicate filter (polytope.py:1113-1118)
__global__ void plan_kernel(const uint64_t* __restrict__ rows1, uint32_t* __restrict__ flags, int P, int d,
                            int early_exit) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    uint32_t f = flags[p];
    if (f & PB200_F_EMPTY) return;
    const int neq = __popcll(rows1[p]);
    if (early_exit && neq <= d + 1) { flags[p] = f; return; }    // nonEmptyBounded early exit, minrep stays False
    if (neq > 3 * d) f |= CTL_NEED_BBOX | PB200_F_BBOX;
    else f |= CTL_ROW_LOOP;
    flags[p] = f;
}

// bounding-box candidate filter (polytope.py:1119-1138). One warp per polytope.
__global__ void candidate_kernel(const double* __restrict__ An, const double* __restrict__ bn,
                                 const uint64_t* __restrict__ rows1, const double* __restrict__ bblo,
                                 const double* __restrict__ bbhi, const int8_t* __restrict__ bbstatus, int P, int m, int d,
                                 int early_exit, uint64_t* __restrict__ rows2, uint32_t* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= P) return;
    const uint32_t f = flags[p];
    const uint64_t mask1 = rows1[p];
    if (!(f & CTL_NEED_BBOX)) {
        if (lane == 0) rows2[p] = mask1;
        return;
    }
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    uint64_t cand = 0;
    bool lpfail = false;
    for (int base = 0; base < m; base += 32) {
        const int i = base + lane;
        bool c = false;
        if (i < m && ((mask1 >> i) & 1ull)) {
            const double* row = An + ((size_t)p * m + i) * d;
            double t1 = 0.0, t2 = 0.0;
            for (int j = 0; j < d; ++j) {
                const int8_t sl = bbstatus[(size_t)p * 2 * d + j], su = bbstatus[(size_t)p * 2 * d + d + j];
                double lo = bblo[(size_t)p * d + j], hi = bbhi[(size_t)p * d + j];
                if (sl == ST_UNBOUNDED) lo = -inf; else if (sl == ST_INFEASIBLE) lo = 0.0; else if (sl != ST_OPTIMAL) lpfail = true;
                if (su == ST_UNBOUNDED) hi = inf; else if (su == ST_INFEASIBLE) hi = lo; else if (su != ST_OPTIMAL) lpfail = true;
                const double a = row[j];
                const double ap = a > 0.0 ? a : __dmul_rn(0.0, a);
                t1 = __dadd_rn(t1, __dmul_rn(ap, __dadd_rn(hi, -lo)));
                t2 = __dadd_rn(t2, __dmul_rn(a, lo));
            }
            const double v = t1 - (bn[(size_t)p * m + i] - t2);
            c = !(v < -1e-4);
        }
        const unsigned bal = __ballot_sync(FULL_MASK, c);
        cand |= (uint64_t)bal << base;
    }
    lpfail = __any_sync(FULL_MASK, lpfail);
    if (lane == 0) {
        const uint64_t mask2 = mask1 & cand;
        rows2[p] = mask2;
        uint32_t g = f & ~CTL_NEED_BBOX;
        if (lpfail) g |= PB200_F_LPFAIL;
        if (!early_exit || __popcll(mask2) > d + 1) g |= CTL_ROW_LOOP;
        flags[p] = g;
    }
}

// assemble keep masks, drifted b, LP counts and public flags
__global__ void finalize_kernel(const double* __restrict__ bn, const uint64_t* __restrict__ rows2,
                                const unsigned long long* __restrict__ keep_lp, uint32_t* __restrict__ flags,
                                int P, int m, int d, uint64_t* __restrict__ keep, double* __restrict__ b_out,
                                int32_t* __restrict__ n_lp) {
    const int lane = threadIdx.x & 31;
    const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= P) return;
    const uint32_t f = flags[p];
    const uint64_t mask2 = rows2[p];
    const bool looped = f & CTL_ROW_LOOP;
    for (int i = lane; i < m; i += 32) {
        double v = bn[(size_t)p * m + i];
        if (looped && ((mask2 >> i) & 1ull)) v = __dadd_rn(__dadd_rn(v, 0.1), -0.1);
        b_out[(size_t)p * m + i] = v;
    }
    if (lane == 0) {
        uint32_t g = f & ~CTL_MASK;
        uint64_t k = 0;
        int lps = 1;
        if (!(f & PB200_F_EMPTY)) {
            if (f & PB200_F_BBOX) lps += 2 * d;
            if (looped) { k = keep_lp[p]; lps += __popcll(mask2); g |= PB200_F_MINREP; }
            else k = mask2;
        }
        keep[p] = k;
        flags[p] = g;
        if (n_lp) n_lp[p] = lps;
    }
}

// bounding_box status conventions (polytope.py:1372-1402).  An optimum that
// sits (to 1e-11) on an axis-aligned facet +-e_j x <= b_i is returned as that
// facet's b_i exactly: the reference's simplex returns such vertices without
// rounding, and callers floor()/ceil() the bounds of boxes
// (enumerate_integral_points, polytope.py:2352-2358).
__global__ void bbox_resolve_kernel(const int8_t* __restrict__ status, const double* __restrict__ A,
                                    const double* __restrict__ b, const int32_t* __restrict__ m_rows, int P, int m, int d,
                                    double* lo, double* hi) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)P * d) return;
    const long long p = t / d;
    const int j = (int)(t - p * d);
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    const int8_t sl = status[p * 2 * d + j], su = status[p * 2 * d + d + j];
    double l = lo[t], u = hi[t];
    const int mm = m_rows ? min(max(m_rows[p], 0), m) : m;
    for (int i = 0; i < mm; ++i) {
        const double* row = A + ((size_t)p * m + i) * d;
        const double a = row[j];
        if (a != 1.0 && a != -1.0) continue;
        bool axis = true;
        for (int k = 0; k < d; ++k) axis = axis && (k == j || row[k] == 0.0);
        if (!axis) continue;
        const double bi = b[(size_t)p * m + i];
        if (a == 1.0 && su == ST_OPTIMAL && fabs(u - bi) <= 1e-11 * fmax(1.0, fabs(bi))) u = bi;
        if (a == -1.0 && sl == ST_OPTIMAL && fabs(l + bi) <= 1e-11 * fmax(1.0, fabs(bi))) l = -bi;
    }
    if (sl == ST_UNBOUNDED) l = -inf; else if (sl == ST_INFEASIBLE) l = 0.0; else if (sl != ST_OPTIMAL) l = nan;
    if (su == ST_UNBOUNDED) u = inf; else if (su == ST_INFEASIBLE) u = l; else if (su != ST_OPTIMAL) u = nan;
    lo[t] = l;
    hi[t] = u;
}

struct ReduceWorkspace {
    double *An, *bn, *bblo, *bbhi;
    uint64_t *valid, *rows1, *rows2;
    unsigned long long* keep_lp;
    int8_t *cheb_status, *bbstatus;
    uint32_t* retry_any;     // [2]: a bounding-box / row LP of the lane kernels ended without a verdict
    size_t bytes;
};
static ReduceWorkspace carve_reduce(void* base, int P, int m, int d) {
    ReduceWorkspace w;
    char* p = (char*)base;
    auto take = [&](size_t n) { char* q = p; p += (n + 255) & ~(size_t)255; return q; };
    w.An = (double*)take(sizeof(double) * P * m * d);
    w.bn = (double*)take(sizeof(double) * P * m);
    w.bblo = (double*)take(sizeof(double) * P * d);
    w.bbhi = (double*)take(sizeof(double) * P * d);
    w.valid = (uint64_t*)take(sizeof(uint64_t) * P);
    w.rows1 = (uint64_t*)take(sizeof(uint64_t) * P);
    w.rows2 = (uint64_t*)take(sizeof(uint64_t) * P);
    w.keep_lp = (unsigned long long*)take(sizeof(uint64_t) * P);
    w.cheb_status = (int8_t*)take(P);
    w.bbstatus = (int8_t*)take((size_t)P * 2 * d);
    w.retry_any = (uint32_t*)take(2 * sizeof(uint32_t));
    w.bytes = (size_t)(p - (char*)base);
    return w;
}

}  // namespace pb200

using namespace pb200;

extern "C" {

const char* pb200_version(void) { return "polytope_b200 0.1 (sm_100a)"; }
const char* pb200_last_error(void) { return g_err; }
long long pb200_launch_count(void) { return g_launches; }

int pb200_lp_batch(const double* G, const double* h, const double* c, const int32_t* m_rows, int B, int m, int n,
                   double* x, double* fun, int8_t* status, int32_t* iters, void* stream) {
    if (B == 0) return PB200_OK;
    if (B < 0 || !G || !h || !c || !x || !fun || !status) return fail(PB200_EINVAL, "pb200_lp_batch: null pointer or negative batch");
    GenericLP prob{G, h, c, m_rows, m, n, x, fun, status, iters};
    return launch_lp(prob, B, m, n, (cudaStream_t)stream);
}

int pb200_normalize_batch(const double* A, const double* b, const int32_t* m_rows, int P, int m, int d, double* An,
                          double* bn, uint64_t* valid, void* stream) {
    if (P == 0) return PB200_OK;
    if (P < 0 || !A || !b || !An || !bn) return fail(PB200_EINVAL, "pb200_normalize_batch: null pointer");
    if (m < 1 || m > 64 || d < 1 || d > 128) return fail(PB200_EUNSUPPORTED, "normalize: need 1<=m<=64, 1<=d<=128");
    if (P == 0) return PB200_OK;
    return launch_normalize(A, b, m_rows, P, m, d, 1, An, bn, valid, (cudaStream_t)stream);
}

void pb200_lane_solver(int on) { g_lane_enabled = on != 0; }

void pb200_normalize_variant(int variant) { g_norm_variant = variant < -2 || variant > 2 ? -2 : variant; }

int pb200_cheby_batch(const double* A, const double* b, const int32_t* m_rows, const uint64_t* rows, int P, int m,
                      int d, double* r, double* xc, int8_t* status, void* stream) {
    if (P == 0) return PB200_OK;
    if (P < 0 || !A || !b || !r || !xc || !status) return fail(PB200_EINVAL, "pb200_cheby_batch: null pointer");
    if (rows && m > 64) return fail(PB200_EUNSUPPORTED, "row masks need m <= 64");
    if (m <= 64 && own_applies(m, d + 1)) {
        ChebyOwn prob{A, b, m_rows, rows, m, d, r, xc, status, nullptr};
        return launch_own(prob, P, m, d + 1, (cudaStream_t)stream);
    }
    ChebyLP prob{A, b, m_rows, rows, nullptr, 0, m, d, r, xc, status, nullptr};
    return launch_lp(prob, P, m, d + 1, (cudaStream_t)stream);
}

int pb200_bbox_batch(const double* A, const double* b, const int32_t* m_rows, int P, int m, int d, double* lo,
                     double* hi, int8_t* status, void* stream) {
    if (P == 0) return PB200_OK;
    if (P < 0 || !A || !b || !lo || !hi || !status) return fail(PB200_EINVAL, "pb200_bbox_batch: null pointer");
    if (d < 1 || d > LP_MAX_N) return fail(PB200_EUNSUPPORTED, "bbox: need 1 <= d <= 32");
    if (P == 0) return PB200_OK;
    // the LP kernel writes the raw optimised coordinates into lo / hi, the
    // resolve kernel then applies the reference's status conventions in place
    int rc;
    if (lane_applies(m, d, P)) {
        BboxLanes prob{A, b, m_rows, nullptr, nullptr, 0, 0, nullptr, m, d, 0, lo, hi, status, nullptr};
        rc = launch_lanes(prob, P, m, (cudaStream_t)stream);
    } else {
        BboxLP prob{A, b, m_rows, nullptr, nullptr, 0, m, d, 0, lo, hi, status, nullptr};
        rc = launch_lp(prob, (long long)P * 2 * d, m, d, (cudaStream_t)stream);
    }
    if (rc) return rc;
    bbox_resolve_kernel<<<blocks_for((long long)P * d, 256), 256, 0, (cudaStream_t)stream>>>(status, A, b, m_rows, P, m, d, lo, hi);
    ++g_launches;
    PB_CHECK_CUDA(cudaGetLastError());
    return PB200_OK;
}

size_t pb200_reduce_workspace_bytes(int P, int m, int d) {
    if (P < 0 || m < 1 || d < 1) return 0;
    return carve_reduce(nullptr, P, m, d).bytes;
}

int pb200_reduce_batch(const double* A, const double* b, const int32_t* m_rows, int P, int m, int d, double abs_tol,
                       int normalize, uint64_t* keep, uint32_t* flags, double* r, double* xc, double* b_out, double* A_out,
                       int32_t* n_lp, int32_t* lp_iters, void* workspace, size_t workspace_bytes, void* stream) {
    if (P == 0) return PB200_OK;
    if (P < 0 || !A || !b || !keep || !flags || !r || !xc || !b_out || !workspace)
        return fail(PB200_EINVAL, "pb200_reduce_batch: null pointer");
    if (m < 1 || m > 64) return fail(PB200_EUNSUPPORTED, "reduce: need 1 <= m <= 64 rows (row sets are 64-bit masks)");
    if (d < 1 || d + 1 > LP_MAX_N) return fail(PB200_EUNSUPPORTED, "reduce: need 1 <= d <= 31");
    if (P == 0) return PB200_OK;
    ReduceWorkspace ws = carve_reduce(workspace, P, m, d);
    if (ws.bytes > workspace_bytes) return fail(PB200_EWORKSPACE, "pb200_reduce_batch: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    double* An = A_out ? A_out : ws.An;
    const int early_exit = (normalize & PB200_REDUCE_NO_EARLY_EXIT) ? 0 : 1;
    normalize &= PB200_REDUCE_NORMALIZE;
    int rc;
    if (lp_iters) PB_CHECK_CUDA(cudaMemsetAsync(lp_iters, 0, sizeof(int32_t) * P, st));
    PB_CHECK_CUDA(cudaMemsetAsync(ws.retry_any, 0, 2 * sizeof(uint32_t), st));
    stage_mark(0, st);
    // 1. constructor normalisation
    rc = launch_normalize(A, b, m_rows, P, m, d, normalize ? 1 : 0, An, ws.bn, ws.valid, st);
    if (rc) return rc;
    stage_mark(1, st);
    // 2. is_fulldim: one Chebyshev LP per polytope
    if (own_applies(m, d + 1)) {
        ChebyOwn cheb{An, ws.bn, nullptr, ws.valid, m, d, r, xc, ws.cheb_status, lp_iters};
        if ((rc = launch_own(cheb, P, m, d + 1, st))) return rc;
    } else {
        ChebyLP cheb{An, ws.bn, nullptr, ws.valid, nullptr, 0, m, d, r, xc, ws.cheb_status, lp_iters};
        if ((rc = launch_lp(cheb, P, m, d + 1, st))) return rc;
    }
    stage_mark(2, st);
    // 3. b == inf drop + duplicate-direction filter, then plan
    {
        const int wpb = 4;
        const size_t sh = (size_t)wpb * (m * (d | 1) + 2 * m) * sizeof(double);
        if (sh > 48 * 1024)
            PB_CHECK_CUDA(cudaFuncSetAttribute(prefilter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        prefilter_kernel<<<blocks_for(P, wpb), wpb * 32, sh, st>>>(An, ws.bn, ws.valid, r, ws.cheb_status, P, m, d,
                                                                 abs_tol, ws.rows1, flags);
        ++g_launches;
        PB_CHECK_CUDA(cudaGetLastError());
        plan_kernel<<<blocks_for(P, 256), 256, 0, st>>>(ws.rows1, flags, P, d, early_exit);
        ++g_launches;
        PB_CHECK_CUDA(cudaGetLastError());
    }
    stage_mark(3, st);
    // 4. bounding box of Polytope(A_arr, b_arr) where neq > 3 nx
    if (lane_applies(m, d, P)) {
        BboxLanes bb{An, ws.bn, nullptr, ws.rows1, flags, CTL_NEED_BBOX, CTL_RETRY_BBOX, ws.retry_any, m, d, 1, ws.bblo, ws.bbhi, ws.bbstatus, lp_iters};
        if ((rc = launch_lanes(bb, P, m, st))) return rc;
        // safety net: polytopes with an LP the lane solver could not finish (none on the BASELINE
        // workloads) go through the warp-per-LP kernel, whose arithmetic differs
        BboxLP again{An, ws.bn, nullptr, ws.rows1, flags, CTL_RETRY_BBOX, m, d, 1, ws.bblo, ws.bbhi, ws.bbstatus, lp_iters};
        if ((rc = launch_lp(again, (long long)P * 2 * d, m, d, st, ws.retry_any))) return rc;
    } else {
        BboxLP bb{An, ws.bn, nullptr, ws.rows1, flags, CTL_NEED_BBOX, m, d, 1, ws.bblo, ws.bbhi, ws.bbstatus, lp_iters};
        if ((rc = launch_lp(bb, (long long)P * 2 * d, m, d, st))) return rc;
    }
    stage_mark(4, st);
    // 5. candidate filter
    candidate_kernel<<<blocks_for((long long)P * 32, 256), 256, 0, st>>>(An, ws.bn, ws.rows1, ws.bblo, ws.bbhi, ws.bbstatus,
                                                                       P, m, d, early_exit, ws.rows2, flags);
    ++g_launches;
    PB_CHECK_CUDA(cudaGetLastError());
    stage_mark(5, st);
    // 6. one LP per surviving row
    PB_CHECK_CUDA(cudaMemsetAsync(ws.keep_lp, 0, sizeof(uint64_t) * P, st));
    if (lane_applies(m, d, P)) {
        RowLanes row{An, ws.bn, ws.rows2, flags, CTL_ROW_LOOP, CTL_RETRY_ROWS, ws.retry_any + 1, m, d, abs_tol, ws.keep_lp, lp_iters};
        if ((rc = launch_lanes(row, P, m, st))) return rc;
        RowLP again{An, ws.bn, ws.rows2, flags, CTL_RETRY_ROWS, m, d, abs_tol, ws.keep_lp, lp_iters};
        if ((rc = launch_lp(again, (long long)P * m, m, d, st, ws.retry_any + 1))) return rc;
    } else {
        RowLP row{An, ws.bn, ws.rows2, flags, CTL_ROW_LOOP, m, d, abs_tol, ws.keep_lp, lp_iters};
        if ((rc = launch_lp(row, (long long)P * m, m, d, st))) return rc;
    }
    stage_mark(6, st);
    // 7. results
    finalize_kernel<<<blocks_for((long long)P * 32, 256), 256, 0, st>>>(ws.bn, ws.rows2, ws.keep_lp, flags, P, m, d, keep,
                                                                      b_out, n_lp);
    ++g_launches;
    PB_CHECK_CUDA(cudaGetLastError());
    stage_mark(7, st);
    return PB200_OK;
}

void pb200_profile_enable(int on) { g_profile = on != 0; g_ev_valid = false; }

int pb200_profile_read(float* stage_ms, int n) {
    if (!stage_ms || n < N_STAGES) return fail(PB200_EINVAL, "pb200_profile_read: need room for 7 stages");
    if (!g_ev_valid) return fail(PB200_EINVAL, "pb200_profile_read: no profiled pb200_reduce_batch call yet");
    PB_CHECK_CUDA(cudaEventSynchronize(g_ev[N_STAGES]));
    for (int k = 0; k < N_STAGES; ++k) PB_CHECK_CUDA(cudaEventElapsedTime(stage_ms + k, g_ev[k], g_ev[k + 1]));
    return PB200_OK;
}

int pb200_adjacent_pairs(const double* A, const double* b, int ncell, int mc, int d, const int32_t* pair_i,
                         const int32_t* pair_j, long long T, double abs_tol, uint8_t* adjacent, double* radius,
                         int8_t* status, void* stream) {
    if (T == 0) return PB200_OK;
    if (!A || !b || !adjacent || T < 0 || ncell < 0) return fail(PB200_EINVAL, "pb200_adjacent_pairs: bad argument");
    if ((pair_i == nullptr) != (pair_j == nullptr)) return fail(PB200_EINVAL, "pair_i and pair_j must both be given or both NULL");
    if (!pair_i && T != (long long)ncell * (ncell - 1) / 2)
        return fail(PB200_EINVAL, "implicit pair enumeration needs T == ncell*(ncell-1)/2");
    if (2 * mc > 64) return fail(PB200_EUNSUPPORTED, "adjacent: need 2*mc <= 64 rows");
    if (own_applies(2 * mc, d + 1)) {
        AdjacentOwn prob{A, b, ncell, mc, d, pair_i, pair_j, 0, 0, abs_tol, adjacent, radius, status};
        return launch_own(prob, T, 2 * mc, d + 1, (cudaStream_t)stream);
    }
    AdjacentLP prob{A, b, ncell, mc, d, pair_i, pair_j, abs_tol, adjacent, radius, status};
    return launch_lp(prob, T, 2 * mc, d + 1, (cudaStream_t)stream);
}

int pb200_adjacent_range(const double* A, const double* b, int ncell, int mc, int d, int order, long long t_begin,
                         long long T, double abs_tol, uint8_t* adjacent, double* radius, int8_t* status, void* stream) {
    if (T == 0) return PB200_OK;
    if (!A || !b || !adjacent || T < 0 || t_begin < 0 || ncell < 2) return fail(PB200_EINVAL, "pb200_adjacent_range: bad argument");
    if (order != 0 && order != 1) return fail(PB200_EINVAL, "pb200_adjacent_range: order must be 0 (j < i) or 1 (all i != j)");
    const long long total = order == 0 ? (long long)ncell * (ncell - 1) / 2 : (long long)ncell * (ncell - 1);
    if (t_begin + T > total) return fail(PB200_EINVAL, "pb200_adjacent_range: pair range exceeds the enumeration");
    if (2 * mc > 64) return fail(PB200_EUNSUPPORTED, "adjacent: need 2*mc <= 64 rows");
    if (!own_applies(2 * mc, d + 1))
        return fail(PB200_EUNSUPPORTED, "pb200_adjacent_range: cells too large for the lane kernel; pass explicit pair lists "
                                        "to pb200_adjacent_pairs");
    AdjacentOwn prob{A, b, ncell, mc, d, nullptr, nullptr, order, t_begin, abs_tol, adjacent, radius, status};
    return launch_own(prob, T, 2 * mc, d + 1, (cudaStream_t)stream);
}

}  // extern "C"

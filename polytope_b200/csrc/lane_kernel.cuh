// Device side of the lane LP solver (lp_lane.cuh): a warp packs up to 32 LPs of up to
// LANE_NG polytopes into its lanes, stages each polytope's rows once into shared memory
// (row-major, read by broadcast) and lets every lane solve its own LP.  Polytopes are handed
// out by a global counter, so warps stay busy whatever the per-polytope LP counts are.
//
// Problems:
//   RowLanes   reduce()'s row loop        polytope/polytope.py:1142-1160
//   BboxLanes  bounding_box's 2d LPs      polytope/polytope.py:1362-1411
#pragma once
#include "common.cuh"
#include "staging.cuh"
#include "normalize.cuh"      // bulk-copy / mbarrier wrappers
#include "lp_lane_wide.cuh"

namespace pb200 {

// The kernels are instantiated for NS = 8 padded columns (lane_solve, everything in registers)
// and NS = 12 / 16 (lane_solve_wide, the factor in a lane-interleaved shared-memory array).
constexpr int LANE_NS = 8;
template <int NS> PBL_CX int lane_ng() { return NS <= 8 ? 3 : 2; }     // polytopes a warp holds per round
template <int NS> PBL_CX int lane_nt() { return NS * (NS + 1) / 2; }
#ifndef PB200_LANE_MINB
#define PB200_LANE_MINB 8
#endif

template <int NS>
__host__ __device__ inline int lane_slot_doubles(int m) { return m * NS + 4; }   // +4: slots land in different banks
// G[NG][slot] | hA[NG][m] | hB[NG][m] | sz[2 m][32] | raw[m NS + m] (landing zone of the bulk copies) | mbarrier
// | (NS > 8) L[NT][32]
template <int NS>
__host__ __device__ inline size_t lane_smem_doubles(int m) {
    return (size_t)lane_ng<NS>() * lane_slot_doubles<NS>(m) + 2 * lane_ng<NS>() * m + 64 * (size_t)m + (size_t)m * NS + m + 2 +
           (NS > 8 ? (size_t)lane_nt<NS>() * 32 : 0);
}

// per-warp staging context: where the TMA engine lands a polytope's (A, b) block
struct LaneStage {
    double* raw;
    uint64_t* bar;
    uint32_t parity;
};

// what a lane sees of its LP
template <int NS>
struct LaneData {
    const double* G;     // [rows][NS] row-major, shared by the lanes of the same polytope
    const double* hA;    // right-hand side
    const double* hB;    // right-hand side after the reference's +0.1 / -0.1 round trip (rows before k)
    double* sz;          // this lane's (s_i, z_i): sz[(2 i) * 32], sz[(2 i + 1) * 32]
    double* Lm;          // NS > 8: this lane's packed factor, entry e at Lm[e * 32]
    int m, k;            // rows; row whose h carries +0.1 (-1: none)
    int cj;              // objective: cs * e_cj, or -G[k] when cj < 0
    double cs;
    __device__ __forceinline__ int rows() const { return m; }
    __device__ __forceinline__ void row(int i, double (&g)[NS]) const {
        const double2* p = reinterpret_cast<const double2*>(G + i * NS);
#pragma unroll
        for (int j = 0; j < NS / 2; ++j) { const double2 v = p[j]; g[2 * j] = v.x; g[2 * j + 1] = v.y; }
    }
    __device__ __forceinline__ double h(int i) const {
        const double* src = i < k ? hB : hA;          // rows before k carry the +0.1 / -0.1 round trip
        const double a = src[i];
        return i == k ? __dadd_rn(a, 0.1) : a;
    }
    __device__ __forceinline__ double c(int j) const { return cj < 0 ? -G[k * NS + j] : (j == cj ? cs : 0.0); }
    __device__ __forceinline__ double& s(int i) { return sz[(2 * i) * 32]; }
    __device__ __forceinline__ double& z(int i) { return sz[(2 * i + 1) * 32]; }
    __device__ __forceinline__ double& L(int e) { return Lm[e * 32]; }
};

// rows selected by `mask` (ascending) of a row-major [m x d] matrix -> row-major [cnt][NS], zero padded.
// The polytope's (A, b) block is one contiguous span each: when the spans are 16-byte aligned they
// are fetched by two 1-D bulk async copies (cp.async.bulk, the TMA engine; SASS UBLKCP) that
// complete on the warp's mbarrier, and the rows are picked / padded out of shared memory;
// otherwise by coalesced loads.
template <int NS>
__device__ __forceinline__ int lane_stage_masked(const double* __restrict__ Ap, const double* __restrict__ bp, int m, int d,
                                                 uint64_t mask, double* G, double* hA, int lane, LaneStage& sg) {
    const int cnt = __popcll(mask);
    const int total = m * d;
    const bool bulk = (((uintptr_t)Ap | (uintptr_t)bp) & 15u) == 0 && (m & 1) == 0;
    if (bulk) {
        if (lane == 0) {
            fence_smem_to_async_proxy();       // earlier reads of the landing zone are done (syncwarp of the previous use)
            mbar_expect_tx(sg.bar, (uint32_t)((total + m) * sizeof(double)));
            bulk_load(sg.raw, Ap, (uint32_t)(total * sizeof(double)), sg.bar);
            bulk_load(sg.raw + total, bp, (uint32_t)(m * sizeof(double)), sg.bar);
        }
    }
    for (int e = lane; e < cnt * NS; e += 32) G[e] = 0.0;
    __syncwarp();
    if (bulk) {
        mbar_wait(sg.bar, sg.parity);
        sg.parity ^= 1u;
        for (int e = lane; e < total; e += 32) {
            const int i = e / d, j = e - i * d;
            if ((mask >> i) & 1ull) G[__popcll(mask & ((1ull << i) - 1ull)) * NS + j] = sg.raw[e];
        }
        for (int i = lane; i < m; i += 32)
            if ((mask >> i) & 1ull) hA[__popcll(mask & ((1ull << i) - 1ull))] = sg.raw[total + i];
    } else {
        for (int e = lane; e < total; e += 32) {
            const int i = e / d, j = e - i * d;
            if ((mask >> i) & 1ull) G[__popcll(mask & ((1ull << i) - 1ull)) * NS + j] = __ldg(Ap + e);
        }
        for (int i = lane; i < m; i += 32)
            if ((mask >> i) & 1ull) hA[__popcll(mask & ((1ull << i) - 1ull))] = __ldg(bp + i);
    }
    __syncwarp();
    return cnt;
}

// Polytope.__init__ normalisation of the staged rows (polytope.py:128-138), numpy's summation order
template <int NS>
__device__ __forceinline__ void lane_renormalize(double* G, double* hA, int cnt, int d, int lane) {
    for (int i = lane; i < cnt; i += 32) {
        double* row = G + i * NS;
        const double nrm = sqrt(np_sum_squares([&](int j) { return row[j]; }, d));
        if (nrm > 1e-10) {
            const double mult = __ddiv_rn(1.0, nrm);
            for (int j = 0; j < d; ++j) row[j] = __dmul_rn(row[j], mult);
            hA[i] = __dmul_rn(hA[i], mult);
        } else {
            // the constructor drops the row (:130): staged as 0'x <= 1, which never binds
            for (int j = 0; j < d; ++j) row[j] = 0.0;
            hA[i] = 1.0;
        }
    }
    __syncwarp();
}

struct RowLanes {
    const double *A, *b;        // constructor-normalised
    const uint64_t* rows;       // surviving rows (after duplicate / bbox filters)
    uint32_t* flags;
    uint32_t run_mask;          // run only when (flags & run_mask)
    uint32_t retry_bit;         // set on flags[p] when an LP ends without a verdict (status 1 / 4): the
                                // polytope's row LPs are then repeated on the warp-per-LP kernel
    uint32_t* retry_any;        // nullable: set to 1 with it (the retry launch returns at once while it is 0)
    int m, d;
    double abs_tol;
    unsigned long long* keep;   // OR-accumulated
    int32_t* lp_iters;          // nullable: += interior-point iterations
    struct Tune {                          // lp_lane.cuh: the row-LP kernel is instruction-fetch bound
        static constexpr int kUnroll = 1;
#ifndef PB200_ROW_EARLY
#define PB200_ROW_EARLY 1e-1
#endif
#ifndef PB200_ROW_NEXT
#define PB200_ROW_NEXT 1e-1
#endif
        static constexpr double kEarly = PB200_ROW_EARLY, kNext = PB200_ROW_NEXT;
    };
    __device__ int n() const { return d; }
    __device__ int count(long long p) const {
        if (!(flags[p] & run_mask)) return 0;
        return __popcll(rows[p] & low_bits(m));
    }
    template <int NS>
    __device__ int stage(long long p, double* G, double* hA, double* hB, int lane, LaneStage& sg) const {
        const int cnt = lane_stage_masked<NS>(A + (size_t)p * m * d, b + (size_t)p * m, m, d, rows[p] & low_bits(m), G, hA, lane, sg);
        for (int i = lane; i < cnt; i += 32) hB[i] = __dadd_rn(__dadd_rn(hA[i], 0.1), -0.1);
        __syncwarp();
        return cnt;
    }
    template <int NS>
    __device__ void setup(LaneData<NS>& dat, int k) const { dat.k = k; dat.cj = -1; dat.cs = 0.0; }
    template <int NS>
    __device__ void store(long long p, int k, const lane::Result<NS>& res) const {
        const uint64_t mask = rows[p] & low_bits(m);
        const int orig = nth_set_bit(mask, k);
        bool kept = false;
        if (res.status == lane::OPTIMAL) {
            const double hk = __dadd_rn(__dadd_rn(b[(size_t)p * m + orig], 0.1), -0.1);
            kept = (-res.fun - hk) > abs_tol;
        } else if (res.status == lane::UNBOUNDED) {
            kept = true;
        } else if (res.status != lane::INFEASIBLE) {
            atomicOr(flags + p, retry_bit);
            if (retry_any) *retry_any = 1u;
        }
        if (kept) atomicOr(keep + p, 1ull << orig);
        if (lp_iters) atomicAdd(lp_iters + p, res.iters);
    }
};

struct BboxLanes {
    const double *A, *b;
    const int32_t* m_rows;
    const uint64_t* rows;        // nullable row masks
    uint32_t* need_flags;        // nullable: run only when (flags & need_mask)
    uint32_t need_mask;
    uint32_t retry_bit;          // with need_flags: set when an LP ends with status 1 / 4 (see RowLanes)
    uint32_t* retry_any;         // nullable, as RowLanes
    int m, d, renorm;
    double *val_lo, *val_hi;     // [P][d] each: optimised coordinate of the lower / upper LP
    int8_t* status;              // [P][2d]
    int32_t* lp_iters;           // nullable: += iterations
    struct Tune {                          // unit objectives: the active set shows after the first iteration
        static constexpr int kUnroll = 2;
        static constexpr double kEarly = 1.0, kNext = 1e-1;
    };
    __device__ int n() const { return d; }
    __device__ int count(long long p) const {
        if (need_flags && !(need_flags[p] & need_mask)) return 0;
        return 2 * d;
    }
    template <int NS>
    __device__ int stage(long long p, double* G, double* hA, double* hB, int lane, LaneStage& sg) const {
        uint64_t mask;
        if (rows) mask = rows[p] & low_bits(m);
        else mask = low_bits(m_rows ? min(max(m_rows[p], 0), m) : m);
        const int cnt = lane_stage_masked<NS>(A + (size_t)p * m * d, b + (size_t)p * m, m, d, mask, G, hA, lane, sg);
        if (renorm) lane_renormalize<NS>(G, hA, cnt, d, lane);
        return cnt;
    }
    template <int NS>
    __device__ void setup(LaneData<NS>& dat, int q) const {
        dat.k = -1;
        dat.cj = q < d ? q : q - d;
        dat.cs = q < d ? 1.0 : -1.0;
    }
    template <int NS>
    __device__ void store(long long p, int q, const lane::Result<NS>& res) const {
        const int i = q < d ? q : q - d;
        double xi = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j)
            if (j == i) xi = res.x[j];
        (q < d ? val_lo : val_hi)[p * d + i] = res.status == lane::OPTIMAL ? xi : 0.0;
        status[p * 2 * d + q] = (int8_t)res.status;
        if (need_flags && (res.status == lane::ITER_LIMIT || res.status == lane::NUMERICAL)) {
            atomicOr(need_flags + p, retry_bit);
            if (retry_any) *retry_any = 1u;
        }
        if (lp_iters) atomicAdd(lp_iters + p, res.iters);
    }
};

template <int NS, class Prob>
__global__ void __launch_bounds__(32, NS <= 8 ? PB200_LANE_MINB : 2) lane_kernel(const Prob prob, long long P, int m, unsigned long long* counter) {
    extern __shared__ __align__(16) double smem[];
    constexpr int LANE_NG = lane_ng<NS>();
    const int lane = threadIdx.x;
    const int GS = lane_slot_doubles<NS>(m);
    double* G = smem;
    double* hA = G + LANE_NG * GS;
    double* hB = hA + LANE_NG * m;
    double* sz = hB + LANE_NG * m;
    LaneStage sg;
    sg.raw = sz + 64 * (size_t)m;
    sg.bar = reinterpret_cast<uint64_t*>(sg.raw + (size_t)m * NS + m + (m & 1));
    sg.parity = 0;
    if (lane == 0) mbar_init(sg.bar, 1);
    __syncwarp();
    long long carry_p = -1;
    int carry_k = 0;
    for (;;) {
        // ---- fill the lanes with LPs of the next polytopes ----
        int nl = 0, ns = 0;
        long long my_p = -1;
        int my_k = 0, my_slot = 0, my_rows = 0;
        while (nl < 32 && ns < LANE_NG) {
            long long p;
            int k0 = 0;
            if (carry_p >= 0) {
                p = carry_p;
                k0 = carry_k;
                carry_p = -1;
            } else {
                unsigned long long t = 0;
                if (lane == 0) t = atomicAdd(counter, 1ull);
                p = (long long)__shfl_sync(FULL_MASK, t, 0);
                if (p >= P) break;
            }
            const int cnt = prob.count(p);
            if (cnt - k0 <= 0) continue;
            const int rows = prob.template stage<NS>(p, G + ns * GS, hA + ns * m, hB + ns * m, lane, sg);
            const int take = min(cnt - k0, 32 - nl);
            if (lane >= nl && lane < nl + take) { my_p = p; my_k = k0 + lane - nl; my_slot = ns; my_rows = rows; }
            nl += take;
            if (k0 + take < cnt) { carry_p = p; carry_k = k0 + take; }
            ++ns;
        }
        if (nl == 0) break;
        __syncwarp();
        LaneData<NS> dat;
        dat.G = G + my_slot * GS;
        dat.hA = hA + my_slot * m;
        dat.hB = hB + my_slot * m;
        dat.sz = sz + lane;
        dat.Lm = reinterpret_cast<double*>(sg.bar + 1) + lane;
        dat.m = my_rows;
        prob.template setup<NS>(dat, my_k);
        lane::Result<NS> res;
        if constexpr (NS <= 8) lane::lane_solve<NS, LaneData<NS>, lane::WarpLanes, typename Prob::Tune>(dat, my_p >= 0, prob.n(), res);
        else lane::lane_solve_wide<NS, LaneData<NS>, lane::WarpLanes>(dat, my_p >= 0, prob.n(), res);
        if (my_p >= 0) prob.template store<NS>(my_p, my_k, res);
        __syncwarp();
    }
}


// ------------------------------------------------------------------------
// Independent small LPs (every LP has its own G): one LP per lane, the lane's rows in a
// lane-interleaved shared-memory array (double2 granules: 128-bit loads, consecutive lanes on
// consecutive granules, no bank conflicts).  Used when the LPs are small enough for a useful
// number of resident warps: is_adjacent pairs of grid cells (prop2partition, cfg5: 8 x 3), the
// Chebyshev LPs of is_fulldim / cheby_ball over a Region (cfg3: 16 x 7), envelope tests.
// ------------------------------------------------------------------------
template <int NS>
struct OwnData {
    double* G;      // this lane's granule column: element (i, j) at G[(i * (NS / 2) + (j >> 1)) * 64 + (j & 1)]
    double* hv;     // hv[i * 32]
    double* sz;
    int m, cj;      // rows; objective = cs * e_cj
    double cs;
    __device__ __forceinline__ int rows() const { return m; }
    __device__ __forceinline__ void row(int i, double (&g)[NS]) const {
        const double2* p = reinterpret_cast<const double2*>(G) + i * (NS / 2) * 32;
#pragma unroll
        for (int j = 0; j < NS / 2; ++j) { const double2 v = p[j * 32]; g[2 * j] = v.x; g[2 * j + 1] = v.y; }
    }
    __device__ __forceinline__ void put_row(int i, const double (&g)[NS]) {
        double2* p = reinterpret_cast<double2*>(G) + i * (NS / 2) * 32;
#pragma unroll
        for (int j = 0; j < NS / 2; ++j) p[j * 32] = make_double2(g[2 * j], g[2 * j + 1]);
    }
    __device__ __forceinline__ double h(int i) const { return hv[i * 32]; }
    __device__ __forceinline__ double c(int j) const { return j == cj ? cs : 0.0; }
    __device__ __forceinline__ double& s(int i) { return sz[(2 * i) * 32]; }
    __device__ __forceinline__ double& z(int i) { return sz[(2 * i + 1) * 32]; }
};

template <int NS>
__host__ __device__ inline size_t own_smem_doubles(int mr) { return (size_t)mr * 32 * (NS + 3); }

// One constraint row of a Chebyshev LP, built the way the reference builds it: the Polytope
// constructor's normalisation (polytope.py:128-138, rows with norm <= 1e-10 dropped -> staged as
// 0'x <= 1) when `renorm`, then the norm column ||a_i|| of cheby_ball (polytope.py:1285-1286).
// Sums of squares in numpy's order (d < 8: plain left-to-right).
template <int NS>
__device__ __forceinline__ void cheby_row(const double* __restrict__ a, double bi, int d, bool renorm, double (&g)[NS], double& h) {
    double ss = 0.0;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        g[j] = j < d ? __ldg(a + j) : 0.0;
        if (j < d) ss = __dadd_rn(ss, __dmul_rn(g[j], g[j]));
    }
    h = bi;
    if (renorm) {
        const double nrm = sqrt(ss);
        if (nrm > 1e-10) {
            const double mult = __ddiv_rn(1.0, nrm);
#pragma unroll
            for (int j = 0; j < NS; ++j) g[j] = j < d ? __dmul_rn(g[j], mult) : 0.0;
            h = __dmul_rn(bi, mult);
        } else {
#pragma unroll
            for (int j = 0; j < NS; ++j) g[j] = 0.0;
            h = 1.0;
        }
        ss = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j)
            if (j < d) ss = __dadd_rn(ss, __dmul_rn(g[j], g[j]));
    }
    const double col = sqrt(ss);
#pragma unroll
    for (int j = 0; j < NS; ++j)
        if (j == d) g[j] = col;
}

// is_adjacent, overlap=True (polytope.py:1856-1866): rows of both cells, b + tol, constructor
// normalisation, Chebyshev LP, radius > tol/10.  Pairs from a list, or enumerated:
// order 0: t = i(i-1)/2 + j, j < i (find_adjacent_regions, prop2partition.py:57-61);
// order 1: all ordered pairs i != j, row-major (MetricPartition.compute_adj, :253-261).
struct AdjacentOwn {
    const double *A, *b;
    int ncell, mc, d;
    const int32_t *pi, *pj;
    int order;
    long long t_begin;
    double abs_tol;
    uint8_t* adjacent;
    double* radius;
    int8_t* status;
    __device__ int n() const { return d + 1; }
    __device__ void pair(long long t, int& i, int& j) const {
        if (pi) { i = pi[t]; j = pj[t]; return; }
        const long long u = t + t_begin;
        if (order == 1) {
            const long long ii = u / (ncell - 1);
            const long long jj = u - ii * (ncell - 1);
            i = (int)ii;
            j = (int)(jj + (jj >= ii ? 1 : 0));
            return;
        }
        long long ii = (long long)((1.0 + sqrt(1.0 + 8.0 * (double)u)) * 0.5);
        while (ii * (ii - 1) / 2 > u) --ii;
        while ((ii + 1) * ii / 2 <= u) ++ii;
        i = (int)ii;
        j = (int)(u - ii * (ii - 1) / 2);
    }
    template <int NS>
    __device__ void stage(long long t, OwnData<NS>& dat) const {
        int ci, cj;
        pair(t, ci, cj);
        dat.m = 2 * mc;
        dat.cj = d;
        dat.cs = -1.0;
        for (int r = 0; r < 2 * mc; ++r) {
            const int cell = r < mc ? ci : cj, i = r < mc ? r : r - mc;
            double g[NS], h;
            cheby_row<NS>(A + ((size_t)cell * mc + i) * d, __dadd_rn(__ldg(b + (size_t)cell * mc + i), abs_tol), d, true, g, h);
            dat.put_row(r, g);
            dat.hv[r * 32] = h;
        }
    }
    template <int NS>
    __device__ void store(long long t, const lane::Result<NS>& res) const {
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        double rr = nan;
#pragma unroll
        for (int j = 0; j < NS; ++j)
            if (j == d && res.status == lane::OPTIMAL) rr = res.x[j];
        adjacent[t] = (res.status == lane::OPTIMAL && rr > abs_tol / 10) ? 1 : 0;
        if (radius) radius[t] = rr;
        if (status) status[t] = (int8_t)res.status;
    }
};

// cheby_ball of P stacked polytopes (polytope.py:1280-1300), rows used as given
struct ChebyOwn {
    const double *A, *b;
    const int32_t* m_rows;
    const uint64_t* rows;     // nullable row masks
    int m, d;
    double *r, *xc;
    int8_t* status;
    int32_t* lp_iters;        // nullable: = iterations of this LP
    __device__ int n() const { return d + 1; }
    template <int NS>
    __device__ void stage(long long p, OwnData<NS>& dat) const {
        const double* Ap = A + (size_t)p * m * d;
        const double* bp = b + (size_t)p * m;
        uint64_t mask = rows ? (rows[p] & low_bits(m)) : low_bits(m_rows ? min(max(m_rows[p], 0), m) : m);
        int cnt = 0;
        for (int i = 0; i < m; ++i) {
            if (!((mask >> i) & 1ull)) continue;
            double g[NS], h;
            cheby_row<NS>(Ap + (size_t)i * d, __ldg(bp + i), d, false, g, h);
            dat.put_row(cnt, g);
            dat.hv[cnt * 32] = h;
            ++cnt;
        }
        dat.m = cnt;
        dat.cj = d;
        dat.cs = -1.0;
    }
    template <int NS>
    __device__ void store(long long p, const lane::Result<NS>& res) const {
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        const bool ok = res.status == lane::OPTIMAL;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            if (j < d) xc[(size_t)p * d + j] = ok ? res.x[j] : nan;
            if (j == d) r[p] = ok ? res.x[j] : nan;
        }
        status[p] = (int8_t)res.status;
        if (lp_iters) lp_iters[p] = res.iters;
    }
};

// (A one-LP-per-lane kernel for Chebyshev LPs of 9 <= n <= 16 columns that read the rows from global memory was
// measured in r02h and dropped: 0.72 vs 0.42 ms on cfg2's 10 000 LPs and 1.8 vs 0.1 ms on cfg4's 1000 against the
// warp-per-LP kernel -- one LP per polytope leaves too few warps, and every row pass waits for L2.)

#ifndef PB200_OWN_MINB4
#define PB200_OWN_MINB4 16      // NS = 4: 10 matrix entries, <= 128 registers -> 16 warps per SM
#endif
template <int NS, class Prob>
__global__ void __launch_bounds__(32, NS <= 4 ? PB200_OWN_MINB4 : 8) lane_own_kernel(const Prob prob, long long T, int mr, unsigned long long* counter) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x;
    OwnData<NS> dat;
    dat.G = smem + 2 * lane;
    dat.hv = smem + (size_t)mr * NS * 32 + lane;
    dat.sz = dat.hv - lane + (size_t)mr * 32 + lane;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, 32ull);
        base = __shfl_sync(FULL_MASK, base, 0);
        if ((long long)base >= T) break;
        const long long t = (long long)base + lane;
        const bool has = t < T;
        dat.m = 0;
        dat.cj = 0;
        dat.cs = 0.0;
        if (has) prob.template stage<NS>(t, dat);
        __syncwarp();
        lane::Result<NS> res;
        lane::lane_solve<NS, OwnData<NS>, lane::WarpLanes>(dat, has, prob.n(), res);
        if (has) prob.template store<NS>(t, res);
        __syncwarp();
    }
}

}  // namespace pb200

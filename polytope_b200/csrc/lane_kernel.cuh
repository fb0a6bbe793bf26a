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
#include "lp_lane.cuh"

namespace pb200 {

constexpr int LANE_NS = 8;       // padded columns of the lane solver
constexpr int LANE_NG = 3;       // polytopes a warp holds per round
#ifndef PB200_LANE_MINB
#define PB200_LANE_MINB 8
#endif

__host__ __device__ inline int lane_slot_doubles(int m) { return m * LANE_NS + 4; }   // +4: slots land in different banks
__host__ __device__ inline size_t lane_smem_doubles(int m) {
    return (size_t)LANE_NG * lane_slot_doubles(m) + 2 * LANE_NG * m + 64 * (size_t)m;
}

// what a lane sees of its LP
struct LaneData {
    const double* G;     // [rows][LANE_NS] row-major, shared by the lanes of the same polytope
    const double* hA;    // right-hand side
    const double* hB;    // right-hand side after the reference's +0.1 / -0.1 round trip (rows before k)
    double* sz;          // this lane's (s_i, z_i): sz[(2 i) * 32], sz[(2 i + 1) * 32]
    int m, k;            // rows; row whose h carries +0.1 (-1: none)
    int cj;              // objective: cs * e_cj, or -G[k] when cj < 0
    double cs;
    __device__ __forceinline__ int rows() const { return m; }
    __device__ __forceinline__ void row(int i, double (&g)[LANE_NS]) const {
        const double2* p = reinterpret_cast<const double2*>(G + i * LANE_NS);
#pragma unroll
        for (int j = 0; j < LANE_NS / 2; ++j) { const double2 v = p[j]; g[2 * j] = v.x; g[2 * j + 1] = v.y; }
    }
    __device__ __forceinline__ double h(int i) const {
        const double* src = i < k ? hB : hA;          // rows before k carry the +0.1 / -0.1 round trip
        const double a = src[i];
        return i == k ? __dadd_rn(a, 0.1) : a;
    }
    __device__ __forceinline__ double c(int j) const { return cj < 0 ? -G[k * LANE_NS + j] : (j == cj ? cs : 0.0); }
    __device__ __forceinline__ double& s(int i) { return sz[(2 * i) * 32]; }
    __device__ __forceinline__ double& z(int i) { return sz[(2 * i + 1) * 32]; }
};

// rows selected by `mask` (ascending) of a row-major [m x d] matrix -> row-major [cnt][LANE_NS], zero padded
__device__ __forceinline__ int lane_stage_masked(const double* __restrict__ Ap, const double* __restrict__ bp, int m, int d,
                                                 uint64_t mask, double* G, double* hA, int lane) {
    const int cnt = __popcll(mask);
    for (int e = lane; e < cnt * LANE_NS; e += 32) G[e] = 0.0;
    __syncwarp();
    const int total = m * d;
    for (int e = lane; e < total; e += 32) {
        const int i = e / d, j = e - i * d;
        if ((mask >> i) & 1ull) G[__popcll(mask & ((1ull << i) - 1ull)) * LANE_NS + j] = __ldg(Ap + e);
    }
    for (int i = lane; i < m; i += 32)
        if ((mask >> i) & 1ull) hA[__popcll(mask & ((1ull << i) - 1ull))] = __ldg(bp + i);
    __syncwarp();
    return cnt;
}

// Polytope.__init__ normalisation of the staged rows (polytope.py:128-138), numpy's summation order
__device__ __forceinline__ void lane_renormalize(double* G, double* hA, int cnt, int d, int lane) {
    for (int i = lane; i < cnt; i += 32) {
        double* row = G + i * LANE_NS;
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
    int m, d;
    double abs_tol;
    unsigned long long* keep;   // OR-accumulated
    int32_t* lp_iters;          // nullable: += interior-point iterations
    __device__ int n() const { return d; }
    __device__ int count(long long p) const {
        if (!(flags[p] & run_mask)) return 0;
        return __popcll(rows[p] & low_bits(m));
    }
    __device__ int stage(long long p, double* G, double* hA, double* hB, int lane) const {
        const int cnt = lane_stage_masked(A + (size_t)p * m * d, b + (size_t)p * m, m, d, rows[p] & low_bits(m), G, hA, lane);
        for (int i = lane; i < cnt; i += 32) hB[i] = __dadd_rn(__dadd_rn(hA[i], 0.1), -0.1);
        __syncwarp();
        return cnt;
    }
    __device__ void setup(LaneData& dat, int k) const { dat.k = k; dat.cj = -1; dat.cs = 0.0; }
    __device__ void store(long long p, int k, const lane::Result<LANE_NS>& res) const {
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
    int m, d, renorm;
    double *val_lo, *val_hi;     // [P][d] each: optimised coordinate of the lower / upper LP
    int8_t* status;              // [P][2d]
    int32_t* lp_iters;           // nullable: += iterations
    __device__ int n() const { return d; }
    __device__ int count(long long p) const {
        if (need_flags && !(need_flags[p] & need_mask)) return 0;
        return 2 * d;
    }
    __device__ int stage(long long p, double* G, double* hA, double* hB, int lane) const {
        uint64_t mask;
        if (rows) mask = rows[p] & low_bits(m);
        else mask = low_bits(m_rows ? min(max(m_rows[p], 0), m) : m);
        const int cnt = lane_stage_masked(A + (size_t)p * m * d, b + (size_t)p * m, m, d, mask, G, hA, lane);
        if (renorm) lane_renormalize(G, hA, cnt, d, lane);
        return cnt;
    }
    __device__ void setup(LaneData& dat, int q) const {
        dat.k = -1;
        dat.cj = q < d ? q : q - d;
        dat.cs = q < d ? 1.0 : -1.0;
    }
    __device__ void store(long long p, int q, const lane::Result<LANE_NS>& res) const {
        const int i = q < d ? q : q - d;
        (q < d ? val_lo : val_hi)[p * d + i] = res.status == lane::OPTIMAL ? res.x[i] : 0.0;
        status[p * 2 * d + q] = (int8_t)res.status;
        if (need_flags && (res.status == lane::ITER_LIMIT || res.status == lane::NUMERICAL)) atomicOr(need_flags + p, retry_bit);
        if (lp_iters) atomicAdd(lp_iters + p, res.iters);
    }
};

template <class Prob>
__global__ void __launch_bounds__(32, PB200_LANE_MINB) lane_kernel(const Prob prob, long long P, int m, unsigned long long* counter) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x;
    const int GS = lane_slot_doubles(m);
    double* G = smem;
    double* hA = G + LANE_NG * GS;
    double* hB = hA + LANE_NG * m;
    double* sz = hB + LANE_NG * m;
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
            const int rows = prob.stage(p, G + ns * GS, hA + ns * m, hB + ns * m, lane);
            const int take = min(cnt - k0, 32 - nl);
            if (lane >= nl && lane < nl + take) { my_p = p; my_k = k0 + lane - nl; my_slot = ns; my_rows = rows; }
            nl += take;
            if (k0 + take < cnt) { carry_p = p; carry_k = k0 + take; }
            ++ns;
        }
        if (nl == 0) break;
        __syncwarp();
        LaneData dat;
        dat.G = G + my_slot * GS;
        dat.hA = hA + my_slot * m;
        dat.hB = hB + my_slot * m;
        dat.sz = sz + lane;
        dat.m = my_rows;
        prob.setup(dat, my_k);
        lane::Result<LANE_NS> res;
        lane::lane_solve<LANE_NS, LaneData, lane::WarpLanes>(dat, my_p >= 0, prob.n(), res);
        if (my_p >= 0) prob.store(my_p, my_k, res);
        __syncwarp();
    }
}

}  // namespace pb200

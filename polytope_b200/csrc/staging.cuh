// Staging helpers shared by the LP-solving kernels: rows of a row-major (A, b)
// into the warp's column-major shared-memory scratch, the Polytope constructor's
// row normalisation and the Chebyshev norm column.  One warp per call.
#pragma once
#include "common.cuh"
#include "lp_warp.cuh"

namespace pb200 {

__device__ __forceinline__ uint64_t low_bits(int m) { return m >= 64 ? ~0ull : ((1ull << m) - 1ull); }

// index of the k-th (0-based) set bit of mask; mask must have > k bits set
__device__ __forceinline__ int nth_set_bit(uint64_t mask, int k) {
    for (int i = 0; i < k; ++i) mask &= mask - 1;
    return __ffsll((long long)mask) - 1;
}

// ------------------------------------------------------------------------
// staging helpers (one warp)
// ------------------------------------------------------------------------
template <int RPL, class S>
__device__ __forceinline__ void zero_G(const S& w, int lane) {
    for (int e = lane; e < w.NC * w.MP; e += 32) w.G[e] = 0.0;
    __syncwarp();
}

// rows [0, cnt) of a row-major [.. x ld] matrix -> column-major slots; h from bp
template <int RPL, class S>
__device__ __forceinline__ void stage_first_rows(const S& w, const double* __restrict__ Ap,
                                                 const double* __restrict__ bp, int cnt, int d,
                                                 int lane, double (&h)[RPL]) {
    zero_G<RPL>(w, lane);
    const int total = cnt * d;
    for (int e = lane; e < total; e += 32) {
        const int i = e / d, j = e - i * d;
        w.G[j * w.MP + i] = __ldg(Ap + e);
    }
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int i = lane + 32 * r;
        h[r] = i < cnt ? __ldg(bp + i) : 0.0;
    }
    __syncwarp();
}

// rows selected by `mask` (ascending) of a row-major [m x d] matrix -> slots 0..cnt-1
template <int RPL, class S>
__device__ __forceinline__ int stage_masked_rows(const S& w, const double* __restrict__ Ap,
                                                 const double* __restrict__ bp, int m, int d,
                                                 uint64_t mask, int lane, double (&h)[RPL]) {
    zero_G<RPL>(w, lane);
    const int total = m * d;
    for (int e = lane; e < total; e += 32) {
        const int i = e / d, j = e - i * d;
        if ((mask >> i) & 1ull) {
            const int slot = __popcll(mask & ((1ull << i) - 1ull));
            w.G[j * w.MP + slot] = __ldg(Ap + e);
        }
    }
    for (int i = lane; i < m; i += 32)
        if ((mask >> i) & 1ull) w.d[__popcll(mask & ((1ull << i) - 1ull))] = __ldg(bp + i);
    __syncwarp();
    const int cnt = __popcll(mask);
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int i = lane + 32 * r;
        h[r] = i < cnt ? w.d[i] : 0.0;
    }
    __syncwarp();
    return cnt;
}

// Polytope.__init__ normalisation of the staged rows (polytope.py:128-138):
// row / ||row||_2, b / ||row||_2; rows with norm <= 1e-10 are dropped (h = +inf
// tells the solver the row does not exist).
template <int RPL, class S>
__device__ __forceinline__ void renormalize_rows(const S& w, int cnt, int d, int lane, double (&h)[RPL]) {
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int i = lane + 32 * r;
        if (i < cnt) {
            const double* row = w.G + i;
            const int MP = w.MP;
            const double nrm = sqrt(np_sum_squares([&](int j) { return row[j * MP]; }, d));
            if (nrm > 1e-10) {
                const double mult = __ddiv_rn(1.0, nrm);
                for (int j = 0; j < d; ++j) w.G[j * MP + i] = __dmul_rn(row[j * MP], mult);
                h[r] = __dmul_rn(h[r], mult);
            } else {
                for (int j = 0; j < d; ++j) w.G[j * MP + i] = 0.0;
                h[r] = 1e308 * 10.0;
            }
        }
    }
    __syncwarp();
}

// extra column d = ||row||_2 of the staged rows (polytope.py:1285-1286)
template <int RPL, class S>
__device__ __forceinline__ void append_norm_column(const S& w, int cnt, int d, int lane) {
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int i = lane + 32 * r;
        if (i < cnt) {
            const double* row = w.G + i;
            const int MP = w.MP;
            w.G[d * MP + i] = sqrt(np_sum_squares([&](int j) { return row[j * MP]; }, d));
        }
    }
    __syncwarp();
}

// lpsolve() callers floor()/ceil() coordinates of optimal vertices (enumerate_integral_points,
// polytope.py:2352-2358, pinned by the reference's test_enumerate_integral_points): a simplex
// returns x_j = h_i / G_ij without rounding error when an active row has the single non-zero
// entry G_ij.  The interior-point + polish solution sits within ~1e-16 of that; this snaps it:
// a solution component within 1e-12 of such a row's h_i / G_ij becomes exactly that quotient
// (lowest row wins).  x_own is the lane-owned component (lanes < n); vec is >= n doubles of scratch.
template <int RPL, class S>
__device__ __forceinline__ double snap_axis_rows(const S& w, double* vec, int m, int n, int lane,
                                                 const double (&h)[RPL], double x_own) {
    if (lane < n) vec[lane] = x_own;
    __syncwarp();
    int cj[RPL];
    double cv[RPL];
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int i = lane + 32 * r;
        cj[r] = -1;
        cv[r] = 0.0;
        if (i < m && h[r] < 1e300) {
            int nz = 0, jj = 0;
            double a = 1.0;
            for (int j = 0; j < n; ++j) {
                const double g = w.G[j * w.MP + i];
                if (g != 0.0) { ++nz; jj = j; a = g; }
            }
            if (nz == 1) {
                const double v = __ddiv_rn(h[r], a);
                if (fabs(v - vec[jj]) <= 1e-12 * fmax(1.0, fabs(v))) { cj[r] = jj; cv[r] = v; }
            }
        }
    }
    double out = x_own;
    for (int j = 0; j < n; ++j) {
        bool found = false;
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const unsigned bal = __ballot_sync(FULL_MASK, !found && cj[r] == j);
            if (bal) {
                const double v = __shfl_sync(FULL_MASK, cv[r], __ffs(bal) - 1);
                if (lane == j) out = v;
                found = true;
            }
        }
    }
    __syncwarp();
    return out;
}

}  // namespace pb200

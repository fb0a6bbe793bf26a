// Host-side plumbing shared by the translation units of libpolytope_b200.so:
// error reporting across the C ABI, the launch counter, numpy-order sums.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/polytope_b200.h"

namespace pb200 {

// defined in pb200.cu
int fail(int code, const char* msg);
char* err_buf(size_t* cap);
void count_launch(int n = 1);
int sm_count();      // SMs of the current device (0 on failure, error recorded)

#define PB_CHECK_CUDA(expr)                                                              \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            size_t cap__;                                                                 \
            char* buf__ = pb200::err_buf(&cap__);                                         \
            snprintf(buf__, cap__, "%s failed: %s (%s:%d)", #expr,                        \
                     cudaGetErrorString(e__), __FILE__, __LINE__);                        \
            return PB200_ECUDA;                                                           \
        }                                                                                 \
    } while (0)

static inline unsigned blocks_for(long long threads, int block) { return (unsigned)((threads + block - 1) / block); }

// ------------------------------------------------------------------------
// numpy-order arithmetic.  np.sum over a contiguous axis uses pairwise
// summation with 8 accumulators (numpy/_core/src/umath/loops_utils.h.src,
// *_pairwise_sum); for n <= 128 that is the code below.  The reference
// computes every row norm that way (polytope.py:129, :1094, :1285), and
// tests/test_gpu_polytope.py checks bit-equality against numpy.
// ------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ double np_sum_squares(F elem, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int j = 0; j < n; ++j) { const double a = elem(j); res = __dadd_rn(res, __dmul_rn(a, a)); }
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { const double a = elem(j); r[j] = __dmul_rn(a, a); }
    int i = 8;
    for (; i < n - (n % 8); i += 8)
#pragma unroll
        for (int j = 0; j < 8; ++j) { const double a = elem(i + j); r[j] = __dadd_rn(r[j], __dmul_rn(a, a)); }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) { const double a = elem(i); res = __dadd_rn(res, __dmul_rn(a, a)); }
    return res;
}

}  // namespace pb200

// The batched (A|b) read: Polytope.__init__'s row normalisation
// (polytope/polytope.py:128-138) of P stacked polytopes as an HBM-streaming
// kernel.  A CTA owns a tile of NW consecutive polytopes, i.e. one contiguous
// span of A and one of b:
//   load    the two spans into shared memory -- as two 1-D bulk async copies
//           (cp.async.bulk, the TMA engine; SASS UBLKCP) completing on an
//           mbarrier, or as coalesced 128-bit / 64-bit LDGs when the spans are
//           not 16-byte aligned;
//   norms   one warp per polytope: 32 rows at a time are copied into a scratch
//           with an odd row stride (conflict-free row-per-lane reads), each lane
//           sums its row's squares in numpy's order (common.cuh) and leaves the
//           multiplier 1/||row|| and the scaled b in shared memory;
//   scale   element-wise in place, consecutive lanes on consecutive elements;
//   store   the two spans back with bulk async copies (or coalesced STGs).
// Every byte of A and b crosses HBM once in each direction; nothing is re-read.
#pragma once
#include "common.cuh"

namespace pb200 {

enum { NORM_LDG64 = 0, NORM_LDG128 = 1, NORM_BULK = 2 };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(__cvta_generic_to_global(src_gmem)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    // try_wait suspends in hardware; a copy that never lands is a programming
    // error and must not hang the GPU
    for (long long spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (spins > (1ll << 24)) __trap();
    }
}
__device__ __forceinline__ void bulk_store(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(dst_gmem)),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_store_commit_and_drain() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_smem_to_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// doubles of dynamic shared memory for nw polytopes per CTA
static inline size_t normalize_smem_doubles(int nw, int m, int d) {
    return (size_t)nw * ((size_t)m * d + 2 * (size_t)m + 32 * (size_t)(d | 1));
}

// span copies of the LDG variants; VEC = 2 needs 16-byte aligned spans of even length
template <int VEC>
__device__ __forceinline__ void span_in(double* dst, const double* __restrict__ src, int n, int tid, int nthr) {
    if (VEC == 2) {
        const double2* s2 = reinterpret_cast<const double2*>(src);
        double2* d2 = reinterpret_cast<double2*>(dst);
        for (int e = tid; e < (n >> 1); e += nthr) d2[e] = __ldcs(s2 + e);
    } else {
        for (int e = tid; e < n; e += nthr) dst[e] = __ldcs(src + e);
    }
}
template <int VEC>
__device__ __forceinline__ void span_out(double* __restrict__ dst, const double* src, int n, int tid, int nthr) {
    if (VEC == 2) {
        const double2* s2 = reinterpret_cast<const double2*>(src);
        double2* d2 = reinterpret_cast<double2*>(dst);
        for (int e = tid; e < (n >> 1); e += nthr) d2[e] = s2[e];
    } else {
        for (int e = tid; e < n; e += nthr) dst[e] = src[e];
    }
}

// blockDim.x = 32 * NW, one CTA per tile of NW polytopes; div_magic = 2^20 / d + 1
template <int MODE>
__global__ void normalize_tile_kernel(const double* __restrict__ A, const double* __restrict__ b,
                                      const int32_t* __restrict__ m_rows, int P, int m, int d, int do_norm,
                                      uint32_t div_magic, double* __restrict__ An, double* __restrict__ bn,
                                      uint64_t* __restrict__ valid) {
    extern __shared__ __align__(128) double nsm[];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, w = tid >> 5, NW = nthr >> 5;
    const int md = m * d, ds = d | 1;
    double* tA = nsm;                       // [NW][m][d]   the A span, scaled in place
    double* tb = tA + (size_t)NW * md;      // [NW][m]      the b span, scaled in place
    double* mult = tb + (size_t)NW * m;     // [NW][m]      row multipliers (< 0: row beyond m_rows)
    double* pad = mult + (size_t)NW * m + (size_t)w * 32 * ds;   // per warp [32][ds]
    const long long p0 = (long long)blockIdx.x * NW;
    const int np = (int)min((long long)NW, (long long)P - p0);
    const int nA = np * md, nb = np * m;
    const double* gA = A + (size_t)p0 * md;
    const double* gb = b + (size_t)p0 * m;

    if (MODE == NORM_BULK) {
        if (tid == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&bar, (uint32_t)(nA + nb) * 8u);
            bulk_load(tA, gA, (uint32_t)nA * 8u, &bar);
            bulk_load(tb, gb, (uint32_t)nb * 8u, &bar);
        }
        mbar_wait(&bar, 0);
    } else {
        span_in<MODE == NORM_LDG128 ? 2 : 1>(tA, gA, nA, tid, nthr);
        span_in<MODE == NORM_LDG128 ? 2 : 1>(tb, gb, nb, tid, nthr);
        __syncthreads();
    }

    if (w < np) {
        double* a = tA + (size_t)w * md;
        double* bb = tb + (size_t)w * m;
        double* mu = mult + (size_t)w * m;
        const long long p = p0 + w;
        const int mm = m_rows ? min(max(m_rows[p], 0), m) : m;
        uint64_t mask = 0;
        for (int base = 0; base < m; base += 32) {
            const int cnt = min(32, m - base) * d;
            const double* src = a + base * d;
            for (int e = lane; e < cnt; e += 32) {
                const int i = (int)(((uint32_t)e * div_magic) >> 20);
                pad[i * ds + (e - i * d)] = src[e];
            }
            __syncwarp();
            const int i = base + lane;
            bool ok = false;
            if (i < m) {
                double mlt = -1.0;
                if (i < mm && !do_norm) {   // rows already normalised by a constructor
                    ok = true;
                    mlt = 1.0;
                } else if (i < mm) {
                    const double* row = pad + lane * ds;
                    const double nrm = sqrt(np_sum_squares([&](int j) { return row[j]; }, d));
                    ok = nrm > 1e-10;
                    mlt = ok ? __ddiv_rn(1.0, nrm) : 0.0;
                    bb[i] = ok ? __dmul_rn(bb[i], mlt) : 0.0;
                } else {
                    bb[i] = 0.0;
                }
                mu[i] = mlt;
            }
            mask |= (uint64_t)__ballot_sync(0xffffffffu, ok) << base;
            __syncwarp();                   // scratch and multipliers are read across lanes
        }
        for (int e = lane; e < md; e += 32) {
            const int i = (int)(((uint32_t)e * div_magic) >> 20);
            const double mlt = mu[i];
            a[e] = mlt < 0.0 ? 0.0 : __dmul_rn(a[e], mlt);
        }
        if (lane == 0 && valid) valid[p] = mask;
    }

    double* oA = An + (size_t)p0 * md;
    double* ob = bn + (size_t)p0 * m;
    if (MODE == NORM_BULK) {
        fence_smem_to_async_proxy();        // generic-proxy writes above -> visible to the copy engine
        __syncthreads();
        if (tid == 0) {
            bulk_store(oA, tA, (uint32_t)nA * 8u);
            bulk_store(ob, tb, (uint32_t)nb * 8u);
            bulk_store_commit_and_drain();  // shared memory must outlive the copies
        }
    } else {
        __syncthreads();
        span_out<MODE == NORM_LDG128 ? 2 : 1>(oA, tA, nA, tid, nthr);
        span_out<MODE == NORM_LDG128 ? 2 : 1>(ob, tb, nb, tid, nthr);
    }
}

}  // namespace pb200

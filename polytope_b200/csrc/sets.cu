// polytope_b200: point-set kernels around the LP path (sm_100a).
//
//   contains_kernel   Polytope.contains / Region.contains   polytope/polytope.py:206-218, :736-748
//   volume_kernel     Monte-Carlo volume() containment test  polytope/polytope.py:1583-1592
//   sweep_kernel      quickhull's point-to-hyperplane sweep  polytope/quickhull.py:117-121, :228-237
//
// These are the streaming members of the family: every point is read from HBM
// exactly once (coalesced, coordinate-major as the reference lays them out:
// `points` is a d x N array of column vectors), the polytope rows are staged in
// shared memory and read by broadcast, and the result is one byte (or one
// counter update) per point.  volume() reads nothing at all: the uniform
// samples are regenerated on the fly from numpy's PCG64 stream, bit for bit.
//
// Bit-exactness: numpy's `A.dot(X)` is an OpenBLAS dgemm whose micro-kernel
// accumulates over k in order with FMAs; tests/test_oracle.py pins that the
// oracle's `A.dot` equals the fma chain `acc = fma(A[i][k], X[k][j], acc)` used
// below, so the strict `< 0` / `< abs_tol` decisions see identical doubles.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace pb200 {

typedef unsigned __int128 u128;

// ---- numpy's PCG64 (XSL-RR 128/64, numpy/random/src/pcg64/pcg64.h) ----
__host__ __device__ __forceinline__ u128 pcg_mult() {
    return ((u128)2549297995355413924ull << 64) | (u128)4865540595714422341ull;
}
struct Pcg {
    u128 state, inc;
};
// (mult, plus) such that advancing `delta` steps is s -> s * mult + plus
__host__ __device__ inline void pcg_jump_coeffs(u128 inc, unsigned long long delta, u128& acc_mult, u128& acc_plus) {
    u128 cur_mult = pcg_mult(), cur_plus = inc;
    acc_mult = 1;
    acc_plus = 0;
    while (delta > 0) {
        if (delta & 1) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
}
// step, then output from the new state (pcg64_random_r), mapped to [0, 1) as
// Generator.random does: (u >> 11) * 2^-53
__device__ __forceinline__ double pcg_next_double(u128& s, u128 inc) {
    s = s * pcg_mult() + inc;
    const uint64_t hi = (uint64_t)(s >> 64), lo = (uint64_t)s;
    const uint64_t x = hi ^ lo;
    const unsigned rot = (unsigned)(hi >> 58);
    const uint64_t u = (x >> rot) | (x << ((64u - rot) & 63u));
    return (double)(u >> 11) * (1.0 / 9007199254740992.0);
}

// rows of one polytope against a point held in registers; `strict_tol`:
// contains() tests  A x - b < abs_tol,  volume() tests  A x - b < 0
template <int D>
__device__ __forceinline__ bool inside_rows(const double* __restrict__ As, const double* __restrict__ bs, int mm, int d,
                                            const double (&x)[D > 0 ? D : 32], double tol) {
    bool ok = true;
    for (int i = 0; i < mm && ok; ++i) {
        double acc = 0.0;
        if (D > 0) {
#pragma unroll
            for (int k = 0; k < (D > 0 ? D : 1); ++k) acc = fma(As[i * D + k], x[k], acc);
        } else {
            for (int k = 0; k < d; ++k) acc = fma(As[i * d + k], x[k], acc);
        }
        ok = __dsub_rn(acc, bs[i]) < tol;
    }
    return ok;
}

// ------------------------------------------------------------------------
// contains: P polytopes x N points.  One thread per point, the polytopes are
// walked in shared-memory chunks.  mode 0: out[P][N]; mode 1: out[N] = OR over p.
// ------------------------------------------------------------------------
constexpr int SET_THREADS = 256;
constexpr int CHUNK_DOUBLES = 4096;      // 32 KB of staged rows per chunk

template <int D>
__global__ void __launch_bounds__(SET_THREADS) contains_kernel(const double* __restrict__ A, const double* __restrict__ b,
                                                               const int32_t* __restrict__ m_rows, int P, int m, int d,
                                                               const double* __restrict__ pts, long long N, double abs_tol,
                                                               int mode, uint8_t* __restrict__ out, int polys_per_chunk) {
    __shared__ double sh[CHUNK_DOUBLES];
    __shared__ int sh_rows[CHUNK_DOUBLES / 2];
    constexpr int DD = D > 0 ? D : 32;
    const long long j = (long long)blockIdx.x * SET_THREADS + threadIdx.x;
    const bool live = j < N;
    double x[DD];
    if (D > 0) {
#pragma unroll
        for (int k = 0; k < DD; ++k) x[k] = live ? __ldg(pts + (size_t)k * N + j) : 0.0;
    } else {
        for (int k = 0; k < d; ++k) x[k] = live ? __ldg(pts + (size_t)k * N + j) : 0.0;
    }
    bool any = false;
    const int per = m * (d + 1);
    for (int p0 = 0; p0 < P; p0 += polys_per_chunk) {
        const int np = min(polys_per_chunk, P - p0);
        __syncthreads();
        // stage [np][m][d] rows then [np][m] right-hand sides
        for (int e = threadIdx.x; e < np * m * d; e += SET_THREADS) sh[e] = __ldg(A + (size_t)p0 * m * d + e);
        for (int e = threadIdx.x; e < np * m; e += SET_THREADS) sh[np * m * d + e] = __ldg(b + (size_t)p0 * m + e);
        for (int e = threadIdx.x; e < np; e += SET_THREADS) sh_rows[e] = m_rows ? min(max(m_rows[p0 + e], 0), m) : m;
        __syncthreads();
        (void)per;
        for (int q = 0; q < np; ++q) {
            if (mode == 1 && any) break;
            const bool in = inside_rows<D>(sh + (size_t)q * m * d, sh + (size_t)np * m * d + q * m, sh_rows[q], d, x, abs_tol);
            if (mode == 0) {
                if (live) out[(size_t)(p0 + q) * N + j] = in ? 1 : 0;
            } else {
                any = any || in;
            }
        }
    }
    if (mode == 1 && live) out[j] = any ? 1 : 0;
}

// ------------------------------------------------------------------------
// volume: count[p] += #{ j < N : A (l + u_j * (hi - lo)) - b < 0 }, with
// u = default_rng(seed).random((d, N)) regenerated from the PCG64 state
// (polytope.py:1583-1591).  grid = (P, S): S CTAs share the samples of one
// polytope, thread t of split s takes samples j = s*T + t, then + S*T, ...
// ------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(SET_THREADS) volume_kernel(const double* __restrict__ A, const double* __restrict__ b,
                                                             const int32_t* __restrict__ m_rows, int m, int d,
                                                             const double* __restrict__ lo, const double* __restrict__ hi,
                                                             long long N, const uint64_t* __restrict__ rng,
                                                             unsigned long long* __restrict__ count) {
    extern __shared__ double sh[];       // [m][d] rows | [m] b
    constexpr int DD = D > 0 ? D : 32;
    __shared__ unsigned long long sh_cnt;
    __shared__ u128 sh_coef[4];          // jump by N: (mult, plus); jump by stride-1: (mult, plus)
    const int p = blockIdx.x;
    const int mm = m_rows ? min(max(m_rows[p], 0), m) : m;
    for (int e = threadIdx.x; e < m * d; e += SET_THREADS) sh[e] = __ldg(A + (size_t)p * m * d + e);
    for (int e = threadIdx.x; e < m; e += SET_THREADS) sh[m * d + e] = __ldg(b + (size_t)p * m + e);
    Pcg g;
    g.state = ((u128)rng[4 * p + 0] << 64) | rng[4 * p + 1];
    g.inc = ((u128)rng[4 * p + 2] << 64) | rng[4 * p + 3];
    const long long stride = (long long)gridDim.y * SET_THREADS;
    if (threadIdx.x == 0) {
        sh_cnt = 0;
        pcg_jump_coeffs(g.inc, (unsigned long long)N, sh_coef[0], sh_coef[1]);
        pcg_jump_coeffs(g.inc, (unsigned long long)(stride - 1), sh_coef[2], sh_coef[3]);
    }
    __syncthreads();
    const long long j0 = (long long)blockIdx.y * SET_THREADS + threadIdx.x;
    double l[DD], w[DD];
    u128 st[DD];
    {
        // stream position of coordinate k for sample j is k*N + j
        u128 am, ap;
        pcg_jump_coeffs(g.inc, (unsigned long long)j0, am, ap);
        u128 s = g.state * am + ap;
        const int dd = D > 0 ? DD : d;
#pragma unroll
        for (int k = 0; k < DD; ++k) {
            if (k < dd) {
                st[k] = s;
                s = s * sh_coef[0] + sh_coef[1];
                l[k] = lo[(size_t)p * d + k];
                w[k] = __dsub_rn(hi[(size_t)p * d + k], l[k]);
            }
        }
    }
    unsigned long long mine = 0;
    for (long long j = j0; j < N; j += stride) {
        double x[DD];
        const int dd = D > 0 ? DD : d;
#pragma unroll
        for (int k = 0; k < DD; ++k) {
            if (k < dd) {
                const double u = pcg_next_double(st[k], g.inc);
                x[k] = __dadd_rn(l[k], __dmul_rn(u, w[k]));
                st[k] = st[k] * sh_coef[2] + sh_coef[3];     // skip the draws of the other threads
            }
        }
        mine += inside_rows<D>(sh, sh + m * d, mm, d, x, 0.0) ? 1ull : 0ull;
    }
    mine = __reduce_add_sync(0xffffffffu, (unsigned)mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&sh_cnt, mine);
    __syncthreads();
    if (threadIdx.x == 0 && sh_cnt) atomicAdd(count + p, sh_cnt);
}

// ------------------------------------------------------------------------
// quickhull's distance sweep (quickhull.py:117-121): dist = sum(n * p) - off
// with numpy's summation order; per point the FIRST facet (in facet order) it
// lies outside of by more than tol -- the assignment rule of quickhull.py:228-246
// -- and the largest distance over all facets.
// points[N][d] row-major (as quickhull takes them), normals[F][d], offsets[F].
// ------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ double np_sum_products(F prod, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int j = 0; j < n; ++j) res = __dadd_rn(res, prod(j));
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = prod(j);
    int i = 8;
    for (; i < n - (n % 8); i += 8)
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], prod(i + j));
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, prod(i));
    return res;
}

__global__ void __launch_bounds__(SET_THREADS) sweep_kernel(const double* __restrict__ pts, const double* __restrict__ nrm,
                                                            const double* __restrict__ off, long long N, int F, int d,
                                                            double tol, int32_t* __restrict__ first_facet,
                                                            int32_t* __restrict__ far_facet, double* __restrict__ far_dist) {
    __shared__ double sh[CHUNK_DOUBLES];
    const long long j = (long long)blockIdx.x * SET_THREADS + threadIdx.x;
    const bool live = j < N;
    double x[32];
    for (int k = 0; k < d; ++k) x[k] = live ? __ldg(pts + (size_t)j * d + k) : 0.0;
    int first = -1, far = -1;
    double best = -__longlong_as_double(0x7ff0000000000000ll);
    const int per_chunk = CHUNK_DOUBLES / (d + 1);
    for (int f0 = 0; f0 < F; f0 += per_chunk) {
        const int nf = min(per_chunk, F - f0);
        __syncthreads();
        for (int e = threadIdx.x; e < nf * d; e += SET_THREADS) sh[e] = __ldg(nrm + (size_t)f0 * d + e);
        for (int e = threadIdx.x; e < nf; e += SET_THREADS) sh[nf * d + e] = __ldg(off + f0 + e);
        __syncthreads();
        for (int q = 0; q < nf; ++q) {
            const double* n = sh + q * d;
            const double dist = __dsub_rn(np_sum_products([&](int k) { return __dmul_rn(n[k], x[k]); }, d), sh[nf * d + q]);
            if (dist > best) { best = dist; far = f0 + q; }
            if (first < 0 && dist > tol) first = f0 + q;
        }
    }
    if (live) {
        if (first_facet) first_facet[j] = first;
        if (far_facet) far_facet[j] = far;
        if (far_dist) far_dist[j] = best;
    }
}

template <int D>
static int launch_contains(const double* A, const double* b, const int32_t* m_rows, int P, int m, int d, const double* pts,
                           long long N, double abs_tol, int mode, uint8_t* out, cudaStream_t st) {
    const int per = m * (d + 1);
    const int ppc = CHUNK_DOUBLES / per;
    contains_kernel<D><<<blocks_for(N, SET_THREADS), SET_THREADS, 0, st>>>(A, b, m_rows, P, m, d, pts, N, abs_tol, mode, out, ppc);
    count_launch();
    PB_CHECK_CUDA(cudaGetLastError());
    return PB200_OK;
}

template <int D>
static int launch_volume(const double* A, const double* b, const int32_t* m_rows, int P, int m, int d, const double* lo,
                         const double* hi, long long N, const uint64_t* rng, unsigned long long* count, cudaStream_t st) {
    const int sms = sm_count();
    if (!sms) return PB200_ECUDA;
    long long splits = (4LL * sms + P - 1) / P;
    const long long max_splits = (N + SET_THREADS - 1) / SET_THREADS;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    const size_t smem = sizeof(double) * (size_t)m * (d + 1);
    PB_CHECK_CUDA(cudaFuncSetAttribute(volume_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    volume_kernel<D><<<dim3((unsigned)P, (unsigned)splits), SET_THREADS, smem, st>>>(A, b, m_rows, m, d, lo, hi, N, rng, count);
    count_launch();
    PB_CHECK_CUDA(cudaGetLastError());
    return PB200_OK;
}

#define PB_DISPATCH_D(d, FN, ...)                                  \
    switch (d) {                                                   \
        case 1: return FN<1>(__VA_ARGS__);                         \
        case 2: return FN<2>(__VA_ARGS__);                         \
        case 3: return FN<3>(__VA_ARGS__);                         \
        case 4: return FN<4>(__VA_ARGS__);                         \
        case 5: return FN<5>(__VA_ARGS__);                         \
        case 6: return FN<6>(__VA_ARGS__);                         \
        case 7: return FN<7>(__VA_ARGS__);                         \
        case 8: return FN<8>(__VA_ARGS__);                         \
        case 9: return FN<9>(__VA_ARGS__);                         \
        case 10: return FN<10>(__VA_ARGS__);                       \
        case 11: return FN<11>(__VA_ARGS__);                       \
        case 12: return FN<12>(__VA_ARGS__);                       \
        case 13: return FN<13>(__VA_ARGS__);                       \
        case 14: return FN<14>(__VA_ARGS__);                       \
        case 15: return FN<15>(__VA_ARGS__);                       \
        case 16: return FN<16>(__VA_ARGS__);                       \
        default: return FN<0>(__VA_ARGS__);                        \
    }

}  // namespace pb200

using namespace pb200;

extern "C" {

int pb200_contains_batch(const double* A, const double* b, const int32_t* m_rows, int P, int m, int d, const double* points,
                         long long N, double abs_tol, int any_of, uint8_t* out, void* stream) {
    if (P < 0 || N < 0 || !A || !b || !points || !out) return fail(PB200_EINVAL, "pb200_contains_batch: null pointer or negative size");
    if (d < 1 || d > 32) return fail(PB200_EUNSUPPORTED, "contains: need 1 <= d <= 32");
    if (m < 1 || m * (d + 1) > CHUNK_DOUBLES) return fail(PB200_EUNSUPPORTED, "contains: need 1 <= m and m*(d+1) <= 4096");
    if (N == 0) return PB200_OK;
    if (P == 0) {
        if (any_of) PB_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)N, (cudaStream_t)stream));
        return PB200_OK;
    }
    PB_DISPATCH_D(d, launch_contains, A, b, m_rows, P, m, d, points, N, abs_tol, any_of ? 1 : 0, out, (cudaStream_t)stream);
}

int pb200_volume_counts(const double* A, const double* b, const int32_t* m_rows, int P, int m, int d, const double* lo,
                        const double* hi, long long N, const uint64_t* rng_state, unsigned long long* count, void* stream) {
    if (P < 0 || N < 1 || !A || !b || !lo || !hi || !rng_state || !count)
        return fail(PB200_EINVAL, "pb200_volume_counts: null pointer, negative batch or nsamples < 1");
    if (d < 1 || d > 32) return fail(PB200_EUNSUPPORTED, "volume: need 1 <= d <= 32");
    if (m < 1 || (size_t)m * (d + 1) * sizeof(double) > 200 * 1024) return fail(PB200_EUNSUPPORTED, "volume: polytope too large for shared memory");
    if (P == 0) return PB200_OK;
    PB_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(unsigned long long) * (size_t)P, (cudaStream_t)stream));
    PB_DISPATCH_D(d, launch_volume, A, b, m_rows, P, m, d, lo, hi, N, rng_state, count, (cudaStream_t)stream);
}

int pb200_point_facet_sweep(const double* points, const double* normals, const double* offsets, long long N, int F, int d,
                            double tol, int32_t* first_facet, int32_t* far_facet, double* far_dist, void* stream) {
    if (N < 0 || F < 0 || !points || !normals || !offsets) return fail(PB200_EINVAL, "pb200_point_facet_sweep: null pointer or negative size");
    if (d < 1 || d > 32) return fail(PB200_EUNSUPPORTED, "sweep: need 1 <= d <= 32");
    if (N == 0) return PB200_OK;
    sweep_kernel<<<blocks_for(N, SET_THREADS), SET_THREADS, 0, (cudaStream_t)stream>>>(points, normals, offsets, N, F, d, tol,
                                                                                    first_facet, far_facet, far_dist);
    count_launch();
    PB_CHECK_CUDA(cudaGetLastError());
    return PB200_OK;
}

}  // extern "C"

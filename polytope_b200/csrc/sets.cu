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

// Rows of one polytope against TWO points held in registers.  The rows sit in
// shared memory with an even leading dimension DP (a zero pad column when d is
// odd) so that one 128-bit broadcast load feeds four DFMAs (2 columns x 2
// points); four rows are in flight per step (independent fma chains).
// contains() tests  A x - b < abs_tol,  volume() tests  A x - b < 0.
// The sums are acc = fma(A[i][k], x[k], acc) in k order -- what numpy's dgemm
// computes; the pad column adds fma(0, 0, acc) = acc.
template <int DP>
__device__ __forceinline__ void inside_rows_x2(const double* __restrict__ As, const double* __restrict__ bs, int mm,
                                               const double (&x0)[DP], const double (&x1)[DP], double tol, bool& in0, bool& in1) {
    constexpr int R = 4;
    bool ok0 = true, ok1 = true;
    int i = 0;
    for (; i + R <= mm && (ok0 || ok1); i += R) {
        double a0[R], a1[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { a0[r] = 0.0; a1[r] = 0.0; }
#pragma unroll
        for (int k = 0; k < DP; k += 2)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double2 g = *reinterpret_cast<const double2*>(As + (i + r) * DP + k);
                a0[r] = fma(g.x, x0[k], a0[r]);
                a1[r] = fma(g.x, x1[k], a1[r]);
                a0[r] = fma(g.y, x0[k + 1], a0[r]);
                a1[r] = fma(g.y, x1[k + 1], a1[r]);
            }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double bi = bs[i + r];
            ok0 = ok0 && (__dsub_rn(a0[r], bi) < tol);
            ok1 = ok1 && (__dsub_rn(a1[r], bi) < tol);
        }
    }
    for (; i < mm && (ok0 || ok1); ++i) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int k = 0; k < DP; k += 2) {
            const double2 g = *reinterpret_cast<const double2*>(As + i * DP + k);
            a0 = fma(g.x, x0[k], a0);
            a1 = fma(g.x, x1[k], a1);
            a0 = fma(g.y, x0[k + 1], a0);
            a1 = fma(g.y, x1[k + 1], a1);
        }
        const double bi = bs[i];
        ok0 = ok0 && (__dsub_rn(a0, bi) < tol);
        ok1 = ok1 && (__dsub_rn(a1, bi) < tol);
    }
    in0 = ok0;
    in1 = ok1;
}

// ------------------------------------------------------------------------
// contains: P polytopes x N points.  One thread per PAIR of points of a
// 512-point tile: adjacent points (2t, 2t+1) fetched with one 128-bit streaming
// load per coordinate when N is even (VEC), else (t, t + 256) with 64-bit loads.
// The polytopes are walked in shared-memory chunks (staged once when they all
// fit).  mode 0: out[P][N]; mode 1: out[N] = OR over p.
// ------------------------------------------------------------------------
constexpr int SET_THREADS = 256;
constexpr int CHUNK_DOUBLES = 4096;      // 32 KB of staged rows per chunk

template <int D, bool VEC>
__global__ void __launch_bounds__(SET_THREADS) contains_kernel(const double* __restrict__ A, const double* __restrict__ b,
                                                               const int32_t* __restrict__ m_rows, int P, int m, int d,
                                                               const double* __restrict__ pts, long long N, double abs_tol,
                                                               int mode, uint8_t* __restrict__ out, int polys_per_chunk) {
    __shared__ __align__(16) double sh[CHUNK_DOUBLES];
    __shared__ int sh_rows[CHUNK_DOUBLES / 2];
    constexpr int DP = (D + 1) & ~1;
    const int per = m * DP;
    const long long ntiles = (N + 2 * SET_THREADS - 1) / (2 * SET_THREADS);
    const bool single_chunk = P <= polys_per_chunk;     // stage once, then stream tiles of points
    bool staged = false;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long j0 = tile * (2 * SET_THREADS) + (VEC ? 2 * threadIdx.x : threadIdx.x);
        const long long j1 = j0 + (VEC ? 1 : SET_THREADS);
        const bool live0 = j0 < N, live1 = j1 < N;
        double x0[DP], x1[DP];
#pragma unroll
        for (int k = 0; k < DP; ++k) {
            if (VEC) {
                // N even and j0 even: (j0, j0 + 1) are both live or both dead, 16-byte aligned
                const double2 v = (k < D && live0) ? __ldcs(reinterpret_cast<const double2*>(pts + (size_t)k * N + j0))
                                                   : make_double2(0.0, 0.0);
                x0[k] = v.x;
                x1[k] = v.y;
            } else {
                x0[k] = (k < D && live0) ? __ldcs(pts + (size_t)k * N + j0) : 0.0;
                x1[k] = (k < D && live1) ? __ldcs(pts + (size_t)k * N + j1) : 0.0;
            }
        }
        bool any0 = false, any1 = false;
        for (int p0 = 0; p0 < P; p0 += polys_per_chunk) {
            const int np = min(polys_per_chunk, P - p0);
            if (!(single_chunk && staged)) {
                __syncthreads();
                // stage [np][m][DP] rows (pad column zero) then [np][m] right-hand sides
                for (int e = threadIdx.x; e < np * per; e += SET_THREADS) {
                    const int row = e / DP, k = e - row * DP;
                    sh[e] = k < D ? __ldg(A + ((size_t)p0 * m + row) * D + k) : 0.0;
                }
                for (int e = threadIdx.x; e < np * m; e += SET_THREADS) sh[np * per + e] = __ldg(b + (size_t)p0 * m + e);
                for (int e = threadIdx.x; e < np; e += SET_THREADS) sh_rows[e] = m_rows ? min(max(m_rows[p0 + e], 0), m) : m;
                __syncthreads();
                staged = true;
            }
            for (int q = 0; q < np; ++q) {
                if (mode == 1 && any0 && any1) break;
                bool in0, in1;
                inside_rows_x2<DP>(sh + (size_t)q * per, sh + (size_t)np * per + q * m, sh_rows[q], x0, x1, abs_tol, in0, in1);
                if (mode == 0) {
                    uint8_t* o = out + (size_t)(p0 + q) * N;
                    if (VEC) {
                        if (live0) __stcs(reinterpret_cast<uchar2*>(o + j0), make_uchar2(in0 ? 1 : 0, in1 ? 1 : 0));
                    } else {
                        if (live0) o[j0] = in0 ? 1 : 0;
                        if (live1) o[j1] = in1 ? 1 : 0;
                    }
                } else {
                    any0 = any0 || in0;
                    any1 = any1 || in1;
                }
            }
        }
        if (mode == 1) {
            if (VEC) {
                if (live0) __stcs(reinterpret_cast<uchar2*>(out + j0), make_uchar2(any0 ? 1 : 0, any1 ? 1 : 0));
            } else {
                if (live0) out[j0] = any0 ? 1 : 0;
                if (live1) out[j1] = any1 ? 1 : 0;
            }
        }
    }
}

// ------------------------------------------------------------------------
// volume: count[p] += #{ j < N : A (l + u_j * (hi - lo)) - b < 0 }, with
// u = default_rng(seed).random((d, N)) regenerated from the PCG64 state
// (polytope.py:1583-1591).  grid = (P, S): S CTAs share the samples of one
// polytope; every thread owns a contiguous run of samples, so after one
// O(log) jump per thread each draw is a single 128-bit multiply-add.
// ------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(SET_THREADS) volume_kernel(const double* __restrict__ A, const double* __restrict__ b,
                                                             const int32_t* __restrict__ m_rows, int m, int d,
                                                             const double* __restrict__ lo, const double* __restrict__ hi,
                                                             long long N, const uint64_t* __restrict__ rng,
                                                             unsigned long long* __restrict__ count) {
    extern __shared__ __align__(16) double sh[];       // [m][DP] rows | [m] b
    constexpr int DP = (D + 1) & ~1;
    __shared__ unsigned long long sh_cnt;
    __shared__ u128 sh_coef[2];          // jump by N: (mult, plus)
    const int p = blockIdx.x;
    const int mm = m_rows ? min(max(m_rows[p], 0), m) : m;
    for (int e = threadIdx.x; e < m * DP; e += SET_THREADS) {
        const int row = e / DP, k = e - row * DP;
        sh[e] = k < D ? __ldg(A + ((size_t)p * m + row) * D + k) : 0.0;
    }
    for (int e = threadIdx.x; e < m; e += SET_THREADS) sh[m * DP + e] = __ldg(b + (size_t)p * m + e);
    Pcg g;
    g.state = ((u128)rng[4 * p + 0] << 64) | rng[4 * p + 1];
    g.inc = ((u128)rng[4 * p + 2] << 64) | rng[4 * p + 3];
    if (threadIdx.x == 0) {
        sh_cnt = 0;
        pcg_jump_coeffs(g.inc, (unsigned long long)N, sh_coef[0], sh_coef[1]);
    }
    __syncthreads();
    // samples [j_begin, j_end) of this thread
    const long long nthreads = (long long)gridDim.y * SET_THREADS;
    const long long per_thread = (N + nthreads - 1) / nthreads;
    const long long j_begin = ((long long)blockIdx.y * SET_THREADS + threadIdx.x) * per_thread;
    const long long j_end = j_begin + per_thread < N ? j_begin + per_thread : N;
    double l[DP], w[DP];
    u128 st[DP];
    {
        // stream position of coordinate k for sample j is k*N + j
        u128 am, ap;
        pcg_jump_coeffs(g.inc, (unsigned long long)(j_begin < N ? j_begin : 0), am, ap);
        u128 s = g.state * am + ap;
#pragma unroll
        for (int k = 0; k < DP; ++k) {
            st[k] = s;
            s = s * sh_coef[0] + sh_coef[1];
            l[k] = k < D ? lo[(size_t)p * d + k] : 0.0;
            w[k] = k < D ? __dsub_rn(hi[(size_t)p * d + k], l[k]) : 0.0;
        }
    }
    unsigned mine = 0;
    for (long long j = j_begin; j < j_end; j += 2) {
        double x0[DP], x1[DP];
#pragma unroll
        for (int k = 0; k < DP; ++k) {
            if (k < D) {
                const double u0 = pcg_next_double(st[k], g.inc);
                const double u1 = pcg_next_double(st[k], g.inc);
                x0[k] = __dadd_rn(l[k], __dmul_rn(u0, w[k]));
                x1[k] = __dadd_rn(l[k], __dmul_rn(u1, w[k]));
            } else {
                x0[k] = 0.0;
                x1[k] = 0.0;
            }
        }
        bool in0, in1;
        inside_rows_x2<DP>(sh, sh + m * DP, mm, x0, x1, 0.0, in0, in1);
        mine += (in0 ? 1u : 0u) + ((in1 && j + 1 < j_end) ? 1u : 0u);
    }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&sh_cnt, (unsigned long long)mine);
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
    const int per = m * (((D + 1) & ~1) + 1);
    const int ppc = CHUNK_DOUBLES / per;
    // persistent-ish grid: a whole number of waves, every CTA streams several 512-point tiles
    const int sms = sm_count();
    if (!sms) return PB200_ECUDA;
    long long grid = (long long)sms * 5 * 4;
    const long long ntiles = (N + 2 * SET_THREADS - 1) / (2 * SET_THREADS);
    if (ntiles < grid) grid = ntiles;
    const bool vec = (N % 2 == 0) && ((uintptr_t)pts % 16 == 0) && ((uintptr_t)out % 2 == 0);
    if (vec) contains_kernel<D, true><<<(unsigned)grid, SET_THREADS, 0, st>>>(A, b, m_rows, P, m, d, pts, N, abs_tol, mode, out, ppc);
    else contains_kernel<D, false><<<(unsigned)grid, SET_THREADS, 0, st>>>(A, b, m_rows, P, m, d, pts, N, abs_tol, mode, out, ppc);
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
    const size_t smem = sizeof(double) * (size_t)m * (((D + 1) & ~1) + 1);
    PB_CHECK_CUDA(cudaFuncSetAttribute(volume_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    volume_kernel<D><<<dim3((unsigned)P, (unsigned)splits), SET_THREADS, smem, st>>>(A, b, m_rows, m, d, lo, hi, N, rng, count);
    count_launch();
    PB_CHECK_CUDA(cudaGetLastError());
    return PB200_OK;
}

#define PB_DISPATCH_D(d, FN, ...)                                  \
    switch (d) {                                                   \
        case 1: return FN<1>(__VA_ARGS__);                      \
        case 2: return FN<2>(__VA_ARGS__);                      \
        case 3: return FN<3>(__VA_ARGS__);                      \
        case 4: return FN<4>(__VA_ARGS__);                      \
        case 5: return FN<5>(__VA_ARGS__);                      \
        case 6: return FN<6>(__VA_ARGS__);                      \
        case 7: return FN<7>(__VA_ARGS__);                      \
        case 8: return FN<8>(__VA_ARGS__);                      \
        case 9: return FN<9>(__VA_ARGS__);                      \
        case 10: return FN<10>(__VA_ARGS__);                      \
        case 11: return FN<11>(__VA_ARGS__);                      \
        case 12: return FN<12>(__VA_ARGS__);                      \
        case 13: return FN<13>(__VA_ARGS__);                      \
        case 14: return FN<14>(__VA_ARGS__);                      \
        case 15: return FN<15>(__VA_ARGS__);                      \
        case 16: return FN<16>(__VA_ARGS__);                      \
        case 17: return FN<17>(__VA_ARGS__);                      \
        case 18: return FN<18>(__VA_ARGS__);                      \
        case 19: return FN<19>(__VA_ARGS__);                      \
        case 20: return FN<20>(__VA_ARGS__);                      \
        case 21: return FN<21>(__VA_ARGS__);                      \
        case 22: return FN<22>(__VA_ARGS__);                      \
        case 23: return FN<23>(__VA_ARGS__);                      \
        case 24: return FN<24>(__VA_ARGS__);                      \
        case 25: return FN<25>(__VA_ARGS__);                      \
        case 26: return FN<26>(__VA_ARGS__);                      \
        case 27: return FN<27>(__VA_ARGS__);                      \
        case 28: return FN<28>(__VA_ARGS__);                      \
        case 29: return FN<29>(__VA_ARGS__);                      \
        case 30: return FN<30>(__VA_ARGS__);                      \
        case 31: return FN<31>(__VA_ARGS__);                      \
        default: return FN<32>(__VA_ARGS__);                       \
    }

}  // namespace pb200

using namespace pb200;

extern "C" {

int pb200_contains_batch(const double* A, const double* b, const int32_t* m_rows, int P, int m, int d, const double* points,
                         long long N, double abs_tol, int any_of, uint8_t* out, void* stream) {
    if (N == 0) return PB200_OK;
    if (P == 0 && out) {          // no polytopes: nothing contains anything
        if (any_of) PB_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)N, (cudaStream_t)stream));
        return PB200_OK;
    }
    if (P < 0 || N < 0 || !A || !b || !points || !out) return fail(PB200_EINVAL, "pb200_contains_batch: null pointer or negative size");
    if (d < 1 || d > 32) return fail(PB200_EUNSUPPORTED, "contains: need 1 <= d <= 32");
    if (m < 1 || m * (d + 3) > CHUNK_DOUBLES) return fail(PB200_EUNSUPPORTED, "contains: need 1 <= m and m*(d+3) <= 4096");
    PB_DISPATCH_D(d, launch_contains, A, b, m_rows, P, m, d, points, N, abs_tol, any_of ? 1 : 0, out, (cudaStream_t)stream);
}

int pb200_volume_counts(const double* A, const double* b, const int32_t* m_rows, int P, int m, int d, const double* lo,
                        const double* hi, long long N, const uint64_t* rng_state, unsigned long long* count, void* stream) {
    if (P == 0) return PB200_OK;
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
    if (N == 0) return PB200_OK;
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

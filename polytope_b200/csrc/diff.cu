// polytope_b200: set difference poly \ region as a device-resident search (sm_100a).
//
// Replaces: region_diff(poly, reg), polytope/polytope.py:2117-2282 -- the routine
// behind mldivide / diff / is_subset / == / union(check_convex) / is_convex.
// The reference walks a depth-first tree over "which facet of which region cell
// cuts the current piece", solving one Chebyshev LP per node through lpsolve
// (300-800 LPs per call, 87-94 % of the time).
//
// Here one warp owns one (poly, region) problem and runs the whole search on
// the device: the node stack (`counter`, `INDICES`, `level` of the reference,
// with numpy's negative-index wrap-around reproduced) lives in shared memory,
// every node's LP is staged from the stacked row table straight into the warp
// solver (lp_warp*.cuh), and the pieces of the difference are appended to an
// output pool.  Pieces the reference passes through reduce() (:2276) are only
// flagged: the caller runs pb200_reduce_batch over all of them at once.
// A batch of problems (cfg3: 50 000 cells minus one polytope) fills the GPU; the
// warps pull problems from a global counter because search depths differ.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "lp_warp.cuh"
#include "lp_warp_small.cuh"
#include "staging.cuh"

namespace pb200 {

constexpr int DIFF_WPC = 4;
constexpr int DIFF_MAX_REG = 64;       // region cells per problem
constexpr int DIFF_MAX_ROWS = 128;     // rows of one LP / one piece
constexpr int DIFF_MAX_ADDED = 2048;   // rows of region cells that are not rows of poly (M)

struct DiffArgs {
    const double *PA, *Pb;     // [T][mp][d], [T][mp]
    const int32_t* p_rows;     // nullable [T]
    int T, mp, d;
    const double *RA, *Rb;     // [T or 1][Nr][mr][d], [..][Nr][mr]
    const int32_t* r_rows;     // nullable [T or 1][Nr]
    const int32_t* n_reg;      // nullable [T or 1]
    int reg_shared;            // 1: one region for all problems
    int Nr, mr;
    double abs_tol, intersect_tol;
    int max_steps;
    double* pieceA;            // [piece_cap][piece_m][d] raw stacked rows of each piece
    double* pieceb;            // [piece_cap][piece_m]
    int32_t* piece_rows;       // [piece_cap]
    int32_t* piece_reduce;     // [piece_cap] 1: the reference reduces this piece
    int32_t* piece_owner;      // [piece_cap] problem index
    int32_t* piece_seq;        // [piece_cap] order within the problem
    long long piece_cap;
    int piece_m;
    unsigned long long* piece_used;
    int32_t* status;           // [T]
    int32_t* n_pieces;         // [T]
    int32_t* n_lp;             // [T]
    int* next_problem;
};

enum : int { DS_PIECES = 0, DS_UNTOUCHED = 1, DS_COVERED = 2, DS_POOL_FULL = 3, DS_INDEX_ERROR = 4, DS_STEP_LIMIT = 5 };

// per-warp search state in shared memory
struct DiffState {
    int indices[DIFF_MAX_ROWS];
    int n_idx;
    int counter[DIFF_MAX_REG], mi[DIFF_MAX_REG], beg[DIFF_MAX_REG], order[DIFF_MAX_REG];
    double rc[DIFF_MAX_REG];
    unsigned char src_cell[DIFF_MAX_ADDED], src_row[DIFF_MAX_ADDED];   // added rows m .. m+M-1
    int m, M, N;
};

struct DiffProblem {
    const double *PA, *Pb, *RA, *Rb;
    const int32_t* r_rows;
    int mr, d;
};

// stacked row table of the reference: [poly rows | added cell rows | their negations]
// (polytope.py:2180-2196).  Returns false for an index numpy would reject.
__device__ __forceinline__ bool resolve_row(const DiffState& s, const DiffProblem& q, int idx, const double*& row, const double*& bb,
                                            double& sign) {
    const int total = s.m + 2 * s.M;
    if (idx < 0) idx += total;                  // numpy wraps negative fancy indices
    if (idx < 0 || idx >= total) return false;
    sign = 1.0;
    if (idx >= s.m + s.M) { idx -= s.M; sign = -1.0; }
    if (idx < s.m) {
        row = q.PA + (size_t)idx * q.d;
        bb = q.Pb + idx;
    } else {
        const int c = s.src_cell[idx - s.m], r = s.src_row[idx - s.m];
        row = q.RA + ((size_t)c * q.mr + r) * q.d;
        bb = q.Rb + (size_t)c * q.mr + r;
    }
    return true;
}

template <int RPL, bool SMALL>
struct Solver;
template <int RPL>
struct Solver<RPL, true> {
    typedef SmallScratch Scratch;
    static __host__ __device__ int doubles(int n) { return lps_scratch_doubles(RPL); }
    static __device__ Scratch carve(double* base, int n) { return lps_carve(base, RPL); }
    static __device__ void solve(const Scratch& w, int m, int n, double c, const double (&h)[RPL], int& status, double& x) {
        const SmallResult r = lp_solve_small<RPL>(w, m, n, c, h);
        status = r.status;
        x = r.x;
    }
};
template <int RPL>
struct Solver<RPL, false> {
    typedef WarpScratch Scratch;
    static __host__ __device__ int doubles(int n) { return lp_scratch_doubles(RPL, n); }
    static __device__ Scratch carve(double* base, int n) { return lp_carve(base, RPL, n); }
    static __device__ void solve(const Scratch& w, int m, int n, double c, const double (&h)[RPL], int& status, double& x) {
        const LpResult r = lp_solve_warp<RPL>(w, m, n, c, h);
        status = r.status;
        x = r.x;
    }
};

// cheby_ball(Polytope(A[idx], B[idx]))[0] for idx = list[0..cnt) followed by the
// `extra_cnt` consecutive table rows extra0.. ; -1 signals an index error.
template <int RPL, bool SMALL>
__device__ __noinline__ double cheby_of_rows(const typename Solver<RPL, SMALL>::Scratch& w, const DiffState& s, const DiffProblem& q,
                                const int* list, int cnt, int extra0, int extra_cnt, int lane, int& n_lp, bool& index_error) {
    const int d = q.d, total = cnt + extra_cnt;
    if (total > 32 * RPL) { index_error = true; return 0.0; }      // outside the kernel envelope
    zero_G<RPL>(w, lane);
    bool bad = false;
    for (int e = lane; e < total * d; e += 32) {
        const int i = e / d, j = e - i * d;
        const double *row, *bb;
        double sign;
        if (!resolve_row(s, q, i < cnt ? list[i] : extra0 + (i - cnt), row, bb, sign)) { bad = true; continue; }
        w.G[j * w.MP + i] = sign * __ldg(row + j);
    }
    double h[RPL];
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int i = lane + 32 * r;
        h[r] = 0.0;
        if (i < total) {
            const double *row, *bb;
            double sign;
            if (resolve_row(s, q, i < cnt ? list[i] : extra0 + (i - cnt), row, bb, sign)) h[r] = sign * __ldg(bb);
            else bad = true;
        }
    }
    __syncwarp();
    if (__any_sync(FULL_MASK, bad)) { index_error = true; return 0.0; }
    renormalize_rows<RPL>(w, total, d, lane, h);          // Polytope(...) constructor, polytope.py:128-138
    append_norm_column<RPL>(w, total, d, lane);          // cheby_ball, polytope.py:1283-1287
    int status;
    double x;
    Solver<RPL, SMALL>::solve(w, total, d + 1, lane == d ? -1.0 : 0.0, h, status, x);
    ++n_lp;
    const double r = __shfl_sync(FULL_MASK, x, d);
    __syncwarp();
    // cheby_ball: status != 0 or r < 0 -> (0, None), polytope.py:1289-1300
    return (status == ST_OPTIMAL && !(r < 0.0)) ? r : 0.0;
}

template <int RPL, bool SMALL>
__global__ void __launch_bounds__(DIFF_WPC * 32) diff_kernel(const DiffArgs a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ DiffState states[DIFF_WPC];
    const int lane = threadIdx.x & 31, wib = __reduce_max_sync(FULL_MASK, threadIdx.x >> 5);   // provably uniform
    const int d = a.d, n = d + 1;
    typedef Solver<RPL, SMALL> S;
    const typename S::Scratch w = S::carve(smem + (size_t)wib * S::doubles(n), n);
    DiffState& s = states[wib];

    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(a.next_problem, 1);
        t = __shfl_sync(FULL_MASK, t, 0);
        if (t >= a.T) break;
        const int rt = a.reg_shared ? 0 : t;
        DiffProblem q;
        q.PA = a.PA + (size_t)t * a.mp * d;
        q.Pb = a.Pb + (size_t)t * a.mp;
        q.RA = a.RA + (size_t)rt * a.Nr * a.mr * d;
        q.Rb = a.Rb + (size_t)rt * a.Nr * a.mr;
        q.r_rows = a.r_rows ? a.r_rows + (size_t)rt * a.Nr : nullptr;
        q.mr = a.mr;
        q.d = d;
        const int m = a.p_rows ? min(max(a.p_rows[t], 0), a.mp) : a.mp;
        const int ncell = a.n_reg ? min(max(a.n_reg[rt], 0), a.Nr) : a.Nr;
        int n_lp = 0, npieces = 0, status = DS_PIECES;
        bool index_error = false;
        auto cell_rows = [&](int c) { return q.r_rows ? min(max(q.r_rows[c], 0), a.mr) : a.mr; };
        __syncwarp();
        if (lane == 0) { s.m = m; s.M = 0; s.N = 0; }
        __syncwarp();

        // ---- which cells intersect poly (polytope.py:2146-2156): stage [poly; cell] directly ----
        int N = 0;
        for (int c = 0; c < ncell; ++c) {
            const int mc = cell_rows(c);
            // temporary table: added rows = all rows of cell c
            if (lane == 0) s.M = mc;
            for (int r = lane; r < mc; r += 32) { s.src_cell[r] = (unsigned char)c; s.src_row[r] = (unsigned char)r; }
            for (int i = lane; i < m; i += 32) s.indices[i] = i;
            __syncwarp();
            const double rc = cheby_of_rows<RPL, SMALL>(w, s, q, s.indices, m, m, mc, lane, n_lp, index_error);
            if (lane == 0) s.rc[c] = rc;
            N += rc >= a.intersect_tol ? 1 : 0;
            __syncwarp();
        }
        if (N == 0) status = DS_UNTOUCHED;
        if (status == DS_PIECES) {
            // ---- argsort(-Rc), stable (numpy sorts short arrays by insertion) ----
            if (lane == 0) {
                for (int c = 0; c < ncell; ++c) s.order[c] = c;
                for (int i = 1; i < ncell; ++i) {
                    const int v = s.order[i];
                    int j = i - 1;
                    while (j >= 0 && s.rc[s.order[j]] < s.rc[v]) { s.order[j + 1] = s.order[j]; --j; }
                    s.order[j + 1] = v;
                }
                s.M = 0;
                s.N = N;
            }
            __syncwarp();
            // ---- rows of the intersecting cells that are not rows of poly (:2166-2182) ----
            bool covered = false;
            int M = 0;
            for (int ii = 0; ii < N; ++ii) {
                const int c = s.order[ii];
                const int mc = cell_rows(c);
                int mi = 0;
                // is_fulldim(cell): implied by Rc well above the threshold, otherwise its own LP
                bool fulldim = s.rc[c] > 1e-7 + 1e-10;
                if (!fulldim) {
                    // LP over the cell alone: table trick -- list = added rows of a temporary table
                    const int keepM = M;
                    if (lane == 0) s.M = keepM + mc;
                    for (int r = lane; r < mc; r += 32) {
                        s.src_cell[keepM + r] = (unsigned char)c;
                        s.src_row[keepM + r] = (unsigned char)r;
                    }
                    __syncwarp();
                    const double r0 = cheby_of_rows<RPL, SMALL>(w, s, q, s.indices, 0, m + keepM, mc, lane, n_lp, index_error);
                    fulldim = r0 > 1e-7;
                    if (lane == 0) s.M = keepM;
                    __syncwarp();
                }
                if (fulldim) {
                    for (int j = 0; j < mc; ++j) {
                        const double* row = q.RA + ((size_t)c * a.mr + j) * d;
                        const double bj = q.Rb[(size_t)c * a.mr + j];
                        bool differs = true;
                        for (int k = lane; k < m; k += 32) {
                            double sum = 0.0;
                            for (int cc = 0; cc < d; ++cc) sum += fabs(q.PA[(size_t)k * d + cc] - row[cc]);
                            sum += fabs(q.Pb[k] - bj);
                            differs = differs && (sum >= a.abs_tol);
                        }
                        if (__all_sync(FULL_MASK, differs)) {
                            if (M + 1 > DIFF_MAX_ADDED) { index_error = true; break; }
                            if (lane == 0) { s.src_cell[M] = (unsigned char)c; s.src_row[M] = (unsigned char)j; }
                            ++M;
                            ++mi;
                        }
                    }
                }
                if (lane == 0) { s.mi[ii] = mi; s.counter[ii] = 0; }
                if (mi == 0) covered = true;
                __syncwarp();
            }
            if (lane == 0) {
                s.M = M;
                int acc = m;
                for (int ii = 0; ii < N; ++ii) { s.beg[ii] = acc; acc += s.mi[ii]; }
                s.n_idx = m;
            }
            for (int i = lane; i < m; i += 32) s.indices[i] = i;
            __syncwarp();
            if (covered) status = DS_COVERED;
        }
        long long pool_base = -1;
        if (status == DS_PIECES && !index_error) {
            const int M = s.M;
            // python-style index into counter / mi / beg (level may be -1, :2218-2224)
            auto py = [&](int level) { return level < 0 ? level + N : level; };
            auto emit = [&](bool reduce_it) {
                // union(res, piece, False): append (polytope.py:2217, :2276)
                unsigned long long slot = 0;
                if (lane == 0) slot = atomicAdd(a.piece_used, 1ull);
                slot = __shfl_sync(FULL_MASK, slot, 0);
                if (s.n_idx > a.piece_m) { status = DS_INDEX_ERROR; ++npieces; return; }     // outside the envelope
                if ((long long)slot >= a.piece_cap) { status = DS_POOL_FULL; ++npieces; return; }
                if (pool_base < 0) pool_base = (long long)slot;
                double* oA = a.pieceA + (size_t)slot * a.piece_m * d;
                double* ob = a.pieceb + (size_t)slot * a.piece_m;
                for (int e = lane; e < a.piece_m * d; e += 32) {
                    const int i = e / d, j = e - i * d;
                    double v = 0.0;
                    if (i < s.n_idx) {
                        const double *row, *bb;
                        double sign;
                        if (resolve_row(s, q, s.indices[i], row, bb, sign)) v = sign * row[j];
                    }
                    oA[e] = v;
                }
                for (int i = lane; i < a.piece_m; i += 32) {
                    double v = 0.0;
                    if (i < s.n_idx) {
                        const double *row, *bb;
                        double sign;
                        if (resolve_row(s, q, s.indices[i], row, bb, sign)) v = sign * bb[0];
                    }
                    ob[i] = v;
                }
                if (lane == 0) {
                    a.piece_rows[slot] = s.n_idx;
                    a.piece_reduce[slot] = reduce_it ? 1 : 0;
                    a.piece_owner[slot] = t;
                    a.piece_seq[slot] = npieces;
                }
                ++npieces;
            };
            auto sum_counter = [&]() { int sc = 0; for (int i = 0; i < N; ++i) sc += s.counter[i]; return sc; };
            auto nonzero_count = [&]() { int sc = 0; for (int i = 0; i < N; ++i) sc += s.counter[i] != 0; return sc; };
            int level = 0, steps = 0;
            bool done = false;
            while (level != -1 && !done && !index_error && status == DS_PIECES) {
                if (++steps > a.max_steps) { status = DS_STEP_LIMIT; break; }
                __syncwarp();
                if (s.counter[py(level)] == 0) {
                    double R = 0.0;
                    for (int j = level; j < N; ++j) {
                        R = cheby_of_rows<RPL, SMALL>(w, s, q, s.indices, s.n_idx, s.beg[j], s.mi[j], lane, n_lp, index_error);
                        if (index_error) break;
                        if (R > a.abs_tol) {
                            level = j;
                            if (lane == 0) {
                                s.counter[level] = 1;
                                if (s.n_idx < DIFF_MAX_ROWS) s.indices[s.n_idx] = s.beg[level] + M;
                                s.n_idx += 1;
                            }
                            __syncwarp();
                            break;
                        }
                    }
                    if (index_error || s.n_idx > DIFF_MAX_ROWS) { index_error = true; break; }
                    if (R < a.abs_tol) {
                        level = level - 1;
                        emit(false);
                        const int nz = nonzero_count();
                        for (int jj = nz - 1; jj >= 0; --jj) {
                            __syncwarp();
                            const int L = py(level);
                            if (L < 0 || L >= N) { index_error = true; break; }
                            if (s.counter[L] <= s.mi[L]) {
                                if (lane == 0) {
                                    s.indices[s.n_idx - 1] -= M;
                                    if (s.n_idx < DIFF_MAX_ROWS) s.indices[s.n_idx] = s.beg[L] + s.counter[L] + M;
                                    s.n_idx += 1;
                                }
                                __syncwarp();
                                break;
                            } else {
                                __syncwarp();
                                if (lane == 0) s.counter[L] = 0;
                                __syncwarp();
                                if (lane == 0) s.n_idx = min(s.n_idx, m + sum_counter());
                                __syncwarp();
                                if (level == -1) { done = true; break; }
                            }
                        }
                        if (done || index_error) break;
                    }
                } else {
                    // the non-zero entries of counter, last first (:2235-2262)
                    int nzl[DIFF_MAX_REG];
                    int nz = 0;
                    for (int i = 0; i < N; ++i)
                        if (s.counter[i] != 0) nzl[nz++] = i;
                    for (int jj = nz - 1; jj >= 0; --jj) {
                        level = nzl[jj];
                        __syncwarp();
                        const int cnt = s.counter[level] + 1;
                        __syncwarp();
                        if (lane == 0) s.counter[level] = cnt;
                        __syncwarp();
                        if (cnt <= s.mi[level]) {
                            if (lane == 0) {
                                s.indices[s.n_idx - 1] -= M;
                                if (s.n_idx < DIFF_MAX_ROWS) s.indices[s.n_idx] = s.beg[level] + cnt + M - 1;
                                s.n_idx += 1;
                            }
                            __syncwarp();
                            break;
                        } else {
                            if (lane == 0) s.counter[level] = 0;
                            __syncwarp();
                            if (lane == 0) s.n_idx = min(s.n_idx, m + sum_counter());
                            __syncwarp();
                            level = level - 1;
                            if (level == -1) { done = true; break; }
                        }
                    }
                    if (done) break;
                }
                __syncwarp();
                if (s.n_idx > DIFF_MAX_ROWS) { index_error = true; break; }
                const double rc = cheby_of_rows<RPL, SMALL>(w, s, q, s.indices, s.n_idx, 0, 0, lane, n_lp, index_error);
                if (index_error) break;
                if (rc > a.abs_tol) {
                    if (level == N - 1) emit(true);
                    else level = level + 1;
                }
            }
        }
        if (index_error && status == DS_PIECES) status = DS_INDEX_ERROR;
        if (lane == 0) {
            a.status[t] = status;
            a.n_pieces[t] = npieces;
            a.n_lp[t] = n_lp;
        }
        __syncwarp();
    }
}

template <int RPL, bool SMALL>
static int launch_diff(const DiffArgs& a, cudaStream_t st) {
    const int n = a.d + 1;
    const size_t smem = (size_t)DIFF_WPC * Solver<RPL, SMALL>::doubles(n) * sizeof(double);
    if (smem + sizeof(DiffState) * DIFF_WPC > 227 * 1024) return fail(PB200_EUNSUPPORTED, "region_diff: LP too large for shared memory");
    auto kern = diff_kernel<RPL, SMALL>;
    PB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int sms = sm_count();
    if (!sms) return PB200_ECUDA;
    int per_sm = 0;
    PB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DIFF_WPC * 32, smem));
    if (per_sm < 1) return fail(PB200_EUNSUPPORTED, "region_diff kernel does not fit on an SM");
    long long grid = (long long)sms * per_sm;
    const long long need = ((long long)a.T + DIFF_WPC - 1) / DIFF_WPC;
    if (need < grid) grid = need;
    kern<<<(unsigned)grid, DIFF_WPC * 32, smem, st>>>(a);
    count_launch();
    PB_CHECK_CUDA(cudaGetLastError());
    return PB200_OK;
}

}  // namespace pb200

using namespace pb200;

extern "C" {

int pb200_region_diff_batch(const double* PA, const double* Pb, const int32_t* p_rows, int T, int mp, int d, const double* RA,
                            const double* Rb, const int32_t* r_rows, const int32_t* n_reg, int reg_shared, int Nr, int mr,
                            double abs_tol, double intersect_tol, double* piece_A, double* piece_b, int32_t* piece_rows,
                            int32_t* piece_reduce, int32_t* piece_owner, int32_t* piece_seq, long long piece_cap, int piece_m,
                            long long* pieces_used, int32_t* status, int32_t* n_pieces, int32_t* n_lp, int* work_counter,
                            void* stream) {
    if (T == 0) return PB200_OK;
    if (T < 0 || !PA || !Pb || !RA || !Rb || !piece_A || !piece_b || !piece_rows || !piece_reduce || !piece_owner || !piece_seq ||
        !pieces_used || !status || !n_pieces || !n_lp || !work_counter)
        return fail(PB200_EINVAL, "pb200_region_diff_batch: null pointer or negative batch");
    if (d < 1 || d + 1 > LP_MAX_N) return fail(PB200_EUNSUPPORTED, "region_diff: need 1 <= d <= 31");
    if (Nr < 1 || Nr > DIFF_MAX_REG) return fail(PB200_EUNSUPPORTED, "region_diff: need 1 <= cells per region <= 64");
    if (mp < 1 || mr < 1 || mp + mr > DIFF_MAX_ROWS || mr > 255 || Nr * mr > DIFF_MAX_ADDED)
        return fail(PB200_EUNSUPPORTED, "region_diff: need mp + mr <= 128 rows, mr <= 255, Nr * mr <= 2048");
    if (piece_m < 1) return fail(PB200_EINVAL, "region_diff: piece_m must be >= 1");
    if (T == 0) return PB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    PB_CHECK_CUDA(cudaMemsetAsync(work_counter, 0, sizeof(int), st));
    PB_CHECK_CUDA(cudaMemsetAsync(pieces_used, 0, sizeof(long long), st));
    DiffArgs a;
    a.PA = PA; a.Pb = Pb; a.p_rows = p_rows; a.T = T; a.mp = mp; a.d = d;
    a.RA = RA; a.Rb = Rb; a.r_rows = r_rows; a.n_reg = n_reg; a.reg_shared = reg_shared ? 1 : 0; a.Nr = Nr; a.mr = mr;
    a.abs_tol = abs_tol; a.intersect_tol = intersect_tol; a.max_steps = 1000000;
    a.pieceA = piece_A; a.pieceb = piece_b; a.piece_rows = piece_rows; a.piece_reduce = piece_reduce;
    a.piece_owner = piece_owner; a.piece_seq = piece_seq; a.piece_cap = piece_cap; a.piece_m = piece_m;
    a.piece_used = (unsigned long long*)pieces_used;
    a.status = status; a.n_pieces = n_pieces; a.n_lp = n_lp; a.next_problem = work_counter;
    // an LP of the search has the rows of poly, the rows of the cells it has been cut by so
    // far and the rows of the cell being probed; searches that go past 32 * RPL rows end
    // with INDEX_ERROR
    const int rows = mp + 2 * Nr * mr;
    const bool small = d + 1 <= NS;
    if (small && rows <= 64) return launch_diff<2, true>(a, st);
    if (rows <= 64) return launch_diff<2, false>(a, st);
    return launch_diff<4, false>(a, st);
}

}  // extern "C"

/* polytope_b200 -- C ABI of the B200 batched-LP engine.
 *
 * Drop-in boundary for the LP hot path of tulip-control/polytope.  Each entry
 * point cites the reference interface (file:line under /root/reference) whose
 * loop of one-at-a-time `lpsolve` calls it replaces.  INTEGRATION.md shows the
 * ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *     the caller owns all buffers, nothing is allocated or freed inside;
 *   - matrices are C-order (row-major) float64, exactly numpy's default;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); all
 *     work is stream-ordered and asynchronous, results are valid after the
 *     caller synchronises the stream;
 *   - return value: 0 on success, a negative PB200_E* code otherwise; nothing
 *     ever throws across the ABI.  pb200_last_error() describes the last failure
 *     of the calling thread;
 *   - LP status bytes use scipy.optimize.linprog's convention, which is what
 *     polytope/solvers.py:92-93 documents: 0 optimal, 1 iteration limit,
 *     2 infeasible, 3 unbounded, 4 numerical trouble.
 *   - supported sizes: 1 <= n <= 32 columns, 1 <= m <= 128 rows per LP
 *     (the reduce / adjacency pipelines additionally need m <= 64 because row
 *     sets travel as 64-bit masks).  Anything else returns PB200_EUNSUPPORTED.
 */
#ifndef POLYTOPE_B200_H
#define POLYTOPE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB200_OK 0
#define PB200_EINVAL (-1)        /* bad argument (null pointer, negative size) */
#define PB200_EUNSUPPORTED (-2)  /* size outside the kernel envelope */
#define PB200_ECUDA (-3)         /* a CUDA call failed; see pb200_last_error() */
#define PB200_EWORKSPACE (-4)    /* workspace too small */

/* flag bits written by pb200_reduce_batch into flags[p] */
#define PB200_F_EMPTY 1u      /* not full-dimensional: reference returns Polytope() (polytope.py:1081-1082) */
#define PB200_F_MINREP 2u     /* went through the per-row LP loop: result has minrep=True (:1161-1162) */
#define PB200_F_BBOX 4u       /* bounding-box prefilter ran (:1118-1134) */
#define PB200_F_LPFAIL 8u     /* some LP ended with status 1/4 (reference: silently drops the row in reduce, raises in bounding_box) */

const char* pb200_version(void);
const char* pb200_last_error(void);

/* B independent LPs  min c'x s.t. Gx <= h, x free.
 * Replaces: polytope.solvers.lpsolve(c, G, h), polytope/solvers.py:76-106,
 * called once per LP (scipy adapter :149-158).
 *   G[B][m][n], h[B][m], c[B][n]; m_rows[B] (nullable) = rows actually used by
 *   LP i (ragged batches; rows >= m_rows[i] are ignored).
 *   x[B][n], fun[B] (valid where status == 0), status[B], iters[B] (nullable). */
int pb200_lp_batch(const double* G, const double* h, const double* c,
                   const int32_t* m_rows, int B, int m, int n,
                   double* x, double* fun, int8_t* status, int32_t* iters,
                   void* stream);

/* The same contract for LPs of any size: m unlimited, n <= 64 (one LP per CTA, G streamed through shared
 * memory in chunks of 256 rows).  pb200_lp_batch covers m <= 128, n <= 32; this entry is what extreme()'s final
 * is_fulldim(Q) (one row per vertex, polytope/polytope.py:1666-1670), intersections / envelopes of polytopes with
 * many rows (:255-275, :1414-1464) and Chebyshev LPs of d >= 32 need.  The workspace holds the per-row iterates
 * of the resident CTAs: pb200_lp_big_workspace_bytes(B, m, n) bytes.  shared_G != 0: G is ONE [m][n] matrix used by
 * all B LPs (the row LPs of reduce(), :1142-1160, differ only in c and h). */
size_t pb200_lp_big_workspace_bytes(int B, int m, int n);
int pb200_lp_batch_big(const double* G, const double* h, const double* c,
                       const int32_t* m_rows, int B, int m, int n, int shared_G,
                       double* x, double* fun, int8_t* status, int32_t* iters,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Row normalisation of the Polytope constructor for P stacked polytopes.
 * Replaces: Polytope.__init__, polytope/polytope.py:128-138 (norm in numpy's
 * summation order, rows with norm <= 1e-10 dropped).
 *   A[P][m][d], b[P][m] in; An, bn out (same shapes); valid[P] = bit i set iff
 *   row i survives (m <= 64).  m_rows nullable as above. */
int pb200_normalize_batch(const double* A, const double* b, const int32_t* m_rows,
                          int P, int m, int d, double* An, double* bn,
                          uint64_t* valid, void* stream);

/* Chebyshev ball LP of P polytopes used as given (no normalisation).
 * Replaces: cheby_ball, polytope/polytope.py:1280-1300 (and is_fulldim :962-985,
 * which thresholds the radius).
 *   rows[P] nullable: bit mask of the rows to use (default: the first m_rows[p]
 *   or m rows).  r[P] = x[-1] of the LP, xc[P][d] = x[:-1], status[P]. */
int pb200_cheby_batch(const double* A, const double* b, const int32_t* m_rows,
                      const uint64_t* rows, int P, int m, int d,
                      double* r, double* xc, int8_t* status, void* stream);

/* Bounding boxes: 2d LPs per polytope.
 * Replaces: bounding_box, polytope/polytope.py:1362-1411 (status 3 -> -/+inf,
 * status 2 -> l = 0, u = l).  lo[P][d], hi[P][d]; status[P][2d] raw LP statuses. */
int pb200_bbox_batch(const double* A, const double* b, const int32_t* m_rows,
                     int P, int m, int d, double* lo, double* hi, int8_t* status,
                     void* stream);

/* Batched reduce(): redundant-row removal for P polytopes given as RAW (A, b),
 * i.e. what `reduce(Polytope(A, b))` computes.
 * Replaces: reduce, polytope/polytope.py:1053-1163, including the constructor
 * normalisation (:128-138), is_fulldim (:1081), the b == inf drop (:1087-1089),
 * the duplicate-direction filter (:1094-1112), both early exits (:1114-1116,
 * :1135-1138), the bounding-box prefilter (:1118-1134) and the per-row LP loop
 * (:1142-1160) with its +0.1/-0.1 one-ulp drift of b.
 *   normalize: bit field.  PB200_REDUCE_NORMALIZE (1): (A, b) are raw constructor
 *             arguments; clear: they are the .A/.b of an existing Polytope and are
 *             used as they are.  PB200_REDUCE_NO_EARLY_EXIT (2): the reference's
 *             `nonEmptyBounded=0` -- the two `neq <= nx + 1` exits are skipped.
 *   keep[P]   bit i set iff input row i is kept
 *   flags[P]  PB200_F_* bits
 *   r[P], xc[P][d]  Chebyshev ball of the input (first is_fulldim call)
 *   b_out[P][m]     constructor-normalised b after the reference's drift
 *   A_out[P][m][d]  constructor-normalised A (nullable)
 *   n_lp[P]         LPs the reference algorithm solves for this polytope
 *   lp_iters[P]     (nullable) interior-point iterations summed over those LPs
 *   workspace: pb200_reduce_workspace_bytes(P, m, d) bytes of device memory. */
#define PB200_REDUCE_NORMALIZE 1
#define PB200_REDUCE_NO_EARLY_EXIT 2
size_t pb200_reduce_workspace_bytes(int P, int m, int d);
int pb200_reduce_batch(const double* A, const double* b, const int32_t* m_rows,
                       int P, int m, int d, double abs_tol, int normalize,
                       uint64_t* keep, uint32_t* flags, double* r, double* xc,
                       double* b_out, double* A_out, int32_t* n_lp, int32_t* lp_iters,
                       void* workspace, size_t workspace_bytes, void* stream);

/* is_adjacent() over T pairs of cells (single polytopes, already normalised by
 * the constructor).
 * Replaces: is_adjacent overlap=True branch, polytope/polytope.py:1856-1866,
 * driven by find_adjacent_regions / compute_adj, prop2partition.py:46-63,
 * :244-261: stack both cells, add abs_tol to b, re-normalise, Chebyshev LP,
 * flag = radius > abs_tol/10.
 *   A[ncell][mc][d], b[ncell][mc]; pair_i/pair_j[T] or both NULL for the
 *   lower-triangular enumeration t -> (i, j), j < i, of find_adjacent_regions
 *   (then T must be ncell*(ncell-1)/2).  adjacent[T] (0/1), radius[T] nullable. */
int pb200_adjacent_pairs(const double* A, const double* b, int ncell, int mc, int d,
                         const int32_t* pair_i, const int32_t* pair_j, long long T,
                         double abs_tol, uint8_t* adjacent, double* radius,
                         int8_t* status, void* stream);

/* The same test over a contiguous range [t_begin, t_begin + T) of an implicit pair
 * enumeration -- what a rank of a sharded partition job runs, no pair lists in memory:
 *   order 0: t = i (i - 1) / 2 + j, j < i   (find_adjacent_regions, prop2partition.py:57-61)
 *   order 1: all ordered pairs i != j, row-major over i  (MetricPartition.compute_adj,
 *            prop2partition.py:253-261)
 * Outputs are indexed from 0 (pair t_begin + k -> adjacent[k]). */
int pb200_adjacent_range(const double* A, const double* b, int ncell, int mc, int d, int order,
                         long long t_begin, long long T, double abs_tol, uint8_t* adjacent,
                         double* radius, int8_t* status, void* stream);

/* ---- point-set kernels ---------------------------- */

/* contains(): which of N points (column vectors: points[d][N], exactly the
 * d x N array the reference takes) satisfy A x - b < abs_tol for every row.
 * Replaces: Polytope.contains, polytope/polytope.py:206-218 (any_of == 0,
 * out[P][N]) and Region.contains, :736-748 (any_of != 0, out[N] = OR over the
 * P member polytopes).  Products are accumulated as numpy's A.dot(points) does
 * (OpenBLAS dgemm: fma chain over k), so the flags are bit-identical. */
int pb200_contains_batch(const double* A, const double* b, const int32_t* m_rows,
                         int P, int m, int d, const double* points, long long N,
                         double abs_tol, int any_of, uint8_t* out, void* stream);

/* Monte-Carlo containment count of volume().
 * Replaces: polytope/polytope.py:1583-1591:
 *   x = l + default_rng(seed).random((d, N)) * (u - l);  count(all(A x - b < 0)).
 * The uniform samples are regenerated on the device from numpy's PCG64 stream:
 * rng_state[P][4] = (state_hi, state_lo, inc_hi, inc_lo) of
 * np.random.default_rng(seed).bit_generator.state for each polytope, so the
 * counts equal the reference's for the same seed.  lo/hi[P][d] = bounding box.
 * volume = prod(hi - lo) * count / N is left to the host (one multiply). */
int pb200_volume_counts(const double* A, const double* b, const int32_t* m_rows,
                        int P, int m, int d, const double* lo, const double* hi,
                        long long N, const uint64_t* rng_state,
                        unsigned long long* count, void* stream);

/* quickhull's point-to-hyperplane distance sweep as a stand-alone reduction.
 * Replaces: distance(), polytope/quickhull.py:117-121, in the assignment loops
 * :226-246 and :316-336: dist = sum(n * p) - off (numpy summation order).
 *   points[N][d] (row-major, as quickhull takes them), normals[F][d], offsets[F]
 *   first_facet[N]: first facet (in order) with dist > tol, else -1   (nullable)
 *   far_facet[N], far_dist[N]: arg-max / max of dist over all facets   (nullable) */
int pb200_point_facet_sweep(const double* points, const double* normals,
                            const double* offsets, long long N, int F, int d,
                            double tol, int32_t* first_facet, int32_t* far_facet,
                            double* far_dist, void* stream);

/* ---- convex hulls and vertex enumeration ---------- */

/* per-hull status written by pb200_hull_batch */
#define PB200_HULL_OK 0
#define PB200_HULL_FEW_POINTS 1   /* npt <= d: reference returns an empty hull (quickhull.py:155-157) */
#define PB200_HULL_FLAT 2         /* not full-dimensional: reference returns an empty hull (:159-165) */
#define PB200_HULL_FACET_CAP 3    /* more live facets than facet_cap: call again with a larger cap */
#define PB200_HULL_OUT_CAP 4      /* output pool too small: *total_facets tells the size needed */
#define PB200_HULL_SINGULAR 5     /* a facet's vertices were affinely dependent */

/* Convex hulls of H point sets.
 * Replaces: quickhull(POINTS, abs_tol), polytope/quickhull.py:141-359 (and so
 * qhull(), polytope/polytope.py:1685-1695).
 *   points[H][Nmax][d] row-major, n_pts[H] nullable (ragged batches), 2 <= d <= 16
 *   facet_cap: facet slots per hull in the workspace
 *   out_A[out_cap][d], out_b[out_cap], out_vid[out_cap][d]: facets of all hulls in
 *     one pool: unit outer normals, offsets (A x <= b), and the input indices of the d
 *     vertices of each (simplicial) facet, ascending
 *   facet_off[H], facet_cnt[H]: where hull h's facets are in the pool
 *   status[H]: PB200_HULL_*;  is_vertex[H][Nmax] (nullable): 1 for hull vertices
 *   stats[H][2] (nullable): points inserted, facets created
 *   total_facets (device, 1 value): facets of all hulls, also when the pool overflowed */
size_t pb200_hull_workspace_bytes(int H, int Nmax, int d, int facet_cap);
int pb200_hull_batch(const double* points, const int32_t* n_pts, int H, int Nmax, int d,
                     double abs_tol, int facet_cap, double* out_A, double* out_b,
                     int32_t* out_vid, long long out_cap, long long* facet_off,
                     int32_t* facet_cnt, int32_t* status, uint8_t* is_vertex,
                     int32_t* stats, long long* total_facets, void* workspace,
                     size_t workspace_bytes, void* stream);

/* extreme(): polar dual points Ai = A_i / (b_i - A_i . xc) of P polytopes.
 * Replaces: polytope/polytope.py:1659-1664.  out[P][m][d] (rows >= m_rows[p] zero). */
int pb200_dual_points(const double* A, const double* b, const int32_t* m_rows,
                      const double* xc, int P, int m, int d, double* out, void* stream);

/* extreme(): vertices from the facets (H, K) of the dual hull: V = H / K + xc
 * after the constructor normalisation qhull()'s Polytope(A, b) applies.
 * Replaces: polytope/polytope.py:1665-1676.  V[pool][d], same indexing as the pool. */
int pb200_dual_facets_to_vertices(const double* hull_A, const double* hull_b,
                                  const long long* facet_off, const int32_t* facet_cnt,
                                  const double* xc, int P, int d, int max_cnt,
                                  double* V, void* stream);

/* ---- set difference -------------------------------- */

/* per-problem status written by pb200_region_diff_batch */
#define PB200_DIFF_PIECES 0       /* result = the n_pieces[t] pieces in the pool (0 pieces: Polytope()) */
#define PB200_DIFF_UNTOUCHED 1    /* no cell intersects poly: the reference returns poly itself (:2155-2157) */
#define PB200_DIFF_COVERED 2      /* a cell has no facet cutting poly: the reference returns Polytope() (:2183-2184) */
#define PB200_DIFF_POOL_FULL 3    /* piece pool (or piece_m) too small: *pieces_used tells the size needed */
#define PB200_DIFF_INDEX_ERROR 4  /* the reference would raise IndexError / sizes outside the envelope */
#define PB200_DIFF_STEP_LIMIT 5

/* T independent set differences poly_t \ region_t, each a depth-first search
 * over Chebyshev LPs run by one warp.
 * Replaces: region_diff(poly, reg, abs_tol, intersect_tol), polytope/polytope.py:2117-2282.
 *   PA[T][mp][d], Pb[T][mp], p_rows[T] (nullable): the minuends, constructor-normalised
 *   RA[T][Nr][mr][d], Rb[T][Nr][mr], r_rows[T][Nr], n_reg[T] (nullable): the cells of each
 *     region; with reg_shared != 0 there is ONE region (leading dimension 1) used by
 *     every problem (cfg3: 50 000 cells minus the same polytope)
 *   piece pool (device, caller-owned): piece_A[piece_cap][piece_m][d], piece_b[..][piece_m]
 *     raw stacked rows of each piece exactly as the reference hands them to
 *     Polytope(A[INDICES], B[INDICES]); piece_rows = row count; piece_reduce = 1 when the
 *     reference passes the piece through reduce() (:2276); piece_owner = t; piece_seq =
 *     position in the reference's union order
 *   pieces_used (device, 1 value): pieces appended (also beyond piece_cap)
 *   status[T] PB200_DIFF_*, n_pieces[T], n_lp[T] (LPs solved for problem t)
 *   work_counter: 4 bytes of device scratch.
 * Limits: every LP of a search (rows of poly + rows of the cells cutting it) <= 128 rows,
 * else that problem ends with PB200_DIFF_INDEX_ERROR; Nr <= 64, Nr * mr <= 2048, d <= 31.
 * piece_m = min(128, mp + 2 * Nr * mr) is always enough. */
int pb200_region_diff_batch(const double* PA, const double* Pb, const int32_t* p_rows,
                            int T, int mp, int d, const double* RA, const double* Rb,
                            const int32_t* r_rows, const int32_t* n_reg, int reg_shared,
                            int Nr, int mr, double abs_tol, double intersect_tol,
                            double* piece_A, double* piece_b, int32_t* piece_rows,
                            int32_t* piece_reduce, int32_t* piece_owner, int32_t* piece_seq,
                            long long piece_cap, int piece_m, long long* pieces_used,
                            int32_t* status, int32_t* n_pieces, int32_t* n_lp,
                            int* work_counter, void* stream);

/* Per-stage device times of the last pb200_reduce_batch call made while
 * profiling was enabled: CUDA events recorded on the launch stream around the
 * 7 stages (normalize, cheby LP, prefilter, bbox LPs, candidates, row LPs,
 * finalize).  pb200_profile_read synchronises on the last event. */
void pb200_profile_enable(int on);
int pb200_profile_read(float* stage_ms, int n);

/* Diagnostics: which kernel variant pb200_normalize_batch / pb200_reduce_batch
 * use for the constructor normalisation (the batched (A|b) read).  -2 (default):
 * tiled streaming kernel, bulk async copies (TMA) when the buffers are 16-byte
 * aligned and m is even, 64-bit loads otherwise; 0 / 1 / 2 force 64-bit loads /
 * 128-bit loads / bulk copies (still degrading to 64-bit when unaligned);
 * -1 = row-per-lane kernel without shared-memory staging.  All variants produce
 * identical bits; tools/normalize_bench.py measures them. */
void pb200_normalize_variant(int variant);

/* Measurement: fp64 FMA throughput of the current device in TFLOP/s (2 flop per DFMA), timed
 * with CUDA events on `stream`; `scratch` is device memory of >= 1024 * #SMs doubles.  bench.py's
 * roofline.fp64 divides by this instead of a datasheet figure. */
int pb200_measure_dfma_tflops(double* scratch, size_t scratch_doubles, double* tflops, void* stream);

/* Diagnostics: the reduce / bounding-box LPs of polytopes with d <= 8 and m <= 64 run on the
 * lane solver (one LP per lane, csrc/lp_lane.cuh); on = 0 sends them back to the warp-per-LP
 * kernels for A/B measurements.  Default on. */
void pb200_lane_solver(int on);

/* Number of kernels this library has launched since load (bench.py's
 * `gpu_launches` evidence). */
long long pb200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* POLYTOPE_B200_H */

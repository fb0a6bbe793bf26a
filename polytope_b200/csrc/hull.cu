// polytope_b200: batched convex hulls (sm_100a).
//
// Replaces: quickhull(POINTS), polytope/quickhull.py:141-359, as called by
// qhull() (polytope/polytope.py:1685-1695) and through it by extreme()
// (:1654-1676, hull of the polar dual of a polytope).
//
// The reference grows the hull one point at a time and spends its time in
// Python loops over facets (93 % in the O(|NV|^2 d^2) `is_neighbor` scan at
// d = 8).  The hull itself is unique, so this kernel keeps the incremental
// structure -- start simplex, furthest outside point, visible set, horizon,
// cone of new facets, reassignment of the orphaned outside points
// (quickhull.py:168-190, :248-347) -- but makes every step a data-parallel sweep
// of one CTA over a structure-of-arrays facet store:
//
//   * every facet keeps D neighbour pointers (nbr[f][i] = the facet across the
//     ridge that omits vertex i), so the visible set is the reference's own walk
//     (quickhull.py:252-266) done level by level by the whole CTA: only the
//     visible facets and their rim are touched (2 % of the facets at d = 12;
//     round 1 swept every live facet per insertion -- 34 MB of DRAM traffic per
//     hull for 2 MB of output);
//   * the horizon is read off the pointers (a ridge of a visible facet whose
//     neighbour is not visible); the cone's internal neighbours are found by
//     hashing the ridges that contain the new point (sorted (d-2)-tuples of
//     vertex ids): two new facets meet in each -- this replaces the O(|NV|^2 d^2)
//     `is_neighbor` scan (quickhull.py:305-310);
//   * facet hyperplanes come from a (d x d) Gauss-Jordan solve of V x = 1 held in
//     the registers of a half warp (the reference solves the equivalent
//     (d+1) x (d+1) system, quickhull.py:66-85), two facets per warp;
//   * one CTA owns one hull; a batch of hulls (cfg4: 1000 duals of 12-D
//     polytopes, ~20 000 facets each) is spread over a persistent grid with a
//     global work counter.
//
// Points keep their input indices, facets come out as (normal, offset, vertex
// ids); the order of the facets is an implementation detail (the reference's
// own order depends on an unseeded random start simplex, quickhull.py:172).
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace pb200 {

constexpr int HT = 256;              // threads per CTA
constexpr int HW = HT / 32;          // warps
constexpr int HULL_MAX_D = 16;
constexpr unsigned FULL = 0xffffffffu;

enum : int { HS_OK = 0, HS_FEW_POINTS = 1, HS_FLAT = 2, HS_FACET_CAP = 3, HS_OUT_CAP = 4, HS_SINGULAR = 5 };

struct HullArgs {
    const double* pts;        // [H][Nmax][d]
    const int32_t* n_pts;     // nullable [H]
    int H, Nmax, d, cap;
    double tol;
    char* ws;                 // per-CTA workspace slices
    size_t ws_stride;
    int x_in_smem;            // shifted points live in shared memory
    double* outA;             // [out_cap][d]
    double* outb;             // [out_cap]
    int32_t* outV;            // [out_cap][d]
    long long out_cap;
    unsigned long long* out_used;
    long long* facet_off;     // [H]
    int32_t* facet_cnt;       // [H]
    int32_t* status;          // [H]
    uint8_t* is_vertex;       // nullable [H][Nmax]
    int32_t* stats;           // nullable [H][2]: points inserted, facets created
    int* next_hull;
};

struct HullWs {
    double* nrm;      // [cap][d]  (array of structures: a facet is touched as a whole)
    double* off;      // [cap]
    int32_t* vid;     // [cap][d] sorted ascending per facet
    int32_t* nbr;     // [cap][d] facet across the ridge that omits vertex i (-1: not wired)
    int32_t* state;   // [cap] 1 = live
    int32_t* mark;    // [cap] visibility stamps of the insertions: 2 it + 2 visible, 2 it + 3 tested, not visible
    int32_t* disc;    // [cap] claim of the walk (lowest discovering item), INT_MAX when idle
    int32_t* ppos;    // [cap] per cone facet k: position of the new point in its sorted vertex list
    int32_t* redo;    // [cap] cone facets whose pencil plane failed the check (rebuilt by Gauss-Jordan)
    unsigned long long* ikey;   // [IKEY_CAP] 64-bit key of the cone ridge (facet k, position i) at k * d + i
    int32_t* vis_list;  // [cap]
    int32_t* hor_list;  // [cap] ridge items f*d+i
    int32_t* new_list;  // [cap] slots of the new facets
    int32_t* free_stack;  // [cap]
    uint32_t* table;  // [tab_cap]
    double* X;        // [d][Nmax] shifted points (unless in shared memory)
    int32_t* owner;   // [Nmax]
    double* odist;    // [Nmax]
    int32_t* orph;    // [Nmax] orphaned outside points of the current step
    int32_t* cand;    // [Nmax] first new facet (index into new_list) an orphan is outside of
};

constexpr int IKEY_CAP = 1 << 16;    // cone ridges with precomputed keys; larger cones hash the vertex rows
__host__ __device__ inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }
// 64-bit mix of a vertex id (splitmix64 finaliser).  The key of a ridge is the SUM of the mixes of its vertices: all
// d - 1 keys of a new facet come from one total by subtraction, and equal vertex sets give equal keys whatever the
// order.  Two different (d-2)-sets collide with probability ~2^-64 per comparison.
__device__ __forceinline__ unsigned long long vmix(int v) {
    unsigned long long x = (unsigned long long)(unsigned)v + 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ inline unsigned table_cap(int cap, int d) {
    unsigned need = 2u * (unsigned)cap * (unsigned)d;
    unsigned t = 1024;
    while (t < need) t <<= 1;
    return t;
}
__host__ __device__ inline size_t hull_ws_bytes(int Nmax, int d, int cap, int x_in_smem) {
    size_t s = 0;
    s += align_up(sizeof(double) * (size_t)d * cap);
    s += align_up(sizeof(double) * (size_t)cap);
    s += 2 * align_up(sizeof(int32_t) * (size_t)d * cap);
    s += 5 * align_up(sizeof(int32_t) * (size_t)cap);
    s += 4 * align_up(sizeof(int32_t) * (size_t)cap);
    s += align_up(sizeof(unsigned long long) * (size_t)IKEY_CAP);
    s += align_up(sizeof(uint32_t) * (size_t)table_cap(cap, d));
    if (!x_in_smem) s += align_up(sizeof(double) * (size_t)d * Nmax);
    s += 3 * align_up(sizeof(int32_t) * (size_t)Nmax);
    s += align_up(sizeof(double) * (size_t)Nmax);
    return s;
}
__device__ inline HullWs hull_carve(char* p, int Nmax, int d, int cap, int x_in_smem, double* smem_x) {
    HullWs w;
    auto take = [&](size_t n) { char* q = p; p += align_up(n); return q; };
    w.nrm = (double*)take(sizeof(double) * (size_t)d * cap);
    w.off = (double*)take(sizeof(double) * (size_t)cap);
    w.vid = (int32_t*)take(sizeof(int32_t) * (size_t)d * cap);
    w.nbr = (int32_t*)take(sizeof(int32_t) * (size_t)d * cap);
    w.state = (int32_t*)take(sizeof(int32_t) * (size_t)cap);
    w.mark = (int32_t*)take(sizeof(int32_t) * (size_t)cap);
    w.disc = (int32_t*)take(sizeof(int32_t) * (size_t)cap);
    w.ppos = (int32_t*)take(sizeof(int32_t) * (size_t)cap);
    w.redo = (int32_t*)take(sizeof(int32_t) * (size_t)cap);
    w.ikey = (unsigned long long*)take(sizeof(unsigned long long) * (size_t)IKEY_CAP);
    w.vis_list = (int32_t*)take(sizeof(int32_t) * (size_t)cap);
    w.hor_list = (int32_t*)take(sizeof(int32_t) * (size_t)cap);
    w.new_list = (int32_t*)take(sizeof(int32_t) * (size_t)cap);
    w.free_stack = (int32_t*)take(sizeof(int32_t) * (size_t)cap);
    w.table = (uint32_t*)take(sizeof(uint32_t) * (size_t)table_cap(cap, d));
    w.X = x_in_smem ? smem_x : (double*)take(sizeof(double) * (size_t)d * Nmax);
    w.owner = (int32_t*)take(sizeof(int32_t) * (size_t)Nmax);
    w.odist = (double*)take(sizeof(double) * (size_t)Nmax);
    w.orph = (int32_t*)take(sizeof(int32_t) * (size_t)Nmax);
    w.cand = (int32_t*)take(sizeof(int32_t) * (size_t)Nmax);
    return w;
}

// ---- block-wide helpers (all HT threads must call) ----
struct BlockScratch {
    double dv[HW];
    int iv[HW];
    int cnt[2][HW];  // double-buffered warp counts of block_compact_pos
};

// arg-max with lowest-index tie break; idx < 0 means "no candidate"
__device__ __forceinline__ void block_argmax(BlockScratch& bs, double v, int idx, double& out_v, int& out_i) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double ov = __shfl_xor_sync(FULL, v, o);
        const int oi = __shfl_xor_sync(FULL, idx, o);
        const bool take = (oi >= 0) && (idx < 0 || ov > v || (ov == v && oi < idx));
        if (take) { v = ov; idx = oi; }
    }
    const int wid = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { bs.dv[wid] = v; bs.iv[wid] = idx; }
    __syncthreads();
    v = bs.dv[0];
    idx = bs.iv[0];
#pragma unroll
    for (int w = 1; w < HW; ++w) {
        const double ov = bs.dv[w];
        const int oi = bs.iv[w];
        if ((oi >= 0) && (idx < 0 || ov > v || (ov == v && oi < idx))) { v = ov; idx = oi; }
    }
    out_v = v;
    out_i = idx;
}

// ordered compaction step: threads with `flag` get consecutive positions from
// `base` on (in thread order); every thread advances its own copy of `base`
// identically.  One barrier per call (the warp counts are double-buffered on
// `parity`, which every thread toggles).
__device__ __forceinline__ int block_compact_pos(BlockScratch& bs, bool flag, int& base, int& parity) {
    const unsigned bal = __ballot_sync(FULL, flag);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) bs.cnt[parity][wid] = __popc(bal);
    __syncthreads();
    int before = base, total = 0;
#pragma unroll
    for (int w = 0; w < HW; ++w) {
        const int c = bs.cnt[parity][w];
        if (w < wid) before += c;
        total += c;
    }
    parity ^= 1;
    base += total;
    return flag ? before + __popc(bal & ((1u << lane) - 1u)) : -1;
}

// Ordered compaction of a whole index range: emit(pos, f) is called for every
// f in [0, n) with flag(f) true, positions counted from `base` on.  Each warp
// walks a contiguous run of 8 x 32 indices per round (coalesced, eight
// independent evaluations of `flag` in flight), keeps the eight ballots in
// registers and needs ONE block barrier per 2048 indices for the cross-warp
// prefix.  The order is deterministic (by warp run, then index).
template <class Flag, class Emit>
__device__ __forceinline__ void block_compact_range(BlockScratch& bs, int n, int& base, int& parity, Flag flag, Emit emit) {
    constexpr int Q = 8;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int c0 = 0; c0 < n; c0 += HT * Q) {
        unsigned masks[Q];
        bool fl[Q];
        int mine = 0;
        // all eight evaluations first (their loads overlap), then the eight ballots
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int f = c0 + wid * (32 * Q) + 32 * q + lane;
            fl[q] = f < n && flag(f);
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            masks[q] = __ballot_sync(FULL, fl[q]);
            mine += __popc(masks[q]);
        }
        if (lane == 0) bs.cnt[parity][wid] = mine;
        __syncthreads();
        int before = base, total = 0;
#pragma unroll
        for (int w = 0; w < HW; ++w) {
            const int c = bs.cnt[parity][w];
            if (w < wid) before += c;
            total += c;
        }
        parity ^= 1;
        base += total;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            if ((masks[q] >> lane) & 1u) emit(before + __popc(masks[q] & ((1u << lane) - 1u)), c0 + wid * (32 * Q) + 32 * q + lane);
            before += __popc(masks[q]);
        }
    }
}

// ---- ridges ----
// element t of the ridge of facet f that omits vertex i
__device__ __forceinline__ int ridge_elem(const int32_t* vid, int d, int f, int i, int t) {
    return vid[(size_t)f * d + t + (t >= i ? 1 : 0)];
}
// element t of the (d-2)-tuple of facet f that omits the vertices at positions i and j
__device__ __forceinline__ int cone_elem(const int32_t* vid, int d, int f, int i, int j, int t) {
    const int lo = min(i, j), hi = max(i, j);
    t += t >= lo ? 1 : 0;
    t += t >= hi ? 1 : 0;
    return vid[(size_t)f * d + t];
}
__device__ __forceinline__ uint32_t cone_hash(const int32_t* vid, int d, int f, int i, int j) {
    uint32_t h = 2166136261u;
    for (int t = 0; t < d - 2; ++t) {
        h ^= (uint32_t)cone_elem(vid, d, f, i, j, t);
        h *= 16777619u;
        h ^= h >> 13;
    }
    h *= 0x85ebca6bu;
    h ^= h >> 16;
    return h;
}
__device__ __forceinline__ bool cone_equal(const int32_t* vid, int d, int f, int i, int j, int g, int k, int l) {
    for (int t = 0; t < d - 2; ++t)
        if (cone_elem(vid, d, f, i, j, t) != cone_elem(vid, d, g, k, l, t)) return false;
    return true;
}
// position of vertex id v in the sorted vertex list of facet f (v must be one of them)
__device__ __forceinline__ int vertex_pos(const int32_t* vid, int d, int f, int v) {
    int pos = 0;
    for (int t = 0; t < d; ++t) pos += vid[(size_t)f * d + t] < v ? 1 : 0;
    return pos;
}

// ---- facet hyperplane from its D vertices, one half warp per facet ----
// Lane r (< D) of the group holds vertex id `myv` (sorted ascending over r).
// Solves V x = 1 by Gauss-Jordan with row pivoting; n = x / |x|, off = 1 / |x|
// (quickhull.py:66-85 solves the same system bordered by one row/column).
// reciprocal: MUFU seed + two Newton steps (<= 1-2 ulp); the IEEE division sequence is ~3x longer
__device__ __forceinline__ double hull_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}

template <int D>
__device__ __forceinline__ bool facet_plane(const double* __restrict__ X, int ldx, int myv, int gl, double (&nk)[1], int& mycol,
                                            double& off) {
    double a[D + 1];
    const bool row = gl < D;
#pragma unroll
    for (int k = 0; k < D; ++k) a[k] = row ? X[(size_t)k * ldx + myv] : 0.0;
    a[D] = 1.0;
    bool used = !row;
    mycol = -1;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        // row pivot: largest |a[k]| among the unused rows of this half warp.  The magnitude
        // (as fp32 bits, order-preserving for non-negative floats) and the lane share one
        // 32-bit key, so the search is one 4-step integer butterfly; ties go to the higher lane.
        const unsigned key = used ? 0u : ((__float_as_uint(__double2float_rd(fabs(a[k]))) & ~0xfu) | (unsigned)gl) + 16u;
        // butterfly over the 16 lanes of the half warp (a redux with two different member masks
        // compiles to WARPSYNC.EXCLUSIVE: the two halves would run one after the other)
        unsigned best = key;
#pragma unroll
        for (int o = 8; o; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o, 16));
        if (best < 32u) ok = false;                       // every candidate was (sub)zero
        const int who = (int)((best - 16u) & 0xfu);
        const double pv = __shfl_sync(FULL, a[k], who, 16);
        const double rinv = hull_rcp(pv);
        const double f = (gl == who) ? 0.0 : a[k] * rinv;
#pragma unroll
        for (int j = k + 1; j <= D; ++j) {
            const double pj = __shfl_sync(FULL, a[j], who, 16);
            a[j] = fma(-f, pj, a[j]);
        }
        if (gl == who) { used = true; mycol = k; }
    }
    // lane that pivoted column k holds x_k = a[D] / a[k]
    double piv = 1.0;
#pragma unroll
    for (int k = 0; k < D; ++k)
        if (mycol == k) piv = a[k];
    const double xk = mycol >= 0 ? a[D] * hull_rcp(piv) : 0.0;
    double s = xk * xk;
#pragma unroll
    for (int o = 8; o; o >>= 1) s += __shfl_xor_sync(FULL, s, o, 16);
    const double mult = sqrt(s);
    if (!(mult > 0.0) || !(mult < 1e300)) ok = false;
    const double minv = hull_rcp(mult);
    nk[0] = xk * minv;
    off = minv;
    return ok;
}

// builds the facets listed by `get_ids` (k -> lane's vertex id) into slots new_list[k]
template <int D, class GetId>
__device__ __forceinline__ bool make_facets(const HullWs& w, int cap, int ldx, int count, GetId get_id, const int32_t* klist = nullptr) {
    const int group = threadIdx.x >> 4, gl = threadIdx.x & 15;
    constexpr int NG = HT / 16;
    bool all_ok = true;
    const int rounds = (count + NG - 1) / NG;
    for (int rnd = 0; rnd < rounds; ++rnd) {
        const int k = rnd * NG + group;
        const bool act = k < count;
        // get_id may use half-warp collectives: every lane calls it (idle groups redo the last item)
        const int v = get_id(act ? k : count - 1, gl);
        const int myv = (act && gl < D) ? v : 0;
        double nk[1];
        int mycol;
        double off;
        const bool ok = facet_plane<D>(w.X, ldx, myv, gl, nk, mycol, off);
        if (act) {
            const int g = w.new_list[klist ? klist[k] : k];
            if (gl < D) {
                w.vid[(size_t)g * D + gl] = myv;
                w.nbr[(size_t)g * D + gl] = -1;
                if (mycol >= 0) w.nrm[(size_t)g * D + mycol] = nk[0];
            }
            if (gl == 0) {
                w.off[g] = off;
                w.state[g] = 1;
            }
            all_ok = all_ok && ok;
        }
    }
    return all_ok;
}

__device__ __forceinline__ double facet_dist(const HullWs& w, int cap, int d, int f, const double* p) {
    double acc = 0.0;
    const double* nf = w.nrm + (size_t)f * d;
    for (int k = 0; k < d; ++k) acc = fma(nf[k], p[k], acc);
    return acc - w.off[f];
}

template <int D>
__device__ void hull_one(const HullArgs& a, const HullWs& w, int h, BlockScratch& bs, double* sh_p, double* sh_c, double* sh_q,
                         int* sh_simplex, int* sh_flag) {
    const int tid = threadIdx.x;
    const int d = D, cap = a.cap, Nmax = a.Nmax;
    const int n = a.n_pts ? min(max(a.n_pts[h], 0), Nmax) : Nmax;
    const double* P = a.pts + (size_t)h * Nmax * d;
    const int ldx = Nmax;
    auto finish = [&](int status, int nfac, long long offs, int inserted, int created) {
        if (tid == 0) {
            a.status[h] = status;
            a.facet_cnt[h] = nfac;
            a.facet_off[h] = offs;
            if (a.stats) { a.stats[2 * h] = inserted; a.stats[2 * h + 1] = created; }
        }
    };
    if (a.is_vertex)
        for (int j = tid; j < Nmax; j += HT) a.is_vertex[(size_t)h * Nmax + j] = 0;
    if (n <= d) { finish(HS_FEW_POINTS, 0, 0, 0, 0); return; }

    // ---- start simplex: lowest first coordinate, then d times the point
    // furthest from the affine span of the points chosen so far ----
    {
        double bv = 0.0;
        int bi = -1;
        for (int j = tid; j < n; j += HT) {
            const double v = -P[(size_t)j * d];
            if (bi < 0 || v > bv) { bv = v; bi = j; }
        }
        double ov;
        int oi;
        block_argmax(bs, bv, bi, ov, oi);
        if (tid == 0) sh_simplex[0] = oi;
        __syncthreads();
    }
    for (int k = 1; k <= d; ++k) {
        const int v0 = sh_simplex[0];
        double bv = 0.0;
        int bi = -1;
        for (int j = tid; j < n; j += HT) {
            double r[D];
#pragma unroll
            for (int c = 0; c < D; ++c) r[c] = P[(size_t)j * d + c] - P[(size_t)v0 * d + c];
            for (int q = 0; q < k - 1; ++q) {
                double dot = 0.0;
#pragma unroll
                for (int c = 0; c < D; ++c) dot = fma(r[c], sh_q[q * D + c], dot);
#pragma unroll
                for (int c = 0; c < D; ++c) r[c] = fma(-dot, sh_q[q * D + c], r[c]);
            }
            double nn = 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c) nn = fma(r[c], r[c], nn);
            if (bi < 0 || nn > bv) { bv = nn; bi = j; }
        }
        double ov;
        int oi;
        block_argmax(bs, bv, bi, ov, oi);
        // singular values of the simplex edge matrix must exceed 1e-10 (quickhull.py:186-188)
        if (!(ov > 1e-20)) { finish(HS_FLAT, 0, 0, 0, 0); return; }
        if (tid == 0) {
            sh_simplex[k] = oi;
            double r[D];
#pragma unroll
            for (int c = 0; c < D; ++c) r[c] = P[(size_t)oi * d + c] - P[(size_t)v0 * d + c];
            for (int q = 0; q < k - 1; ++q) {
                double dot = 0.0;
#pragma unroll
                for (int c = 0; c < D; ++c) dot = fma(r[c], sh_q[q * D + c], dot);
#pragma unroll
                for (int c = 0; c < D; ++c) r[c] = fma(-dot, sh_q[q * D + c], r[c]);
            }
            double nn = 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c) nn = fma(r[c], r[c], nn);
            const double inv = 1.0 / sqrt(nn);
#pragma unroll
            for (int c = 0; c < D; ++c) sh_q[(k - 1) * D + c] = r[c] * inv;
        }
        __syncthreads();
    }
    // sort the simplex ids, centre = mean of the simplex (quickhull.py:192-196)
    if (tid == 0) {
        for (int i = 1; i <= d; ++i) {
            const int v = sh_simplex[i];
            int j = i - 1;
            while (j >= 0 && sh_simplex[j] > v) { sh_simplex[j + 1] = sh_simplex[j]; --j; }
            sh_simplex[j + 1] = v;
        }
        for (int c = 0; c < d; ++c) {
            double s = 0.0;
            for (int i = 0; i <= d; ++i) s += P[(size_t)sh_simplex[i] * d + c] / (double)(d + 1);
            sh_c[c] = s;
        }
    }
    __syncthreads();
    for (int e = tid; e < n * d; e += HT) {
        const int j = e / d, c = e - j * d;
        w.X[(size_t)c * ldx + j] = P[e] - sh_c[c];
    }
    for (int j = tid; j < n; j += HT) { w.owner[j] = -1; w.odist[j] = 0.0; }
    for (int f = tid; f < cap; f += HT) { w.state[f] = 0; w.mark[f] = 0; w.disc[f] = 0x7fffffff; }
    __syncthreads();
    if (tid <= d) w.owner[sh_simplex[tid]] = -2;
    if (tid <= d) w.new_list[tid] = tid;
    __syncthreads();
    // initial facets: facet i omits simplex vertex i; across the ridge that also omits simplex
    // vertex j lies facet j
    bool ok = make_facets<D>(w, cap, ldx, d + 1, [&](int k, int gl) { return sh_simplex[gl + (gl >= k ? 1 : 0)]; });
    __syncthreads();
    for (int t = tid; t < (d + 1) * d; t += HT) {
        const int i = t / d, idx = t - i * d;
        w.nbr[(size_t)i * D + idx] = idx + (idx >= i ? 1 : 0);
    }
    int hi = d + 1, nfree = 0, inserted = d + 1, created = d + 1;
    __syncthreads();
    // assign every other point to the first facet it is outside of (quickhull.py:226-246)
    for (int j = tid; j < n; j += HT) {
        if (w.owner[j] == -2) continue;
        double p[D];
#pragma unroll
        for (int c = 0; c < D; ++c) p[c] = w.X[(size_t)c * ldx + j];
        for (int f = 0; f <= d; ++f) {
            const double dist = facet_dist(w, cap, d, f, p);
            if (dist > a.tol) { w.owner[j] = f; w.odist[j] = dist; break; }
        }
    }
    __syncthreads();

    int status = HS_OK, parity = 0;
    for (int iter = 0; iter < n; ++iter) {
        // (a) furthest outside point
        double bv = 0.0;
        int bi = -1;
        for (int j = tid; j < n; j += HT)
            if (w.owner[j] >= 0) {
                const double v = w.odist[j];
                if (bi < 0 || v > bv) { bv = v; bi = j; }
            }
        double ov;
        int pstar;
        block_argmax(bs, bv, bi, ov, pstar);
        if (pstar < 0) break;
        const int f0 = w.owner[pstar];            // visible by construction (quickhull.py:256)
        const int VIS = 2 * iter + 2, NOTVIS = 2 * iter + 3;
        __syncthreads();
        if (tid < d) sh_p[tid] = w.X[(size_t)tid * ldx + pstar];
        if (tid == 0) { w.owner[pstar] = -2; w.vis_list[0] = f0; w.mark[f0] = VIS; }
        __syncthreads();
        double p[D];
#pragma unroll
        for (int c = 0; c < D; ++c) p[c] = sh_p[c];
        // (b) visible facets: the neighbour walk of quickhull.py:252-266, one level per round.  A facet
        // reached from several sides is claimed by the lowest item (atomicMin), so the order of vis_list
        // -- and with it the order of the cone -- does not depend on thread timing.
        int nV = 1, lo = 0;
        while (lo < nV) {
            const int items = (nV - lo) * d;
            for (int t = tid; t < items; t += HT) {
                const int f = w.vis_list[lo + t / d], i = t % d;
                const int g = w.nbr[(size_t)f * D + i];
                if (g < 0) { ok = false; continue; }
                const int mk = w.mark[g];
                if (mk == VIS || mk == NOTVIS) continue;
                if (facet_dist(w, cap, d, g, p) > a.tol) atomicMin(w.disc + g, t);
                else w.mark[g] = NOTVIS;
            }
            __syncthreads();
            int nNew = nV;
            block_compact_range(bs, items, nNew, parity,
                                [&](int t) {
                                    const int g = w.nbr[(size_t)w.vis_list[lo + t / d] * D + t % d];
                                    return g >= 0 && w.disc[g] == t;
                                },
                                [&](int pos, int t) {
                                    const int g = w.nbr[(size_t)w.vis_list[lo + t / d] * D + t % d];
                                    w.vis_list[pos] = g;
                                    w.mark[g] = VIS;
                                    w.disc[g] = 0x7fffffff;
                                });
            __syncthreads();
            lo = nV;
            nV = nNew;
        }
        // (c) horizon: ridges of visible facets whose neighbour is not visible
        const int items = nV * d;
        int nH = 0;
        block_compact_range(bs, items, nH, parity,
                            [&](int t) {
                                const int g = w.nbr[(size_t)w.vis_list[t / d] * D + t % d];
                                return g >= 0 && w.mark[g] != VIS;
                            },
                            [&](int pos, int t) {
                                if (pos < cap) w.hor_list[pos] = w.vis_list[t / d] * d + t % d;
                            });
        __syncthreads();
        // (d) slots for the cone of new facets: recycled ones first, then fresh ones
        const int from_free = min(nH, nfree);
        const int fresh = nH - from_free;
        if (hi + fresh > cap || nH > cap) { status = HS_FACET_CAP; break; }
        for (int k = tid; k < nH; k += HT) w.new_list[k] = k < from_free ? w.free_stack[nfree - 1 - k] : hi + (k - from_free);
        const unsigned tcap = table_cap(cap, d);
        unsigned tsize = 1024;
        while (tsize < 2u * (unsigned)nH * (unsigned)d) tsize <<= 1;
        if (tsize > tcap) tsize = tcap;
        for (unsigned e = tid; e < tsize; e += HT) w.table[e] = 0;
        __syncthreads();
        // (e) new facets = horizon ridge + the new point, one thread per facet.  The hyperplanes through the
        // ridge form a pencil spanned by the planes of the two facets that meet in it (the visible one and the
        // one outside), so the plane through pstar is  dist_out(p) (n_f, off_f) - dist_f(p) (n_out, off_out):
        // O(d) instead of the d x d solve.  The result is checked against the facet's own vertices
        // (|n . v - off| <= 1e-13 max(1, |n|.|v|) for all d of them, so no error is inherited from the parent planes); a facet
        // that fails -- nearly coplanar parents -- is rebuilt from its vertices by Gauss-Jordan below.
        int nRedo = 0;
        const bool keyed = (long long)nH * d <= IKEY_CAP;
        {
            for (int k0 = 0; k0 < nH; k0 += HT) {
                const int k = k0 + tid;
                bool bad = false;
                if (k < nH) {
                    const int item = w.hor_list[k];
                    const int f = item / d, i = item - f * d;
                    const int gout = w.nbr[(size_t)f * D + i];
                    const int g = w.new_list[k];
                    int vv[D];
                    int pos = 0;
                    unsigned long long ktot = 0;
#pragma unroll
                    for (int t = 0; t < D - 1; ++t) {
                        const int e = ridge_elem(w.vid, d, f, i, t);
                        pos += e < pstar ? 1 : 0;
                        vv[t] = e;
                        ktot += vmix(e);
                    }
                    if (keyed) {
                        // key of the cone ridge that omits ridge vertex t (and pstar), stored at the vertex's
                        // position in the new facet's sorted list
#pragma unroll
                        for (int t = 0; t < D - 1; ++t) w.ikey[(size_t)k * d + t + (t >= pos ? 1 : 0)] = ktot - vmix(vv[t]);
                    }
                    const double* nf = w.nrm + (size_t)f * D;
                    const double* ng = w.nrm + (size_t)gout * D;
                    double af[D], ag[D];
                    double df = -w.off[f], dg = -w.off[gout];
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        af[c] = nf[c];
                        ag[c] = ng[c];
                        df = fma(af[c], p[c], df);
                        dg = fma(ag[c], p[c], dg);
                    }
                    double nn[D];
                    double n2 = 0.0;
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        nn[c] = dg * af[c] - df * ag[c];
                        n2 = fma(nn[c], nn[c], n2);
                    }
                    double off = dg * w.off[f] - df * w.off[gout];
                    // outward: the shifted origin is strictly inside, so the offset must come out positive
                    const double sc = (off < 0.0 ? -1.0 : 1.0) / sqrt(n2);
                    off *= sc;
#pragma unroll
                    for (int c = 0; c < D; ++c) nn[c] *= sc;
                    double worst = 0.0, mag = 1.0;
#pragma unroll
                    for (int t = 0; t < D; ++t) {
                        const int v = t < D - 1 ? vv[t] : pstar;
                        double acc = -off, ab = 0.0;
#pragma unroll
                        for (int c = 0; c < D; ++c) {
                            const double xv = w.X[(size_t)c * ldx + v];
                            acc = fma(nn[c], xv, acc);
                            ab = fma(fabs(nn[c]), fabs(xv), ab);
                        }
                        worst = fmax(worst, fabs(acc));
                        mag = fmax(mag, ab);
                    }
                    bad = !(worst <= 1e-13 * mag) || !(off > 0.0) || !(n2 > 0.0);
                    // sorted vertex list with pstar merged in
#pragma unroll
                    for (int t = 0; t < D; ++t) {
                        const int src = t < pos ? t : t - 1;
                        int e = pstar;
#pragma unroll
                        for (int u = 0; u < D - 1; ++u)
                            if (u == src && t != pos) e = vv[u];
                        w.vid[(size_t)g * D + t] = e;
                        w.nbr[(size_t)g * D + t] = -1;
                        w.nrm[(size_t)g * D + t] = nn[t];
                    }
                    w.off[g] = off;
                    w.state[g] = 1;
                    w.ppos[k] = pos;
                }
                const int at = block_compact_pos(bs, bad, nRedo, parity);
                if (bad) w.redo[at] = k;
            }
        }
        __syncthreads();
        if (nRedo > 0) {
            ok = make_facets<D>(w, cap, ldx, nRedo, [&](int q, int gl) {
                const int k = w.redo[q];
                return gl < d ? w.vid[(size_t)w.new_list[k] * D + gl] : 0;
            }, w.redo) && ok;
            __syncthreads();
        }
        // (measured r02t / r02u, 1000 cfg4 duals: one thread per new FACET with the vertex row in registers 75 ms,
        // the ridge table in shared memory 38 ms -- 64 KB per CTA cost more L1 than the faster atomics gained --
        // against 28 ms for one thread per (facet, ridge) with the table in global memory; precomputed 64-bit
        // ridge keys then shorten every probe from five dependent memory round trips to three: 25 ms, 23.9 ms with
        // four CTAs per SM; four ridges in flight per thread on top of that changed nothing, r02ag)
        // (e2) neighbour pointers of the cone.  Across the horizon ridge (the one that omits pstar): the
        // facet outside, whose own pointer moves from the visible facet to the new one.  Across the d - 1
        // ridges that contain pstar: another new facet, found by hashing the ridge without pstar -- the first
        // of the two inserts itself, the second wires both (quickhull.py:305-310).
        for (int t = tid; t < nH * d; t += HT) {
            const int k = t / d, i = t - k * d;
            const int g = w.new_list[k];
            const int pos = w.ppos[k];
            if (keyed && i != pos) {
                // the key of this ridge was left by the thread that built the facet: the probe touches the
                // table and the other item's key only, no vertex rows
                const unsigned long long key = w.ikey[t];
                uint32_t slot = ((uint32_t)(key >> 32) ^ (uint32_t)key) * 0x85ebca6bu;
                slot = (slot ^ (slot >> 15)) & (tsize - 1);
                const uint32_t me = (uint32_t)t + 1u;
                for (unsigned probe = 0; probe < tsize; ++probe) {
                    const uint32_t cur = atomicCAS(w.table + slot, 0u, me);
                    if (cur == 0u) break;
                    const int t2 = (int)(cur - 1u);
                    if (w.ikey[t2] == key) {
                        const int k2 = t2 / d;
                        const int g2 = w.new_list[k2];
                        w.nbr[(size_t)g * D + i] = g2;
                        w.nbr[(size_t)g2 * D + (t2 - k2 * d)] = g;
                        break;
                    }
                    slot = (slot + 1) & (tsize - 1);
                }
                continue;
            }
            if (i == pos) {
                const int item = w.hor_list[k];
                const int f = item / d;
                const int gout = w.nbr[(size_t)f * D + (item - f * d)];
                w.nbr[(size_t)g * D + pos] = gout;
                for (int j = 0; j < d; ++j)
                    if (w.nbr[(size_t)gout * D + j] == f) w.nbr[(size_t)gout * D + j] = g;
                continue;
            }
            uint32_t slot = cone_hash(w.vid, d, g, i, pos) & (tsize - 1);
            const uint32_t me = (uint32_t)t + 1u;
            for (unsigned probe = 0; probe < tsize; ++probe) {
                const uint32_t cur = atomicCAS(w.table + slot, 0u, me);
                if (cur == 0u) break;
                const int t2 = (int)(cur - 1u);
                const int k2 = t2 / d, i2 = t2 - k2 * d;
                const int g2 = w.new_list[k2];
                if (cone_equal(w.vid, d, g, i, pos, g2, i2, w.ppos[k2])) {
                    w.nbr[(size_t)g * D + i] = g2;
                    w.nbr[(size_t)g2 * D + i2] = g;
                    break;
                }
                slot = (slot + 1) & (tsize - 1);
            }
        }
        // (f) orphaned outside points go to the first new facet (in cone order) they are
        // outside of (quickhull.py:316-336): all (facet, orphan) pairs in parallel
        int nO = 0;
        block_compact_range(bs, n, nO, parity,
                            [&](int j) {
                                const int o = w.owner[j];
                                return o >= 0 && w.mark[o] == VIS;
                            },
                            [&](int pos, int j) { w.orph[pos] = j; w.cand[j] = 0x7fffffff; });
        __syncthreads();
        if (nO > 0) {
            const long long pairs = (long long)nH * nO;
            for (long long t = tid; t < pairs; t += HT) {
                const int k = (int)(t / nO);
                const int j = w.orph[(int)(t - (long long)k * nO)];
                if (w.cand[j] < k) continue;
                double q[D];
#pragma unroll
                for (int c = 0; c < D; ++c) q[c] = w.X[(size_t)c * ldx + j];
                if (facet_dist(w, cap, d, w.new_list[k], q) > a.tol) atomicMin(w.cand + j, k);
            }
            __syncthreads();
            for (int o = tid; o < nO; o += HT) {
                const int j = w.orph[o];
                const int k = w.cand[j];
                if (k == 0x7fffffff) { w.owner[j] = -1; w.odist[j] = 0.0; continue; }
                double q[D];
#pragma unroll
                for (int c = 0; c < D; ++c) q[c] = w.X[(size_t)c * ldx + j];
                const int g = w.new_list[k];
                w.owner[j] = g;
                w.odist[j] = facet_dist(w, cap, d, g, q);
            }
        }
        __syncthreads();
        // (g) retire the visible facets
        for (int k = tid; k < nV; k += HT) {
            const int f = w.vis_list[k];
            w.state[f] = 0;
            w.free_stack[nfree - from_free + k] = f;
        }
        nfree = nfree - from_free + nV;
        hi += fresh;
        ++inserted;
        created += nH;
        __syncthreads();
    }
    // a singular facet system or an unwired ridge seen by any thread fails the whole hull
    ok = !__syncthreads_or(ok ? 0 : 1);
    if (status == HS_OK && !ok) status = HS_SINGULAR;
    if (status != HS_OK) { finish(status, 0, 0, inserted, created); return; }

    // ---- output: live facets, b = off + n . centre (quickhull.py:348-359) ----
    int nF = 0;
    block_compact_range(bs, hi, nF, parity, [&](int f) { return w.state[f] == 1; },
                        [&](int pos, int f) { w.vis_list[pos] = f; });
    __syncthreads();
    if (tid == 0) {
        const unsigned long long o = atomicAdd(a.out_used, (unsigned long long)nF);
        *reinterpret_cast<volatile long long*>(sh_flag) = (long long)o;
    }
    __syncthreads();
    const long long o = *reinterpret_cast<volatile long long*>(sh_flag);
    if (o + nF > a.out_cap) { finish(HS_OUT_CAP, nF, o, inserted, created); return; }
    // row-major outputs written element-wise: consecutive threads write consecutive addresses
    for (long long e = tid; e < (long long)nF * d; e += HT) {
        const int k = (int)(e / d), c = (int)(e - (long long)k * d);
        const int f = w.vis_list[k];
        a.outA[(size_t)o * d + e] = w.nrm[(size_t)f * D + c];
        const int v = w.vid[(size_t)f * D + c];
        a.outV[(size_t)o * d + e] = v;
        if (a.is_vertex) a.is_vertex[(size_t)h * Nmax + v] = 1;
    }
    for (int k = tid; k < nF; k += HT) {
        const int f = w.vis_list[k];
        double dot = 0.0;
        for (int c = 0; c < d; ++c) dot = fma(w.nrm[(size_t)f * D + c], sh_c[c], dot);
        a.outb[o + k] = w.off[f] + dot;
    }
    finish(HS_OK, nF, o, inserted, created);
}

#ifndef PB200_HULL_MINB
#define PB200_HULL_MINB 4
#endif
template <int D>
__global__ void __launch_bounds__(HT, PB200_HULL_MINB) hull_kernel(const HullArgs a) {
    extern __shared__ __align__(16) double smem_x[];
    __shared__ BlockScratch bs;
    __shared__ double sh_p[HULL_MAX_D], sh_c[HULL_MAX_D], sh_q[HULL_MAX_D * HULL_MAX_D];
    __shared__ int sh_simplex[HULL_MAX_D + 1];
    __shared__ __align__(8) int sh_flag[2];
    __shared__ int sh_h;
    const HullWs w = hull_carve(a.ws + (size_t)blockIdx.x * a.ws_stride, a.Nmax, a.d, a.cap, a.x_in_smem, smem_x);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) sh_h = atomicAdd(a.next_hull, 1);
        __syncthreads();
        const int h = sh_h;
        if (h >= a.H) break;
        hull_one<D>(a, w, h, bs, sh_p, sh_c, sh_q, sh_simplex, sh_flag);
    }
}

template <int D>
static int launch_hull(const HullArgs& a, int grid, size_t smem, cudaStream_t st) {
    PB_CHECK_CUDA(cudaFuncSetAttribute(hull_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hull_kernel<D><<<grid, HT, smem, st>>>(a);
    count_launch();
    PB_CHECK_CUDA(cudaGetLastError());
    return PB200_OK;
}

static int hull_grid(int H) {
    const int sms = sm_count();
    if (!sms) return 0;
    const int g = PB200_HULL_MINB * sms;
    return H < g ? H : g;
}
static int hull_x_in_smem(int Nmax, int d) { return (size_t)Nmax * d * sizeof(double) <= 96 * 1024 ? 1 : 0; }

// polar dual points of extreme(): Ai = A_i / (b_i - A_i . xc), polytope.py:1659-1664
__global__ void dual_points_kernel(const double* __restrict__ A, const double* __restrict__ b, const int32_t* __restrict__ m_rows,
                                   const double* __restrict__ xc, int P, int m, int d, double* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)P * m) return;
    const long long p = t / m;
    const int i = (int)(t - p * m);
    const int mm = m_rows ? min(max(m_rows[p], 0), m) : m;
    const double* row = A + (size_t)t * d;
    double* o = out + (size_t)t * d;
    if (i >= mm) {
        for (int k = 0; k < d; ++k) o[k] = 0.0;
        return;
    }
    // np.dot(A[ii, :], xmid): a length-d ddot; numpy uses pairwise-free BLAS ddot, decisions do not depend on its last bit
    double dot = 0.0;
    for (int k = 0; k < d; ++k) dot = fma(row[k], xc[(size_t)p * d + k], dot);
    const double den = __dsub_rn(b[t], dot);
    for (int k = 0; k < d; ++k) o[k] = __ddiv_rn(row[k], den);
}

// vertices of the primal from the facets of the dual hull: V = H / K + xmid, polytope.py:1671-1676
__global__ void dual_facets_to_vertices_kernel(const double* __restrict__ HA, const double* __restrict__ Hb,
                                               const long long* __restrict__ facet_off, const int32_t* __restrict__ facet_cnt,
                                               const double* __restrict__ xc, int P, int d, double* __restrict__ V) {
    // one thread per facet of the dual hull = vertex of the primal: the row norm and its reciprocal once per row
    // (one thread per element recomputed them d times: 1.7 ms for cfg4's 13.1 M vertices, 4x the HBM bound)
    const int p = blockIdx.y;
    const long long o = facet_off[p];
    const int cnt = facet_cnt[p];
    const double* centre = xc + (size_t)p * d;
    for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < cnt; f += (long long)gridDim.x * blockDim.x) {
        // Polytope(A, b) re-normalises the hull rows (polytope.py:128-138) before extreme() divides them
        const double* row = HA + (size_t)(o + f) * d;
        double v[HULL_MAX_D];
#pragma unroll
        for (int k = 0; k < HULL_MAX_D; ++k) v[k] = k < d ? row[k] : 0.0;
        const double nrm = sqrt(np_sum_squares([&](int j) { return row[j]; }, d));
        const double mult = __ddiv_rn(1.0, nrm);
        const double kk = __dmul_rn(Hb[o + f], mult);
        double* out = V + (size_t)(o + f) * d;
#pragma unroll
        for (int k = 0; k < HULL_MAX_D; ++k)
            if (k < d) out[k] = __dadd_rn(__ddiv_rn(__dmul_rn(v[k], mult), kk), centre[k]);
    }
}

}  // namespace pb200

using namespace pb200;

extern "C" {

size_t pb200_hull_workspace_bytes(int H, int Nmax, int d, int facet_cap) {
    if (H < 0 || Nmax < 1 || d < 2 || d > HULL_MAX_D || facet_cap < d + 1) return 0;
    const int sms = sm_count();
    if (!sms) return 0;
    const int grid = hull_grid(H > 0 ? H : 1);
    return (size_t)grid * hull_ws_bytes(Nmax, d, facet_cap, hull_x_in_smem(Nmax, d)) + 256;
}

int pb200_hull_batch(const double* points, const int32_t* n_pts, int H, int Nmax, int d, double abs_tol, int facet_cap,
                     double* out_A, double* out_b, int32_t* out_vid, long long out_cap, long long* facet_off, int32_t* facet_cnt,
                     int32_t* status, uint8_t* is_vertex, int32_t* stats, long long* total_facets, void* workspace,
                     size_t workspace_bytes, void* stream) {
    if (H == 0) return PB200_OK;
    if (H < 0 || !points || !out_A || !out_b || !out_vid || !facet_off || !facet_cnt || !status || !total_facets || !workspace)
        return fail(PB200_EINVAL, "pb200_hull_batch: null pointer or negative batch");
    if (d < 2 || d > HULL_MAX_D) return fail(PB200_EUNSUPPORTED, "hull: need 2 <= d <= 16");
    if (Nmax < 1 || facet_cap < d + 1 || (long long)facet_cap * d >= (1ll << 30))
        return fail(PB200_EUNSUPPORTED, "hull: need Nmax >= 1 and d+1 <= facet_cap, facet_cap*d < 2^30");
    if (H == 0) return PB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = hull_grid(H);
    if (!grid) return PB200_ECUDA;
    const int xs = hull_x_in_smem(Nmax, d);
    const size_t per = hull_ws_bytes(Nmax, d, facet_cap, xs);
    if ((size_t)grid * per + 256 > workspace_bytes) return fail(PB200_EWORKSPACE, "pb200_hull_batch: workspace too small");
    // first 256 bytes: work counter; total_facets doubles as the output-pool cursor
    PB_CHECK_CUDA(cudaMemsetAsync(workspace, 0, 256, st));
    PB_CHECK_CUDA(cudaMemsetAsync(total_facets, 0, sizeof(long long), st));
    HullArgs a;
    a.pts = points; a.n_pts = n_pts; a.H = H; a.Nmax = Nmax; a.d = d; a.cap = facet_cap; a.tol = abs_tol;
    a.ws = (char*)workspace + 256; a.ws_stride = per; a.x_in_smem = xs;
    a.outA = out_A; a.outb = out_b; a.outV = out_vid; a.out_cap = out_cap;
    a.out_used = (unsigned long long*)total_facets;
    a.facet_off = facet_off; a.facet_cnt = facet_cnt; a.status = status; a.is_vertex = is_vertex; a.stats = stats;
    a.next_hull = (int*)workspace;
    const size_t smem = xs ? sizeof(double) * (size_t)Nmax * d : 0;
    switch (d) {
        case 2: return launch_hull<2>(a, grid, smem, st);
        case 3: return launch_hull<3>(a, grid, smem, st);
        case 4: return launch_hull<4>(a, grid, smem, st);
        case 5: return launch_hull<5>(a, grid, smem, st);
        case 6: return launch_hull<6>(a, grid, smem, st);
        case 7: return launch_hull<7>(a, grid, smem, st);
        case 8: return launch_hull<8>(a, grid, smem, st);
        case 9: return launch_hull<9>(a, grid, smem, st);
        case 10: return launch_hull<10>(a, grid, smem, st);
        case 11: return launch_hull<11>(a, grid, smem, st);
        case 12: return launch_hull<12>(a, grid, smem, st);
        case 13: return launch_hull<13>(a, grid, smem, st);
        case 14: return launch_hull<14>(a, grid, smem, st);
        case 15: return launch_hull<15>(a, grid, smem, st);
        default: return launch_hull<16>(a, grid, smem, st);
    }
}

int pb200_dual_points(const double* A, const double* b, const int32_t* m_rows, const double* xc, int P, int m, int d,
                      double* out, void* stream) {
    if (P == 0) return PB200_OK;
    if (P < 0 || !A || !b || !xc || !out) return fail(PB200_EINVAL, "pb200_dual_points: null pointer");
    if (P == 0) return PB200_OK;
    dual_points_kernel<<<blocks_for((long long)P * m, 256), 256, 0, (cudaStream_t)stream>>>(A, b, m_rows, xc, P, m, d, out);
    count_launch();
    PB_CHECK_CUDA(cudaGetLastError());
    return PB200_OK;
}

int pb200_dual_facets_to_vertices(const double* hull_A, const double* hull_b, const long long* facet_off,
                                  const int32_t* facet_cnt, const double* xc, int P, int d, int max_cnt, double* V, void* stream) {
    if (P < 0 || !hull_A || !hull_b || !facet_off || !facet_cnt || !xc || !V) return fail(PB200_EINVAL, "pb200_dual_facets_to_vertices: null pointer");
    if (P == 0 || max_cnt <= 0) return PB200_OK;
    if (P > 65535) return fail(PB200_EUNSUPPORTED, "dual_facets_to_vertices: at most 65535 polytopes per call");
    unsigned gx = blocks_for((long long)max_cnt, 256);
    if (gx > 1024) gx = 1024;
    dual_facets_to_vertices_kernel<<<dim3(gx, (unsigned)P), 256, 0, (cudaStream_t)stream>>>(hull_A, hull_b, facet_off, facet_cnt, xc, P, d, V);
    count_launch();
    PB_CHECK_CUDA(cudaGetLastError());
    return PB200_OK;
}

}  // extern "C"

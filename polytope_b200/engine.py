"""Batched device operations of the LP hot path (host side of the C ABI).

Every function takes either torch CUDA float64 tensors (results stay on the
device, nothing synchronises) or numpy arrays / CPU tensors (inputs are copied
to the device on the current CUDA stream, results come back as numpy arrays).
torch is plumbing only: device memory, streams.  All arithmetic happens in
libpolytope_b200.so; there is no fallback.
"""
import logging
import threading

import numpy as np
import torch

from polytope_b200 import _capi

logger = logging.getLogger(__name__)

F_EMPTY, F_MINREP, F_BBOX, F_LPFAIL = 1, 2, 4, 8
PB200_EUNSUPPORTED = -2         # include/polytope_b200.h
ABS_TOL = 1e-7          # polytope/polytope.py:83


def _require_cuda():
    if not torch.cuda.is_available():
        raise _capi.Pb200Error(
            'polytope_b200 needs a CUDA device (B200, sm_100a); none is visible '
            'and there is no CPU fallback')


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _dev(x, dtype=torch.float64):
    """-> (contiguous CUDA tensor, was_host)."""
    if isinstance(x, torch.Tensor):
        host = not x.is_cuda
        t = x.to(device='cuda', dtype=dtype, non_blocking=True) if host or x.dtype != dtype else x
        return t.contiguous(), host
    a = np.ascontiguousarray(x, dtype={torch.float64: np.float64, torch.int32: np.int32,
                                       torch.int64: np.int64}[dtype])
    return torch.from_numpy(a).to('cuda', non_blocking=True), True


def _opt(x, dtype):
    if x is None:
        return None, 0
    t, _ = _dev(x, dtype)
    return t, t.data_ptr()


def _out(host, *tensors):
    if not host:
        return tensors
    return tuple(t.cpu().numpy() for t in tensors)


LP_MAX_M, LP_MAX_N = 128, 32          # envelope of pb200_lp_batch (lane / warp kernels)
LP_BIG_MAX_N = 64                      # pb200_lp_batch_big


def lp_batch(C, G, H, m_rows=None):
    """B independent LPs min c'x s.t. Gx <= h (solvers.lpsolve, one per row of the batch).

    C[B,n], G[B,m,n] (or G[m,n]: one matrix shared by all B LPs), H[B,m]
    -> status[B] int8, X[B,n], fun[B], iters[B] int32.
    LPs with more than 128 rows or 32 columns (or a shared G) take the one-LP-per-CTA solver.
    """
    _require_cuda()
    lib = _capi.lib()
    G, host = _dev(G)
    C, _ = _dev(C)
    H, _ = _dev(H)
    shared = G.dim() == 2
    B = C.shape[0]
    m, n = G.shape[-2:]
    assert C.shape == (B, n) and H.shape == (B, m) and (shared or G.shape[0] == B), (C.shape, G.shape, H.shape)
    mr, mr_ptr = _opt(m_rows, torch.int32)
    X = torch.empty((B, n), dtype=torch.float64, device='cuda')
    fun = torch.empty(B, dtype=torch.float64, device='cuda')
    status = torch.empty(B, dtype=torch.int8, device='cuda')
    iters = torch.empty(B, dtype=torch.int32, device='cuda')
    if shared or m > LP_MAX_M or n > LP_MAX_N:
        if n > LP_BIG_MAX_N:
            raise _capi.Pb200Error('lp_batch: at most %d columns, got %d' % (LP_BIG_MAX_N, n))
        nbytes = int(lib.pb200_lp_big_workspace_bytes(B, m, n))
        if B and not nbytes:
            raise _capi.Pb200Error('pb200_lp_big_workspace_bytes: unsupported sizes (m=%d, n=%d)' % (m, n))
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device='cuda')
        _capi.check(lib.pb200_lp_batch_big(G.data_ptr(), H.data_ptr(), C.data_ptr(), mr_ptr, B, m, n, int(shared),
                                           X.data_ptr(), fun.data_ptr(), status.data_ptr(), iters.data_ptr(),
                                           ws.data_ptr(), nbytes, _stream()), 'pb200_lp_batch_big')
    else:
        _capi.check(lib.pb200_lp_batch(G.data_ptr(), H.data_ptr(), C.data_ptr(), mr_ptr, B, m, n,
                                       X.data_ptr(), fun.data_ptr(), status.data_ptr(),
                                       iters.data_ptr(), _stream()), 'pb200_lp_batch')
    return _out(host, status, X, fun, iters)


def normalize_batch(A, b, m_rows=None):
    """Polytope.__init__ row normalisation for stacked polytopes -> (An, bn, valid u64 mask)."""
    _require_cuda()
    lib = _capi.lib()
    A, host = _dev(A)
    b, _ = _dev(b)
    P, m, d = A.shape
    mr, mr_ptr = _opt(m_rows, torch.int32)
    An = torch.empty_like(A)
    bn = torch.empty_like(b)
    valid = torch.empty(P, dtype=torch.int64, device='cuda')
    _capi.check(lib.pb200_normalize_batch(A.data_ptr(), b.data_ptr(), mr_ptr, P, m, d,
                                          An.data_ptr(), bn.data_ptr(), valid.data_ptr(),
                                          _stream()), 'pb200_normalize_batch')
    return _out(host, An, bn, valid)


def cheby_batch(A, b, m_rows=None, rows=None):
    """Chebyshev LP of each polytope, rows used as given -> (r, xc, status)."""
    _require_cuda()
    lib = _capi.lib()
    A, host = _dev(A)
    b, _ = _dev(b)
    P, m, d = A.shape
    if m > LP_MAX_M or d + 1 > LP_MAX_N:
        # beyond the lane / warp kernels: the LP of polytope.py:1280-1300 assembled here, solved one per CTA
        if rows is not None:
            raise _capi.Pb200Error('cheby_batch: row masks need m <= 64')
        G = torch.cat([A, torch.sqrt((A * A).sum(-1, keepdim=True))], -1).contiguous()
        C = torch.zeros((P, d + 1), dtype=torch.float64, device='cuda')
        C[:, d] = -1.0
        status, X, _, _ = lp_batch(C, G, b, _dev(m_rows)[0] if m_rows is not None else None)
        return _out(host, X[:, d].contiguous(), X[:, :d].contiguous(), status)
    mr, mr_ptr = _opt(m_rows, torch.int32)
    rw, rw_ptr = _opt(rows, torch.int64)
    r = torch.empty(P, dtype=torch.float64, device='cuda')
    xc = torch.empty((P, d), dtype=torch.float64, device='cuda')
    status = torch.empty(P, dtype=torch.int8, device='cuda')
    _capi.check(lib.pb200_cheby_batch(A.data_ptr(), b.data_ptr(), mr_ptr, rw_ptr, P, m, d,
                                      r.data_ptr(), xc.data_ptr(), status.data_ptr(),
                                      _stream()), 'pb200_cheby_batch')
    return _out(host, r, xc, status)


def bbox_batch(A, b, m_rows=None):
    """bounding_box of each polytope -> (lo[P,d], hi[P,d], status[P,2d])."""
    _require_cuda()
    lib = _capi.lib()
    A, host = _dev(A)
    b, _ = _dev(b)
    P, m, d = A.shape
    if m > LP_MAX_M or d > LP_MAX_N:
        # beyond the lane / warp kernels: 2 d passes of the one-LP-per-CTA solver (every pass one objective for all
        # P polytopes), then the status conventions of polytope.py:1372-1402
        mrd = _dev(m_rows)[0] if m_rows is not None else None
        inf, nan = float('inf'), float('nan')
        lo = torch.empty((P, d), dtype=torch.float64, device='cuda')
        hi = torch.empty((P, d), dtype=torch.float64, device='cuda')
        status = torch.empty((P, 2 * d), dtype=torch.int8, device='cuda')
        for q in range(2 * d):
            j = q % d
            C = torch.zeros((P, d), dtype=torch.float64, device='cuda')
            C[:, j] = 1.0 if q < d else -1.0
            st, X, _, _ = lp_batch(C, A, b, mrd)
            status[:, q] = st
            (lo if q < d else hi)[:, j] = X[:, j]
        sl, su = status[:, :d], status[:, d:]
        lo = torch.where(sl == 3, -inf, torch.where(sl == 2, 0.0, torch.where(sl == 0, lo, nan)))
        hi = torch.where(su == 3, inf, torch.where(su == 2, lo, torch.where(su == 0, hi, nan)))
        return _out(host, lo, hi, status)
    mr, mr_ptr = _opt(m_rows, torch.int32)
    lo = torch.empty((P, d), dtype=torch.float64, device='cuda')
    hi = torch.empty((P, d), dtype=torch.float64, device='cuda')
    status = torch.empty((P, 2 * d), dtype=torch.int8, device='cuda')
    _capi.check(lib.pb200_bbox_batch(A.data_ptr(), b.data_ptr(), mr_ptr, P, m, d,
                                     lo.data_ptr(), hi.data_ptr(), status.data_ptr(),
                                     _stream()), 'pb200_bbox_batch')
    return _out(host, lo, hi, status)


class ReduceResult(object):
    """Outputs of reduce_batch (tensors or numpy arrays, see module docstring).

    keep   int64[P]   bit i set iff input row i is kept (view as uint64)
    flags  int32[P]   F_EMPTY | F_MINREP | F_BBOX | F_LPFAIL
    r, xc             Chebyshev ball of each input polytope
    A, b              constructor-normalised rows; b carries the reference's drift
    n_lp   int32[P]   LPs the reference algorithm needs for this polytope
    lp_iters int32[P] interior-point iterations summed over those LPs
    """
    __slots__ = ('keep', 'flags', 'r', 'xc', 'A', 'b', 'n_lp', 'lp_iters')

    def keep_lists(self):
        keep = np.asarray(self.keep.cpu() if isinstance(self.keep, torch.Tensor) else self.keep)
        m = self.b.shape[1] if self.b is not None else 64
        bits = (keep.astype(np.uint64)[:, None] >> np.arange(m, dtype=np.uint64)) & np.uint64(1)
        return [np.nonzero(row)[0].tolist() for row in bits]


def reduce_batch(A, b, m_rows=None, abs_tol=ABS_TOL, normalize=True, want_A=True, want_b=True,
                 non_empty_bounded=True, results_on_device=False):
    """reduce(Polytope(A[p], b[p])) for every p (polytope.py:1053-1163).

    normalize=False reproduces reduce(poly) on rows that a constructor already
    normalised (poly.A, poly.b are used as they are).  For host-resident batches
    want_A / want_b = False skip the device-to-host copies of the row data (A; b,
    r, xc) when only the keep masks, flags and LP counts are wanted; results_on_device=True
    leaves the results of a host-resident batch on the GPU (for a collective that follows:
    the chunked H2D / kernel overlap stays, the D2H is the caller's).
    """
    _require_cuda()
    lib = _capi.lib()
    if _is_host(A) and len(A) >= 2 * PIPELINE_MIN_CHUNK:
        return _reduce_batch_host_pipelined(A, b, m_rows, abs_tol, normalize, want_A, want_b, non_empty_bounded,
                                            results_on_device)
    A, host = _dev(A)
    host = host and not results_on_device
    b, _ = _dev(b)
    P, m, d = A.shape
    mr, mr_ptr = _opt(m_rows, torch.int32)
    res = ReduceResult()
    keep = torch.empty(P, dtype=torch.int64, device='cuda')
    flags = torch.empty(P, dtype=torch.int32, device='cuda')
    r = torch.empty(P, dtype=torch.float64, device='cuda')
    xc = torch.empty((P, d), dtype=torch.float64, device='cuda')
    b_out = torch.empty((P, m), dtype=torch.float64, device='cuda')
    A_out = torch.empty((P, m, d), dtype=torch.float64, device='cuda') if want_A else None
    n_lp = torch.empty(P, dtype=torch.int32, device='cuda')
    lp_iters = torch.empty(P, dtype=torch.int32, device='cuda')
    ws_bytes = lib.pb200_reduce_workspace_bytes(P, m, d)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device='cuda')
    _capi.check(lib.pb200_reduce_batch(
        A.data_ptr(), b.data_ptr(), mr_ptr, P, m, d, float(abs_tol),
        int(bool(normalize)) | (0 if non_empty_bounded else REDUCE_NO_EARLY_EXIT),
        keep.data_ptr(), flags.data_ptr(), r.data_ptr(), xc.data_ptr(), b_out.data_ptr(),
        A_out.data_ptr() if want_A else 0, n_lp.data_ptr(), lp_iters.data_ptr(), ws.data_ptr(),
        ws_bytes, _stream()),
        'pb200_reduce_batch')
    outs = [keep, flags, r, xc, b_out, n_lp, lp_iters] + ([A_out] if want_A else [])
    outs = _out(host, *outs)
    res.keep, res.flags, res.r, res.xc, res.b, res.n_lp, res.lp_iters = outs[:7]
    res.A = outs[7] if want_A else None
    if logger.isEnabledFor(logging.DEBUG):
        it = np.asarray(res.lp_iters.cpu() if isinstance(res.lp_iters, torch.Tensor) else res.lp_iters)
        nl = np.asarray(res.n_lp.cpu() if isinstance(res.n_lp, torch.Tensor) else res.n_lp)
        per_lp = it / np.maximum(nl, 1)
        logger.debug('reduce_batch: P=%d m=%d d=%d LPs=%d mean ipm iterations/LP=%.2f histogram(per polytope, '
                     'bins 0..10+)=%s', P, m, d, int(nl.sum()), float(it.sum()) / max(int(nl.sum()), 1),
                     np.bincount(np.minimum(per_lp.astype(int), 10), minlength=11).tolist())
    return res


REDUCE_NO_EARLY_EXIT = 2        # bit 1 of pb200_reduce_batch's `normalize` argument (include/polytope_b200.h)
PIPELINE_MIN_CHUNK = 1024      # polytopes per chunk below which pipelining does not pay
PIPELINE_CHUNKS = int(__import__('os').environ.get('PB200_PIPELINE_CHUNKS', '2'))
PIPELINE_FRACTIONS = None      # e.g. (0.25, 0.75): uneven chunk sizes (overrides PIPELINE_CHUNKS)
PINNED_CACHE_BYTES = 1 << 30   # cap on cached page-locked staging memory (per process)
_pinned = {}                   # tag -> flat uint8 pinned buffer (grown geometrically, reused through views)
_pinned_lock = threading.Lock()
_pipeline_lock = threading.Lock()   # the two pipeline streams and the staging buffers are shared
_streams = []


def _is_host(x):
    return isinstance(x, np.ndarray) or (isinstance(x, torch.Tensor) and not x.is_cuda)


def _pinned_buffer(tag, shape, dtype):
    """Reusable pinned host buffer (cudaHostAlloc is far too slow to do per call): one flat
    page-locked allocation per tag, grown geometrically and handed out as a view, so varying
    batch shapes do not accumulate allocations.  Requests that would push the cache past
    PINNED_CACHE_BYTES are served by a one-off pinned allocation that is not kept."""
    shape = tuple(int(v) for v in shape)
    numel = 1
    for v in shape:
        numel *= v
    nbytes = numel * torch.empty((), dtype=dtype).element_size()
    with _pinned_lock:
        flat = _pinned.get(tag)
        if flat is None or flat.numel() < nbytes:
            others = sum(v.numel() for k, v in _pinned.items() if k != tag)
            want = max(nbytes, 2 * flat.numel() if flat is not None else 0)
            if others + want > PINNED_CACHE_BYTES:
                want = nbytes
            if others + want > PINNED_CACHE_BYTES:
                return torch.empty(shape, dtype=dtype).pin_memory()
            flat = torch.empty(max(want, 16), dtype=torch.uint8).pin_memory()
            _pinned[tag] = flat
        return flat[:nbytes].view(dtype).view(shape)


def _reduce_batch_host_pipelined(A, b, m_rows, abs_tol, normalize, want_A, want_b, non_empty_bounded=True,
                                 on_device=False):
    """reduce_batch for host-resident batches: the batch is cut into chunks that
    alternate between two CUDA streams, so the H2D copy of chunk k+1 and the D2H
    of chunk k-1 overlap the kernels of chunk k.  Results land in pinned host
    buffers and come back as numpy arrays."""
    with _pipeline_lock:
        return _reduce_batch_host_pipelined_locked(A, b, m_rows, abs_tol, normalize, want_A, want_b,
                                                   non_empty_bounded, on_device)


def _reduce_batch_host_pipelined_locked(A, b, m_rows, abs_tol, normalize, want_A, want_b, non_empty_bounded,
                                        on_device=False):
    if not _streams:
        _streams.extend([torch.cuda.Stream(), torch.cuda.Stream()])
    At = torch.as_tensor(A)
    bt = torch.as_tensor(b)
    if At.dtype != torch.float64 or not At.is_contiguous():
        At = At.to(torch.float64).contiguous()
    if bt.dtype != torch.float64 or not bt.is_contiguous():
        bt = bt.to(torch.float64).contiguous()
    mt = None if m_rows is None else torch.as_tensor(np.ascontiguousarray(m_rows, dtype=np.int32))
    P, m, d = At.shape
    nchunk = max(2, min(PIPELINE_CHUNKS, P // PIPELINE_MIN_CHUNK))
    if PIPELINE_FRACTIONS:
        # uneven chunks: a small first chunk shortens the only copy nothing overlaps
        cuts = [0] + [int(round(P * f)) for f in np.cumsum(PIPELINE_FRACTIONS)]
        cuts[-1] = P
        bounds = [(lo, hi) for lo, hi in zip(cuts[:-1], cuts[1:]) if hi > lo]
    else:
        bounds = [(P * k // nchunk, P * (k + 1) // nchunk) for k in range(nchunk)]
    names = ['keep', 'flags', 'n_lp', 'lp_iters'] + (['r', 'xc', 'b'] if want_b else []) + (['A'] if want_A else [])
    shapes = {'keep': (P,), 'flags': (P,), 'r': (P,), 'xc': (P, d), 'b': (P, m), 'n_lp': (P,), 'lp_iters': (P,),
              'A': (P, m, d)}
    dtypes = {'keep': torch.int64, 'flags': torch.int32, 'r': torch.float64, 'xc': torch.float64, 'b': torch.float64,
              'n_lp': torch.int32, 'lp_iters': torch.int32, 'A': torch.float64}
    if on_device:
        out = {n: torch.empty(shapes[n], dtype=dtypes[n], device='cuda') for n in names}
    else:
        out = {n: _pinned_buffer('reduce_' + n, shapes[n], dtypes[n]) for n in names}
    cur = torch.cuda.current_stream()
    keepalive = []
    for k, (lo, hi) in enumerate(bounds):
        st = _streams[k % 2]
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            Ad = At[lo:hi].to('cuda', non_blocking=True)
            bd = bt[lo:hi].to('cuda', non_blocking=True)
            md = None if mt is None else mt[lo:hi].to('cuda', non_blocking=True)
            res = reduce_batch(Ad, bd, md, abs_tol=abs_tol, normalize=normalize, want_A=want_A,
                               non_empty_bounded=non_empty_bounded)
            for n in names:
                out[n][lo:hi].copy_(getattr(res, n), non_blocking=True)
            keepalive.append((Ad, bd, md, res))
    for st in _streams:
        cur.wait_stream(st)
    res = ReduceResult()
    if on_device:
        # stream-ordered: the caller's stream waits for both pipeline streams, nothing blocks the host; the
        # chunk buffers are handed to the allocator only after their streams have been joined
        for Ad, bd, md, _ in keepalive:
            for t in (Ad, bd, md):
                if t is not None:
                    t.record_stream(cur)
        for n in ReduceResult.__slots__:
            setattr(res, n, out.get(n))
        return res
    cur.synchronize()
    for n in ReduceResult.__slots__:
        setattr(res, n, out[n].numpy().copy() if n in out else None)
    return res


def adjacent_pairs(A, b, pair_i=None, pair_j=None, abs_tol=ABS_TOL):
    """is_adjacent(cell_i, cell_j) for a list of pairs (polytope.py:1856-1866).

    A[ncell,mc,d], b[ncell,mc] are constructor-normalised cells.  With no pair
    list, all pairs j < i in find_adjacent_regions order (prop2partition.py:57-61).
    -> (adjacent uint8[T], radius[T], status int8[T])
    """
    _require_cuda()
    lib = _capi.lib()
    A, host = _dev(A)
    b, _ = _dev(b)
    ncell, mc, d = A.shape
    pi, pi_ptr = _opt(pair_i, torch.int32)
    pj, pj_ptr = _opt(pair_j, torch.int32)
    T = ncell * (ncell - 1) // 2 if pi is None else int(pi.numel())
    adj = torch.empty(T, dtype=torch.uint8, device='cuda')
    rad = torch.empty(T, dtype=torch.float64, device='cuda')
    status = torch.empty(T, dtype=torch.int8, device='cuda')
    _capi.check(lib.pb200_adjacent_pairs(A.data_ptr(), b.data_ptr(), ncell, mc, d, pi_ptr, pj_ptr,
                                         T, float(abs_tol), adj.data_ptr(), rad.data_ptr(),
                                         status.data_ptr(), _stream()), 'pb200_adjacent_pairs')
    return _out(host, adj, rad, status)


def adjacent_range(A, b, order, t_begin, count, abs_tol=ABS_TOL):
    """is_adjacent over pairs [t_begin, t_begin + count) of an implicit enumeration of the cells'
    pairs -- the block a rank of a sharded partition job owns, no pair lists in memory.
    order 0: t = i (i - 1) / 2 + j, j < i (find_adjacent_regions, prop2partition.py:57-61);
    order 1: all ordered pairs i != j, row-major (compute_adj, prop2partition.py:253-261).
    -> (adjacent uint8[count], radius[count], status int8[count])"""
    _require_cuda()
    lib = _capi.lib()
    A, host = _dev(A)
    b, _ = _dev(b)
    ncell, mc, d = A.shape
    adj = torch.empty(count, dtype=torch.uint8, device='cuda')
    rad = torch.empty(count, dtype=torch.float64, device='cuda')
    status = torch.empty(count, dtype=torch.int8, device='cuda')
    rc = lib.pb200_adjacent_range(A.data_ptr(), b.data_ptr(), ncell, mc, d, int(order), int(t_begin), int(count),
                                  float(abs_tol), adj.data_ptr(), rad.data_ptr(), status.data_ptr(), _stream())
    if rc == PB200_EUNSUPPORTED:
        # cells too large for the one-LP-per-lane kernel: explicit pair lists, warp-per-LP kernel
        from polytope_b200 import sharding
        block = sharding.pair_block if order == 0 else sharding.ordered_pair_block
        pi, pj = block(ncell, t_begin, t_begin + count, device='cuda')
        return adjacent_pairs(A if not host else A.cpu(), b if not host else b.cpu(), pi, pj, abs_tol=abs_tol)
    _capi.check(rc, 'pb200_adjacent_range')
    return _out(host, adj, rad, status)


# ---------------------------------------------------------------------------
# point-set kernels
# ---------------------------------------------------------------------------
def contains_batch(A, b, points, m_rows=None, abs_tol=ABS_TOL, any_of=False):
    """Polytope.contains for P stacked polytopes (polytope.py:206-218), or
    Region.contains (:736-748) with any_of=True.

    points[d, N] are column vectors, as in the reference.
    -> bool[P, N], or bool[N] = OR over the polytopes.
    """
    _require_cuda()
    lib = _capi.lib()
    A, host = _dev(A)
    b, _ = _dev(b)
    pts, phost = _dev(points)
    P, m, d = A.shape
    assert pts.shape[0] == d, (pts.shape, d)
    N = pts.shape[1]
    mr, mr_ptr = _opt(m_rows, torch.int32)
    out = torch.empty((N,) if any_of else (P, N), dtype=torch.uint8, device='cuda')
    _capi.check(lib.pb200_contains_batch(A.data_ptr(), b.data_ptr(), mr_ptr, P, m, d, pts.data_ptr(), N,
                                         float(abs_tol), int(bool(any_of)), out.data_ptr(), _stream()),
                'pb200_contains_batch')
    out = out.view(torch.bool)
    return out.cpu().numpy() if (host and phost) else out


def pcg64_state_words(bit_generator):
    """(state_hi, state_lo, inc_hi, inc_lo) of a numpy PCG64 bit generator."""
    st = bit_generator.state
    if st.get('bit_generator') != 'PCG64':
        raise NotImplementedError('volume(): only numpy PCG64 generators (np.random.default_rng) can be '
                                  'regenerated on the device; got ' + str(st.get('bit_generator')))
    s, inc = int(st['state']['state']), int(st['state']['inc'])
    mask = (1 << 64) - 1
    return [s >> 64, s & mask, inc >> 64, inc & mask]


def volume_counts(A, b, lo, hi, nsamples, rng_words, m_rows=None):
    """Monte-Carlo containment counts of volume() (polytope.py:1583-1591).

    rng_words: uint64[P, 4] from pcg64_state_words, one generator per polytope.
    -> int64[P] counts of samples strictly inside.
    """
    _require_cuda()
    lib = _capi.lib()
    A, host = _dev(A)
    b, _ = _dev(b)
    lo, _ = _dev(lo)
    hi, _ = _dev(hi)
    P, m, d = A.shape
    words = np.ascontiguousarray(np.asarray(rng_words, dtype=np.uint64).reshape(P, 4))
    rw = torch.from_numpy(words.view(np.int64)).to('cuda', non_blocking=True)
    mr, mr_ptr = _opt(m_rows, torch.int32)
    count = torch.empty(P, dtype=torch.int64, device='cuda')
    _capi.check(lib.pb200_volume_counts(A.data_ptr(), b.data_ptr(), mr_ptr, P, m, d, lo.data_ptr(), hi.data_ptr(),
                                        int(nsamples), rw.data_ptr(), count.data_ptr(), _stream()),
                'pb200_volume_counts')
    return count.cpu().numpy() if host else count


def point_facet_sweep(points, normals, offsets, tol=ABS_TOL):
    """quickhull's distance sweep (quickhull.py:117-121, :226-246): points[N, d]
    against facets (normals[F, d], offsets[F]).
    -> (first_facet int32[N], far_facet int32[N], far_dist[N])"""
    _require_cuda()
    lib = _capi.lib()
    pts, host = _dev(points)
    nrm, _ = _dev(normals)
    off, _ = _dev(offsets)
    N, d = pts.shape
    F = nrm.shape[0]
    first = torch.empty(N, dtype=torch.int32, device='cuda')
    far = torch.empty(N, dtype=torch.int32, device='cuda')
    dist = torch.empty(N, dtype=torch.float64, device='cuda')
    _capi.check(lib.pb200_point_facet_sweep(pts.data_ptr(), nrm.data_ptr(), off.data_ptr(), N, F, d, float(tol),
                                            first.data_ptr(), far.data_ptr(), dist.data_ptr(), _stream()),
                'pb200_point_facet_sweep')
    return _out(host, first, far, dist)


# ---------------------------------------------------------------------------
# convex hulls / vertex enumeration
# ---------------------------------------------------------------------------
HULL_OK, HULL_FEW_POINTS, HULL_FLAT, HULL_FACET_CAP, HULL_OUT_CAP, HULL_SINGULAR = range(6)


class HullResult(object):
    """Facets of H hulls in one pool (device tensors, or numpy if the input was host).

    A[F, d], b[F]      unit outer normals and offsets, A x <= b
    vid[F, d]          input indices of the d vertices of each simplicial facet
    facet_off[H], facet_cnt[H]   slice of hull h in the pool
    status[H]          HULL_*; is_vertex[H, Nmax] bool; stats[H, 2] (points inserted, facets created)
    """
    __slots__ = ('A', 'b', 'vid', 'facet_off', 'facet_cnt', 'status', 'is_vertex', 'stats', 'facet_cap')

    def facets(self, h):
        o, c = int(self.facet_off[h]), int(self.facet_cnt[h])
        return self.A[o:o + c], self.b[o:o + c], self.vid[o:o + c]


def _default_facet_cap(nmax, d):
    if d == 2:
        return nmax + 16
    if d == 3:
        return 4 * nmax + 64
    # facets grow like nmax^floor(d/2): start generous (the workspace is (20 d + 24) bytes per
    # slot and CTA), hull_batch retries with 4x on HULL_FACET_CAP
    base = 4096 if d <= 6 else 16384 if d <= 9 else 65536
    return int(max(base, 8 * nmax))


def hull_batch(points, n_pts=None, abs_tol=ABS_TOL, facet_cap=None, out_cap=None, max_tries=6):
    """Convex hulls of H point sets (quickhull.py:141-359): points[H, Nmax, d].

    Capacities are grown and the call repeated when a hull reports
    HULL_FACET_CAP / HULL_OUT_CAP; anything else is returned in `status`.
    """
    _require_cuda()
    lib = _capi.lib()
    pts, host = _dev(points)
    H, nmax, d = pts.shape
    npt, npt_ptr = _opt(n_pts, torch.int32)
    cap = int(facet_cap) if facet_cap else _default_facet_cap(nmax, d)
    pool = int(out_cap) if out_cap else None
    res = HullResult()
    for _ in range(max_tries):
        if pool is None:
            pool = int(min(H * cap, max(1 << 20, (4 << 30) // (8 * (d + 1)))))
        ws_bytes = lib.pb200_hull_workspace_bytes(H, nmax, d, cap)
        if ws_bytes == 0:
            raise _capi.Pb200Error('pb200_hull_workspace_bytes: unsupported sizes (need 2 <= d <= 16)')
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device='cuda')
        A = torch.empty((pool, d), dtype=torch.float64, device='cuda')
        b = torch.empty(pool, dtype=torch.float64, device='cuda')
        vid = torch.empty((pool, d), dtype=torch.int32, device='cuda')
        off = torch.empty(H, dtype=torch.int64, device='cuda')
        cnt = torch.empty(H, dtype=torch.int32, device='cuda')
        status = torch.empty(H, dtype=torch.int32, device='cuda')
        isv = torch.empty((H, nmax), dtype=torch.uint8, device='cuda')
        stats = torch.empty((H, 2), dtype=torch.int32, device='cuda')
        total = torch.zeros(1, dtype=torch.int64, device='cuda')
        _capi.check(lib.pb200_hull_batch(pts.data_ptr(), npt_ptr, H, nmax, d, float(abs_tol), cap, A.data_ptr(),
                                         b.data_ptr(), vid.data_ptr(), pool, off.data_ptr(), cnt.data_ptr(),
                                         status.data_ptr(), isv.data_ptr(), stats.data_ptr(), total.data_ptr(),
                                         ws.data_ptr(), ws_bytes, _stream()), 'pb200_hull_batch')
        st = status.cpu()
        if bool((st == HULL_FACET_CAP).any()):
            cap *= 4
            pool = int(out_cap) if out_cap else None
            continue
        if bool((st == HULL_OUT_CAP).any()):
            pool = int(total.item())
            continue
        break
    else:
        raise _capi.Pb200Error('hull_batch: capacities still too small after %d tries (facet_cap=%d)' % (max_tries, cap))
    n_used = int(total.item())
    outs = _out(host, A[:n_used], b[:n_used], vid[:n_used], off, cnt, status, isv.view(torch.bool), stats)
    res.A, res.b, res.vid, res.facet_off, res.facet_cnt, res.status, res.is_vertex, res.stats = outs
    res.facet_cap = cap
    return res


def dual_points(A, b, xc, m_rows=None):
    """extreme(): Ai = A_i / (b_i - A_i . xc) (polytope.py:1659-1664) -> [P, m, d]."""
    _require_cuda()
    lib = _capi.lib()
    A, host = _dev(A)
    b, _ = _dev(b)
    xc, _ = _dev(xc)
    P, m, d = A.shape
    mr, mr_ptr = _opt(m_rows, torch.int32)
    out = torch.empty_like(A)
    _capi.check(lib.pb200_dual_points(A.data_ptr(), b.data_ptr(), mr_ptr, xc.data_ptr(), P, m, d, out.data_ptr(),
                                      _stream()), 'pb200_dual_points')
    return out.cpu().numpy() if host else out


def dual_facets_to_vertices(hull, xc):
    """extreme(): V = H / K + xc over the facet pool of a dual hull_batch result
    (polytope.py:1665-1676) -> V[F, d] indexed like the pool."""
    _require_cuda()
    lib = _capi.lib()
    HA, host = _dev(hull.A)
    Hb, _ = _dev(hull.b)
    off, _ = _dev(hull.facet_off, torch.int64)
    cnt, _ = _dev(hull.facet_cnt, torch.int32)
    xc, _ = _dev(xc)
    P, d = xc.shape
    V = torch.empty_like(HA)
    max_cnt = int(cnt.max().item()) if P else 0
    for p0 in range(0, P, 65535):
        p1 = min(P, p0 + 65535)
        _capi.check(lib.pb200_dual_facets_to_vertices(HA.data_ptr(), Hb.data_ptr(), off[p0:p1].data_ptr(),
                                                      cnt[p0:p1].data_ptr(), xc[p0:p1].data_ptr(), p1 - p0, d,
                                                      max_cnt, V.data_ptr(), _stream()),
                    'pb200_dual_facets_to_vertices')
    return V.cpu().numpy() if host else V


def extreme_pipeline(A, b, facet_cap=None, out_cap=None):
    """extreme() of P stacked full-dimensional polytopes of dimension d >= 3, device tensors in and
    out (polytope.py:1597-1682): reduce -> Chebyshev centre of the reduced rows -> polar dual ->
    dual hull -> V = H / K + xc.  No host bookkeeping per polytope (polytope.extreme_batch is the
    object-level form with the reference's caching and empty / low-dimensional branches).
    -> (counts int32[P], V[sum(counts), d] in polytope order, hull HullResult, reduce ReduceResult)"""
    _require_cuda()
    A, _ = _dev(A)
    b, _ = _dev(b)
    P, m, d = A.shape
    res = reduce_batch(A, b)
    W = res.keep.shape[1] if res.keep.dim() == 2 else 1
    shifts = torch.arange(m, device='cuda')
    if W == 1:
        bits = ((res.keep.unsqueeze(1) >> shifts) & 1).bool()
    else:
        bits = ((res.keep[:, shifts // 64] >> (shifts % 64)) & 1).bool()
    rows = bits.sum(1).to(torch.int32)
    order = torch.argsort((~bits).to(torch.int8), dim=1, stable=True)       # kept rows first, original order
    Ar = torch.gather(res.A, 1, order.unsqueeze(-1).expand(-1, -1, d)).contiguous()
    br = torch.gather(res.b, 1, order).contiguous()
    r, xc, st = cheby_batch(Ar, br, rows)
    dual = dual_points(Ar, br, xc, rows)
    hull = hull_batch(dual, rows, facet_cap=facet_cap, out_cap=out_cap)
    V = dual_facets_to_vertices(hull, xc)
    return hull.facet_cnt, V, hull, res


# ---------------------------------------------------------------------------
# set difference
# ---------------------------------------------------------------------------
DIFF_PIECES, DIFF_UNTOUCHED, DIFF_COVERED, DIFF_POOL_FULL, DIFF_INDEX_ERROR, DIFF_STEP_LIMIT = range(6)


class DiffResult(object):
    """Pieces of T set differences (device tensors, or numpy if the input was host).

    status[T], n_pieces[T], n_lp[T]; the pool is sorted by (owner, seq):
    piece_off[T] = first piece of problem t; A[F, piece_m, d], b[F, piece_m],
    rows[F], reduce[F] (the reference passes this piece through reduce()).
    """
    __slots__ = ('status', 'n_pieces', 'n_lp', 'piece_off', 'A', 'b', 'rows', 'reduce', 'owner')


def region_diff_batch(PA, Pb, RA, Rb, p_rows=None, r_rows=None, n_reg=None, abs_tol=ABS_TOL,
                      intersect_tol=ABS_TOL, piece_cap=None, max_tries=16):
    """poly_t \\ region_t for T problems (polytope.py:2117-2282).

    PA[T, mp, d], Pb[T, mp]; RA[T, Nr, mr, d], Rb[T, Nr, mr] -- or RA[Nr, mr, d],
    Rb[Nr, mr] for one region shared by all problems.
    """
    _require_cuda()
    lib = _capi.lib()
    PA, host = _dev(PA)
    Pb, _ = _dev(Pb)
    RA, _ = _dev(RA)
    Rb, _ = _dev(Rb)
    T, mp, d = PA.shape
    shared = RA.dim() == 3
    Nr, mr = (RA.shape[0], RA.shape[1]) if shared else (RA.shape[1], RA.shape[2])
    pr, pr_ptr = _opt(p_rows, torch.int32)
    rr, rr_ptr = _opt(r_rows, torch.int32)
    nr, nr_ptr = _opt(n_reg, torch.int32)
    piece_m = min(128, mp + 2 * Nr * mr)
    cap = int(piece_cap) if piece_cap else max(4 * T, 1024)
    status = torch.empty(T, dtype=torch.int32, device='cuda')
    npieces = torch.empty(T, dtype=torch.int32, device='cuda')
    nlp = torch.empty(T, dtype=torch.int32, device='cuda')
    used = torch.zeros(1, dtype=torch.int64, device='cuda')
    counter = torch.empty(1, dtype=torch.int32, device='cuda')
    for _ in range(max_tries):
        pA = torch.empty((cap, piece_m, d), dtype=torch.float64, device='cuda')
        pb = torch.empty((cap, piece_m), dtype=torch.float64, device='cuda')
        prow = torch.empty(cap, dtype=torch.int32, device='cuda')
        pred = torch.empty(cap, dtype=torch.int32, device='cuda')
        pown = torch.empty(cap, dtype=torch.int32, device='cuda')
        pseq = torch.empty(cap, dtype=torch.int32, device='cuda')
        _capi.check(lib.pb200_region_diff_batch(
            PA.data_ptr(), Pb.data_ptr(), pr_ptr, T, mp, d, RA.data_ptr(), Rb.data_ptr(), rr_ptr, nr_ptr,
            int(shared), Nr, mr, float(abs_tol), float(intersect_tol), pA.data_ptr(), pb.data_ptr(),
            prow.data_ptr(), pred.data_ptr(), pown.data_ptr(), pseq.data_ptr(), cap, piece_m, used.data_ptr(),
            status.data_ptr(), npieces.data_ptr(), nlp.data_ptr(), counter.data_ptr(), _stream()),
            'pb200_region_diff_batch')
        n_used = int(used.item())
        if n_used <= cap:
            break
        # a search stops at its first overflowing piece, so `used` under-reports the need
        # (cap + one per overflowed problem): grow geometrically, not to `used`
        logger.debug('region_diff_batch: piece pool of %d overflowed (%d requested), retrying with %d',
                     cap, n_used, max(n_used, 4 * cap))
        cap = max(n_used, 4 * cap)
        del pA, pb, prow, pred, pown, pseq
    else:
        raise _capi.Pb200Error('region_diff_batch: piece pool still too small after %d tries (cap=%d)'
                               % (max_tries, cap))
    # pool order is arrival order: sort by (owner, seq)
    key = pown[:n_used].to(torch.int64) * (1 << 31) + pseq[:n_used].to(torch.int64)
    order = torch.argsort(key)
    res = DiffResult()
    ok_pieces = torch.where(status == DIFF_POOL_FULL, torch.zeros_like(npieces), npieces)
    off = torch.cumsum(ok_pieces.to(torch.int64), 0) - ok_pieces.to(torch.int64)
    outs = _out(host, status, npieces, nlp, off, pA[:n_used][order], pb[:n_used][order], prow[:n_used][order],
                pred[:n_used][order], pown[:n_used][order])
    (res.status, res.n_pieces, res.n_lp, res.piece_off, res.A, res.b, res.rows, res.reduce, res.owner) = outs
    return res


REDUCE_STAGES = ('normalize', 'cheby_lp', 'prefilter', 'bbox_lp', 'candidates', 'row_lp', 'finalize')


def profile_enable(on=True):
    """Record CUDA events around the stages of the next reduce_batch calls."""
    _capi.lib().pb200_profile_enable(int(bool(on)))


def profile_read():
    """-> dict stage -> milliseconds of the last profiled reduce_batch call."""
    import ctypes
    buf = (ctypes.c_float * len(REDUCE_STAGES))()
    _capi.check(_capi.lib().pb200_profile_read(ctypes.cast(buf, ctypes.c_void_p), len(REDUCE_STAGES)),
                'pb200_profile_read')
    return dict(zip(REDUCE_STAGES, [float(v) for v in buf]))


def launch_count():
    return int(_capi.lib().pb200_launch_count())


def measure_dfma_tflops():
    """fp64 FMA peak of the current device, measured now (TFLOP/s)."""
    import ctypes
    _require_cuda()
    n = 1024 * torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    scratch = torch.empty(n, dtype=torch.float64, device='cuda')
    out = ctypes.c_double(0.0)
    _capi.check(_capi.lib().pb200_measure_dfma_tflops(scratch.data_ptr(), n, ctypes.cast(ctypes.byref(out), ctypes.c_void_p),
                                                      _stream()), 'pb200_measure_dfma_tflops')
    return float(out.value)


def lane_solver(on=True):
    """A/B switch: reduce / bounding-box LPs (d <= 8, m <= 64) on the one-LP-per-lane solver
    (default) or on the warp-per-LP kernels."""
    _capi.lib().pb200_lane_solver(int(bool(on)))


if __import__('os').environ.get('PB200_LANE_SOLVER', '') == '0':
    lane_solver(False)

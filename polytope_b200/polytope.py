"""`Polytope` / `Region` shell and the LP-backed set operations of the hot path.

Host-side mirror of the part of /root/reference/polytope/polytope.py that
SURVEY.md section 8(a) puts on the hot path: the constructor normalisation and
caches, `is_empty`, `is_fulldim`, `cheby_ball`, `bounding_box`, `reduce`,
`Polytope.intersect`, `is_adjacent`, and of the next rows of SURVEY.md 8(f):
`contains`, `volume`, `grid_region`, `enumerate_integral_points`, `qhull`,
`extreme`.  Names, argument meaning, caching side effects and error behaviour
follow the reference; every LP, hull and point sweep goes to the sm_100a
kernels through `polytope_b200.engine` (no scipy, no CPU fallback).

Each single-object function has a batched sibling (`*_batch`) that runs the
whole list in a handful of kernel launches -- that is the form the throughput
numbers are quoted on.  Operations outside the hot path (projection, esp,
rotation, plotting) are not provided.

Attribution.  The class shells and the sequential host logic that SURVEY.md 8(f)
leaves on the host restate code of tulip-control/polytope (polytope/polytope.py,
Copyright (c) 2011-2014 California Institute of Technology, BSD 3-clause; see
the reference's LICENSE): `is_convex`, `union`, `is_interior`, `grid_region`,
`enumerate_integral_points`, the d <= 2 branch of `extreme`
(`_extreme_low_dim`), the `overlap=False` branch of `is_adjacent`, and the step
sequence of `reduce` in `_reduce_wide` follow the reference line by line (each
cites its line range) because the API contract is "the same results for the same
calls".  Redistribution of this file is under the terms of that licence for
those parts.  Everything that carries the design -- the `*_batch` functions,
`region_diff_batch`, `extreme_batch`, `separate`, the device pipelines they
drive -- is original to this repository.
"""
import logging
import math
import warnings

import numpy as np

from polytope_b200 import engine
from polytope_b200.solvers import lpsolve  # noqa: F401  (bound at import, as polytope.py:69)

logger = logging.getLogger(__name__)

np.set_printoptions(precision=5, suppress=True)   # polytope.py:78, pinned by test_polytope_str

ABS_TOL = 1e-7                                     # polytope.py:83


class Polytope(object):
    """Convex polytope {x : A x <= b} in half-space representation.

    Same constructor contract as the reference (polytope.py:122-148): arrays are
    cast to float, and unless `normalize=False` every row is scaled to unit
    2-norm and rows with norm <= 1e-10 are dropped.
    """

    def __init__(self, A=np.array([]), b=np.array([]), minrep=False, chebR=0, chebX=None,
                 fulldim=None, volume=None, vertices=None, normalize=True):
        self.A = A.astype(float)
        self.b = b.astype(float).flatten()
        if A.size > 0 and normalize:
            norms = np.sqrt(np.sum(A * A, 1)).flatten()
            pos = np.nonzero(norms > 1e-10)[0]
            mult = 1 / norms[pos]
            self.A = self.A[pos, :] * mult[:, np.newaxis]
            self.b = self.b[pos] * mult
        self.minrep = minrep
        self._chebXc = chebX
        self._chebR = chebR
        self.bbox = None
        self.fulldim = fulldim
        self._volume = None if volume is None else float(volume)
        self.vertices = vertices

    def __str__(self):
        A_rows = str(self.A).split('\n')
        b_rows = str(self.b.reshape(self.b.shape[0], 1)).split('\n')
        n = len(A_rows)
        mid = int((n - 1) / 2)
        sep = [' |    '] * mid + [' x <= '] + [' |    '] * (n - mid - 2) + (['|    '] if n > 1 else [])
        lines = [A_rows[k] + sep[k] + b_rows[k] for k in range(n)]
        return 'Single polytope \n  {lines}\n'.format(lines='\n  '.join(lines))

    def __len__(self):
        return 0

    def __copy__(self):
        P = Polytope(self.A.copy(), self.b.copy())
        P._chebXc, P._chebR = self._chebXc, self._chebR
        P.minrep, P.bbox, P.fulldim = self.minrep, self.bbox, self.fulldim
        return P

    copy = __copy__

    def __contains__(self, point):
        point = np.asarray(point)
        return np.all(self.A.dot(point.flatten()) - self.b < ABS_TOL)

    def contains(self, points, abs_tol=ABS_TOL):
        """Boolean array: which column vectors of `points` satisfy A x - b < abs_tol
        (polytope.py:206-218), one thread per point on the device."""
        points = np.asarray(points, dtype=float)
        if points.ndim != 2 or self.A.size == 0:
            return np.all(self.A.dot(points) - self.b[:, np.newaxis] < abs_tol, axis=0)
        return engine.contains_batch(self.A[None], self.b[None], points, abs_tol=abs_tol)[0]

    @property
    def volume(self):
        if self._volume is None:
            self._volume = volume(self)
        return self._volume

    def _set_volume(self, polytope_volume):
        if polytope_volume < 0.0:
            raise ValueError('`polytope_volume` must be >= 0, given:  {v}'.format(v=polytope_volume))
        self._volume = float(polytope_volume)

    def intersect(self, other, abs_tol=ABS_TOL):
        """Intersection with another Polytope (polytope.py:255-275)."""
        if isinstance(other, Region):
            return other.intersect(self, abs_tol=abs_tol)
        if not isinstance(other, Polytope):
            raise Exception('Polytope intersection defined only with other Polytope. '
                            'Got instead: ' + str(type(other)))
        if (not is_fulldim(self)) or (not is_fulldim(other)):
            return Polytope()
        if self.dim != other.dim:
            raise Exception("polytopes have different dimension")
        iA = np.vstack([self.A, other.A])
        ib = np.hstack([self.b, other.b])
        return reduce(Polytope(iA, ib), abs_tol=abs_tol)

    def __eq__(self, other):
        return self <= other and other <= self

    def __ne__(self, other):
        return not self == other

    def __le__(self, other):
        return is_subset(self, other)

    def __ge__(self, other):
        return is_subset(other, self)

    __hash__ = object.__hash__

    def __bool__(self):
        return bool(self.volume > 0)

    __nonzero__ = __bool__

    def union(self, other, check_convex=False):
        return union(self, other, check_convex)

    def diff(self, other):
        return mldivide(self, other)

    @classmethod
    def from_box(cls, intervals=[]):
        """Hyperrectangle from [[x0_min, x0_max], ...] (polytope.py:311-354)."""
        if not isinstance(intervals, np.ndarray):
            try:
                intervals = np.array(intervals)
            except Exception:
                raise Exception('Polytope.from_box:intervals must be a numpy ndarray or '
                                'convertible as arg to numpy.array')
        if intervals.ndim != 2:
            raise Exception('Polytope.from_box: intervals must be 2 dimensional')
        if intervals.shape[1] != 2:
            raise Exception('Polytope.from_box: intervals must have 2 columns')
        if (intervals[:, 0] > intervals[:, 1]).any():
            raise Exception('Polytope.from_box: Invalid interval in from_box method.\n'
                            'First element of an interval must not be larger than the second.')
        n = intervals.shape[0]
        A = np.vstack([np.eye(n), -np.eye(n)])
        b = np.hstack([intervals[:, 1], -intervals[:, 0]])
        return cls(A, b, minrep=True)

    def scale(self, factor):
        self.b = factor * self.b

    @property
    def dim(self):
        try:
            return np.shape(self.A)[1]
        except Exception:
            return 0.0

    @property
    def chebR(self):
        cheby_ball(self)
        return self._chebR

    @property
    def chebXc(self):
        cheby_ball(self)
        return self._chebXc

    @property
    def cheby(self):
        return cheby_ball(self)

    @property
    def bounding_box(self):
        if self.bbox is None:
            self.bbox = bounding_box(self)
        return self.bbox


class Region(object):
    """Possibly non-convex set: a list of Polytopes plus `props` (polytope.py:650-703)."""

    def __init__(self, list_poly=None, props=None):
        if list_poly is None:
            list_poly = []
        if props is None:
            props = set()
        if isinstance(list_poly, str):
            self.list_poly = list_poly
            self.props = set(props)
            return
        if isinstance(list_poly, Region):
            dim = list_poly[0].dim
            for poly in list_poly:
                if poly.dim != dim:
                    raise Exception("Region error: Polytopes must be of same dimension!")
        self.list_poly = [p for p in list_poly if not is_empty(p)]
        self.props = set(props)
        self.bbox = None
        self.fulldim = None
        self._volume = None
        self._chebXc = None
        self._chebR = None

    def __iter__(self):
        return iter(self.list_poly)

    def __getitem__(self, key):
        return self.list_poly[key]

    def __len__(self):
        return len(self.list_poly)

    def __contains__(self, point):
        point = np.asarray(point)
        return any(point in u for u in self.list_poly)

    def contains(self, points, abs_tol=ABS_TOL):
        points = np.asarray(points)
        if points.shape[0] != self.dim:
            raise ValueError('points should be column vectors')
        if len(self.list_poly) == 0:
            return np.full(points.shape[1], False, dtype=bool)
        # one pass over the points for all member polytopes (polytope.py:742-747)
        A, b, rows = _stack(self.list_poly)
        return engine.contains_batch(A, b, np.asarray(points, dtype=float), rows, abs_tol=abs_tol, any_of=True)

    @property
    def volume(self):
        if self._volume is None:
            self._volume = volume(self)
        return self._volume

    def _set_volume(self, region_volume):
        if region_volume < 0.0:
            raise ValueError('`region_volume` must be >= 0, given:  {v}'.format(v=region_volume))
        self._volume = float(region_volume)

    def __eq__(self, other):
        return self <= other and other <= self

    def __ne__(self, other):
        return not self == other

    def __le__(self, other):
        return is_subset(self, other)

    def __ge__(self, other):
        return is_subset(other, self)

    __hash__ = object.__hash__

    def __add__(self, other):
        """Union with convex simplification (polytope.py:762-775)."""
        return union(self, other, check_convex=True)

    def __bool__(self):
        return bool(self.volume > 0)

    __nonzero__ = __bool__

    def union(self, other, check_convex=False):
        return union(self, other, check_convex)

    def __sub__(self, other):
        return mldivide(self, other)

    def diff(self, other):
        return mldivide(self, other)

    def __and__(self, other):
        return intersect(self, other)

    def intersect(self, other, abs_tol=ABS_TOL):
        """Intersection with a Polytope or an iterable of Polytopes (polytope.py:815-830):
        all pairwise intersections in one device pass, then the reference's union
        accumulation."""
        if isinstance(other, Polytope):
            other = [other]
        pairs = [(poly0, poly1) for poly0 in self for poly1 in other]
        P = Region()
        if not pairs:
            return P
        isects = intersect_batch([a for a, _ in pairs], [b for _, b in pairs], abs_tol)
        balls = cheby_ball_batch(isects)
        for isect, (rp, _) in zip(isects, balls):
            if rp > abs_tol:
                P = union(P, isect, check_convex=True)
        return P

    def __copy__(self):
        return Region(list_poly=self.list_poly[:], props=self.props.copy())

    copy = __copy__

    @property
    def dim(self):
        return np.shape(self.list_poly[0].A)[1]

    @property
    def chebR(self):
        cheby_ball(self)
        return self._chebR

    @property
    def chebXc(self):
        cheby_ball(self)
        return self._chebXc

    @property
    def cheby(self):
        return cheby_ball(self)

    @property
    def bounding_box(self):
        if self.bbox is None:
            self.bbox = bounding_box(self)
        return self.bbox


def box2poly(box):
    """Hyperrectangle Polytope from [[x0_min, x0_max], ...] (polytope.py:2285-2298)."""
    return Polytope.from_box(box)


def _bounding_box_to_polytope(lower, upper):
    """Polytope of the box with the given corner columns (polytope.py:1303-1311)."""
    return box2poly([(a[0], b[0]) for a, b in zip(lower, upper)])


def is_empty(polyreg):
    """Structural emptiness, no LP (polytope.py:939-959)."""
    n = len(polyreg)
    if n == 0:
        try:
            return len(polyreg.A) == 0
        except Exception:
            return True
    return bool(np.all([is_empty(p) for p in polyreg.list_poly]))


# ---------------------------------------------------------------------------
# batching helpers
# ---------------------------------------------------------------------------
def _stack(polys):
    """Pad a list of same-dimension polytopes to one (A[P,m,d], b[P,m], m_rows[P])."""
    d = polys[0].A.shape[1]
    rows = np.array([p.A.shape[0] for p in polys], dtype=np.int32)
    m = max(int(rows.max()), 1)
    A = np.zeros((len(polys), m, d))
    b = np.zeros((len(polys), m))
    for k, p in enumerate(polys):
        if p.A.shape[1] != d:
            raise Exception("polytopes have different dimension")
        A[k, :rows[k]] = p.A
        b[k, :rows[k]] = p.b
    return A, b, rows


def _store_cheby(poly, status, r, xc):
    """Caching side effects of cheby_ball (polytope.py:1289-1300)."""
    if status == 0 and not r < 0:
        poly._chebXc = np.array(xc)
        poly._chebR = np.double(r)
        return poly._chebR, poly._chebXc
    return 0, None


def cheby_ball_batch(polys):
    """[cheby_ball(p) for p in polys] with one kernel launch for the uncached ones."""
    out = [None] * len(polys)
    todo = []
    for k, p in enumerate(polys):
        if (p._chebXc is not None) and (p._chebR is not None):
            out[k] = (p._chebR, p._chebXc)
        elif isinstance(p, Region):
            out[k] = cheby_ball(p)
        elif is_empty(p):
            out[k] = (0, None)
        else:
            todo.append(k)
    if todo:
        A, b, rows = _stack([polys[k] for k in todo])
        r, xc, status = engine.cheby_batch(A, b, rows)
        for t, k in enumerate(todo):
            out[k] = _store_cheby(polys[k], int(status[t]), r[t], xc[t])
    return out


def is_fulldim_batch(polys, abs_tol=ABS_TOL):
    """[is_fulldim(p, abs_tol) for p in polys], batched."""
    need = [p for p in polys if p.fulldim is None and not isinstance(p, Region)]
    balls = dict(zip(map(id, need), cheby_ball_batch(need)))
    out = []
    for p in polys:
        if p.fulldim is None:
            if isinstance(p, Region):
                is_fulldim(p, abs_tol)
            else:
                p.fulldim = balls[id(p)][0] > abs_tol
        out.append(p.fulldim)
    return out


def bounding_box_batch(polys):
    """[bounding_box(p) for p in polys] with 2*d LPs per polytope in one launch."""
    todo = [k for k, p in enumerate(polys) if p.bbox is None]
    if todo:
        A, b, rows = _stack([polys[k] for k in todo])
        lo, hi, status = engine.bbox_batch(A, b, rows)
        for t, k in enumerate(todo):
            bad = np.nonzero((status[t] == 1) | (status[t] == 4))[0]
            if len(bad):
                raise RuntimeError('bounding_box: `polytope_b200.solvers.lpsolve` returned status '
                                   '{v} for LP {i}'.format(v=int(status[t][bad[0]]), i=int(bad[0])))
            polys[k].bbox = lo[t].reshape(-1, 1).copy(), hi[t].reshape(-1, 1).copy()
    return [p.bbox for p in polys]


REDUCE_MAX_ROWS, REDUCE_MAX_DIM = 64, 31      # envelope of pb200_reduce_batch (row sets are 64-bit masks)


def _reduce_wide(poly, nonEmptyBounded, abs_tol):
    """reduce() of one polytope outside the envelope of the fused device pipeline (more than 64 rows or 31
    columns).  The steps of polytope.py:1081-1163 on the host, every LP on the device: the Chebyshev LP of
    is_fulldim, the 2 d bounding-box LPs, then ALL row LPs in one call that shares G (one LP per CTA for LPs
    of more than 128 rows).  The in-place `h[k] += 0.1 ... h[k] -= 0.1` of the reference leaves rows before k
    one rounding away from b (:1147-1149); the right-hand sides below carry exactly that."""
    if not is_fulldim(poly):
        return Polytope()
    finite = poly.b != np.inf                                   # :1087-1089
    R, rhs = poly.A[finite], poly.b[finite]
    # duplicate directions (:1094-1109).  The reference scans all pairs with a 1-D np.dot; here one matrix
    # product proposes the pairs and the reference's own dot confirms each, so the removed set is identical.
    scale = 1 / np.sqrt(np.sum(R.T**2, 0))
    unit = np.dot(R.T, np.diag(scale)).T
    proposed = np.triu(unit.dot(unit.T) > 1 - abs_tol - 1e-9, 1)
    drop = set()
    for i, j in zip(*np.nonzero(proposed)):
        if np.dot(unit[i].T, unit[j]) > 1 - abs_tol:
            drop.add(j if rhs[i] * scale[i] < rhs[j] * scale[j] else i)
    left = [k for k in range(len(rhs)) if k not in drop]
    R, rhs = R[left], rhs[left]
    nrow, nx = R.shape
    small = lambda count: bool(nonEmptyBounded) and count <= nx + 1      # noqa: E731  (:1113-1116, :1135-1138)
    if small(nrow):
        return Polytope(R, rhs)
    if nrow > 3 * nx:                                            # rows that cannot touch the bounding box (:1118-1134)
        lo, hi = Polytope(R, rhs).bounding_box
        reach = np.dot((R > 0) * R, hi - lo) - (np.array([rhs]).T - np.dot(R, lo))
        inside = ~(reach < -1e-4)
        R, rhs = R[inside.squeeze()], rhs[inside.squeeze()]
        nrow = R.shape[0]
    if small(nrow):
        return Polytope(R, rhs)
    # one LP per remaining row, all in calls that share G (:1142-1160): objective -R[k], row k relaxed by 0.1;
    # rows before k already carry the (b + 0.1) - 0.1 rounding of the reference's in-place update
    R = np.ascontiguousarray(R)
    settled = (rhs + 0.1) - 0.1
    kept = []
    per_call = max(1, (64 << 20) // (8 * nrow))                 # right-hand sides of one call: <= 64 MB
    for first in range(0, nrow, per_call):
        ks = np.arange(first, min(nrow, first + per_call))
        H = np.where(np.arange(nrow)[None, :] < ks[:, None], settled[None, :], rhs[None, :])
        H[np.arange(len(ks)), ks] = rhs[ks] + 0.1
        status, _, fun, _ = engine.lp_batch(-R[ks], R, H)
        for t, k in enumerate(ks):
            if status[t] == 3 or (status[t] == 0 and -fun[t] - settled[k] > abs_tol):
                kept.append(int(k))
    out = Polytope(R[kept], settled[kept])
    out.minrep = True
    return out


def reduce_batch(polys, abs_tol=ABS_TOL, nonEmptyBounded=1):
    """[reduce(p, nonEmptyBounded, abs_tol) for p in polys] through the device pipeline."""
    out = [None] * len(polys)
    todo = []
    for k, p in enumerate(polys):
        if isinstance(p, Region):
            out[k] = reduce(p, abs_tol=abs_tol)
        elif p.minrep:
            out[k] = p
        elif p.fulldim is False or is_empty(p):
            out[k] = Polytope()
        elif p.A.shape[0] > REDUCE_MAX_ROWS or p.A.shape[1] > REDUCE_MAX_DIM:
            out[k] = _reduce_wide(p, nonEmptyBounded, abs_tol)
        else:
            todo.append(k)
    if todo:
        A, b, rows = _stack([polys[k] for k in todo])
        res = engine.reduce_batch(A, b, rows, abs_tol=abs_tol, normalize=False,
                                  non_empty_bounded=bool(nonEmptyBounded))
        keeps = res.keep_lists()
        for t, k in enumerate(todo):
            p = polys[k]
            if res.flags[t] & engine.F_LPFAIL and not (res.flags[t] & engine.F_EMPTY):
                # reduce() -> Polytope(A_arr, b_arr).bounding_box raises on LP status 1 / 4
                # (polytope.py:1382, :1404); a failed row LP is dropped silently there (:1152-1160),
                # but it is surfaced here rather than returning a polytope built on a failed solve
                raise RuntimeError('reduce: `polytope_b200.solvers.lpsolve` did not converge for an LP of '
                                   'polytope {k} of the batch (status 1 or 4)'.format(k=k))
            if p.fulldim is None:      # is_fulldim(poly) side effects, polytope.py:1081
                rr = res.r[t]
                _store_cheby(p, 0 if rr == rr else 4, rr, res.xc[t])
                p.fulldim = bool(rr > ABS_TOL)
            if res.flags[t] & engine.F_EMPTY:
                out[k] = Polytope()
                continue
            red = Polytope(res.A[t][keeps[t]], res.b[t][keeps[t]])
            red.minrep = bool(res.flags[t] & engine.F_MINREP)
            out[k] = red
    return out


# ---------------------------------------------------------------------------
# reference-named single-object functions
# ---------------------------------------------------------------------------
def is_fulldim(polyreg, abs_tol=ABS_TOL):
    """True if the polytope / region has interior points (polytope.py:962-985)."""
    if polyreg.fulldim is not None:
        return polyreg.fulldim
    if len(polyreg) == 0:
        rc, _ = cheby_ball(polyreg)
        status = rc > abs_tol
    else:
        balls = cheby_ball_batch(polyreg.list_poly)
        status = bool(np.sum([rc > abs_tol for rc, _ in balls]) > 0)
    polyreg.fulldim = status
    return status


def cheby_ball(poly1):
    """Chebyshev radius and (a) centre (polytope.py:1241-1300)."""
    if (poly1._chebXc is not None) and (poly1._chebR is not None):
        return poly1._chebR, poly1._chebXc
    if isinstance(poly1, Region):
        maxr, maxx = 0, None
        for rc, xc in cheby_ball_batch(poly1.list_poly):
            if rc > maxr:
                maxr, maxx = rc, xc
        poly1._chebXc, poly1._chebR = maxx, maxr
        return maxr, maxx
    if is_empty(poly1):
        return 0, None
    r, xc, status = engine.cheby_batch(poly1.A[None], poly1.b[None])
    return _store_cheby(poly1, int(status[0]), r[0], xc[0])


def bounding_box(polyreg):
    """Smallest hyperbox (l, u) containing the polytope / region (polytope.py:1314-1411)."""
    if polyreg.bbox is not None:
        return polyreg.bbox
    if isinstance(polyreg, Region):
        boxes = bounding_box_batch(polyreg.list_poly)
        l = np.min(np.hstack([bb[0] for bb in boxes]), axis=1).reshape(-1, 1)
        u = np.max(np.hstack([bb[1] for bb in boxes]), axis=1).reshape(-1, 1)
        polyreg.bbox = l, u
        return l, u
    return bounding_box_batch([polyreg])[0]


def reduce(poly, nonEmptyBounded=1, abs_tol=ABS_TOL):
    """Remove redundant inequalities (polytope.py:1053-1163), one LP per facet."""
    if isinstance(poly, Region):
        lst = [red for red in reduce_batch(poly.list_poly) if is_fulldim(red)]
        if len(lst) > 0:
            return Region(lst, poly.props)
        return Polytope()
    return reduce_batch([poly], abs_tol=abs_tol, nonEmptyBounded=nonEmptyBounded)[0]


def intersect(poly1, poly2, abs_tol=ABS_TOL):
    """Intersection of two polytopes or regions (polytope.py:1508-1526)."""
    if isinstance(poly1, Region):
        return poly1.intersect(poly2, abs_tol=abs_tol)
    if isinstance(poly2, Region):
        return poly2.intersect(poly1, abs_tol=abs_tol)
    if not isinstance(poly1, Polytope):
        raise Exception('poly1 not Region nor Polytope.Got instead: ' + str(type(poly1)))
    return poly1.intersect(poly2, abs_tol)


def intersect_batch(polys1, polys2, abs_tol=ABS_TOL):
    """[p.intersect(q) for p, q in zip(polys1, polys2)]: the pairwise, batchable
    part of Region.intersect (polytope.py:823-826)."""
    fd1 = is_fulldim_batch(polys1)
    fd2 = is_fulldim_batch(polys2)
    stacked, where = [], []
    out = [None] * len(polys1)
    for k, (p, q) in enumerate(zip(polys1, polys2)):
        if not (fd1[k] and fd2[k]):
            out[k] = Polytope()
            continue
        if p.dim != q.dim:
            raise Exception("polytopes have different dimension")
        stacked.append(Polytope(np.vstack([p.A, q.A]), np.hstack([p.b, q.b])))
        where.append(k)
    for k, red in zip(where, reduce_batch(stacked, abs_tol=abs_tol)):
        out[k] = red
    return out


def is_adjacent(poly1, poly2, overlap=True, abs_tol=ABS_TOL):
    """True if the two polytopes / regions touch or overlap (polytope.py:1827-1885)."""
    if poly1.dim != poly2.dim:
        raise Exception("is_adjacent: polytopes do not have the same dimension")
    if isinstance(poly1, Region):
        return any(is_adjacent(p, poly2, overlap=overlap, abs_tol=abs_tol) for p in poly1)
    if isinstance(poly2, Region):
        return any(is_adjacent(poly1, p, overlap=overlap, abs_tol=abs_tol) for p in poly2)
    b1_arr = poly1.b.copy()
    b2_arr = poly2.b.copy()
    if overlap:
        b1_arr += abs_tol
        b2_arr += abs_tol
    else:
        M1 = np.concatenate((poly1.A, np.array([poly1.b]).T), 1).T
        M1n = np.dot(M1, np.diag(1 / np.sqrt(np.sum(M1**2, 0))))
        M2 = np.concatenate((poly2.A, np.array([poly2.b]).T), 1).T
        M2n = np.dot(M2, np.diag(1 / np.sqrt(np.sum(M2**2, 0))))
        prod = np.dot(M1n.T, M2n)
        if not np.any(prod < -0.99):
            return False
        row, col = np.nonzero(np.isclose(prod, prod.min()))
        for i, j in zip(row, col):
            b1_arr[i] += abs_tol
            b2_arr[j] += abs_tol
    dummy = Polytope(np.concatenate((poly1.A, poly2.A)), np.concatenate((b1_arr, b2_arr)))
    return is_fulldim(dummy, abs_tol=abs_tol / 10)


def adjacency_matrix(cells, abs_tol=ABS_TOL):
    """Dense int8 adjacency of a list of single Polytopes -- the loop of
    find_adjacent_regions (prop2partition.py:46-63) as one kernel launch of
    n(n-1)/2 Chebyshev LPs (cells with fewer rows are padded with zero rows,
    which the constructor normalisation inside the kernel drops, polytope.py:130)."""
    n = len(cells)
    adj = np.eye(n, dtype=np.int8)
    if n < 2:
        return adj
    i, j = np.tril_indices(n, -1)
    if max(c.A.shape[0] for c in cells) > 32:
        # cells with more rows than the pair kernel stacks (2 x 32): the stacked polytopes of is_adjacent
        # (polytope.py:1856-1866) built here, one batch of Chebyshev LPs
        pairs = [Polytope(np.vstack([cells[a].A, cells[c].A]), np.hstack([cells[a].b, cells[c].b]) + abs_tol)
                 for a, c in zip(i, j)]
        flags = np.array([rc > abs_tol / 10 for rc, _ in cheby_ball_batch(pairs)], dtype=np.int8)
        adj[i, j] = flags
        adj[j, i] = flags
        return adj
    A, b, _ = _stack(cells)
    flags, _, _ = engine.adjacent_pairs(A, b, abs_tol=abs_tol)
    adj[i, j] = flags
    adj[j, i] = flags
    return adj


def separate(reg1, abs_tol=ABS_TOL):
    """Divide a Region into connected Regions (polytope.py:1795-1824).

    The reference grows each part with one is_adjacent(part, polytope) call per
    remaining polytope -- O(n^2) sequential LPs.  Here all n(n-1)/2 pair LPs run
    as one `adjacency_matrix` launch and the reference's grouping is replayed on
    the flags: a single ordered pass per part (a polytope skipped before the
    part reached it is not revisited, so a part is not always a full connected
    component -- same as the reference), members in the order they were added,
    props copied.  As in the reference (:1815) the pair test runs at the default
    tolerance; `abs_tol` is accepted and unused.
    """
    polys = list(reg1.list_poly)
    adj = adjacency_matrix(polys)
    final = []
    left = list(range(len(polys)))
    while left:
        members = [left[0]]
        for j in left[1:]:
            if adj[members, j].any():
                members.append(j)
        part = Region([polys[k] for k in members], [])
        part.props = reg1.props.copy()
        final.append(part)
        taken = set(members)
        left = [k for k in left if k not in taken]
    return final


def is_interior(r0, r1, abs_tol=ABS_TOL):
    """polytope.py:1888-1909, kept verbatim in meaning: True as soon as some polytope of r1,
    enlarged by abs_tol, is NOT a subset of r0 (the reference's docstring promises the opposite;
    callers get what the reference returns).  One `<=` (region_diff search + volume) per member."""
    if isinstance(r0, Polytope):
        r0 = Region([r0])
    if isinstance(r1, Polytope):
        r1 = Region([r1])
    for p in r1:
        dummy = Polytope(p.A.copy(), p.b.copy() + abs_tol)
        if not dummy <= r0:
            return True
    return False


def is_inside(polyreg, point, abs_tol=ABS_TOL):
    """`point in polyreg` (deprecated in the reference too, polytope.py:1017-1029)."""
    warnings.warn('Write `point in polyreg` instead of calling this function.', DeprecationWarning)
    if not isinstance(point, np.ndarray):
        point = np.array(point)
    return polyreg.contains(point[:, np.newaxis], abs_tol)[0]


# ---------------------------------------------------------------------------
# volume, grid_region, enumerate_integral_points
# ---------------------------------------------------------------------------
def _volume_nsamples(n, nsamples):
    """Sample-count rule of volume() (polytope.py:1563-1582)."""
    N = 50 if n == 1 else 500 if n == 2 else 3000 if n == 3 else 10000
    if nsamples is not None and nsamples < 1:
        raise ValueError('`nsamples` must be >= 1, given:  {v}'.format(v=nsamples))
    if nsamples is not None:
        N = nsamples
    if N != int(N):
        raise ValueError(('it appears that a noninteger number of samples '
                          'has been given, namely:  {v}').format(v=nsamples))
    return int(N)


def volume_batch(polys, nsamples=None, seeds=None):
    """[volume(p, nsamples, seed) for p, seed in zip(polys, seeds)] for single
    Polytopes of one dimension: bounding boxes and the Monte-Carlo containment
    test (polytope.py:1583-1592) each run as one launch for the whole list.  The
    samples are numpy's `default_rng(seed).random((n, N))`, regenerated on the
    device from the PCG64 state, so a seeded volume equals the reference's."""
    if seeds is None:
        seeds = [None] * len(polys)
    out = [None] * len(polys)
    fd = is_fulldim_batch(polys)
    todo = [k for k, p in enumerate(polys) if fd[k]]
    for k, p in enumerate(polys):
        if not fd[k]:
            out[k] = 0.0
    if not todo:
        return out
    n = polys[todo[0]].A.shape[1]
    N = _volume_nsamples(n, nsamples)
    boxes = bounding_box_batch([polys[k] for k in todo])
    gens = []
    for k in todo:
        seed = seeds[k]
        gens.append(np.random.default_rng(seed))
    words = [engine.pcg64_state_words(g.bit_generator) for g in gens]
    A, b, rows = _stack([polys[k] for k in todo])
    lo = np.array([bb[0].flatten() for bb in boxes])
    hi = np.array([bb[1].flatten() for bb in boxes])
    counts = engine.volume_counts(A, b, lo, hi, N, words, rows)
    for t, k in enumerate(todo):
        gens[t].bit_generator.advance(n * N)     # a Generator passed as `seed` ends where numpy leaves it
        l_b, u_b = boxes[t]
        vol = np.prod(u_b - l_b) * int(counts[t]) / N
        polys[k]._set_volume(vol)
        out[k] = vol
    return out


def volume(polyreg, nsamples=None, seed=None):
    """Monte-Carlo volume of a Polytope or Region (polytope.py:1529-1594)."""
    if not is_fulldim(polyreg):
        return 0.0
    if isinstance(polyreg, Region):
        # the reference recurses without `nsamples` / `seed` (polytope.py:1558-1561)
        tot_vol = 0.0
        for v in volume_batch(polyreg.list_poly):
            tot_vol += v
        polyreg._set_volume(tot_vol)
        return tot_vol
    return volume_batch([polyreg], nsamples, [seed])[0]


def grid_region(polyreg, res=None):
    """Bounding-box grid points inside `polyreg` (polytope.py:2364-2399)."""
    bbox = polyreg.bounding_box
    if res is None:
        density = 8
        res = [math.ceil(density * (b[0] - a[0])) for a, b in zip(*bbox)]
    if len(res) != polyreg.dim:
        raise ValueError(("`len(res)` must equal the polytope's dimension "
                          "(which is {dim}), but instead `res` is:  {res}").format(dim=polyreg.dim, res=res))
    if any(n < 1 for n in res):
        raise ValueError(('`res` must contain `int` values >= 1, '
                          'instead `res` equals:  {res}').format(res=res))
    linspaces = [np.linspace(a, b, num=n) for a, b, n in zip(*bbox, res)]
    points = np.meshgrid(*linspaces)
    x = np.vstack(list(map(np.ravel, points)))
    x = x[:, polyreg.contains(x)]
    return (x, res)


def enumerate_integral_points(poly):
    """All points of `poly` with integer coordinates, d x m (polytope.py:2343-2361)."""
    a, b = poly.bounding_box
    a_int = np.floor(a)
    b_int = np.ceil(b)
    intervals = list(zip(a_int.flatten(), b_int.flatten()))
    box = box2poly(intervals)
    res = [int(b - a + 1) for a, b in intervals]
    grid, _ = grid_region(box, res=res)
    inside = poly.contains(grid)
    return grid[:, inside]


# ---------------------------------------------------------------------------
# qhull, extreme
# ---------------------------------------------------------------------------
def qhull_batch(point_sets, abs_tol=ABS_TOL):
    """[qhull(v, abs_tol) for v in point_sets] (same dimension) with all hulls in
    one launch (polytope.py:1685-1695 over quickhull.py:141-359)."""
    sets = [np.asarray(v, dtype=float) for v in point_sets]
    out = [None] * len(sets)
    if not sets:
        return out
    d = sets[0].shape[1]
    if d < 2 or d > 16:
        raise NotImplementedError('qhull: the B200 hull kernel covers 2 <= d <= 16, got d = %d' % d)
    nmax = max(v.shape[0] for v in sets)
    pts = np.zeros((len(sets), max(nmax, 1), d))
    n_pts = np.array([v.shape[0] for v in sets], dtype=np.int32)
    for k, v in enumerate(sets):
        pts[k, :v.shape[0]] = v
    res = engine.hull_batch(pts, n_pts, abs_tol=abs_tol)
    for k, v in enumerate(sets):
        st = int(res.status[k])
        if st in (engine.HULL_FEW_POINTS, engine.HULL_FLAT):
            if st == engine.HULL_FLAT:
                print("Warning: convex hull is not fully dimensional, returning empty polytope")
            out[k] = Polytope()
            continue
        if st != engine.HULL_OK:
            raise RuntimeError('qhull: hull %d failed with status %d' % (k, st))
        A, b, _ = res.facets(k)
        vert = np.unique(v[res.is_vertex[k][:v.shape[0]]], axis=0)
        out[k] = Polytope(A, b, minrep=True, vertices=vert)
    return out


def qhull(vertices, abs_tol=ABS_TOL):
    """Convex hull of N x d `vertices` as a Polytope (polytope.py:1685-1695)."""
    return qhull_batch([vertices], abs_tol)[0]


def _extreme_low_dim(poly1):
    """nx == 1 and nx == 2 branches of extreme() (polytope.py:1622-1653), host numpy."""
    A = poly1.A.copy()
    b = poly1.b.copy()
    nc, nx = A.shape
    V = np.array([])
    if nx == 1:
        for ii in range(nc):
            V = np.append(V, b[ii] / A[ii])
        if len(A) == 1:
            raise Exception("extreme: polytope is unbounded")
        return V
    alf = np.angle(A[:, 0] + 1j * A[:, 1])
    I = np.argsort(alf)
    H = np.vstack([A, A[0, :]])
    K = np.hstack([b, b[0]])
    I = np.hstack([I, I[0]])
    for ii in range(nc):
        HH = np.vstack([H[I[ii], :], H[I[ii + 1], :]])
        KK = np.hstack([K[I[ii]], K[I[ii + 1]]])
        if np.linalg.cond(HH) == np.inf:
            raise Exception("extreme: polytope is unbounded")
        try:
            v = np.linalg.solve(HH, KK)
        except Exception:
            raise Exception('Finding extreme points failed, Check if any unbounded Polytope is causing this.')
        V = np.append(V, v) if len(V) == 0 else np.vstack([V, v])
    return V


def extreme_batch(polys):
    """[extreme(p) for p in polys] (single Polytopes of one dimension).

    reduce, Chebyshev centres, polar duals, the dual hulls and the map back to
    vertices (polytope.py:1597-1682) each run as one launch for the whole list.
    """
    out = [None] * len(polys)
    todo = []
    for k, p in enumerate(polys):
        if isinstance(p, Region):
            raise Exception("extreme: not executable for regions")
        if p.vertices is not None:
            out[k] = p.vertices
        else:
            todo.append(k)
    if not todo:
        return out
    reduced = reduce_batch([polys[k] for k in todo])
    fd = is_fulldim_batch(reduced)
    work = []
    for t, k in enumerate(todo):
        if not fd[t]:
            out[k] = None
            continue
        nx = reduced[t].A.shape[1]
        if nx <= 2:
            V = _extreme_low_dim(reduced[t])
            polys[k].vertices = V.reshape((int(V.size / nx), nx))
            out[k] = polys[k].vertices
        else:
            work.append((t, k))
    if not work:
        return out
    red = [reduced[t] for t, _ in work]
    balls = cheby_ball_batch(red)
    A, b, rows = _stack(red)
    xc = np.array([bb[1] for bb in balls])
    d = A.shape[2]
    if d > 16:
        raise NotImplementedError('extreme: the B200 hull kernel covers d <= 16, got d = %d' % d)
    dual = engine.dual_points(A, b, xc, rows)
    hull = engine.hull_batch(dual, rows)
    V = engine.dual_facets_to_vertices(hull, xc)
    for w, (t, k) in enumerate(work):
        st = int(hull.status[w])
        if st in (engine.HULL_FEW_POINTS, engine.HULL_FLAT):
            out[k] = None            # qhull returned Polytope(): not full-dimensional (polytope.py:1666-1667)
            continue
        if st != engine.HULL_OK:
            raise RuntimeError('extreme: dual hull %d failed with status %d' % (k, st))
        o, c = int(hull.facet_off[w]), int(hull.facet_cnt[w])
        # is_fulldim(Q): the dual hull contains the ball of radius min(b) around the origin
        if not float(hull.b[o:o + c].min()) > ABS_TOL:
            Q = Polytope(hull.A[o:o + c], hull.b[o:o + c], minrep=True)
            if not is_fulldim(Q):
                out[k] = None
                continue
        polys[k].vertices = V[o:o + c].copy()
        out[k] = polys[k].vertices
    return out


def extreme(poly1):
    """Vertices of a bounded polytope, N x d (polytope.py:1597-1682)."""
    if poly1.vertices is not None:
        return poly1.vertices
    if isinstance(poly1, Region):
        raise Exception("extreme: not executable for regions")
    return extreme_batch([poly1])[0]


# ---------------------------------------------------------------------------
# envelope, is_convex, union, region_diff, mldivide
# ---------------------------------------------------------------------------
def envelope(reg, abs_tol=ABS_TOL):
    """Envelope of a region (polytope.py:1414-1464): the nP (nP-1) m Chebyshev LPs
    "does cell j cross facet ii of cell i" are independent and run as one batch."""
    cells = reg.list_poly
    nP = len(cells)
    jobs = []           # (i, ii, j)
    for i in range(nP):
        for ii in range(cells[i].A.shape[0]):
            for j in range(nP):
                if i != j:
                    jobs.append((i, ii, j))
    outer = [np.ones(c.A.shape[0]) for c in cells]
    if jobs:
        tests = [Polytope(np.vstack([cells[j].A, -cells[i].A[ii, :]]), np.hstack([cells[j].b, -cells[i].b[ii]]))
                 for i, ii, j in jobs]
        for (i, ii, j), (rc, _) in zip(jobs, cheby_ball_batch(tests)):
            if rc > abs_tol:
                outer[i][ii] = 0
    Ae = np.vstack([c.A[np.nonzero(o)[0], :] for c, o in zip(cells, outer)])
    be = np.hstack([c.b[np.nonzero(o)[0]] for c, o in zip(cells, outer)])
    ret = reduce(Polytope(Ae, be), abs_tol=abs_tol)
    if is_fulldim(ret):
        return ret
    return Polytope()


def is_convex(reg, abs_tol=ABS_TOL):
    """(convex?, envelope or None) for a region (polytope.py:988-1014)."""
    if len(reg) == 0:
        return True, None
    outer = envelope(reg)
    if is_empty(outer):
        return False, None
    Pl, Pu = reg.bounding_box
    Ol, Ou = outer.bounding_box
    bboxP = np.hstack([Pl, Pu])
    bboxO = np.hstack([Ol, Ou])
    if (np.any(abs(bboxP[:, 0] - bboxO[:, 0]) > abs_tol) or
            np.any(abs(bboxP[:, 1] - bboxO[:, 1]) > abs_tol)):
        return False, None
    if is_fulldim(outer.diff(reg)):
        return False, None
    return True, outer


def is_subset(small, big, abs_tol=ABS_TOL):
    """small is a subset of big, by the volume of the difference (polytope.py:1034-1052)."""
    for x in [small, big]:
        if not isinstance(x, (Polytope, Region)):
            raise TypeError('Not a Polytope or Region, got instead:\n\t' + str(type(x)))
    diff = small.diff(big)
    return bool(diff.volume < abs_tol)


def union(polyreg1, polyreg2, check_convex=False):
    """Union of polytopes / regions as a Region of non-overlapping polytopes
    (polytope.py:1166-1238)."""
    if is_empty(polyreg1):
        return polyreg2
    if is_empty(polyreg2):
        return polyreg1
    if check_convex:
        s1 = intersect(polyreg1, polyreg2)
        if is_fulldim(s1):
            s2 = polyreg2.diff(polyreg1)
            s3 = polyreg1.diff(polyreg2)
        else:
            s2 = polyreg1
            s3 = polyreg2
    else:
        s1 = polyreg1
        s2 = polyreg2
        s3 = None
    lst = []
    for part in (s1, s2, s3):
        if part is None:
            continue
        if len(part) == 0:
            if not is_empty(part):
                lst.append(part)
        else:
            for poly in part.list_poly:
                if not is_empty(poly):
                    lst.append(poly)
    if check_convex:
        final = []
        N = len(lst)
        if N > 1:
            while N > 0:
                templist = [lst[0]]
                for ii in range(1, N):
                    templist.append(lst[ii])
                    is_conv, env = is_convex(Region(templist))
                    if not is_conv:
                        templist.remove(lst[ii])
                for poly in templist:
                    lst.remove(poly)
                cvxpoly = reduce(envelope(Region(templist)))
                if not is_empty(cvxpoly):
                    final.append(reduce(cvxpoly))
                N = len(lst)
        else:
            final = lst
        return Region(final)
    return Region(lst)


def region_diff_batch(polys, regs, abs_tol=ABS_TOL, intersect_tol=ABS_TOL):
    """[region_diff(p, r) for p, r in zip(polys, regs)] with every depth-first
    search running on its own warp (polytope.py:2117-2282).  `regs` may be one
    Region / Polytope shared by all minuends."""
    if isinstance(regs, (Polytope, Region)):
        regs = [regs] * len(polys)
    out = [None] * len(polys)
    todo, cells_of = [], []
    for k, (poly, reg) in enumerate(zip(polys, regs)):
        if not isinstance(poly, Polytope):
            raise Exception('poly not a Polytope, but: ' + str(type(poly)))
        if isinstance(reg, Polytope):
            reg = Region([reg])
        if not isinstance(reg, Region):
            raise Exception('reg not a Region, but: ' + str(type(reg)))
        if is_empty(reg):
            out[k] = poly.copy()
        elif is_empty(poly):
            out[k] = Polytope()
        else:
            todo.append(k)
            cells_of.append(reg.list_poly)
    if not todo:
        return out
    # Regions beyond the kernel's per-problem envelope (64 cells, 2048 cell rows): only the
    # cells that intersect the minuend take part in the search (polytope.py:2146-2157), so
    # find those first with one batch of Chebyshev LPs and hand the kernel the survivors
    # (original order kept: the kernel's stable sort by radius then matches the reference's).
    if any(len(c) > 64 or sum(p.A.shape[0] for p in c) > 2048 for c in cells_of):
        tests, where = [], []
        for t, k in enumerate(todo):
            for i, cell in enumerate(cells_of[t]):
                tests.append(Polytope(np.vstack([polys[k].A, cell.A]), np.hstack([polys[k].b, cell.b])))
                where.append((t, i))
        hit = [[] for _ in todo]
        for (t, i), (rc, _) in zip(where, cheby_ball_batch(tests)):
            if rc >= intersect_tol:
                hit[t].append(cells_of[t][i])
        keep_t = []
        for t, k in enumerate(todo):
            if hit[t]:
                keep_t.append(t)
            else:
                out[k] = polys[k].copy()       # no cell intersects poly: the reference returns poly
        todo = [todo[t] for t in keep_t]
        cells_of = [hit[t] for t in keep_t]
        if not todo:
            return out
    PA, Pb, prow = _stack([polys[k] for k in todo])
    shared = all(c is cells_of[0] for c in cells_of)
    groups = [cells_of[0]] if shared else cells_of
    Nr = max(len(c) for c in groups)
    d = PA.shape[2]
    mr = max(max(p.A.shape[0] for p in c) for c in groups)
    RA = np.zeros((len(groups), Nr, mr, d))
    Rb = np.zeros((len(groups), Nr, mr))
    rrow = np.zeros((len(groups), Nr), dtype=np.int32)
    nreg = np.array([len(c) for c in groups], dtype=np.int32)
    for g, c in enumerate(groups):
        for i, p in enumerate(c):
            rrow[g, i] = p.A.shape[0]
            RA[g, i, :rrow[g, i]] = p.A
            Rb[g, i, :rrow[g, i]] = p.b
    if shared:
        RA, Rb = RA[0], Rb[0]
    res = engine.region_diff_batch(PA, Pb, RA, Rb, prow, rrow, nreg, abs_tol, intersect_tol)
    # pieces the reference reduces: one device pass for all of them
    red_idx = np.nonzero(res.reduce)[0]
    reduced = {}
    if len(red_idx):
        mx = int(res.rows[red_idx].max())          # the pool is padded to piece_m rows
        if mx > REDUCE_MAX_ROWS:
            # pieces with more rows than the fused pipeline's row masks hold: reduce() object by object
            pieces = [Polytope(res.A[f][:res.rows[f]], res.b[f][:res.rows[f]]) for f in red_idx]
            for f, q in zip(red_idx, reduce_batch(pieces)):
                reduced[f] = q
            red_idx = []
        else:
            rr = engine.reduce_batch(np.ascontiguousarray(res.A[red_idx][:, :mx]), np.ascontiguousarray(res.b[red_idx][:, :mx]),
                                     res.rows[red_idx], normalize=True)
            keeps = rr.keep_lists()
        for t, f in enumerate(red_idx):
            if rr.flags[t] & engine.F_EMPTY:
                reduced[f] = Polytope()
            else:
                q = Polytope(rr.A[t][keeps[t]], rr.b[t][keeps[t]])
                q.minrep = bool(rr.flags[t] & engine.F_MINREP)
                reduced[f] = q
    for t, k in enumerate(todo):
        st = int(res.status[t])
        if st == engine.DIFF_UNTOUCHED:
            out[k] = polys[k].copy()
        elif st == engine.DIFF_COVERED:
            out[k] = Polytope()
        elif st == engine.DIFF_PIECES:
            acc = Polytope()
            o = int(res.piece_off[t])
            for f in range(o, o + int(res.n_pieces[t])):
                if res.reduce[f]:
                    piece = reduced[f]
                else:
                    n = int(res.rows[f])
                    piece = Polytope(res.A[f][:n], res.b[f][:n])
                acc = union(acc, piece, False)
            out[k] = acc
        elif st == engine.DIFF_INDEX_ERROR:
            raise IndexError('region_diff: problem %d is outside the kernel envelope '
                             '(rows per LP <= 128, <= 64 cells) or indexes past the stacked rows' % k)
        else:
            raise RuntimeError('region_diff: problem %d ended with status %d' % (k, st))
    return out


def region_diff(poly, reg, abs_tol=ABS_TOL, intersect_tol=ABS_TOL, save=False):
    """poly minus reg (polytope.py:2117-2282)."""
    if not isinstance(poly, Polytope):
        raise Exception('poly not a Polytope, but: ' + str(type(poly)))
    if isinstance(reg, Polytope):
        reg = Region([reg])
    if not isinstance(reg, Region):
        raise Exception('reg not a Region, but: ' + str(type(reg)))
    return region_diff_batch([poly], [reg], abs_tol, intersect_tol)[0]


def mldivide(a, b, save=False):
    """Set difference a minus b (polytope.py:1469-1505)."""
    if isinstance(b, Polytope):
        b = Region([b])
    if isinstance(a, Region):
        P = Region()
        for poly in a:
            Pdiff = poly
            for poly1 in b:
                Pdiff = mldivide(Pdiff, poly1, save=save)
            P = union(P, Pdiff, check_convex=True)
    elif isinstance(a, Polytope):
        P = region_diff(a, b)
    else:
        raise Exception('a neither Region nor Polytope')
    return P

"""CPU oracle for the batched-LP hot path of tulip-control/polytope.

TEST INFRASTRUCTURE ONLY.  Nothing under ``polytope_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the CPU arm being timed -- never as the product path.

What it restates (all citations relative to /root/reference):

* ``lpsolve``            -- polytope/solvers.py:76-106, scipy adapter :149-158
* ``normalize_rows``     -- Polytope.__init__, polytope/polytope.py:122-138
* ``cheby_ball``         -- polytope/polytope.py:1241-1300
* ``is_fulldim``         -- polytope/polytope.py:962-985
* ``bounding_box``       -- polytope/polytope.py:1314-1411
* ``reduce``             -- polytope/polytope.py:1053-1163
* ``intersect``          -- Polytope.intersect, polytope/polytope.py:255-275
* ``is_adjacent``        -- polytope/polytope.py:1827-1866 (overlap=True branch)
* ``adjacency_matrix``   -- prop2partition.py:46-63 (find_adjacent_regions)
* ``contains``           -- Polytope.contains, polytope/polytope.py:206-218
* ``region_contains``    -- Region.contains, polytope/polytope.py:736-748
* ``volume``             -- polytope/polytope.py:1529-1594 (single polytope)
* ``grid_region``        -- polytope/polytope.py:2364-2399
* ``pcg64_uniform``      -- numpy's PCG64 stream behind default_rng(seed).random()
                            (the kernel regenerates it; this restatement pins the
                            algorithm against numpy itself in tests/test_oracle.py)
* ``qhull`` / ``extreme`` -- polytope/polytope.py:1685-1695, :1597-1682.  The
                            reference's own quickhull.py (Barber et al. 1996) is
                            O(|NV|^2 d^2) Python and cannot finish d >= 9
                            (SURVEY.md section 0), so the hull itself is taken
                            from Qhull (scipy.spatial.ConvexHull, the published
                            implementation of the same algorithm); the hull of
                            a point set is unique, and tests/golden/hull_cases.npz
                            pins this oracle to the reference's outputs as sets
                            at the sizes the reference can run (d <= 6).

The arithmetic of the path lives in a third-party dependency that is NOT
vendored under /root/reference: ``scipy.optimize.linprog`` -> HiGHS
(reference pins scipy==1.10.0, requirements/default.txt:2; this image has
scipy 1.18.1 / HiGHS 1.12.0, inside the range pyproject.toml:28-32 allows).
The oracle therefore calls the same ``linprog`` entry with the same arguments
the reference does and restates only the reference's own numpy logic around it.

Parity pinning: ``tests/golden/make_golden.py`` imports the unmodified
reference from /root/reference in the build container, records its outputs on
seeded inputs (including the inputs of the reference's own tests) and
``tests/test_oracle.py`` asserts this module reproduces them bit for bit.

Data model: a polytope is a pair ``(A, b)`` of float64 arrays (m x d, m); the
reference's ``Polytope`` object is not rebuilt here.  Every function returns
plain arrays / index lists so results can be compared element by element.
"""
import numpy as np
from scipy import optimize

ABS_TOL = 1e-7  # polytope/polytope.py:83

# number of LPs solved since the last reset (the unit BASELINE.json's metric
# counts: one lpsolve-equivalent solve the reference algorithm requires)
lp_count = 0


def lpsolve(c, G, h):
    """min c'x s.t. Gx <= h, x free.  solvers.py:149-158."""
    global lp_count
    lp_count += 1
    sol = optimize.linprog(c, G, np.transpose(h), None, None,
                           bounds=(None, None))
    return dict(status=sol.status, x=sol.x, fun=sol.fun)


def normalize_rows(A, b):
    """Row normalisation of Polytope.__init__ (polytope.py:122-138).

    Returns (A_n, b_n, pos): the normalised float64 copies and the indices of
    the input rows that survive (rows with 2-norm <= 1e-10 are dropped).
    """
    A = np.asarray(A)
    b = np.asarray(b)
    An = A.astype(float)
    bn = b.astype(float).flatten()
    pos = np.arange(An.shape[0]) if An.ndim == 2 else np.arange(0)
    if A.size > 0:
        # the norm is taken on the caller's array (its dtype), :129
        nrm = np.sqrt(np.sum(A * A, 1)).flatten()
        pos = np.nonzero(nrm > 1e-10)[0]
        An = An[pos, :]
        bn = bn[pos]
        mult = 1 / nrm[pos]
        An = An * mult[:, None]      # same products as the row loop :136-137
        bn = bn * mult
    return An, bn, pos


def cheby_lp_data(A, b):
    """(c, G, h) of the Chebyshev LP, polytope.py:1283-1287."""
    c = np.negative(np.r_[np.zeros(A.shape[1]), 1])
    G = np.c_[A, np.sqrt(np.sum(A * A, axis=1))]
    return c, G, b


def cheby_ball(A, b):
    """Chebyshev radius/centre of {x: Ax<=b} (polytope.py:1241-1300).

    Returns (r, xc); (0, None) when A has no rows, the LP is not optimal or the
    radius is negative.  (A, b) is used as given (no normalisation here: the
    reference object was normalised by its constructor).
    """
    if len(A) == 0:                     # is_empty, :939-948
        return 0, None
    c, G, h = cheby_lp_data(A, b)
    sol = lpsolve(c, G, h)
    if sol['status'] == 0:
        r = sol['x'][-1]
        if r < 0:
            return 0, None
        return np.double(r), np.array(sol['x'][0:-1])
    return 0, None


def is_fulldim(A, b, abs_tol=ABS_TOL):
    """polytope.py:962-985 for a single polytope."""
    r, _ = cheby_ball(A, b)
    return bool(r > abs_tol)


def bounding_box(A, b):
    """(l, u) column vectors, polytope.py:1362-1411."""
    m, n = np.shape(A)
    l = np.zeros([n, 1])
    u = np.zeros([n, 1])
    for sign, out in ((1.0, l), (-1.0, u)):
        for i in range(n):
            c = np.zeros(n)
            c[i] = sign
            sol = lpsolve(c, A, b)
            st = sol['status']
            if st == 0:
                out[i] = sol['x'][i]
            elif st == 3:
                out[i] = -sign * np.inf
            elif st == 2:
                out[i] = 0 if sign > 0 else l[i]
            else:
                raise RuntimeError('bounding_box: lpsolve returned %r' % sol)
    return l, u


def duplicate_rows(A, b, abs_tol=ABS_TOL):
    """Rows `reduce` removes as parallel duplicates, polytope.py:1094-1110.

    Returns the sorted unique kept indices (what np.setdiff1d gives).
    """
    neq = A.shape[0]
    a_norm = 1 / np.sqrt(np.sum(A.T**2, 0))
    a_normed = np.dot(A.T, np.diag(a_norm)).T
    removed = []
    for i in range(neq):
        for j in range(i + 1, neq):
            if np.dot(a_normed[i].T, a_normed[j]) > 1 - abs_tol:
                if b[i] * a_norm[i] < b[j] * a_norm[j]:
                    removed.append(j)
                else:
                    removed.append(i)
    return np.setdiff1d(range(neq), removed).tolist()


def bbox_candidates(A, b, lb, ub):
    """Boolean mask of rows that may touch the bounding box, :1118-1132."""
    cand = ~(np.dot((A > 0) * A, ub - lb)
             - (np.array([b]).T - np.dot(A, lb)) < -1e-4)
    return cand.squeeze()


def reduce(A, b, abs_tol=ABS_TOL, normalize=True, non_empty_bounded=True):
    """Redundant-row removal, polytope.py:1053-1163, on raw (A, b).

    Returns a dict:
      empty   -- True when the reference returns the empty Polytope() (:1082)
      keep    -- indices (into the INPUT rows) of the rows the reference keeps
      A, b    -- the arrays handed to the final Polytope(...) constructor
                 (b carries the reference's +0.1/-0.1 one-ulp drift, :1149-1151)
      minrep  -- value of the returned object's `minrep`
      r, xc   -- Chebyshev ball found by the leading is_fulldim() call
      n_lp    -- LPs this call solved
      margins -- decision quantity `-fun - h[k]` of every row LP that ended with status 0
                 (compared with abs_tol at :1153; lets tests count "ambiguous" LPs)
    """
    global lp_count
    lp0 = lp_count
    if normalize:
        A, b, idx = normalize_rows(A, b)
    else:
        A = np.array(A, dtype=float)
        b = np.array(b, dtype=float).flatten()
        idx = np.arange(A.shape[0])
    r, xc = cheby_ball(A, b)
    out = dict(empty=False, keep=[], A=None, b=None, minrep=False,
               r=r, xc=xc, n_lp=0, margins=[])
    if not r > ABS_TOL:                 # is_fulldim(poly) uses its default, :1081
        out['empty'] = True
        out['n_lp'] = lp_count - lp0
        return out
    fin = np.nonzero(b != np.inf)[0]                      # :1087-1089
    A, b, idx = A[fin], b[fin], idx[fin]
    keep = duplicate_rows(A, b, abs_tol)                  # :1094-1112
    A, b, idx = A[keep], b[keep], idx[keep]
    neq, nx = A.shape
    done = non_empty_bounded and neq <= nx + 1            # :1113-1116
    if not done and neq > 3 * nx:                         # :1118-1134
        An, bn, _ = normalize_rows(A, b)                  # Polytope(A_arr,b_arr)
        lb, ub = bounding_box(An, bn)
        cand = bbox_candidates(A, b, lb, ub)
        A, b, idx = A[cand], b[cand], idx[cand]
        neq, nx = A.shape
    done = done or (non_empty_bounded and neq <= nx + 1)  # :1135-1138
    if done:
        out.update(keep=idx.tolist(), A=A, b=b, n_lp=lp_count - lp0)
        return out
    kept = []
    h = b                                                 # aliased, as in :1147
    for k in range(neq):                                  # :1142-1160
        h[k] += 0.1
        sol = lpsolve(-A[k, :], A, h)
        h[k] -= 0.1
        if sol['status'] == 0:
            out['margins'].append(float(-sol['fun'] - h[k]))
            if (-sol['fun'] - h[k]) > abs_tol:
                kept.append(k)
        elif sol['status'] == 3:
            kept.append(k)
    out.update(keep=idx[kept].tolist(), A=A[kept], b=b[kept], minrep=True,
               n_lp=lp_count - lp0)
    return out


def intersect(A1, b1, A2, b2, abs_tol=ABS_TOL):
    """Polytope.intersect, polytope.py:255-275, on normalised operands.

    Returns the `reduce` dict of the stacked system (empty=True when either
    operand is not full-dimensional, :268-269); `keep` indexes the stacked rows.
    """
    if not is_fulldim(A1, b1) or not is_fulldim(A2, b2):
        return dict(empty=True, keep=[], A=None, b=None, minrep=False,
                    r=0, xc=None, n_lp=2)
    if A1.shape[1] != A2.shape[1]:
        raise Exception('polytopes have different dimension')
    return reduce(np.vstack([A1, A2]), np.hstack([b1, b2]), abs_tol=abs_tol)


def adjacent_lp_data(A1, b1, A2, b2, abs_tol=ABS_TOL):
    """Stacked, inflated, re-normalised system of is_adjacent, :1856-1865."""
    A = np.concatenate((A1, A2))
    b = np.concatenate((b1 + abs_tol, b2 + abs_tol))
    An, bn, _ = normalize_rows(A, b)
    return An, bn


def is_adjacent(A1, b1, A2, b2, abs_tol=ABS_TOL):
    """polytope.py:1827-1866, overlap=True, two single polytopes."""
    An, bn = adjacent_lp_data(A1, b1, A2, b2, abs_tol)
    return is_fulldim(An, bn, abs_tol=abs_tol / 10)


def adjacency_matrix(cells, abs_tol=ABS_TOL):
    """find_adjacent_regions, prop2partition.py:46-63, over single polytopes.

    `cells` is a list of normalised (A, b); returns a dense int8 matrix.
    """
    n = len(cells)
    adj = np.zeros((n, n), dtype=np.int8)
    for i in range(n):
        adj[i, i] = 1
        for j in range(i):
            adj[i, j] = adj[j, i] = is_adjacent(*cells[i], *cells[j],
                                                abs_tol=abs_tol)
    return adj


# ---------------------------------------------------------------------------
# contains / volume / grid_region / qhull / extreme
# ---------------------------------------------------------------------------
def contains(A, b, points, abs_tol=ABS_TOL):
    """Polytope.contains, polytope.py:206-218: points are column vectors (d x N)."""
    test = A.dot(points) - b[:, np.newaxis] < abs_tol
    return np.all(test, axis=0)


def region_contains(cells, points, abs_tol=ABS_TOL):
    """Region.contains, polytope.py:736-748, over a list of (A, b)."""
    contained = np.full(points.shape[1], False, dtype=bool)
    for A, b in cells:
        contained = np.logical_or(contains(A, b, points, abs_tol), contained)
    return contained


def volume(A, b, nsamples=None, seed=None):
    """volume() of one normalised polytope, polytope.py:1529-1594.

    Returns (vol, count, N): the estimate, the number of samples strictly
    inside, and the number of samples drawn.
    """
    if not is_fulldim(A, b):
        return 0.0, 0, 0
    n = A.shape[1]
    N = 50 if n == 1 else 500 if n == 2 else 3000 if n == 3 else 10000
    if nsamples is not None:
        N = nsamples
    l_b, u_b = bounding_box(A, b)
    x = (np.tile(l_b, (1, N))
         + np.random.default_rng(seed).random((n, N))
         * np.tile(u_b - l_b, (1, N)))
    aux = (np.dot(A, x) - np.tile(np.array([b]).T, (1, N)))
    aux = np.nonzero(np.all(aux < 0, 0))[0].shape[0]
    vol = np.prod(u_b - l_b) * aux / N
    return vol, aux, N


def grid_points(l_b, u_b, res):
    """The unfiltered grid of grid_region, polytope.py:2391-2396."""
    linspaces = [np.linspace(a, bb, num=n) for a, bb, n in zip(l_b, u_b, res)]
    points = np.meshgrid(*linspaces)
    return np.vstack(list(map(np.ravel, points)))


PCG_MULT = (2549297995355413924 << 64) | 4865540595714422341   # numpy/random/src/pcg64/pcg64.h
_M128 = (1 << 128) - 1


def pcg64_uniform(state, inc, count, skip=0):
    """`count` doubles of Generator.random() from a PCG64 (state, inc), after
    skipping `skip` draws: step the 128-bit LCG, output XSL-RR 128/64, take the
    top 53 bits (numpy: pcg64_random_r + random_standard_uniform)."""
    # jump ahead (pcg_advance_lcg_128)
    acc_mult, acc_plus, cur_mult, cur_plus, delta = 1, 0, PCG_MULT, inc, skip
    while delta > 0:
        if delta & 1:
            acc_mult = (acc_mult * cur_mult) & _M128
            acc_plus = (acc_plus * cur_mult + cur_plus) & _M128
        cur_plus = ((cur_mult + 1) * cur_plus) & _M128
        cur_mult = (cur_mult * cur_mult) & _M128
        delta >>= 1
    state = (acc_mult * state + acc_plus) & _M128
    out = np.empty(count)
    for i in range(count):
        state = (state * PCG_MULT + inc) & _M128
        hi, lo = state >> 64, state & ((1 << 64) - 1)
        x = hi ^ lo
        rot = hi >> 58
        u = ((x >> rot) | (x << ((64 - rot) & 63))) & ((1 << 64) - 1)
        out[i] = (u >> 11) * (1.0 / 9007199254740992.0)
    return out


def qhull(points, abs_tol=ABS_TOL):
    """Facets (A, b) of the convex hull of N x d points, rows normalised as the
    Polytope constructor leaves them (polytope.py:1685-1695).  Empty arrays when
    the reference returns Polytope() (too few points / not full-dimensional,
    quickhull.py:155-165).  Hull from Qhull, see the module docstring."""
    from scipy.spatial import ConvexHull, QhullError
    points = np.asarray(points, dtype=float)
    n, d = points.shape
    if n <= d:
        return np.zeros((0, d)), np.zeros(0), np.zeros((0, d))
    s = np.linalg.svd(np.transpose(points - points[0, :]), compute_uv=False)
    if np.sum(s > 1e-15) < d:
        return np.zeros((0, d)), np.zeros(0), np.zeros((0, d))
    try:
        hull = ConvexHull(points)
    except QhullError:
        return np.zeros((0, d)), np.zeros(0), np.zeros((0, d))
    eq = hull.equations                       # [normal | offset], normal . x + offset <= 0
    A, b, _ = normalize_rows(eq[:, :-1], -eq[:, -1])
    return A, b, points[np.sort(hull.vertices)]


def extreme(A, b):
    """Vertices of a bounded polytope given as raw (A, b): reduce, polar dual
    around the Chebyshev centre, hull of the dual, map the dual facets back
    (polytope.py:1597-1682, the nx >= 3 branch).  Returns an N x d array with
    one row per (simplicial) facet of the dual hull, or None."""
    red = reduce(A, b)
    if red['empty']:
        return None
    Ar, br, _ = normalize_rows(red['A'], red['b'])        # Polytope(A_arr[keep], b_arr[keep]), :1161
    rmid, xmid = cheby_ball(Ar, br)
    if not rmid > ABS_TOL:
        return None
    Ai = np.zeros(Ar.shape)
    for ii in range(Ar.shape[0]):
        Ai[ii, :] = Ar[ii, :] / (br[ii] - np.dot(Ar[ii, :], xmid))
    H, K, _ = qhull(Ai)
    if len(H) == 0:
        return None
    return H / K[:, None] + xmid


# ---------------------------------------------------------------------------
# envelope, region_diff
# ---------------------------------------------------------------------------
def cheby_radius_of_rows(A, b):
    """cheby_ball(Polytope(A, b))[0]: constructor normalisation, then the LP."""
    An, bn, _ = normalize_rows(A, b)
    return cheby_ball(An, bn)[0]


def envelope_outer_rows(cells, abs_tol=ABS_TOL):
    """The 'outer' rows of envelope(), polytope.py:1430-1458: for every cell i the
    indices of its rows that no other cell crosses.  `cells` = list of (A, b)."""
    out = []
    for i, (A1, b1) in enumerate(cells):
        outer = np.ones(A1.shape[0])
        for ii in range(A1.shape[0]):
            for j, (A2, b2) in enumerate(cells):
                if i == j:
                    continue
                rc = cheby_radius_of_rows(np.vstack([A2, -A1[ii, :]]), np.hstack([b2, -b1[ii]]))
                if rc > abs_tol:
                    outer[ii] = 0
        out.append(np.nonzero(outer)[0])
    return out


def envelope(cells, abs_tol=ABS_TOL):
    """envelope() of a region given as normalised cells, polytope.py:1414-1464.
    Returns the `reduce` dict of the stacked outer rows (empty=True when the
    reference returns Polytope())."""
    outer = envelope_outer_rows(cells, abs_tol)
    Ae = np.vstack([A[idx, :] for (A, b), idx in zip(cells, outer)])
    be = np.hstack([b[idx] for (A, b), idx in zip(cells, outer)])
    if len(be) == 0:
        return dict(empty=True, keep=[], A=None, b=None)
    red = reduce(Ae, be, abs_tol=abs_tol)
    if not red['empty']:
        Ar, br, _ = normalize_rows(red['A'], red['b'])
        if not is_fulldim(Ar, br):
            red['empty'] = True
    return red


class _PyList(list):
    """INDICES of region_diff: numpy fancy indexing wraps negative entries."""


def region_diff(poly, cells, abs_tol=ABS_TOL, intersect_tol=ABS_TOL, max_steps=100000):
    """region_diff(poly, reg), polytope.py:2117-2282, on normalised (A, b) pairs.

    Returns (kind, pieces):
      kind 'poly'    -- the reference returns `poly` itself (no cell intersects it)
      kind 'empty'   -- the reference returns Polytope() (a cell covers poly)
      kind 'pieces'  -- pieces = [(A_rows, b_rows, reduced)] in the order the
                        reference unions them; `reduced` pieces went through
                        reduce() (:2276), the others are used as they are (:2217).
                        Rows are the RAW stacked rows handed to Polytope(...).
    """
    PA, Pb = poly
    N = len(cells)
    Rc = np.zeros(N)
    for i, (A1, b1) in enumerate(cells):
        Rc[i] = cheby_radius_of_rows(np.vstack([PA, A1]), np.hstack([Pb, b1]))
    N = int(np.sum(Rc >= intersect_tol))
    if N == 0:
        return 'poly', []
    ind = np.argsort(-Rc)
    A = PA.copy()
    B = Pb.copy()
    m = A.shape[0]
    mi = np.zeros(N, dtype=int)
    HK = np.hstack([A, np.array([B]).T])
    for ii in range(N):
        i = ind[ii]
        if not is_fulldim(*cells[i]):
            continue
        Hni, Kni = cells[i]
        for j in range(Hni.shape[0]):
            HKnij = np.hstack([Hni[j, :], Kni[j]])
            if np.all(np.sum(np.abs(HK - np.tile(HKnij, [m, 1])), axis=1) >= abs_tol):
                mi[ii] += 1
                A = np.vstack([A, Hni[j, :]])
                B = np.hstack([B, Kni[j]])
    if np.any(mi == 0):
        return 'empty', []
    M = int(np.sum(mi))
    beg_mi = np.cumsum(np.hstack([0, mi[:-1]])) + m if N > 1 else np.array([m])
    A = np.vstack([A, -A[m:m + M, :]])
    B = np.hstack([B, -B[m:m + M]])
    counter = np.zeros(N, dtype=int)
    INDICES = list(range(m))
    level = 0
    pieces = []

    def radius(idx):
        idx = np.array(idx, dtype=int)
        return cheby_radius_of_rows(A[idx, :], B[idx])

    for _ in range(max_steps):
        if level == -1:
            break
        if counter[level] == 0:
            for j in range(level, N):
                R = radius(INDICES + list(range(beg_mi[j], beg_mi[j] + mi[j])))
                if R > abs_tol:
                    level = j
                    counter[level] = 1
                    INDICES = INDICES + [beg_mi[level] + M]
                    break
            if R < abs_tol:
                level = level - 1
                idx = np.array(INDICES, dtype=int)
                pieces.append((A[idx, :], B[idx], False))
                nz = len(np.nonzero(counter)[0])
                for _jj in range(nz - 1, -1, -1):
                    if counter[level] <= mi[level]:
                        INDICES[-1] = INDICES[-1] - M
                        INDICES = INDICES + [beg_mi[level] + counter[level] + M]
                        break
                    else:
                        counter[level] = 0
                        INDICES = INDICES[0:m + int(np.sum(counter))]
                        if level == -1:
                            return 'pieces', pieces
        else:
            nzcount = np.nonzero(counter)[0]
            for jj in range(len(nzcount) - 1, -1, -1):
                level = nzcount[jj]
                counter[level] += 1
                if counter[level] <= mi[level]:
                    INDICES[-1] = INDICES[-1] - M
                    INDICES = INDICES + [beg_mi[level] + counter[level] + M - 1]
                    break
                else:
                    counter[level] = 0
                    INDICES = INDICES[0:m + int(np.sum(counter))]
                    level = level - 1
                    if level == -1:
                        return 'pieces', pieces
        idx = np.array(INDICES, dtype=int)
        rc = cheby_radius_of_rows(A[idx, :], B[idx])
        if rc > abs_tol:
            if level == N - 1:
                pieces.append((A[idx, :], B[idx], True))
            else:
                level = level + 1
    return 'pieces', pieces

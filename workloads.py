"""Synthetic workloads named by BASELINE.json / SURVEY.md section 8(d).

Shared by bench.py, the tests and tests/golden/make_golden.py so that the GPU
arm, the CPU arm and the golden fixtures all see identical (A, b) inputs.
Pure numpy; no reference or oracle import.
"""
import numpy as np


def box_cuts(seed, m, d, shift_scale=False):
    """G1 "box+cuts": the box [-1,1]^d plus m-2d random unit-normal cuts
    a.x <= t*||a||_1, t ~ U(0.6, 1.4), rows permuted.
    A cut with t >= 1 is redundant w.r.t. the box by construction.
    """
    rng = np.random.default_rng(seed)
    k = m - 2 * d
    if k < 0:
        raise ValueError('box_cuts needs m >= 2d')
    C = rng.standard_normal((k, d))
    C /= np.sqrt(np.sum(C * C, axis=1))[:, None]
    t = rng.uniform(0.6, 1.4, k)
    A = np.vstack([np.eye(d), -np.eye(d), C])
    b = np.hstack([np.ones(2 * d), t * np.sum(np.abs(C), axis=1)])
    perm = rng.permutation(m)
    A, b = A[perm], b[perm]
    if shift_scale:
        x0 = rng.uniform(-1, 1, d)
        s = rng.uniform(0.3, 1.0)
        b = s * b + A @ x0
    return np.ascontiguousarray(A), np.ascontiguousarray(b)


def box_cuts_batch(cfg, n_poly, m, d, shift_scale=False, first=0):
    """Stacked batch: polytope i uses seed 1000*cfg + i."""
    A = np.empty((n_poly, m, d))
    b = np.empty((n_poly, m))
    for i in range(n_poly):
        A[i], b[i] = box_cuts(1000 * cfg + first + i, m, d, shift_scale)
    return A, b


def box_grid(shape, origin=0.0, cell=1.0):
    """Regular grid of axis-aligned boxes as rows [I; -I] x <= [hi; -lo]
    (cfg5: prop2partition adjacency).  Returns (A[n,2d,d], b[n,2d], index[n,d]).
    """
    shape = tuple(shape)
    d = len(shape)
    idx = np.stack(np.meshgrid(*[np.arange(s) for s in shape],
                               indexing='ij'), -1).reshape(-1, d)
    lo = origin + cell * idx
    hi = lo + cell
    n = idx.shape[0]
    A = np.broadcast_to(np.vstack([np.eye(d), -np.eye(d)]), (n, 2 * d, d)).copy()
    b = np.hstack([hi, -lo])
    return A, b, idx


def unit_cube3():
    """cfg1: Polytope(vstack(I3,-I3), [1,1,1,0,0,0]) (not from_box)."""
    return np.vstack([np.eye(3), -np.eye(3)]), np.array([1., 1, 1, 0, 0, 0])


def contains_points(seed, d, n):
    """Test points for contains(): uniform in [-1.6, 1.6]^d, a few exactly on the
    faces of the box [-1, 1]^d and one at the origin (column vectors, d x n)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1.6, 1.6, (d, n))
    x[0, :5] = 1.0
    x[:, 5] = 0.0
    return x


def hull_points(seed, n, d):
    """Gaussian point cloud for qhull() (n x d)."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, d))


# (m, d) / (N, d) of the golden fixtures in tests/golden/sets_cases.npz, hull_cases.npz
VOLUME_SPECS = [(6, 2), (10, 3), (16, 4), (16, 6), (32, 8)]
HULL_SPECS = [(20, 2), (30, 3), (40, 4), (30, 5), (24, 6)]
EXTREME_SPECS = [(6, 2), (10, 3), (12, 4), (15, 5), (18, 6)]


def diff_case(i):
    """Seeded (poly, [cells]) inputs for region_diff / envelope / union tests:
    raw (A, b) pairs.  d cycles 2, 3, 4; the cells are shifted copies of box+cuts
    polytopes so that some intersect the minuend, some cover it, some miss it."""
    d = 2 + i % 3
    m = 2 * d + 2 + (i % 2)
    rng = np.random.default_rng(5000 + i)
    A, b = box_cuts(5100 + i, m, d)
    ncell = 1 + i % 3
    cells = []
    for c in range(ncell):
        Ac, bc = box_cuts(5200 + 10 * i + c, m, d)
        kind = (i + c) % 5
        if kind == 4:
            shift, scale = np.zeros(d), 3.0                  # covers the minuend
        elif kind == 3:
            shift, scale = 4.0 * np.ones(d), 1.0             # misses it
        else:
            shift, scale = rng.uniform(-1.2, 1.2, d), rng.uniform(0.5, 1.3)
        cells.append((Ac, scale * bc + Ac @ shift))
    return (A, b), cells


def box_rows(intervals):
    """Rows [I; -I] x <= [hi; -lo] of a hyperrectangle (what box2poly builds)."""
    iv = np.asarray(intervals, dtype=float)
    n = iv.shape[0]
    return np.vstack([np.eye(n), -np.eye(n)]), np.hstack([iv[:, 1], -iv[:, 0]])


DIFF_CASES = 30


def partition_scenarios():
    """Inputs of `separate` / `find_adjacent_regions` (tests/golden/make_golden_partition.py):
    a list of (name, cells, groups) with cells = [(A, b), ...] and groups = ordered lists of
    cell indices, one per region.  `separate` is recorded for every group, the adjacency matrix
    for the partition formed by all groups of a scenario."""
    out = []

    def boxes(coords):
        return [box_rows([[c, c + 1.0] for c in xy]) for xy in coords]

    # chain listed out of order: the reference's single ordered pass splits a connected set
    out.append(('chain_out_of_order', boxes([(0, 0), (2, 0), (1, 0), (5, 5)]), [[0, 1, 2], [3]]))
    # corner contact counts as adjacent (both polytopes are inflated by abs_tol)
    out.append(('corner_contact', boxes([(0, 0), (1, 1), (3, 0), (2, 1), (0, 3)]), [[0, 1, 2, 3, 4]]))
    rng = np.random.default_rng(77)
    for k, shape in enumerate([(4, 4), (3, 3, 3), (5, 4), (2, 2, 2, 2)]):
        A, b, _ = box_grid(shape)
        n = len(A)
        pick = rng.permutation(n)[: max(6, (2 * n) // 3)]
        label = rng.integers(0, 4, len(pick))
        groups = [[int(i) for i in range(len(pick)) if label[i] == g] for g in range(4)]
        out.append(('grid%d' % k, [(A[i], b[i]) for i in pick], [g for g in groups if g]))
    for k, (m, d, n) in enumerate([(6, 2, 9), (8, 3, 8)]):      # overlapping / disjoint general polytopes
        cells = []
        for i in range(n):
            A, b = box_cuts(9100 + 10 * k + i, m, d)
            cells.append((A, 0.45 * b + A @ rng.uniform(-1.6, 1.6, d)))
        label = rng.integers(0, 3, n)
        groups = [[int(i) for i in range(n) if label[i] == g] for g in range(3)]
        out.append(('general%d' % k, cells, [g for g in groups if g]))
    return out

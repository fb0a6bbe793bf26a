"""The adjacency pair loop of /root/reference/polytope/prop2partition.py as one kernel launch.

Only `find_adjacent_regions` (prop2partition.py:46-63; `Partition.compute_adj`, :244-261, fills
the same matrix) is on the hot path (BASELINE cfg5); the Partition classes
themselves are bookkeeping and are not provided.
"""
import numpy as np

from polytope_b200.polytope import ABS_TOL, Region, adjacency_matrix


def find_adjacent_regions(partition, abs_tol=ABS_TOL):
    """Which regions of a partition touch or overlap (prop2partition.py:46-63).

    `partition` is anything with a `.regions` list (the reference's Partition) or a plain
    sequence of Regions / Polytopes.  The reference calls is_adjacent(region_i, region_j) for
    every j < i, i.e. one LP per pair of member polytopes until one is adjacent; here every
    member polytope of every region becomes a cell of one `adjacency_matrix` launch and the
    flags are OR-ed per region pair.  Returns a dense symmetric int8 ndarray with ones on the
    diagonal -- the reference returns the same entries as a scipy `lil_matrix`; wrap the result
    in `scipy.sparse.lil_matrix(...)` if a sparse container is needed (scipy is not a dependency
    of this package).
    """
    regions = list(getattr(partition, 'regions', partition))
    n = len(regions)
    adj = np.eye(n, dtype=np.int8)
    cells, owner = [], []
    for i, reg in enumerate(regions):
        for poly in (reg.list_poly if isinstance(reg, Region) else [reg]):
            cells.append(poly)
            owner.append(i)
    if any(c.dim != cells[0].dim for c in cells):
        raise Exception("is_adjacent: polytopes do not have the same dimension")
    if n < 2 or len(cells) < 2:
        return adj
    member = np.zeros((n, len(cells)), dtype=np.int32)
    member[owner, np.arange(len(cells))] = 1
    cell_adj = adjacency_matrix(cells, abs_tol=abs_tol).astype(np.int32)
    touching = (member @ cell_adj @ member.T) > 0
    adj[touching] = 1
    return adj

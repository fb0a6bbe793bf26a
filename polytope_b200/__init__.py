"""polytope_b200: B200-native batched-LP engine behind the polytope API.

Drop-in for the LP hot path of tulip-control/polytope:
`solvers.lpsolve`, `Polytope`/`Region`, `is_fulldim`, `cheby_ball`,
`bounding_box`, `reduce`, `intersect`, `is_adjacent`, `contains`, `volume`,
`grid_region`, `qhull`, `extreme`, `separate`, `find_adjacent_regions`, plus batched forms.
All LPs, hulls and point sweeps run in hand-written sm_100a kernels (libpolytope_b200.so, C ABI in
include/polytope_b200.h); there is no CPU fallback.
"""
from polytope_b200 import solvers
from polytope_b200.polytope import (
    ABS_TOL, Polytope, Region, box2poly, is_empty, is_fulldim, cheby_ball,
    bounding_box, reduce, intersect, is_adjacent,
    cheby_ball_batch, is_fulldim_batch, bounding_box_batch, reduce_batch,
    intersect_batch, adjacency_matrix,
    volume, volume_batch, grid_region, enumerate_integral_points,
    qhull, qhull_batch, extreme, extreme_batch,
    envelope, is_convex, is_subset, union, region_diff, region_diff_batch, mldivide,
    separate, is_inside, is_interior)
from polytope_b200.prop2partition import find_adjacent_regions

__version__ = '0.1.0'

"""ctypes binding of libpolytope_b200.so (C ABI: include/polytope_b200.h).

The library is the product; there is no Python/numpy/CPU fallback.  Importing
this module on a box where the library has not been built, or calling into it
without a CUDA device, raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PB200_LIB lets a developer A/B-test another build of the same library
LIB_PATH = os.environ.get('PB200_LIB') or os.path.join(_HERE, 'libpolytope_b200.so')

c_void_p, c_int, c_double = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
c_size_t, c_longlong, c_char_p = ctypes.c_size_t, ctypes.c_longlong, ctypes.c_char_p

# name -> (restype, argtypes); mirrors include/polytope_b200.h one to one
SIGNATURES = {
    'pb200_version': (c_char_p, []),
    'pb200_last_error': (c_char_p, []),
    'pb200_launch_count': (c_longlong, []),
    'pb200_lp_batch': (c_int, [c_void_p] * 4 + [c_int] * 3 + [c_void_p] * 5),
    'pb200_lp_big_workspace_bytes': (c_size_t, [c_int] * 3),
    'pb200_lp_batch_big': (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_void_p] * 5 + [c_size_t, c_void_p]),
    'pb200_normalize_batch': (c_int, [c_void_p] * 3 + [c_int] * 3 + [c_void_p] * 4),
    'pb200_cheby_batch': (c_int, [c_void_p] * 4 + [c_int] * 3 + [c_void_p] * 4),
    'pb200_bbox_batch': (c_int, [c_void_p] * 3 + [c_int] * 3 + [c_void_p] * 4),
    'pb200_reduce_workspace_bytes': (c_size_t, [c_int] * 3),
    'pb200_reduce_batch': (c_int, [c_void_p] * 3 + [c_int] * 3 + [c_double, c_int]
                           + [c_void_p] * 9 + [c_size_t, c_void_p]),
    'pb200_profile_enable': (None, [c_int]),
    'pb200_normalize_variant': (None, [c_int]),
    'pb200_lane_solver': (None, [c_int]),
    'pb200_measure_dfma_tflops': (c_int, [c_void_p, c_size_t, c_void_p, c_void_p]),
    'pb200_profile_read': (c_int, [c_void_p, c_int]),
    'pb200_contains_batch': (c_int, [c_void_p] * 3 + [c_int] * 3 + [c_void_p, c_longlong, c_double, c_int,
                                     c_void_p, c_void_p]),
    'pb200_volume_counts': (c_int, [c_void_p] * 3 + [c_int] * 3 + [c_void_p, c_void_p, c_longlong]
                            + [c_void_p] * 3),
    'pb200_point_facet_sweep': (c_int, [c_void_p] * 3 + [c_longlong, c_int, c_int, c_double] + [c_void_p] * 4),
    'pb200_hull_workspace_bytes': (c_size_t, [c_int] * 4),
    'pb200_hull_batch': (c_int, [c_void_p] * 2 + [c_int] * 3 + [c_double, c_int] + [c_void_p] * 3
                         + [c_longlong] + [c_void_p] * 7 + [c_size_t, c_void_p]),
    'pb200_dual_points': (c_int, [c_void_p] * 4 + [c_int] * 3 + [c_void_p] * 2),
    'pb200_dual_facets_to_vertices': (c_int, [c_void_p] * 5 + [c_int] * 3 + [c_void_p] * 2),
    'pb200_region_diff_batch': (c_int, [c_void_p] * 3 + [c_int] * 3 + [c_void_p] * 4 + [c_int] * 3
                                + [c_double] * 2 + [c_void_p] * 6 + [c_longlong, c_int] + [c_void_p] * 5
                                + [c_void_p]),
    'pb200_adjacent_pairs': (c_int, [c_void_p] * 2 + [c_int] * 3 + [c_void_p] * 2
                             + [c_longlong, c_double] + [c_void_p] * 4),
    'pb200_adjacent_range': (c_int, [c_void_p] * 2 + [c_int] * 4 + [c_longlong, c_longlong, c_double] + [c_void_p] * 4),
}


class Pb200Error(RuntimeError):
    pass


def load():
    if not os.path.exists(LIB_PATH):
        raise Pb200Error(
            'polytope_b200: %s is missing -- build it with '
            '`python -c "import __graft_entry__ as g; g.build()"` or '
            '`make -C polytope_b200/csrc`.  There is no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = load()
    return _lib


def check(rc, what):
    if rc != 0:
        raise Pb200Error('%s failed (%d): %s' % (
            what, rc, lib().pb200_last_error().decode()))

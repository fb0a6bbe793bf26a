"""Worker functions of the full-size parity tests (tests/test_gpu_fullsize.py): the CPU oracle
over every unit of a BASELINE config, spread over all host cores with a `spawn` pool (the parent
has a CUDA context, so no fork).  Test infrastructure; never imported by polytope_b200/.

Every worker regenerates its inputs from the seed rule of workloads.py, so only seeds travel to
the workers and plain ints / floats travel back.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for _p in (ROOT, HERE):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def _mask(idx):
    out = 0
    for k in idx:
        out |= 1 << int(k)
    return out


def reduce_chunk(args):
    """reduce() of polytopes [first, first+count) of a box+cuts config.
    -> list of (keep mask, n_lp, empty, r, xc list or None, min |margin - abs_tol|)."""
    cfg, first, count, m, d, shift_scale = args
    import workloads as wl
    from oracle import polytope_oracle as orc
    out = []
    for i in range(first, first + count):
        A, b = wl.box_cuts(1000 * cfg + i, m, d, shift_scale)
        o = orc.reduce(A, b)
        amb = min([abs(v - orc.ABS_TOL) for v in o['margins']] + [abs(float(o['r']) - orc.ABS_TOL)])
        out.append((_mask(o['keep']), int(o['n_lp']), bool(o['empty']), float(o['r']),
                    None if o['xc'] is None else [float(v) for v in o['xc']], float(amb)))
    return out


def fulldim_chunk(args):
    """cheby_ball of the constructor-normalised polytopes [first, first+count). -> list of r."""
    cfg, first, count, m, d, shift_scale = args
    import workloads as wl
    from oracle import polytope_oracle as orc
    out = []
    for i in range(first, first + count):
        A, b = wl.box_cuts(1000 * cfg + i, m, d, shift_scale)
        An, bn, _ = orc.normalize_rows(A, b)
        out.append(float(orc.cheby_ball(An, bn)[0]))
    return out


def intersect_chunk(args):
    """Polytope.intersect(member_i, Q) for members [first, first+count) of cfg3.
    -> list of (keep mask over the stacked rows, n_lp of the reduce, empty)."""
    cfg, first, count, m, d, qseed = args
    import workloads as wl
    from oracle import polytope_oracle as orc
    Q = orc.normalize_rows(*wl.box_cuts(qseed, m, d, True))[:2]
    out = []
    for i in range(first, first + count):
        An, bn, _ = orc.normalize_rows(*wl.box_cuts(1000 * cfg + i, m, d, True))
        o = orc.intersect(An, bn, Q[0], Q[1])
        out.append((_mask(o['keep']), int(o['n_lp']), bool(o['empty'])))
    return out


def adjacent_chunk(args):
    """is_adjacent of box-grid cell pairs. -> list of (flag, radius of the stacked Chebyshev LP)."""
    shape, pairs = args
    import workloads as wl
    from oracle import polytope_oracle as orc
    A, b, _ = wl.box_grid(shape)
    cells = {}

    def cell(i):
        if i not in cells:
            cells[i] = orc.normalize_rows(A[i], b[i])[:2]
        return cells[i]
    out = []
    for i, j in pairs:
        An, bn = orc.adjacent_lp_data(*cell(i), *cell(j))
        r, _ = orc.cheby_ball(An, bn)
        out.append((bool(r > orc.ABS_TOL / 10), float(r)))
    return out


def run_pool(fn, jobs, workers=None):
    """Map `fn` over `jobs` on a spawn pool; results in job order."""
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    workers = workers or os.cpu_count() or 1
    env_path = os.environ.get('PYTHONPATH', '')
    os.environ['PYTHONPATH'] = os.pathsep.join([ROOT, HERE] + ([env_path] if env_path else []))
    # one BLAS/OpenMP thread per worker: the LPs are tiny and the pool already fills the cores
    for var in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ.setdefault(var, '1')
    with ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context('spawn')) as ex:
        return list(ex.map(fn, jobs))


def chunks(total, per):
    return [(s, min(per, total - s)) for s in range(0, total, per)]

"""Drop-in at the plugin surface: the UNMODIFIED reference (oracle/_ref, installed
by oracle/make_ref.sh) runs its own L3 code -- reduce, cheby_ball, bounding_box, is_adjacent,
Polytope.intersect, region_diff, extreme -- with `polytope.polytope.lpsolve` (the name bound at
polytope/polytope.py:69) replaced by `polytope_b200.solvers.lpsolve`, i.e. every LP goes through
the C ABI to the sm_100a kernel, one LP per call.  Results are compared with the same reference
code on its stock scipy/HiGHS path, and the reference's own test-suite (tests/polytope_test.py,
including the `operations_test` class and the rotation tests that pytest never collects,
SURVEY.md section 4) is run on the patched package.
"""
import importlib.util
import os
import sys

import numpy as np
import pytest

import workloads as wl
from oracle import ref_loader

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_loader.available(), reason='oracle/_ref not built (oracle/make_ref.sh)')]


def _gpu_lpsolve():
    from polytope_b200 import solvers
    calls = [0]

    def fn(c, G, h, solver=None):
        calls[0] += 1
        return solvers.lpsolve(c, G, h)
    return fn, calls


def _rows_of(poly):
    return np.array(poly.A, dtype=float), np.array(poly.b, dtype=float)


def test_reference_reduce_bbox_adjacent_intersect_through_the_gpu_solver():
    pc = ref_loader.load()
    fn, calls = _gpu_lpsolve()
    cases = [(2000 + i, 32, 8, False) for i in range(6)] + [(3000 + i, 16, 6, True) for i in range(6)] \
        + [(4000 + i, 64, 12, False) for i in range(2)] + [(5000 + i, 10, 3, True) for i in range(6)]
    n_lp_ref = 0
    for seed, m, d, ss in cases:
        A, b = wl.box_cuts(seed, m, d, ss)
        with ref_loader.count_lps() as n:
            ref = pc.reduce(pc.Polytope(A.copy(), b.copy()))
            l0, u0 = pc.Polytope(A.copy(), b.copy()).bounding_box
            r0, x0 = pc.cheby_ball(pc.Polytope(A.copy(), b.copy()))
        n_lp_ref += n[0]
        before = calls[0]
        with ref_loader.patched_lpsolve(fn):
            got = pc.reduce(pc.Polytope(A.copy(), b.copy()))
            l1, u1 = pc.Polytope(A.copy(), b.copy()).bounding_box
            r1, x1 = pc.cheby_ball(pc.Polytope(A.copy(), b.copy()))
        assert calls[0] - before == n[0]                      # LP for LP the same call sequence
        assert np.array_equal(got.A, ref.A) and np.array_equal(got.b, ref.b), seed     # same rows kept, same drift
        assert got.minrep == ref.minrep
        np.testing.assert_allclose(l1, l0, rtol=1e-7, atol=1e-7)
        np.testing.assert_allclose(u1, u0, rtol=1e-7, atol=1e-7)
        assert abs(r1 - r0) <= 1e-9 * max(1.0, abs(r0))
    assert n_lp_ref > 500
    # Polytope.intersect and is_adjacent on pairs (polytope.py:255-275, :1827-1866)
    for i in range(8):
        A1, b1 = wl.box_cuts(3100 + i, 16, 6, True)
        A2, b2 = wl.box_cuts(3200 + i, 16, 6, True)
        ref = pc.Polytope(A1, b1).intersect(pc.Polytope(A2, b2))
        with ref_loader.patched_lpsolve(fn):
            got = pc.Polytope(A1, b1).intersect(pc.Polytope(A2, b2))
        assert np.array_equal(got.A, ref.A) and np.array_equal(got.b, ref.b)
    Ag, bg, idx = wl.box_grid((4, 4))
    cells = [pc.Polytope(Ag[i], bg[i]) for i in range(len(Ag))]
    ref = np.array([[pc.is_adjacent(p, q) for q in cells] for p in cells])
    with ref_loader.patched_lpsolve(fn):
        cells2 = [pc.Polytope(Ag[i], bg[i]) for i in range(len(Ag))]
        got = np.array([[pc.is_adjacent(p, q) for q in cells2] for p in cells2])
    assert np.array_equal(got, ref)
    assert np.array_equal(got, np.abs(idx[:, None] - idx[None]).max(2) <= 1)


def test_reference_region_ops_and_extreme_through_the_gpu_solver():
    """The sequential L3 algorithms that call cheby_ball / reduce in a data-dependent order
    (region_diff's search, extreme's hull) take the same path on both solvers."""
    pc = ref_loader.load()
    fn, calls = _gpu_lpsolve()
    for i in range(8):
        (A, b), cells = wl.diff_case(i)
        P = pc.Polytope(A, b)
        R = pc.Region([pc.Polytope(*c) for c in cells])
        ref = P.diff(R)
        with ref_loader.patched_lpsolve(fn):
            got = pc.Polytope(A, b).diff(pc.Region([pc.Polytope(*c) for c in cells]))
        ref_list = list(ref) if len(ref) else []
        got_list = list(got) if len(got) else []
        assert len(ref_list) == len(got_list), i
        for p, q in zip(got_list, ref_list):
            assert p.A.shape == q.A.shape
            np.testing.assert_allclose(p.A, q.A, atol=1e-9)
            np.testing.assert_allclose(p.b, q.b, atol=1e-9)
    for m, d in [(6, 2), (10, 3), (12, 4)]:
        A, b = wl.box_cuts(7000 + d, m, d)
        ref = pc.extreme(pc.Polytope(A, b))
        with ref_loader.patched_lpsolve(fn):
            got = pc.extreme(pc.Polytope(A, b))
        assert ref.shape == got.shape
        key = lambda V: V[np.lexsort(np.round(V, 6).T)]
        np.testing.assert_allclose(key(got), key(ref), atol=1e-7)
    assert calls[0] > 100


def _load_reference_tests(name):
    path = os.path.join(ref_loader.REF_TESTS, name + '.py')
    spec = importlib.util.spec_from_file_location('reference_' + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_own_testsuite_runs_on_the_gpu_solver():
    """/root/reference/tests/polytope_test.py, :26-651, unmodified, with the L3 code solving its LPs on
    the GPU.  pytest would only collect the `test_*` functions; the nose-style `operations_test`
    methods (:57-296) and the rotation functions (:315-418) are called by hand as SURVEY.md section 4
    describes.  projection_test.py's Fourier-Motzkin tests run too (reduce after each elimination);
    `test_projection_iterhull` needs basic (vertex) LP solutions and is out of scope."""
    pc = ref_loader.load()
    fn, calls = _gpu_lpsolve()
    with ref_loader.patched_lpsolve(fn):
        mod = _load_reference_tests('polytope_test')
        assert mod.pc is pc
        ran = []
        for name in sorted(dir(mod)):
            obj = getattr(mod, name)
            if not callable(obj) or getattr(obj, '__module__', None) != mod.__name__:
                continue
            if name.startswith('test_'):
                if name == 'test_gurobipy_return_same_result_as_scipy':
                    continue                                   # skipif-marked: gurobipy is not installed
                obj()
                ran.append(name)
            elif name.startswith(('solve_rotation_test', 'givens_rotation_test')):
                obj()
                ran.append(name)
        ops = mod.operations_test()
        for name in sorted(dir(ops)):
            if name.endswith('_test'):
                ops.setUp()
                getattr(ops, name)()
                ops.tearDown()
                ran.append('operations_test.' + name)
        proj = _load_reference_tests('projection_test')
        proj.test_fourier_motzkin_square()
        proj.test_fourier_motzkin_triangle()
        ran += ['test_fourier_motzkin_square', 'test_fourier_motzkin_triangle']
    assert len(ran) >= 30, ran
    assert calls[0] > 50, calls[0]                             # the suite really solved its LPs on the GPU


def test_oracle_restatement_equals_the_live_reference_on_this_box():
    """Pin of the oracle on the GPU box itself (the build container pins it through
    tests/test_ref_pin.py): reduce / is_adjacent of oracle/polytope_oracle.py against the installed
    reference, fresh seeds, bit for bit."""
    from test_ref_pin import pin_oracle_against_reference
    pin_oracle_against_reference(n_reduce=24, grid=(4, 3))

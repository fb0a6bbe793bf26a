"""CPU-side tests: the C ABI library loads and exports every declared symbol,
the host-side mirror of the reference interface behaves like the reference
where no LP is needed, and the product path fails loudly without a GPU."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import REPO

HAS_CUDA = torch.cuda.is_available()


def test_abi_exports_every_declared_symbol():
    from polytope_b200 import _capi
    header = open(os.path.join(REPO, 'include', 'polytope_b200.h')).read()
    declared = set(re.findall(r'\b(pb200_\w+)\s*\(', header))
    assert declared == set(_capi.SIGNATURES), declared ^ set(_capi.SIGNATURES)
    lib = _capi.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert b'sm_100a' in lib.pb200_version()
    assert lib.pb200_reduce_workspace_bytes(10, 32, 8) > 10 * 32 * 8 * 8


def test_library_is_sm100a_with_fp64_mma():
    """The shipped binary is sm_100a SASS and the hot reductions are DMMA."""
    import subprocess
    from polytope_b200 import _capi
    out = subprocess.run(['cuobjdump', '-lelf', _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out, out
    sass = subprocess.run(['cuobjdump', '-sass', '-fun', '_ZN5pb2009lp_kernelILi1ENS_5RowLPEEEvT0_xPKj',
                           _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert 'DMMA' in sass


def test_solver_selection_errors_match_reference():
    """solvers.py:103-105 (unknown name) and :200-207 (absent solver)."""
    from polytope_b200 import solvers
    c, A, b = np.array([1.]), np.array([[-1.]]), np.array([1.])
    assert solvers.installed_solvers == {'b200'} and solvers.default_solver == 'b200'
    for name in ('glpk', 'mosek', 'scipy', 'gurobi'):
        with pytest.raises(RuntimeError):
            solvers.lpsolve(c, A, b, solver=name)
    with pytest.raises(Exception, match='unknown LP solver "foo"'):
        solvers.lpsolve(c, A, b, solver='foo')


@pytest.mark.skipif(HAS_CUDA, reason='checks the no-GPU failure mode')
def test_no_gpu_means_loud_failure_not_fallback():
    from polytope_b200 import solvers, _capi
    import polytope_b200 as pb
    with pytest.raises(_capi.Pb200Error, match='no CPU fallback'):
        solvers.lpsolve(np.array([1.]), np.array([[-1.]]), np.array([1.]))
    with pytest.raises(_capi.Pb200Error):
        pb.reduce(pb.Polytope(np.vstack([np.eye(2), -np.eye(2)]), np.ones(4)))


def test_product_never_imports_oracle_or_scipy():
    """The package must not route through the oracle, scipy or any CPU LP solver."""
    pkg = os.path.join(REPO, 'polytope_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r'^\s*(from|import)\s+(oracle|scipy)', src, re.M), f
                assert not re.search(r'linprog\s*\(', src), f


def test_oracle_is_only_reachable_from_the_checkers():
    """oracle/ is test infrastructure: besides tests/, only smoke() and bench.py's CPU arm import it,
    and bench.py does so inside the CPU-arm functions, never at module level."""
    for f in os.listdir(os.path.join(REPO, 'tools')) + ['workloads.py']:
        if f.endswith('.py'):
            path = os.path.join(REPO, 'tools', f) if f != 'workloads.py' else os.path.join(REPO, f)
            assert not re.search(r'^\s*(from|import)\s+oracle', open(path).read(), re.M), f
    bench = open(os.path.join(REPO, 'bench.py')).read()
    assert not re.search(r'^(from|import)\s+oracle', bench, re.M)
    users = re.findall(r'^def (\w+)\([^\n]*\n(?:(?!^def ).*\n)*?\s+from oracle import', bench, re.M)
    assert users and all('cpu' in u or 'reference' in u or 'oracle' in u for u in users), users


def test_polytope_constructor_matches_reference_normalisation(golden):
    """Polytope.__init__ vs the oracle's restatement of polytope.py:122-138."""
    import polytope_b200 as pb
    from oracle import polytope_oracle as orc
    import workloads as wl
    for seed in range(20):
        A, b = wl.box_cuts(seed, 20, 5, True)
        A[3] = 0
        A[7] *= 1e-12
        p = pb.Polytope(A, b)
        An, bn, pos = orc.normalize_rows(A, b)
        assert np.array_equal(p.A, An) and np.array_equal(p.b, bn)
    q = pb.Polytope(np.array([[2, 0], [0, 4]]), np.array([2, 4]))     # integer input, :126-129
    assert np.array_equal(q.A, np.eye(2)) and np.array_equal(q.b, [1., 1.])
    raw = pb.Polytope(np.array([[2., 0]]), np.array([2.]), normalize=False)
    assert raw.A[0, 0] == 2.0


def test_polytope_str_pinned_like_reference():
    """tests/polytope_test.py:26-54 of the reference."""
    import polytope_b200 as pc
    p = pc.Polytope(np.array([[1]]), np.array([1]))
    assert str(p) == 'Single polytope \n  [[1.]] x <= [[1.]]\n'
    strings = dict(
        p1d='Single polytope \n  [[ 1.] x <= [[1.]\n   [-1.]]|     [0.]]\n',
        p2d=('Single polytope \n  [[ 1.  0.] |    [[1.]\n   [ 0.  1.] '
             'x <=  [2.]\n   [-1. -0.] |     [0.]\n   [-0. -1.]]|'
             '     [0.]]\n'))
    assert str(pc.Polytope.from_box([[0, 1]])) == strings['p1d']
    assert str(pc.Polytope.from_box([[0, 1], [0, 2]])) == strings['p2d']


def test_structural_host_logic():
    import polytope_b200 as pc
    assert pc.is_empty(pc.Polytope()) and not pc.is_empty(pc.box2poly([[0, 1]]))
    reg = pc.Region()
    reg.list_poly = [pc.Polytope(), pc.Polytope()]
    assert len(reg) > 0 and pc.is_empty(reg)               # region_empty_test
    reg = pc.Region([pc.box2poly([[0, 1]]), pc.Polytope()])
    assert len(reg) == 1                                    # ctor drops empty members
    box = pc.box2poly([[0.0, 1.0], [0.0, 2.0]])
    assert box.minrep and pc.reduce(box) is box             # minrep short-circuit, :1079-1080
    assert [0.1, 0.3] in box and [2, 0] not in box
    pts = np.array([[-1.0, 0.0, 0.5, 1.0, 2.0]])
    r1 = pc.Region([pc.Polytope(np.array([[1.0], [-1.0]]), np.array([1.0, 0.0]))])
    # contains() streams the points through the device kernel: loud failure without a GPU
    from polytope_b200._capi import Pb200Error
    with pytest.raises(Pb200Error):
        r1.contains(pts)
    with pytest.raises(ValueError):
        r1.contains(np.zeros((2, 3)))
    assert pc.cheby_ball(pc.Polytope()) == (0, None)
    with pytest.raises(Exception):
        pc.Polytope.from_box([[1, 0]])


def test_ipm_model_matches_scipy_on_golden_lps(golden):
    """The numpy model of the kernel algorithm (tests/ipm_model.py) against the
    reference's recorded LP results -- the CPU-side check of the algorithm."""
    import ipm_model
    g = golden('lp_cases')
    idx = list(range(int(g['n_hand']))) + list(range(int(g['n_hand']), len(g['status']), 9))
    for i in idx:
        m, n = g['shape'][i]
        r = ipm_model.solve_lp(g['C'][i, :n], g['G'][i, :m, :n], g['H'][i, :m])
        assert r['status'] == g['status'][i], i
        if r['status'] == 0:
            assert abs(r['fun'] - g['fun'][i]) <= 1e-9 * (1 + abs(g['fun'][i]))
    r = ipm_model.solve_lp(np.array([1.]), np.array([[-1.]]), np.array([1.]))
    assert r['x'][0] == -1.0


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: the header must compile as C (and as C++) on its own."""
    import subprocess
    hdr = os.path.join(REPO, 'include', 'polytope_b200.h')
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', '-fsyntax-only', '-x', 'c', hdr])
    subprocess.check_call(['g++', '-std=c++11', '-Wall', '-Werror', '-fsyntax-only', '-x', 'c++', hdr])

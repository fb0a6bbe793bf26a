"""Pin of the oracle against the LIVE unmodified reference installed in oracle/_ref (the golden
fixtures pin it against recorded outputs; this pins it on fresh seeds, on whatever box runs the
suite -- scipy/HiGHS is the LP solver on both sides, so results are compared bit for bit)."""
import numpy as np
import pytest

import workloads as wl
from oracle import polytope_oracle as orc
from oracle import ref_loader

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason='oracle/_ref not built (oracle/make_ref.sh)')


def _match_rows(A_in, b_in, A_out, b_out):
    keep = []
    for a, bb in zip(A_out, b_out):
        dist = np.abs(A_in - a).sum(1) + np.abs(b_in - bb)
        k = int(np.argmin(dist))
        assert dist[k] < 1e-12
        keep.append(k)
    return keep


def pin_oracle_against_reference(n_reduce=12, grid=(3, 3), seed0=881000):
    pc = ref_loader.load()
    for i in range(n_reduce):
        m, d, ss = [(32, 8, False), (16, 6, True), (20, 5, True), (64, 12, False)][i % 4]
        if m == 64 and i >= 8:
            m, d = 24, 7
        A, b = wl.box_cuts(seed0 + i, m, d, ss)
        with ref_loader.count_lps() as n:
            poly = pc.Polytope(A.copy(), b.copy())
            red = pc.reduce(poly)
        o = orc.reduce(A, b)
        assert o['n_lp'] == n[0], (i, o['n_lp'], n[0])
        An, bn, _ = orc.normalize_rows(A, b)
        # row matching on the normalised rows (+ the reference's one-ulp b drift, :1149-1151)
        assert _match_rows(An, bn, red.A, red.b) == o['keep'], i
        On, obn, _ = orc.normalize_rows(o['A'], o['b'])
        assert np.array_equal(On, red.A) and np.array_equal(obn, red.b)
        assert o['minrep'] == red.minrep
        r, xc = pc.cheby_ball(pc.Polytope(A.copy(), b.copy()))
        assert r == o['r']
    Ag, bg, idx = wl.box_grid(grid)
    cells = [pc.Polytope(Ag[i], bg[i]) for i in range(len(Ag))]
    ref = np.array([[pc.is_adjacent(p, q) for q in cells] for p in cells])
    got = np.array([[orc.is_adjacent(p.A, p.b, q.A, q.b) for q in cells] for p in cells])
    assert np.array_equal(ref, got)
    return True


@needs_ref
def test_oracle_equals_the_live_reference_on_fresh_seeds():
    assert pin_oracle_against_reference()


@needs_ref
def test_reference_is_the_unmodified_install():
    """oracle/_ref holds what oracle/make_ref.sh installed: the package resolves there, and its
    own test-suite passes on the stock scipy path (same result as SURVEY.md section 4: the 13 that
    do not need matplotlib / gurobipy)."""
    import os
    pc = ref_loader.load()
    assert os.path.dirname(pc.__file__).startswith(ref_loader.REF_DIR)
    assert pc.solvers.default_solver == 'scipy' and pc.solvers.installed_solvers == {'scipy'}
    assert os.path.exists(os.path.join(ref_loader.REF_DIR, 'SOURCES.sha256'))

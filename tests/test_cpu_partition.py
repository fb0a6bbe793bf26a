"""Host logic of `separate` and `find_adjacent_regions` against fixtures recorded from the
reference (tests/golden/make_golden_partition.py).

Both functions are pure bookkeeping over one `engine.adjacent_pairs` launch (the cfg5 kernel, whose
flags the GPU tests pin to the oracle).  No GPU is needed here: the launch is replaced by the
oracle's pairwise test in the same pair order, so what is checked is the replay of the reference's
grouping / OR-reduction, not the kernel.
"""
import numpy as np
import pytest

import workloads as wl


@pytest.fixture()
def oracle_adjacent_pairs(monkeypatch):
    from oracle import polytope_oracle as orc
    from polytope_b200 import engine

    def fake(A, b, pair_i=None, pair_j=None, abs_tol=1e-7):
        assert pair_i is None and pair_j is None
        flags = [orc.is_adjacent(A[i], b[i], A[j], b[j], abs_tol=abs_tol)
                 for i in range(len(A)) for j in range(i)]          # prop2partition.py:57-61 order
        return np.array(flags, dtype=np.uint8), None, None

    monkeypatch.setattr(engine, 'adjacent_pairs', fake)


def _build(pb, cells, groups):
    polys = [pb.Polytope(A, b) for A, b in cells]
    regions = [pb.Region([polys[i] for i in g], ['p%d' % k]) for k, g in enumerate(groups)]
    return polys, regions


def test_separate_replays_the_reference_grouping(golden, oracle_adjacent_pairs):
    import polytope_b200 as pb
    g = golden('partition_cases')
    for name, cells, groups in wl.partition_scenarios():
        polys, regions = _build(pb, cells, groups)
        for k, (grp, reg) in enumerate(zip(groups, regions)):
            parts = pb.separate(reg)
            ident = {id(polys[i]): i for i in grp}
            assert [len(p) for p in parts] == g['%s_sep%d_sizes' % (name, k)].tolist(), (name, k)
            flat = [ident[id(p)] for part in parts for p in part.list_poly]
            assert flat == g['%s_sep%d_members' % (name, k)].tolist(), (name, k)
            assert all(part.props == reg.props and part.props is not reg.props for part in parts)
    # the chain listed out of order is split although it is connected (single ordered pass)
    assert g['chain_out_of_order_sep0_sizes'].tolist() == [2, 1]


def test_find_adjacent_regions_matches_reference(golden, oracle_adjacent_pairs):
    import polytope_b200 as pb
    g = golden('partition_cases')

    class Part(object):
        def __init__(self, regions):
            self.regions = regions

    for name, cells, groups in wl.partition_scenarios():
        _, regions = _build(pb, cells, groups)
        adj = pb.find_adjacent_regions(Part(regions))
        assert adj.dtype == np.int8 and np.array_equal(adj, g['%s_adj' % name]), name
        assert np.array_equal(pb.find_adjacent_regions(regions), adj)      # plain sequence form
    one = pb.find_adjacent_regions([pb.box2poly([[0, 1], [0, 1]])])
    assert np.array_equal(one, np.eye(1, dtype=np.int8))
    with pytest.raises(Exception, match='same dimension'):
        pb.find_adjacent_regions([pb.box2poly([[0, 1]]), pb.box2poly([[0, 1], [0, 1]])])


def test_is_inside_warns_and_goes_to_the_device():
    import torch
    import polytope_b200 as pb
    from polytope_b200 import _capi
    box = pb.box2poly([[0, 1], [0, 1]])
    with pytest.warns(DeprecationWarning):
        if torch.cuda.is_available():
            assert pb.is_inside(box, [0.5, 0.5]) and not pb.is_inside(box, [1.5, 0.5])
        else:                                   # no CPU fallback: the call must fail loudly
            with pytest.raises(_capi.Pb200Error):
                pb.is_inside(box, [0.5, 0.5])


def test_bounding_box_to_polytope_rebuilds_the_box():
    """Structure of reference test_bounding_box_to_polytope (tests/polytope_test.py:299-312); the
    bounding-box LPs themselves are GPU work (tests/test_gpu_polytope.py)."""
    import polytope_b200 as pb
    from polytope_b200 import polytope as alg
    for intervals in ([[0, 1]], [[0, 1], [0, 2]], [[-1, 2], [3, 5], [-5, -3]]):
        iv = np.array(intervals, dtype=float)
        poly = pb.box2poly(intervals)
        back = alg._bounding_box_to_polytope(iv[:, :1], iv[:, 1:])
        assert np.array_equal(back.A, poly.A) and np.array_equal(back.b, poly.b) and back.minrep

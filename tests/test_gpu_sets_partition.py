"""GPU run of the is_adjacent callers `separate` (polytope.py:1795-1824) and `find_adjacent_regions`
(prop2partition.py:46-63) against the fixtures recorded from the unmodified reference
(tests/golden/make_golden_partition.py): one adjacency launch per call, then the host replay that
tests/test_cpu_partition.py pins.  Part membership, part order and matrix entries must be identical.
"""
import numpy as np
import pytest

import workloads as wl

pytestmark = pytest.mark.gpu


def test_separate_and_find_adjacent_regions_vs_reference(golden):
    import polytope_b200 as pb
    from polytope_b200 import engine
    g = golden('partition_cases')
    launches0 = engine.launch_count()
    calls = 0
    for name, cells, groups in wl.partition_scenarios():
        polys = [pb.Polytope(A, b) for A, b in cells]
        regions = [pb.Region([polys[i] for i in grp], ['p%d' % k]) for k, grp in enumerate(groups)]
        for k, (grp, reg) in enumerate(zip(groups, regions)):
            parts = pb.separate(reg)
            calls += len(grp) > 1
            ident = {id(polys[i]): i for i in grp}
            assert [len(p) for p in parts] == g['%s_sep%d_sizes' % (name, k)].tolist(), (name, k)
            flat = [ident[id(p)] for part in parts for p in part.list_poly]
            assert flat == g['%s_sep%d_members' % (name, k)].tolist(), (name, k)
            assert all(part.props == reg.props for part in parts)
        adj = pb.find_adjacent_regions(regions)
        calls += len(regions) > 1          # a single region needs no pair test
        assert np.array_equal(adj, g['%s_adj' % name]), name
    # every call is exactly one kernel launch (no per-pair LP calls)
    assert engine.launch_count() - launches0 == calls

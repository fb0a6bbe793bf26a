#!/usr/bin/env python
"""Golden fixtures for `separate` (polytope.py:1795-1824) and `find_adjacent_regions`
(prop2partition.py:46-63), recorded from the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    cd /tmp && python -u /root/repo/tests/golden/make_golden_partition.py

Inputs come from workloads.partition_scenarios(); stored are, per scenario, the member indices of
every part `separate` returns for every region (in the reference's order) and the dense adjacency
matrix of the partition.
"""
import os
import sys
import logging

import numpy as np
import scipy

logging.disable(logging.CRITICAL)
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, '/root/reference')
sys.path.insert(1, REPO)

import polytope as pc                                       # noqa: E402  (the reference)
from polytope.prop2partition import find_adjacent_regions  # noqa: E402
import workloads as wl                                      # noqa: E402

assert pc.__file__.startswith('/root/reference'), pc.__file__


class _Partition(object):      # what find_adjacent_regions needs of a Partition: len() and .regions
    def __init__(self, regions):
        self.regions = regions

    def __len__(self):
        return len(self.regions)


def main():
    out = {'meta': np.array(repr(dict(scipy=scipy.__version__, numpy=np.__version__,
                                      reference='tulip-control/polytope @ /root/reference (v0.2.6.dev0)')))}
    for name, cells, groups in wl.partition_scenarios():
        polys = [pc.Polytope(A, b) for A, b in cells]
        regions = [pc.Region([polys[i] for i in g], ['p%d' % k]) for k, g in enumerate(groups)]
        for k, (g, reg) in enumerate(zip(groups, regions)):
            parts = pc.separate(reg)
            ident = {id(polys[i]): i for i in g}
            flat, sizes = [], []
            for part in parts:
                assert part.props == reg.props
                sizes.append(len(part))
                flat += [ident[id(p)] for p in part.list_poly]
            out['%s_sep%d_members' % (name, k)] = np.array(flat, dtype=np.int64)
            out['%s_sep%d_sizes' % (name, k)] = np.array(sizes, dtype=np.int64)
        adj = find_adjacent_regions(_Partition(regions))
        out['%s_adj' % name] = np.asarray(adj.todense(), dtype=np.int8)
        print(name, [out['%s_sep%d_sizes' % (name, k)].tolist() for k in range(len(groups))],
              int(out['%s_adj' % name].sum()))
    np.savez_compressed(os.path.join(HERE, 'partition_cases.npz'), **out)


if __name__ == '__main__':
    main()

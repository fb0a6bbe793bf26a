#!/usr/bin/env python
"""Golden fixtures for region_diff / mldivide / envelope / is_convex / union /
is_subset, recorded from the UNMODIFIED reference.

    cd /tmp && python -u /root/repo/tests/golden/make_golden_diff.py
"""
import os
import signal
import sys
import logging

import numpy as np
import scipy

logging.disable(logging.CRITICAL)
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, '/root/reference')
sys.path.insert(1, REPO)

import polytope as pc                     # noqa: E402  (the reference)
import workloads as wl                    # noqa: E402

assert pc.__file__.startswith('/root/reference'), pc.__file__
META = dict(scipy=scipy.__version__, numpy=np.__version__,
            reference='tulip-control/polytope @ /root/reference (v0.2.6.dev0)')


def pieces_of(x):
    """A Polytope / Region result as a list of (A, b)."""
    if len(x) == 0:
        return [] if len(x.A) == 0 else [(x.A, x.b)]
    return [(p.A, p.b) for p in x.list_poly]


def pack(out, tag, x):
    ps = pieces_of(x)
    out[tag + '_kind'] = np.array([0 if len(x) == 0 else 1])       # 0: Polytope, 1: Region
    out[tag + '_n'] = np.array([len(ps)])
    mr = max([len(b) for _, b in ps] + [1])
    d = ps[0][0].shape[1] if ps else 1
    A = np.full((len(ps), mr, d), np.nan)
    b = np.full((len(ps), mr), np.nan)
    for k, (a_, b_) in enumerate(ps):
        A[k, :len(b_)] = a_
        b[k, :len(b_)] = b_
    out[tag + '_A'], out[tag + '_b'] = A, b


def main():
    out = dict(meta=str(META))
    for i in range(wl.DIFF_CASES):
        (A, b), cells = wl.diff_case(i)
        poly = pc.Polytope(A, b)
        reg = pc.Region([pc.Polytope(a_, b_) for a_, b_ in cells])
        signal.alarm(300)
        res = pc.polytope.region_diff(poly, reg)
        signal.alarm(0)
        pack(out, 'diff%d' % i, res)
        out['diff%d_same' % i] = np.array([res is not poly and len(res) == 0 and np.array_equal(res.A, poly.A)])
        print('diff', i, 'd', A.shape[1], 'cells', len(cells), '->', 'Region' if len(res) else 'Polytope', len(pieces_of(res)))
    # boxes: the textbook cases
    B = lambda iv: pc.box2poly(iv)
    sq = B([[0, 2], [0, 2]])
    pack(out, 'box_corner', sq.diff(B([[1, 3], [1, 3]])))
    pack(out, 'box_hole', sq.diff(B([[0.5, 1.5], [0.5, 1.5]])))
    pack(out, 'box_covered', sq.diff(B([[-1, 3], [-1, 3]])))
    pack(out, 'box_far', sq.diff(B([[5, 6], [5, 6]])))
    pack(out, 'box_two', pc.polytope.region_diff(sq, pc.Region([B([[0.5, 1], [0.5, 1]]), B([[1.2, 1.8], [-1, 3]])])))
    pack(out, 'box3', B([[0, 1], [0, 1], [0, 1]]).diff(B([[0.5, 2], [0.5, 2], [-1, 2]])))
    # Region minus Polytope, with the convex re-merge of union(check_convex=True)
    L = pc.Region([B([[0, 1], [0, 2]]), B([[1, 2], [0, 1]])])
    signal.alarm(600)
    pack(out, 'reg_minus', L.diff(B([[0.5, 1.5], [0.5, 1.5]])))
    # envelope / is_convex / union
    two = pc.Region([B([[0, 1], [0, 1]]), B([[1, 2], [0, 1]])])
    pack(out, 'env_two', pc.envelope(two))
    pack(out, 'env_L', pc.envelope(L))
    conv, env = pc.is_convex(two)
    out['convex_two'] = np.array([conv])
    pack(out, 'convex_two_env', env)
    out['convex_L'] = np.array([pc.is_convex(L)[0]])
    pack(out, 'union_cc', pc.union(B([[0, 1], [0, 1]]), B([[1, 2], [0, 1]]), check_convex=True))
    pack(out, 'union_overlap_cc', pc.union(B([[0, 2], [0, 2]]), B([[1, 3], [1, 3]]), check_convex=True))
    pack(out, 'union_plain', pc.union(B([[0, 1], [0, 1]]), B([[3, 4], [0, 1]])))
    for i in range(4):
        (A, b), cells = wl.diff_case(i)
        reg = pc.Region([pc.Polytope(A, b)] + [pc.Polytope(a_, b_) for a_, b_ in cells])
        pack(out, 'env%d' % i, pc.envelope(reg))
    # is_subset / == / Region.intersect
    out['subset'] = np.array([B([[0.2, 0.8], [0.2, 0.8]]) <= sq, sq <= B([[0.2, 0.8], [0.2, 0.8]]),
                              L <= sq, sq <= L, two == B([[0, 2], [0, 1]]), sq == sq.copy()])
    pack(out, 'reg_isect', L.intersect(B([[0.5, 1.5], [0.5, 1.5]])))
    pack(out, 'reg_and', L & pc.Region([B([[0.5, 3], [0.25, 0.75]])]))
    signal.alarm(0)
    np.savez_compressed(os.path.join(HERE, 'diff_cases.npz'), **out)
    print('diff_cases.npz', os.path.getsize(os.path.join(HERE, 'diff_cases.npz')), 'bytes')


if __name__ == '__main__':
    main()

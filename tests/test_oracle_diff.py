"""The oracle's region_diff / envelope restatements against the golden vectors
recorded from the unmodified reference (tests/golden/make_golden_diff.py)."""
import numpy as np

import workloads as wl
from oracle import polytope_oracle as orc


def final_pieces(kind, pieces, poly):
    """What the reference's union(res, piece, False) chain holds at the end."""
    if kind == 'poly':
        return [poly]
    out = []
    for A, b, reduced in pieces:
        if reduced:
            red = orc.reduce(A, b)
            if red['empty']:
                continue
            An, bn, _ = orc.normalize_rows(red['A'], red['b'])
        else:
            An, bn, _ = orc.normalize_rows(A, b)
        out.append((An, bn))
    return out


def check_against_golden(g, tag, pieces):
    assert int(g[tag + '_n'][0]) == len(pieces), (tag, int(g[tag + '_n'][0]), len(pieces))
    assert int(g[tag + '_kind'][0]) == (1 if len(pieces) >= 2 else 0), tag
    for k, (A, b) in enumerate(pieces):
        rows = int(np.sum(~np.isnan(g[tag + '_b'][k])))
        assert rows == len(b), (tag, k, rows, len(b))
        np.testing.assert_allclose(A, g[tag + '_A'][k][:rows], atol=1e-9)
        np.testing.assert_allclose(b, g[tag + '_b'][k][:rows], atol=1e-9)


def test_region_diff_oracle_matches_reference(golden):
    g = golden('diff_cases')
    for i in range(wl.DIFF_CASES):
        (A, b), cells = wl.diff_case(i)
        poly = orc.normalize_rows(A, b)[:2]
        cells = [orc.normalize_rows(a_, b_)[:2] for a_, b_ in cells]
        kind, pieces = orc.region_diff(poly, cells)
        check_against_golden(g, 'diff%d' % i, final_pieces(kind, pieces, poly))
    sq = wl.box_rows([[0, 2], [0, 2]])
    for tag, cells in [('box_corner', [[[1, 3], [1, 3]]]), ('box_hole', [[[0.5, 1.5], [0.5, 1.5]]]),
                       ('box_covered', [[[-1, 3], [-1, 3]]]), ('box_far', [[[5, 6], [5, 6]]]),
                       ('box_two', [[[0.5, 1], [0.5, 1]], [[1.2, 1.8], [-1, 3]]])]:
        kind, pieces = orc.region_diff(sq, [wl.box_rows(c) for c in cells])
        check_against_golden(g, tag, final_pieces(kind, pieces, sq))


def test_envelope_oracle_matches_reference(golden):
    g = golden('diff_cases')
    for i in range(4):
        (A, b), cells = wl.diff_case(i)
        cells = [orc.normalize_rows(A, b)[:2]] + [orc.normalize_rows(a_, b_)[:2] for a_, b_ in cells]
        red = orc.envelope(cells)
        pieces = [] if red['empty'] else [orc.normalize_rows(red['A'], red['b'])[:2]]
        check_against_golden(g, 'env%d' % i, pieces)
    two = [wl.box_rows([[0, 1], [0, 1]]), wl.box_rows([[1, 2], [0, 1]])]
    red = orc.envelope(two)
    check_against_golden(g, 'env_two', [orc.normalize_rows(red['A'], red['b'])[:2]])

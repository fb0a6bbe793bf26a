"""The oracle's contains / volume / grid / qhull / extreme restatements against
the golden vectors recorded from the unmodified reference
(tests/golden/make_golden_sets.py), plus the two facts the kernels rely on:
numpy's A.dot(X) is an in-order fma chain, and default_rng().random() is the
PCG64 stream restated in oracle.pcg64_uniform."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

import workloads as wl
from oracle import polytope_oracle as orc


def rows_as_set_close(X, Y, tol):
    """Every row of X has a row of Y within tol (max-norm) and vice versa."""
    X, Y = np.atleast_2d(X), np.atleast_2d(Y)
    D = np.abs(X[:, None, :] - Y[None, :, :]).max(-1)
    return D.min(1).max() <= tol and D.min(0).max() <= tol


def test_pcg64_restatement_is_numpy_stream():
    for seed in (0, 1, 12345, 2**63 + 11):
        st = np.random.default_rng(seed).bit_generator.state['state']
        ref = np.random.default_rng(seed).random(64)
        assert np.array_equal(orc.pcg64_uniform(st['state'], st['inc'], 64), ref)
        assert np.array_equal(orc.pcg64_uniform(st['state'], st['inc'], 10, skip=54), ref[54:])


def test_numpy_dot_is_inorder_fma_chain():
    """contains()/volume() kernels accumulate A.x as acc = fma(A[i,k], x[k], acc);
    this is bit-identical to numpy's A.dot(X) (OpenBLAS dgemm micro-kernel) for
    d <= 15, m <= 64 and N > 1 (probed exhaustively on this image's OpenBLAS
    0.3.30; for d >= 16 the k-unrolled tail of the micro-kernel reorders a few
    columns by <= 4 ulp, which can flip a decision only within 1e-15 of the
    threshold)."""
    src = ('#include <math.h>\n'
           'void chain(const double*A,const double*X,int m,int d,int N,double*o){'
           'for(int i=0;i<m;i++)for(int j=0;j<N;j++){double a=0;for(int k=0;k<d;k++)a=fma(A[i*d+k],X[k*N+j],a);o[i*N+j]=a;}}')
    with tempfile.TemporaryDirectory() as tmp:
        c = os.path.join(tmp, 'c.c')
        so = os.path.join(tmp, 'c.so')
        open(c, 'w').write(src)
        subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-mfma', '-shared', '-fPIC', c, '-o', so, '-lm'])
        lib = ctypes.CDLL(so)
        rng = np.random.default_rng(0)
        for m, d, N in [(6, 3, 3000), (32, 8, 10000), (64, 12, 2000), (5, 1, 50), (16, 15, 777), (64, 6, 500)]:
            A = rng.standard_normal((m, d))
            X = rng.random((d, N))
            out = np.empty((m, N))
            lib.chain(A.ctypes.data_as(ctypes.c_void_p), X.ctypes.data_as(ctypes.c_void_p), m, d, N,
                      out.ctypes.data_as(ctypes.c_void_p))
            assert np.array_equal(out, A.dot(X)), (m, d, N)


def test_volume_contains_match_reference(golden):
    g = golden('sets_cases')
    for m, d in wl.VOLUME_SPECS:
        x = wl.contains_points(50 + d, d, 2000)
        cells = []
        for i in range(4):
            A, b = wl.box_cuts(8000 + 10 * d + i, m, d, True)
            An, bn, _ = orc.normalize_rows(A, b)
            vol, cnt, N = orc.volume(An, bn, seed=100 + i)
            assert vol == g['vol_d%d' % d][i]
            if i < 3:
                cells.append((An, bn))
                assert np.array_equal(orc.contains(An, bn, x), g['contains_d%d' % d][i])
        assert np.array_equal(orc.region_contains(cells, x), g['region_contains_d%d' % d])
        assert np.array_equal(orc.contains(*cells[0], x, abs_tol=0), g['contains_tol0_d%d' % d])
        assert orc.volume(*cells[0], nsamples=777, seed=5)[0] == g['vol_d%d_n777' % d]


def test_qhull_oracle_matches_reference_as_sets(golden):
    g = golden('hull_cases')
    for n, d in wl.HULL_SPECS:
        for i in range(3):
            A, b, vert = orc.qhull(wl.hull_points(600 + 10 * d + i, n, d))
            ref = np.c_[g['hull_d%d_%d_A' % (d, i)], g['hull_d%d_%d_b' % (d, i)]]
            assert len(b) == len(ref)
            assert rows_as_set_close(np.c_[A, b], ref, 1e-9)
            assert rows_as_set_close(vert, g['hull_d%d_%d_vert' % (d, i)], 1e-12)
    assert len(orc.qhull(np.eye(3))[1]) == int(g['hull_few_empty'][0]) == 0
    flat = np.c_[wl.hull_points(1, 10, 2), np.zeros(10)]
    assert len(orc.qhull(flat)[1]) == int(g['hull_flat_empty'][0]) == 0


def test_extreme_oracle_matches_reference_as_sets(golden):
    g = golden('hull_cases')
    for m, d in wl.EXTREME_SPECS:
        if d < 3:
            continue
        for i in range(3):
            V = orc.extreme(*wl.box_cuts(8500 + 10 * d + i, m, d, True))
            ref = g['ext_d%d_%d' % (d, i)]
            assert V.shape == ref.shape
            assert rows_as_set_close(V, ref, 1e-7)
    V = orc.extreme(*wl.unit_cube3())
    assert rows_as_set_close(V, g['ext_cube3'], 1e-9)

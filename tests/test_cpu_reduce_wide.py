"""Host logic of `reduce()` for polytopes beyond the fused pipeline's envelope (`polytope._reduce_wide`: more than
64 rows) against the oracle, without a GPU: the three device LP batches it issues (`engine.cheby_batch`,
`engine.bbox_batch`, `engine.lp_batch` with a shared G) are replaced by the oracle's `lpsolve`, so what is
checked is the restated step sequence of polytope.py:1081-1163 -- duplicate filter, early exits, bounding-box
candidates, the drifted right-hand sides of the row LPs -- not the kernels (tests/test_gpu_envelope.py pins those)."""
import numpy as np
import pytest

import workloads as wl


@pytest.fixture()
def oracle_lp_batches(monkeypatch):
    from oracle import polytope_oracle as orc
    from polytope_b200 import engine
    calls = {'lp': 0, 'shared': 0}

    def cheby_batch(A, b, m_rows=None, rows=None):
        r, xc, st = [], [], []
        for k in range(len(A)):
            mk = len(b[k]) if m_rows is None else int(m_rows[k])
            c, G, h = orc.cheby_lp_data(A[k][:mk], b[k][:mk])
            sol = orc.lpsolve(c, G, h)
            st.append(sol['status'])
            ok = sol['status'] == 0
            r.append(sol['x'][-1] if ok else np.nan)
            xc.append(sol['x'][:-1] if ok else np.full(A.shape[2], np.nan))
        return np.array(r), np.array(xc), np.array(st, dtype=np.int8)

    def bbox_batch(A, b, m_rows=None):
        lo, hi, st = [], [], []
        for k in range(len(A)):
            mk = len(b[k]) if m_rows is None else int(m_rows[k])
            l, u = orc.bounding_box(A[k][:mk], b[k][:mk])
            lo.append(l.ravel())
            hi.append(u.ravel())
            st.append(np.zeros(2 * A.shape[2], dtype=np.int8))
        return np.array(lo), np.array(hi), np.array(st)

    def lp_batch(C, G, H, m_rows=None):
        assert m_rows is None
        calls['lp'] += 1
        calls['shared'] += int(np.ndim(G) == 2)
        st, X, fun = [], [], []
        for k in range(len(C)):
            sol = orc.lpsolve(C[k], G if np.ndim(G) == 2 else G[k], H[k])
            st.append(sol['status'])
            ok = sol['status'] == 0
            X.append(sol['x'] if ok else np.full(C.shape[1], np.nan))
            fun.append(sol['fun'] if ok else np.nan)
        return np.array(st, dtype=np.int8), np.array(X), np.array(fun), np.zeros(len(C), dtype=np.int32)

    monkeypatch.setattr(engine, 'cheby_batch', cheby_batch)
    monkeypatch.setattr(engine, 'bbox_batch', bbox_batch)
    monkeypatch.setattr(engine, 'lp_batch', lp_batch)
    return calls


@pytest.mark.parametrize('seed,m,d', [(9100, 100, 5), (9102, 72, 9), (9104, 80, 2), (9105, 66, 30)])
def test_reduce_wide_replays_the_reference_steps(oracle_lp_batches, seed, m, d):
    import polytope_b200 as pc
    from oracle import polytope_oracle as orc
    A, b = wl.box_cuts(seed, m, d, True)
    p = pc.Polytope(A, b)
    red = pc.reduce(p)
    o = orc.reduce(A, b)
    An, bn, _ = orc.normalize_rows(A, b)
    Ak, bk, _ = orc.normalize_rows(An[o['keep']], o['b'])      # the reference's final Polytope(A_arr[keep], b_arr[keep])
    assert red.minrep == o['minrep']
    assert np.array_equal(red.A, Ak) and np.array_equal(red.b, bk)
    # every row LP went out in calls that share one G
    assert oracle_lp_batches['lp'] == oracle_lp_batches['shared']


def test_reduce_wide_duplicates_and_early_exit(oracle_lp_batches):
    """Duplicate hyperplanes beyond 64 rows: the tighter copy survives (polytope.py:1097-1109), and a polytope that
    is left with at most d + 1 rows returns before any row LP (:1113-1116)."""
    import polytope_b200 as pc
    from oracle import polytope_oracle as orc
    A0 = np.vstack([np.eye(2), -np.eye(2)])
    b0 = np.array([1.0, 2.0, 0.0, 0.0])
    A = np.vstack([A0] + [A0 * (1.0 + 0.01 * k) for k in range(1, 20)])
    b = np.hstack([b0 + 0.5] + [(b0 + 0.01 * k) * (1.0 + 0.01 * k) for k in range(1, 20)])
    p = pc.Polytope(A, b)
    assert p.A.shape[0] == 80
    red = pc.reduce(p)
    o = orc.reduce(A, b)
    An, _, _ = orc.normalize_rows(A, b)
    Ak, bk, _ = orc.normalize_rows(An[o['keep']], o['b'])
    assert np.array_equal(red.A, Ak) and np.array_equal(red.b, bk)
    assert len(red.b) == 4


def test_adjacency_of_wide_cells_stacks_the_pairs_on_the_host(oracle_lp_batches):
    """Cells with more than 32 rows: `adjacency_matrix` builds the stacked, inflated polytope of is_adjacent
    (polytope.py:1856-1866) per pair and asks for one batch of Chebyshev LPs; flags equal the oracle's."""
    import polytope_b200 as pc
    from oracle import polytope_oracle as orc
    rng = np.random.default_rng(7)
    cells = []
    for k in range(4):
        lo = np.array([float(k if k < 3 else 6), 0.0, 0.0])
        C = rng.standard_normal((34, 3))
        C /= np.linalg.norm(C, axis=1)[:, None]
        A = np.vstack([np.eye(3), -np.eye(3), C])
        b = np.hstack([lo + 1.0, -lo, C @ (lo + 0.5) + 3.0])
        cells.append(pc.Polytope(A, b))
    assert all(c.A.shape[0] == 40 for c in cells)
    adj = pc.adjacency_matrix(cells)
    assert np.array_equal(adj, orc.adjacency_matrix([(c.A, c.b) for c in cells]))
    assert adj[0, 1] == 1 and adj[1, 2] == 1 and adj[0, 2] == 0 and adj[2, 3] == 0

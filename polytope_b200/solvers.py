"""LP plugin surface, mirroring `polytope.solvers` of the reference.

Same call, same result dictionary, same module attributes
(/root/reference/polytope/solvers.py:39, :66-73, :76-106), one backend:
`'b200'`, the hand-written sm_100a interior-point kernel in
libpolytope_b200.so.  There is deliberately no scipy / GLPK / CPU path here:
asking for one raises exactly like the reference does for a solver that is not
installed (`RuntimeError`, solvers.py:200-207), and an unknown name raises
`Exception('unknown LP solver ...')` (solvers.py:103-105).

`lpsolve_batch` is the batched sibling the reference lacks: B independent LPs
in one kernel launch, one LP per warp.
"""
import logging

import numpy as np

from polytope_b200 import engine

logger = logging.getLogger(__name__)

installed_solvers = {'b200'}
default_solver = 'b200'
# names the reference knows (solvers.py:96-101); none of them is installed here
_reference_solvers = ('glpk', 'mosek', 'scipy', 'gurobi')


def lpsolve(c, G, h, solver=None):
    """Solve min c'x s.t. Gx <= h (x free) with the given or default solver.

    @return: `dict(status=int, x=argmin or None, fun=min_value or None)` with
        status as in `scipy.optimize.linprog` (0 optimal, 1 iteration limit,
        2 infeasible, 3 unbounded, 4 numerical difficulties).
    """
    if solver is None:
        solver = default_solver
    if solver == 'b200':
        return _solve_lp_using_b200(c, G, h)
    if solver in _reference_solvers:
        _assert_have_solver(solver)
    raise Exception('unknown LP solver "{s}".'.format(s=solver))


def _solve_lp_using_b200(c, G, h):
    _assert_have_solver('b200')
    c = np.asarray(c, dtype=np.float64).reshape(-1)
    G = np.asarray(G, dtype=np.float64)
    h = np.asarray(h, dtype=np.float64).reshape(-1)
    if G.ndim != 2 or G.shape != (h.shape[0], c.shape[0]):
        raise ValueError('lpsolve: inconsistent shapes c%s G%s h%s' % (c.shape, G.shape, h.shape))
    status, X, fun, _ = engine.lp_batch(c[None], G[None], h[None])
    st = int(status[0])
    if st != 0:
        return dict(status=st, x=None, fun=None)
    return dict(status=0, x=X[0].copy(), fun=float(fun[0]))


def lpsolve_batch(C, G, H, m_rows=None):
    """B independent LPs: C[B,n], G[B,m,n], H[B,m] -> (status[B], X[B,n], fun[B]).

    Arguments may be numpy arrays (results are numpy arrays; rows of X / entries
    of fun are NaN where status != 0) or torch CUDA tensors (results stay on the
    device).  `m_rows[B]` gives ragged row counts.
    """
    status, X, fun, _ = engine.lp_batch(C, G, H, m_rows)
    return status, X, fun


def _assert_have_solver(solver):
    """Raise `RuntimeError` if `solver` is absent."""
    if solver in installed_solvers:
        return
    raise RuntimeError((
        'solver {solver} not in '
        'installed solvers: {have}').format(
            solver=solver, have=installed_solvers))

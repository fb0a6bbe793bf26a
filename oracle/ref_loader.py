"""Test infrastructure: import the UNMODIFIED reference installed by oracle/make_ref.sh.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs
may import this module; polytope_b200/ never does (tests/test_cpu_host.py pins that).

The reference binds `lpsolve` into `polytope.polytope` at import
(/root/reference/polytope/polytope.py:69), so `count_lps()` and `patched_lpsolve()` wrap /
replace THAT name, as SURVEY.md section 8(b) prescribes.
"""
import contextlib
import importlib
import os
import sys

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
REF_TESTS = os.path.join(REF_DIR, 'reference_tests')


def available():
    return os.path.isdir(os.path.join(REF_DIR, 'polytope'))


def load():
    """-> the reference's top-level `polytope` module (from oracle/_ref, never from /root/reference)."""
    if not available():
        raise ImportError('oracle/_ref/polytope is missing: run oracle/make_ref.sh in the build container')
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    mod = importlib.import_module('polytope')
    where = os.path.dirname(os.path.abspath(mod.__file__))
    if not where.startswith(REF_DIR):
        raise ImportError('`polytope` resolved to %s, not to oracle/_ref' % where)
    return mod


@contextlib.contextmanager
def count_lps():
    """Counts calls of the reference's own `lpsolve` binding; yields a one-element list."""
    pc = load()
    inner = pc.polytope.lpsolve
    n = [0]

    def counting(c, G, h, solver=None):
        n[0] += 1
        return inner(c, G, h, solver) if solver is not None else inner(c, G, h)
    pc.polytope.lpsolve = counting
    try:
        yield n
    finally:
        pc.polytope.lpsolve = inner


@contextlib.contextmanager
def patched_lpsolve(fn):
    """Routes every LP of the reference's L3 code (reduce, cheby_ball, bounding_box, ...)
    through `fn(c, G, h)` -- the drop-in point of SURVEY.md section 8(b)."""
    pc = load()
    inner = pc.polytope.lpsolve
    pc.polytope.lpsolve = fn
    try:
        yield pc
    finally:
        pc.polytope.lpsolve = inner

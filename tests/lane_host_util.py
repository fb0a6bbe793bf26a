"""Build and run tests/lane_host.cpp (the lane LP solver compiled for the CPU)."""
import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NS = 8


def build(extra=(), ns=8):
    """ns = 8: lane_solve (n <= 8); ns = 12 / 16: lane_solve_wide."""
    extra = tuple(extra) + (('-DLANE_HOST_NS=%d' % ns,) if ns != 8 else ())
    exe = os.path.join(tempfile.gettempdir(), 'pb200_lane_host_%d' % os.getuid() + ''.join(extra).replace('-', '_').replace('=', '_'))
    src = os.path.join(HERE, 'lane_host.cpp')
    csrc = os.path.join(os.path.dirname(HERE), 'polytope_b200', 'csrc')
    hdrs = [os.path.join(csrc, 'lp_lane.cuh'), os.path.join(csrc, 'lp_lane_wide.cuh')]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in hdrs]):
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-o', exe, src] + list(extra))
    return exe


def solve(lps, exe=None, waiting=False, ns=8):
    """lps: list of (c, G, h).  -> status[B], iters[B], polishes[B], fun[B], X[B, n]"""
    exe = exe or build(ns=ns)
    m = max(len(h) for _, _, h in lps)
    n = len(lps[0][0])
    assert n <= ns
    NS = ns
    with tempfile.TemporaryDirectory() as tmp:
        fin, fout = os.path.join(tmp, 'in.bin'), os.path.join(tmp, 'out.bin')
        with open(fin, 'wb') as f:
            f.write(struct.pack('iii', len(lps), m, n))
            for c, G, h in lps:
                rows = len(h)
                Gp = np.zeros((m, n))
                Gp[:rows] = G
                hp = np.zeros(m)
                hp[:rows] = h
                f.write(struct.pack('i', rows))
                f.write(np.ascontiguousarray(Gp, dtype=np.float64).tobytes())
                f.write(np.ascontiguousarray(hp, dtype=np.float64).tobytes())
                f.write(np.ascontiguousarray(c, dtype=np.float64).tobytes())
        subprocess.check_call([exe, fin, fout] + (['w'] if waiting else []))
        raw = np.fromfile(fout, dtype=np.uint8).reshape(len(lps), 16 + 8 + 8 * NS)
    meta = raw[:, :16].copy().view(np.int32)
    vals = raw[:, 16:].copy().view(np.float64)
    return meta[:, 0], meta[:, 1], meta[:, 2], vals[:, 0], vals[:, 1:1 + n]

#!/usr/bin/env python
"""Timing of the extreme() pipeline stages (reduce, cheby, dual points, dual hull,
vertices) and of plain qhull batches.  Usage: bench_extreme.py [m d n]..."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as wl                      # noqa: E402
from polytope_b200 import engine            # noqa: E402


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def run(m, d, n, cap=None):
    A, b = wl.box_cuts_batch(4, n, m, d)
    Ad, bd = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    out = {}
    for rep in range(2):
        t0 = ev()
        res = engine.reduce_batch(Ad, bd)
        keep = res.keep
        # compact kept rows to the front (host-free): build padded reduced polytopes
        bits = ((keep.unsqueeze(1) >> torch.arange(m, device='cuda')) & 1).bool()
        rows = bits.sum(1).to(torch.int32)
        order = torch.argsort((~bits).to(torch.int8), dim=1, stable=True)
        Ar = torch.gather(res.A, 1, order.unsqueeze(-1).expand(-1, -1, d)).contiguous()
        br = torch.gather(res.b, 1, order).contiguous()
        t1 = ev()
        r, xc, st = engine.cheby_batch(Ar, br, rows)
        t2 = ev()
        dual = engine.dual_points(Ar, br, xc, rows)
        t3 = ev()
        hull = engine.hull_batch(dual, rows, facet_cap=cap)
        cap = hull.facet_cap
        t4 = ev()
        V = engine.dual_facets_to_vertices(hull, xc)
        t5 = ev()
        torch.cuda.synchronize()
        out = {'m': m, 'd': d, 'n_poly': n, 'reduce_ms': t0.elapsed_time(t1), 'cheby_ms': t1.elapsed_time(t2),
               'dual_ms': t2.elapsed_time(t3), 'hull_ms': t3.elapsed_time(t4), 'vertices_ms': t4.elapsed_time(t5),
               'facet_cap': cap, 'vertices_total': int(hull.facet_cnt.sum()),
               'vertices_per_poly': float(hull.facet_cnt.float().mean()),
               'points_inserted_mean': float(hull.stats[:, 0].float().mean()),
               'facets_created_mean': float(hull.stats[:, 1].float().mean()),
               'status_bad': int((hull.status != 0).sum())}
    out['vertices_per_s'] = out['vertices_total'] / (out['hull_ms'] * 1e-3)
    return out


if __name__ == '__main__':
    args = [int(a) for a in sys.argv[1:]]
    specs = [tuple(args[i:i + 3]) for i in range(0, len(args), 3)] or [(32, 8, 256), (64, 12, 32)]
    for m, d, n in specs:
        t = time.time()
        o = run(m, d, n)
        o['wall_s'] = time.time() - t
        print(json.dumps(o), flush=True)

#!/usr/bin/env python
"""Throughput + parity sample for the BASELINE configs other than the headline
(cfg3 Region members / pairwise intersect, cfg4 d=12 m=64 reduce, cfg5 adjacency
grid).  Device-timed with CUDA events, batch resident in HBM, 3 warm-ups.
Prints one JSON object; commit the output under profiles/.

Lives under tests/ because it samples the oracle as its parity checker (only tests/, smoke() and
bench.py's CPU arm may touch oracle/); the timed region never does."""
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
from oracle import polytope_oracle as orc   # noqa: E402


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def cfg3(n=50000):
    A, b = wl.box_cuts_batch(3, n, 16, 6, shift_scale=True)
    Ad, bd = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    An, bn, _ = engine.normalize_batch(Ad, bd)
    ms_fd, (r, xc, st) = timed(lambda: engine.cheby_batch(An, bn))
    Q = orc.normalize_rows(*wl.box_cuts(3999, 16, 6, True))[:2]
    Qa = torch.from_numpy(Q[0]).cuda().expand(n, 16, 6)
    Qb = torch.from_numpy(Q[1]).cuda().expand(n, 16)
    SA = torch.cat([An, Qa], 1).contiguous()
    Sb = torch.cat([bn, Qb], 1).contiguous()
    ms_is, res = timed(lambda: engine.reduce_batch(SA, Sb, want_A=False))
    lps = int(res.n_lp.sum())
    keeps = res.keep.cpu().numpy().astype(np.uint64)
    flags = res.flags.cpu().numpy()
    bad = 0
    rng = np.random.default_rng(0)
    sample = rng.choice(n, 48, replace=False)
    for p in sample:
        o = orc.reduce(np.vstack([An[p].cpu().numpy(), Q[0]]), np.hstack([bn[p].cpu().numpy(), Q[1]]))
        mask = sum(1 << k for k in o['keep'])
        bad += int(mask != int(keeps[p])) + int(bool(flags[p] & 1) != o['empty'])
    return {'workload': 'cfg3: %d polytopes d=6 m=16 (shift/scale)' % n,
            'is_fulldim_LPs_per_s': n / (ms_fd * 1e-3), 'is_fulldim_ms': ms_fd,
            'fulldim_count': int((r > 1e-7).sum()),
            'pairwise_intersect_LPs_per_s': lps / (ms_is * 1e-3), 'pairwise_intersect_ms': ms_is,
            'pairwise_intersect_LPs': lps, 'nonempty_intersections': int((flags & 1 == 0).sum()),
            'oracle_sample': len(sample), 'oracle_mismatches': bad}


def cfg4(n=1000, m=64, d=12):
    A, b = wl.box_cuts_batch(4, n, m, d)
    Ad, bd = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    ms, res = timed(lambda: engine.reduce_batch(Ad, bd, want_A=False))
    lps = int(res.n_lp.sum())
    keeps = res.keep.cpu().numpy().astype(np.uint64)
    bad = 0
    sample = list(range(0, n, n // 12))[:12]
    for p in sample:
        o = orc.reduce(A[p], b[p])
        bad += int(sum(1 << k for k in o['keep']) != int(keeps[p])) + int(o['n_lp'] != int(res.n_lp[p]))
    return {'workload': 'cfg4 LP part: reduce() of %d polytopes d=%d m=%d' % (n, d, m),
            'LPs_per_s': lps / (ms * 1e-3), 'ms': ms, 'LPs': lps,
            'mean_iters': float(res.lp_iters.sum()) / lps, 'oracle_sample': len(sample), 'oracle_mismatches': bad}


def cfg5(shape=(32, 32)):
    A, b, idx = wl.box_grid(shape)
    n = len(A)
    cells = [orc.normalize_rows(A[i], b[i])[:2] for i in range(n)]
    An = torch.from_numpy(np.stack([c[0] for c in cells])).cuda()
    bn = torch.from_numpy(np.stack([c[1] for c in cells])).cuda()
    ii, jj = np.nonzero(~np.eye(n, dtype=bool))            # compute_adj: all ordered pairs i != j
    pi = torch.from_numpy(ii.astype(np.int32)).cuda()
    pj = torch.from_numpy(jj.astype(np.int32)).cuda()
    ms, (adj, rad, st) = timed(lambda: engine.adjacent_pairs(An, bn, pi, pj))
    touch = np.abs(idx[ii] - idx[jj]).max(1) <= 1
    adj = adj.cpu().numpy().astype(bool)
    return {'workload': 'cfg5: %s grid of unit boxes, all ordered pairs (compute_adj)' % (shape,),
            'pairs': int(len(ii)), 'LPs_per_s': len(ii) / (ms * 1e-3), 'ms': ms,
            'adjacent_pairs': int(adj.sum()), 'flag_mismatches_vs_geometry': int((adj != touch).sum()),
            'lp_status_nonzero': int((st != 0).sum())}


def cfg3_diff(n=50000):
    """cfg3 'diff' LPs: every member of the Region minus one fixed polytope Q
    (the batchable part of Region.diff)."""
    A, b = wl.box_cuts_batch(3, n, 16, 6, shift_scale=True)
    Ad, bd = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    An, bn, _ = engine.normalize_batch(Ad, bd)
    Q = orc.normalize_rows(*wl.box_cuts(3999, 16, 6, True))[:2]
    QA = torch.from_numpy(Q[0]).cuda()[None].contiguous()
    Qb = torch.from_numpy(Q[1]).cuda()[None].contiguous()
    ms, res = timed(lambda: engine.region_diff_batch(An, bn, QA, Qb, piece_cap=8 * n), reps=3, warm=1)
    lps = int(res.n_lp.sum())
    st = res.status.cpu().numpy()
    npc = res.n_pieces.cpu().numpy()
    bad = 0
    sample = np.random.default_rng(1).choice(n, 24, replace=False)
    kinds = {'pieces': 0, 'poly': 1, 'empty': 2}
    Anh, bnh = An.cpu().numpy(), bn.cpu().numpy()
    for p in sample:
        kind, pieces = orc.region_diff((Anh[p], bnh[p]), [Q])
        bad += int(kinds[kind] != int(st[p])) + int(kind == 'pieces' and len(pieces) != int(npc[p]))
    return {'workload': 'cfg3 diff: %d polytopes d=6 m=16, each minus one fixed polytope (region_diff)' % n,
            'LPs_per_s': lps / (ms * 1e-3), 'ms': ms, 'LPs': lps, 'pieces': int(npc[st == 0].sum()),
            'status_counts': {int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))},
            'oracle_sample': len(sample), 'oracle_mismatches': bad}


def cfg4_extreme(n=1000, m=64, d=12):
    """cfg4 in full: extreme() = reduce + cheby + polar dual + dual hull + vertices."""
    import polytope_b200 as pc
    A, b = wl.box_cuts_batch(4, n, m, d)
    Ad, bd = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()

    def pipeline():
        res = engine.reduce_batch(Ad, bd)
        bits = ((res.keep.unsqueeze(1) >> torch.arange(m, device='cuda')) & 1).bool()
        rows = bits.sum(1).to(torch.int32)
        order = torch.argsort((~bits).to(torch.int8), dim=1, stable=True)
        Ar = torch.gather(res.A, 1, order.unsqueeze(-1).expand(-1, -1, d)).contiguous()
        br = torch.gather(res.b, 1, order).contiguous()
        r, xc, st = engine.cheby_batch(Ar, br, rows)
        dual = engine.dual_points(Ar, br, xc, rows)
        hull = engine.hull_batch(dual, rows, facet_cap=pipeline.cap, out_cap=pipeline.pool)
        pipeline.cap, pipeline.pool = hull.facet_cap, int(hull.facet_cnt.sum()) + 1024
        V = engine.dual_facets_to_vertices(hull, xc)
        return res, hull, V
    pipeline.cap, pipeline.pool = None, None
    ms, (res, hull, V) = timed(pipeline, reps=2, warm=2)
    nv = int(hull.facet_cnt.sum())
    bad = 0
    sample = [0, n // 2] if d >= 12 else [0, n // 3, 2 * n // 3]
    from scipy.spatial import cKDTree
    for p in sample:
        ref = orc.extreme(A[p], b[p])
        o, c = int(hull.facet_off[p]), int(hull.facet_cnt[p])
        got = V[o:o + c].cpu().numpy()
        ok = got.shape == ref.shape and cKDTree(ref).query(got, p=np.inf)[0].max() < 1e-7 \
            and cKDTree(got).query(ref, p=np.inf)[0].max() < 1e-7
        bad += int(not ok)
    return {'workload': 'cfg4: extreme() of %d polytopes d=%d m=%d (reduce + cheby + dual hull + vertices)' % (n, d, m),
            'ms': ms, 'polytopes_per_s': n / (ms * 1e-3), 'vertices': nv, 'vertices_per_s': nv / (ms * 1e-3),
            'LPs': int(res.n_lp.sum()) + n, 'facets_created_mean': float(hull.stats[:, 1].float().mean()),
            'hull_status_bad': int((hull.status != 0).sum()), 'oracle_sample': len(sample), 'oracle_mismatches': bad}


def volume_bench(n=10000, m=32, d=8, N=10000):
    A, b = wl.box_cuts_batch(2, n, m, d)
    Ad, bd = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    An, bn, _ = engine.normalize_batch(Ad, bd)
    lo, hi, st = engine.bbox_batch(An, bn)
    words = np.array([engine.pcg64_state_words(np.random.default_rng(i).bit_generator) for i in range(n)], dtype=np.uint64)
    ms, cnt = timed(lambda: engine.volume_counts(An, bn, lo, hi, N, words))
    bad = 0
    for p in (0, n // 2, n - 1):
        l, u = lo[p].cpu().numpy()[:, None], hi[p].cpu().numpy()[:, None]
        x = l + np.random.default_rng(p).random((d, N)) * (u - l)
        ref = int(np.sum(np.all(An[p].cpu().numpy().dot(x) - bn[p].cpu().numpy()[:, None] < 0, 0)))
        bad += int(ref != int(cnt[p]))
    return {'workload': 'volume(): %d polytopes d=%d m=%d, %d PCG64 samples each (regenerated on device)' % (n, d, m, N),
            'ms': ms, 'samples_per_s': n * N / (ms * 1e-3), 'polytopes_per_s': n / (ms * 1e-3),
            'fp64_tflops_upper': 2.0 * n * N * m * d / (ms * 1e-3) / 1e12, 'count_mismatches_vs_numpy': bad}


def contains_bench(m=16, d=8, N=1 << 25):
    """Region.contains-style stream: N points (d x N column vectors, > L2) against one polytope."""
    A, b = wl.box_cuts(2000, m, d)
    An, bn, _ = orc.normalize_rows(A, b)
    Ad, bd = torch.from_numpy(An).cuda()[None].contiguous(), torch.from_numpy(bn).cuda()[None].contiguous()
    pts = (torch.rand((d, N), dtype=torch.float64, device='cuda') * 3.2 - 1.6).contiguous()
    ms, out = timed(lambda: engine.contains_batch(Ad, bd, pts))
    sub = pts[:, :100000].cpu().numpy()
    ref = orc.contains(An, bn, sub)
    bytes_ = N * (8 * d + 1)
    return {'workload': 'contains(): one polytope d=%d m=%d against %d points (%.2f GB streamed once)' % (d, m, N, bytes_ / 1e9),
            'ms': ms, 'points_per_s': N / (ms * 1e-3), 'algorithmic_bytes': bytes_, 'achieved_GBps': bytes_ / (ms * 1e-3) / 1e9,
            'hbm_peak_GBps': PEAK_HBM, 'frac_of_measured_hbm_peak': bytes_ / (ms * 1e-3) / 1e9 / PEAK_HBM,
            'flag_mismatches_vs_oracle': int((out[0, :100000].cpu().numpy() != ref).sum())}


def cfg5_partition(shape=(32, 32), block=4, ref_regions=6):
    """The is_adjacent callers on a box grid cut into block x block regions: find_adjacent_regions
    (one launch over all n(n-1)/2 cell pairs, OR-ed per region pair) and separate() of the union of
    every other region (disconnected by construction).  Wall-clock, host bookkeeping included; the
    oracle's sequential loops are timed on the first `ref_regions` regions only."""
    import polytope_b200 as pb
    A, b, idx = wl.box_grid(shape)
    cells = [pb.Polytope(A[i], b[i]) for i in range(len(A))]
    label = (idx[:, 0] // block) * ((shape[1] + block - 1) // block) + idx[:, 1] // block
    regions = [pb.Region([cells[i] for i in np.nonzero(label == g)[0]]) for g in range(int(label.max()) + 1)]
    n_cells = len(cells)

    def wall(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3, out

    ms_adj, adj = wall(lambda: pb.find_adjacent_regions(regions))
    every_other = pb.Region([p for g, reg in enumerate(regions) if g % 2 == 0 for p in reg.list_poly])
    ms_sep, parts = wall(lambda: pb.separate(every_other))
    # oracle: the reference's loop of is_adjacent(region_i, region_j), j < i, with its early exits
    sub = regions[:ref_regions]
    t0 = time.perf_counter()
    ref = np.eye(len(sub), dtype=np.int8)
    n_ref_lp = 0
    for i in range(len(sub)):
        for j in range(i):
            hit = False
            for p in sub[i]:
                for q in sub[j]:
                    n_ref_lp += 1
                    if orc.is_adjacent(p.A, p.b, q.A, q.b):
                        hit = True
                        break
                if hit:
                    break
            ref[i, j] = ref[j, i] = hit
    ref_s = time.perf_counter() - t0
    return {'workload': 'find_adjacent_regions / separate on a %s box grid in %dx%d regions (%d regions, %d cells)'
                        % ('x'.join(map(str, shape)), block, block, len(regions), n_cells),
            'find_adjacent_regions_ms': ms_adj, 'pair_LPs': n_cells * (n_cells - 1) // 2,
            'pair_LPs_per_s_wall': n_cells * (n_cells - 1) // 2 / (ms_adj * 1e-3),
            'separate_ms': ms_sep, 'separate_cells': len(every_other), 'separate_parts': len(parts),
            'oracle_regions': len(sub), 'oracle_LPs': n_ref_lp, 'oracle_s': ref_s,
            'oracle_mismatches': int((adj[:len(sub), :len(sub)] != ref).sum())}


try:
    PEAK_HBM = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except (OSError, ValueError, KeyError):
    PEAK_HBM = 6650.0


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'next':
        t0 = time.time()
        out = {'cfg3_diff': cfg3_diff(), 'cfg4_extreme_d8': cfg4_extreme(1000, 32, 8), 'cfg4_extreme': cfg4_extreme(),
               'volume': volume_bench(), 'contains_m16': contains_bench(16, 8), 'contains_m32': contains_bench(32, 8),
               'contains_d3': contains_bench(6, 3, 1 << 26)}
        out['wall_s'] = time.time() - t0
        print(json.dumps(out, indent=1))
        sys.exit(0)
    t0 = time.time()
    out = {'cfg3': cfg3(), 'cfg4': cfg4(), 'd16': cfg4(500, 64, 16), 'cfg5': cfg5(),
           'cfg5_4d': cfg5((6, 6, 6, 6)), 'cfg5_partition': cfg5_partition()}
    out['wall_s'] = time.time() - t0
    print(json.dumps(out, indent=1))

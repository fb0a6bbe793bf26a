#!/usr/bin/env python
"""Headline benchmark: LPs/s on batched reduce() of random H-polytopes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step is one pass of the hot path over one batch: `reduce(Polytope(A, b))` for
BASELINE.json configs[1] -- 10 000 random H-polytopes, d = 8, m = 32 (generator
"box+cuts") -- per GPU.  An "LP" is one lpsolve-equivalent solve
the reference algorithm needs on that input (counted by the kernels exactly as
the reference's own lpsolve call counter counts them; tests pin the equality).

Printed JSON (one line, rank 0):
  value     whole-job LPs/s with the batch resident in HBM (device-timed)
  e2e       same metric through the public host-buffer API: pinned host (A, b)
            -> H2D -> pipeline -> (N > 1: one NCCL all-gather on the device) -> one D2H
            of masks/flags/counts, every step
  roofline  the dominant kernel (the row LPs of reduce): algorithmic bytes / its mean
            device time (CUDA events on the launch stream inside the library) against
            the measured HBM peak, plus the fp64 view that actually bounds it (peak
            measured in this run by the library's DFMA kernel) and `ab_read`: the
            batched (A|b) read in isolation (constructor normalisation, bulk async copies)
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, installed by oracle/make_ref.sh;
            scipy/HiGHS path) on a bounded sample of the same workload, all host cores
            (N = 1 only); kind "port" = the oracle restatement when oracle/_ref is absent
  config.strong  strong scaling of the two BASELINE configs that ask for it: cfg5
            (1 047 552 ordered-pair is_adjacent LPs, compute_adj semantics) and cfg4
            (extreme() of 1 000 polytopes d = 12, m = 64), the units split over the N
            ranks with their one all-gather, device-timed, max over ranks

N > 1: launched by torchrun, one rank per GPU, weak scaling (each rank reduces
its own 10 000 polytopes), one NCCL all-gather of the 64-bit keep masks per
step; time = max over ranks.

--impl reference times the reference's own CPU implementation of the path
(`polytope.reduce(polytope.Polytope(A, b))` of oracle/_ref with its lpsolve calls
counted) on a bounded sample per step, with every host core.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import workloads as wl  # noqa: E402

METRIC = 'LPs/sec on batched reduce() of random H-polytopes'
CFG = dict(cfg=2, n_poly=10000, m=32, d=8)
WORKLOAD = 'cfg2: reduce() of 10000 box+cuts H-polytopes, d=8, m=32, per GPU'


def algorithmic_bytes(n_poly, m, d):
    """in 8*m*(d+1) B per polytope (A|b read once), out 8 B mask + m B status."""
    return n_poly * (8 * m * (d + 1) + 8 + m)


def ipm_flops_per_iteration(m, n):
    """2*(m n^2/2 + n^3/6 + 6 m n + 2 n^2) flop per interior-point iteration."""
    return 2.0 * (m * n * n / 2.0 + n ** 3 / 6.0 + 6.0 * m * n + 2.0 * n * n)


# --------------------------------------------------------------------------
# CPU arm: the unmodified reference (or, without oracle/_ref, the oracle port), all host cores
# --------------------------------------------------------------------------
def reference_kind():
    from oracle import ref_loader
    return 'reference' if ref_loader.available() else 'port'


def _cpu_chunk(args):
    first, count, kind = args
    if kind == 'reference':
        from oracle import ref_loader
        pc = ref_loader.load()
        with ref_loader.count_lps() as n:
            for i in range(first, first + count):
                A, b = wl.box_cuts(1000 * CFG['cfg'] + i, CFG['m'], CFG['d'])
                pc.reduce(pc.Polytope(A, b))
        return n[0]
    from oracle import polytope_oracle as orc
    n = 0
    for i in range(first, first + count):
        A, b = wl.box_cuts(1000 * CFG['cfg'] + i, CFG['m'], CFG['d'])
        n += orc.reduce(A, b)['n_lp']
    return n


def cpu_reduce_sample(n_poly, cores, first=0, kind='reference'):
    """reduce() on `n_poly` polytopes of the workload on the host; -> (LPs, seconds)."""
    import logging
    import multiprocessing as mp
    logging.disable(logging.WARNING)                     # the reference warns about cvxopt at import
    per = max(1, n_poly // (cores * 4))
    chunks = [(first + s, min(per, n_poly - s), kind) for s in range(0, n_poly, per)]
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_chunk, [(0, 1, kind)] * cores)     # warm the workers (imports)
        t0 = time.perf_counter()
        lps = sum(pool.map(_cpu_chunk, chunks))
        dt = time.perf_counter() - t0
    logging.disable(logging.NOTSET)
    return lps, dt


def _baseline_text(kind, sample, lps, secs=None):
    what = ('the unmodified reference (oracle/_ref: polytope.reduce(polytope.Polytope(A, b)), its lpsolve calls '
            'counted)' if kind == 'reference' else 'oracle port of the reference reduce()')
    return '%d polytopes of cfg2 (%d LPs%s): %s over scipy.optimize.linprog/HiGHS, multiprocessing over all host ' \
           'cores' % (sample, lps, '' if secs is None else ', %.1f s wall' % secs, what)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    kind = reference_kind()
    sample = 16 * cores                                  # ~25 CPU-seconds per step
    for _ in range(args.warmup):
        cpu_reduce_sample(cores, cores, kind=kind)
    lps, secs = 0, 0.0
    for k in range(args.steps):
        a, b = cpu_reduce_sample(sample, cores, first=k * sample, kind=kind)
        lps += a
        secs += b
    value = lps / secs
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'LPs/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * secs / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': {'workload': WORKLOAD, 'sample_polytopes_per_step': sample},
        'cpu_baseline': {'value': value, 'unit': 'LPs/s', 'cores': cores, 'kind': kind,
                         'sample': _baseline_text(kind, sample, lps) + ', per step'},
        'e2e': {'value': value, 'unit': 'LPs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------
class ClockSampler(object):
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                 '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            f = [x.strip() for x in r.split(',')]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def measured_traffic():
    """ncu DRAM bytes per launch of the dominant kernel, only if the capture belongs to the
    sources the loaded library was built from (tools/update_traffic.py records the digest)."""
    path = os.path.join(ROOT, 'profiles', 'row_lp_dram_bytes.json')
    stamp = os.path.join(ROOT, 'polytope_b200', 'libpolytope_b200.so.srchash')
    try:
        doc = json.load(open(path))
        built = open(stamp).read().strip()
    except (OSError, ValueError):
        return None, 'no ncu capture recorded for this build (profiles/row_lp_dram_bytes.json or the .srchash is missing)'
    if doc.get('srchash') != built:
        return None, 'stale: profiles/row_lp_dram_bytes.json was captured on source digest %s, the library is %s' % (
            str(doc.get('srchash'))[:12], built[:12])
    return doc.get('dram_bytes_per_launch'), doc.get('source')


# --------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------
def run_gpu(args):
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    P, m, d = CFG['n_poly'], CFG['m'], CFG['d']

    # CPU baseline first (fork-based pool must not run after CUDA is initialised)
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        kind = reference_kind()
        sample = 16 * cores
        lps, secs = cpu_reduce_sample(sample, cores, kind=kind)
        cpu_baseline = {'value': lps / secs, 'unit': 'LPs/s', 'cores': cores, 'kind': kind,
                        'sample': _baseline_text(kind, sample, lps, secs)}

    import torch
    import torch.distributed as dist
    from polytope_b200 import engine, sharding
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except (OSError, ValueError):
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback 6.65 TB/s (B200_PROFILING.md)'

    # weak scaling: rank r owns polytopes [r*P, (r+1)*P) of the seed sequence
    A_h, b_h = wl.box_cuts_batch(CFG['cfg'], P, m, d, first=rank * P)
    A_pin = torch.from_numpy(A_h).pin_memory()
    b_pin = torch.from_numpy(b_h).pin_memory()
    A_dev = A_pin.to('cuda')
    b_dev = b_pin.to('cuda')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        res = engine.reduce_batch(A_dev, b_dev, want_A=False)
        keep = res.keep
        if world > 1:
            keep = sharding.allgather_blocks(res.keep, world * P)
        return res, keep

    def step_e2e():
        # the public host-buffer call: pinned (A, b) in, masks / flags / counts out.  N = 1: the H2D
        # and D2H copies happen inside (chunked over two streams, overlapping the kernels), results
        # are numpy arrays.  N > 1: the same chunked call with the results left on the device for the one
        # all-gather, then one D2H.
        if world == 1:
            res = engine.reduce_batch(A_pin, b_pin, want_A=False, want_b=False)
            return (torch.from_numpy(res.keep), torch.from_numpy(res.flags), torch.from_numpy(res.n_lp),
                    torch.from_numpy(res.lp_iters))
        res = engine.reduce_batch(A_pin, b_pin, want_A=False, want_b=False, results_on_device=True)
        packed = torch.stack([res.keep, res.flags.to(torch.int64), res.n_lp.to(torch.int64),
                              res.lp_iters.to(torch.int64)], 1)
        allp = sharding.allgather_blocks(packed, world * P).cpu()
        return allp[:, 0], allp[:, 1], allp[:, 2], allp[:, 3]

    def timed(step, steps, profile=False):
        """K steps, each bracketed by its own CUDA events on the launch stream,
        L2 flushed (untimed) before every step; returns (total ms, last result, stage ms)."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        stages = {}
        out = None
        barrier()
        for k in range(steps):
            flush.zero_()
            if world > 1:
                dist.barrier()
            ev[k][0].record()
            out = step()
            ev[k][1].record()
            if profile:
                for name, v in engine.profile_read().items():
                    stages[name] = stages.get(name, 0.0) + v / steps
        barrier()
        total = sum(a.elapsed_time(b) for a, b in ev)
        return total, out, stages

    for _ in range(max(args.warmup, 3)):
        step_device()
        step_e2e()
    barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = engine.launch_count()
    t_wall0 = time.perf_counter()
    engine.profile_enable(True)
    ms_dev, (res, _), stages = timed(step_device, args.steps, profile=True)
    engine.profile_enable(False)
    launches = engine.launch_count() - launches0
    ms_e2e, e2e_out, _ = timed(step_e2e, args.steps)
    t_wall1 = time.perf_counter()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    lps_step = int(res.n_lp.sum().item())
    iters_step = int(res.lp_iters.sum().item())
    t = torch.tensor([ms_dev, ms_e2e, float(lps_step)], dtype=torch.float64, device='cuda')
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_dev, ms_e2e, lps_all = tmax[0].item(), tmax[1].item(), tsum[2].item()
    else:
        lps_all = float(lps_step)

    strong = None if args.no_strong else strong_scaling(torch, dist, engine, sharding, rank, world, barrier)
    dfma_peak = engine.measure_dfma_tflops() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = lps_all * args.steps / (ms_dev * 1e-3)
    e2e_value = lps_all * args.steps / (ms_e2e * 1e-3)
    row_ms = stages.get('row_lp', float('nan'))
    alg_bytes = algorithmic_bytes(P, m, d)
    achieved = alg_bytes / (row_ms * 1e-3) / 1e9
    # fp64 view: algorithmic flops of the row LPs (SURVEY 8d formula x measured iterations)
    kept_rows = int(res.n_lp.sum().item()) - P * (1 + 2 * d)
    row_iters = iters_step * kept_rows / max(lps_step, 1)
    flops = row_iters * ipm_flops_per_iteration(int(round(kept_rows / P)), d)
    h2d = A_pin.numel() * 8 + b_pin.numel() * 8
    d2h = sum(x.numel() * x.element_size() for x in e2e_out)          # what one rank copies back
    traffic, traffic_src = measured_traffic()
    line = {
        'metric': METRIC, 'value': value, 'unit': 'LPs/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_dev / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'n_poly_per_gpu': P, 'm': m, 'd': d, 'lps_per_step_per_gpu': lps_step,
                   'polytopes_per_s': world * P * args.steps / (ms_dev * 1e-3),
                   'mean_ipm_iterations_per_lp': iters_step / max(lps_step, 1),
                   'l2': 'flushed before every step (256 MiB memset, untimed); each step timed with its own '
                         'CUDA event pair on the launch stream',
                   'collective': 'none' if world == 1 else 'one NCCL all_gather of int64 keep masks per step',
                   'strong': strong},
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': 'LPs/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': launches,
        'roofline': {'bound': 'hbm', 'kernel': 'lane_kernel<RowLanes> (reduce row LPs, one LP per lane)', 'achieved': achieved,
                     'peak': hbm_peak, 'unit': 'GB/s', 'frac': achieved / hbm_peak, 'traffic': traffic,
                     'traffic_source': traffic_src,
                     'peak_source': peak_src, 'kernel_ms': row_ms, 'algorithmic_bytes': alg_bytes,
                     'kernel_share_of_step': row_ms / (ms_dev / args.steps),
                     'stage_ms': stages,
                     'fp64': {'achieved_tflops': flops / (row_ms * 1e-3) / 1e12, 'peak_tflops': dfma_peak,
                              'peak_source': 'measured in this run: pb200_measure_dfma_tflops (DFMA chains, 32 warps/SM, CUDA events)',
                              'frac': flops / (row_ms * 1e-3) / 1e12 / dfma_peak,
                              'note': 'the path is fp64 latency/issue bound, not HBM bound (SURVEY.md 8d)'}},
        'cpu_baseline': cpu_baseline,
    }
    line['roofline']['ab_read'] = ab_read_roofline(P, m, d, stages.get('normalize'), hbm_peak)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def strong_scaling(torch, dist, engine, sharding, rank, world, barrier, reps=5):
    """The two BASELINE configs quoted as sharded over the GPUs of one box, total work fixed:
    cfg5 (prop2partition adjacency grid, all ordered pairs as MetricPartition.compute_adj walks
    them, prop2partition.py:253-261) and cfg4 (extreme() of 1 000 polytopes d = 12, m = 64).
    Each rank does its contiguous block of pairs / polytopes; one all-gather of the flags, and of
    the vertex counts + (ragged) vertices.  Device-timed with CUDA events (median of the repetitions), max over ranks."""
    out = {}

    def run(fn, reps):
        # every repetition has its own event pair and the median is reported: these calls are a few milliseconds
        # long, and one host-side hiccup (allocator growth, garbage collection) inside a single bracket around all
        # repetitions moved the figure by an order of magnitude between runs
        fn()
        fn()
        times = []
        for _ in range(reps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = fn()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = torch.tensor([sorted(times)[len(times) // 2]], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), res

    # cfg5
    A, b, idx = wl.box_grid((32, 32))
    n = len(A)
    Ad, bd = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()     # unit rows: already constructor-normalised
    ms, flags = run(lambda: sharding.adjacency_ordered_sharded(Ad, bd), reps)
    pairs = n * (n - 1)
    ii, jj = np.nonzero(~np.eye(n, dtype=bool))
    touch = np.abs(idx[ii] - idx[jj]).max(1) <= 1
    out['cfg5'] = {'workload': '32x32 grid of unit boxes, all %d ordered pairs (compute_adj), one is_adjacent LP (8x3) each, '
                               'pair range split over %d GPU(s), one all-gather of the flags' % (pairs, world),
                   'value': pairs / (ms * 1e-3), 'unit': 'LPs/s', 'ms': ms, 'scaling': 'strong',
                   'flag_mismatches_vs_geometry': int((flags.cpu().numpy().astype(bool) != touch).sum())}
    # cfg4
    A4, b4 = wl.box_cuts_batch(4, 1000, 64, 12)
    A4d, b4d = torch.from_numpy(A4).cuda(), torch.from_numpy(b4).cuda()
    state = {'caps': None}

    def cfg4():
        counts, V, caps = sharding.extreme_tensor_sharded(A4d, b4d, caps=state['caps'])
        state['caps'] = caps
        return counts, V
    ms, (counts, V) = run(cfg4, 3)
    out['cfg4'] = {'workload': 'extreme() of 1000 box+cuts polytopes d=12 m=64 (reduce + cheby + polar dual + dual hull + '
                               'vertices), polytopes split over %d GPU(s), all-gather of counts and vertices' % world,
                   'value': 1000 / (ms * 1e-3), 'unit': 'polytopes/s', 'ms': ms, 'scaling': 'strong',
                   'vertices': int(counts.sum().item()), 'vertices_per_s': float(counts.sum().item()) / (ms * 1e-3)}
    return out


def ab_read_roofline(P, m, d, stage_ms, hbm_peak, scale=50, reps=6):
    """The batched (A|b) read in isolation (north_star's HBM target applies to this phase):
    the constructor-normalisation kernel reads A, b once and writes An, bn, valid once.  Reported at
    the bench batch (stage time of the timed steps, launch-latency sized: 46 MB) and, outside the
    timed region, on a batch `scale` times larger that streams from HBM (inputs >> L2)."""
    import torch
    from polytope_b200 import _capi
    per_poly = 2 * 8 * m * (d + 1) + 8
    out = {'kernel': 'normalize_tile_kernel<bulk copies> (Polytope.__init__ normalisation of the stacked batch)',
           'algorithmic_bytes_per_polytope': per_poly, 'peak': hbm_peak, 'unit': 'GB/s'}
    if stage_ms:
        out['bench_batch'] = {'n_poly': P, 'ms': stage_ms, 'achieved': per_poly * P / stage_ms / 1e6,
                              'frac': per_poly * P / stage_ms / 1e6 / hbm_peak}
    Pb = P * scale
    lib = _capi.lib()
    A = torch.randn(Pb, m, d, dtype=torch.float64, device='cuda')
    b = torch.randn(Pb, m, dtype=torch.float64, device='cuda')
    An, bn = torch.empty_like(A), torch.empty_like(b)
    valid = torch.empty(Pb, dtype=torch.int64, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    ms = []
    for k in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.pb200_normalize_batch(A.data_ptr(), b.data_ptr(), None, Pb, m, d, An.data_ptr(), bn.data_ptr(),
                                       valid.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        if rc != 0:
            raise RuntimeError('pb200_normalize_batch failed: %d' % rc)
        if k >= 2:
            ms.append(e0.elapsed_time(e1))
    mean = sum(ms) / len(ms)
    out['streaming_batch'] = {'n_poly': Pb, 'bytes': per_poly * Pb, 'ms': mean, 'achieved': per_poly * Pb / mean / 1e6,
                              'frac': per_poly * Pb / mean / 1e6 / hbm_peak,
                              'note': 'untimed extra after the steps; inputs 1.15 GB >> L2, CUDA events per launch'}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-strong', action='store_true', help='skip the cfg4 / cfg5 strong-scaling extras')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            # convenience: relaunch under torchrun on this node
            cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(args.gpus),
                   '--master-addr', '127.0.0.1', '--master-port', '29533', os.path.abspath(__file__)] + sys.argv[1:]
            return subprocess.call(cmd)
        raise SystemExit('--gpus %d but WORLD_SIZE=%d' % (args.gpus, world))
    return run_gpu(args)


if __name__ == '__main__':
    sys.exit(main())

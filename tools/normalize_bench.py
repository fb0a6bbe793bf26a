"""A/B of the constructor-normalisation kernels (the batched (A|b) read of the reduce pipeline).

    python tools/normalize_bench.py [out.json]

For every variant of pb200_normalize_variant (-1 row-per-lane without staging, 0 tiled + 64-bit
loads, 1 tiled + 128-bit loads, 2 tiled + bulk async copies / TMA) a child process checks that the
outputs equal the row-per-lane kernel's bit for bit (that kernel is pinned to numpy by
tests/test_gpu_polytope.py) on aligned, odd, ragged and partial-tile shapes, then times the kernel
on cfg2's batch (10 000 x 32 x 8, L2 flushed before every launch) and on a batch far larger than
L2 (500 000 x 32 x 8 = 1.15 GB each way), and reports algorithmic bytes / time against the
measured HBM peak of MEASURED_PEAKS.json.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = [(37, 20, 1), (37, 20, 3), (37, 20, 8), (37, 20, 13), (37, 20, 31), (37, 33, 5), (5, 1, 4),
          (10, 64, 128), (9, 64, 16), (1001, 32, 8), (100, 64, 12), (64, 16, 6)]


def child(variants):
    for v in variants:
        child_one(v)


def child_one(variant):
    import numpy as np
    import torch
    from polytope_b200 import _capi, engine
    lib = _capi.lib()
    rng = np.random.default_rng(5)
    out = {'variant': variant, 'mismatches': 0, 'cases': 0}
    for (P, m, d) in SHAPES:
        A = rng.standard_normal((P, m, d)) * 10 ** rng.uniform(-3, 3, (P, m, 1))
        b = rng.standard_normal((P, m))
        A[P // 2, m // 2] = 0.0
        for mr in (None, rng.integers(0, m + 1, P).astype(np.int32)):
            lib.pb200_normalize_variant(-1)
            want = engine.normalize_batch(A, b, mr)
            lib.pb200_normalize_variant(variant)
            got = engine.normalize_batch(A, b, mr)
            out['cases'] += 1
            for w_, g_ in zip(want, got):
                if not np.array_equal(np.asarray(w_).view(np.uint64), np.asarray(g_).view(np.uint64)):
                    out['mismatches'] += 1
            if mr is None:   # numpy itself, batched form of polytope.py:129-138
                nrm = np.sqrt(np.sum(A * A, axis=2))
                pos = nrm > 1e-10
                mult = np.where(pos, 1.0 / np.where(pos, nrm, 1.0), 0.0)
                ok = np.array_equal(np.asarray(got[0])[pos], (A * mult[..., None])[pos]) and \
                    np.array_equal(np.asarray(got[1])[pos], (b * mult)[pos])
                out['mismatches'] += 0 if ok else 1
    peak = 6550.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for name, P, reps, do_flush in (('cfg2_10k', 10000, 15, True), ('large_500k', 500000, 8, False)):
        m, d = 32, 8
        A = torch.randn(P, m, d, dtype=torch.float64, device='cuda')
        b = torch.randn(P, m, dtype=torch.float64, device='cuda')
        An, bn = torch.empty_like(A), torch.empty_like(b)
        valid = torch.empty(P, dtype=torch.int64, device='cuda')
        st = torch.cuda.current_stream().cuda_stream
        ms = []
        for k in range(reps + 3):
            if do_flush:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.pb200_normalize_batch(A.data_ptr(), b.data_ptr(), None, P, m, d, An.data_ptr(), bn.data_ptr(),
                                           valid.data_ptr(), st)
            e1.record()
            torch.cuda.synchronize()
            assert rc == 0
            if k >= 3:
                ms.append(e0.elapsed_time(e1))
        ms.sort()
        med = ms[len(ms) // 2]
        nbytes = P * (2 * 8 * m * (d + 1) + 8)       # A, b in; An, bn, valid out
        out[name] = {'ms_median': med, 'ms_min': ms[0], 'algorithmic_bytes': nbytes, 'GBps': nbytes / med / 1e6,
                     'frac_of_measured_hbm_peak': nbytes / med / 1e6 / peak, 'peak_GBps': peak}
        del A, b, An, bn
    print('RESULT ' + json.dumps(out), flush=True)


def profile(variant, P=500000, m=32, d=8):
    """Two launches on the large batch (run under ncu with --launch-skip 1 --launch-count 1)."""
    import torch
    from polytope_b200 import _capi
    lib = _capi.lib()
    lib.pb200_normalize_variant(variant)
    A = torch.randn(P, m, d, dtype=torch.float64, device='cuda')
    b = torch.randn(P, m, dtype=torch.float64, device='cuda')
    An, bn = torch.empty_like(A), torch.empty_like(b)
    valid = torch.empty(P, dtype=torch.int64, device='cuda')
    for _ in range(2):
        assert lib.pb200_normalize_batch(A.data_ptr(), b.data_ptr(), None, P, m, d, An.data_ptr(), bn.data_ptr(),
                                         valid.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()


def main():
    results = []
    # the bulk-copy variant gets its own process: a device trap there must not take the others down
    for vs in ('-1,0,1', '2'):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--child', vs], capture_output=True,
                               text=True, timeout=60)
            lines = [l for l in r.stdout.splitlines() if l.startswith('RESULT ')]
            results += [json.loads(l[7:]) for l in lines]
            if len(lines) < len(vs.split(',')) or r.returncode:
                results.append({'variants': vs, 'error': (r.stderr or r.stdout)[-600:], 'rc': r.returncode})
        except subprocess.TimeoutExpired:
            results.append({'variants': vs, 'error': 'timeout'})
    txt = json.dumps(results, indent=1)
    print(txt)
    if len(sys.argv) > 1:
        open(sys.argv[1], 'w').write(txt + '\n')


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == '--child':
        child([int(v) for v in sys.argv[2].split(',')])
    elif len(sys.argv) > 2 and sys.argv[1] == '--profile':
        profile(int(sys.argv[2]))
    else:
        main()

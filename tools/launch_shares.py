#!/usr/bin/env python
"""Per-kernel time shares from an ncu launch list (the --metrics gpu__time_duration.sum --csv pass).
Usage: tools/launch_shares.py launches.csv > profiles/rNN_launch_shares.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[ki] == 'Kernel Name':
        continue
    scale = {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'msecond': 1.0, 'ms': 1.0, 'nsecond': 1e-6, 'second': 1e3}[r[ui]]
    t = float(r[vi].replace(',', '')) * scale
    n, s = agg.get(r[ki], (0, 0.0))
    agg[r[ki]] = (n + 1, s + t)
own = {k: v for k, v in agg.items() if 'pb200' in k or 'lane_kernel' in k or 'lane_own' in k}
import re
# kernels of bench.py's untimed extras (cfg4 extreme(), cfg5 pair flags, the DFMA peak measurement), not of the cfg2 step
extras = re.compile(r'hull_kernel|lp_kernel<2|AdjacentOwn|dual_|dfma_peak|sweep_kernel|<1[026],|lane_kernel<1[26]')
step = {k: v for k, v in own.items() if not extras.search(k)}
tot = sum(s for _, s in step.values())
print('# launch shares from %s (ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are' % sys.argv[1])
print('# cold-cache and serialised).  Shares are of the kernels of the cfg2 step; the extras of bench.py (cfg4 / cfg5 strong-scaling passes, peak measurement, 500k-polytope ab_read pass) and the fill / random kernels are listed without a share.')
for k, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    share = '%5.1f%%' % (100 * s / tot) if k in step else '   -  '
    print('%-110s n=%3d %10.3f ms %s' % (k[:110], n, s, share))

#!/usr/bin/env python
"""Regenerate profiles/row_lp_dram_bytes.json (bench.py's `roofline.traffic`) from an
`ncu --set full` capture of the dominant kernel, tagged with the source digest of the library
that was profiled so that bench.py can refuse a stale figure.

  on the GPU box:   ncu --set full --clock-control none --import-source on -k regex:lane_kernel \
                        -s 2 -c 2 -o gpurun_out/rowlp python tools/profile_reduce.py 10000 32 8 2
  here:             python tools/update_traffic.py gpurun_out/rowlp.ncu-rep
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
pattern = sys.argv[2] if len(sys.argv) > 2 else 'RowLanes'
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
hit = [dict(zip(hdr, r)) for r in rows[2:] if pattern in dict(zip(hdr, r)).get('Kernel Name', '')]
if not hit:
    raise SystemExit('no kernel matching %r in %s' % (pattern, rep))
unit = dict(zip(hdr, units))


def to_bytes(name, row):
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit[name]]
    return float(row[name]) * scale


k = hit[0]
rd, wr = to_bytes('dram__bytes_read.sum', k), to_bytes('dram__bytes_write.sum', k)
stamp = os.path.join(ROOT, 'polytope_b200', 'libpolytope_b200.so.srchash')
doc = {'kernel': k['Kernel Name'], 'dram_bytes_per_launch': int(rd + wr), 'dram_bytes_read': int(rd),
       'dram_bytes_write': int(wr), 'gpu_time_ms': float(k['gpu__time_duration.sum']) * {'us': 1e-3, 'ms': 1.0, 'ns': 1e-6}[unit['gpu__time_duration.sum']],
       'srchash': open(stamp).read().strip(),
       'source': '%s: dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full, cfg2 (10000 polytopes); '
                 'algorithmic bytes 23.44 MB' % os.path.basename(rep)}
json.dump(doc, open(os.path.join(ROOT, 'profiles', 'row_lp_dram_bytes.json'), 'w'), indent=1)
print(json.dumps(doc, indent=1))

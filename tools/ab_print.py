#!/usr/bin/env python
"""Pretty-print the log of tools/ab_stage_bench.py runs."""
import json
import sys
for line in open(sys.argv[1]):
    if line.startswith('=='):
        print(line.strip())
        continue
    name, js = line.split(' ', 1)
    try:
        d = json.loads(js)
    except Exception:
        print(line[:200])
        continue
    print('%-24s ms %.3f MLPs %.2f it %.2f bbox %.3f row %.3f cheby %.3f keepsum %d' % (
        name.split('/')[-1], d['ms'], d['MLPs'], d['iters_per_lp'], d['stages']['bbox_lp'], d['stages']['row_lp'],
        d['stages']['cheby_lp'], d['keepsum'] % 100000))

#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key raw metrics, stall mix and
the hottest source lines.  Usage: tools/ncu_summary.py rep.ncu-rep [n_units] [kernel-name substring] > profiles/x.txt
(with a substring the raw metrics are those of the first launch whose name contains it; the source page is
then restricted to that launch)"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 and float(sys.argv[2]) > 0 else None
pattern = sys.argv[3] if len(sys.argv) > 3 else None


def ncu(*args):
    return subprocess.run(['ncu', '-i', rep] + list(args), capture_output=True, text=True).stdout


rows = list(csv.reader(io.StringIO(ncu('--page', 'raw', '--csv'))))
hdr, unit = rows[0], rows[1]
vals, launch = rows[2], 0
if pattern:
    launch = [k for k, r in enumerate(rows[2:]) if pattern in dict(zip(hdr, r)).get('Kernel Name', '')][0]
    vals = rows[2 + launch]
raw = dict(zip(hdr, zip(unit, vals)))
print('# ncu summary of', rep)
print('kernel:', raw.get('Kernel Name', ('', '?'))[1], ' grid', raw.get('Grid Size', ('', '?'))[1],
      ' block', raw.get('Block Size', ('', '?'))[1])
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'lts__t_sector_hit_rate.pct']
for k in keys:
    if k in raw:
        print('%-78s %12s %s' % (k, raw[k][1], raw[k][0]))
if units and 'smsp__inst_executed.sum' in raw:
    print('warp instructions per unit: %.0f' % (float(raw['smsp__inst_executed.sum'][1]) / units))
print('\n# warp stall reasons (cycles per issued instruction)')
for k in sorted(raw):
    if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio'):
        v = float(raw[k][1])
        if v >= 0.02:
            print('  %-28s %.3f' % (k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], v))

sel = ['--launch-skip', str(launch), '--launch-count', '1']
src = list(csv.reader(io.StringIO(ncu('--page', 'source', '--print-source', 'cuda,sass', '--csv', *sel))))
hdr = None
cur = None
lines = []
for r in src:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or not r or r[0] == '':
        continue
    try:
        lines.append((int(r[4]), int(r[hdr.index('Instructions Executed')]), cur, r[0], r[1].strip()[:90]))
    except (ValueError, IndexError):
        pass
tot = sum(l[0] for l in lines) or 1
toti = sum(l[1] for l in lines) or 1
print('\n# hottest source lines (stall samples %, executed instructions %)')
for s, i, f, ln, text in sorted(lines, reverse=True)[:30]:
    print('%5.1f%% %5.1f%%  %s:%s  %s' % (100.0 * s / tot, 100.0 * i / toti, f, ln, text))

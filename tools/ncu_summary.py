#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) per kernel: duration, DRAM traffic, throughput %, occupancy, stall reasons."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
def col(n): return h.index(n) if n in h else None
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active']
seen = collections.Counter()
for r in rows[2:]:
    name = r[col('Kernel Name')].split('(')[0].replace('void ', '').replace('kasa::<unnamed>::', '').replace('unnamed>::', '')
    seen[name] += 1
    if seen[name] > int(sys.argv[2]) if len(sys.argv) > 2 else seen[name] > 1:
        continue
    print('==== %s  (launch #%d)' % (name, seen[name]))
    for k in keys:
        c = col(k)
        if c is not None:
            print('   %-72s %s %s' % (k, r[c], units[c]))
    stalls = []
    for i, n in enumerate(h):
        if n.startswith('smsp__average_warps_issue_stalled_') and n.endswith('_per_issue_active.ratio'):
            try: stalls.append((float(r[i]), n[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
            except ValueError: pass
    stalls.sort(reverse=True)
    print('   stalls (warps per issue):', ', '.join('%s=%.2f' % (n, v) for v, n in stalls[:7]))

#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: per kernel duration, pipe utilisation, stall reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_blocks', 'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'sm__cycles_active.avg', 'sm__cycles_elapsed.max',
        'smsp__cycles_active.avg']
st = [h for h in hdr if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    print('-----', r[hdr.index('Kernel Name')])
    for k in keys:
        if k in hdr:
            print('   %-70s %s' % (k, r[hdr.index(k)]))
    vals = sorted([(float(r[hdr.index(n)].replace(',', '') or 0), n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for n in st], reverse=True)[:8]
    print('   stalls/issue:', ', '.join('%s %.2f' % (n, v) for v, n in vals))

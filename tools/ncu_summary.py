#!/usr/bin/env python3
"""Print the handful of ncu metrics that matter for the matching kernels from a `--page raw --csv` dump."""
import csv
import sys

WANT = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_op_read.sum',
    'lts__t_sectors_op_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
    'sm__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
    'sm__cycles_elapsed.avg', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
    'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'sm__maximum_warps_per_active_cycle_pct',
    'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
    for i, h in enumerate(hdr):
        if h in WANT or ('issue_stalled' in h and h.endswith('per_issue_active.ratio')):
            print('%-90s %-14s %s' % (h, units[i], r[i]))

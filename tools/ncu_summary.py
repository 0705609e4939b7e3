#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of counters DESIGN.md / profiles/ quote."""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu', 'launch__registers_per_thread', 'launch__occupancy_limit', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_bytes.sum', 'smsp__average_warp', 'smsp__warp_issue_stalled', 'smsp__average_warps_issue_stalled', 'launch__shared_mem',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__sass_thread_inst_executed_op_d', 'lts__t_sectors_op_red', 'lts__t_sectors_op_atom',
        'smsp__thread_inst_executed_per_inst_executed', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg', 'smsp__cycles_active.avg',
        'sm__pipe_fp64', 'sm__inst_executed_pipe_fp64', 'smsp__inst_executed_pipe_fp64', 'sm__sass_inst_executed_op_global', 'sm__sass_inst_executed_op_shared',
        'local_', 'lmem']
def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('== kernel', r[hdr.index('Kernel Name')][:60], 'id', r[0])
        for h, u, v in zip(hdr, units, r):
            if any(k in h for k in KEYS) and v not in ('', 'no data'):
                print('  %-90s %-10s %s' % (h, u, v))
main()

#!/usr/bin/env python
"""Print the handful of ncu raw-page metrics we track, one block per profiled kernel.  usage: ncu_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.max', 'smsp__average_warp_latency_per_inst_issued.ratio']
stalls = [k for k in hdr if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    print('-----')
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f'{k} = {r[i]} {units[i]}')
    st = sorted(((float(r[hdr.index(k)] or 0), k) for k in stalls), reverse=True)[:7]
    print('stalls: ' + ', '.join(f"{k.split('stalled_')[1].split('_per_issue')[0]}={v:.2f}" for v, k in st))

#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into markdown: duration, DRAM bytes, pipe utilisation, stall reasons.
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep [title] > profiles/summary.md"""
import csv
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'smsp__sass_inst_executed_op_shared_ld.sum', 'smsp__sass_inst_executed_op_shared_st.sum',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum',
        'smsp__sass_inst_executed_op_global_ld.sum', 'smsp__sass_inst_executed_op_global_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__average_warp_latency_per_inst_issued.ratio']
print('# %s\n' % title)
print('source: `%s` (ncu --set full --clock-control none)\n' % rep)
print('| metric | value | unit |\n|---|---|---|')
for k in keys:
    if k in d:
        print('| %s | %s | %s |' % (k, d[k][0], d[k][1]))
st = [(float(v[0]), k) for k, v in d.items() if 'average_warps_issue_stalled' in k and k.endswith('per_issue_active.ratio') and v[0]]
print('\nwarp stall reasons (warps stalled per issue-active cycle):\n')
print('| reason | ratio |\n|---|---|')
for v, k in sorted(st, reverse=True)[:9]:
    print('| %s | %.2f |' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))

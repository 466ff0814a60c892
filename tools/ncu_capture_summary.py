#!/usr/bin/env python
"""profiles/ncu_captures.json: per-kernel figures that only a profiler can measure (DRAM bytes per launch, pipe utilisation),
extracted from `ncu --set full --clock-control none` captures of the CURRENT build at the headline configuration.  bench.py
reads this file instead of carrying literals; every entry names its .ncu-rep, the commit and the date.

   python tools/ncu_capture_summary.py key=path.ncu-rep[:launch] [key=path ...]      keys: rollout_kernel_t4, gae, oz_gemm_kernel
"""
import csv
import datetime
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles', 'ncu_captures.json')


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return [{h: (v, u) for h, u, v in zip(hdr, units, r)} for r in rows[2:]]


def num(d, k):
    v, u = d.get(k, ('', ''))
    if v == '':
        return None
    x = float(v.replace(',', ''))
    scale = {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1.0, 'ms': 1.0, 'us': 1e-3, 'ns': 1e-6, 's': 1e3}.get(u, 1.0)
    return x * scale


def main():
    commit = subprocess.run(['git', '-C', ROOT, 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
    cur = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for arg in sys.argv[1:]:
        key, rep = arg.split('=', 1)
        sel = None
        if ':' in rep:                  # path.ncu-rep:i  -> launch i of the capture only
            rep, sel = rep.rsplit(':', 1)
        launches = raw(rep)
        if sel is not None:
            launches = [launches[int(sel)]]
        # several launches of one logical call (the GAE scan = 3 kernels) are summed
        ent = {'source': os.path.relpath(rep, ROOT), 'commit': commit, 'date': datetime.date.today().isoformat(),
               'kernels': [l['Kernel Name'][0] for l in launches],
               'dram_bytes': sum((num(l, 'dram__bytes_read.sum') or 0) + (num(l, 'dram__bytes_write.sum') or 0) for l in launches),
               'dram_bytes_read': sum(num(l, 'dram__bytes_read.sum') or 0 for l in launches),
               'dram_bytes_write': sum(num(l, 'dram__bytes_write.sum') or 0 for l in launches),
               'ms': sum(num(l, 'gpu__time_duration.sum') or 0 for l in launches)}
        l0 = launches[0]
        for name, k in (('fp64_pipe_active_pct', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'),
                        ('issue_active_pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
                        ('warps_active_pct', 'sm__warps_active.avg.pct_of_peak_sustained_active'),
                        ('imma_pipe_active_pct', 'sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active'),
                        ('dram_throughput_pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
                        ('registers_per_thread', 'launch__registers_per_thread'), ('grid', 'launch__grid_size'),
                        ('block', 'launch__block_size')):
            ent[name] = num(l0, k)
        cur[key] = ent
        print(key, json.dumps(ent)[:400])
    json.dump(cur, open(OUT, 'w'), indent=1)


if __name__ == '__main__':
    main()

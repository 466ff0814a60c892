#!/usr/bin/env python
"""Warp-state samples of an .ncu-rep (ncu --set full --import-source on) by source line:
   python tools/ncu_source.py gpurun_out/x.ncu-rep [min_pct] > profiles/x_src.md
Only the per-source-line rows of `--page source --print-source cuda,sass` are used (the SASS rows under them are
already summed into their line)."""
import csv
import io
import os
import subprocess
import sys

rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
KEYS = ['barrier', 'wait', 'short_sb', 'long_sb', 'selected', 'math', 'mio', 'lg', 'no_inst', 'branch_resolving', 'dispatch', 'not_selected']
cur_file, col = None, None
lines = []          # (file, line_no, source, samples, inst, {stall: n})
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = os.path.basename(r[1]); col = None; continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No':
        col = {h: i for i, h in enumerate(r)}
        # duplicate 'Source' header: first = cuda, second = sass
        continue
    if col is None or not r[0]:
        continue
    try:
        s = float(r[col['# Samples']] or 0)
        ins = float(r[col['Instructions Executed']] or 0)
    except (ValueError, IndexError):
        continue
    st = {}
    for k in KEYS:
        i = col.get('stall_' + k)
        if i is not None and i < len(r) and r[i]:
            st[k] = float(r[i])
    lines.append((cur_file, int(r[0]), r[1].strip(), s, ins, st))
tot = sum(l[3] for l in lines) or 1.0
tot_i = sum(l[4] for l in lines) or 1.0
print('# warp-state samples by source line: `%s`\n' % rep)
print('total samples %d, warp instructions %d; lines with >= %.1f%% of the samples\n' % (tot, tot_i, min_pct))
print('```')
print(' smp%   inst%  file:line  dominant stalls | source')
for f, ln, src, s, ins, st in lines:
    if 100.0 * s / tot >= min_pct:
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print('%5.2f  %5.2f  %s:%d  %s | %s' % (100.0 * s / tot, 100.0 * ins / tot_i, f, ln,
                                              ' '.join('%s=%.0f%%' % (k, 100.0 * v / max(s, 1)) for k, v in top), src[:100]))
print('```\n')
# totals per file and per stall
byf = {}
for f, ln, src, s, ins, st in lines:
    a = byf.setdefault(f, [0.0, 0.0]); a[0] += s; a[1] += ins
print('| file | samples % | instructions % |\n|---|---|---|')
for f, (s, i) in sorted(byf.items(), key=lambda kv: -kv[1][0]):
    print('| %s | %.1f | %.1f |' % (f, 100 * s / tot, 100 * i / tot_i))
agg = {}
for l in lines:
    for k, v in l[5].items():
        agg[k] = agg.get(k, 0.0) + v
print('\n| stall | % of samples |\n|---|---|')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print('| %s | %.1f |' % (k, 100 * v / tot))

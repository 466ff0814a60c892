#!/usr/bin/env python
"""Summarise an ncu launch list (ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file x.csv <cmd>)
into a per-kernel table: launches, total time, share.  Per-launch times are cold-cache and serialised: compare SHARES.
   python tools/launch_summary.py gpurun_out/launches.csv "title" > profiles/summary.md"""
import collections
import csv
import re
import sys

path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else path
rows = []
with open(path, newline='') as f:
    lines = [ln for ln in f if ln.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
tot = collections.Counter()
cnt = collections.Counter()
for r in rd:
    if len(r) != len(hdr) or r[ix['Metric Name']] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r[ix['Kernel Name']])
    v = float(r[ix['Metric Value']].replace(',', ''))
    unit = r[ix['Metric Unit']]
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(unit, 1e-6)
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print('# %s\n' % title)
print('source: `%s` (%d launches, %.1f ms total device time); per-launch times are cold-cache and serialised: compare SHARES\n'
      % (path, sum(cnt.values()), total))
print('| kernel | launches | total ms | share |\n|---|---|---|---|')
for name, v in tot.most_common(28):
    print('| %s | %d | %.1f | %.1f %% |' % (name[:90], cnt[name], v, 100 * v / total))

#!/usr/bin/env python
"""Blackwell-specific SASS mnemonics per kernel of libegopose_b200.so (cuobjdump -sass): evidence that the kernels use the
sm_100a units they claim (UTCIMMA = tcgen05.mma kind::i8, UTMALDG / UTMASTG = TMA load / store, LDTM / STTM = tcgen05.ld / st,
DMMA = FP64 tensor-core mma.sync, DFMA = FP64 pipe).   python tools/sass_summary.py > profiles/r2_sass_summary.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'egopose_b200', 'libegopose_b200.so')
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
KEYS = ['UTCIMMA', 'UTCHMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'STTM', 'UTCBAR', 'SYNCS', 'DMMA', 'DFMA', 'IMMA', 'LDGSTS', 'BAR.SYNC', 'REDUX']
cur, counts, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
print('# SASS mnemonic counts per kernel (cuobjdump -sass egopose_b200/libegopose_b200.so, sm_100a)\n')
print('| kernel | instructions | ' + ' | '.join(KEYS) + ' |')
print('|---|---|' + '---|' * len(KEYS))
for fn in sorted(total, key=lambda f: -total[f]):
    if total[fn] < 200 and not counts[fn]:
        continue
    name = re.sub(r'\(.*', '', fn)[:70]
    print('| `%s` | %d | ' % (name, total[fn]) + ' | '.join(str(counts[fn][k]) if counts[fn][k] else '' for k in KEYS) + ' |')

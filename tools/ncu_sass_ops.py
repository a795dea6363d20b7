#!/usr/bin/env python
"""Opcode histogram (warp instructions executed, stall samples) from `ncu --page source --csv --print-source sass`.
usage: ncu_sass_ops.py sass.csv [draws]   -- with `draws`, also warp instructions per 32 draws"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {k: i for i, k in enumerate(hdr)}
ops = collections.Counter()
smp = collections.Counter()
tot = 0
for r in rows[2:]:
    if r and r[0] in ("Address", "Kernel Name"):
        continue
    if len(r) != len(hdr):
        continue
    s = r[ix['Source']].strip()
    s = re.sub(r'^@!?U?P\d+\s+', '', s)
    op = s.split()[0].split('.')[0] if s else '?'
    full = s.split()[0] if s else '?'
    if op in ('LDS', 'STS', 'LDG', 'STG', 'ST', 'LD'):
        op = full.split('.')[0]
    n = float(r[ix['Instructions Executed']])
    ops[op] += n
    smp[op] += float(r[ix['# Samples']])
    tot += n
draws = float(sys.argv[2]) if len(sys.argv) > 2 else None
stot = sum(smp.values())
print('total warp instructions %.4g' % tot, ('= %.1f per 32 draws' % (tot * 32 / draws)) if draws else '')
for op, n in ops.most_common(40):
    print('%-10s %6.2f%%  %s samples %5.1f%%' % (op, 100 * n / tot, ('%5.2f/32 draws' % (n * 32 / draws)) if draws else '', 100 * smp[op] / stot))

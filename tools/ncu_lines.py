#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source cuda,sass` output per CUDA source line.
usage: ncu -i rep --page source --csv --print-source cuda,sass > src.csv; python tools/ncu_lines.py src.csv [section] [top]"""
import collections
import csv
import sys


def f(v):
    try:
        return float(v)
    except ValueError:
        return 0.0


rows = list(csv.reader(open(sys.argv[1])))
want = int(sys.argv[2]) if len(sys.argv) > 2 else -1
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
sections, hdr, data, path = [], None, None, None
for r in rows:
    if r and r[0] == 'File Path':
        path = r[1]
    if r and r[0] == 'Line No':
        hdr, data = r, []
        sections.append((path, hdr, data))
        continue
    if hdr is not None and len(r) == len(hdr):
        data.append(r)
for si, (path, h, d) in enumerate(sections):
    ix = {}
    for n, k in enumerate(h):
        ix.setdefault(k, n)
    tot = sum(f(r[ix['Instructions Executed']]) for r in d)
    smp_tot = sum(f(r[ix['# Samples']]) for r in d)
    print('section', si, path, 'rows', len(d), 'warp inst %.4g' % tot, 'samples', smp_tot)
    if want >= 0 and si != want:
        continue
    agg = collections.OrderedDict()
    for r in d:
        a = agg.setdefault(r[0], [r[1], 0, 0, 0])
        a[1] += f(r[ix['Instructions Executed']])
        a[2] += f(r[ix['Thread Instructions Executed']])
        a[3] += f(r[ix['# Samples']])
    for ln, (s, n, t, smp) in sorted(agg.items(), key=lambda kv: -kv[1][3])[:top_n]:
        print('%5s inst %5.1f%% lanes %4.1f samples %5.1f%%  %s' % (ln, 100 * n / max(tot, 1), t / max(n, 1), 100 * smp / max(smp_tot, 1), s.strip()[:100]))

#!/usr/bin/env python
"""Per-kernel totals of the last step of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: launch_shares.py launches.csv [first-kernel-of-a-step substring, default k_colsums]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
ix = {k: i for i, k in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
key = sys.argv[2] if len(sys.argv) > 2 else 'k_colsums'
starts = [i for i, r in enumerate(data) if key in r[ix['Kernel Name']]]
last = data[starts[-1]:]
agg = collections.OrderedDict()
tot = 0.0
for r in last:
    n = re.sub(r'\(.*', '', r[ix['Kernel Name']])
    t = float(r[ix['Metric Value']].replace(',', ''))
    u = r[ix['Metric Unit']]
    t = t / 1000 if u == 'ns' else (t * 1000 if u == 'ms' else t)
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += t
    tot += t
print('kernel,launches,total_us,share')
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%s,%d,%.1f,%.3f' % (n[:70], c, t, t / tot))
print('TOTAL,%d,%.1f,1' % (len(last), tot))

#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python scripts/launch_summary.py launches.csv > summary.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == 'ID':
        hdr, start = r, i + 1
        break
iK, iV, iU = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[start:]:
    if len(r) <= iV:
        continue
    us = float(r[iV].replace(',', '')) * scale.get(r[iU].strip().replace('usecond', 'us').replace('nsecond', 'ns').replace('msecond', 'ms'), 1.0)
    tot[r[iK]] += us
    cnt[r[iK]] += 1
T = sum(tot.values())
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches", "total_us", "avg_us", "share"])
for k, v in tot.most_common():
    w.writerow([k, cnt[k], "%.1f" % v, "%.1f" % (v / cnt[k]), "%.4f" % (v / T)])

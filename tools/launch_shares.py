#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys

rows = [x for x in csv.reader(open(sys.argv[1])) if len(x) > 5]
hdr, agg = None, {}
for x in rows:
    if x[0] == 'ID':
        hdr = x
        continue
    if hdr is None:
        continue
    n = x[hdr.index('Kernel Name')].split('(')[0]
    v = float(x[hdr.index('Metric Value')].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}[x[hdr.index('Metric Unit')]]
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-52s n=%5d total=%10.1f us  avg=%9.1f us %5.1f%%" % (n[:52], a[0], a[1], a[1] / a[0], 100 * a[1] / tot))

#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
iK = hdr.index('Kernel Name'); iV = hdr.index('Metric Value'); iU = hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[iV].replace(',', ''))
    except ValueError:
        continue
    u = r[iU]
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else (v * 1e6 if u in ('s', 'second') else v))
    a = agg.setdefault(r[iK][:78], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("%-80s %5s %12s %6s %10s" % ("kernel", "n", "total us", "share", "mean us"))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-80s %5d %12.1f %5.1f%% %10.1f" % (k, a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))
print("total %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))

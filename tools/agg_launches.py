"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
Usage: python tools/agg_launches.py gpurun_out/launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i
        break
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    a = agg.setdefault(r[ki][:100], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%6d %10.1f us %5.1f%%  avg %8.1f us  %s" % (a[0], a[1] / 1e3, 100 * a[1] / tot, a[1] / a[0] / 1e3, n))

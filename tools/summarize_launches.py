"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel count, total and share.
Usage: python tools/summarize_launches.py gpurun_out/launches.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
tot = collections.defaultdict(float)
cnt = collections.Counter()
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    name = re.sub(r"^void ", "", name)
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print("%d launches, %.1f us of kernel time (cold-cache, serialised: compare shares)" % (sum(cnt.values()), total))
for name, t in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%6d x %8.2f us = %9.1f us  %5.1f%%  %s" % (cnt[name], t / cnt[name], t, 100 * t / total, name[:110]))

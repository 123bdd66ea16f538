"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --steps 1 --warmup 3 --quick` by kernel
(last of the 4 identical steps).  usage: python tools/launch_summary.py gpurun_out/launches.csv [n_rows]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
ls = []
for r in rows[start:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    ls.append((r[ki], v))
per = len(ls) // 4
step = ls[-per:]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in step:
    k = re.sub(r"\(.*", "", k)
    k = re.sub(r"^void ", "", k)
    agg[k[:120]][0] += 1
    agg[k[:120]][1] += v
tot = sum(v for _, v in step)
print("%d launches, %.1f us" % (per, tot))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:n]:
    print("%9.1f %5.1f%% %4d %8.1f  %s" % (t, 100 * t / tot, c, t / c, k))

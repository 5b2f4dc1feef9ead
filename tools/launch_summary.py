"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals over the LAST
`frac` of the launches (skip warm-up).  usage: launch_summary.py launches.csv [frac_or_count]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
data = rows[hdr + 1:]
arg = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
n = int(arg) if arg > 1 else int(len(data) * arg)
data = data[-n:]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    v = float(r[vi].replace(",", ""))
    if r[ui] in ("ns", "nsecond"):
        v /= 1e3
    elif r[ui] in ("ms", "msecond"):
        v *= 1e3
    k = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:60]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print("launches: %d  total %.1f us" % (len(data), tot))
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print("%9.1f us %5.1f%% x%4d %s" % (v, 100 * v / tot, c, k))

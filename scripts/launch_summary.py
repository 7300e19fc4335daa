"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    a = agg.setdefault(r[ki][:70], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for k, a in agg.items() if "dfma_peak" not in k)
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} {n:4d} {t / n:9.1f} us  {100 * t / tot:5.1f}%")
print("total us (without the peak-measurement kernel)", tot)

"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for row in rows[1:]:
    if len(row) <= vi:
        continue
    k = row[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    v = float(row[vi].replace(",", ""))
    u = row[ui]
    v = v / 1e3 if u in ("usecond", "us") else v / 1e6 if u in ("nsecond", "ns") else v
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
for k, (n, t) in agg.items():
    print(f"{k:40s} launches={n:3d} total_ms={t:9.3f} avg_ms={t / n:8.4f} share={t / tot * 100:5.1f}%")
print(f"{'total':40s} {tot:.3f} ms")

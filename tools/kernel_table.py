"""Per-kernel averages from an `ncu --metrics ... --csv` launch list with several metrics per launch.
usage: python tools/kernel_table.py launches.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
h = rows[0]
ki, mi, vi, ui, idi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    k = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    per.setdefault((r[idi], k), {})[r[mi]] = float(r[vi].replace(",", ""))
agg = collections.OrderedDict()
for (_i, k), d in per.items():
    a = agg.setdefault(k, collections.defaultdict(float))
    a["n"] += 1
    for m, v in d.items():
        a[m] += v
tot_t = tot_i = 0.0
for k, a in agg.items():
    n = a["n"]
    t, ins = a["gpu__time_duration.sum"] / n / 1e3, a["smsp__inst_executed.sum"] / n / 1e6
    tot_t += a["gpu__time_duration.sum"] / 1e3
    tot_i += a["smsp__inst_executed.sum"] / 1e6
    print(f"{k:24s} n={int(n):3d} t={t:8.1f}us inst={ins:7.2f}M rd={a['dram__bytes_read.sum'] / n / 1e6:8.2f}MB wr={a['dram__bytes_write.sum'] / n / 1e6:7.2f}MB "
          f"warps={a['sm__warps_active.avg.pct_of_peak_sustained_active'] / n:5.1f}% issue={a['smsp__issue_active.avg.pct_of_peak_sustained_active'] / n:5.1f}%")
print(f"all launches: {tot_t:.1f} us, {tot_i:.1f} M warp instructions")

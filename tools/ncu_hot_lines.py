"""Top CUDA source lines by warp-stall samples for one kernel of an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_hot_lines.py report.ncu-rep kernel_regex [launch_index] [top_n]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kern}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
agg = []
for r in csv.reader(out.splitlines()):
    if len(r) > 6 and r[0].isdigit():
        try:
            agg.append((int(r[4]), int(r[5]), int(r[0]), r[1].strip()[:105]))
        except ValueError:
            pass
tot = sum(a[0] for a in agg) or 1
print(f"{kern}: {tot} samples")
for s, ex, ln, src in sorted(agg, reverse=True)[:top]:
    print(f"{s:7d} {s / tot * 100:5.1f}%  inst={ex:9d}  L{ln:<4d} {src}")

"""Developer script: per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/summarize_launches.py gpurun_out/launches.csv "<command line that was profiled>" > profiles/rNN_launches_summary.txt"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.DictReader(lines))
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
    name = r["Kernel Name"]
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
ours = sum(v for k, v in tot.items() if "hb::" in k)
print("ncu --metrics gpu__time_duration.sum --clock-control none ; %s (1 GPU, B200)" % (sys.argv[2] if len(sys.argv) > 2 else ""))
print("per-launch times are cold-cache and serialised: compare SHARES, not absolutes.  share_hb = share among this library's kernels")
print()
print("%-110s %5s %12s %8s %9s" % ("kernel", "n", "total_us", "share", "share_hb"))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%-110s %5d %12.1f %7.2f%% %9s" % (k[:110], cnt[k], v, 100 * v / total, ("%.2f%%" % (100 * v / ours)) if "hb::" in k else "-"))

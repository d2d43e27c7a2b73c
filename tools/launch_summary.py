"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/rN_launches.csv) by kernel:
launches, total and average duration, share.   python tools/launch_summary.py profiles/r1_launches.csv"""
import csv, sys, collections

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4]
    short = name.split("(")[0].replace("void ", "").replace("at::", "").replace("rc::", "")[:70]
    ns = float(r[14].replace(",", ""))
    unit = r[13]
    us = ns * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':72s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} {n:8d} {us:12.1f} {us / n:10.1f} {100 * us / tot:6.1f}%")

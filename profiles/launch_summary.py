"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X.csv <command>).
    python profiles/launch_summary.py profiles/r01_launches.csv "<command>" > profiles/r01_launches_summary.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = next(r for r in rows if "Kernel Name" in r)
k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    if r is hdr or r[k] == "Kernel Name":
        continue
    try:
        t = float(r[v].replace(",", ""))
    except ValueError:
        continue
    t *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[u], 1.0)
    tot[r[k]] += t
    cnt[r[k]] += 1
all_t = sum(tot.values())
print(f"# ncu --metrics gpu__time_duration.sum --clock-control none, {sys.argv[2] if len(sys.argv) > 2 else ''} (cold, serialised)")
for name, t in tot.most_common(25):
    print(f"{t:12.1f} us total {cnt[name]:5d} launches {t / cnt[name]:12.1f} us avg {100 * t / all_t:6.1f}%  {name[:70]}")

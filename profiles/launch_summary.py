#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python profiles/launch_summary.py gpurun_out/launches_r1.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
tot = collections.OrderedDict()
for r in rows:
    name, val = r[4], float(r[-1])          # "Kernel Name", metric value (ns)
    t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += val / 1e3
allus = sum(v[1] for v in tot.values())
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:100]:100s} n={n:3d} total={us:12.1f}us mean={us / n:12.1f}us share={100 * us / allus:5.1f}%")

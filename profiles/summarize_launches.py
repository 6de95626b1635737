#!/usr/bin/env python
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --csv --log-file X.csv ...`):
per (kernel, grid): launches, total / mean device time, warp instructions per launch, share of all captured time.
usage: python profiles/summarize_launches.py X.csv [--ours-csv OUT.csv] > summary.md
--ours-csv also writes the rows of this library's kernels (k_*) only, to keep the committed list small."""
import collections, csv, sys
src = sys.argv[1]
ours_out = sys.argv[sys.argv.index("--ours-csv") + 1] if "--ours-csv" in sys.argv else None
rows = list(csv.reader(open(src, errors="replace")))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[start]
ki, gi, mi, vi, ii = h.index("Kernel Name"), h.index("Grid Size"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
per = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    d = per.setdefault(r[ii], {"name": r[ki], "grid": r[gi]})
    d[r[mi]] = float(r[vi].replace(",", ""))
agg = collections.OrderedDict()
for d in per.values():
    name = d["name"].split("(")[0]
    a = agg.setdefault((name, d["grid"]), [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0) / 1e3
    a[2] += d.get("smsp__inst_executed.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print("| kernel | grid | launches | total us | mean us | warp instr / launch | share of all captured time |")
print("|---|---|---|---|---|---|---|")
for (name, grid), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if a[1] < 0.002 * tot:
        continue
    print(f"| {name[:140]} | {grid} | {a[0]} | {a[1]:.0f} | {a[1] / a[0]:.1f} | {a[2] / a[0]:.3g} | {100 * a[1] / tot:.1f} % |")
if ours_out:
    w = csv.writer(open(ours_out, "w"))
    w.writerow(["ID", "Kernel Name", "Grid Size", "gpu__time_duration.sum [ns]", "smsp__inst_executed.sum"])
    for k, d in per.items():
        n = d["name"]
        if n.startswith("k_") or n.startswith("void k_"):
            w.writerow([k, n.split("(")[0], d["grid"], int(d.get("gpu__time_duration.sum", 0)), int(d.get("smsp__inst_executed.sum", 0))])

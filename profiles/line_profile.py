#!/usr/bin/env python
"""Attribute an ncu source-page dump to CUDA source lines: joins `ncu -i X.ncu-rep --page source --csv` (per-SASS executed
counts and stall samples, SASS order) with `nvdisasm -g -c` line info of the same cubin.
usage: python profiles/line_profile.py <src.csv> <libm17b200.so> <kernel-substring> [units_per_launch]"""
import collections, csv, os, re, subprocess, sys, tempfile
srccsv, lib, kern = sys.argv[1:4]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines, cur, on = [], None, False
for l in dis.split("\n"):
    if l.startswith("//---") and ".text." in l:
        on = kern in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append(cur)
rows = list(csv.reader(open(srccsv)))
ins, hdr, k = [], None, None
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        k = r[1]; hdr = None; continue
    if r[0] == "Address":
        hdr = r; continue
    if hdr and k and kern.split("I")[0].replace("_Z", "").lstrip("0123456789") in k and len(r) == len(hdr):
        ins.append(dict(zip(hdr, r)))
n = min(len(ins), len(lines))
agg, smp = collections.Counter(), collections.Counter()
for i in range(n):
    agg[lines[i]] += int(ins[i]["Instructions Executed"]); smp[lines[i]] += int(ins[i]["# Samples"])
tot, ts = sum(agg.values()), sum(smp.values())
print(f"# {kern}: {len(ins)} SASS / {len(lines)} disassembled; {tot / units:.1f} warp-instr per unit; {ts} samples")
srcdir = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc")
cache = {}
for (f, ln), v in sorted(smp.items(), key=lambda x: -x[1])[:40]:
    if f not in cache:
        p = os.path.join(srcdir, f)
        cache[f] = open(p).read().split("\n") if os.path.exists(p) else []
    text = cache[f][ln - 1].strip()[:100] if ln - 1 < len(cache[f]) else ""
    print(f"{v * 100 / ts:5.1f}% time {agg[(f, ln)] / units:8.1f} instr/unit  {f}:{ln}  {text}")

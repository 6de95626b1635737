#!/usr/bin/env python
"""Summarise an ncu --page source --csv dump: per kernel, stall-reason totals and the hottest SASS instructions.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python tools_ncu_summary.py src.csv [topN]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
kern = None; hdr = None; data = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name": kern = r[1].split("(")[0]; data[kern] = []; hdr = None; continue
    if r[0] == "Address": hdr = r; continue
    if hdr and kern: data[kern].append(dict(zip(hdr, r)))
for k, ins in data.items():
    tot = sum(int(i["# Samples"]) for i in ins)
    ex = sum(int(i["Instructions Executed"]) for i in ins)
    print(f"== {k}: {len(ins)} SASS instrs, {ex} warp-instr executed, {tot} samples")
    st = collections.Counter()
    for i in ins:
        for key, v in i.items():
            if key.startswith("stall_") and "Not Issued" not in key: st[key] += int(v or 0)
    print("   stalls:", ", ".join(f"{a[6:]}={b*100//max(tot,1)}%" for a, b in st.most_common(8)))
    mix = collections.Counter()
    for i in ins:
        op = i["Source"].split()[0] if not i["Source"].strip().startswith("@") else i["Source"].split()[1]
        mix[op.split(".")[0]] += int(i["Instructions Executed"])
    print("   mix:", ", ".join(f"{a}={b*100//max(ex,1)}%" for a, b in mix.most_common(14)))
    for i in sorted(ins, key=lambda x: -int(x["# Samples"]))[:top]:
        s = {kk[6:]: int(v) for kk, v in i.items() if kk.startswith("stall_") and "Not Issued" not in kk and int(v or 0) > 0}
        s = sorted(s.items(), key=lambda x: -x[1])[:3]
        print(f"   {int(i['# Samples'])*100/max(tot,1):5.1f}%  ex={i['Instructions Executed']:>9s}  {i['Source'].strip()[:70]:70s} {s}")

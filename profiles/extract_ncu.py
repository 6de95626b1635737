#!/usr/bin/env python
"""Turn `ncu -i X.ncu-rep --page raw --csv` into the per-kernel table kept under profiles/.
usage: python profiles/extract_ncu.py raw.csv > table.md"""
import csv, sys
KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__inst_executed.sum", "warp instr"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/instr"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_static", "smem static"), ("launch__shared_mem_per_block_dynamic", "smem dyn"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
print("| kernel | " + " | ".join(n for _, n in KEYS) + " |")
print("|---|" + "---|" * len(KEYS))
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    cells = []
    for k, _ in KEYS:
        if k in idx:
            v = r[idx[k]]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            cells.append(f"{v} {units[idx[k]]}".strip())
        else:
            cells.append("n/a")
    print(f"| {name} | " + " | ".join(cells) + " |")

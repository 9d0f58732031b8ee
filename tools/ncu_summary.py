#!/usr/bin/env python
"""ncu report -> a small JSON summary (the numbers DESIGN.md and bench.py quote).

    python tools/ncu_summary.py gpurun_out/X.ncu-rep [fields_per_launch] > profiles/X_summary.json"""
import csv
import json
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
fields = int(sys.argv[2]) if len(sys.argv) > 2 else 16
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
get = lambda k: float(vals[hdr.index(k)]) if k in hdr and vals[hdr.index(k)] not in ("", "n/a") else None
name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
keep = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
out = {"kernel": name, "report": rep, "fields_per_launch": fields, "metrics": {}}
for k in keep:
    if k in hdr:
        out["metrics"][k] = {"value": get(k), "unit": units[hdr.index(k)]}
for i, h in enumerate(hdr):
    if "stalled" in h and "per_issue_active" in h:
        v = float(vals[i] or 0)
        if v >= 0.2:
            out["metrics"][h] = {"value": v, "unit": units[i]}
dr, dw = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
if dr is not None and dw is not None:
    out["dram_bytes_per_launch"] = dr * scale.get(units[hdr.index("dram__bytes_read.sum")], 1.0) + dw * scale.get(units[hdr.index("dram__bytes_write.sum")], 1.0)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
if len(srows) > 2 and "Instructions Executed" in srows[1]:
    h2 = srows[1]
    isrc, iex = h2.index("Source"), h2.index("Instructions Executed")
    c = Counter()
    tot = 0
    for r in srows[2:]:
        if len(r) <= iex:
            continue
        s = r[isrc].strip()
        op = (s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]
        c[op] += int(r[iex])
        tot += int(r[iex])
    out["warp_instructions"] = tot
    out["dynamic_opcode_share_pct"] = {op: round(100.0 * n / tot, 2) for op, n in c.most_common(30)}
print(json.dumps(out, indent=1))

#!/usr/bin/env python
"""tools/ncu_summary.py REP.ncu-rep OUT.csv [TRAFFIC.json] -- the columns of an `ncu --set full` capture that profiles/*_ncu_summary.csv keep
(one row per profiled launch), and per-kernel DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum per launch) for bench.py's roofline.traffic."""
import csv
import json
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
KEEP = ["Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
KEEP += sorted(h for h in hdr if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", h))
ix = [hdr.index(k) for k in KEEP if k in hdr]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in ix]); w.writerow([units[i] for i in ix])
    for r in data:
        w.writerow([r[i] for i in ix])
if len(sys.argv) > 3:
    def to_bytes(v, u):
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    tr = {}
    for r in data:
        name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", r[ik]).split("(")[0]
        tr.setdefault(name, []).append(to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]))
    tj = {k: int(sum(v) / len(v)) for k, v in tr.items()}
    tj["_source"] = f"dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full capture {rep.split('/')[-1]} (summary: {out}); not measured by the bench run itself"
    json.dump(tj, open(sys.argv[3], "w"), indent=1)

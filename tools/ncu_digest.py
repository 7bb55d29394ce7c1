#!/usr/bin/env python
"""tools/ncu_digest.py RAW.csv SOURCE.csv -- the numbers read off an `ncu --set full --import-source on` capture in this repo: stall reasons per
issue, pipe / issue utilisation, then stall samples per 512-instruction region of the SASS and the hottest instructions.
(RAW = `ncu -i rep --page raw --csv`, SOURCE = `ncu -i rep --page source --csv`.)"""
import collections
import csv
import re
import sys

raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
m = {h: v for h, v in zip(rows[0], rows[2])}
for k in sorted(m):
    if re.search(r"^(gpu__time_duration.sum|l1tex__t_sector_hit_rate.pct|smsp__average_warps_issue_stalled.*per_issue_active|smsp__issue_active.avg.pct|sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active|sm__warps_active.avg.per_cycle_active|smsp__inst_executed.sum$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|launch__registers_per_thread$|launch__grid_size|dram__bytes_read.sum$|dram__bytes_write.sum$)", k):
        if m[k] not in ("0", ""):
            print(f"{k:95s} {m[k]:>16s}")
rows = list(csv.reader(open(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0


tot = sum(f(r, "# Samples") for r in data)
print("samples", tot, "instructions", len(data))
B = 512
for b in range(0, len(data), B):
    ch = data[b:b + B]
    if sum(f(r, '# Samples') for r in ch) == 0:
        continue
    print(f"  [{b:5d}] samples {sum(f(r,'# Samples') for r in ch):6.0f} no_inst {sum(f(r,'stall_no_inst') for r in ch):5.0f} lsb {sum(f(r,'stall_long_sb') for r in ch):5.0f} wait {sum(f(r,'stall_wait') for r in ch):5.0f} sel {sum(f(r,'stall_selected') for r in ch):5.0f} ssb {sum(f(r,'stall_short_sb') for r in ch):5.0f} math {sum(f(r,'stall_math') for r in ch):4.0f} mio {sum(f(r,'stall_mio') for r in ch):4.0f} exec/tile {sum(f(r,'Instructions Executed') for r in ch)/max(1,float(m.get('launch__grid_size','1')) if float(m.get('launch__grid_size','1'))>1000 else 15625):7.1f}")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    print(f"{r[ix['Address']][-5:]:>6s} {f(r,'# Samples'):6.0f} lsb {f(r,'stall_long_sb'):5.0f} ssb {f(r,'stall_short_sb'):5.0f} wait {f(r,'stall_wait'):5.0f} noi {f(r,'stall_no_inst'):4.0f} {r[ix['Source']][:80]}")

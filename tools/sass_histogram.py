#!/usr/bin/env python
"""tools/sass_histogram.py [lib.so] -- static SASS opcode counts of the hot kernels (cuobjdump -sass), the evidence for which instructions a
kernel is made of: DFMA / DMUL / DADD (fp64 pipe), UBLKCP (TMA bulk copies), LDGSTS (cp.async), STL / LDL (local-memory spills), DMMA ...
Writes a markdown table; committed per round under profiles/ (round 2: profiles/r2_sass_histogram.md)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "kontiki_b200", "lib", "libkontiki_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KERNELS = ["k_static_rs", "k_static_rs_local", "k_landmark_ref", "k_imu<0>", "k_imu<1>", "k_short_batch", "k_pair_prepass", "k_static_rs_split", "k_imu_split<0>", "k_imu_split<1>",
           "k_newton_rs", "k_newton_rs_fast", "k_newton_rs_rev", "k_newton_rs_two_w", "k_lifting_rs", "k_span_rs_split", "k_span_sensor", "k_gn_gather<7>", "k_gn_blocks<7>", "k_gn_rows_apply"]
COLS = ["total", "DFMA", "DMUL", "DADD", "DMMA", "MUFU", "LDG", "STG", "LDS", "STS", "LDGSTS", "UBLKCP", "LDL", "STL", "SHFL", "BRA", "CALL", "IMAD", "MOV", "LDC", "LDCU", "UMOV"]
demangle = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
funcs = re.split(r"\n\s*Function : ", sass)[1:]
rows = {}
for name, body in zip(demangle, funcs):
    short = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name).replace("(int)", "").split("(")[0].replace("void ", "")
    ops = collections.Counter()
    for m in re.finditer(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", body):
        ops[m.group(1)] += 1
    ops["total"] = sum(v for k, v in ops.items() if k != "total")
    rows[short] = ops
print("| kernel | " + " | ".join(COLS) + " |")
print("|---|" + "---:|" * len(COLS))
for k in KERNELS:
    if k in rows:
        print(f"| `{k}` | " + " | ".join(str(rows[k].get(c, 0)) for c in COLS) + " |")

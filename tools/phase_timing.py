#!/usr/bin/env python
"""tools/phase_timing.py -- where a k_static_rs tile spends its time: clock64 sums per phase (lane 0 of every warp), from an experimental build
(tools/build_variant.sh phases -DKTK_PHASE_TIMING; run with KTK_LIB=gpurun_variants/libktk_phases.so python tools/phase_timing.py)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kontiki_b200 import _lib, synthetic as syn      # noqa: E402

cfg = syn.make_config("C3")
c = cfg["cam"]
p = _lib.Problem(0)
p.set_se3_spline(cfg["dt"], 0.0, len(cfg["knots"]))
p.add_static_rs(_lib.make_camera(c["rows"], c["cols"], c["readout"], c["K"]), c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"], c["weight"], c["huber_c"])
flags = _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS | _lib.EVAL_ROBUST
import torch      # noqa: E402
dev = torch.device("cuda", 0)
n = len(c["lm_idx"])
r, J = torch.empty((n, 2), dtype=torch.float64, device=dev), torch.empty((n, 114), dtype=torch.float64, device=dev)
i0, i0b = torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.int32, device=dev)
outs = [dict(r=r.data_ptr(), J=J.data_ptr(), i0=i0.data_ptr(), i0_b=i0b.data_ptr(), i0_c=None, i0_d=None)]
dk, dr = torch.from_numpy(cfg["knots"].reshape(-1)).to(dev), torch.from_numpy(c["rho"]).to(dev)
L = _lib.lib()
out = (C.c_ulonglong * 8)()
for rep in range(3):
    p.evaluate_device(dk.data_ptr(), dr.data_ptr(), len(c["rho"]), flags, outs)
    p.synchronize()
    L.ktk_debug_read_phases(out, 1)
names = ["inputs issued", "wait inputs + gather issue", "observation pose", "wait records", "projection + ref half", "obs half + r/idx stores", "store issue", "wait TMA read"]
tot = sum(out)
tiles = (n + 31) // 32
for k, nm in enumerate(names):
    print(f"{nm:32s} {out[k] / tiles:9.0f} cycles/tile  {100.0 * out[k] / tot:5.1f} %")
print(f"{'total':32s} {tot / tiles:9.0f} cycles/tile")

"""tools/newton_soak.py -- CPU soak of the closed-form NewtonRs rows (host-compiled csrc/newton_math.cuh, the text the kernels run) against forward mode through
the iteration: 300 seeded cases x 240 rows, observed rows anywhere in the image / on the fence of the half-row break test / 30 rows off, pinhole and atan,
with and without the Huber corrector.  Prints the number of rows by evaluations of the iteration, status / iteration-count mismatches (expected: none) and
the worst relative deviation of a Jacobian row.  Takes half a minute."""
import sys, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import hostcheck as hc
import test_math_host as T
tot = np.zeros(7, int); worst = np.zeros(7); bad_iter = 0; bad_status = 0; nrows = 0
for seed in range(100, 400):
    dt = (0.02, 0.05, 0.1, 0.03)[seed % 4]; atan = bool(seed & 1)
    knots, s, cam = T._camera_case_model(dt, seed, atan, "newton")
    rng = np.random.default_rng(seed)
    uv = s["obs_uv"].copy()
    mode = seed % 3
    if mode == 0: uv[:,1] = rng.uniform(0.0, cam.rows - 1.0, len(uv))          # anywhere, including the image border (clamping)
    elif mode == 1: uv += rng.normal(0, 0.6, uv.shape)                          # around half a row: the break test sits on the fence
    else: uv[:,1] = np.clip(uv[:,1] + rng.normal(0, 30, len(uv)), 0, cam.rows - 1.0)
    hub = np.full(len(uv), 5.0) if seed % 5 == 0 else None
    args = (knots, dt, 0.0, cam, uv, s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"], s["weight"])
    h0 = hc.newton_rs(*args, huber_c=hub); h2 = hc.newton_rs(*args, huber_c=hub, fast=2)
    nrows += len(uv)
    bad_status += int((h0["status"] != h2["status"]).sum())
    ok = (h0["status"]==0)&(h2["status"]==0)
    bad_iter += int((h0["iterations"] != h2["iterations"])[ok].sum())
    same = ok & (h0["iterations"] == h2["iterations"])
    sc = np.abs(h0["J"]).max(axis=1, keepdims=True) + 1e-300
    err = (np.abs(h2["J"] - h0["J"]) / sc).max(axis=1)
    rerr = np.abs(h2["r"] - h0["r"]).max(axis=1)
    for k in range(1,6):
        m = same&(h0["iterations"]==k)
        tot[k] += m.sum()
        if m.any(): worst[k] = max(worst[k], err[m].max()); worst[6] = max(worst[6], rerr[m].max())
print("rows", nrows, "by evaluations", tot[1:6], "status mismatches", bad_status, "iteration mismatches", bad_iter)
print("worst rel J err by evaluations", worst[1:6], "worst |dr| px", worst[6])

"""The sampled parity gate (oracle/gate.py, behind bench.py's `parity` key) checked on the CPU: rows from the product's mathematics
compiled for the host pass it; a perturbed row, a shifted index or a wrong Huber factor fail it."""
import numpy as np

import hostcheck as hc
from kontiki_b200 import synthetic as syn
from oracle import gate, kto


def test_gate_accepts_correct_rows_and_rejects_wrong_ones():
    cfg = syn.make_config("H1", scale=0.002)
    c = cfg["cam"]
    rng = np.random.default_rng(3)
    bad = rng.random(len(c["lm_idx"])) < 0.2
    c["obs_uv"][bad] += rng.normal(0, 40, (int(bad.sum()), 2))
    sel = np.sort(rng.permutation(len(c["lm_idx"]))[:300])
    cam = kto.Camera(c["rows"], c["cols"], c["readout"], K=c["K"])
    h = hc.static_rs(cfg["knots"], cfg["dt"], 0.0, cam, c["obs_uv"][sel], c["obs_t0"][sel], c["ref_uv"][sel], c["ref_t0"][sel], c["lm_idx"][sel], c["rho"], c["weight"][sel],
                     huber_c=c["huber_c"][sel])
    idx = [h["i0_ref"], h["i0_obs"], None, None]
    res = gate.check_rows(cfg, "cam", sel, h["r"], h["J"], idx, c["rho"], robust=True)
    assert res["idx_exact"] and res["rel_J"] < gate.TOL and res["abs_r_cam_px"] < gate.CAM_R_TOL, res
    J2 = h["J"].copy(); J2[7, 60] += 1e-7 * np.abs(J2[7]).max()
    assert gate.check_rows(cfg, "cam", sel, h["r"], J2, idx, c["rho"], robust=True)["rel_J"] > gate.TOL
    i2 = [h["i0_ref"].copy(), h["i0_obs"], None, None]; i2[0][5] += 1
    assert not gate.check_rows(cfg, "cam", sel, h["r"], h["J"], i2, c["rho"], robust=True)["idx_exact"]
    assert gate.check_rows(cfg, "cam", sel, h["r"], h["J"], idx, c["rho"], robust=False)["rel_J"] > gate.TOL      # rows carry the corrector
    for which, name in ((0, "gyro"), (1, "accel")):
        m = cfg[name]
        s2 = np.arange(len(m["t"]))[::3]
        hi = hc.imu(which, cfg["knots"], cfg["dt"], 0.0, m["t"][s2], m["y"][s2], m["weight"][s2])
        res = gate.check_rows(cfg, name, s2, hi["r"], hi["J"], [hi["i0"], None, None, None])
        assert res["idx_exact"] and res["rel_r"] < gate.TOL and res["rel_J"] < gate.TOL, res
        r2 = hi["r"].copy(); r2[0, 0] += 1e-6
        assert gate.check_rows(cfg, name, s2, r2, hi["J"], [hi["i0"], None, None, None])["rel_r"] > gate.TOL


def test_gate_accepts_span_rows_and_rejects_wrong_ones():
    """NewtonRs / LiftingRs rows (span layout) through the same gate: host-compiled closed-form rows pass, perturbed ones fail."""
    cfg = syn.make_config("C3", scale=0.002)
    c = cfg["cam"]
    rng = np.random.default_rng(4)
    c["obs_uv"] += rng.normal(0, 1.0, c["obs_uv"].shape)
    bad = rng.random(len(c["lm_idx"])) < 0.2
    c["obs_uv"][bad] += rng.normal(0, 40, (int(bad.sum()), 2))
    sel = np.sort(rng.permutation(len(c["lm_idx"]))[:200])
    cam = kto.Camera(c["rows"], c["cols"], c["readout"], K=c["K"])
    args = (cfg["knots"], cfg["dt"], 0.0, cam, c["obs_uv"][sel], c["obs_t0"][sel], c["ref_uv"][sel], c["ref_t0"][sel], c["lm_idx"][sel], c["rho"])
    h = hc.newton_rs(*args, c["weight"][sel], huber_c=c["huber_c"][sel], fast=2)
    assert (h["status"] == 0).all() and (h["iterations"] >= 2).sum() > 10
    idx = [h["i0_ref"], h["i0_obs"], None, None]
    res = gate.check_rows(cfg, "cam", sel, h["r"], h["J"], idx, c["rho"], robust=True, method="newton")
    assert res["idx_exact"] and res["rel_J"] < gate.TOL and res["abs_r_cam_px"] < gate.CAM_R_TOL, res
    J2 = h["J"].copy(); J2[7, 60] += 1e-7 * np.abs(J2[7]).max()
    assert gate.check_rows(cfg, "cam", sel, h["r"], J2, idx, c["rho"], robust=True, method="newton")["rel_J"] > gate.TOL
    i2 = [h["i0_ref"], h["i0_obs"].copy(), None, None]; i2[1][5] += 1
    assert not gate.check_rows(cfg, "cam", sel, h["r"], h["J"], i2, c["rho"], robust=True, method="newton")["idx_exact"]
    vt = np.clip(c["obs_uv"][:, 1] / c["rows"] + rng.uniform(-0.2, 0.2, len(c["lm_idx"])), 0.0, 1.0)
    hl = hc.lifting_rs(*args, vt=vt[sel], w=c["weight"][sel])
    assert (hl["status"] == 0).all()
    idx = [hl["i0_ref"], hl["i0_obs"], None, None]
    res = gate.check_rows(cfg, "cam", sel, hl["r"], hl["J"], idx, c["rho"], method="lifting", vt=vt)
    assert res["idx_exact"] and res["rel_J"] < gate.TOL and res["abs_r_cam_px"] < gate.CAM_R_TOL, res
    r2 = hl["r"].copy(); r2[3, 2] += 1e-6
    assert gate.check_rows(cfg, "cam", sel, r2, hl["J"], idx, c["rho"], method="lifting", vt=vt)["abs_r_cam_px"] > gate.CAM_R_TOL

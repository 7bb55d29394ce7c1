"""BASELINE.json's configurations at FULL size through the C ABI (-m gpu): C2 (100k gyro + 100k accel), C4 (500k static-RS + 200k IMU) and
C5 (split trajectory, 10 000 knots, 1 M rows).  The oracle needs minutes for a million rows, so the gate is the sampled one bench.py
prints with every line (oracle/gate.py: 4096 random rows per group -- indices bit-exact, values to 1e-9) plus the size-independent
property that a row does not depend on the batch it is evaluated in (a random subset evaluated alone is bit-identical)."""
import numpy as np
import pytest

from kontiki_b200 import _lib, synthetic as syn
from oracle import gate

pytestmark = pytest.mark.gpu
FLAGS = _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS | _lib.EVAL_ROBUST


def build(cfg):
    p = _lib.Problem(0)
    if cfg.get("split"):
        p.set_split_spline(cfg["dt"], cfg["t0"], len(cfg["r3"]), cfg["dt"], cfg["t0"], len(cfg["so3"]))
    else:
        p.set_se3_spline(cfg["dt"], cfg["t0"], len(cfg["knots"]))
    g = {}
    imu = _lib.make_sensor()
    if cfg["gyro"]:
        g["gyro"] = p.add_gyroscope(imu, cfg["gyro"]["t"], cfg["gyro"]["y"], cfg["gyro"]["weight"])
    if cfg["accel"]:
        g["accel"] = p.add_accelerometer(imu, cfg["accel"]["t"], cfg["accel"]["y"], cfg["accel"]["weight"])
    if cfg["cam"]:
        c = cfg["cam"]
        g["cam"] = p.add_static_rs(_lib.make_camera(c["rows"], c["cols"], c["readout"], c["K"]), c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"],
                                   c["weight"], c["huber_c"])
    return p, g


def knots_of(cfg):
    return (cfg["r3"], cfg["so3"]) if cfg.get("split") else cfg["knots"]


def subset(cfg, name, sel):
    sub = dict(cfg, gyro=None, accel=None, cam=None)
    if name == "cam":
        c = cfg["cam"]
        sub["cam"] = dict(c, **{k: c[k][sel] for k in ("obs_uv", "obs_t0", "ref_uv", "ref_t0", "lm_idx", "weight", "huber_c")})
    else:
        sub[name] = {k: v[sel] for k, v in cfg[name].items()}
    return sub


@pytest.mark.parametrize("workload", ["C2", "C4", "C5"])
def test_full_size_sampled_parity_and_batch_independence(workload):
    cfg = syn.make_config(workload)
    if cfg["cam"]:      # gross outliers so that the Huber corrector's linear region is on the path
        rng = np.random.default_rng(17)
        bad = rng.random(len(cfg["cam"]["lm_idx"])) < 0.05
        cfg["cam"]["obs_uv"][bad] += rng.normal(0, 40, (int(bad.sum()), 2))
        # the observed row fixes the evaluation time: outside the image it leaves the residual's spans (the reference throws there too)
        cfg["cam"]["obs_uv"][:, 1] = np.clip(cfg["cam"]["obs_uv"][:, 1], 0.0, cfg["cam"]["rows"] - 1.0)
    rho = cfg["cam"]["rho"] if cfg["cam"] else None
    p, g = build(cfg)
    outs = p.evaluate(knots_of(cfg), rho, FLAGS)
    rng = np.random.default_rng(5)
    keys = ("i0", "i0_b", "i0_c", "i0_d")
    for name, gi in g.items():
        o = outs[gi]
        assert np.isfinite(o["r"]).all() and np.isfinite(o["J"]).all()
        sel = np.sort(rng.permutation(p.group_size(gi))[:4096])
        res = gate.check_rows(cfg, name, sel, o["r"][sel], o["J"][sel], [o[k][sel] if o.get(k) is not None else None for k in keys], rho, robust=True)
        assert res["idx_exact"], (workload, name)
        assert res["rel_r"] <= gate.TOL and res["rel_J"] <= gate.TOL and res["abs_r_cam_px"] <= gate.CAM_R_TOL, (workload, name, res)
        # the same rows evaluated alone, in shuffled order: bit-identical
        sh = rng.permutation(sel)
        p2, g2 = build(subset(cfg, name, sh))
        o2 = p2.evaluate(knots_of(cfg), rho, FLAGS)[g2[name]]
        for k in ("r", "J") + keys:
            if o.get(k) is not None and o2.get(k) is not None:
                assert np.array_equal(o2[k], o[k][sh]), (workload, name, k)

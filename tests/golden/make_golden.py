"""Regenerates tests/golden/oracle_v1.npz -- frozen input/output vectors of the CPU ORACLE (oracle/, the restated reference).

The reference itself cannot be built or imported here (it needs Ceres 1.x + Sophus @00f3fd91 + Eigen, oracle/README.md), so these are
NOT outputs of the reference: they freeze the oracle at the state in which it passed the reference's own property tests
(tests/test_oracle_pinning.py), so that any later change of the oracle's arithmetic is caught (tests/test_golden.py), and they give the
product a second, file-based target.  Inputs are the reference's own fixtures (python/tests/conftest.py:32-105, fixtures/camera_fixtures.py)
plus one seeded smooth trajectory.

    python tests/golden/make_golden.py          # rewrites oracle_v1.npz; commit the result
    python tests/golden/make_golden.py v2       # rewrites oracle_v2.npz (OrientationMeasurement rows)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import fixtures_ref as fx  # noqa: E402
from kontiki_b200 import synthetic as syn  # noqa: E402
from oracle import kto  # noqa: E402


def build():
    out = {}
    rng = np.random.default_rng(20261017)
    # --- IMU / position rows on the reference's SE3 fixture and on a smooth 40-knot trajectory
    for name, knots, dt, t0 in (("fix", fx.SE3_KNOTS, fx.SE3_DT, fx.SE3_T0), ("smooth", syn.smooth_se3_knots(40, 0.1), 0.1, 0.0)):
        n = 24
        t = np.sort(rng.uniform(t0 + 1e-6, t0 + (len(knots) - 3) * dt - 1e-6, n))
        y, w = rng.uniform(-1, 1, (n, 3)), rng.uniform(0.5, 2, n)
        out[f"se3_{name}_knots"], out[f"se3_{name}_meta"] = np.asarray(knots, float), np.array([dt, t0])
        out[f"se3_{name}_t"], out[f"se3_{name}_y"], out[f"se3_{name}_w"] = t, y, w
        for which, tag in ((0, "gyro"), (1, "accel"), (2, "pos")):
            for compat in ((False, True) if which == 1 else (False,)):
                o = kto.imu_residuals(kto.Traj(kto.SE3, dt, t0, knots, compat_zero_dB=compat), kto.Sensor(), which, t, y, w, jac_mode=2)
                sfx = f"se3_{name}_{tag}{'_compat' if compat else ''}"
                out[sfx + "_r"], out[sfx + "_J"], out[sfx + "_i0"] = o["r"], o["Ja"][:, :4], o["i0_a"]
    # --- split trajectory (reference R3 / SO3 fixtures)
    t = np.linspace(max(fx.R3_T0, fx.SO3_T0) + 1e-6, min(fx.R3_T0 + (len(fx.R3_KNOTS) - 3) * fx.R3_DT, fx.SO3_T0 + (len(fx.SO3_KNOTS) - 3) * fx.SO3_DT) - 1e-6, 16)
    y = rng.uniform(-1, 1, (16, 3))
    traj = kto.Traj(kto.SPLIT, fx.R3_DT, fx.R3_T0, fx.R3_KNOTS, fx.SO3_DT, fx.SO3_T0, fx.SO3_KNOTS)
    out["split_t"], out["split_y"] = t, y
    for which, tag in ((0, "gyro"), (1, "accel"), (2, "pos")):
        o = kto.imu_residuals(traj, kto.Sensor(), which, t, y, jac_mode=2)
        out[f"split_{tag}_r"], out[f"split_{tag}_Ja"], out[f"split_{tag}_Jb"] = o["r"], o["Ja"][:, :4], o["Jb"][:, :4]
        out[f"split_{tag}_i0a"], out[f"split_{tag}_i0b"] = o["i0_a"], o["i0_b"]
    # --- camera rows: static / Newton x pinhole / atan (camera_fixtures.py constants), relative pose set
    dt = 0.05
    knots = syn.smooth_se3_knots(80, dt)
    s = syn.make_static_rs(knots, dt, 8, obs_per_landmark=4, seed=77, noise_px=1.0)
    q_ct, p_ct = fx.so3_exp_xyzw(np.array([0.1, -0.2, 0.05])), np.array([0.05, -0.02, 0.1])
    out["cam_knots"], out["cam_dt"] = knots, np.array([dt])
    for k in ("obs_uv", "obs_t0", "ref_uv", "ref_t0", "lm_idx", "rho", "K"):
        out["cam_" + k] = np.asarray(s[k])
    out["cam_meta"] = np.array([s["rows"], s["cols"], s["readout"]], float)
    out["cam_q_ct"], out["cam_p_ct"] = q_ct, p_ct
    atan = dict(wc=(0.0029110778971412417, 0.0004189670467132041), gamma=0.8894355177968156)
    out["cam_atan"] = np.array([*atan["wc"], atan["gamma"]])
    for method in ("static", "newton"):
        for model in ("pinhole", "atan"):
            cam = kto.Camera(s["rows"], s["cols"], s["readout"], K=s["K"], method=method, q_ct=q_ct, p_ct=p_ct, **(atan if model == "atan" else {}))
            o = kto.static_rs_residuals(kto.Traj(kto.SE3, dt, 0.0, knots), cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"],
                                        jac_mode=2, cap=16)
            tag = f"cam_{method}_{model}"
            out[tag + "_r"], out[tag + "_ids"], out[tag + "_Ja"], out[tag + "_Jrho"] = o["r"], o["ids_a"], o["Ja"], o["Jrho"]
            out[tag + "_i0_ref"], out[tag + "_i0_obs"] = o["i0_ref_a"], o["i0_obs_a"]
    return out


def build_v2():
    """oracle_v2.npz: OrientationMeasurement and LiftingRsCameraMeasurement rows (added after v1 was frozen; v1 is not regenerated)."""
    out = {}
    rng = np.random.default_rng(20261018)

    def rotated(q):
        n = len(q)
        ang = rng.uniform(0.05, 2.5, n)
        ax = rng.normal(size=(n, 3)); ax /= np.linalg.norm(ax, axis=1)[:, None]
        dq = np.concatenate([ax * np.sin(ang / 2)[:, None], np.cos(ang / 2)[:, None]], axis=1)
        x1, y1, z1, w1 = q.T; x2, y2, z2, w2 = dq.T
        qm = np.stack([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
                       w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2], axis=1)
        qm[::2] *= -1.0
        return qm, ang

    for name, knots, dt, t0 in (("fix", fx.SE3_KNOTS, fx.SE3_DT, fx.SE3_T0), ("smooth", syn.smooth_se3_knots(40, 0.1), 0.1, 0.0)):
        n = 24
        traj = kto.Traj(kto.SE3, dt, t0, knots)
        t = np.sort(rng.uniform(t0 + 1e-6, t0 + (len(knots) - 3) * dt - 1e-6, n))
        qm, ang = rotated(kto.traj_evaluate(traj, t, 0xff)["orientation"])
        o = kto.imu_residuals(traj, kto.Sensor(), 3, t, qm, jac_mode=2)
        out[f"se3_{name}_knots"], out[f"se3_{name}_meta"] = np.asarray(knots, float), np.array([dt, t0])
        out[f"se3_{name}_t"], out[f"se3_{name}_q"], out[f"se3_{name}_angle"] = t, qm, ang
        out[f"se3_{name}_ori_r"], out[f"se3_{name}_ori_J"], out[f"se3_{name}_ori_i0"] = o["r"], o["Ja"][:, :4], o["i0_a"]
    t = np.linspace(max(fx.R3_T0, fx.SO3_T0) + 1e-6, min(fx.R3_T0 + (len(fx.R3_KNOTS) - 3) * fx.R3_DT, fx.SO3_T0 + (len(fx.SO3_KNOTS) - 3) * fx.SO3_DT) - 1e-6, 16)
    traj = kto.Traj(kto.SPLIT, fx.R3_DT, fx.R3_T0, fx.R3_KNOTS, fx.SO3_DT, fx.SO3_T0, fx.SO3_KNOTS)
    qm, ang = rotated(kto.traj_evaluate(traj, t, 0xff)["orientation"])
    o = kto.imu_residuals(traj, kto.Sensor(), 3, t, qm, jac_mode=2)
    out["split_t"], out["split_q"], out["split_angle"] = t, qm, ang
    out["split_ori_r"], out["split_ori_Jb"], out["split_ori_i0a"], out["split_ori_i0b"] = o["r"], o["Jb"][:, :4], o["ids_a"][:, 0], o["i0_b"]
    # --- LiftingRsCameraMeasurement rows on the camera case of v1 (same structure, relative pose set), at displaced row times
    dt = 0.05
    knots = syn.smooth_se3_knots(80, dt)
    s = syn.make_static_rs(knots, dt, 8, obs_per_landmark=4, seed=77, noise_px=1.0)
    q_ct, p_ct = fx.so3_exp_xyzw(np.array([0.1, -0.2, 0.05])), np.array([0.05, -0.02, 0.1])
    cam = kto.Camera(s["rows"], s["cols"], s["readout"], K=s["K"], q_ct=q_ct, p_ct=p_ct)
    vt = np.clip(np.asarray(s["obs_uv"])[:, 1] / s["rows"] + rng.uniform(-0.2, 0.2, len(s["lm_idx"])), 0.0, 1.0)
    o = kto.lifting_rs_residuals(kto.Traj(kto.SE3, dt, 0.0, knots), cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"], vt=vt,
                                 jac_mode=2, cap=16)
    out["lift_vt"] = vt
    out["lift_r"], out["lift_ids"], out["lift_Ja"], out["lift_Jvt"], out["lift_Jrho"], out["lift_i0_ref"] = o["r"], o["ids_a"], o["Ja"], o["Jvt"], o["Jrho"], o["i0_ref_a"]
    return out


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "v1"
    data = build_v2() if which == "v2" else build()
    path = os.path.join(HERE, f"oracle_{which}.npz")
    np.savez_compressed(path, **data)
    print(path, f"{os.path.getsize(path) / 1024:.0f} KiB, {len(data)} arrays")

"""Split (R3 + SO3) trajectory mathematics (kontiki_b200/csrc/split_math.cuh), compiled for the host, against the oracle."""
import numpy as np
import pytest

import fixtures_ref as fx
import hostcheck as hc
import parity
from kontiki_b200 import synthetic as syn
from oracle import kto


def _traj(dt_b=0.04, t0_b=0.01):
    k = syn.smooth_se3_knots(150, 0.05)
    vecs = k[:, 4:7].copy()
    quats = syn.smooth_se3_knots(190, dt_b)[:, :4].copy()
    return vecs, quats, kto.Traj(kto.SPLIT, 0.05, 0.0, vecs, dt_b, t0_b, quats), k


@pytest.mark.parametrize("which", [0, 1])
def test_split_imu_rows_match_oracle(which):
    vecs, quats, traj, _ = _traj()
    rng = np.random.default_rng(which)
    t, y, w = rng.uniform(0.02, 7.2, 300), rng.uniform(-1, 1, (300, 3)), rng.uniform(0.5, 2, 300)
    o = kto.imu_residuals(traj, kto.Sensor(), which, t, y, w, jac_mode=2)
    h = hc.imu_split(which, vecs, 0.05, 0.0, quats, 0.04, 0.01, t, y, w)
    assert (h["status"] == 0).all() and h["prepass_status"] == 0
    assert (h["i0_so3"] == o["i0_b"]).all()
    assert parity.rel_err(h["r"], o["r"]) < parity.TOL
    assert parity.rel_err(h["J"][:, -48:].reshape(-1, 4, 3, 4), o["Jb"][:, :4]) < parity.TOL
    if which == 1:
        assert (h["i0_r3"] == o["i0_a"]).all()
        assert parity.rel_err(h["J"][:, :36].reshape(-1, 4, 3, 3), o["Ja"][:, :4]) < parity.TOL


def test_split_reference_fixtures():
    """The reference's R3 / SO3 fixtures (python/tests/conftest.py:32-45, :52-67): constant-rate rotation => gyro == rate."""
    t = np.linspace(max(fx.R3_T0, fx.SO3_T0) + 1e-6, fx.SO3_T0 + (len(fx.SO3_KNOTS) - 3) * fx.SO3_DT - 1e-6, 50)
    h = hc.imu_split(0, fx.R3_KNOTS, fx.R3_DT, fx.R3_T0, fx.SO3_KNOTS, fx.SO3_DT, fx.SO3_T0, t, np.zeros((50, 3)))
    assert (h["status"] == 0).all()
    assert np.abs(-h["r"] - fx.SO3_RATE * fx.SO3_AXIS).max() < 1e-12     # body rate of a constant rotation about a fixed axis
    traj = kto.Traj(kto.SPLIT, fx.R3_DT, fx.R3_T0, fx.R3_KNOTS, fx.SO3_DT, fx.SO3_T0, fx.SO3_KNOTS)
    for which in (0, 1):
        o = kto.imu_residuals(traj, kto.Sensor(), which, t, np.zeros((50, 3)), jac_mode=2)
        h = hc.imu_split(which, fx.R3_KNOTS, fx.R3_DT, fx.R3_T0, fx.SO3_KNOTS, fx.SO3_DT, fx.SO3_T0, t, np.zeros((50, 3)))
        assert parity.rel_err(h["r"], o["r"]) < parity.TOL
        assert parity.rel_err(h["J"][:, -48:].reshape(-1, 4, 3, 4), o["Jb"][:, :4]) < parity.TOL


def test_split_camera_rows_match_oracle():
    vecs, quats, traj, k = _traj(dt_b=0.04, t0_b=0.0)
    s = syn.make_static_rs(k, 0.05, 40, obs_per_landmark=6, seed=5, noise_px=1.0)
    rng = np.random.default_rng(1)
    n = len(s["lm_idx"])
    out = rng.random(n) < 0.2
    s["obs_uv"][out] += rng.normal(0, 40, (out.sum(), 2))
    w = rng.uniform(0.5, 2, n)
    cam = kto.Camera(s["rows"], s["cols"], s["readout"], K=s["K"], q_ct=fx.so3_exp_xyzw(np.array([0.1, -0.2, 0.05])), p_ct=np.array([0.05, -0.02, 0.1]))
    o = kto.static_rs_residuals(traj, cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"], w, jac_mode=2, cap=24)
    h = hc.static_rs_split(vecs, 0.05, 0.0, quats, 0.04, 0.0, cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"], w)
    assert (h["status"] == 0).all()
    idx = h["idx"]
    assert (idx[:, 0] == o["i0_ref_a"]).all() and (idx[:, 1] == o["i0_obs_a"]).all()
    assert (idx[:, 2] == o["i0_ref_b"]).all() and (idx[:, 3] == o["i0_obs_b"]).all()
    assert np.abs(h["r"] - o["r"]).max() < parity.CAM_R_TOL
    Ja, Jb = np.zeros_like(o["Ja"]), np.zeros_like(o["Jb"])
    for i in range(n):
        pa = {int(kk): j for j, kk in enumerate(o["ids_a"][i]) if kk >= 0}
        pb = {int(kk): j for j, kk in enumerate(o["ids_b"][i]) if kk >= 0}
        J = h["J"][i]
        for k4 in range(4):
            Ja[i, pa[idx[i, 0] + k4]] += J[6 * k4:6 * k4 + 6].reshape(2, 3)
            Ja[i, pa[idx[i, 1] + k4]] += J[56 + 6 * k4:62 + 6 * k4].reshape(2, 3)
            Jb[i, pb[idx[i, 2] + k4]] += J[24 + 8 * k4:32 + 8 * k4].reshape(2, 4)
            Jb[i, pb[idx[i, 3] + k4]] += J[80 + 8 * k4:88 + 8 * k4].reshape(2, 4)
    assert parity.rel_err(Ja, o["Ja"]) < parity.TOL
    assert parity.rel_err(Jb, o["Jb"]) < parity.TOL
    assert parity.rel_err(h["J"][:, 112:114], o["Jrho"]) < parity.TOL


def test_split_non_unit_quaternion_flags_runtime_error():
    vecs, quats, traj, _ = _traj()
    q = quats.copy()
    q[10] *= 1.001
    _, _, _, st = hc.split_prepass(vecs, q)
    assert st == -2        # std::runtime_error in the reference (quaternion_math.h:19-23)


def test_split_position_rows_match_oracle():
    """PositionMeasurement (position_measurement.h:24-31) on the split trajectory: only the R3 spline is evaluated."""
    vecs, quats, traj, _ = _traj()
    rng = np.random.default_rng(5)
    t, y, w = rng.uniform(0.02, 7.2, 200), rng.uniform(-5, 5, (200, 3)), rng.uniform(0.5, 2, 200)
    o = kto.imu_residuals(traj, kto.Sensor(), 2, t, y, w, jac_mode=2)
    h = hc.imu_split(2, vecs, 0.05, 0.0, quats, 0.04, 0.01, t, y, w)
    assert (h["status"] == 0).all()
    assert (h["i0_r3"] == o["i0_a"]).all() and (h["i0_so3"] == o["ids_b"][:, 0]).all()
    assert parity.rel_err(h["r"], o["r"]) < parity.TOL
    assert parity.rel_err(h["J"].reshape(-1, 4, 3, 3), o["Ja"][:, :4]) < parity.TOL
    assert not o["Jb"].any()                                  # the SO3 blocks are structurally present but zero


def test_split_orientation_rows_match_oracle():
    """OrientationMeasurement on the split trajectory: only the SO3 spline is evaluated, one residual, J [4 SO3 knots][1][4]."""
    vecs, quats, traj, _ = _traj()
    rng = np.random.default_rng(6)
    t = rng.uniform(0.02, 7.2, 150)
    q = kto.traj_evaluate(traj, t, 0xff)["orientation"]
    ang = rng.uniform(0.05, 2.5, len(t))
    ax = rng.normal(size=(len(t), 3)); ax /= np.linalg.norm(ax, axis=1)[:, None]
    dq = np.concatenate([ax * np.sin(ang / 2)[:, None], np.cos(ang / 2)[:, None]], axis=1)
    x1, y1, z1, w1 = q.T; x2, y2, z2, w2 = dq.T
    qm = np.stack([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
                   w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2], axis=1)
    qm[::2] *= -1.0
    o = kto.imu_residuals(traj, kto.Sensor(), 3, t, qm, jac_mode=2)
    h = hc.imu_split(3, vecs, 0.05, 0.0, quats, 0.04, 0.01, t, qm)
    assert (h["status"] == 0).all() and (h["i0_so3"] == o["i0_b"]).all() and (h["i0_r3"] == o["ids_a"][:, 0]).all()
    assert np.abs(o["r"][:, 0] - ang).max() < 1e-9 and np.abs(h["r"] - o["r"]).max() < parity.TOL
    assert parity.rel_err(h["J"].reshape(-1, 4, 1, 4), o["Jb"][:, :4]) < parity.TOL
    assert not o["Ja"].any()


@pytest.mark.parametrize("lifting,atan,robust", [(False, False, False), (False, True, True), (True, False, True), (True, True, False)])
def test_split_span_camera_rows_match_oracle(lifting, atan, robust):
    """NewtonRsCameraMeasurement / LiftingRsCameraMeasurement on a SplitTrajectory (the reference instantiates every measurement with every
    trajectory, python/src/kontiki/measurements/measurement_defs.h:40-85): forward mode through the R3 spline and the reference's SO3
    evaluation (uniform_so3_spline_trajectory.h:46-125) on the hoisted pair logs, different grids for the two splines."""
    vecs, quats, traj, k = _traj(dt_b=0.04, t0_b=0.0)
    s = syn.make_static_rs(k, 0.05, 30, obs_per_landmark=5, seed=7, noise_px=1.5)
    rng = np.random.default_rng(2)
    n = len(s["lm_idx"])
    out = rng.random(n) < 0.2
    s["obs_uv"][out] += rng.normal(0, 40, (out.sum(), 2))
    s["obs_uv"][:, 1] = np.clip(s["obs_uv"][:, 1], 0, s["rows"] - 1e-6)
    w = rng.uniform(0.5, 2, n)
    kw = dict(wc=(0.02, -0.01), gamma=0.9) if atan else {}
    cam = kto.Camera(s["rows"], s["cols"], s["readout"], K=s["K"], method="static" if lifting else "newton",
                     q_ct=fx.so3_exp_xyzw(np.array([0.1, -0.2, 0.05])), p_ct=np.array([0.05, -0.02, 0.1]), **kw)
    args = (s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"])
    nres = 3 if lifting else 2
    if lifting:
        vt = np.clip(s["obs_uv"][:, 1] / s["rows"] + rng.uniform(-0.2, 0.2, n), 0.0, 1.0)
        o = kto.lifting_rs_residuals(traj, cam, *args, vt=vt, weight=w, jac_mode=2, cap=24)
        tail_o = np.concatenate([o["Jvt"], o["Jrho"]], axis=1)
    else:
        vt = None
        o = kto.static_rs_residuals(traj, cam, *args, w, jac_mode=2, cap=24)
        tail_o = o["Jrho"]
    c = np.full(n, 2.0) if robust else None
    cam.q_locked = cam.p_locked = False          # the knot columns do not depend on the lock state; the oracle then also returns the sensor blocks
    h = hc.span_rs_split(vecs, 0.05, 0.0, quats, 0.04, 0.0, cam, *args, lifting=lifting, vt=vt, w=w, huber_c=c, sensor=True)
    assert (h["status"] == 0).all()
    o2 = kto.lifting_rs_residuals(traj, cam, *args, vt=vt, weight=w, jac_mode=2, cap=24) if lifting else kto.static_rs_residuals(traj, cam, *args, w, jac_mode=2, cap=24)
    Jo = o2["Js"].copy()
    if robust:
        for i in range(n):
            cols = np.concatenate([Jo[i, :4 * nres].reshape(nres, 4), Jo[i, 4 * nres:7 * nres].reshape(nres, 3), Jo[i, 7 * nres:].reshape(nres, 1)], axis=1)
            _, _, J2 = kto.huber_correct(2.0, o2["r"][i], cols)
            Jo[i] = np.concatenate([J2[:, :4].reshape(-1), J2[:, 4:7].reshape(-1), J2[:, 7].reshape(-1)])
    assert np.abs(Jo[:, :7 * nres]).max() > 1.0 and parity.rel_err(h["Js"][:, None, :7 * nres], Jo[:, None, :7 * nres]) < parity.TOL
    idx = h["idx"]
    assert (idx[:, 0] == o["i0_ref_a"]).all() and (idx[:, 2] == o["i0_ref_b"]).all()
    Ja, Jb, tail = parity.scatter_span_split(h["J"], idx, o["ids_a"], o["ids_b"], h["Wa"], h["Wb"], nres)
    if not robust:
        assert np.abs(h["r"] - o["r"]).max() < parity.CAM_R_TOL
        assert parity.rel_err(Ja, o["Ja"]) < parity.TOL and parity.rel_err(Jb, o["Jb"]) < parity.TOL and parity.rel_err(tail, tail_o) < parity.TOL
        return
    n_out = 0
    for i in range(n):
        ma, mb = int((o["ids_a"][i] >= 0).sum()), int((o["ids_b"][i] >= 0).sum())
        Jfull = np.concatenate([o["Ja"][i, kk] for kk in range(ma)] + [o["Jb"][i, kk] for kk in range(mb)] + [tail_o[i].reshape(-1, nres).T], axis=1)
        _, r2, J2 = kto.huber_correct(2.0, o["r"][i], Jfull)
        Jmine = np.concatenate([Ja[i, kk] for kk in range(ma)] + [Jb[i, kk] for kk in range(mb)] + [tail[i].reshape(-1, nres).T], axis=1)
        assert np.abs(Jmine - J2).max() <= parity.TOL * np.abs(J2).max()
        assert np.abs(h["r"][i] - r2).max() <= parity.CAM_R_TOL
        n_out += float(np.dot(o["r"][i], o["r"][i])) > 4.0
    assert n_out > 3

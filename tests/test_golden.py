"""Committed golden vectors (tests/golden/oracle_v1.npz, made by tests/golden/make_golden.py from the CPU oracle):
  * the oracle still reproduces them (indices exactly, values to 1e-13 -- libm may differ in the last bit between host CPUs): its arithmetic
    has not drifted since it passed the reference's property tests;
  * the product's mathematics, compiled for the host, matches them to the parity tolerance (CPU suite);
  * the CUDA path through the C ABI matches them (-m gpu).
They are outputs of the RESTATED reference: the reference itself cannot be built here (oracle/README.md)."""
import os

import numpy as np
import pytest

import hostcheck as hc
import parity
from kontiki_b200 import _lib
from oracle import kto

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_v1.npz"))
ATAN = dict(wc=tuple(G["cam_atan"][:2]), gamma=float(G["cam_atan"][2]))
IMU_CASES = [(name, which, tag, compat) for name in ("fix", "smooth") for which, tag, compat in
             ((0, "gyro", False), (1, "accel", False), (1, "accel_compat", True), (2, "pos", False))]
CAM_CASES = [(m, c) for m in ("static", "newton") for c in ("pinhole", "atan")]


def _oracle_cam(method, model):
    rows, cols, readout = G["cam_meta"]
    cam = kto.Camera(int(rows), int(cols), float(readout), K=G["cam_K"], method=method, q_ct=G["cam_q_ct"], p_ct=G["cam_p_ct"], **(ATAN if model == "atan" else {}))
    return cam


def _same(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() <= 1e-13 * max(1.0, np.abs(np.asarray(b)).max())


def test_oracle_reproduces_the_golden_vectors():
    for name, which, tag, compat in IMU_CASES:
        dt, t0 = G[f"se3_{name}_meta"]
        o = kto.imu_residuals(kto.Traj(kto.SE3, dt, t0, G[f"se3_{name}_knots"], compat_zero_dB=compat), kto.Sensor(), which, G[f"se3_{name}_t"],
                              G[f"se3_{name}_y"], G[f"se3_{name}_w"], jac_mode=2)
        assert _same(o["r"], G[f"se3_{name}_{tag}_r"]) and _same(o["Ja"][:, :4], G[f"se3_{name}_{tag}_J"])
        assert np.array_equal(o["i0_a"], G[f"se3_{name}_{tag}_i0"])
    for method, model in CAM_CASES:
        o = kto.static_rs_residuals(kto.Traj(kto.SE3, float(G["cam_dt"][0]), 0.0, G["cam_knots"]), _oracle_cam(method, model), G["cam_obs_uv"], G["cam_obs_t0"],
                                    G["cam_ref_uv"], G["cam_ref_t0"], G["cam_lm_idx"], G["cam_rho"], jac_mode=2, cap=16)
        tag = f"cam_{method}_{model}"
        assert _same(o["r"], G[tag + "_r"]) and _same(o["Ja"], G[tag + "_Ja"]) and _same(o["Jrho"], G[tag + "_Jrho"])
        assert np.array_equal(o["ids_a"], G[tag + "_ids"]) and np.array_equal(o["i0_obs_a"], G[tag + "_i0_obs"])


@pytest.mark.parametrize("name,which,tag,compat", IMU_CASES)
def test_host_math_matches_golden_imu(name, which, tag, compat):
    dt, t0 = G[f"se3_{name}_meta"]
    h = hc.imu(which, G[f"se3_{name}_knots"], dt, t0, G[f"se3_{name}_t"], G[f"se3_{name}_y"], G[f"se3_{name}_w"], compat=compat)
    assert (h["i0"] == G[f"se3_{name}_{tag}_i0"]).all()
    assert parity.rel_err(h["r"], G[f"se3_{name}_{tag}_r"]) < parity.TOL and parity.rel_err(h["J"], G[f"se3_{name}_{tag}_J"]) < parity.TOL


@pytest.mark.parametrize("method,model", CAM_CASES)
def test_host_math_matches_golden_camera(method, model):
    cam = _oracle_cam(method, model)
    fn = hc.newton_rs if method == "newton" else hc.static_rs
    h = fn(G["cam_knots"], float(G["cam_dt"][0]), 0.0, cam, G["cam_obs_uv"], G["cam_obs_t0"], G["cam_ref_uv"], G["cam_ref_t0"], G["cam_lm_idx"], G["cam_rho"])
    tag = f"cam_{method}_{model}"
    assert (h["status"] == 0).all() and (h["i0_ref"] == G[tag + "_i0_ref"]).all() and (h["i0_obs"] == G[tag + "_i0_obs"]).all()
    assert np.abs(h["r"] - G[tag + "_r"]).max() < parity.CAM_R_TOL
    Js, Jrho = parity.scatter_cam(h["J"], h["i0_ref"], h["i0_obs"], G[tag + "_ids"], h.get("W", 4))
    assert parity.rel_err(Js, G[tag + "_Ja"]) < parity.TOL and parity.rel_err(Jrho, G[tag + "_Jrho"]) < parity.TOL


@pytest.mark.parametrize("tag,which", [("gyro", 0), ("accel", 1), ("pos", 2)])
def test_host_math_matches_golden_split(tag, which):
    import fixtures_ref as fx
    h = hc.imu_split(which, fx.R3_KNOTS, fx.R3_DT, fx.R3_T0, fx.SO3_KNOTS, fx.SO3_DT, fx.SO3_T0, G["split_t"], G["split_y"])
    assert parity.rel_err(h["r"], G[f"split_{tag}_r"]) < parity.TOL
    if which != 2:
        assert parity.rel_err(h["J"][:, -48:].reshape(-1, 4, 3, 4), G[f"split_{tag}_Jb"]) < parity.TOL
    if which != 0:
        assert parity.rel_err(h["J"][:, :36].reshape(-1, 4, 3, 3), G[f"split_{tag}_Ja"]) < parity.TOL


@pytest.mark.gpu
def test_cuda_path_matches_golden():
    for name, which, tag, compat in IMU_CASES:
        dt, t0 = G[f"se3_{name}_meta"]
        p = _lib.Problem(0)
        p.set_se3_spline(dt, t0, len(G[f"se3_{name}_knots"]), compat_zero_dB=compat)
        args = (G[f"se3_{name}_t"], G[f"se3_{name}_y"], G[f"se3_{name}_w"])
        g = p.add_position(*args) if which == 2 else (p.add_gyroscope if which == 0 else p.add_accelerometer)(_lib.make_sensor(), *args)
        o = p.evaluate(G[f"se3_{name}_knots"])[g]
        assert (o["i0"] == G[f"se3_{name}_{tag}_i0"]).all()
        assert parity.rel_err(o["r"], G[f"se3_{name}_{tag}_r"]) < parity.TOL and parity.rel_err(o["J"].reshape(-1, 4, 3, 7), G[f"se3_{name}_{tag}_J"]) < parity.TOL
    rows, cols, readout = G["cam_meta"]
    for method, model in CAM_CASES:
        p = _lib.Problem(0)
        p.set_se3_spline(float(G["cam_dt"][0]), 0.0, len(G["cam_knots"]))
        cam = _lib.make_camera(int(rows), int(cols), float(readout), G["cam_K"], q_ct=G["cam_q_ct"], p_ct=G["cam_p_ct"], **(ATAN if model == "atan" else {}))
        add = p.add_newton_rs if method == "newton" else p.add_static_rs
        g = add(cam, G["cam_obs_uv"], G["cam_obs_t0"], G["cam_ref_uv"], G["cam_ref_t0"], G["cam_lm_idx"])
        o = p.evaluate(G["cam_knots"], G["cam_rho"], _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS)[g]
        tag = f"cam_{method}_{model}"
        assert (o["i0"] == G[tag + "_i0_ref"]).all() and (o["i0_b"] == G[tag + "_i0_obs"]).all()          # bit-exact indexing
        ids, _ = p.get_structure(g, cap=16)
        assert (ids == G[tag + "_ids"]).all()
        Js = p.expand_static_rs(g, ids, o["J"], o["i0"], o["i0_b"])
        assert np.abs(o["r"] - G[tag + "_r"]).max() < parity.CAM_R_TOL
        assert parity.rel_err(Js, G[tag + "_Ja"]) < parity.TOL and parity.rel_err(o["J"][:, -2:], G[tag + "_Jrho"]) < parity.TOL


# ---- oracle_v2.npz: OrientationMeasurement rows (tests/golden/make_golden.py v2) ---------------------------------------------------------
G2 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_v2.npz"))


def _split_traj():
    import fixtures_ref as fx
    return fx, kto.Traj(kto.SPLIT, fx.R3_DT, fx.R3_T0, fx.R3_KNOTS, fx.SO3_DT, fx.SO3_T0, fx.SO3_KNOTS)


def test_oracle_reproduces_the_golden_orientation_vectors():
    for name in ("fix", "smooth"):
        dt, t0 = G2[f"se3_{name}_meta"]
        o = kto.imu_residuals(kto.Traj(kto.SE3, dt, t0, G2[f"se3_{name}_knots"]), kto.Sensor(), 3, G2[f"se3_{name}_t"], G2[f"se3_{name}_q"], jac_mode=2)
        assert _same(o["r"], G2[f"se3_{name}_ori_r"]) and _same(o["Ja"][:, :4], G2[f"se3_{name}_ori_J"]) and np.array_equal(o["i0_a"], G2[f"se3_{name}_ori_i0"])
        assert np.abs(o["r"][:, 0] - G2[f"se3_{name}_angle"]).max() < 1e-9        # the residual IS the angle the vectors were rotated by
    fx, traj = _split_traj()
    o = kto.imu_residuals(traj, kto.Sensor(), 3, G2["split_t"], G2["split_q"], jac_mode=2)
    assert _same(o["r"], G2["split_ori_r"]) and _same(o["Jb"][:, :4], G2["split_ori_Jb"]) and np.array_equal(o["i0_b"], G2["split_ori_i0b"])


def test_host_math_matches_golden_orientation():
    for name in ("fix", "smooth"):
        dt, t0 = G2[f"se3_{name}_meta"]
        h = hc.imu(3, G2[f"se3_{name}_knots"], dt, t0, G2[f"se3_{name}_t"], G2[f"se3_{name}_q"])
        assert (h["i0"] == G2[f"se3_{name}_ori_i0"]).all()
        assert np.abs(h["r"] - G2[f"se3_{name}_ori_r"]).max() < parity.TOL and parity.rel_err(h["J"], G2[f"se3_{name}_ori_J"]) < parity.TOL
    fx, _ = _split_traj()
    h = hc.imu_split(3, fx.R3_KNOTS, fx.R3_DT, fx.R3_T0, fx.SO3_KNOTS, fx.SO3_DT, fx.SO3_T0, G2["split_t"], G2["split_q"])
    assert (h["i0_so3"] == G2["split_ori_i0b"]).all() and (h["i0_r3"] == G2["split_ori_i0a"]).all()
    assert np.abs(h["r"] - G2["split_ori_r"]).max() < parity.TOL and parity.rel_err(h["J"].reshape(-1, 4, 1, 4), G2["split_ori_Jb"]) < parity.TOL


@pytest.mark.gpu
def test_cuda_path_matches_golden_orientation():
    for name in ("fix", "smooth"):
        dt, t0 = G2[f"se3_{name}_meta"]
        p = _lib.Problem(0)
        p.set_se3_spline(dt, t0, len(G2[f"se3_{name}_knots"]))
        g = p.add_orientation(G2[f"se3_{name}_t"], G2[f"se3_{name}_q"])
        o = p.evaluate(G2[f"se3_{name}_knots"])[g]
        assert (o["i0"] == G2[f"se3_{name}_ori_i0"]).all()
        assert np.abs(o["r"] - G2[f"se3_{name}_ori_r"]).max() < parity.TOL and parity.rel_err(o["J"].reshape(-1, 4, 1, 7), G2[f"se3_{name}_ori_J"]) < parity.TOL
    fx, _ = _split_traj()
    p = _lib.Problem(0)
    p.set_split_spline(fx.R3_DT, fx.R3_T0, len(fx.R3_KNOTS), fx.SO3_DT, fx.SO3_T0, len(fx.SO3_KNOTS))
    g = p.add_orientation(G2["split_t"], G2["split_q"])
    o = p.evaluate((np.asarray(fx.R3_KNOTS, float), np.asarray(fx.SO3_KNOTS, float)))[g]
    assert (o["i0_c"] == G2["split_ori_i0b"]).all() and (o["i0"] == G2["split_ori_i0a"]).all()
    assert np.abs(o["r"] - G2["split_ori_r"]).max() < parity.TOL and parity.rel_err(o["J"].reshape(-1, 4, 1, 4), G2["split_ori_Jb"]) < parity.TOL


# ---- oracle_v2.npz: LiftingRsCameraMeasurement rows on the v1 camera case (structure arrays from oracle_v1.npz) ------------------------------
def _lifting_args():
    return (G["cam_obs_uv"], G["cam_obs_t0"], G["cam_ref_uv"], G["cam_ref_t0"], G["cam_lm_idx"], G["cam_rho"])


def test_oracle_reproduces_the_golden_lifting_vectors():
    cam = _oracle_cam("static", "pinhole")
    o = kto.lifting_rs_residuals(kto.Traj(kto.SE3, float(G["cam_dt"][0]), 0.0, G["cam_knots"]), cam, *_lifting_args(), vt=G2["lift_vt"], jac_mode=2, cap=16)
    assert _same(o["r"], G2["lift_r"]) and _same(o["Ja"], G2["lift_Ja"]) and _same(o["Jvt"], G2["lift_Jvt"]) and _same(o["Jrho"], G2["lift_Jrho"])
    assert np.array_equal(o["ids_a"], G2["lift_ids"]) and np.array_equal(o["i0_ref_a"], G2["lift_i0_ref"])


def test_host_math_matches_golden_lifting():
    cam = _oracle_cam("static", "pinhole")
    h = hc.lifting_rs(G["cam_knots"], float(G["cam_dt"][0]), 0.0, cam, *_lifting_args(), vt=G2["lift_vt"])
    assert (h["status"] == 0).all() and (h["i0_ref"] == G2["lift_i0_ref"]).all()
    Js, Jvt, Jrho = parity.scatter_lifting(h["J"], h["i0_ref"], h["i0_obs"], G2["lift_ids"], h["W"])
    assert np.abs(h["r"] - G2["lift_r"]).max() < parity.CAM_R_TOL
    assert parity.rel_err(Js, G2["lift_Ja"]) < parity.TOL and parity.rel_err(Jvt, G2["lift_Jvt"]) < parity.TOL and parity.rel_err(Jrho, G2["lift_Jrho"]) < parity.TOL


@pytest.mark.gpu
def test_cuda_path_matches_golden_lifting():
    rows, cols, readout = G["cam_meta"]
    p = _lib.Problem(0)
    p.set_se3_spline(float(G["cam_dt"][0]), 0.0, len(G["cam_knots"]))
    cam = _lib.make_camera(int(rows), int(cols), float(readout), G["cam_K"], q_ct=G["cam_q_ct"], p_ct=G["cam_p_ct"])
    g = p.add_lifting_rs(cam, *_lifting_args()[:5])
    p.set_group_vt(g, G2["lift_vt"])
    o = p.evaluate(G["cam_knots"], G["cam_rho"], _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS)[g]
    assert (o["i0"] == G2["lift_i0_ref"]).all()
    ids, _ = p.get_structure(g, cap=16)
    assert (ids == G2["lift_ids"]).all()
    Js = p.expand_static_rs(g, ids, o["J"], o["i0"], o["i0_b"])
    assert np.abs(o["r"] - G2["lift_r"]).max() < parity.CAM_R_TOL
    assert parity.rel_err(Js, G2["lift_Ja"]) < parity.TOL and parity.rel_err(o["J"][:, -6:-3], G2["lift_Jvt"]) < parity.TOL
    assert parity.rel_err(o["J"][:, -3:], G2["lift_Jrho"]) < parity.TOL

"""The product's __host__ __device__ mathematics (kontiki_b200/csrc/spline_math.cuh -- the text the CUDA kernels execute),
compiled for the host by tests/hostcheck.py, against the CPU oracle.  Runs without a GPU; the same comparisons are
repeated through the C ABI on the device in test_gpu_parity.py."""
import numpy as np
import pytest

import fixtures_ref as fx
import hostcheck as hc
import parity
from kontiki_b200 import synthetic as syn
from oracle import kto


@pytest.mark.parametrize("which", [0, 1])
@pytest.mark.parametrize("compat", [False, True])
def test_imu_rows_match_oracle(which, compat):
    knots = fx.smooth_se3_knots(60, 0.1)
    rng = np.random.default_rng(which)
    t = rng.uniform(0.0, 5.69, 300)
    y, w = rng.uniform(-1, 1, (300, 3)), rng.uniform(0.5, 2, 300)
    o = parity.oracle_imu(kto.Traj(kto.SE3, 0.1, 0.0, knots, compat_zero_dB=compat), which, t, y, w)
    h = hc.imu(which, knots, 0.1, 0.0, t, y, w, compat=compat)
    assert (h["status"] == 0).all()
    assert (h["i0"] == o["i0"]).all()
    assert parity.rel_err(h["r"], o["r"]) < parity.TOL
    assert parity.rel_err(h["J"], o["J"]) < parity.TOL


@pytest.mark.parametrize("which", [0, 1])
def test_imu_rows_reference_fixture(which):
    """The reference's own SE3 fixture (python/tests/conftest.py:83-105): large relative rotations between knots."""
    t = np.linspace(fx.SE3_T0, fx.SE3_T0 + 3 * fx.SE3_DT - 1e-9, 64)
    y = np.random.default_rng(3).uniform(-1, 1, (64, 3))
    o = parity.oracle_imu(kto.Traj(kto.SE3, fx.SE3_DT, fx.SE3_T0, fx.SE3_KNOTS), which, t, y)
    h = hc.imu(which, fx.SE3_KNOTS, fx.SE3_DT, fx.SE3_T0, t, y)
    assert (h["i0"] == o["i0"]).all()
    assert parity.rel_err(h["r"], o["r"]) < parity.TOL
    assert parity.rel_err(h["J"], o["J"]) < parity.TOL


def test_imu_out_of_range_status():
    knots = fx.smooth_se3_knots(20, 0.1)
    h = hc.imu(0, knots, 0.1, 0.0, [-0.01, 1.71, 1.6999], np.zeros((3, 3)))
    assert list(h["status"]) == [-1, -1, 0]
    with pytest.raises(kto.OracleError):
        kto.imu_residuals(kto.Traj(kto.SE3, 0.1, 0.0, knots), kto.Sensor(), 0, [1.71], np.zeros((1, 3)))


def _camera_case(dt, seed, q_ct=None, p_ct=None, n_lm=40):
    knots = syn.smooth_se3_knots(150, dt)
    s = syn.make_static_rs(knots, dt, n_lm, obs_per_landmark=6, seed=seed, noise_px=1.0)
    rng = np.random.default_rng(seed)
    out = rng.random(len(s["lm_idx"])) < 0.2           # 20 % gross outliers: beyond the Huber threshold
    s["obs_uv"][out] += rng.normal(0, 40, (out.sum(), 2))
    s["weight"] = rng.uniform(0.5, 2, len(s["lm_idx"]))
    cam = kto.Camera(s["rows"], s["cols"], s["readout"], K=s["K"], q_ct=(0, 0, 0, 1) if q_ct is None else q_ct, p_ct=(0, 0, 0) if p_ct is None else p_ct)
    return knots, s, cam


@pytest.mark.parametrize("dt", [0.02, 0.1])
@pytest.mark.parametrize("rel_pose", [False, True])
def test_static_rs_rows_match_oracle(dt, rel_pose):
    q_ct = fx.so3_exp_xyzw(np.array([0.1, -0.2, 0.05])) if rel_pose else None
    p_ct = np.array([0.05, -0.02, 0.1]) if rel_pose else None
    knots, s, cam = _camera_case(dt, 5, q_ct, p_ct)
    o = kto.static_rs_residuals(kto.Traj(kto.SE3, dt, 0.0, knots), cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"],
                                s["weight"], jac_mode=2, cap=24)
    h = hc.static_rs(knots, dt, 0.0, cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"], s["weight"])
    assert (h["status"] == 0).all()
    assert (h["i0_ref"] == o["i0_ref_a"]).all() and (h["i0_obs"] == o["i0_obs_a"]).all()
    # residuals are differences of ~1e3 px coordinates: tolerance relative to the pixel scale, Jacobians to their block norm
    assert np.abs(h["r"] - o["r"]).max() < parity.CAM_R_TOL
    Js, Jrho = parity.scatter_cam(h["J"], h["i0_ref"], h["i0_obs"], o["ids_a"])
    assert parity.rel_err(Js, o["Ja"]) < parity.TOL
    assert parity.rel_err(Jrho, o["Jrho"]) < parity.TOL


def test_static_rs_huber_corrector_matches_oracle():
    dt = 0.05
    knots, s, cam = _camera_case(dt, 9)
    n = len(s["lm_idx"])
    c = np.full(n, 5.0)
    o = kto.static_rs_residuals(kto.Traj(kto.SE3, dt, 0.0, knots), cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"],
                                s["weight"], jac_mode=2, cap=24)
    h = hc.static_rs(knots, dt, 0.0, cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"], s["weight"], huber_c=c)
    n_out = 0
    for i in range(n):
        m = int((o["ids_a"][i] >= 0).sum())
        Jfull = np.concatenate([o["Ja"][i, k] for k in range(m)] + [o["Jrho"][i].reshape(2, 1)], axis=1)
        _, r2, J2 = kto.huber_correct(5.0, o["r"][i], Jfull)
        Js, Jr = parity.scatter_cam(h["J"][i:i + 1], h["i0_ref"][i:i + 1], h["i0_obs"][i:i + 1], o["ids_a"][i:i + 1])
        Jmine = np.concatenate([Js[0, k] for k in range(m)] + [Jr[0].reshape(2, 1)], axis=1)
        assert np.abs(Jmine - J2).max() <= parity.TOL * np.abs(J2).max()
        assert np.abs(h["r"][i] - r2).max() <= parity.CAM_R_TOL
        n_out += np.linalg.norm(o["r"][i]) > 5.0
    assert 5 < n_out < n - 5          # both branches of the loss exercised


def test_pair_prepass_small_angle_branch():
    """Identical adjacent knots: log hits Sophus' small-angle branches; values and derivatives must stay finite."""
    k = np.tile(np.array([0.1, -0.2, 0.3, 0.0, 1.0, 2.0, 3.0]), (6, 1))
    k[:, 3] = np.sqrt(1 - (k[0, :3] ** 2).sum())
    k8, pairs = hc.prepass(k)
    assert np.isfinite(pairs).all()
    assert np.abs(pairs[1:, :6]).max() < 1e-15


# ---- sensor-block Jacobians (kontiki_b200/csrc/sensor_jac.cuh) against the oracle's autodiff over the sensor blocks ---------
@pytest.mark.parametrize("which", [0, 1])
@pytest.mark.parametrize("offset", [0.0, 0.013])
def test_imu_time_offset_jacobian_matches_oracle(which, offset):
    knots = syn.smooth_se3_knots(60, 0.1)
    rng = np.random.default_rng(4)
    t, y, w = rng.uniform(0.2, 5.4, 150), rng.uniform(-1, 1, (150, 3)), rng.uniform(0.5, 2, 150)
    o = kto.imu_residuals(kto.Traj(kto.SE3, 0.1, 0.0, knots), kto.Sensor(time_offset=offset), which, t, y, w, jac_mode=2, raise_on_error=False)
    out, st = hc.imu_time_offset_se3(which, knots, 0.1, 0.0, t, w, time_offset=offset)
    assert ((o["status"] == 0) == (st == 0)).all()          # a locked non-zero offset that crosses a knot throws in both (SURVEY 8b edge case i)
    ok = st == 0
    assert ok.sum() > 100
    assert parity.rel_err(out[ok], o["Js"][ok, 21:24]) < parity.TOL
    assert np.abs(o["Js"][ok, :21]).max() == 0.0            # IMU relative pose is not applied by the reference (TODO.md:6)


@pytest.mark.parametrize("which", [0, 1])
def test_imu_time_offset_jacobian_split_matches_oracle(which):
    k = syn.smooth_se3_knots(80, 0.05)
    vecs, qb = k[:, 4:7].copy(), syn.smooth_se3_knots(100, 0.04)[:, :4].copy()
    rng = np.random.default_rng(5)
    t, y, w = rng.uniform(0.05, 3.7, 150), rng.uniform(-1, 1, (150, 3)), rng.uniform(0.5, 2, 150)
    o = kto.imu_residuals(kto.Traj(kto.SPLIT, 0.05, 0.0, vecs, 0.04, 0.01, qb), kto.Sensor(), which, t, y, w, jac_mode=2)
    out, st = hc.imu_time_offset_split(which, vecs, 0.05, 0.0, qb, 0.04, 0.01, t, w)
    assert (st == 0).all()
    assert parity.rel_err(out, o["Js"][:, 21:24]) < parity.TOL


@pytest.mark.parametrize("robust", [False, True])
def test_camera_sensor_jacobians_match_oracle(robust):
    dt = 0.05
    knots, s, _ = _camera_case(dt, 5)
    cam = kto.Camera(s["rows"], s["cols"], s["readout"], K=s["K"], q_ct=fx.so3_exp_xyzw(np.array([0.1, -0.2, 0.05])), p_ct=np.array([0.05, -0.02, 0.1]),
                     time_offset=0.004)
    n = len(s["lm_idx"])
    o = kto.static_rs_residuals(kto.Traj(kto.SE3, dt, 0.0, knots), cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"],
                                s["weight"], jac_mode=2, cap=24)
    out, st = hc.static_rs_sensor_se3(knots, dt, 0.0, cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"], s["weight"],
                                      huber_c=np.full(n, 5.0) if robust else None)
    assert (st == 0).all()
    Js = o["Js"].copy()
    if robust:
        for i in range(n):
            _, _, J2 = kto.huber_correct(5.0, o["r"][i], np.concatenate([Js[i, 0:8].reshape(2, 4), Js[i, 8:14].reshape(2, 3), Js[i, 14:16].reshape(2, 1)], 1))
            Js[i, 0:8], Js[i, 8:14], Js[i, 14:16] = J2[:, 0:4].reshape(-1), J2[:, 4:7].reshape(-1), J2[:, 7]
    for a, b in ((0, 8), (8, 14), (14, 16)):
        assert parity.rel_err(out[:, a:b], Js[:, a:b]) < parity.TOL


@pytest.mark.parametrize("robust", [False, True])
def test_camera_sensor_jacobians_split_match_oracle(robust):
    """SURVEY.md 8f-2 remainder: the camera's q_ct / p_ct / time_offset columns on a SPLIT trajectory (sensors.h:135-165, split_trajectory.h:41-58)."""
    dt = 0.05
    knots, s, _ = _camera_case(dt, 7)
    vecs, quats = knots[:, 4:7].copy(), knots[:, :4].copy()
    cam = kto.Camera(s["rows"], s["cols"], s["readout"], K=s["K"], q_ct=fx.so3_exp_xyzw(np.array([0.1, -0.2, 0.05])), p_ct=np.array([0.05, -0.02, 0.1]),
                     time_offset=0.004)
    n = len(s["lm_idx"])
    o = kto.static_rs_residuals(kto.Traj(kto.SPLIT, dt, 0.0, vecs, dt, 0.0, quats), cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"],
                                s["weight"], jac_mode=2, cap=24)
    out, st = hc.static_rs_sensor_split(vecs, dt, 0.0, quats, dt, 0.0, cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"], s["weight"],
                                        huber_c=np.full(n, 5.0) if robust else None)
    assert (st == 0).all()
    Js = o["Js"].copy()
    if robust:
        for i in range(n):
            _, _, J2 = kto.huber_correct(5.0, o["r"][i], np.concatenate([Js[i, 0:8].reshape(2, 4), Js[i, 8:14].reshape(2, 3), Js[i, 14:16].reshape(2, 1)], 1))
            Js[i, 0:8], Js[i, 8:14], Js[i, 14:16] = J2[:, 0:4].reshape(-1), J2[:, 4:7].reshape(-1), J2[:, 7]
    for a, b in ((0, 8), (8, 14), (14, 16)):
        assert parity.rel_err(out[:, a:b], Js[:, a:b]) < parity.TOL


def test_se3_evaluate_matrices_match_oracle():
    """UniformSE3SplineTrajectory.evaluate(t) -> (P, P', P'') (py_uniform_se3_spline_trajectory.cc:53-60) on the reference's fixture."""
    t = np.linspace(fx.SE3_T0, fx.SE3_T0 + 3 * fx.SE3_DT - 1e-9, 33)
    P, Pp, Pb = kto.se3_evaluate_matrices(kto.Traj(kto.SE3, fx.SE3_DT, fx.SE3_T0, fx.SE3_KNOTS), t)
    out, st = hc.se3_matrices(fx.SE3_KNOTS, fx.SE3_DT, fx.SE3_T0, t)
    assert (st == 0).all()
    for a, b in ((out[:, 0], P), (out[:, 1], Pp), (out[:, 2], Pb)):
        assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(b).max())


def test_traj_point_queries_match_oracle():
    Pf, V, A, Q, W = 1, 2, 4, 8, 16
    traj = kto.Traj(kto.SE3, fx.SE3_DT, fx.SE3_T0, fx.SE3_KNOTS)
    t = np.linspace(traj.min_time, traj.max_time - 1e-6, 40)
    o = kto.traj_evaluate(traj, t, Pf | V | A | Q | W)
    out, st = hc.traj_eval_se3(fx.SE3_KNOTS, fx.SE3_DT, fx.SE3_T0, t)
    assert (st == 0).all()
    q = out[:, 9:13] * np.sign((out[:, 9:13] * o["orientation"]).sum(1))[:, None]
    for a, b in ((out[:, 0:3], o["position"]), (out[:, 3:6], o["velocity"]), (out[:, 6:9], o["acceleration"]), (q, o["orientation"]),
                 (out[:, 13:16], o["angular_velocity"])):
        assert np.abs(a - b).max() < 1e-12 * max(1.0, np.abs(b).max())
    # compat_zero_dB: what the reference's Jet path gives for acceleration(t) alone (flags = EvalAcceleration, dB unassigned)
    oc = kto.traj_evaluate(kto.Traj(kto.SE3, fx.SE3_DT, fx.SE3_T0, fx.SE3_KNOTS, compat_zero_dB=True), t, A)
    outc, _ = hc.traj_eval_se3(fx.SE3_KNOTS, fx.SE3_DT, fx.SE3_T0, t, compat=True)
    assert np.abs(outc[:, 6:9] - oc["acceleration"]).max() < 1e-12 * np.abs(oc["acceleration"]).max()
    straj = kto.Traj(kto.SPLIT, fx.R3_DT, fx.R3_T0, fx.R3_KNOTS, fx.SO3_DT, fx.SO3_T0, fx.SO3_KNOTS)
    t = np.linspace(straj.min_time, straj.max_time - 1e-6, 40)
    o = kto.traj_evaluate(straj, t, Pf | V | A | Q | W)
    out, st = hc.traj_eval_split(fx.R3_KNOTS, fx.R3_DT, fx.R3_T0, fx.SO3_KNOTS, fx.SO3_DT, fx.SO3_T0, t)
    assert (st == 0).all()
    for a, b in ((out[:, 0:3], o["position"]), (out[:, 3:6], o["velocity"]), (out[:, 6:9], o["acceleration"]), (out[:, 9:13], o["orientation"]),
                 (out[:, 13:16], o["angular_velocity"])):
        assert np.abs(a - b).max() < 1e-12 * max(1.0, np.abs(b).max())


# ---- AtanCamera and NewtonRsCameraMeasurement ---------------------------------------------------------
def _camera_case_model(dt, seed, atan, method):
    knots, s, _ = _camera_case(dt, seed)
    kw = dict(wc=(0.02, -0.01), gamma=0.9) if atan else {}
    cam = kto.Camera(s["rows"], s["cols"], s["readout"], K=s["K"], method=method, q_ct=fx.so3_exp_xyzw(np.array([0.1, -0.2, 0.05])),
                     p_ct=np.array([0.05, -0.02, 0.1]), **kw)
    return knots, s, cam


@pytest.mark.parametrize("dt", [0.02, 0.1])
def test_static_rs_atan_camera_rows_match_oracle(dt):
    """AtanCamera (sensors/atan_camera.h:54-103) under the static measurement: analytic projection Jacobian vs the oracle's autodiff."""
    knots, s, cam = _camera_case_model(dt, 5, True, "static")
    o = kto.static_rs_residuals(kto.Traj(kto.SE3, dt, 0.0, knots), cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"],
                                s["weight"], jac_mode=2, cap=24)
    h = hc.static_rs(knots, dt, 0.0, cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"], s["weight"])
    assert (h["status"] == 0).all()
    assert (h["i0_ref"] == o["i0_ref_a"]).all() and (h["i0_obs"] == o["i0_obs_a"]).all()
    assert np.abs(h["r"] - o["r"]).max() < parity.CAM_R_TOL
    Js, Jrho = parity.scatter_cam(h["J"], h["i0_ref"], h["i0_obs"], o["ids_a"])
    assert parity.rel_err(Js, o["Ja"]) < parity.TOL and parity.rel_err(Jrho, o["Jrho"]) < parity.TOL


@pytest.mark.parametrize("dt", [0.02, 0.1])
@pytest.mark.parametrize("atan", [False, True])
def test_newton_rs_rows_match_oracle(dt, atan):
    """NewtonRsCameraMeasurement (newton_rscamera_measurement.h:23-120): residual and the derivative THROUGH the iteration."""
    knots, s, cam = _camera_case_model(dt, 5, atan, "newton")
    o = kto.static_rs_residuals(kto.Traj(kto.SE3, dt, 0.0, knots), cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"],
                                s["weight"], jac_mode=2, cap=24)
    h = hc.newton_rs(knots, dt, 0.0, cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"], s["weight"])
    assert (h["status"] == 0).all()
    assert (h["i0_ref"] == o["i0_ref_a"]).all() and (h["i0_obs"] == o["i0_obs_a"]).all()       # bit-exact: i0_obs = first knot of the span
    assert np.abs(h["r"] - o["r"]).max() < parity.CAM_R_TOL
    Js, Jrho = parity.scatter_cam(h["J"], h["i0_ref"], h["i0_obs"], o["ids_a"], h["W"])
    assert parity.rel_err(Js, o["Ja"]) < parity.TOL and parity.rel_err(Jrho, o["Jrho"]) < parity.TOL
    it = np.bincount(h["iterations"], minlength=6)
    assert it[1] > 0 and it[2:].sum() > 0          # rows that stop after one step AND rows that iterate
    # ... and the rows as the two kernels produce them: closed form where the iteration stops after one or two evaluations (static row at
    # t_1 + pi'(t_1) d t_1 / d theta, one dual evaluation per direction instead of two)
    hf = hc.newton_rs(knots, dt, 0.0, cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"], s["weight"], fast=True)
    assert (hf["status"] == 0).all() and (hf["i0_ref"] == o["i0_ref_a"]).all()
    assert np.abs(hf["r"] - o["r"]).max() < parity.CAM_R_TOL
    Jf, Jrf = parity.scatter_cam(hf["J"], hf["i0_ref"], hf["i0_obs"], o["ids_a"], hf["W"])
    assert parity.rel_err(Jf, o["Ja"]) < parity.TOL and parity.rel_err(Jrf, o["Jrho"]) < parity.TOL
    assert (h["iterations"] == 2).sum() > 5


@pytest.mark.parametrize("seed", [21, 24, 48, 22, 23])
def test_newton_rs_rows_closed_form_for_any_number_of_evaluations(seed):
    """Reverse mode (newton_math.cuh "NewtonRs rows in CLOSED FORM for any number of evaluations"): static row at t_last + jfin (x) d t_last / d theta, with
    d t_last / d theta carried through the iteration by one reverse sweep (pose + body twist) per evaluation.  Observed rows anywhere in the image, so that
    rows with two, three and more evaluations occur; against forward mode through the iteration (same harness) and against the oracle's autodiff."""
    dt, atan = (0.02, 0.05, 0.1)[seed % 3], bool(seed & 1)
    knots, s, cam = _camera_case_model(dt, seed, atan, "newton")
    rng = np.random.default_rng(seed)
    uv = s["obs_uv"].copy()
    uv[:, 1] = rng.uniform(1.0, cam.rows - 2.0, len(uv))
    args = (knots, dt, 0.0, cam, uv, s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"], s["weight"])
    h0, h2 = hc.newton_rs(*args), hc.newton_rs(*args, fast=2)
    assert (h0["status"] == h2["status"]).all()
    ok = h0["status"] == 0
    assert ok.sum() > 0.8 * len(ok) and (h0["iterations"][ok] == h2["iterations"][ok]).all()
    assert dt != 0.02 or (h0["iterations"][ok] >= 3).sum() >= 10
    sc = np.abs(h0["J"][ok]).max(axis=1, keepdims=True)
    assert (np.abs(h2["J"][ok] - h0["J"][ok]) / sc).max() < 1e-11
    assert np.abs(h2["r"][ok] - h0["r"][ok]).max() < parity.CAM_R_TOL
    o = kto.static_rs_residuals(kto.Traj(kto.SE3, dt, 0.0, knots), cam, uv[ok], s["obs_t0"][ok], s["ref_uv"][ok], s["ref_t0"][ok], s["lm_idx"][ok], s["rho"],
                                s["weight"][ok], jac_mode=2, cap=24)
    assert (h2["i0_ref"][ok] == o["i0_ref_a"]).all() and (h2["i0_obs"][ok] == o["i0_obs_a"]).all()
    Js, Jrho = parity.scatter_cam(h2["J"][ok], h2["i0_ref"][ok], h2["i0_obs"][ok], o["ids_a"], h2["W"])
    assert parity.rel_err(Js, o["Ja"]) < parity.TOL and parity.rel_err(Jrho, o["Jrho"]) < parity.TOL


def test_newton_rs_one_step_rows_equal_static_rows():
    """A row whose first Newton step is below half a row time returns the projection at the observed row: the static-RS row."""
    dt = 0.05
    knots, s, cam = _camera_case_model(dt, 11, False, "newton")
    args = (knots, dt, 0.0, cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"], s["weight"])
    hn, hs = hc.newton_rs(*args), hc.static_rs(*args)
    one = hn["iterations"] == 1
    assert one.sum() > 5
    assert np.abs(hn["r"][one] - hs["r"][one]).max() < 1e-9
    Jn = hn["J"][one]
    W = hn["W"]
    off = (hs["i0_obs"][one] - hn["i0_obs"][one])          # position of the static row's 4-knot window inside the span
    for a, (jn, js, k) in enumerate(zip(Jn, hs["J"][one], off)):
        assert np.abs(jn[:56] - js[:56]).max() <= 1e-9 * np.abs(js[:56]).max()
        assert np.abs(jn[56 + 14 * k:56 + 14 * (k + 4)] - js[56:112]).max() <= 1e-9 * np.abs(js[56:112]).max()
        assert np.abs(jn[-2:] - js[112:114]).max() <= 1e-9 * np.abs(js[112:114]).max()


def test_newton_rs_huber_corrector_matches_oracle():
    dt = 0.05
    knots, s, cam = _camera_case_model(dt, 9, False, "newton")
    n = len(s["lm_idx"])
    o = kto.static_rs_residuals(kto.Traj(kto.SE3, dt, 0.0, knots), cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"],
                                s["weight"], jac_mode=2, cap=24)
    for fast in (0, 1, 2):
        h = hc.newton_rs(knots, dt, 0.0, cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"], s["weight"], huber_c=np.full(n, 5.0), fast=fast)
        for i in range(n):
            m = int((o["ids_a"][i] >= 0).sum())
            Jfull = np.concatenate([o["Ja"][i, k] for k in range(m)] + [o["Jrho"][i].reshape(2, 1)], axis=1)
            _, r2, J2 = kto.huber_correct(5.0, o["r"][i], Jfull)
            Js, Jr = parity.scatter_cam(h["J"][i:i + 1], h["i0_ref"][i:i + 1], h["i0_obs"][i:i + 1], o["ids_a"][i:i + 1], h["W"])
            Jmine = np.concatenate([Js[0, k] for k in range(m)] + [Jr[0].reshape(2, 1)], axis=1)
            assert np.abs(Jmine - J2).max() <= parity.TOL * np.abs(J2).max()
            assert np.abs(h["r"][i] - r2).max() <= parity.CAM_R_TOL


def test_position_rows_match_oracle():
    """PositionMeasurement (position_measurement.h:24-31) on SE3: r = p - position(t), Jacobian through the cumulative spline."""
    knots = fx.smooth_se3_knots(60, 0.1)
    rng = np.random.default_rng(8)
    t, y, w = rng.uniform(0.0, 5.69, 200), rng.uniform(-5, 5, (200, 3)), rng.uniform(0.5, 2, 200)
    o = parity.oracle_imu(kto.Traj(kto.SE3, 0.1, 0.0, knots), 2, t, y, w)
    h = hc.imu(2, knots, 0.1, 0.0, t, y, w)
    assert (h["status"] == 0).all() and (h["i0"] == o["i0"]).all()
    assert parity.rel_err(h["r"], o["r"]) < parity.TOL
    assert parity.rel_err(h["J"], o["J"]) < parity.TOL
    # ... and on the reference's own SE3 fixture (large relative rotations between knots)
    t = np.linspace(fx.SE3_T0, fx.SE3_T0 + 3 * fx.SE3_DT - 1e-9, 40)
    o = parity.oracle_imu(kto.Traj(kto.SE3, fx.SE3_DT, fx.SE3_T0, fx.SE3_KNOTS), 2, t, np.zeros((40, 3)))
    h = hc.imu(2, fx.SE3_KNOTS, fx.SE3_DT, fx.SE3_T0, t, np.zeros((40, 3)))
    assert parity.rel_err(h["r"], o["r"]) < parity.TOL and parity.rel_err(h["J"], o["J"]) < parity.TOL


def _orientation_case(knots, dt, t0, n, seed):
    """Measured orientations = the trajectory's own, rotated by a random 0.05 .. 2.5 rad (both hemispheres of the quaternion)."""
    rng = np.random.default_rng(seed)
    t = rng.uniform(t0 + 1e-3, t0 + (len(knots) - 3) * dt - 1e-3, n)
    q = kto.traj_evaluate(kto.Traj(kto.SE3, dt, t0, knots), t, 0xff)["orientation"]
    ang = rng.uniform(0.05, 2.5, n)
    ax = rng.normal(size=(n, 3)); ax /= np.linalg.norm(ax, axis=1)[:, None]
    dq = np.concatenate([ax * np.sin(ang / 2)[:, None], np.cos(ang / 2)[:, None]], axis=1)
    x1, y1, z1, w1 = q.T; x2, y2, z2, w2 = dq.T
    qm = np.stack([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
                   w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2], axis=1)
    qm[::2] *= -1.0                     # q and -q are the same rotation: angularDistance uses |d.w|
    qm[::3] *= 1.7                      # ... and is scale invariant
    return t, qm, ang


def test_orientation_rows_match_oracle():
    """OrientationMeasurement (orientation_measurement.h:27-31): one residual = angular distance, analytic Jacobian vs the oracle's autodiff."""
    knots = fx.smooth_se3_knots(60, 0.1)
    t, qm, ang = _orientation_case(knots, 0.1, 0.0, 200, 3)
    o = kto.imu_residuals(kto.Traj(kto.SE3, 0.1, 0.0, knots), kto.Sensor(), 3, t, qm, jac_mode=2)
    assert np.abs(o["r"][:, 0] - ang).max() < 1e-9          # the oracle returns the angle the case was built with
    h = hc.imu(3, knots, 0.1, 0.0, t, qm)
    assert (h["status"] == 0).all() and (h["i0"] == o["i0_a"]).all()
    assert np.abs(h["r"] - o["r"]).max() < parity.TOL
    assert parity.rel_err(h["J"], o["Ja"][:, :4]) < parity.TOL


@pytest.mark.parametrize("atan,robust", [(False, False), (True, True)])
def test_lifting_rs_rows_match_oracle(atan, robust):
    """LiftingRsCameraMeasurement (lifting_rscamera_measurement.h:21-56, :98-149): 3 residuals, the observation evaluated at the lifted time
    t0 + vt readout with vt its own parameter block; forward-mode columns on the hoisted structure vs the oracle's autodiff, at the initial
    vt (= v / rows, where rows 0..1 are the static rows and row 2 is zero) and at displaced vt (all three rows alive)."""
    dt = 0.05
    knots, s, cam = _camera_case_model(dt, 11, atan, "static")
    n = len(s["lm_idx"])
    rng = np.random.default_rng(4)
    args = (s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"])
    for vt in (None, np.clip(s["obs_uv"][:, 1] / s["rows"] + rng.uniform(-0.2, 0.2, n), 0.0, 1.0)):
        o = kto.lifting_rs_residuals(kto.Traj(kto.SE3, dt, 0.0, knots), cam, *args, vt=vt, weight=s["weight"], jac_mode=2, cap=24)
        c = np.full(n, 2.0) if robust else None
        h = hc.lifting_rs(knots, dt, 0.0, cam, *args, vt=vt, w=s["weight"], huber_c=c)
        hf = hc.lifting_rs(knots, dt, 0.0, cam, *args, vt=vt, w=s["weight"], huber_c=c, analytic=False)
        # the closed-form rows and the forward-mode rows are the same rows
        assert (hf["status"] == 0).all() and np.abs(hf["r"] - h["r"]).max() < parity.CAM_R_TOL and parity.rel_err(hf["J"][:, None], h["J"][:, None]) < parity.TOL
        assert (h["status"] == 0).all() and (h["i0_ref"] == o["i0_ref_a"]).all()
        Js, Jvt, Jrho = parity.scatter_lifting(h["J"], h["i0_ref"], h["i0_obs"], o["ids_a"], h["W"])
        if not robust:
            assert np.abs(h["r"] - o["r"]).max() < parity.CAM_R_TOL
            assert parity.rel_err(Js, o["Ja"]) < parity.TOL and parity.rel_err(Jvt, o["Jvt"]) < parity.TOL and parity.rel_err(Jrho, o["Jrho"]) < parity.TOL
            if vt is None:
                assert not h["r"][:, 2].any() and np.allclose(h["J"][:, -6:-3], o["Jvt"])
        else:
            n_out = 0
            for i in range(n):
                m = int((o["ids_a"][i] >= 0).sum())
                Jfull = np.concatenate([o["Ja"][i, k] for k in range(m)] + [o["Jvt"][i].reshape(3, 1), o["Jrho"][i].reshape(3, 1)], axis=1)
                _, r2, J2 = kto.huber_correct(2.0, o["r"][i], Jfull)
                Jmine = np.concatenate([Js[i, k] for k in range(m)] + [Jvt[i].reshape(3, 1), Jrho[i].reshape(3, 1)], axis=1)
                assert np.abs(h["r"][i] - r2).max() < parity.CAM_R_TOL and parity.rel_err(Jmine[None], J2[None]) < parity.TOL
                n_out += float(np.dot(o["r"][i], o["r"][i])) > 4.0
            assert n_out > 0


def test_optimised_cpu_variant_matches_oracle():
    """oracle/analytic_cpu.cpp -- the OPTIMISED CPU baseline bench.py reports next to the restated reference: the product's
    closed-form math on the host with OpenMP.  It is only a baseline if it computes the same rows: gyro / accel / static-RS vs the oracle."""
    from kontiki_b200 import synthetic as syn
    cfg = syn.make_config("H1", scale=0.002)
    c = cfg["cam"]
    camd = {k: c[k] for k in ("K", "readout", "rows", "obs_uv", "obs_t0", "ref_uv", "ref_t0", "lm_idx", "rho", "weight")}
    res = kto.analytic_se3_evaluate(cfg["dt"], cfg["t0"], cfg["knots"], gyro=cfg["gyro"], accel=cfg["accel"], cam=camd, nthreads=2)
    assert res["bad"] == 0 and res["seconds"] > 0
    traj = kto.Traj(kto.SE3, cfg["dt"], cfg["t0"], cfg["knots"])
    for which, key in ((0, "gyro"), (1, "accel")):
        m = cfg[key]
        o = kto.imu_residuals(traj, kto.Sensor(), which, m["t"], m["y"], m["weight"], jac_mode=2)
        assert parity.rel_err(res[key][0], o["r"]) < parity.TOL and parity.rel_err(res[key][1].reshape(-1, 4, 3, 7), o["Ja"][:, :4]) < parity.TOL
    ocam = kto.Camera(c["rows"], c["cols"], c["readout"], K=c["K"])
    o = kto.static_rs_residuals(traj, ocam, c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"], c["rho"], c["weight"], jac_mode=2, cap=24)
    r, J = res["cam"]
    assert np.abs(r - o["r"]).max() < parity.CAM_R_TOL
    Js, Jrho = parity.scatter_cam(J, o["i0_ref_a"], o["i0_obs_a"], o["ids_a"])
    assert parity.rel_err(Js, o["Ja"]) < parity.TOL and parity.rel_err(Jrho, o["Jrho"]) < parity.TOL



@pytest.mark.parametrize("atan,robust", [(False, False), (True, True)])
def test_span_camera_sensor_columns_match_oracle(atan, robust):
    """Unlocked relative pose of the camera (sensors.h:135-165) under NewtonRs / LiftingRs rows: the q_ct (ambient) and p_ct columns, forward mode
    through BOTH sides of the row (the landmark depends on the camera pose too) vs the oracle's autodiff over the sensor blocks."""
    dt = 0.05
    n_checked = 0
    for method, lifting in (("newton", False), ("static", True)):
        knots, s, cam = _camera_case_model(dt, 11, atan, method)
        cam.q_locked = cam.p_locked = False
        n = len(s["lm_idx"])
        args = (s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"])
        traj = kto.Traj(kto.SE3, dt, 0.0, knots)
        nres = 3 if lifting else 2
        if lifting:
            vt = np.clip(s["obs_uv"][:, 1] / s["rows"] + np.random.default_rng(4).uniform(-0.2, 0.2, n), 0.0, 1.0)
            o = kto.lifting_rs_residuals(traj, cam, *args, vt=vt, weight=s["weight"], jac_mode=2, cap=24)
        else:
            vt = None
            o = kto.static_rs_residuals(traj, cam, *args, s["weight"], jac_mode=2, cap=24)
        c = np.full(n, 2.0) if robust else None
        h = hc.span_sensor(knots, dt, 0.0, cam, *args, lifting=lifting, vt=vt, w=s["weight"], huber_c=c)
        assert (h["status"] == 0).all()
        Jo = o["Js"]
        if robust:      # ceres::Corrector on the sensor columns of every row
            Jo = Jo.copy()
            for i in range(n):
                cols = np.concatenate([Jo[i, :4 * nres].reshape(nres, 4), Jo[i, 4 * nres:7 * nres].reshape(nres, 3), Jo[i, 7 * nres:].reshape(nres, 1)], axis=1)
                _, _, J2 = kto.huber_correct(2.0, o["r"][i], cols)
                Jo[i] = np.concatenate([J2[:, :4].reshape(-1), J2[:, 4:7].reshape(-1), J2[:, 7].reshape(-1)])
        assert np.abs(Jo[:, :7 * nres]).max() > 1.0
        assert parity.rel_err(h["Js"][:, None, :7 * nres], Jo[:, None, :7 * nres]) < parity.TOL
        assert not h["Js"][:, 7 * nres:].any()          # the time offset of these measurements stays locked
        n_checked += n
    assert n_checked > 50

"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs;
plus size-independent properties at BASELINE.json's full sizes."""
import numpy as np
import pytest

import fixtures_ref as fx
import parity
from kontiki_b200 import _lib, synthetic as syn
from oracle import kto

pytestmark = pytest.mark.gpu


def _problem(cfg, compat=False):
    p = _lib.Problem(0)
    p.set_se3_spline(cfg["dt"], cfg["t0"], len(cfg["knots"]), compat_zero_dB=compat)
    groups = {}
    imu = _lib.make_sensor()
    if cfg["gyro"]:
        groups["gyro"] = p.add_gyroscope(imu, cfg["gyro"]["t"], cfg["gyro"]["y"], cfg["gyro"]["weight"])
    if cfg["accel"]:
        groups["accel"] = p.add_accelerometer(imu, cfg["accel"]["t"], cfg["accel"]["y"], cfg["accel"]["weight"])
    if cfg["cam"]:
        c = cfg["cam"]
        cam = _lib.make_camera(c["rows"], c["cols"], c["readout"], c["K"], q_ct=c.get("q_ct", (0, 0, 0, 1)), p_ct=c.get("p_ct", (0, 0, 0)))
        groups["cam"] = p.add_static_rs(cam, c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"], c["weight"], c["huber_c"])
    return p, groups


def _check_imu(out, o):
    assert (out["i0"] == o["i0"]).all()                               # bit-exact indexing
    assert parity.rel_err(out["r"], o["r"]) < parity.TOL
    assert parity.rel_err(out["J"], o["J"]) < parity.TOL


def _check_cam(p, g, out, cfg, robust):
    c = cfg["cam"]
    traj = kto.Traj(kto.SE3, cfg["dt"], cfg["t0"], cfg["knots"])
    ocam = kto.Camera(c["rows"], c["cols"], c["readout"], K=c["K"], q_ct=c.get("q_ct", (0, 0, 0, 1)), p_ct=c.get("p_ct", (0, 0, 0)))
    o = kto.static_rs_residuals(traj, ocam, c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"], c["rho"], c["weight"], jac_mode=2, cap=24)
    assert (out["i0"] == o["i0_ref_a"]).all() and (out["i0_b"] == o["i0_obs_a"]).all()
    ids, nids = p.get_structure(g, cap=24)
    assert (ids == o["ids_a"]).all()
    Js = p.expand_static_rs(g, ids, out["J"], out["i0"], out["i0_b"])
    Jrho = out["J"][:, 112:114]
    if not robust:
        assert np.abs(out["r"] - o["r"]).max() < parity.CAM_R_TOL
        assert parity.rel_err(Js, o["Ja"]) < parity.TOL
        assert parity.rel_err(Jrho, o["Jrho"]) < parity.TOL
        return
    for i in range(len(c["lm_idx"])):
        m = int(nids[i])
        Jfull = np.concatenate([o["Ja"][i, k] for k in range(m)] + [o["Jrho"][i].reshape(2, 1)], axis=1)
        _, r2, J2 = kto.huber_correct(c["huber_c"][i], o["r"][i], Jfull)
        Jmine = np.concatenate([Js[i, k] for k in range(m)] + [Jrho[i].reshape(2, 1)], axis=1)
        assert np.abs(Jmine - J2).max() <= parity.TOL * np.abs(J2).max()
        assert np.abs(out["r"][i] - r2).max() <= parity.CAM_R_TOL


def test_c1_gyro_matches_oracle():
    """BASELINE.json configs[0]: SE3, 200 knots, 5k gyroscope measurements."""
    cfg = syn.make_config("C1")
    p, g = _problem(cfg)
    outs = p.evaluate(cfg["knots"])
    o = parity.oracle_imu(kto.Traj(kto.SE3, cfg["dt"], 0.0, cfg["knots"]), 0, cfg["gyro"]["t"], cfg["gyro"]["y"], cfg["gyro"]["weight"])
    _check_imu(outs[g["gyro"]], o)


@pytest.mark.parametrize("compat", [False, True])
def test_c2_scaled_gyro_accel_match_oracle(compat):
    cfg = syn.make_config("C2", scale=0.05)
    rng = np.random.default_rng(12)
    cfg["gyro"]["weight"] = rng.uniform(0.5, 2, len(cfg["gyro"]["t"]))
    cfg["accel"]["weight"] = rng.uniform(0.5, 2, len(cfg["accel"]["t"]))
    p, g = _problem(cfg, compat)
    outs = p.evaluate(cfg["knots"])
    traj = kto.Traj(kto.SE3, cfg["dt"], 0.0, cfg["knots"], compat_zero_dB=compat)
    _check_imu(outs[g["gyro"]], parity.oracle_imu(traj, 0, cfg["gyro"]["t"], cfg["gyro"]["y"], cfg["gyro"]["weight"]))
    _check_imu(outs[g["accel"]], parity.oracle_imu(traj, 1, cfg["accel"]["t"], cfg["accel"]["y"], cfg["accel"]["weight"]))


def test_reference_fixture_knots():
    """The reference's SE3 fixture (python/tests/conftest.py:83-105): large relative rotations between knots."""
    t = np.linspace(fx.SE3_T0, fx.SE3_T0 + 3 * fx.SE3_DT - 1e-9, 257)
    y = np.random.default_rng(3).uniform(-1, 1, (257, 3))
    cfg = dict(knots=fx.SE3_KNOTS, dt=fx.SE3_DT, t0=fx.SE3_T0, gyro=dict(t=t, y=y, weight=np.ones(257)), accel=dict(t=t, y=y, weight=np.ones(257)), cam=None)
    p, g = _problem(cfg)
    outs = p.evaluate(cfg["knots"])
    traj = kto.Traj(kto.SE3, fx.SE3_DT, fx.SE3_T0, fx.SE3_KNOTS)
    _check_imu(outs[g["gyro"]], parity.oracle_imu(traj, 0, t, y))
    _check_imu(outs[g["accel"]], parity.oracle_imu(traj, 1, t, y))


@pytest.mark.parametrize("robust", [False, True])
@pytest.mark.parametrize("rel_pose", [False, True])
def test_h1_scaled_matches_oracle(robust, rel_pose):
    """North-star problem (C3 + IMU) at 1/100 scale: 5k static-RS + 1k IMU rows, with gross outliers for the Huber branch."""
    cfg = syn.make_config("H1", scale=0.01)
    c = cfg["cam"]
    rng = np.random.default_rng(21)
    out = rng.random(len(c["lm_idx"])) < 0.2
    c["obs_uv"][out] += rng.normal(0, 40, (out.sum(), 2))
    c["weight"] = rng.uniform(0.5, 2, len(c["lm_idx"]))
    if rel_pose:
        c["q_ct"] = fx.so3_exp_xyzw(np.array([0.1, -0.2, 0.05]))
        c["p_ct"] = np.array([0.05, -0.02, 0.1])
    p, g = _problem(cfg)
    flags = _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS | (_lib.EVAL_ROBUST if robust else 0)
    outs = p.evaluate(cfg["knots"], c["rho"], flags)
    traj = kto.Traj(kto.SE3, cfg["dt"], 0.0, cfg["knots"])
    _check_imu(outs[g["gyro"]], parity.oracle_imu(traj, 0, cfg["gyro"]["t"], cfg["gyro"]["y"]))
    _check_imu(outs[g["accel"]], parity.oracle_imu(traj, 1, cfg["accel"]["t"], cfg["accel"]["y"]))
    _check_cam(p, g["cam"], outs[g["cam"]], cfg, robust)


def test_out_of_range_is_an_error_not_garbage():
    cfg = syn.make_config("C1", scale=0.01)
    cfg["gyro"]["t"][7] = 19.71                 # valid time is [0, 19.7)
    p, g = _problem(cfg)
    with pytest.raises(ValueError):
        p.evaluate(cfg["knots"])


def test_empty_and_ragged_groups():
    cfg = syn.make_config("C1", scale=0.0134)   # 67 rows: one full CTA + a ragged tail
    p = _lib.Problem(0)
    p.set_se3_spline(cfg["dt"], 0.0, len(cfg["knots"]))
    imu = _lib.make_sensor()
    g0 = p.add_gyroscope(imu, np.zeros(0), np.zeros((0, 3)))
    g1 = p.add_gyroscope(imu, cfg["gyro"]["t"], cfg["gyro"]["y"])
    outs = p.evaluate(cfg["knots"])
    assert outs[g0]["r"].shape == (0, 3)
    o = parity.oracle_imu(kto.Traj(kto.SE3, cfg["dt"], 0.0, cfg["knots"]), 0, cfg["gyro"]["t"], cfg["gyro"]["y"])
    _check_imu(outs[g1], o)
    # empty groups of every other kind sit in the same problem without disturbing it
    cam = _lib.make_camera(1080, 1920, 0.026, np.array([[900., 0, 960], [0, 900, 540], [0, 0, 1]]))
    e = np.zeros(0)
    g2 = p.add_orientation(e, np.zeros((0, 4)))
    g3 = p.add_position(e, np.zeros((0, 3)))
    g4 = p.add_lifting_rs(cam, np.zeros((0, 2)), e, np.zeros((0, 2)), e, np.zeros(0, np.int32))
    g5 = p.add_newton_rs(cam, np.zeros((0, 2)), e, np.zeros((0, 2)), e, np.zeros(0, np.int32))
    outs = p.evaluate(cfg["knots"], np.zeros(1))
    assert outs[g2]["r"].shape == (0, 1) and outs[g3]["r"].shape == (0, 3) and outs[g4]["r"].shape == (0, 3) and outs[g5]["r"].shape == (0, 2)
    _check_imu(outs[g1], o)


def test_full_size_properties_h1():
    """At full H1 size (600k rows) the oracle is too slow; use size-independent properties:
    (1) rows are independent of the batch they are evaluated in and of its order (a random 2k subset evaluated alone, in
        shuffled order, gives bit-identical rows), (2) that subset matches the oracle, (3) residuals scale linearly with
        the weight (reference test_measurements.py:73-89), (4) J is the derivative of r: r(x + h d) - r(x - h d) ~ 2h J d."""
    cfg = syn.make_config("H1")
    c = cfg["cam"]
    p, g = _problem(cfg)
    outs = p.evaluate(cfg["knots"], c["rho"])
    assert all(np.isfinite(o["r"]).all() and np.isfinite(o["J"]).all() for o in outs)
    rng = np.random.default_rng(5)
    # (1) + (2): subsets
    sel = rng.permutation(len(cfg["gyro"]["t"]))[:2000]
    sub = dict(cfg, gyro={k: v[sel] for k, v in cfg["gyro"].items()}, accel={k: v[sel] for k, v in cfg["accel"].items()}, cam=None)
    p2, g2 = _problem(sub)
    outs2 = p2.evaluate(cfg["knots"])
    for name in ("gyro", "accel"):
        assert np.array_equal(outs2[g2[name]]["r"], outs[g[name]]["r"][sel])
        assert np.array_equal(outs2[g2[name]]["J"], outs[g[name]]["J"][sel])
        assert np.array_equal(outs2[g2[name]]["i0"], outs[g[name]]["i0"][sel])
    traj = kto.Traj(kto.SE3, cfg["dt"], 0.0, cfg["knots"])
    _check_imu(outs2[g2["gyro"]], parity.oracle_imu(traj, 0, sub["gyro"]["t"], sub["gyro"]["y"]))
    _check_imu(outs2[g2["accel"]], parity.oracle_imu(traj, 1, sub["accel"]["t"], sub["accel"]["y"]))
    csel = rng.permutation(len(c["lm_idx"]))[:3000]
    csub = dict(c, **{k: c[k][csel] for k in ("obs_uv", "obs_t0", "ref_uv", "ref_t0", "lm_idx", "weight", "huber_c")})
    sub = dict(cfg, gyro=None, accel=None, cam=csub)
    p3, g3 = _problem(sub)
    outs3 = p3.evaluate(cfg["knots"], c["rho"])
    for k in ("r", "J", "i0", "i0_b"):
        assert np.array_equal(outs3[g3["cam"]][k], outs[g["cam"]][k][csel])
    _check_cam(p3, g3["cam"], outs3[g3["cam"]], sub, robust=False)
    # (3) weight linearity, exact in fp64 for a power-of-two factor
    cfg4 = dict(cfg, gyro=dict(cfg["gyro"], weight=cfg["gyro"]["weight"] * 4.0), accel=None, cam=dict(c, weight=c["weight"] * 0.5))
    p4, g4 = _problem(cfg4)
    outs4 = p4.evaluate(cfg["knots"], c["rho"])
    assert np.array_equal(outs4[g4["gyro"]]["r"], 4.0 * outs[g["gyro"]]["r"])
    assert np.array_equal(outs4[g4["cam"]]["r"], 0.5 * outs[g["cam"]]["r"])
    # (4) directional derivative along a random tangent-space direction of all knots and all rho
    h = 1e-6
    d = rng.normal(0, 1, cfg["knots"].shape)
    # quaternion part: tangent directions only (d_q orthogonal to q).  The residual is defined on unit quaternions; along
    # q itself the reference's polynomial R(q) / q*v arithmetic gives the Jacobian a component (reproduced, see
    # spline_math.cuh "radial") that a finite difference of the closed-form residual does not see.
    q = cfg["knots"][:, :4]
    d[:, :4] -= (d[:, :4] * q).sum(1, keepdims=True) * q
    drho = rng.normal(0, 1, c["rho"].shape) * c["rho"]
    op = p.evaluate(cfg["knots"] + h * d, c["rho"] + h * drho, _lib.EVAL_RESIDUALS)
    om = p.evaluate(cfg["knots"] - h * d, c["rho"] - h * drho, _lib.EVAL_RESIDUALS)
    for name in ("gyro", "accel"):
        o = outs[g[name]]
        idx = o["i0"][:, None] + np.arange(4)[None, :]
        Jd = np.einsum("nkrc,nkc->nr", o["J"], d[idx])
        num = (op[g[name]]["r"] - om[g[name]]["r"]) / (2 * h)
        assert np.abs(Jd - num).max() < 1e-5 * max(1.0, np.abs(num).max())
    o = outs[g["cam"]]
    Jr = o["J"][:, :56].reshape(-1, 4, 2, 7)
    Jo = o["J"][:, 56:112].reshape(-1, 4, 2, 7)
    Jd = np.einsum("nkrc,nkc->nr", Jr, d[o["i0"][:, None] + np.arange(4)]) + np.einsum("nkrc,nkc->nr", Jo, d[o["i0_b"][:, None] + np.arange(4)])
    Jd += o["J"][:, 112:114] * drho[c["lm_idx"]][:, None]
    num = (op[g["cam"]]["r"] - om[g["cam"]]["r"]) / (2 * h)
    assert np.abs(Jd - num).max() < 1e-5 * max(1.0, np.abs(num).max())


# ---- split (R3 + SO3) trajectory: BASELINE.json configs[4] ---------------------------------------------------------------
def _split_case(n_knots=400, dt=0.02, scale_imu=3000, n_lm=300, seed=31, same_grid=True):
    k = syn.smooth_se3_knots(n_knots, dt)
    vecs = k[:, 4:7].copy()
    if same_grid:
        quats, dt_b, t0_b = k[:, :4].copy(), dt, 0.0
    else:      # different knot spacing / origin for the orientation spline (split_trajectory.h allows it)
        dt_b, t0_b = dt * 0.8, 0.003
        quats = syn.smooth_se3_knots(int(n_knots / 0.8) + 2, dt_b)[:, :4].copy()
    cam = syn.make_static_rs(k, dt, n_lm, obs_per_landmark=8, seed=seed, noise_px=1.0)
    rng = np.random.default_rng(seed)
    lo = max(0.0, t0_b) + 0.01
    hi = min((n_knots - 3) * dt, t0_b + (len(quats) - 3) * dt_b) - 0.05
    keep = (np.minimum(cam["ref_t0"], cam["obs_t0"]) > lo) & (np.maximum(cam["ref_t0"], cam["obs_t0"]) + cam["readout"] < hi)
    for key in ("obs_uv", "obs_t0", "ref_uv", "ref_t0", "lm_idx", "weight", "huber_c"):
        cam[key] = cam[key][keep]
    out = rng.random(len(cam["lm_idx"])) < 0.2
    cam["obs_uv"][out] += rng.normal(0, 40, (out.sum(), 2))
    cam["weight"] = rng.uniform(0.5, 2, len(cam["lm_idx"]))
    t = rng.uniform(lo, hi, scale_imu)
    return dict(vecs=vecs, dt_a=dt, t0_a=0.0, quats=quats, dt_b=dt_b, t0_b=t0_b, cam=cam, t=t, y=rng.uniform(-1, 1, (scale_imu, 3)),
                w=rng.uniform(0.5, 2, scale_imu))


@pytest.mark.parametrize("same_grid", [True, False])
@pytest.mark.parametrize("robust", [False, True])
def test_split_trajectory_matches_oracle(same_grid, robust):
    c = _split_case(same_grid=same_grid)
    cam = c["cam"]
    p = _lib.Problem(0)
    p.set_split_spline(c["dt_a"], c["t0_a"], len(c["vecs"]), c["dt_b"], c["t0_b"], len(c["quats"]))
    imu = _lib.make_sensor()
    gg = p.add_gyroscope(imu, c["t"], c["y"], c["w"])
    ga = p.add_accelerometer(imu, c["t"], c["y"], c["w"])
    q_ct, p_ct = fx.so3_exp_xyzw(np.array([0.1, -0.2, 0.05])), np.array([0.05, -0.02, 0.1])
    gc = p.add_static_rs(_lib.make_camera(cam["rows"], cam["cols"], cam["readout"], cam["K"], q_ct=q_ct, p_ct=p_ct), cam["obs_uv"], cam["obs_t0"],
                         cam["ref_uv"], cam["ref_t0"], cam["lm_idx"], cam["weight"], cam["huber_c"])
    flags = _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS | (_lib.EVAL_ROBUST if robust else 0)
    outs = p.evaluate((c["vecs"], c["quats"]), cam["rho"], flags)
    traj = kto.Traj(kto.SPLIT, c["dt_a"], c["t0_a"], c["vecs"], c["dt_b"], c["t0_b"], c["quats"])
    # IMU
    for g, which in ((gg, 0), (ga, 1)):
        o = kto.imu_residuals(traj, kto.Sensor(), which, c["t"], c["y"], c["w"], jac_mode=2)
        out = outs[g]
        assert (out["i0_c"] == o["i0_b"]).all()
        assert parity.rel_err(out["r"], o["r"]) < parity.TOL
        assert parity.rel_err(out["J"][:, -48:].reshape(-1, 4, 3, 4), o["Jb"][:, :4]) < parity.TOL
        if which == 1:
            assert (out["i0"] == o["i0_a"]).all()
            assert parity.rel_err(out["J"][:, :36].reshape(-1, 4, 3, 3), o["Ja"][:, :4]) < parity.TOL
        else:
            assert np.abs(o["Ja"]).max() == 0.0        # gyro: R3 blocks structurally present, identically zero
        ids_a, _ = p.get_structure(g, cap=4)
        ids_b, _ = p.get_structure_so3(g, cap=4)
        assert (ids_a == o["ids_a"]).all() and (ids_b == o["ids_b"]).all()
    # camera
    ocam = kto.Camera(cam["rows"], cam["cols"], cam["readout"], K=cam["K"], q_ct=q_ct, p_ct=p_ct)
    o = kto.static_rs_residuals(traj, ocam, cam["obs_uv"], cam["obs_t0"], cam["ref_uv"], cam["ref_t0"], cam["lm_idx"], cam["rho"], cam["weight"],
                                jac_mode=2, cap=24)
    out = outs[gc]
    assert (out["i0"] == o["i0_ref_a"]).all() and (out["i0_b"] == o["i0_obs_a"]).all()
    assert (out["i0_c"] == o["i0_ref_b"]).all() and (out["i0_d"] == o["i0_obs_b"]).all()
    ids_a, _ = p.get_structure(gc, cap=24)
    ids_b, _ = p.get_structure_so3(gc, cap=24)
    assert (ids_a == o["ids_a"]).all() and (ids_b == o["ids_b"]).all()
    n = len(cam["lm_idx"])
    Ja, Jb = np.zeros_like(o["Ja"]), np.zeros_like(o["Jb"])
    for i in range(n):
        pa = {int(k): j for j, k in enumerate(o["ids_a"][i]) if k >= 0}
        pb = {int(k): j for j, k in enumerate(o["ids_b"][i]) if k >= 0}
        J = out["J"][i]
        for k in range(4):
            Ja[i, pa[out["i0"][i] + k]] += J[6 * k:6 * k + 6].reshape(2, 3)
            Ja[i, pa[out["i0_b"][i] + k]] += J[56 + 6 * k:62 + 6 * k].reshape(2, 3)
            Jb[i, pb[out["i0_c"][i] + k]] += J[24 + 8 * k:32 + 8 * k].reshape(2, 4)
            Jb[i, pb[out["i0_d"][i] + k]] += J[80 + 8 * k:88 + 8 * k].reshape(2, 4)
    Jrho = out["J"][:, 112:114]
    if not robust:
        assert np.abs(out["r"] - o["r"]).max() < parity.CAM_R_TOL
        assert parity.rel_err(Ja, o["Ja"]) < parity.TOL and parity.rel_err(Jb, o["Jb"]) < parity.TOL and parity.rel_err(Jrho, o["Jrho"]) < parity.TOL
        return
    n_out = 0
    for i in range(n):
        ma, mb = int((o["ids_a"][i] >= 0).sum()), int((o["ids_b"][i] >= 0).sum())
        Jfull = np.concatenate([o["Ja"][i, k] for k in range(ma)] + [o["Jb"][i, k] for k in range(mb)] + [o["Jrho"][i].reshape(2, 1)], axis=1)
        _, r2, J2 = kto.huber_correct(cam["huber_c"][i], o["r"][i], Jfull)
        Jmine = np.concatenate([Ja[i, k] for k in range(ma)] + [Jb[i, k] for k in range(mb)] + [Jrho[i].reshape(2, 1)], axis=1)
        assert np.abs(Jmine - J2).max() <= parity.TOL * np.abs(J2).max()
        assert np.abs(out["r"][i] - r2).max() <= parity.CAM_R_TOL
        n_out += np.linalg.norm(o["r"][i]) > cam["huber_c"][i]
    assert n_out > 5


def test_split_non_unit_quaternion_is_a_runtime_error():
    """quaternion_math.h:19-23: logq throws std::runtime_error on a non-unit quaternion -> RuntimeError in Python."""
    c = _split_case(n_knots=60, scale_imu=64, n_lm=5)
    q = c["quats"].copy()
    q[20] *= 1.01
    p = _lib.Problem(0)
    p.set_split_spline(c["dt_a"], 0.0, len(c["vecs"]), c["dt_b"], c["t0_b"], len(q))
    p.add_gyroscope(_lib.make_sensor(), c["t"][:64], c["y"][:64])
    with pytest.raises(RuntimeError):
        p.evaluate((c["vecs"], q))


# ---- unlocked sensor parameters --------------------------------------------------------------------
def test_unlocked_sensors_se3_match_oracle():
    """Time offset / camera relative pose unlocked: wider spans (structure), same knot blocks, plus the sensor-block columns.

    Reference quirk reproduced (static_rscamera_measurement.h:153-157): with the camera's time offset unlocked the FIRST span is
    shifted earlier and the SECOND later by max_time_offset (instead of both being widened), so an evaluation time can fall
    outside every segment and the reference throws from inside Evaluate.  Rows for which the oracle throws are dropped here;
    test_out_of_range_is_an_error_not_garbage covers the error path."""
    dt, n_knots = 0.1, 200
    knots = syn.smooth_se3_knots(n_knots, dt)
    rng = np.random.default_rng(8)
    t = rng.uniform(0.3, (n_knots - 3) * dt - 0.4, 3000)
    y = rng.uniform(-1, 1, (3000, 3))
    c = syn.make_static_rs(knots, dt, 300, obs_per_landmark=8, seed=8, noise_px=1.0)
    keep = (np.minimum(c["ref_t0"], c["obs_t0"]) > 0.3) & (np.maximum(c["ref_t0"], c["obs_t0"]) < (n_knots - 3) * dt - 0.4)
    q_ct, p_ct = fx.so3_exp_xyzw(np.array([0.1, -0.2, 0.05])), np.array([0.05, -0.02, 0.1])
    traj = kto.Traj(kto.SE3, dt, 0.0, knots)
    ocam = kto.Camera(c["rows"], c["cols"], c["readout"], K=c["K"], q_ct=q_ct, p_ct=p_ct, time_offset=-0.002, max_time_offset=0.004, q_locked=False,
                      p_locked=False, d_locked=False)
    o_all = kto.static_rs_residuals(traj, ocam, c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"], c["rho"], c["weight"], jac_mode=0, cap=32,
                                    raise_on_error=False)
    keep &= o_all["status"] == 0
    assert 0.5 * len(keep) < keep.sum() < len(keep)          # the quirk bites some rows, not most
    for a in ("obs_uv", "obs_t0", "ref_uv", "ref_t0", "lm_idx", "weight", "huber_c"):
        c[a] = c[a][keep]
    p = _lib.Problem(0)
    p.set_se3_spline(dt, 0.0, n_knots)
    imu = _lib.make_sensor(time_offset=0.003, max_time_offset=0.05, time_offset_locked=False)
    gg = p.add_gyroscope(imu, t, y)
    ga = p.add_accelerometer(imu, t, y)
    p.set_group_bias(ga, [0.01, -0.02, 0.03])
    cam = _lib.make_camera(c["rows"], c["cols"], c["readout"], c["K"], q_ct=q_ct, p_ct=p_ct, time_offset=-0.002, max_time_offset=0.004,
                           q_locked=False, p_locked=False, time_offset_locked=False)
    gc = p.add_static_rs(cam, c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"], c["weight"], c["huber_c"])
    outs = p.evaluate(knots, c["rho"], _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS | _lib.EVAL_SENSOR_JACOBIANS)
    for g, which, bias in ((gg, 0, None), (ga, 1, [0.01, -0.02, 0.03])):
        osen = kto.Sensor(time_offset=0.003, max_time_offset=0.05, d_locked=False, abias=bias, gbias=None if bias is None else [0, 0, 0])
        o = kto.imu_residuals(traj, osen, which, t, y, jac_mode=2, cap=16)
        out = outs[g]
        assert (out["i0"] == o["i0_a"]).all()
        ids, nids = p.get_structure(g, cap=16)
        assert (ids == o["ids_a"]).all() and nids.min() > 4
        assert parity.rel_err(out["r"], o["r"]) < parity.TOL
        pos = np.array([list(ids[i]).index(out["i0"][i]) for i in range(len(ids))])
        Jo = np.stack([o["Ja"][i, pos[i]:pos[i] + 4] for i in range(len(ids))])
        assert parity.rel_err(out["J"], Jo) < parity.TOL
        assert parity.rel_err(out["Js"], o["Js"][:, 21:24]) < parity.TOL
        if bias is not None:
            assert np.allclose(o["Js"][:, 24:33].reshape(-1, 3, 3), -np.eye(3), atol=1e-15)       # d r / d abias = -weight I
    o = kto.static_rs_residuals(traj, ocam, c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"], c["rho"], c["weight"], jac_mode=2, cap=32)
    out = outs[gc]
    assert (out["i0"] == o["i0_ref_a"]).all() and (out["i0_b"] == o["i0_obs_a"]).all()
    ids, _ = p.get_structure(gc, cap=32)
    assert (ids == o["ids_a"]).all()
    Js = p.expand_static_rs(gc, ids, out["J"], out["i0"], out["i0_b"])
    assert np.abs(out["r"] - o["r"]).max() < parity.CAM_R_TOL
    assert parity.rel_err(Js, o["Ja"]) < parity.TOL
    for a, b in ((0, 8), (8, 14), (14, 16)):
        assert parity.rel_err(out["Js"][:, a:b], o["Js"][:, a:b]) < parity.TOL


def test_unlocked_sensors_split_match_oracle():
    """SURVEY.md 8f-2 on a SPLIT trajectory: unlocked IMU time offset (+ ConstantBiasImu bias) and every camera sensor block."""
    c0 = _split_case(n_knots=300, dt=0.05, scale_imu=2000, n_lm=200, seed=41)
    cam = c0["cam"]
    traj = kto.Traj(kto.SPLIT, c0["dt_a"], c0["t0_a"], c0["vecs"], c0["dt_b"], c0["t0_b"], c0["quats"])
    q_ct, p_ct = fx.so3_exp_xyzw(np.array([0.1, -0.2, 0.05])), np.array([0.05, -0.02, 0.1])
    hi = (300 - 3) * 0.05
    t = c0["t"][(c0["t"] > 0.3) & (c0["t"] < hi - 0.4)]
    y = c0["y"][:len(t)]
    keep = (np.minimum(cam["ref_t0"], cam["obs_t0"]) > 0.3) & (np.maximum(cam["ref_t0"], cam["obs_t0"]) < hi - 0.4)
    ocam = kto.Camera(cam["rows"], cam["cols"], cam["readout"], K=cam["K"], q_ct=q_ct, p_ct=p_ct, time_offset=-0.002, max_time_offset=0.004, q_locked=False,
                      p_locked=False, d_locked=False)
    o_all = kto.static_rs_residuals(traj, ocam, cam["obs_uv"], cam["obs_t0"], cam["ref_uv"], cam["ref_t0"], cam["lm_idx"], cam["rho"], cam["weight"], jac_mode=0, cap=32,
                                    raise_on_error=False)
    keep &= o_all["status"] == 0
    assert keep.sum() > 200
    for a in ("obs_uv", "obs_t0", "ref_uv", "ref_t0", "lm_idx", "weight", "huber_c"):
        cam[a] = cam[a][keep]
    p = _lib.Problem(0)
    p.set_split_spline(c0["dt_a"], c0["t0_a"], len(c0["vecs"]), c0["dt_b"], c0["t0_b"], len(c0["quats"]))
    imu = _lib.make_sensor(time_offset=0.003, max_time_offset=0.05, time_offset_locked=False)
    gg = p.add_gyroscope(imu, t, y)
    ga = p.add_accelerometer(imu, t, y)
    p.set_group_bias(ga, [0.01, -0.02, 0.03])
    gc = p.add_static_rs(_lib.make_camera(cam["rows"], cam["cols"], cam["readout"], cam["K"], q_ct=q_ct, p_ct=p_ct, time_offset=-0.002, max_time_offset=0.004,
                                          q_locked=False, p_locked=False, time_offset_locked=False),
                         cam["obs_uv"], cam["obs_t0"], cam["ref_uv"], cam["ref_t0"], cam["lm_idx"], cam["weight"], cam["huber_c"])
    outs = p.evaluate((c0["vecs"], c0["quats"]), cam["rho"], _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS | _lib.EVAL_SENSOR_JACOBIANS)
    for g, which, bias in ((gg, 0, None), (ga, 1, [0.01, -0.02, 0.03])):
        osen = kto.Sensor(time_offset=0.003, max_time_offset=0.05, d_locked=False, abias=bias, gbias=None if bias is None else [0, 0, 0])
        o = kto.imu_residuals(traj, osen, which, t, y, jac_mode=2, cap=16)
        out = outs[g]
        assert (out["i0_c"] == o["i0_b"]).all()
        assert parity.rel_err(out["r"], o["r"]) < parity.TOL
        assert parity.rel_err(out["Js"], o["Js"][:, 21:24]) < parity.TOL
        ids_b, nids = p.get_structure_so3(g, cap=16)
        assert (ids_b == o["ids_b"]).all() and nids.min() > 4
        pos = np.array([list(ids_b[i]).index(out["i0_c"][i]) for i in range(len(ids_b))])
        Jo = np.stack([o["Jb"][i, pos[i]:pos[i] + 4] for i in range(len(ids_b))])
        assert parity.rel_err(out["J"][:, -48:].reshape(-1, 4, 3, 4), Jo) < parity.TOL
    o = kto.static_rs_residuals(traj, ocam, cam["obs_uv"], cam["obs_t0"], cam["ref_uv"], cam["ref_t0"], cam["lm_idx"], cam["rho"], cam["weight"], jac_mode=2, cap=32)
    out = outs[gc]
    assert (out["i0"] == o["i0_ref_a"]).all() and (out["i0_b"] == o["i0_obs_a"]).all() and (out["i0_c"] == o["i0_ref_b"]).all() and (out["i0_d"] == o["i0_obs_b"]).all()
    assert np.abs(out["r"] - o["r"]).max() < parity.CAM_R_TOL
    for a, b in ((0, 8), (8, 14), (14, 16)):
        assert parity.rel_err(out["Js"][:, a:b], o["Js"][:, a:b]) < parity.TOL


# ---- KTK_EVAL_LOCAL: knot blocks after the knots' LocalParameterization ------------------------------------------------------
def test_local_coordinates_equal_ambient_times_plus_jacobian():
    """J_local = J_ambient * dPlus/ddelta (LocalParameterizationSE3 uniform_se3_spline_trajectory.h:25-48; EigenQuaternionParameterization
    for SO3 knots), which is what Ceres forms after Evaluate; the ambient rows are already parity-gated against the oracle."""
    from kontiki_b200.estimator import _quat_plus_jacobian, _se3_plus_jacobian
    cfg = syn.make_config("H1", scale=0.004)
    c = cfg["cam"]
    p, g = _problem(cfg)
    amb = p.evaluate(cfg["knots"], c["rho"], _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS | _lib.EVAL_ROBUST)
    loc = p.evaluate(cfg["knots"], c["rho"], _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS | _lib.EVAL_ROBUST | _lib.EVAL_LOCAL)
    P = _se3_plus_jacobian(cfg["knots"])
    for name in ("gyro", "accel"):
        a, l = amb[g[name]], loc[g[name]]
        assert np.array_equal(a["r"], l["r"])
        ref = np.einsum("nkra,nkad->nkrd", a["J"], P[a["i0"][:, None] + np.arange(4)])
        assert parity.rel_err(l["J"].reshape(-1, 4, 3, 6), ref) < 1e-12
    a, l = amb[g["cam"]], loc[g["cam"]]
    n = len(a["r"])
    assert l["J"].shape == (n, 98)
    ref_r = np.einsum("nkra,nkad->nkrd", a["J"][:, :56].reshape(n, 4, 2, 7), P[a["i0"][:, None] + np.arange(4)])
    ref_o = np.einsum("nkra,nkad->nkrd", a["J"][:, 56:112].reshape(n, 4, 2, 7), P[a["i0_b"][:, None] + np.arange(4)])
    assert parity.rel_err(l["J"][:, :48].reshape(n, 4, 2, 6), ref_r) < 1e-12
    assert parity.rel_err(l["J"][:, 48:96].reshape(n, 4, 2, 6), ref_o) < 1e-12
    assert np.array_equal(l["J"][:, 96:98], a["J"][:, 112:114])
    # split trajectory
    s = _split_case(scale_imu=500, n_lm=60)
    cam = s["cam"]
    p = _lib.Problem(0)
    p.set_split_spline(s["dt_a"], s["t0_a"], len(s["vecs"]), s["dt_b"], s["t0_b"], len(s["quats"]))
    imu = _lib.make_sensor()
    gg = p.add_gyroscope(imu, s["t"], s["y"], s["w"])
    ga = p.add_accelerometer(imu, s["t"], s["y"], s["w"])
    gc = p.add_static_rs(_lib.make_camera(cam["rows"], cam["cols"], cam["readout"], cam["K"]), cam["obs_uv"], cam["obs_t0"], cam["ref_uv"], cam["ref_t0"],
                         cam["lm_idx"], cam["weight"], cam["huber_c"])
    amb = p.evaluate((s["vecs"], s["quats"]), cam["rho"], 3)
    loc = p.evaluate((s["vecs"], s["quats"]), cam["rho"], 3 | _lib.EVAL_LOCAL)
    Pq = _quat_plus_jacobian(s["quats"])
    n = len(s["t"])
    ref = np.einsum("nkra,nkad->nkrd", amb[gg]["J"].reshape(n, 4, 3, 4), Pq[amb[gg]["i0_c"][:, None] + np.arange(4)])
    assert loc[gg]["J"].shape == (n, 36) and parity.rel_err(loc[gg]["J"].reshape(n, 4, 3, 3), ref) < 1e-12
    ref = np.einsum("nkra,nkad->nkrd", amb[ga]["J"][:, 36:].reshape(n, 4, 3, 4), Pq[amb[ga]["i0_c"][:, None] + np.arange(4)])
    assert np.array_equal(loc[ga]["J"][:, :36], amb[ga]["J"][:, :36]) and parity.rel_err(loc[ga]["J"][:, 36:].reshape(n, 4, 3, 3), ref) < 1e-12
    a, l = amb[gc], loc[gc]
    n = len(a["r"])
    assert np.array_equal(l["J"][:, 0:24], a["J"][:, 0:24]) and np.array_equal(l["J"][:, 48:72], a["J"][:, 56:80]) and np.array_equal(l["J"][:, 96:98], a["J"][:, 112:114])
    ref = np.einsum("nkra,nkad->nkrd", a["J"][:, 24:56].reshape(n, 4, 2, 4), Pq[a["i0_c"][:, None] + np.arange(4)])
    assert parity.rel_err(l["J"][:, 24:48].reshape(n, 4, 2, 3), ref) < 1e-12
    ref = np.einsum("nkra,nkad->nkrd", a["J"][:, 80:112].reshape(n, 4, 2, 4), Pq[a["i0_d"][:, None] + np.arange(4)])
    assert parity.rel_err(l["J"][:, 72:96].reshape(n, 4, 2, 3), ref) < 1e-12


def test_device_order_is_a_bit_exact_permutation():
    """KTK_EVAL_DEVICE_ORDER: same values, rows in device order; ktk_get_row_order maps device row k to its insertion index."""
    cfg = syn.make_config("H1", scale=0.01)
    c = cfg["cam"]
    p, g = _problem(cfg)
    base = _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS | _lib.EVAL_ROBUST
    a = p.evaluate(cfg["knots"], c["rho"], base)
    d = p.evaluate(cfg["knots"], c["rho"], base | _lib.EVAL_DEVICE_ORDER)
    for name in ("gyro", "accel", "cam"):
        order = p.get_row_order(g[name])
        assert sorted(order.tolist()) == list(range(len(order)))
        for key in a[g[name]]:
            assert np.array_equal(d[g[name]][key], a[g[name]][key][order]), (name, key)
        i0 = d[g[name]]["i0_b" if name == "cam" else "i0"]
        assert (np.diff(i0) >= 0).all()            # device order = sorted by first active knot (of the observation for camera rows)


# ---- AtanCamera and NewtonRsCameraMeasurement ---------------------------------------------------------
def _camera_group(cfg, method, atan, p=None):
    c = cfg["cam"]
    kw = dict(wc=(0.02, -0.01), gamma=0.9) if atan else {}
    q_ct, p_ct = fx.so3_exp_xyzw(np.array([0.1, -0.2, 0.05])), np.array([0.05, -0.02, 0.1])
    if p is None:
        p = _lib.Problem(0)
        p.set_se3_spline(cfg["dt"], cfg["t0"], len(cfg["knots"]))
    cam = _lib.make_camera(c["rows"], c["cols"], c["readout"], c["K"], q_ct=q_ct, p_ct=p_ct, **kw)
    add = {"newton": p.add_newton_rs, "lifting": p.add_lifting_rs}.get(method, p.add_static_rs)
    g = add(cam, c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"], c["weight"], c["huber_c"])
    ocam = kto.Camera(c["rows"], c["cols"], c["readout"], K=c["K"], method="static" if method == "lifting" else method, q_ct=q_ct, p_ct=p_ct, **kw)
    return p, g, ocam


def _check_span_rows_local(p, g, knots, rho, flags, nres):
    """KTK_EVAL_LOCAL rows of a NewtonRs / LiftingRs group: [ref 4 x (nres x 6) | obs W x (nres x 6) | tail] = the ambient row's knot blocks times
    dPlus/ddelta of their knots (k_span_localize), in caller and in device order."""
    from kontiki_b200.estimator import _se3_plus_jacobian
    amb = p.evaluate(knots, rho, flags)[g]
    loc = p.evaluate(knots, rho, flags | _lib.EVAL_LOCAL)[g]
    n, la = amb["J"].shape
    tail = 2 if nres == 2 else 6
    W = (la - tail) // (7 * nres) - 4
    assert loc["J"].shape == (n, 6 * nres * (4 + W) + tail) and np.array_equal(loc["r"], amb["r"]) and np.array_equal(loc["i0_b"], amb["i0_b"])
    P = _se3_plus_jacobian(knots)
    ok = amb["i0"] >= 0
    assert ok.sum() > 0.9 * n
    kr = amb["i0"][ok][:, None] + np.arange(4)
    ko = np.minimum(amb["i0_b"][ok][:, None] + np.arange(W), len(knots) - 1)
    ref_r = np.einsum("nkra,nkad->nkrd", amb["J"][ok, :28 * nres].reshape(-1, 4, nres, 7), P[kr])
    ref_o = np.einsum("nkra,nkad->nkrd", amb["J"][ok, 28 * nres:la - tail].reshape(-1, W, nres, 7), P[ko])
    assert parity.rel_err(loc["J"][ok, :24 * nres].reshape(-1, 4, nres, 6), ref_r) < 1e-12
    assert parity.rel_err(loc["J"][ok, 24 * nres:-tail].reshape(-1, W, nres, 6), ref_o) < 1e-12
    assert np.array_equal(loc["J"][:, -tail:], amb["J"][:, -tail:])
    dev = p.evaluate(knots, rho, flags | _lib.EVAL_LOCAL | _lib.EVAL_DEVICE_ORDER)[g]
    assert np.array_equal(dev["J"], loc["J"][p.get_row_order(g)], equal_nan=True)


@pytest.mark.parametrize("method,atan,robust", [("static", True, False), ("static", True, True), ("newton", False, False), ("newton", True, True)])
def test_newton_and_atan_rows_match_oracle(method, atan, robust):
    cfg = syn.make_config("C3", scale=0.004)                      # 2000 rows on 5k knots (dt 0.02: Newton spans cover 5-6 knots)
    c = cfg["cam"]
    rng = np.random.default_rng(33)
    out = rng.random(len(c["lm_idx"])) < 0.2
    c["obs_uv"][out] += rng.normal(0, 40, (out.sum(), 2))
    c["obs_uv"] += rng.normal(0, 1.0, c["obs_uv"].shape)          # > half a row: the Newton iteration takes more than one step
    c["weight"] = rng.uniform(0.5, 2, len(c["lm_idx"]))
    p, g, ocam = _camera_group(cfg, method, atan)
    flags = _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS | (_lib.EVAL_ROBUST if robust else 0)
    out = p.evaluate(cfg["knots"], c["rho"], flags)[g]
    traj = kto.Traj(kto.SE3, cfg["dt"], cfg["t0"], cfg["knots"])
    o = kto.static_rs_residuals(traj, ocam, c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"], c["rho"], c["weight"], jac_mode=2, cap=24)
    assert (out["i0"] == o["i0_ref_a"]).all() and (out["i0_b"] == o["i0_obs_a"]).all()            # bit-exact indexing
    row = p.group_row_size(g)
    assert out["J"].shape[1] == row and (row == 114 if method == "static" else (row - 58) % 14 == 0 and row >= 114)
    ids, nids = p.get_structure(g, cap=24)
    assert (ids == o["ids_a"]).all()
    Js = p.expand_static_rs(g, ids, out["J"], out["i0"], out["i0_b"])
    Jrho = out["J"][:, row - 2:row]
    if not robust:
        assert np.abs(out["r"] - o["r"]).max() < parity.CAM_R_TOL
        assert parity.rel_err(Js, o["Ja"]) < parity.TOL and parity.rel_err(Jrho, o["Jrho"]) < parity.TOL
    else:
        for i in range(0, len(c["lm_idx"]), 7):
            m = int(nids[i])
            Jfull = np.concatenate([o["Ja"][i, k] for k in range(m)] + [o["Jrho"][i].reshape(2, 1)], axis=1)
            _, r2, J2 = kto.huber_correct(c["huber_c"][i], o["r"][i], Jfull)
            Jmine = np.concatenate([Js[i, k] for k in range(m)] + [Jrho[i].reshape(2, 1)], axis=1)
            assert np.abs(Jmine - J2).max() <= parity.TOL * np.abs(J2).max()
            assert np.abs(out["r"][i] - r2).max() <= parity.CAM_R_TOL
    # device order is the same rows, permuted, bit for bit
    dev = p.evaluate(cfg["knots"], c["rho"], flags | _lib.EVAL_DEVICE_ORDER)[g]
    order = p.get_row_order(g)
    assert np.array_equal(dev["J"], out["J"][order]) and np.array_equal(dev["r"], out["r"][order]) and np.array_equal(dev["i0_b"], out["i0_b"][order])


@pytest.mark.parametrize("atan,robust", [(False, True), (True, False)])
def test_newton_rows_closed_form_equals_forward_mode_for_any_number_of_evaluations(atan, robust, monkeypatch):
    """Observed rows anywhere in the image: the iteration takes two, three and more evaluations and is clamped to the readout interval for some rows.
    The closed-form rows (k_newton_rs_fast + k_newton_rs_rev, reverse mode: D = d t_last / d theta carried through the iteration) against the oracle's
    autodiff through the iteration, and against the library's own forward-mode kernel (KTK_NEWTON_FAST=0) and the 32-dual-evaluation variant (3)."""
    cfg = syn.make_config("C3", scale=0.004)
    c = cfg["cam"]
    rng = np.random.default_rng(71)
    c["obs_uv"][:, 1] = rng.uniform(1.0, c["rows"] - 2.0, len(c["lm_idx"]))
    c["weight"] = rng.uniform(0.5, 2, len(c["lm_idx"]))
    flags = _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS | (_lib.EVAL_ROBUST if robust else 0)
    outs = {}
    for mode in ("4", "0", "3"):
        monkeypatch.setenv("KTK_NEWTON_FAST", mode)                # read when the problem is created
        p, g, ocam = _camera_group(cfg, "newton", atan)
        outs[mode] = p.evaluate(cfg["knots"], c["rho"], flags)[g]
    monkeypatch.delenv("KTK_NEWTON_FAST")
    out = outs["4"]
    ok = out["i0"] >= 0
    assert ok.sum() > 0.9 * len(ok)
    for mode in ("0", "3"):
        o2 = outs[mode]
        assert np.array_equal(o2["i0"], out["i0"]) and np.array_equal(o2["i0_b"], out["i0_b"])
        assert np.abs(o2["r"][ok] - out["r"][ok]).max() < parity.CAM_R_TOL
        sc = np.abs(o2["J"][ok]).max(axis=1, keepdims=True)
        assert (np.abs(o2["J"][ok] - out["J"][ok]) / sc).max() < parity.TOL
        assert np.array_equal(np.isnan(o2["J"]), np.isnan(out["J"]))
    traj = kto.Traj(kto.SE3, cfg["dt"], cfg["t0"], cfg["knots"])
    sub = np.flatnonzero(ok)[::3]
    o = kto.static_rs_residuals(traj, ocam, c["obs_uv"][sub], c["obs_t0"][sub], c["ref_uv"][sub], c["ref_t0"][sub], c["lm_idx"][sub], c["rho"], c["weight"][sub],
                                jac_mode=2, cap=24)
    assert (out["i0"][sub] == o["i0_ref_a"]).all() and (out["i0_b"][sub] == o["i0_obs_a"]).all()
    row = p.group_row_size(g)
    ids, nids = p.get_structure(g, cap=24)
    Js = p.expand_static_rs(g, ids, out["J"], out["i0"], out["i0_b"])[sub]
    Jrho = out["J"][sub, row - 2:row]
    for a, i in enumerate(sub):
        m = int(nids[i])
        Jfull = np.concatenate([o["Ja"][a, k] for k in range(m)] + [o["Jrho"][a].reshape(2, 1)], axis=1)
        r2, J2 = o["r"][a], Jfull
        if robust:
            _, r2, J2 = kto.huber_correct(c["huber_c"][i], o["r"][a], Jfull)
        Jmine = np.concatenate([Js[a, k] for k in range(m)] + [Jrho[a].reshape(2, 1)], axis=1)
        assert np.abs(Jmine - J2).max() <= parity.TOL * np.abs(J2).max()
        assert np.abs(out["r"][i] - r2).max() <= parity.CAM_R_TOL


def test_newton_rows_without_noise_equal_static_rows_and_unsupported_modes():
    """Exact observations: the first Newton step is below half a row, so every row is the static row inside a wider span."""
    cfg = syn.make_config("C3", scale=0.002)
    c = cfg["cam"]
    traj = kto.Traj(kto.SE3, cfg["dt"], cfg["t0"], cfg["knots"])
    # regenerate exact (noise-free, converged) observations with the oracle's Newton projection
    cam0 = kto.Camera(c["rows"], c["cols"], c["readout"], K=c["K"], method="newton")
    o = kto.static_rs_residuals(traj, cam0, c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"], c["rho"], jac_mode=0, cap=24)
    c["obs_uv"] = c["obs_uv"] - o["r"]                            # uv - r = projection
    c["q_ct"], c["p_ct"] = (0, 0, 0, 1), (0, 0, 0)
    p = _lib.Problem(0)
    p.set_se3_spline(cfg["dt"], cfg["t0"], len(cfg["knots"]))
    cam = _lib.make_camera(c["rows"], c["cols"], c["readout"], c["K"])
    gs = p.add_static_rs(cam, c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"])
    gn = p.add_newton_rs(cam, c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"])
    outs = p.evaluate(cfg["knots"], c["rho"])
    s, n = outs[gs], outs[gn]
    assert np.abs(n["r"] - s["r"]).max() < 1e-6 and np.abs(s["r"]).max() < 0.5
    k = s["i0_b"] - n["i0_b"]
    assert (k >= 0).all() and (k + 4 <= (n["J"].shape[1] - 58) // 14).all()
    for i in range(0, len(k), 5):
        a, b = n["J"][i, 56 + 14 * k[i]:56 + 14 * (k[i] + 4)], s["J"][i, 56:112]
        assert np.abs(a - b).max() <= 1e-6 * np.abs(b).max()
    _check_span_rows_local(p, gn, cfg["knots"], c["rho"], _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS, 2)
    with pytest.raises(NotImplementedError):      # the time offset of these measurements stays locked (the reference moves both spans with it)
        p.add_newton_rs(_lib.make_camera(c["rows"], c["cols"], c["readout"], c["K"], time_offset_locked=False), c["obs_uv"], c["obs_t0"], c["ref_uv"],
                        c["ref_t0"], c["lm_idx"])


@pytest.mark.parametrize("split", [False, True])
def test_position_rows_match_oracle(split):
    """PositionMeasurement (measurements/position_measurement.h) through the C ABI, SE3 and split trajectories; ragged tile."""
    rng = np.random.default_rng(12)
    n = 1000 + 17
    p = _lib.Problem(0)
    if not split:
        knots = syn.smooth_se3_knots(300, 0.05)
        traj = kto.Traj(kto.SE3, 0.05, 0.0, knots)
        p.set_se3_spline(0.05, 0.0, len(knots))
        tmax = 0.05 * (len(knots) - 3)
    else:
        k = syn.smooth_se3_knots(300, 0.05)
        vecs, quats = k[:, 4:7].copy(), syn.smooth_se3_knots(380, 0.04)[:, :4].copy()
        traj = kto.Traj(kto.SPLIT, 0.05, 0.0, vecs, 0.04, 0.01, quats)
        p.set_split_spline(0.05, 0.0, len(vecs), 0.04, 0.01, len(quats))
        knots = (vecs, quats)
        tmax = min(0.05 * (len(vecs) - 3), 0.01 + 0.04 * (len(quats) - 3))
    t, y, w = rng.uniform(0.02, tmax - 1e-6, n), rng.uniform(-5, 5, (n, 3)), rng.uniform(0.5, 2, n)
    g = p.add_position(t, y, w)
    assert p.group_kind(g) == _lib.POSITION and p.group_row_size(g) == (36 if split else 84)
    out = p.evaluate(knots)[g]
    o = kto.imu_residuals(traj, kto.Sensor(), 2, t, y, w, jac_mode=2)
    assert (out["i0"] == o["i0_a"]).all()
    assert parity.rel_err(out["r"], o["r"]) < parity.TOL
    sa = 3 if split else 7
    assert parity.rel_err(out["J"].reshape(n, 4, 3, sa), o["Ja"][:, :4]) < parity.TOL
    if split:
        assert (out["i0_c"] == o["ids_b"][:, 0]).all() and not o["Jb"].any()
    ids, nids = p.get_structure(g, cap=8)
    assert (ids[:, :4] == o["ids_a"][:, :4]).all() and (nids == 4).all()
    with pytest.raises(ValueError):
        p2 = _lib.Problem(0)
        p2.set_se3_spline(0.05, 0.0, 300)
        p2.add_position([tmax + 1.0], np.zeros((1, 3)))
        p2.evaluate(syn.smooth_se3_knots(300, 0.05))


def _rotated_orientations(q_xyzw, rng):
    """q * dq with dq a rotation by 0.05 .. 2.5 rad about a random axis; every second sign-flipped, every third rescaled."""
    n = len(q_xyzw)
    ang = rng.uniform(0.05, 2.5, n)
    ax = rng.normal(size=(n, 3)); ax /= np.linalg.norm(ax, axis=1)[:, None]
    dq = np.concatenate([ax * np.sin(ang / 2)[:, None], np.cos(ang / 2)[:, None]], axis=1)
    x1, y1, z1, w1 = q_xyzw.T; x2, y2, z2, w2 = dq.T
    qm = np.stack([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
                   w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2], axis=1)
    qm[::2] *= -1.0
    qm[::3] *= 1.7
    return qm, ang


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("local", [False, True])
def test_orientation_rows_match_oracle(split, local):
    """OrientationMeasurement (measurements/orientation_measurement.h:27-31, :57-79) through the C ABI: ONE residual per row (Eigen's
    angularDistance), rows [4][1][7] on SE3 and [4 SO3 knots][1][4] on a split trajectory; ragged tile; tangent rows = ambient rows x Plus."""
    rng = np.random.default_rng(21)
    n = 700 + 13
    p = _lib.Problem(0)
    if not split:
        knots = syn.smooth_se3_knots(300, 0.05)
        traj = kto.Traj(kto.SE3, 0.05, 0.0, knots)
        p.set_se3_spline(0.05, 0.0, len(knots))
        tmax = 0.05 * (len(knots) - 3)
    else:
        k = syn.smooth_se3_knots(300, 0.05)
        vecs, quats = k[:, 4:7].copy(), syn.smooth_se3_knots(380, 0.04)[:, :4].copy()
        traj = kto.Traj(kto.SPLIT, 0.05, 0.0, vecs, 0.04, 0.01, quats)
        p.set_split_spline(0.05, 0.0, len(vecs), 0.04, 0.01, len(quats))
        knots = (vecs, quats)
        tmax = min(0.05 * (len(vecs) - 3), 0.01 + 0.04 * (len(quats) - 3))
    t = rng.uniform(0.02, tmax - 1e-6, n)
    qm, ang = _rotated_orientations(kto.traj_evaluate(traj, t, 0xff)["orientation"], rng)
    g = p.add_orientation(t, qm)
    assert p.group_kind(g) == _lib.ORIENTATION and p.group_row_size(g) == (16 if split else 28)
    o = kto.imu_residuals(traj, kto.Sensor(), 3, t, qm, jac_mode=2)
    assert np.abs(o["r"][:, 0] - ang).max() < 1e-9
    sa = 4 if split else 7
    Jo = (o["Jb"] if split else o["Ja"])[:, :4]                        # (n, 4, 1, sa)
    if not local:
        out = p.evaluate(knots)[g]
        assert out["r"].shape == (n, 1) and np.abs(out["r"] - o["r"]).max() < parity.TOL
        assert parity.rel_err(out["J"].reshape(n, 4, 1, sa), Jo) < parity.TOL
    else:
        out = p.evaluate(knots, None, _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS | _lib.EVAL_LOCAL)[g]
        from kontiki_b200.estimator import _quat_plus_jacobian, _se3_plus_jacobian
        i0 = o["i0_b"] if split else o["i0_a"]
        P = _quat_plus_jacobian(quats) if split else _se3_plus_jacobian(knots)
        want = np.stack([np.einsum("nra,nad->nrd", Jo[:, k], P[i0 + k]) for k in range(4)], axis=1)
        assert np.abs(out["r"] - o["r"]).max() < parity.TOL
        assert parity.rel_err(out["J"].reshape(n, 4, 1, -1), want) < parity.TOL
    if split:
        assert (out["i0_c"] == o["i0_b"]).all() and (out["i0"] == o["ids_a"][:, 0]).all() and not o["Ja"].any()
        ids, nids = p.get_structure(g, cap=8)
        assert (ids[:, :4] == o["ids_a"][:, :4]).all() and (nids == 4).all()
    else:
        assert (out["i0"] == o["i0_a"]).all()
        ids, nids = p.get_structure(g, cap=8)
        assert (ids[:, :4] == o["ids_a"][:, :4]).all() and (nids == 4).all()
    with pytest.raises(ValueError):
        p2 = _lib.Problem(0)
        p2.set_se3_spline(0.05, 0.0, 300)
        p2.add_orientation([tmax + 1.0], np.array([[0.0, 0.0, 0.0, 1.0]]))
        p2.evaluate(syn.smooth_se3_knots(300, 0.05))


@pytest.mark.parametrize("atan,robust", [(False, False), (True, True)])
def test_lifting_rows_match_oracle(atan, robust):
    """LiftingRsCameraMeasurement (lifting_rscamera_measurement.h) through the C ABI: 3 residuals, rows [ref 4x(3x7) | obs W x(3x7) | vt 3 | rho 3];
    at the initial row times (where rows 0..1 are the static residuals and row 2 is zero) and after ktk_set_group_vt has moved them."""
    cfg = syn.make_config("C3", scale=0.004)
    c = cfg["cam"]
    rng = np.random.default_rng(35)
    n = len(c["lm_idx"])
    out_l = rng.random(n) < 0.2
    c["obs_uv"][out_l] += rng.normal(0, 40, (out_l.sum(), 2))
    c["weight"] = rng.uniform(0.5, 2, n)
    p, g, ocam = _camera_group(cfg, "lifting", atan)
    assert p.group_kind(g) == _lib.LIFTING_RS
    row = p.group_row_size(g)
    W = (row - 90) // 21
    assert (row - 90) % 21 == 0 and W >= 4
    flags = _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS | (_lib.EVAL_ROBUST if robust else 0)
    traj = kto.Traj(kto.SE3, cfg["dt"], cfg["t0"], cfg["knots"])
    vt0 = c["obs_uv"][:, 1] / c["rows"]
    for vt in (None, np.clip(vt0 + rng.uniform(-0.3, 0.3, n), 0.0, 1.0)):
        if vt is not None:
            p.set_group_vt(g, vt)
        out = p.evaluate(cfg["knots"], c["rho"], flags)[g]
        o = kto.lifting_rs_residuals(traj, ocam, c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"], c["rho"], vt=vt, weight=c["weight"], jac_mode=2, cap=24)
        assert out["r"].shape == (n, 3) and out["J"].shape[1] == row
        assert (out["i0"] == o["i0_ref_a"]).all()                                               # bit-exact indexing (reference window)
        ids, nids = p.get_structure(g, cap=24)
        assert (ids == o["ids_a"]).all()
        assert ((o["i0_obs_a"] >= out["i0_b"]) & (o["i0_obs_a"] + 4 <= out["i0_b"] + W)).all()     # the active window lies inside the row's span
        Js = p.expand_static_rs(g, ids, out["J"], out["i0"], out["i0_b"])                        # (n, cap, 3, 7)
        Jvt, Jrho = out["J"][:, row - 6:row - 3], out["J"][:, row - 3:row]
        if not robust:
            assert np.abs(out["r"] - o["r"]).max() < parity.CAM_R_TOL
            assert parity.rel_err(Js, o["Ja"]) < parity.TOL and parity.rel_err(Jvt, o["Jvt"]) < parity.TOL and parity.rel_err(Jrho, o["Jrho"]) < parity.TOL
            if vt is None:
                assert not out["r"][:, 2].any()
        else:
            for i in range(0, n, 7):
                m = int(nids[i])
                Jfull = np.concatenate([o["Ja"][i, k] for k in range(m)] + [o["Jvt"][i].reshape(3, 1), o["Jrho"][i].reshape(3, 1)], axis=1)
                _, r2, J2 = kto.huber_correct(c["huber_c"][i], o["r"][i], Jfull)
                Jmine = np.concatenate([Js[i, k] for k in range(m)] + [Jvt[i].reshape(3, 1), Jrho[i].reshape(3, 1)], axis=1)
                assert np.abs(Jmine - J2).max() <= parity.TOL * np.abs(J2).max()
                assert np.abs(out["r"][i] - r2).max() <= parity.CAM_R_TOL
        dev = p.evaluate(cfg["knots"], c["rho"], flags | _lib.EVAL_DEVICE_ORDER)[g]
        order = p.get_row_order(g)
        assert np.array_equal(dev["J"], out["J"][order]) and np.array_equal(dev["r"], out["r"][order])
    _check_span_rows_local(p, g, cfg["knots"], c["rho"], flags, 3)



@pytest.mark.gpu
@pytest.mark.parametrize("method,atan,robust", [("newton", False, False), ("newton", True, True), ("lifting", False, True), ("lifting", True, False)])
def test_span_camera_unlocked_relative_pose_matches_oracle(method, atan, robust):
    """NewtonRs / LiftingRs rows with the camera's relative pose unlocked (sensors.h:135-165): KTK_EVAL_SENSOR_JACOBIANS fills
    Js = [d r/d q_ct (nres x 4, ambient) | d r/d p_ct (nres x 3) | time offset (zero: it stays locked)] (k_span_sensor, forward mode through both sides of
    the row) -- against the oracle's autodiff over the sensor blocks; the knot / rho columns are untouched by the flag."""
    cfg = syn.make_config("C3", scale=0.003)
    c = cfg["cam"]
    rng = np.random.default_rng(41)
    n = len(c["lm_idx"])
    out_l = rng.random(n) < 0.2
    c["obs_uv"][out_l] += rng.normal(0, 40, (out_l.sum(), 2))
    c["obs_uv"] += rng.normal(0, 1.0, c["obs_uv"].shape)
    c["weight"] = rng.uniform(0.5, 2, n)
    kw = dict(wc=(0.02, -0.01), gamma=0.9) if atan else {}
    q_ct, p_ct = fx.so3_exp_xyzw(np.array([0.1, -0.2, 0.05])), np.array([0.05, -0.02, 0.1])
    p = _lib.Problem(0)
    p.set_se3_spline(cfg["dt"], cfg["t0"], len(cfg["knots"]))
    cam = _lib.make_camera(c["rows"], c["cols"], c["readout"], c["K"], q_ct=q_ct, p_ct=p_ct, q_locked=False, p_locked=False, **kw)
    add = p.add_newton_rs if method == "newton" else p.add_lifting_rs
    g = add(cam, c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"], c["weight"], c["huber_c"])
    ocam = kto.Camera(c["rows"], c["cols"], c["readout"], K=c["K"], method="newton" if method == "newton" else "static", q_ct=q_ct, p_ct=p_ct,
                      q_locked=False, p_locked=False, **kw)
    traj = kto.Traj(kto.SE3, cfg["dt"], cfg["t0"], cfg["knots"])
    nres = 2 if method == "newton" else 3
    vt = None
    if method == "lifting":
        vt = np.clip(c["obs_uv"][:, 1] / c["rows"] + rng.uniform(-0.2, 0.2, n), 0.0, 1.0)
        p.set_group_vt(g, vt)
        o = kto.lifting_rs_residuals(traj, ocam, c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"], c["rho"], vt=vt, weight=c["weight"], jac_mode=2, cap=24)
    else:
        o = kto.static_rs_residuals(traj, ocam, c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"], c["rho"], c["weight"], jac_mode=2, cap=24)
    base = _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS | (_lib.EVAL_ROBUST if robust else 0)
    plain = p.evaluate(cfg["knots"], c["rho"], base)[g]
    out = p.evaluate(cfg["knots"], c["rho"], base | _lib.EVAL_SENSOR_JACOBIANS)[g]
    assert np.array_equal(out["J"], plain["J"]) and np.array_equal(out["r"], plain["r"])
    assert out["Js"].shape == (n, 8 * nres)
    Jo = o["Js"].copy()
    if robust:
        for i in range(n):
            cols = np.concatenate([Jo[i, :4 * nres].reshape(nres, 4), Jo[i, 4 * nres:7 * nres].reshape(nres, 3), Jo[i, 7 * nres:].reshape(nres, 1)], axis=1)
            _, _, J2 = kto.huber_correct(c["huber_c"][i], o["r"][i], cols)
            Jo[i] = np.concatenate([J2[:, :4].reshape(-1), J2[:, 4:7].reshape(-1), J2[:, 7].reshape(-1)])
    assert np.abs(Jo[:, :7 * nres]).max() > 1.0
    assert parity.rel_err(out["Js"][:, None, :7 * nres], Jo[:, None, :7 * nres]) < parity.TOL
    assert not out["Js"][:, 7 * nres:].any()
    dev = p.evaluate(cfg["knots"], c["rho"], base | _lib.EVAL_SENSOR_JACOBIANS | _lib.EVAL_DEVICE_ORDER)[g]
    assert np.array_equal(dev["Js"], out["Js"][p.get_row_order(g)])


@pytest.mark.gpu
@pytest.mark.parametrize("method,same_grid,robust", [("newton", True, False), ("newton", False, True), ("lifting", False, False), ("lifting", True, True)])
def test_span_camera_rows_on_split_trajectory_match_oracle(method, same_grid, robust):
    """NewtonRsCameraMeasurement / LiftingRsCameraMeasurement on a SplitTrajectory (python/src/kontiki/measurements/measurement_defs.h:40-85: every
    measurement is instantiated with every trajectory) through the C ABI: k_landmark_ref_split + k_span_rs_split (forward mode, one thread per
    row and direction), rows [ref R3 | ref SO3 | obs R3 span | obs SO3 span | (vt) | rho], windows bit-exact, structure vs the oracle's block lists."""
    c = _split_case(n_lm=120, scale_imu=10, same_grid=same_grid)
    cam = c["cam"]
    rng = np.random.default_rng(5)
    n = len(cam["lm_idx"])
    cam["obs_uv"] += rng.normal(0, 1.0, cam["obs_uv"].shape)
    cam["obs_uv"][:, 1] = np.clip(cam["obs_uv"][:, 1], 0, cam["rows"] - 1e-6)
    p = _lib.Problem(0)
    p.set_split_spline(c["dt_a"], c["t0_a"], len(c["vecs"]), c["dt_b"], c["t0_b"], len(c["quats"]))
    q_ct, p_ct = fx.so3_exp_xyzw(np.array([0.1, -0.2, 0.05])), np.array([0.05, -0.02, 0.1])
    add = p.add_newton_rs if method == "newton" else p.add_lifting_rs
    g = add(_lib.make_camera(cam["rows"], cam["cols"], cam["readout"], cam["K"], q_ct=q_ct, p_ct=p_ct), cam["obs_uv"], cam["obs_t0"], cam["ref_uv"],
            cam["ref_t0"], cam["lm_idx"], cam["weight"], cam["huber_c"])
    traj = kto.Traj(kto.SPLIT, c["dt_a"], c["t0_a"], c["vecs"], c["dt_b"], c["t0_b"], c["quats"])
    ocam = kto.Camera(cam["rows"], cam["cols"], cam["readout"], K=cam["K"], method="newton" if method == "newton" else "static", q_ct=q_ct, p_ct=p_ct)
    nres = 2 if method == "newton" else 3
    args = (cam["obs_uv"], cam["obs_t0"], cam["ref_uv"], cam["ref_t0"], cam["lm_idx"], cam["rho"])
    if method == "lifting":
        vt = np.clip(cam["obs_uv"][:, 1] / cam["rows"] + rng.uniform(-0.2, 0.2, n), 0.0, 1.0)
        p.set_group_vt(g, vt)
        o = kto.lifting_rs_residuals(traj, ocam, *args, vt=vt, weight=cam["weight"], jac_mode=2, cap=24)
        tail_o = np.concatenate([o["Jvt"], o["Jrho"]], axis=1)
    else:
        o = kto.static_rs_residuals(traj, ocam, *args, cam["weight"], jac_mode=2, cap=24)
        tail_o = o["Jrho"]
    flags = _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS | (_lib.EVAL_ROBUST if robust else 0)
    out = p.evaluate((c["vecs"], c["quats"]), cam["rho"], flags)[g]
    Wa, Wb = p.group_span_windows(g)
    assert out["J"].shape == (n, nres * (28 + 3 * Wa + 4 * Wb) + (2 if nres == 2 else 6)) and out["r"].shape == (n, nres)
    assert (out["i0"] == o["i0_ref_a"]).all() and (out["i0_c"] == o["i0_ref_b"]).all()                  # bit-exact indexing of the reference window
    ids_a, _ = p.get_structure(g, cap=24)
    ids_b, _ = p.get_structure_so3(g, cap=24)
    assert (ids_a == o["ids_a"]).all() and (ids_b == o["ids_b"]).all()
    idx = np.stack([out["i0"], out["i0_b"], out["i0_c"], out["i0_d"]], axis=1)
    Ja, Jb, tail = parity.scatter_span_split(out["J"], idx, o["ids_a"], o["ids_b"], Wa, Wb, nres)
    if not robust:
        assert np.abs(out["r"] - o["r"]).max() < parity.CAM_R_TOL
        assert parity.rel_err(Ja, o["Ja"]) < parity.TOL and parity.rel_err(Jb, o["Jb"]) < parity.TOL and parity.rel_err(tail, tail_o) < parity.TOL
    else:
        n_out = 0
        for i in range(0, n, 3):
            ma, mb = int((o["ids_a"][i] >= 0).sum()), int((o["ids_b"][i] >= 0).sum())
            Jfull = np.concatenate([o["Ja"][i, k] for k in range(ma)] + [o["Jb"][i, k] for k in range(mb)] + [tail_o[i].reshape(-1, nres).T], axis=1)
            _, r2, J2 = kto.huber_correct(cam["huber_c"][i], o["r"][i], Jfull)
            Jmine = np.concatenate([Ja[i, k] for k in range(ma)] + [Jb[i, k] for k in range(mb)] + [tail[i].reshape(-1, nres).T], axis=1)
            assert np.abs(Jmine - J2).max() <= parity.TOL * np.abs(J2).max()
            assert np.abs(out["r"][i] - r2).max() <= parity.CAM_R_TOL
            n_out += np.linalg.norm(o["r"][i]) > cam["huber_c"][i]
        assert n_out > 3
    dev = p.evaluate((c["vecs"], c["quats"]), cam["rho"], flags | _lib.EVAL_DEVICE_ORDER)[g]
    order = p.get_row_order(g)
    assert np.array_equal(dev["J"], out["J"][order]) and np.array_equal(dev["r"], out["r"][order]) and np.array_equal(dev["i0_d"], out["i0_d"][order])
    # KTK_EVAL_LOCAL: R3 blocks unchanged, SO3 blocks times dPlus/ddelta of their knots (k_span_localize_split)
    from kontiki_b200.estimator import _quat_plus_jacobian
    loc = p.evaluate((c["vecs"], c["quats"]), cam["rho"], flags | _lib.EVAL_LOCAL)[g]
    tl = 2 if nres == 2 else 6
    assert loc["J"].shape == (n, nres * 3 * (8 + Wa + Wb) + tl) and np.array_equal(loc["r"], out["r"])
    Pq = _quat_plus_jacobian(c["quats"])
    A, L = out["J"], loc["J"]
    assert np.array_equal(L[:, :nres * 12], A[:, :nres * 12]) and np.array_equal(L[:, -tl:], A[:, -tl:])
    assert np.array_equal(L[:, nres * 24:nres * (24 + 3 * Wa)], A[:, nres * 28:nres * (28 + 3 * Wa)])
    kr = out["i0_c"][:, None] + np.arange(4)
    ko = np.minimum(out["i0_d"][:, None] + np.arange(Wb), len(c["quats"]) - 1)
    ref = np.einsum("nkra,nkad->nkrd", A[:, nres * 12:nres * 28].reshape(n, 4, nres, 4), Pq[kr])
    assert parity.rel_err(L[:, nres * 12:nres * 24].reshape(n, 4, nres, 3), ref) < 1e-12
    ref = np.einsum("nkra,nkad->nkrd", A[:, nres * (28 + 3 * Wa):-tl].reshape(n, Wb, nres, 4), Pq[ko])
    assert parity.rel_err(L[:, nres * (24 + 3 * Wa):-tl].reshape(n, Wb, nres, 3), ref) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("method,robust", [("newton", True), ("lifting", False)])
def test_span_camera_unlocked_relative_pose_on_split_trajectory(method, robust):
    """... and the same sensor blocks when the trajectory is a SplitTrajectory (k_span_sensor, split branch)."""
    c = _split_case(n_lm=100, scale_imu=10, same_grid=False)
    cam = c["cam"]
    rng = np.random.default_rng(6)
    n = len(cam["lm_idx"])
    cam["obs_uv"] += rng.normal(0, 1.0, cam["obs_uv"].shape)
    cam["obs_uv"][:, 1] = np.clip(cam["obs_uv"][:, 1], 0, cam["rows"] - 1e-6)
    p = _lib.Problem(0)
    p.set_split_spline(c["dt_a"], c["t0_a"], len(c["vecs"]), c["dt_b"], c["t0_b"], len(c["quats"]))
    q_ct, p_ct = fx.so3_exp_xyzw(np.array([0.1, -0.2, 0.05])), np.array([0.05, -0.02, 0.1])
    add = p.add_newton_rs if method == "newton" else p.add_lifting_rs
    g = add(_lib.make_camera(cam["rows"], cam["cols"], cam["readout"], cam["K"], q_ct=q_ct, p_ct=p_ct, q_locked=False, p_locked=False), cam["obs_uv"],
            cam["obs_t0"], cam["ref_uv"], cam["ref_t0"], cam["lm_idx"], cam["weight"], cam["huber_c"])
    traj = kto.Traj(kto.SPLIT, c["dt_a"], c["t0_a"], c["vecs"], c["dt_b"], c["t0_b"], c["quats"])
    ocam = kto.Camera(cam["rows"], cam["cols"], cam["readout"], K=cam["K"], method="newton" if method == "newton" else "static", q_ct=q_ct, p_ct=p_ct,
                      q_locked=False, p_locked=False)
    nres = 2 if method == "newton" else 3
    args = (cam["obs_uv"], cam["obs_t0"], cam["ref_uv"], cam["ref_t0"], cam["lm_idx"], cam["rho"])
    if method == "lifting":
        vt = np.clip(cam["obs_uv"][:, 1] / cam["rows"] + rng.uniform(-0.2, 0.2, n), 0.0, 1.0)
        p.set_group_vt(g, vt)
        o = kto.lifting_rs_residuals(traj, ocam, *args, vt=vt, weight=cam["weight"], jac_mode=2, cap=24)
    else:
        o = kto.static_rs_residuals(traj, ocam, *args, cam["weight"], jac_mode=2, cap=24)
    flags = _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS | _lib.EVAL_SENSOR_JACOBIANS | (_lib.EVAL_ROBUST if robust else 0)
    out = p.evaluate((c["vecs"], c["quats"]), cam["rho"], flags)[g]
    Jo = o["Js"].copy()
    if robust:
        for i in range(n):
            cols = np.concatenate([Jo[i, :4 * nres].reshape(nres, 4), Jo[i, 4 * nres:7 * nres].reshape(nres, 3), Jo[i, 7 * nres:].reshape(nres, 1)], axis=1)
            _, _, J2 = kto.huber_correct(cam["huber_c"][i], o["r"][i], cols)
            Jo[i] = np.concatenate([J2[:, :4].reshape(-1), J2[:, 4:7].reshape(-1), J2[:, 7].reshape(-1)])
    assert out["Js"].shape == (n, 8 * nres) and np.abs(Jo[:, :7 * nres]).max() > 1.0
    assert parity.rel_err(out["Js"][:, None, :7 * nres], Jo[:, None, :7 * nres]) < parity.TOL
    assert not out["Js"][:, 7 * nres:].any()

"""Pins the CPU oracle against every value-level check the reference's own test-suite holds for the
hot path.  Each test names the reference test it restates.  The reference has
no golden vectors and never reads a Jacobian, so Jacobians are additionally cross-checked here by
central differences of the oracle's own double path (self-consistency, not reference parity).
"""
import numpy as np
import pytest
from numpy.testing import assert_almost_equal, assert_allclose
from scipy.interpolate import BSpline

import fixtures_ref as fx
from oracle import kto

P, V, A, Q, W = kto.EvalPosition, kto.EvalVelocity, kto.EvalAcceleration, kto.EvalOrientation, kto.EvalAngularVelocity


def make_traj(name, **kw):
    if name == "r3":
        return kto.Traj(kto.R3, fx.R3_DT, fx.R3_T0, fx.R3_KNOTS, **kw)
    if name == "so3":
        return kto.Traj(kto.SO3, dt_b=fx.SO3_DT, t0_b=fx.SO3_T0, knots_b=fx.SO3_KNOTS, **kw)
    if name == "se3":
        return kto.Traj(kto.SE3, fx.SE3_DT, fx.SE3_T0, fx.SE3_KNOTS, **kw)
    if name == "split":
        return kto.Traj(kto.SPLIT, fx.R3_DT, fx.R3_T0, fx.R3_KNOTS, fx.SO3_DT, fx.SO3_T0, fx.SO3_KNOTS, **kw)
    raise KeyError(name)


ALL = ["r3", "so3", "se3", "split"]


def safe_time_span(traj, length):
    """python/kontiki/utils.py safe_time_span: a random span of given length inside the valid time."""
    t1 = traj.min_time
    t2 = traj.max_time
    assert t2 - t1 >= length
    return t1, t1 + length


# --- trajectories/test_spline_trajectories.py:181-219 -------------------------------------------------
def test_r3_matches_scipy_bspline():
    traj = make_traj("r3")
    n, dt, t0 = len(fx.R3_KNOTS), fx.R3_DT, fx.R3_T0
    knots = t0 + dt * np.arange(-3, n + 1)            # uniform knot vector; spline valid on [t0, t0+(n-3)dt)
    spl = BSpline(knots, fx.R3_KNOTS, 3)
    t = np.linspace(traj.min_time, traj.max_time - 1e-6, 57)
    out = kto.traj_evaluate(traj, t, P | V | A)
    ts = t                                             # same convention as the reference's scipy_bspline helper (:10-14)
    assert_allclose(out["position"], spl(ts), atol=1e-12)
    assert_allclose(out["velocity"], spl.derivative(1)(ts), atol=1e-12)
    assert_allclose(out["acceleration"], spl.derivative(2)(ts), atol=1e-12)


# --- trajectories/test_general.py:155-161 (velocity ~ central difference of position, all trajectories)
@pytest.mark.parametrize("name", ["r3", "se3", "split"])
def test_velocity_numerical(name):
    traj = make_traj(name)
    times = np.linspace(traj.min_time + 0.01, traj.max_time - 0.01, 40)
    h = 1e-6
    v = kto.traj_evaluate(traj, times, V)["velocity"]
    p1 = kto.traj_evaluate(traj, times - h, P)["position"]
    p2 = kto.traj_evaluate(traj, times + h, P)["position"]
    assert_allclose(v, (p2 - p1) / (2 * h), atol=1e-5)      # reference: decimal=3


# --- trajectories/test_general.py:164-174 (xfail for SE3 in the reference; passes with the intended dB) ---
@pytest.mark.parametrize("name", ["r3", "se3", "split"])
def test_acceleration_numerical(name):
    traj = make_traj(name)
    times = np.linspace(traj.min_time + 0.01, traj.max_time - 0.01, 40)
    h = 1e-5
    a = kto.traj_evaluate(traj, times, A)["acceleration"]
    v1 = kto.traj_evaluate(traj, times - h, V)["velocity"]
    v2 = kto.traj_evaluate(traj, times + h, V)["velocity"]
    assert_allclose(a, (v2 - v1) / (2 * h), atol=1e-5)


def test_se3_accel_compat_zero_dB_differs_from_intended():
    """with the reference's Jet-path behaviour (dB == 0) the SE3 'acceleration'
    is NOT the second derivative of position -- which is what the reference's xfail hides."""
    times = np.array([3.7])
    a_int = kto.traj_evaluate(make_traj("se3"), times, A | Q)["acceleration"]
    a_cmp = kto.traj_evaluate(make_traj("se3", compat_zero_dB=True), times, A | Q)["acceleration"]
    assert np.abs(a_int - a_cmp).max() > 1e-2
    # plain trajectory.acceleration(t) requests only EvalAcceleration -> same flags class (num_derivatives=2, dB unset)
    h = 1e-4
    p = kto.traj_evaluate(make_traj("se3"), np.array([3.7 - h, 3.7, 3.7 + h]), P)["position"]
    assert_allclose(a_int[0], (p[0] - 2 * p[1] + p[2]) / h ** 2, atol=1e-5)


# --- trajectories/test_general.py:177-189 (w_world ~ 2 dq q^-1, SO3 / SE3 / Split) ---------------------
@pytest.mark.parametrize("name", ["so3", "se3", "split"])
def test_angular_velocity_numerical(name):
    traj = make_traj(name)
    times = np.linspace(traj.min_time + 0.01, traj.max_time - 0.01, 40)
    h = 1e-6
    w = kto.traj_evaluate(traj, times, W)["angular_velocity"]
    q = kto.traj_evaluate(traj, times, Q)["orientation"]
    q1 = kto.traj_evaluate(traj, times - h, Q)["orientation"]
    q2 = kto.traj_evaluate(traj, times + h, Q)["orientation"]
    for i in range(len(times)):
        dq = (q2[i] - q1[i]) / (2 * h)
        w_num = 2 * fx.qmul_xyzw(dq, fx.qconj_xyzw(q[i]))
        assert_allclose(w[i], w_num[:3], atol=1e-5)          # reference: decimal=4
        assert abs(w_num[3]) < 1e-5


# --- conftest.py:52-81 + trajectories/test_general.py:40-67,107-120: the only closed-form known answer ---
@pytest.mark.parametrize("name", ["so3", "split"])
def test_so3_constant_rate_known_answer(name):
    traj = make_traj(name)
    times = np.linspace(traj.min_time, traj.max_time - 1e-9, 33)
    out = kto.traj_evaluate(traj, times, Q | W)
    w_true = fx.SO3_RATE * fx.SO3_AXIS
    assert_allclose(out["angular_velocity"], np.tile(w_true, (len(times), 1)), atol=1e-12)
    for t, q in zip(times, out["orientation"]):
        # knot i sits at t0+(i-3)dt (conftest.py:55-56) while a cubic B-spline places it at t0+(i-1)dt: constant lag 2dt
        # (the reference comments its orientation example out, test_general.py:61; the rate check above is the pinned part)
        th = fx.SO3_RATE * (t - 2 * fx.SO3_DT)
        q_true = np.concatenate([np.sin(th / 2) * fx.SO3_AXIS, [np.cos(th / 2)]])
        assert_allclose(q, q_true, atol=1e-12)


def test_se3_constant_twist_known_answer():
    """Same idea for SE3 (not in the reference, whose SE3 example data is xfail test_general.py:79-80):
    knots on a one-parameter subgroup exp(tau*xi) must reproduce exp(t*xi) and its body twist exactly."""
    xi_w, xi_v = np.array([0.2, -0.1, 0.3]), np.array([1.0, 0.5, -0.2])
    dt, t0, n = 0.5, 0.0, 12

    def pose(tau):
        w = xi_w * tau
        th = np.linalg.norm(w)
        Wm = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        Vm = np.eye(3) + (1 - np.cos(th)) / th ** 2 * Wm + (th - np.sin(th)) / th ** 3 * Wm @ Wm if th > 1e-12 else np.eye(3)
        return fx.so3_exp_xyzw(w), Vm @ (xi_v * tau)

    # cumulative cubic B-spline with knots exp((i-1)*dt*xi) interpolates exp((t-t0)... ) with a fixed lag of 1 knot:
    knots = np.array([np.concatenate(pose((i - 1) * dt)) for i in range(n)])
    traj = kto.Traj(kto.SE3, dt, t0, knots)
    times = np.linspace(0.0, (n - 3) * dt - 1e-9, 29)
    out = kto.traj_evaluate(traj, times, P | V | Q | W)
    for t, q, p, w, v in zip(times, out["orientation"], out["position"], out["angular_velocity"], out["velocity"]):
        q_true, p_true = pose(t)
        if np.dot(q, q_true) < 0:
            q_true = -q_true
        assert_allclose(q, q_true, atol=1e-12)
        assert_allclose(p, p_true, atol=1e-12)
        R = fx.rot_xyzw(q_true)
        assert_allclose(w, R @ xi_w, atol=1e-12)             # world-frame angular velocity
        assert_allclose(v, R @ xi_v, atol=1e-12)             # world-frame velocity = R * body velocity


# --- trajectories/test_general.py:133-152 pose convention x_w = q x_b + p --------------------------------
@pytest.mark.parametrize("name", ["se3", "split"])
def test_pose_convention(name):
    traj = make_traj(name)
    t = np.array([traj.min_time + 1.3])
    out = kto.traj_evaluate(traj, t, P | Q)
    R = fx.rot_xyzw(out["orientation"][0])
    assert_allclose(R @ R.T, np.eye(3), atol=1e-12)
    assert abs(np.linalg.det(R) - 1) < 1e-12


# --- test_measurements.py:108-126 gyro == R^T w_world (+bias) ----------------------------------------------
@pytest.mark.parametrize("name", ["so3", "se3", "split"])
@pytest.mark.parametrize("bias", [False, True])
def test_gyroscope_measurement(name, bias):
    traj = make_traj(name)
    rng = np.random.default_rng(5)
    imu = kto.Sensor(gbias=rng.uniform(-.1, .1, 3), abias=rng.uniform(-.1, .1, 3)) if bias else kto.Sensor()
    times = np.linspace(traj.min_time + 0.2, traj.max_time - 0.2, 15)
    y = rng.uniform(-1, 1, (15, 3))
    res = kto.imu_residuals(traj, imu, 0, times, y, jac_mode=0)
    out = kto.traj_evaluate(traj, times, Q | W)
    for i in range(15):
        R = fx.rot_xyzw(out["orientation"][i])
        expect = R.T @ out["angular_velocity"][i] + (imu.gbias if bias else 0)
        assert_allclose(y[i] - res["r"][i], expect, atol=1e-12)     # error = w - measure (weight 1)


# --- test_measurements.py:129-153 accel == R^T (a - (0,0,9.80665)); xfail for SE3 in the reference ---------------
@pytest.mark.parametrize("name", ["se3", "split"])
def test_accelerometer_measurement(name):
    traj = make_traj(name)
    imu = kto.Sensor()
    times = np.linspace(traj.min_time + 0.2, traj.max_time - 0.2, 15)
    y = np.zeros((15, 3))
    res = kto.imu_residuals(traj, imu, 1, times, y, jac_mode=0)
    out = kto.traj_evaluate(traj, times, Q | A)
    for i in range(15):
        R = fx.rot_xyzw(out["orientation"][i])
        expect = R.T @ (out["acceleration"][i] - np.array([0, 0, 9.80665]))
        assert_allclose(-res["r"][i], expect, atol=1e-9)


# --- test_measurements.py:73-89 error scales linearly with weight (exact) ------------------------------------
@pytest.mark.parametrize("which", [0, 1])
def test_weight_scaling(which):
    traj = make_traj("se3")
    times = np.linspace(traj.min_time + 0.2, traj.max_time - 0.2, 7)
    y = np.random.default_rng(1).uniform(-1, 1, (7, 3))
    r1 = kto.imu_residuals(traj, kto.Sensor(), which, times, y, jac_mode=0)["r"]
    r2 = kto.imu_residuals(traj, kto.Sensor(), which, times, y, weight=np.full(7, 2.0), jac_mode=0)["r"]
    assert np.array_equal(r2, 2.0 * r1)


# --- test_measurements.py:164-204 time offset: measure(t; d) == measure(t - (-d); 0) ---------------------------
def test_time_offset_equivalence():
    traj = make_traj("split")
    d = 0.05
    times = np.linspace(traj.min_time + 0.3, traj.max_time - 0.3, 9)
    y = np.zeros((9, 3))
    # offset unlocked => spans widen by +-max_time_offset so that t+d stays inside the segment
    r_off = kto.imu_residuals(traj, kto.Sensor(time_offset=d, d_locked=False), 0, times, y, jac_mode=0)["r"]
    r_ref = kto.imu_residuals(traj, kto.Sensor(), 0, times + d, y, jac_mode=0)["r"]
    assert_allclose(r_off, r_ref, atol=1e-13)


def test_locked_nonzero_offset_crossing_knot_raises():
    """SURVEY.md section 8b edge case (i): locked non-zero offset that moves t across a knot boundary -> range_error."""
    traj = make_traj("se3")
    t = np.array([fx.SE3_T0 + fx.SE3_DT - 0.01])
    with pytest.raises(kto.OracleError) as e:
        kto.imu_residuals(traj, kto.Sensor(time_offset=0.05, d_locked=True), 0, t, np.zeros((1, 3)), jac_mode=0)
    assert e.value.code == kto.RANGE_ERROR


# --- test_estimator.py:75,99 / test_imu.py:89 style: structure counts + the segment rule ----------------------
def test_structure_rule_imu_locked_gives_exactly_four_knots():
    traj = make_traj("se3")
    t = np.array([2.0, 4.0, 5.9])
    res = kto.imu_residuals(traj, kto.Sensor(), 0, t, np.zeros((3, 3)), jac_mode=0)
    for i, ti in enumerate(t):
        i1 = int(np.floor((ti - fx.SE3_T0) / fx.SE3_DT))
        assert list(res["ids_a"][i]) == [i1, i1 + 1, i1 + 2, i1 + 3]
        assert res["i0_a"][i] == i1


def test_structure_rule_two_spans_merge_and_split():
    # python transcription of spline_base.h:371-403 vs the C++ restatement
    def rule(dt, t0, spans):
        ids, segs = [], []
        start, end = 0, -1
        for ta, tb in spans:
            i1 = int(np.floor((ta - t0) / dt))
            i2 = int(np.floor((tb - t0) / dt))
            if i1 > end:
                segs.append([t0 + dt * i1, 0])
                start = i1
            else:
                i1 = end + 1
            for i in range(i1, i2 + 4):
                ids.append(i)
                segs[-1][1] += 1
            end = start + segs[-1][1] - 1
        return ids, segs
    rng = np.random.default_rng(3)
    for _ in range(200):
        dt, t0 = rng.uniform(0.01, 0.5), rng.uniform(-1, 1)
        a = t0 + rng.uniform(0, 5)          # CheckTimeSpans guarantees t >= MinTime = t0 (trajectory_estimator.h:106)
        b = a + rng.uniform(0, 1.0)
        spans = [(a, a + rng.uniform(0, 0.2)), (b, b + rng.uniform(0, 0.2))]
        ids, seg_t0, seg_n = kto.spline_structure(dt, t0, spans, cap=256)
        ids_py, segs_py = rule(dt, t0, spans)
        assert list(ids) == ids_py
        assert [s[1] for s in segs_py] == list(seg_n)
        assert [s[0] for s in segs_py] == list(seg_t0)        # bit-exact segment origin


# --- test_cameras.py:32-45 project(unproject(y)) == y ----------------------------------------------------------
# --- test_measurements.py:16-32 static RS reprojection consistency on RS-consistent synthetic structure ------
def _rs_project(traj, cam, X_world, t0):
    """fixtures/sfm_fixtures.py:12-31: exact rolling-shutter projection by root finding on the row time."""
    from scipy.optimize import brentq
    R_ct = fx.rot_xyzw(cam.q_ct)

    def project(t):
        ev = kto.traj_evaluate(traj, [t0 + t], P | Q)
        R = fx.rot_xyzw(ev["orientation"][0])
        X_traj = R.T @ (X_world - ev["position"][0])
        X_cam = R_ct @ X_traj + cam.p_ct
        if X_cam[2] <= 0:
            raise ValueError("Behind camera")
        return kto.camera_project(cam, X_cam)[0]

    def rootfunc(t):
        u, v = project(t)
        return t - v * cam.readout / cam.rows
    t = brentq(rootfunc, 0, cam.readout, xtol=1e-15)
    return project(t)


def make_sfm(traj, cam, rng, nviews=8, nlm=12):
    """fixtures/sfm_fixtures.py:34-84 restated on flat arrays."""
    fps = 30
    t1 = traj.min_time + 1e-2
    view_t0 = t1 + np.arange(nviews) / fps
    R_ct = fx.rot_xyzw(cam.q_ct)
    ref_uv, ref_t0, rho, obs = [], [], [], []
    tries = 0
    while len(rho) < nlm and tries < 5000:
        tries += 1
        i = rng.integers(0, nviews - 1)
        y0 = np.array([rng.uniform(0, cam.cols), rng.uniform(0, cam.rows)])
        z0 = rng.uniform(0.5, 100)
        X_cam = z0 * kto.camera_unproject(cam, y0)
        X_traj = R_ct.T @ (X_cam - cam.p_ct)
        t = view_t0[i] + y0[1] * cam.readout / cam.rows
        ev = kto.traj_evaluate(traj, [t], P | Q)
        X_world = fx.rot_xyzw(ev["orientation"][0]) @ X_traj + ev["position"][0]
        mine = []
        for j in range(i + 1, nviews):
            try:
                x, y = _rs_project(traj, cam, X_world, view_t0[j])
            except ValueError:
                continue
            if 0 <= x < cam.cols and 0 <= y < cam.rows:
                mine.append((j, x, y))
        if mine:
            lm = len(rho)
            rho.append(1 / z0); ref_uv.append(y0); ref_t0.append(view_t0[i])
            obs += [(lm, view_t0[j], x, y) for j, x, y in mine]
    obs = np.array(obs)
    lm_idx = obs[:, 0].astype(np.int32)
    return dict(lm_idx=lm_idx, obs_t0=obs[:, 1].copy(), obs_uv=obs[:, 2:4].copy(), ref_uv=np.array(ref_uv)[lm_idx], ref_t0=np.array(ref_t0)[lm_idx],
                rho=np.array(rho))


ATAN = dict(K=fx.ATAN_K, wc=fx.ATAN_WC, gamma=fx.ATAN_GAMMA)          # fixtures/camera_fixtures.py:12-16
PINHOLE = dict(K=np.array([[900., 0, 960], [0, 900, 540], [0, 0, 1]]))


def make_camera(model, method="static", **kw):
    return kto.Camera(fx.IMAGE_ROWS, fx.IMAGE_COLS, fx.CAMERA_READOUT, method=method, **(ATAN if model == "atan" else PINHOLE), **kw)


def _sfm_case(name, model, method, seed=11, **kw):
    rng = np.random.default_rng(seed)
    k = fx.smooth_se3_knots(40, 0.1)
    traj = kto.Traj(kto.SE3, 0.1, 0.0, k) if name == "se3" else kto.Traj(kto.SPLIT, 0.1, 0.0, k[:, 4:7].copy(), 0.1, 0.0, k[:, 0:4].copy())
    cam = make_camera(model, method, q_ct=fx.so3_exp_xyzw(np.array([0.1, -0.2, 0.05])), p_ct=np.array([0.05, -0.02, 0.1]))
    return traj, cam, make_sfm(traj, cam, rng, **kw)


# --- test_measurements.py:16-32 over projection_types x camera fixture (Pinhole, Atan) -------------------------
@pytest.mark.parametrize("name", ["se3", "split"])
@pytest.mark.parametrize("model", ["pinhole", "atan"])
@pytest.mark.parametrize("method", ["static", "newton", "lifting"])
def test_rscamera_measurements(name, model, method):
    traj, cam, s = _sfm_case(name, model, "static" if method == "lifting" else method)
    assert len(s["lm_idx"]) >= 10
    if method == "lifting":       # LiftingRsCameraMeasurement.project at its initial row time vt = v / rows (lifting_rscamera_measurement.h:68, :90-96)
        res = kto.lifting_rs_residuals(traj, cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"], jac_mode=0)
        assert res["r"].shape[1] == 3 and not res["r"][:, 2].any()
        res["r"] = res["r"][:, :2]
    else:
        res = kto.static_rs_residuals(traj, cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"], jac_mode=0)
    assert_almost_equal(res["r"], 0 * res["r"])       # np.testing.assert_almost_equal(yhat, obs.uv)  (decimal=7)


# --- test_measurements.py:34-54: Newton lands within half a row of the noise-free row ---------------------------
@pytest.mark.parametrize("model", ["pinhole", "atan"])
def test_newton_rscamera_measurements_with_noise(model):
    traj, cam, s = _sfm_case("se3", model, "newton", seed=12)
    noisy = s["obs_uv"] + np.random.default_rng(5).normal(0, 2.0, size=s["obs_uv"].shape)
    res = kto.static_rs_residuals(traj, cam, noisy, s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"], jac_mode=0)
    yhat = noisy - res["r"]
    assert np.abs(yhat[:, 1] - s["obs_uv"][:, 1]).max() <= 0.5


# --- test_cameras.py:32-37 project(unproject(y) * scale) == y;  :40-66 dy against numerical differentiation ---------
@pytest.mark.parametrize("model", ["pinhole", "atan"])
def test_camera_project_unproject_and_derivative(model):
    cam = make_camera(model)
    rng = np.random.default_rng(8)
    for _ in range(50):
        y = np.array([rng.uniform(0, cam.cols), rng.uniform(0, cam.rows)])
        X = kto.camera_unproject(cam, y) * rng.uniform(0.01, 10)
        assert_almost_equal(kto.camera_project(cam, X)[0], y)
        X = kto.camera_unproject(cam, y) * rng.uniform(3, 10)
        dX = X + rng.normal(size=3)
        _, dy = kto.camera_project(cam, X, dX)
        Jn = _numdiff(lambda x: kto.camera_project(cam, x)[0], X, h=1e-3)
        assert_almost_equal(Jn @ dX, dy, decimal=3)


@pytest.mark.parametrize("name", ["se3", "split"])
def test_static_rs_reprojection_consistency(name):
    rng = np.random.default_rng(11)
    if name == "se3":   # gentle motion so that landmarks stay in view (the reference uses its handcrafted fixture + many tries)
        traj = kto.Traj(kto.SE3, 0.1, 0.0, fx.smooth_se3_knots(40, 0.1))
    else:
        k = fx.smooth_se3_knots(40, 0.1)
        traj = kto.Traj(kto.SPLIT, 0.1, 0.0, k[:, 4:7].copy(), 0.1, 0.0, k[:, 0:4].copy())
    q_ct = fx.so3_exp_xyzw(np.array([0.1, -0.2, 0.05]))
    cam = kto.Camera(fx.IMAGE_ROWS, fx.IMAGE_COLS, fx.CAMERA_READOUT, K=np.array([[900., 0, 960], [0, 900, 540], [0, 0, 1]]),
                     q_ct=q_ct, p_ct=np.array([0.05, -0.02, 0.1]))
    s = make_sfm(traj, cam, rng)
    assert len(s["lm_idx"]) >= 10
    res = kto.static_rs_residuals(traj, cam, s["obs_uv"], s["obs_t0"], s["ref_uv"], s["ref_t0"], s["lm_idx"], s["rho"], jac_mode=0)
    # Static RS uses the OBSERVED row for the projection time, which is exact for noise-free RS-consistent data
    assert np.abs(res["r"]).max() < 1e-6           # reference: decimal=... "project(traj) ~ obs.uv"


def test_pinhole_project_unproject_identity():
    # exercised through the static-RS chain with ref == obs view and identity motion: y_hat == uv_ref
    n = 8
    knots = np.tile(np.array([0, 0, 0, 1, 0.3, -0.2, 0.5]), (n, 1)).astype(float)
    traj = kto.Traj(kto.SE3, 0.5, 0.0, knots)
    cam = kto.Camera(fx.IMAGE_ROWS, fx.IMAGE_COLS, fx.CAMERA_READOUT, K=np.array([[853.1, 0., 988.1], [0., 873.5, 525.7], [0., 0., 1.]]))
    uv = np.array([[100.5, 200.25], [1900.0, 1000.0]])
    t0 = np.array([0.6, 0.9])
    res = kto.static_rs_residuals(traj, cam, uv, t0, uv, t0, np.array([0, 1]), np.array([0.1, 2.0]), jac_mode=0)
    assert np.abs(res["r"]).max() < 1e-10


# --- Jacobian self-consistency: autodiff (multipass stride-4) == central differences of the double path ---------
def _numdiff(f, x, h=1e-6):
    x = x.copy()
    g = []
    for i in range(x.size):
        old = x.flat[i]
        x.flat[i] = old + h; f1 = f(x).copy()
        x.flat[i] = old - h; f2 = f(x).copy()
        x.flat[i] = old
        g.append((f1 - f2) / (2 * h))
    return np.stack(g, -1)


@pytest.mark.parametrize("which", [0, 1])
@pytest.mark.parametrize("compat", [False, True])
def test_imu_jacobian_vs_numdiff_se3(which, compat):
    knots = fx.smooth_se3_knots(12, 0.1)
    t = np.array([0.437])
    y = np.array([[0.1, -0.2, 0.3]])
    imu = kto.Sensor()
    res = kto.imu_residuals(kto.Traj(kto.SE3, 0.1, 0.0, knots, compat_zero_dB=compat), imu, which, t, y, jac_mode=2)
    i0 = res["i0_a"][0]
    assert i0 == 4

    def f(kn):
        kk = knots.copy(); kk[i0:i0 + 4] = kn
        return kto.imu_residuals(kto.Traj(kto.SE3, 0.1, 0.0, kk, compat_zero_dB=compat), imu, which, t, y, jac_mode=0)["r"][0]
    if compat and which == 1:
        # In the reference, T=double reads an uninitialised dB here (undefined behaviour); the oracle's double path
        # uses dB = 0 like the Jet path, so autodiff and numdiff agree by construction.
        pass
    Jn = _numdiff(f, knots[i0:i0 + 4])                     # (3, 28)
    Ja = np.concatenate([res["Ja"][0, k] for k in range(4)], axis=1)
    scale = np.abs(Ja).max()
    assert_allclose(Ja, Jn, atol=2e-6 * scale)


def test_camera_jacobian_vs_numdiff_se3():
    knots = fx.smooth_se3_knots(40, 0.1)
    traj = kto.Traj(kto.SE3, 0.1, 0.0, knots)
    cam = kto.Camera(fx.IMAGE_ROWS, fx.IMAGE_COLS, fx.CAMERA_READOUT, K=np.array([[900., 0, 960], [0, 900, 540], [0, 0, 1]]),
                     q_ct=fx.so3_exp_xyzw(np.array([0.1, -0.2, 0.05])), p_ct=np.array([0.05, -0.02, 0.1]))
    s = make_sfm(traj, cam, np.random.default_rng(2), nlm=3)
    sel = slice(0, 1)
    args = [s[k][sel] for k in ("obs_uv", "obs_t0", "ref_uv", "ref_t0", "lm_idx")]
    res = kto.static_rs_residuals(traj, cam, *args, s["rho"], jac_mode=2)
    ids = res["ids_a"][0]
    ids = ids[ids >= 0]

    def f(kn):
        kk = knots.copy(); kk[ids] = kn
        return kto.static_rs_residuals(kto.Traj(kto.SE3, 0.1, 0.0, kk), cam, *args, s["rho"], jac_mode=0)["r"][0]
    Jn = _numdiff(f, knots[ids])
    Ja = np.concatenate([res["Ja"][0, k] for k in range(len(ids))], axis=1)
    assert_allclose(Ja, Jn, atol=2e-6 * np.abs(Ja).max())

    def frho(r):
        return kto.static_rs_residuals(traj, cam, *args, r, jac_mode=0)["r"][0]
    Jr = _numdiff(frho, s["rho"])[:, s["lm_idx"][0]]
    assert_allclose(res["Jrho"][0], Jr, atol=1e-5 * max(1, np.abs(Jr).max()))


@pytest.mark.parametrize("name", ["se3", "split"])
def test_position_known_answer_and_jacobian_vs_numdiff(name):
    """PositionMeasurement (position_measurement.h:24-31): r = p - trajectory.Position(t); autodiff vs central differences."""
    k = fx.smooth_se3_knots(40, 0.1)
    traj = kto.Traj(kto.SE3, 0.1, 0.0, k) if name == "se3" else kto.Traj(kto.SPLIT, 0.1, 0.0, k[:, 4:7].copy(), 0.1, 0.0, k[:, 0:4].copy())
    t = np.array([2.3456])
    pos = kto.traj_evaluate(traj, t, 0xff)["position"][0]
    y = pos + np.array([0.5, -0.25, 2.0])
    res = kto.imu_residuals(traj, kto.Sensor(), 2, t, y[None, :], jac_mode=2)
    assert_allclose(res["r"][0], [0.5, -0.25, 2.0], atol=1e-12)
    i0 = res["i0_a"][0]
    if name == "se3":
        def f(kn):
            kk = k.copy(); kk[i0:i0 + 4] = kn
            return kto.imu_residuals(kto.Traj(kto.SE3, 0.1, 0.0, kk), kto.Sensor(), 2, t, y[None, :], jac_mode=0)["r"][0]
        Jn = _numdiff(f, k[i0:i0 + 4])
        Ja = np.concatenate([res["Ja"][0, j] for j in range(4)], axis=1)
        assert_allclose(Ja, Jn, atol=2e-6 * np.abs(Ja).max())
    else:
        vecs = k[:, 4:7].copy()

        def f(vn):
            vv = vecs.copy(); vv[i0:i0 + 4] = vn
            return kto.imu_residuals(kto.Traj(kto.SPLIT, 0.1, 0.0, vv, 0.1, 0.0, k[:, 0:4].copy()), kto.Sensor(), 2, t, y[None, :], jac_mode=0)["r"][0]
        Jn = _numdiff(f, vecs[i0:i0 + 4])
        Ja = np.concatenate([res["Ja"][0, j] for j in range(4)], axis=1)
        assert_allclose(Ja, Jn, atol=1e-8)                                          # linear in the R3 knots: minus the basis weights
        assert not res["Jb"].any()


@pytest.mark.parametrize("name", ["se3", "split"])
def test_orientation_known_answer_and_jacobian_vs_numdiff(name):
    """OrientationMeasurement (orientation_measurement.h:27-31): the residual is Eigen's angularDistance -- the rotation angle between the
    measured and the trajectory's orientation, whatever the sign or scale of the measured quaternion -- and the multipass autodiff of the
    restated residual agrees with central differences of its double path."""
    k = fx.smooth_se3_knots(40, 0.1)
    traj = kto.Traj(kto.SE3, 0.1, 0.0, k) if name == "se3" else kto.Traj(kto.SPLIT, 0.1, 0.0, k[:, 4:7].copy(), 0.1, 0.0, k[:, 0:4].copy())
    t = np.array([1.2345])
    q = kto.traj_evaluate(traj, t, 0xff)["orientation"][0]                      # x, y, z, w
    ang = 0.7
    ax = np.array([0.3, -0.5, 0.8]); ax /= np.linalg.norm(ax)
    dq = np.array([*(np.sin(ang / 2) * ax), np.cos(ang / 2)])
    x1, y1, z1, w1 = q; x2, y2, z2, w2 = dq
    qm = np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
                   w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])
    for scale in (1.0, -1.0, 3.0):                                               # q and -q are the same rotation; the distance is scale invariant
        res = kto.imu_residuals(traj, kto.Sensor(), 3, t, (scale * qm)[None, :], jac_mode=2)
        assert abs(res["r"][0, 0] - ang) < 1e-12
    res = kto.imu_residuals(traj, kto.Sensor(), 3, t, qm[None, :], jac_mode=2)
    if name == "se3":
        i0 = res["i0_a"][0]
        knots = k.copy()

        def f(kn):
            kk = knots.copy(); kk[i0:i0 + 4] = kn
            return kto.imu_residuals(kto.Traj(kto.SE3, 0.1, 0.0, kk), kto.Sensor(), 3, t, qm[None, :], jac_mode=0)["r"][0]
        Jn = _numdiff(f, knots[i0:i0 + 4])
        Ja = np.concatenate([res["Ja"][0, j] for j in range(4)], axis=1)
        assert_allclose(Ja, Jn, atol=2e-6 * np.abs(Ja).max())
        assert not Ja[:, [4, 5, 6, 11, 12, 13]].any()                              # the angle does not see the translations
    else:
        i0 = res["i0_b"][0]
        quats = k[:, 0:4].copy()

        def f(qn):
            qq = quats.copy(); qq[i0:i0 + 4] = qn
            return kto.imu_residuals(kto.Traj(kto.SPLIT, 0.1, 0.0, k[:, 4:7].copy(), 0.1, 0.0, qq), kto.Sensor(), 3, t, qm[None, :], jac_mode=0)["r"][0]
        Jn = _numdiff(f, quats[i0:i0 + 4], h=1e-7)
        Jb = np.concatenate([res["Jb"][0, j] for j in range(4)], axis=1)
        # logq throws for |q| off the unit sphere by more than 1e-5 (quaternion_math.h:19-23): the step stays far inside
        assert_allclose(Jb, Jn, atol=2e-5 * np.abs(Jb).max())
        assert not res["Ja"].any()


@pytest.mark.parametrize("name", ["se3", "split"])
def test_lifting_jacobian_vs_numdiff(name):
    """LiftingRsCameraMeasurement: the multipass autodiff of the restated residual (3 rows; blocks [knots | camera | vt | rho]) against central
    differences of its double path, at a displaced row time."""
    traj, cam, s = _sfm_case(name, "pinhole", "static", seed=2, nlm=3)
    sel = slice(0, 1)
    args = [s[k][sel] for k in ("obs_uv", "obs_t0", "ref_uv", "ref_t0", "lm_idx")]
    vt = np.array([s["obs_uv"][0, 1] / cam.rows + 0.05])
    res = kto.lifting_rs_residuals(traj, cam, *args, s["rho"], vt=vt, jac_mode=2)
    assert abs(res["r"][0, 2] - cam.rows * 0.05) < 1e-9
    Jv = _numdiff(lambda v: kto.lifting_rs_residuals(traj, cam, *args, s["rho"], vt=v, jac_mode=0)["r"][0], vt)[:, 0]
    assert_allclose(res["Jvt"][0], Jv, atol=2e-6 * np.abs(Jv).max())
    Jr = _numdiff(lambda r: kto.lifting_rs_residuals(traj, cam, *args, r, vt=vt, jac_mode=0)["r"][0], s["rho"])[:, s["lm_idx"][0]]
    assert_allclose(res["Jrho"][0], Jr, atol=1e-5 * max(1, np.abs(Jr).max()))
    if name == "se3":
        knots = traj.knots_a.copy()
        ids = res["ids_a"][0]; ids = ids[ids >= 0]

        def f(kn):
            kk = knots.copy(); kk[ids] = kn
            return kto.lifting_rs_residuals(kto.Traj(kto.SE3, 0.1, 0.0, kk), cam, *args, s["rho"], vt=vt, jac_mode=0)["r"][0]
        Jn = _numdiff(f, knots[ids])
        Ja = np.concatenate([res["Ja"][0, k] for k in range(len(ids))], axis=1)
        assert_allclose(Ja, Jn, atol=2e-6 * np.abs(Ja).max())
        assert not Ja[2].any()                       # the timing residual does not see the trajectory


@pytest.mark.parametrize("model", ["pinhole", "atan"])
@pytest.mark.parametrize("method", ["static", "newton"])
def test_camera_jacobian_vs_numdiff_models(model, method):
    """Atan / Newton variants of the check above (self-consistency of the multipass autodiff)."""
    traj, cam, s = _sfm_case("se3", model, method, seed=2, nlm=3)
    knots = traj.knots_a.copy()
    rng = np.random.default_rng(4)
    s["obs_uv"] = s["obs_uv"] + rng.normal(0, 0.5, size=s["obs_uv"].shape)
    for row in range(min(3, len(s["lm_idx"]))):
        sel = slice(row, row + 1)
        args = [s[k][sel] for k in ("obs_uv", "obs_t0", "ref_uv", "ref_t0", "lm_idx")]
        res = kto.static_rs_residuals(traj, cam, *args, s["rho"], jac_mode=2)
        ids = res["ids_a"][0]
        ids = ids[ids >= 0]

        def f(kn):
            kk = knots.copy(); kk[ids] = kn
            return kto.static_rs_residuals(kto.Traj(kto.SE3, 0.1, 0.0, kk), cam, *args, s["rho"], jac_mode=0)["r"][0]
        Jn = _numdiff(f, knots[ids], h=1e-7)
        Ja = np.concatenate([res["Ja"][0, k] for k in range(len(ids))], axis=1)
        assert_allclose(Ja, Jn, atol=2e-5 * np.abs(Ja).max())

        def frho(r):
            return kto.static_rs_residuals(traj, cam, *args, r, jac_mode=0)["r"][0]
        Jr = _numdiff(frho, s["rho"], h=1e-7)[:, s["lm_idx"][row]]
        assert_allclose(res["Jrho"][0], Jr, atol=1e-4 * max(1, np.abs(Jr).max()))


def test_huber_corrector_matches_definition():
    r = np.array([30.0, -40.0])        # |r| = 50 > c = 5
    J = np.arange(6, dtype=float).reshape(2, 3)
    rho, r2, J2 = kto.huber_correct(5.0, r, J)
    assert_almost_equal(rho, 2 * 5 * 50 - 25)
    s = np.sqrt(5.0 / 50.0)
    assert_allclose(r2, s * r)
    assert_allclose(J2, s * J)
    rho, r2, J2 = kto.huber_correct(5.0, np.array([1.0, 2.0]), J)
    assert_almost_equal(rho, 5.0)
    assert_allclose(r2, [1.0, 2.0]); assert_allclose(J2, J)

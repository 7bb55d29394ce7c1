"""The oracle (oracle/*.hpp, dual-number autodiff in doubles) against an INDEPENDENT 60-digit transcription of the reference
(tests/mp_reference.py: mpmath, written from the reference headers + Sophus' published formulas, Jacobians by central differences,
SE3 poses additionally by 4x4 expm / logm).  Inputs: the reference's own fixture knots (python/tests/conftest.py:32-45, :52-67, :83-105,
restated in tests/fixtures_ref.py) and random non-degenerate trajectories.  Values AND Jacobians, SE3 + SO3 + R3 (split),
gyroscope / accelerometer / static-RS camera.  CPU only."""
import mpmath as mp
import numpy as np
import pytest

import fixtures_ref as fx
import mp_reference as mr
from oracle import kto

VAL_TOL = 5e-13      # values: double rounding of the oracle against a 60-digit number
JAC_TOL = 2e-11      # Jacobians: relative to the largest entry of the block row (north_star gate is 1e-9)


def f(x):
    return np.array([[float(v) for v in row] for row in x]) if isinstance(x[0], (list, tuple)) else np.array([float(v) for v in x])


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def random_se3_knots(n, seed, step=0.35):
    """Random walk on SE3 with relative rotations up to ~step rad and unit-scale translations: far from every small-angle / pi branch."""
    rng = np.random.default_rng(seed)
    q, p, out = np.array([0.1, -0.2, 0.3, 0.9]), np.zeros(3), []
    q /= np.linalg.norm(q)
    for _ in range(n):
        out.append(np.concatenate([q, p]))
        q = fx.qmul_xyzw(q, fx.so3_exp_xyzw(rng.normal(0, step, 3)))
        q /= np.linalg.norm(q)
        p = p + rng.normal(0, 1.0, 3)
    return np.array(out)


SE3_CASES = {"fixture": (fx.SE3_KNOTS, fx.SE3_DT, fx.SE3_T0), "random": (random_se3_knots(9, 7), 0.37, -0.4)}


def times_in(knots, dt, t0, k, seed):
    rng = np.random.default_rng(seed)
    return t0 + rng.uniform(0.02, len(knots) - 3.02, k) * dt


@pytest.mark.parametrize("case", list(SE3_CASES))
def test_se3_values_three_routes(case):
    knots, dt, t0 = SE3_CASES[case]
    traj = kto.Traj(kto.SE3, dt, t0, knots)
    mt = mr.Trajectory("se3", dt, t0, knots=knots)
    ts = times_in(knots, dt, t0, 6, 1)
    o = kto.traj_evaluate(traj, ts, kto.EvalPosition | kto.EvalVelocity | kto.EvalAcceleration | kto.EvalOrientation | kto.EvalAngularVelocity)
    for k, t in enumerate(ts):
        e = mt.evaluate(mp.mpf(float(t)), acc=True)
        assert rel(o["position"][k], f(e["position"])) < VAL_TOL
        assert rel(o["velocity"][k], f(e["velocity"])) < VAL_TOL
        assert rel(o["acceleration"][k], f(e["acceleration"])) < VAL_TOL
        assert rel(o["angular_velocity"][k], f(e["angular_velocity"])) < VAL_TOL
        assert rel(o["orientation"][k], f(e["orientation"])) < VAL_TOL
        # second independent route: 4x4 matrix exponential / logarithm, no closed forms at all
        P = mr.se3_pose_expm(mt.knots, mt.t0, mt.dt, mp.mpf(float(t)))
        Pm = np.array([[float(P[i, j]) for j in range(4)] for i in range(4)])
        assert rel(o["position"][k], Pm[:3, 3]) < VAL_TOL
        assert rel(fx.rot_xyzw(o["orientation"][k]), Pm[:3, :3]) < VAL_TOL
        # velocity / acceleration are the time derivatives of that pose (what the reference's xfail-ed test wanted to check)
        h = mp.mpf("1e-20")
        Pp = mr.se3_pose_expm(mt.knots, mt.t0, mt.dt, mp.mpf(float(t)) + h)
        Pn = mr.se3_pose_expm(mt.knots, mt.t0, mt.dt, mp.mpf(float(t)) - h)
        vel = [float((Pp[i, 3] - Pn[i, 3]) / (2 * h)) for i in range(3)]
        acc = [float((Pp[i, 3] - 2 * P[i, 3] + Pn[i, 3]) / (h * h)) for i in range(3)]
        assert rel(o["velocity"][k], vel) < VAL_TOL
        assert rel(o["acceleration"][k], acc) < 1e-10


@pytest.mark.parametrize("which", [0, 1])
@pytest.mark.parametrize("case", list(SE3_CASES))
def test_se3_imu_residual_and_jacobian(case, which):
    knots, dt, t0 = SE3_CASES[case]
    rng = np.random.default_rng(11 + which)
    ts = times_in(knots, dt, t0, 3, 2 + which)
    y = rng.uniform(-1, 1, (len(ts), 3))
    w = rng.uniform(0.5, 2.0, len(ts))
    for compat in ([False, True] if which == 1 else [False]):
        o = kto.imu_residuals(kto.Traj(kto.SE3, dt, t0, knots, compat_zero_dB=compat), kto.Sensor(), which, ts, y, w, jac_mode=2)
        mt = mr.Trajectory("se3", dt, t0, knots=knots, compat_zero_dB=compat)
        for k, t in enumerate(ts):
            fun = lambda: mr.imu_residual(mt, which, float(t), y[k], float(w[k]))
            assert rel(o["r"][k], f(fun())) < VAL_TOL * 10
            i0 = int(o["i0_a"][k])
            assert i0 == mt.evaluate(mp.mpf(float(t)))["i0"]
            J = np.array(mr.jacobian(fun, [(mt.knots, i0 + b, c) for b in range(4) for c in range(7)])).reshape(3, 4, 7).transpose(1, 0, 2)
            assert (o["ids_a"][k, :4] == i0 + np.arange(4)).all()
            assert rel(o["Ja"][k, :4], J) < JAC_TOL, (case, which, compat, k)
            # a knot outside the window does not enter
            other = [(mt.knots, b, 0) for b in range(len(knots)) if b < i0 or b > i0 + 3][:1]
            if other:
                assert np.abs(np.array(mr.jacobian(fun, other))).max() == 0.0


def camera_case(knots, dt, t0, seed, q_ct=(0, 0, 0, 1), p_ct=(0, 0, 0), time_offset=0.0):
    """A few reference / observation pairs whose projections land near the image (rho and uv from a forward simulation in doubles)."""
    rng = np.random.default_rng(seed)
    K = np.array([[900., 0, 960], [0, 900., 540], [0, 0, 1]])
    cam = dict(K=K, rows=1080, readout=0.026, q_ct=q_ct, p_ct=p_ct, time_offset=time_offset)
    lo, hi = t0 + 0.3 * dt, t0 + (len(knots) - 3.3) * dt - 0.03
    n = 3
    ref_t0 = rng.uniform(lo, hi, n)
    obs_t0 = np.clip(ref_t0 + rng.uniform(-0.8, 0.8, n) * dt, lo, hi)
    ref_uv = rng.uniform([100, 100], [1800, 1000], (n, 2))
    obs_uv = rng.uniform([100, 100], [1800, 1000], (n, 2))
    rho = rng.uniform(0.05, 0.6, n)
    return cam, obs_uv, obs_t0, ref_uv, ref_t0, rho


@pytest.mark.parametrize("relpose", [False, True])
@pytest.mark.parametrize("case", list(SE3_CASES))
def test_se3_static_rs_residual_and_jacobian(case, relpose):
    knots, dt, t0 = SE3_CASES[case]
    kw = dict(q_ct=tuple(np.array([0.1, -0.05, 0.2, 0.97]) / np.linalg.norm([0.1, -0.05, 0.2, 0.97])), p_ct=(0.05, -0.02, 0.1), time_offset=0.004) if relpose else {}
    cam, obs_uv, obs_t0, ref_uv, ref_t0, rho = camera_case(knots, dt, t0, 5, **kw)
    n = len(rho)
    ocam = kto.Camera(cam["rows"], 1920, cam["readout"], K=cam["K"], **({} if not relpose else dict(q_ct=kw["q_ct"], p_ct=kw["p_ct"], time_offset=kw["time_offset"])))
    o = kto.static_rs_residuals(kto.Traj(kto.SE3, dt, t0, knots), ocam, obs_uv, obs_t0, ref_uv, ref_t0, np.arange(n, dtype=np.int32), rho, jac_mode=2, cap=24)
    mt = mr.Trajectory("se3", dt, t0, knots=knots)
    for k in range(n):
        box = [[mp.mpf(float(rho[k]))]]
        fun = lambda: mr.static_rs_residual(mt, cam, obs_uv[k], float(obs_t0[k]), ref_uv[k], float(ref_t0[k]), box[0][0])[0]
        r, ir, io = mr.static_rs_residual(mt, cam, obs_uv[k], float(obs_t0[k]), ref_uv[k], float(ref_t0[k]), box[0][0])
        assert rel(o["r"][k], f(r)) < 1e-12                                   # |r| ~ 1e3 px: 1e-12 relative = 1e-9 px
        assert (int(o["i0_ref_a"][k]), int(o["i0_obs_a"][k])) == (ir, io)
        ids = [int(v) for v in o["ids_a"][k] if v >= 0]
        J = np.array(mr.jacobian(fun, [(mt.knots, b, c) for b in ids for c in range(7)])).reshape(2, len(ids), 7).transpose(1, 0, 2)
        assert rel(o["Ja"][k, :len(ids)], J) < JAC_TOL, (case, relpose, k)
        assert rel(o["Jrho"][k], np.array(mr.jacobian(fun, [(box, 0, 0)])).reshape(2)) < JAC_TOL
        # every knot the residual depends on is in the structural list
        rest = [(mt.knots, b, c) for b in range(len(knots)) if b not in ids for c in (0, 5)]
        if rest:
            assert np.abs(np.array(mr.jacobian(fun, rest))).max() == 0.0


def split_fixture():
    """conftest.py:107-113: the split fixture is the R3 and the SO3 fixture side by side."""
    return fx.R3_KNOTS, fx.R3_DT, fx.R3_T0, fx.SO3_KNOTS, fx.SO3_DT, fx.SO3_T0


def test_split_values_and_so3_known_answer():
    r3, dta, t0a, so3, dtb, t0b = split_fixture()
    for t in (2.0, 3.3, 4.9):
        e = mr.so3_spline([mr.mpv(k) for k in so3], mp.mpf(t0b), mp.mpf(dtb), mp.mpf(t))
        # conftest.py:52-81 + test_general.py:52-62: constant 10 deg/s about (1,0,1)/sqrt2
        assert rel(f(e["angular_velocity"]), fx.SO3_RATE * fx.SO3_AXIS) < 1e-5      # the knots sample the motion; the cumulative spline of a one-parameter subgroup reproduces it
        o = kto.traj_evaluate(kto.Traj(kto.SO3, knots_b=so3, dt_b=dtb, t0_b=t0b), [t], kto.EvalOrientation | kto.EvalAngularVelocity)
        assert rel(o["orientation"][0], f(e["orientation"])) < VAL_TOL and rel(o["angular_velocity"][0], f(e["angular_velocity"])) < VAL_TOL
        a = mr.r3_spline([mr.mpv(k) for k in r3], mp.mpf(t0a), mp.mpf(dta), mp.mpf(t + 1.5))
        o = kto.traj_evaluate(kto.Traj(kto.R3, dta, t0a, r3), [t + 1.5], kto.EvalPosition | kto.EvalVelocity | kto.EvalAcceleration)
        for key in ("position", "velocity", "acceleration"):
            assert rel(o[key][0], f(a[key])) < VAL_TOL


def random_split(n, seed):
    se3 = random_se3_knots(n, seed, step=0.3)
    return se3[:, 4:].copy(), se3[:, :4].copy()


@pytest.mark.parametrize("which", [0, 1])
def test_split_imu_residual_and_jacobian(which):
    r3, so3 = random_split(9, 21)
    dt, t0 = 0.41, 0.3
    traj = kto.Traj(kto.SPLIT, dt, t0, r3, dt, t0, so3)
    mt = mr.Trajectory("split", dt, t0, r3=r3, so3=so3)
    rng = np.random.default_rng(3)
    ts = t0 + rng.uniform(0.1, 5.9, 3) * dt
    y, w = rng.uniform(-1, 1, (3, 3)), rng.uniform(0.5, 2, 3)
    o = kto.imu_residuals(traj, kto.Sensor(), which, ts, y, w, jac_mode=2)
    for k, t in enumerate(ts):
        fun = lambda: mr.imu_residual(mt, which, float(t), y[k], float(w[k]))
        assert rel(o["r"][k], f(fun())) < VAL_TOL * 10
        ia, ib = int(o["i0_a"][k]), int(o["i0_b"][k])
        Jb = np.array(mr.jacobian(fun, [(mt.so3, ib + b, c) for b in range(4) for c in range(4)])).reshape(3, 4, 4).transpose(1, 0, 2)
        assert rel(o["Jb"][k, :4], Jb) < JAC_TOL
        Ja = np.array(mr.jacobian(fun, [(mt.r3, ia + b, c) for b in range(4) for c in range(3)])).reshape(3, 4, 3).transpose(1, 0, 2)
        if which == 0:
            assert np.abs(Ja).max() == 0.0 and np.abs(o["Ja"][k]).max() == 0.0      # R3 blocks structurally present, identically zero
        else:
            assert rel(o["Ja"][k, :4], Ja) < JAC_TOL


def test_split_static_rs_residual_and_jacobian():
    r3, so3 = random_split(10, 33)
    dt, t0 = 0.23, -0.1
    cam, obs_uv, obs_t0, ref_uv, ref_t0, rho = camera_case(r3, dt, t0, 8)
    n = len(rho)
    ocam = kto.Camera(cam["rows"], 1920, cam["readout"], K=cam["K"])
    o = kto.static_rs_residuals(kto.Traj(kto.SPLIT, dt, t0, r3, dt, t0, so3), ocam, obs_uv, obs_t0, ref_uv, ref_t0, np.arange(n, dtype=np.int32), rho, jac_mode=2, cap=24)
    mt = mr.Trajectory("split", dt, t0, r3=r3, so3=so3)
    for k in range(n):
        box = [[mp.mpf(float(rho[k]))]]
        fun = lambda: mr.static_rs_residual(mt, cam, obs_uv[k], float(obs_t0[k]), ref_uv[k], float(ref_t0[k]), box[0][0])[0]
        assert rel(o["r"][k], f(fun())) < 1e-12
        ida = [int(v) for v in o["ids_a"][k] if v >= 0]
        idb = [int(v) for v in o["ids_b"][k] if v >= 0]
        Ja = np.array(mr.jacobian(fun, [(mt.r3, b, c) for b in ida for c in range(3)])).reshape(2, len(ida), 3).transpose(1, 0, 2)
        Jb = np.array(mr.jacobian(fun, [(mt.so3, b, c) for b in idb for c in range(4)])).reshape(2, len(idb), 4).transpose(1, 0, 2)
        assert rel(o["Ja"][k, :len(ida)], Ja) < JAC_TOL and rel(o["Jb"][k, :len(idb)], Jb) < JAC_TOL
        assert rel(o["Jrho"][k], np.array(mr.jacobian(fun, [(box, 0, 0)])).reshape(2)) < JAC_TOL


# ---- widening rows: AtanCamera, NewtonRs / LiftingRs camera measurements, Position / Orientation ------------
ATAN = dict(model="atan", wc=(0.02, -0.01), gamma=0.9)
REL = dict(q_ct=tuple(np.array([0.1, -0.05, 0.2, 0.97]) / np.linalg.norm([0.1, -0.05, 0.2, 0.97])), p_ct=(0.05, -0.02, 0.1), time_offset=0.004)


def span_camera_case(knots, dt, t0, seed, atan, noise_rows, n=3, split=False):
    """Reference / observation pairs of ONE landmark each whose observation is (nearly) where the landmark projects: rho and the observed pixel from a
    forward simulation with the 60-digit reference, the observed row then displaced by `noise_rows` so that the Newton iteration has work to do."""
    rng = np.random.default_rng(seed)
    K = np.array([[900., 0, 960], [0, 900., 540], [0, 0, 1]])
    cam = dict(K=K, rows=1080, readout=0.026, **REL, **(ATAN if atan else {}))
    mt = mr.Trajectory("split", dt, t0, r3=knots[:, 4:], so3=knots[:, :4]) if split else mr.Trajectory("se3", dt, t0, knots=knots)
    lo, hi = t0 + 0.3 * dt, t0 + (len(knots) - 3.3) * dt - 0.03
    rows = []
    while len(rows) < n:
        ref_t0 = rng.uniform(lo, hi)
        obs_t0 = float(np.clip(ref_t0 + rng.uniform(-0.8, 0.8) * dt, lo, hi))
        ref_uv = rng.uniform([300, 200], [1600, 900])
        rho = rng.uniform(0.05, 0.4)
        # where does it project at mid-frame?  use the lifting projection (a plain projection at a given row time) to find a consistent pixel
        r, _, _ = mr.lifting_rs_residual(mt, cam, (0.0, 540.0), obs_t0, ref_uv, ref_t0, mp.mpf(rho), mp.mpf(0.5))
        u, v = -float(r[0]), 540.0 - float(r[1])
        if not (50 < u < 1870 and 100 < v < 980):
            continue
        rows.append((np.array([u + rng.normal(0, 0.5), v + rng.normal(0, noise_rows)]), obs_t0, ref_uv, ref_t0, rho))
    obs_uv, obs_t0, ref_uv, ref_t0, rho = (np.array(x) for x in zip(*rows))
    return cam, obs_uv, obs_t0, ref_uv, ref_t0, rho


def oracle_camera(cam, method):
    kw = dict(wc=cam["wc"], gamma=cam["gamma"]) if cam.get("model") == "atan" else {}
    return kto.Camera(cam["rows"], 1920, cam["readout"], K=cam["K"], method=method, q_ct=cam["q_ct"], p_ct=cam["p_ct"], time_offset=cam["time_offset"], **kw)


def test_atan_camera_projection_and_static_rs_rows():
    knots, dt, t0 = SE3_CASES["random"]
    cam, obs_uv, obs_t0, ref_uv, ref_t0, rho = span_camera_case(knots, dt, t0, 3, True, 2.0)
    # projection / unprojection round trip of the transcription itself, then the oracle's camera against it
    y = mr.mpv((700.0, 300.0))
    Y = mr.camera_unproject(cam, y)
    y2, _ = mr.camera_project(cam, [2 * v for v in Y])
    assert max(abs(float(a - b)) for a, b in zip(y, y2)) < 1e-25
    ocam = oracle_camera(cam, "static")
    yo = kto.camera_project(ocam, np.array([float(v) for v in Y]) * 2)
    assert rel(np.asarray(yo).reshape(-1)[:2], f(y2)) < 1e-13


@pytest.mark.parametrize("atan", [False, True])
@pytest.mark.parametrize("case", list(SE3_CASES))
def test_se3_newton_rs_residual_and_jacobian(case, atan):
    """NewtonRsCameraMeasurement: value of the iteration AND its derivative through the iteration (central differences of the whole loop at 60 digits:
    the perturbation is far too small to change the number of evaluations, exactly the function a Jet differentiates)."""
    knots, dt, t0 = SE3_CASES[case]
    cam, obs_uv, obs_t0, ref_uv, ref_t0, rho = span_camera_case(knots, dt, t0, 11 + atan, atan, 8.0)
    n = len(rho)
    o = kto.static_rs_residuals(kto.Traj(kto.SE3, dt, t0, knots), oracle_camera(cam, "newton"), obs_uv, obs_t0, ref_uv, ref_t0, np.arange(n, dtype=np.int32), rho,
                                jac_mode=2, cap=32)
    mt = mr.Trajectory("se3", dt, t0, knots=knots)
    evals = []
    for k in range(n):
        box = [[mp.mpf(float(rho[k]))]]
        fun = lambda: mr.newton_rs_residual(mt, cam, obs_uv[k], float(obs_t0[k]), ref_uv[k], float(ref_t0[k]), box[0][0])[0]
        r, ir, ne = mr.newton_rs_residual(mt, cam, obs_uv[k], float(obs_t0[k]), ref_uv[k], float(ref_t0[k]), box[0][0])
        evals.append(ne)
        assert np.abs(o["r"][k] - f(r)).max() < 1e-9                           # pixels
        assert int(o["i0_ref_a"][k]) == ir
        ids = [int(v) for v in o["ids_a"][k] if v >= 0]
        J = np.array(mr.jacobian(fun, [(mt.knots, b, c) for b in ids for c in range(7)])).reshape(2, len(ids), 7).transpose(1, 0, 2)
        assert rel(o["Ja"][k, :len(ids)], J) < JAC_TOL, (case, atan, k)
        assert rel(o["Jrho"][k], np.array(mr.jacobian(fun, [(box, 0, 0)])).reshape(2)) < JAC_TOL
        rest = [(mt.knots, b, c) for b in range(len(knots)) if b not in ids for c in (0, 5)]
        if rest:
            assert np.abs(np.array(mr.jacobian(fun, rest))).max() == 0.0
    assert max(evals) >= 2                                                      # the derivative through a step of the iteration was exercised


@pytest.mark.parametrize("atan", [False, True])
def test_se3_lifting_rs_residual_and_jacobian(atan):
    knots, dt, t0 = SE3_CASES["random"]
    cam, obs_uv, obs_t0, ref_uv, ref_t0, rho = span_camera_case(knots, dt, t0, 21 + atan, atan, 3.0)
    n = len(rho)
    vt = np.clip(obs_uv[:, 1] / cam["rows"] + np.array([0.07, -0.05, 0.11]), 0.0, 1.0)
    o = kto.lifting_rs_residuals(kto.Traj(kto.SE3, dt, t0, knots), oracle_camera(cam, "static"), obs_uv, obs_t0, ref_uv, ref_t0, np.arange(n, dtype=np.int32), rho,
                                 vt=vt, jac_mode=2, cap=32)
    mt = mr.Trajectory("se3", dt, t0, knots=knots)
    for k in range(n):
        box = [[mp.mpf(float(rho[k])), mp.mpf(float(vt[k]))]]
        fun = lambda: mr.lifting_rs_residual(mt, cam, obs_uv[k], float(obs_t0[k]), ref_uv[k], float(ref_t0[k]), box[0][0], box[0][1])[0]
        r, ir, io = fun(), *mr.lifting_rs_residual(mt, cam, obs_uv[k], float(obs_t0[k]), ref_uv[k], float(ref_t0[k]), box[0][0], box[0][1])[1:]
        assert np.abs(o["r"][k] - f(r)).max() < 1e-9 and int(o["i0_ref_a"][k]) == ir
        ids = [int(v) for v in o["ids_a"][k] if v >= 0]
        assert io in ids and io + 3 in ids
        J = np.array(mr.jacobian(fun, [(mt.knots, b, c) for b in ids for c in range(7)])).reshape(3, len(ids), 7).transpose(1, 0, 2)
        assert rel(o["Ja"][k, :len(ids)], J) < JAC_TOL, (atan, k)
        assert rel(o["Jrho"][k], np.array(mr.jacobian(fun, [(box, 0, 0)])).reshape(3)) < JAC_TOL
        assert rel(o["Jvt"][k], np.array(mr.jacobian(fun, [(box, 0, 1)])).reshape(3)) < JAC_TOL


@pytest.mark.parametrize("kind", ["se3", "split"])
def test_position_and_orientation_residual_and_jacobian(kind):
    """PositionMeasurement / OrientationMeasurement (the reference has no test of its own for the latter): values and ambient Jacobians."""
    rng = np.random.default_rng(8)
    if kind == "se3":
        knots, dt, t0 = SE3_CASES["random"]
        ot, mt = kto.Traj(kto.SE3, dt, t0, knots), mr.Trajectory("se3", dt, t0, knots=knots)
        t = times_in(knots, dt, t0, 3, 5)
    else:
        r3, so3 = random_split(9, 23)
        dt, t0 = 0.41, 0.3
        ot, mt = kto.Traj(kto.SPLIT, dt, t0, r3, dt, t0, so3), mr.Trajectory("split", dt, t0, r3=r3, so3=so3)
        t = times_in(r3, dt, t0, 3, 6)
    imu = kto.Sensor()
    p_meas = rng.normal(0, 1, (3, 3))
    q_meas = rng.normal(0, 1, (3, 4))
    q_meas /= np.linalg.norm(q_meas, axis=1, keepdims=True)
    for which, y, fun_of in ((2, p_meas, mr.position_residual), (3, q_meas, mr.orientation_residual)):
        o = kto.imu_residuals(ot, imu, which, t, y, jac_mode=2)
        for k in range(3):
            fun = lambda: fun_of(mt, float(t[k]), y[k])
            assert rel(o["r"][k], f(fun())) < 1e-12
            if kind == "se3":
                ids = [int(v) for v in o["ids_a"][k] if v >= 0]
                J = np.array(mr.jacobian(fun, [(mt.knots, b, c) for b in ids for c in range(7)])).reshape(-1, len(ids), 7).transpose(1, 0, 2)
                assert rel(o["Ja"][k, :len(ids)], J) < JAC_TOL, (which, k)
            else:
                ida, idb = [int(v) for v in o["ids_a"][k] if v >= 0], [int(v) for v in o["ids_b"][k] if v >= 0]
                Ja = np.array(mr.jacobian(fun, [(mt.r3, b, c) for b in ida for c in range(3)])).reshape(-1, len(ida), 3).transpose(1, 0, 2)
                Jb = np.array(mr.jacobian(fun, [(mt.so3, b, c) for b in idb for c in range(4)])).reshape(-1, len(idb), 4).transpose(1, 0, 2)
                if which == 2:
                    assert rel(o["Ja"][k, :len(ida)], Ja) < JAC_TOL and np.abs(Jb).max() == 0.0
                else:
                    assert rel(o["Jb"][k, :len(idb)], Jb) < JAC_TOL and np.abs(Ja).max() == 0.0


@pytest.mark.parametrize("kind", ["newton", "lifting"])
def test_split_newton_and_lifting_residual_and_jacobian(kind):
    """The span camera rows on a SplitTrajectory (the reference instantiates every measurement with every trajectory, measurement_defs.h:40-85):
    R3 and SO3 blocks of the oracle against central differences of the 60-digit residual."""
    se3 = random_se3_knots(9, 17, step=0.25)
    r3, so3 = se3[:, 4:].copy(), se3[:, :4].copy()
    dt, t0 = 0.41, 0.3
    cam, obs_uv, obs_t0, ref_uv, ref_t0, rho = span_camera_case(se3, dt, t0, 41, kind == "lifting", 6.0, n=2, split=True)
    n = len(rho)
    vt = np.clip(obs_uv[:, 1] / cam["rows"] + np.array([0.05, -0.07]), 0.0, 1.0)
    traj = kto.Traj(kto.SPLIT, dt, t0, r3, dt, t0, so3)
    lm = np.arange(n, dtype=np.int32)
    if kind == "newton":
        o = kto.static_rs_residuals(traj, oracle_camera(cam, "newton"), obs_uv, obs_t0, ref_uv, ref_t0, lm, rho, jac_mode=2, cap=32)
    else:
        o = kto.lifting_rs_residuals(traj, oracle_camera(cam, "static"), obs_uv, obs_t0, ref_uv, ref_t0, lm, rho, vt=vt, jac_mode=2, cap=32)
    mt = mr.Trajectory("split", dt, t0, r3=r3, so3=so3)
    nres = 2 if kind == "newton" else 3
    for k in range(n):
        box = [[mp.mpf(float(rho[k])), mp.mpf(float(vt[k]))]]
        if kind == "newton":
            fun = lambda: mr.newton_rs_residual(mt, cam, obs_uv[k], float(obs_t0[k]), ref_uv[k], float(ref_t0[k]), box[0][0])[0]
        else:
            fun = lambda: mr.lifting_rs_residual(mt, cam, obs_uv[k], float(obs_t0[k]), ref_uv[k], float(ref_t0[k]), box[0][0], box[0][1])[0]
        assert np.abs(o["r"][k] - f(fun())).max() < 1e-9
        ida, idb = [int(v) for v in o["ids_a"][k] if v >= 0], [int(v) for v in o["ids_b"][k] if v >= 0]
        Ja = np.array(mr.jacobian(fun, [(mt.r3, b, c) for b in ida for c in range(3)])).reshape(nres, len(ida), 3).transpose(1, 0, 2)
        Jb = np.array(mr.jacobian(fun, [(mt.so3, b, c) for b in idb for c in range(4)])).reshape(nres, len(idb), 4).transpose(1, 0, 2)
        assert rel(o["Ja"][k, :len(ida)], Ja) < JAC_TOL and rel(o["Jb"][k, :len(idb)], Jb) < JAC_TOL, (kind, k)
        assert rel(o["Jrho"][k], np.array(mr.jacobian(fun, [(box, 0, 0)])).reshape(nres)) < JAC_TOL
        if kind == "lifting":
            assert rel(o["Jvt"][k], np.array(mr.jacobian(fun, [(box, 0, 1)])).reshape(3)) < JAC_TOL


# ---- sensor parameter blocks: relative pose and time offset of the camera, time offset and biases of the IMU ------------
def _mp_cam(cam):
    """camera dict whose relative pose and time offset are lists of mpf that jacobian() can nudge: [[q_ct (4)], [p_ct (3)], [time_offset]]"""
    box = [mr.mpv(cam["q_ct"]), mr.mpv(cam["p_ct"]), [mp.mpf(float(cam["time_offset"]))]]
    live = dict(cam)
    return box, live


@pytest.mark.parametrize("method", ["static", "newton", "lifting"])
def test_camera_sensor_blocks_against_central_differences(method):
    """d r / d (q_ct, p_ct, time offset) of the three camera measurements (sensors.h:135-165 puts them into the residual's parameter list when unlocked):
    the oracle's Js against central differences of the 60-digit residual, ambient in q_ct."""
    knots, dt, t0 = SE3_CASES["random"]
    cam, obs_uv, obs_t0, ref_uv, ref_t0, rho = span_camera_case(knots, dt, t0, 51, method == "newton", 4.0, n=2)
    n = len(rho)
    vt = np.clip(obs_uv[:, 1] / cam["rows"] + np.array([0.05, -0.06]), 0.0, 1.0)
    kw = dict(wc=cam["wc"], gamma=cam["gamma"]) if cam.get("model") == "atan" else {}
    ocam = kto.Camera(cam["rows"], 1920, cam["readout"], K=cam["K"], method="newton" if method == "newton" else "static", q_ct=cam["q_ct"], p_ct=cam["p_ct"],
                      time_offset=cam["time_offset"], max_time_offset=0.02, q_locked=False, p_locked=False, d_locked=False, **kw)
    traj, lm = kto.Traj(kto.SE3, dt, t0, knots), np.arange(n, dtype=np.int32)
    if method == "lifting":
        o = kto.lifting_rs_residuals(traj, ocam, obs_uv, obs_t0, ref_uv, ref_t0, lm, rho, vt=vt, jac_mode=2, cap=40)
    else:
        o = kto.static_rs_residuals(traj, ocam, obs_uv, obs_t0, ref_uv, ref_t0, lm, rho, jac_mode=2, cap=40)
    nres = 3 if method == "lifting" else 2
    mt = mr.Trajectory("se3", dt, t0, knots=knots)
    box, live = _mp_cam(cam)
    for k in range(n):
        def fun():
            live["q_ct"], live["p_ct"], live["time_offset"] = box[0], box[1], box[2][0]
            a = (mt, live, obs_uv[k], float(obs_t0[k]), ref_uv[k], float(ref_t0[k]), mp.mpf(float(rho[k])))
            if method == "static":
                return mr.static_rs_residual(*a)[0]
            if method == "newton":
                return mr.newton_rs_residual(*a)[0]
            return mr.lifting_rs_residual(*a, mp.mpf(float(vt[k])))[0]
        assert np.abs(o["r"][k] - f(fun())).max() < 1e-9
        J = np.array(mr.jacobian(fun, [(box, 0, c) for c in range(4)] + [(box, 1, c) for c in range(3)] + [(box, 2, 0)]))      # (nres, 8)
        Js = o["Js"][k]
        mine = np.concatenate([Js[:4 * nres].reshape(nres, 4), Js[4 * nres:7 * nres].reshape(nres, 3), Js[7 * nres:8 * nres].reshape(nres, 1)], axis=1)
        for blk in (slice(0, 4), slice(4, 7), slice(7, 8)):
            assert rel(mine[:, blk], J[:, blk]) < JAC_TOL, (method, k, blk)


@pytest.mark.parametrize("which", [0, 1])
def test_imu_sensor_blocks_against_central_differences(which):
    """Time offset of the IMU (imu.h:47-59: the trajectory is evaluated at t + time_offset) and the biases of a ConstantBiasImu
    (constant_bias_imu.h:100-119): the oracle's Js [q_ct 3x4 | p_ct 3x3 | time offset 3 | accelerometer bias 3x3 | gyroscope bias 3x3]."""
    knots, dt, t0 = SE3_CASES["random"]
    t = times_in(knots, dt, t0, 3, 9)
    rng = np.random.default_rng(2)
    y = rng.normal(0, 1, (3, 3))
    ab, gb, d = np.array([0.05, -0.02, 0.03]), np.array([-0.01, 0.02, 0.015]), 0.003
    imu = kto.Sensor(time_offset=d, max_time_offset=0.02, d_locked=False, abias=ab, gbias=gb, abias_locked=False, gbias_locked=False)
    o = kto.imu_residuals(kto.Traj(kto.SE3, dt, t0, knots), imu, which, t, y, jac_mode=2)
    mt = mr.Trajectory("se3", dt, t0, knots=knots)
    box = [[mp.mpf(d)], mr.mpv(ab), mr.mpv(gb)]
    for k in range(3):
        def fun():
            m = mr.gyroscope(mt, mp.mpf(float(t[k])), box[0][0]) if which == 0 else mr.accelerometer(mt, mp.mpf(float(t[k])), box[0][0])
            bias = box[2] if which == 0 else box[1]                      # constant_bias_imu.h: measurement = standard model + bias
            return [mp.mpf(float(y[k][i])) - (m[i] + bias[i]) for i in range(3)]
        assert rel(o["r"][k], f(fun())) < 1e-12
        J = np.array(mr.jacobian(fun, [(box, 0, 0)] + [(box, 1, c) for c in range(3)] + [(box, 2, c) for c in range(3)]))      # (3, 7)
        Js = o["Js"][k]
        assert rel(Js[21:24].reshape(3, 1), J[:, 0:1]) < JAC_TOL, (which, k)
        assert np.abs(Js[24:33].reshape(3, 3) - J[:, 1:4]).max() < 1e-12 and np.abs(Js[33:42].reshape(3, 3) - J[:, 4:7]).max() < 1e-12


def test_local_parameterization_jacobians_against_central_differences():
    """What Ceres multiplies the ambient blocks with, and what KTK_EVAL_LOCAL / the device Gauss-Newton step apply on the GPU: d Plus(x, delta) / d delta at 0.
    SE3 knots: Plus(T, delta) = T * SE3::exp(delta), delta = [upsilon; omega] (uniform_se3_spline_trajectory.h:20-31; the reference takes the Jacobian from
    Sophus' Dx_this_mul_exp_x_at_0).  SO3 knots: ceres::EigenQuaternionParameterization, Plus(q, d) = (sin|d| d/|d|, cos|d|) * q
    (uniform_so3_spline_trajectory.h:21).  The closed forms of kontiki_b200.estimator (the reference of the GPU local-row tests) against central differences
    of the 60-digit group operations."""
    from kontiki_b200.estimator import _quat_plus_jacobian, _se3_plus_jacobian
    knots = random_se3_knots(4, 31)
    P = _se3_plus_jacobian(knots)
    h = mp.mpf("1e-25")
    for k, kn in enumerate(knots):
        q, t = mr.mpv(kn[:4]), mr.mpv(kn[4:])
        num = np.zeros((7, 6))
        for c in range(6):
            out = []
            for s in (1, -1):
                d = [mp.mpf(0)] * 6
                d[c] = s * h
                qe, te = mr.se3_exp(d)
                qq, tt = mr.se3_mul(q, t, qe, te)
                out.append(list(qq) + list(tt))
            num[:, c] = [float((a - b) / (2 * h)) for a, b in zip(*out)]
        assert np.abs(P[k] - num).max() < 1e-14, k
    Q = _quat_plus_jacobian(knots[:, :4])
    for k, kn in enumerate(knots):
        q = mr.mpv(kn[:4])
        num = np.zeros((4, 3))
        for c in range(3):
            out = []
            for s in (1, -1):
                d = [mp.mpf(0)] * 3
                d[c] = s * h
                n = mp.sqrt(sum(v * v for v in d))
                qd = [mp.sin(n) * v / n for v in d] + [mp.cos(n)]
                out.append(mr.q_mul(qd, q))
            num[:, c] = [float((a - b) / (2 * h)) for a, b in zip(*out)]
        assert np.abs(Q[k] - num).max() < 1e-14, k


def test_split_sensor_blocks_against_central_differences():
    """The same sensor blocks on a SplitTrajectory (split_trajectory.h:117-123; the verdict's f-2 remainder): camera relative pose / time offset of static rows,
    time offset and bias of the IMU."""
    se3 = random_se3_knots(9, 19, step=0.25)
    r3, so3 = se3[:, 4:].copy(), se3[:, :4].copy()
    dt, t0 = 0.41, 0.3
    traj, mt = kto.Traj(kto.SPLIT, dt, t0, r3, dt, t0, so3), mr.Trajectory("split", dt, t0, r3=r3, so3=so3)
    cam, obs_uv, obs_t0, ref_uv, ref_t0, rho = span_camera_case(se3, dt, t0, 61, False, 2.0, n=2, split=True)
    ocam = kto.Camera(cam["rows"], 1920, cam["readout"], K=cam["K"], q_ct=cam["q_ct"], p_ct=cam["p_ct"], time_offset=cam["time_offset"], max_time_offset=0.02,
                      q_locked=False, p_locked=False, d_locked=False)
    o = kto.static_rs_residuals(traj, ocam, obs_uv, obs_t0, ref_uv, ref_t0, np.arange(2, dtype=np.int32), rho, jac_mode=2, cap=40)
    box, live = _mp_cam(cam)
    for k in range(2):
        def fun():
            live["q_ct"], live["p_ct"], live["time_offset"] = box[0], box[1], box[2][0]
            return mr.static_rs_residual(mt, live, obs_uv[k], float(obs_t0[k]), ref_uv[k], float(ref_t0[k]), mp.mpf(float(rho[k])))[0]
        assert np.abs(o["r"][k] - f(fun())).max() < 1e-9
        J = np.array(mr.jacobian(fun, [(box, 0, c) for c in range(4)] + [(box, 1, c) for c in range(3)] + [(box, 2, 0)]))
        Js = o["Js"][k]
        mine = np.concatenate([Js[:8].reshape(2, 4), Js[8:14].reshape(2, 3), Js[14:16].reshape(2, 1)], axis=1)
        for blk in (slice(0, 4), slice(4, 7), slice(7, 8)):
            assert rel(mine[:, blk], J[:, blk]) < JAC_TOL, (k, blk)
    t = times_in(r3, dt, t0, 2, 4)
    y = np.random.default_rng(5).normal(0, 1, (2, 3))
    ab, gb, d = np.array([0.05, -0.02, 0.03]), np.array([-0.01, 0.02, 0.015]), 0.003
    imu = kto.Sensor(time_offset=d, max_time_offset=0.02, d_locked=False, abias=ab, gbias=gb, abias_locked=False, gbias_locked=False)
    for which in (0, 1):
        oi = kto.imu_residuals(traj, imu, which, t, y, jac_mode=2)
        ibox = [[mp.mpf(d)]]
        for k in range(2):
            def fun():
                m = mr.gyroscope(mt, mp.mpf(float(t[k])), ibox[0][0]) if which == 0 else mr.accelerometer(mt, mp.mpf(float(t[k])), ibox[0][0])
                bias = gb if which == 0 else ab
                return [mp.mpf(float(y[k][i])) - (m[i] + mp.mpf(float(bias[i]))) for i in range(3)]
            assert rel(oi["r"][k], f(fun())) < 1e-12
            assert rel(oi["Js"][k][21:24].reshape(3, 1), np.array(mr.jacobian(fun, [(ibox, 0, 0)]))) < JAC_TOL, (which, k)

"""Host-side part of the reference's Python surface (no GPU): the reference's own tests restated where they only exercise containers,
constructors, locks and error behaviour -- trajectories/test_spline_trajectories.py:81-180, 224-245, trajectories/test_split.py, test_imu.py:26-67."""
import numpy as np
import pytest

from kontiki_b200.sensors import AtanCamera, BasicImu, ConstantBiasImu, PinholeCamera
from kontiki_b200.trajectories import SplitTrajectory, UniformR3SplineTrajectory, UniformSE3SplineTrajectory, UniformSO3SplineTrajectory

spline_classes = (UniformR3SplineTrajectory, UniformSO3SplineTrajectory)


def _random_spline(cls, rng):
    dt, t0 = rng.uniform(0.1, 2.0), rng.uniform(-2, 2)
    traj = cls(dt, t0)
    if cls is UniformR3SplineTrajectory:
        cps = rng.uniform(-5, 5, size=(10, 3))
        new = lambda: rng.uniform(-5, 5, size=3)
    else:
        cps = rng.uniform(-3, 3, size=(10, 4))
        cps /= np.linalg.norm(cps, axis=1).reshape(-1, 1)

        def new():
            q = rng.uniform(-3, 3, size=4)
            return q / np.linalg.norm(q)
    for cp in cps:
        traj.append_knot(cp)
    return traj, cps, new


@pytest.mark.parametrize("cls", spline_classes + (UniformSE3SplineTrajectory,))
def test_construct_default(cls):                                   # test_spline_trajectories.py:81-86
    traj = cls()
    assert traj.dt == 1.0 and traj.t0 == 0.0 and len(traj) == 0


@pytest.mark.parametrize("cls", spline_classes)
@pytest.mark.parametrize("dt,t0", [(1.0, 0.0), (0.5, 1.0), (0.5, -1.0), (1.0, None), (0.5, None)])
def test_construct_params_ok(cls, dt, t0):                         # :89-105
    traj = cls(dt) if t0 is None else cls(dt, t0)
    assert len(traj) == 0 and traj.dt == dt and traj.t0 == (0.0 if t0 is None else t0)


@pytest.mark.parametrize("cls", spline_classes)
def test_control_points_indices_and_assignment(cls):               # :107-146
    traj, cps, new = _random_spline(cls, np.random.default_rng(0))
    assert len(traj) == len(cps) == 10
    np.testing.assert_allclose(np.vstack([cp for cp in traj]), cps, rtol=0, atol=0)
    n = len(traj)
    for neg in range(-n, 0):
        np.testing.assert_equal(traj[neg], traj[neg + n])
    for i in (-(n + 1), -(n + 2), n, n + 1):
        with pytest.raises(IndexError):
            traj[i]
    for i in (0, 2, -2):
        cp = new()
        traj[i] = cp
        np.testing.assert_equal(traj[i], cp)


@pytest.mark.parametrize("cls", spline_classes)
def test_empty_spline_invalid_times(cls):                          # :148-156
    instance = cls()
    with pytest.raises(ValueError):
        instance.min_time
    with pytest.raises(ValueError):
        instance.max_time


@pytest.mark.parametrize("cls", spline_classes)
def test_extend_to_fill(cls):                                      # :159-176
    rng = np.random.default_rng(1)
    dt, t0 = rng.uniform(0.05, 2.0), rng.uniform(-3, 3)
    instance = cls(dt, t0)
    n = int(rng.integers(6, 12))
    new_tmax = t0 + (n - 3) * dt
    instance.extend_to(new_tmax, np.zeros(3) if cls is UniformR3SplineTrajectory else np.array([1.0, 0.0, 0.0, 0.0]))
    assert len(instance) == n
    np.testing.assert_almost_equal(instance.max_time, new_tmax)


def test_so3_require_unit_quaternion():                            # :224-229
    traj = UniformSO3SplineTrajectory()
    traj.append_knot(np.array([1.0, 0.0, 0.0, 0.0]))
    with pytest.raises(ValueError):
        traj.append_knot(np.array([1.0, 1.0, 1.0, 1.0]))


def test_se3_require_se3_elements():                               # :233-end
    traj = UniformSE3SplineTrajectory()
    c, s = np.cos(0.3), np.sin(0.3)
    good = np.eye(4)
    good[:3, :3] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    good[:3, 3] = [0.1, -0.2, 0.3]
    traj.append_knot(good)
    np.testing.assert_allclose(traj[0], good, atol=1e-15)
    bad_det = good.copy()
    bad_det[:3, :3] = np.array([[-1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    with pytest.raises(ValueError):
        traj.append_knot(bad_det)
    bad_row = good.copy()
    bad_row[3] = [0, 0, 1, 1]
    with pytest.raises(ValueError):
        traj.append_knot(bad_row)
    assert len(traj) == 1


def test_split_constructors():                                     # trajectories/test_split.py:7-52
    t = SplitTrajectory()
    assert (t.R3_spline.dt, t.R3_spline.t0, t.SO3_spline.dt, t.SO3_spline.t0) == (1.0, 0.0, 1.0, 0.0)
    with pytest.raises(ValueError):
        t.valid_time
    t = SplitTrajectory(0.3, 0.7)
    assert t.R3_spline.dt == 0.3 and t.SO3_spline.dt == 0.7
    t = SplitTrajectory(0.3, 0.7, -1.5, 2.5)
    assert (t.R3_spline.dt, t.SO3_spline.dt, t.R3_spline.t0, t.SO3_spline.t0) == (0.3, 0.7, -1.5, 2.5)
    r3, so3 = UniformR3SplineTrajectory(0.4, 1.0), UniformSO3SplineTrajectory(0.6, -1.0)
    t = SplitTrajectory(r3, so3)
    assert t.R3_spline is r3 and t.SO3_spline is so3


def test_locks_and_clone():                                        # trajectory_helper.h:25-33, spline_helpers.h
    traj, _, _ = _random_spline(UniformR3SplineTrajectory, np.random.default_rng(2))
    assert not traj.locked
    traj.locked = True
    c = traj.clone()
    assert c.locked and c is not traj and len(c) == len(traj) and c.dt == traj.dt and c.t0 == traj.t0
    c[0] = np.ones(3)
    assert not np.array_equal(c[0], traj[0])                       # a deep copy
    s = SplitTrajectory(traj, UniformSO3SplineTrajectory())
    with pytest.raises(RuntimeError):
        s.locked                                                   # split_trajectory.h: different lock status of the two halves
    s.locked = False
    assert not s.R3_spline.locked and not s.SO3_spline.locked


def test_constant_bias_imu_surface():                              # test_imu.py:26-67
    imu = ConstantBiasImu()
    np.testing.assert_equal(imu.accelerometer_bias, 0)
    np.testing.assert_equal(imu.gyroscope_bias, 0)
    rng = np.random.default_rng(3)
    a, g = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
    imu = ConstantBiasImu(a, g)
    np.testing.assert_equal(imu.accelerometer_bias, a)
    np.testing.assert_equal(imu.gyroscope_bias, g)
    imu.accelerometer_bias, imu.gyroscope_bias = g, a
    np.testing.assert_equal(imu.accelerometer_bias, g)
    np.testing.assert_equal(imu.gyroscope_bias, a)
    assert imu.gyroscope_bias_locked and imu.accelerometer_bias_locked
    imu.gyroscope_bias_locked = False
    assert not imu.gyroscope_bias_locked
    imu.accelerometer_bias_locked = False
    assert not imu.accelerometer_bias_locked
    for cls in (BasicImu, ConstantBiasImu):
        assert callable(cls().gyroscope) and callable(cls().accelerometer)      # imu_helper.h:24-31


def test_sensor_defaults_and_relative_pose():                      # sensors.h:91-109, sensors_helper.h:12-35
    for s in (BasicImu(), PinholeCamera(480, 640, 0.03), AtanCamera(480, 640, 0.03, np.eye(3), (0.1, 0.2), 0.9)):
        assert s.relative_orientation_locked and s.relative_position_locked and s.time_offset_locked
        assert s.time_offset == 0.0 and s.max_time_offset == 0.1
        q, p = s.relative_pose
        np.testing.assert_equal(q, [1.0, 0.0, 0.0, 0.0])
        np.testing.assert_equal(p, 0.0)
        s.relative_pose = (np.array([0.0, 1.0, 0.0, 0.0]), np.array([1.0, 2.0, 3.0]))      # 180 degrees about x
        X = np.array([0.5, -1.0, 2.0])
        np.testing.assert_allclose(s.from_trajectory(X), [1.5, 3.0, 1.0], atol=1e-15)
        np.testing.assert_allclose(s.to_trajectory(s.from_trajectory(X)), X, atol=1e-15)


def test_camera_project_unproject_match_the_oracle():             # test_cameras.py:32-37 on the Python classes, and against the restated reference
    from oracle import kto
    import fixtures_ref as fx
    rng = np.random.default_rng(5)
    K = np.array([[900.0, 0, 960], [0, 900, 540], [0, 0, 1]])
    cams = [(PinholeCamera(fx.IMAGE_ROWS, fx.IMAGE_COLS, fx.CAMERA_READOUT, K), kto.Camera(fx.IMAGE_ROWS, fx.IMAGE_COLS, fx.CAMERA_READOUT, K=K)),
            (AtanCamera(fx.IMAGE_ROWS, fx.IMAGE_COLS, fx.CAMERA_READOUT, fx.ATAN_K, fx.ATAN_WC, fx.ATAN_GAMMA),
             kto.Camera(fx.IMAGE_ROWS, fx.IMAGE_COLS, fx.CAMERA_READOUT, K=fx.ATAN_K, wc=fx.ATAN_WC, gamma=fx.ATAN_GAMMA))]
    for cam, ocam in cams:
        assert (cam.rows, cam.cols, cam.readout) == (fx.IMAGE_ROWS, fx.IMAGE_COLS, fx.CAMERA_READOUT)
        for _ in range(25):
            y = np.array([rng.uniform(0, cam.cols), rng.uniform(0, cam.rows)])
            X = cam.unproject(y) * rng.uniform(0.01, 10)
            np.testing.assert_almost_equal(cam.project(X), y)
            np.testing.assert_allclose(cam.unproject(y), kto.camera_unproject(ocam, y), rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(cam.project(X), kto.camera_project(ocam, X)[0], rtol=1e-12, atol=1e-9)


def test_safe_time_helpers():                                      # python/kontiki/utils.py, used throughout the reference's tests
    from kontiki_b200.utils import safe_time, safe_time_span
    traj, _, _ = _random_spline(UniformR3SplineTrajectory, np.random.default_rng(7))
    t1, t2 = traj.valid_time
    assert t1 < safe_time(traj) < t2
    a, b = safe_time_span(traj, 0.5 * (t2 - t1))
    assert a >= t1 and b <= t2 and abs((b - a) - 0.5 * (t2 - t1)) < 1e-12
    with pytest.raises(ValueError):
        safe_time_span(traj, 2.0 * (t2 - t1))
    assert safe_time_span(traj, 2.0 * (t2 - t1), allow_shorter=True) == (t1, t2)

    class Unbounded:
        valid_time = (-np.inf, np.inf)
    assert np.isfinite(safe_time(Unbounded())) and safe_time_span(Unbounded(), 3.0)[1] - safe_time_span(Unbounded(), 3.0)[0] == 3.0
    with pytest.raises(ValueError):
        safe_time(UniformR3SplineTrajectory())                     # empty spline: no valid time

"""Measurement classes with the reference's Python surface (python/src/kontiki/measurements/py_gyroscope_measurement.cc:23-32,
py_accelerometer_measurement.cc, py_static_rscamera_measurement.cc:41-56, measurement_helper.h:19-25).  error / measure /
project evaluate ONE row through the CUDA library; the batched path is TrajectoryEstimator."""
import numpy as np

from . import _lib
from .trajectories import SplitTrajectory, UniformSE3SplineTrajectory, evaluable


def _problem_for(traj):
    """A ktk problem bound to `traj` and the knots argument for it."""
    traj = evaluable(traj)
    from . import trajectories as _tj
    p = _lib.Problem(_tj.DEFAULT_DEVICE)
    if isinstance(traj, UniformSE3SplineTrajectory):
        traj._check()
        p.set_se3_spline(traj.dt, traj.t0, len(traj), traj.compat_zero_dB)
        return p, traj.control_points
    if isinstance(traj, SplitTrajectory):
        r, s = traj.R3_spline, traj.SO3_spline
        r._check()
        s._check()
        p.set_split_spline(r.dt, r.t0, len(r), s.dt, s.t0, len(s))
        return p, (r.control_points, s.control_points)
    raise TypeError(f"No CUDA evaluation path for {type(traj).__name__} (UniformSE3SplineTrajectory and SplitTrajectory are built)")


class _ImuMeasurement:
    _add = None

    def __init__(self, imu, t, x, weight=1.0):
        self.imu, self.t, self.weight = imu, float(t), float(weight)
        self._x = np.asarray(x, float).reshape(3).copy()

    def _bias(self):
        name = "gyroscope_bias" if self._add == "add_gyroscope" else "accelerometer_bias"
        return getattr(self.imu, name, None)

    def _residual(self, trajectory, x, weight):
        p, knots = _problem_for(trajectory)
        g = getattr(p, self._add)(self.imu._c_sensor(), [self.t], np.asarray(x, float).reshape(1, 3), [weight])
        if self._bias() is not None:
            p.set_group_bias(g, self._bias())
        return p.evaluate(knots, None, _lib.EVAL_RESIDUALS)[0]["r"][0]

    def error(self, trajectory):
        return self._residual(trajectory, self._x, self.weight)

    def measure(self, trajectory):
        # error = weight (x - measure)   (gyroscope_measurement.h:36-38)
        return -self._residual(trajectory, np.zeros(3), 1.0)


class GyroscopeMeasurement(_ImuMeasurement):
    """GyroscopeMeasurement(imu, t, w[, weight])"""
    _add = "add_gyroscope"
    w = property(lambda self: self._x.copy())


class AccelerometerMeasurement(_ImuMeasurement):
    """AccelerometerMeasurement(imu, t, a[, weight])"""
    _add = "add_accelerometer"
    a = property(lambda self: self._x.copy())


class PositionMeasurement:
    """PositionMeasurement(t, p)  (measurements/position_measurement.h:17-31, py_position_measurement.cc): error = p - position(t)."""

    def __init__(self, t, p):
        self.t = float(t)
        self.p = np.asarray(p, float).reshape(3).copy()

    def _residual(self, trajectory, p):
        prob, knots = _problem_for(trajectory)
        prob.add_position([self.t], np.asarray(p, float).reshape(1, 3))
        return prob.evaluate(knots, None, _lib.EVAL_RESIDUALS)[0]["r"][0]

    def error(self, trajectory):
        return self._residual(trajectory, self.p)

    def measure(self, trajectory):
        return -self._residual(trajectory, np.zeros(3))


class OrientationMeasurement:
    """OrientationMeasurement(t, q)  (measurements/orientation_measurement.h:17-31, py_orientation_measurement.cc): q = (w, x, y, z) as the
    reference's Eigen::Quaterniond(qvec(0..3)) constructor reads it; error = q.angularDistance(orientation(t)), a scalar."""

    def __init__(self, t, q):
        self.t = float(t)
        self.q = np.asarray(q, float).reshape(4).copy()

    @property
    def _q_xyzw(self):
        return np.array([self.q[1], self.q[2], self.q[3], self.q[0]])

    def error(self, trajectory):
        prob, knots = _problem_for(trajectory)
        prob.add_orientation([self.t], self._q_xyzw[None, :])
        return float(prob.evaluate(knots, None, _lib.EVAL_RESIDUALS)[0]["r"][0, 0])

    def measure(self, trajectory):
        """trajectory.Orientation(t) as (w, x, y, z) (orientation_measurement.h:24-26)."""
        return trajectory.orientation(self.t)


class StaticRsCameraMeasurement:
    """StaticRsCameraMeasurement(camera, observation[, huber_c=5[, weight=1]])  (static_rscamera_measurement.h:62-69)"""

    _add = "add_static_rs"

    def __init__(self, camera, observation, huber_c=5.0, weight=1.0):
        self.camera, self.observation, self.huber_c, self.weight = camera, observation, float(huber_c), float(weight)

    def _row(self):
        obs = self.observation
        ref = obs.landmark.reference
        return dict(obs_uv=obs.uv[None, :], obs_t0=[obs.view.t0], ref_uv=ref.uv[None, :], ref_t0=[ref.view.t0], rho=obs.landmark.inverse_depth)

    def _residual(self, trajectory, weight):
        p, knots = _problem_for(trajectory)
        r = self._row()
        getattr(p, self._add)(self.camera._c_camera(), r["obs_uv"], r["obs_t0"], r["ref_uv"], r["ref_t0"], [0], [weight])
        return p.evaluate(knots, np.array([r["rho"]]), _lib.EVAL_RESIDUALS)[0]["r"][0]      # no loss: error() is the raw residual

    def error(self, trajectory):
        return self._residual(trajectory, self.weight)

    def project(self, trajectory):
        # error = weight (uv - project)   (static_rscamera_measurement.h:89-94)
        return self.observation.uv - self._residual(trajectory, 1.0)

    measure = project


class NewtonRsCameraMeasurement(StaticRsCameraMeasurement):
    """NewtonRsCameraMeasurement(camera, observation[, huber_c=5[, weight=1]])  (newton_rscamera_measurement.h:127-141): the row
    time of the projection is found by a 5-step Newton iteration instead of being read off the observed row."""
    _add = "add_newton_rs"


class LiftingRsCameraMeasurement(StaticRsCameraMeasurement):
    """LiftingRsCameraMeasurement(camera, observation[, huber_c=5[, weight=1]])  (lifting_rscamera_measurement.h:60-79): the row time of the
    observation is a parameter of the measurement, `vt` in [0, 1] (frame-normalised, initially observation.v / rows); error is the
    3-vector weight * [uv - project ; rows * (vt - vt_orig)] (:98-118)."""
    _add = "add_lifting_rs"

    def __init__(self, camera, observation, huber_c=5.0, weight=1.0):
        super().__init__(camera, observation, huber_c, weight)
        self.vt_orig = float(observation.uv[1]) / float(camera.rows)
        self.vt = self.vt_orig

    def _residual(self, trajectory, weight):
        p, knots = _problem_for(trajectory)
        r = self._row()
        g = p.add_lifting_rs(self.camera._c_camera(), r["obs_uv"], r["obs_t0"], r["ref_uv"], r["ref_t0"], [0], [weight])
        p.set_group_vt(g, [self.vt])
        return p.evaluate(knots, np.array([r["rho"]]), _lib.EVAL_RESIDUALS)[0]["r"][0]      # 3 residuals, no loss

    def project(self, trajectory):
        return self.observation.uv - self._residual(trajectory, 1.0)[:2]

    measure = project


"""Sensor classes with the reference's Python surface (python/src/kontiki/sensors/sensors_helper.h:12-35, camera_help.h:25-49,
py_pinhole_camera.cc, py_atan_camera.cc, py_basic_imu.cc): BasicImu, ConstantBiasImu, PinholeCamera, AtanCamera."""
import numpy as np

from . import _lib


class _Sensor:
    def __init__(self):
        self._q_ct = np.array([0.0, 0.0, 0.0, 1.0])          # x, y, z, w
        self._p_ct = np.zeros(3)
        self.time_offset = 0.0
        self.max_time_offset = 0.1                           # sensors.h:107
        self.relative_orientation_locked = True
        self.relative_position_locked = True
        self.time_offset_locked = True

    @property
    def relative_pose(self):
        q = self._q_ct
        return (np.array([q[3], q[0], q[1], q[2]]), self._p_ct.copy())      # (w, x, y, z), p

    @relative_pose.setter
    def relative_pose(self, pose):
        q, p = pose
        q = np.asarray(q, float)
        self._q_ct = np.array([q[1], q[2], q[3], q[0]])
        self._p_ct = np.asarray(p, float).copy()

    def from_trajectory(self, X):                            # sensors.h:79-81
        from .trajectories import _quat_xyzw_to_rot
        return _quat_xyzw_to_rot(self._q_ct) @ np.asarray(X, float) + self._p_ct

    def to_trajectory(self, X):                              # sensors.h:83-85
        from .trajectories import _quat_xyzw_to_rot
        return _quat_xyzw_to_rot(self._q_ct).T @ (np.asarray(X, float) - self._p_ct)

    def _c_sensor(self):
        return _lib.make_sensor(self._q_ct, self._p_ct, self.time_offset, self.max_time_offset, self.relative_orientation_locked,
                                self.relative_position_locked, self.time_offset_locked)


class _Imu(_Sensor):
    """python/src/kontiki/sensors/imu_helper.h: every IMU exposes accelerometer(trajectory, t) / gyroscope(trajectory, t) -- the model value that
    the measurement classes compare against (sensors/imu.h:32-45), evaluated by the CUDA path like measurement.measure()."""

    def gyroscope(self, trajectory, t):
        from .measurements import GyroscopeMeasurement
        return GyroscopeMeasurement(self, t, np.zeros(3)).measure(trajectory)

    def accelerometer(self, trajectory, t):
        from .measurements import AccelerometerMeasurement
        return AccelerometerMeasurement(self, t, np.zeros(3)).measure(trajectory)


class BasicImu(_Imu):
    """sensors/basic_imu.h:26-33."""


class ConstantBiasImu(_Imu):
    """sensors/constant_bias_imu.h: BasicImu + constant accelerometer / gyroscope biases (both locked by default, :83-97)."""

    def __init__(self, accelerometer_bias=(0, 0, 0), gyroscope_bias=(0, 0, 0)):
        super().__init__()
        self.accelerometer_bias = np.asarray(accelerometer_bias, float).copy()
        self.gyroscope_bias = np.asarray(gyroscope_bias, float).copy()
        self.accelerometer_bias_locked = True
        self.gyroscope_bias_locked = True


class PinholeCamera(_Sensor):
    """sensors/pinhole_camera.h; PinholeCamera(rows, cols, readout[, camera_matrix])."""

    def __init__(self, rows, cols, readout, camera_matrix=None):
        super().__init__()
        self.rows, self.cols, self.readout = int(rows), int(cols), float(readout)
        self.camera_matrix = np.eye(3) if camera_matrix is None else np.asarray(camera_matrix, float).reshape(3, 3).copy()

    def project(self, X):                                    # pinhole_camera.h:47-51
        p = self.camera_matrix @ np.asarray(X, float)
        return p[:2] / p[2]

    def unproject(self, y):                                  # pinhole_camera.h:63-67
        return np.linalg.inv(self.camera_matrix) @ np.array([y[0], y[1], 1.0])

    def _c_camera(self):
        return _lib.make_camera(self.rows, self.cols, self.readout, self.camera_matrix, q_ct=self._q_ct, p_ct=self._p_ct, time_offset=self.time_offset,
                                max_time_offset=self.max_time_offset, q_locked=self.relative_orientation_locked,
                                p_locked=self.relative_position_locked, time_offset_locked=self.time_offset_locked)


class AtanCamera(PinholeCamera):
    """sensors/atan_camera.h; AtanCamera(rows, cols, readout, camera_matrix, wc, gamma) (py_atan_camera.cc:23-27): pinhole camera
    matrix plus the FOV / arctangent distortion model with centre wc and parameter gamma."""

    def __init__(self, rows, cols, readout, camera_matrix=None, wc=(0.0, 0.0), gamma=1.0):
        super().__init__(rows, cols, readout, camera_matrix)
        self.wc = np.asarray(wc, float).reshape(2).copy()
        self.gamma = float(gamma)

    def project(self, X):                                    # atan_camera.h:54-75
        X = np.asarray(X, float)
        eps = 1e-32
        L = X[:2] / (X[2] + eps) - self.wc
        r = np.sqrt(L @ L + eps)
        f = np.arctan(r * self.gamma) / self.gamma
        Y = np.array([*(self.wc + f * L / r), 1.0])
        return (self.camera_matrix @ Y)[:2]

    def unproject(self, y):                                  # atan_camera.h:92-103
        eps = 1e-32
        phn = np.linalg.inv(self.camera_matrix) @ np.array([y[0], y[1], 1.0])
        L = phn[:2] - self.wc
        r = np.sqrt(L @ L + eps)
        f = np.tan(r * self.gamma) / self.gamma
        return np.array([*(self.wc + f * L / r), 1.0])

    def _c_camera(self):
        c = super()._c_camera()
        c.model = _lib.CAMERA_ATAN
        c.wc[:] = [float(self.wc[0]), float(self.wc[1])]
        c.gamma = self.gamma
        return c

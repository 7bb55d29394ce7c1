"""Trajectory classes with the reference's Python surface (python/src/kontiki/trajectories/*.cc, spline_helpers.h:26-48,
trajectory_helper.h:12-34): control points live in host numpy arrays (the reference keeps one heap block per knot,
entity/paramstore/dynamic_pstore.h:27); every evaluation goes through the CUDA library (ktk_traj_evaluate)."""
import numpy as np

from . import _lib

DEFAULT_DEVICE = 0           # CUDA device of the point queries below (estimators take their own `device`); see set_default_device()


def set_default_device(device):
    """Selects the GPU that trajectory / measurement point queries (position(t), m.error(traj), ...) run on."""
    global DEFAULT_DEVICE
    DEFAULT_DEVICE = int(device)


_EPS_SOPHUS = 1e-10          # Sophus::Constants<double>::epsilon(), py_uniform_se3_spline_trajectory.cc:24-38
_EPS_UNIT = 1e-5             # math/quaternion_math.h:11


def _rot_to_quat_xyzw(R):
    """Rotation matrix -> unit quaternion (x,y,z,w), Eigen's algorithm (what Sophus::SE3d(Matrix4d) does)."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0)
        w = 0.5 * s
        s = 0.5 / s
        q = np.array([(R[2, 1] - R[1, 2]) * s, (R[0, 2] - R[2, 0]) * s, (R[1, 0] - R[0, 1]) * s, w])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q = np.zeros(4)
        q[i] = 0.5 * s
        s = 0.5 / s
        q[3] = (R[k, j] - R[j, k]) * s
        q[j] = (R[j, i] + R[i, j]) * s
        q[k] = (R[k, i] + R[i, k]) * s
    return q / np.linalg.norm(q)


def _quat_xyzw_to_rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


class _Trajectory:
    """trajectory_helper.h:12-34.  Every accessor also takes an ARRAY of times (one batched ktk_traj_evaluate; the reference's
    accessors are scalar, so a scalar t returns exactly what the reference returns)."""
    _locked = False

    def _query(self, t):
        raise NotImplementedError

    def _cached_problem(self, key, setup):
        """One ktk problem per (trajectory shape, device), re-used by the point queries: creating a problem costs cudaMalloc + pinned
        allocations, far more than the query."""
        key = (DEFAULT_DEVICE,) + tuple(key)
        cur = getattr(self, "_qp", None)
        if cur is None or cur[0] != key:
            if cur is not None:
                cur[1].close()
            p = _lib.Problem(DEFAULT_DEVICE)
            setup(p)
            self._qp = cur = (key, p)
        return cur[1]

    def evaluate_many(self, t):
        """dict(position, velocity, acceleration, orientation (x,y,z,w), angular_velocity) at an array of times, one batched call."""
        return self._query(np.atleast_1d(np.asarray(t, float)))

    def position(self, t):
        return self._query(t)["position"][0].copy()

    def velocity(self, t):
        return self._query(t)["velocity"][0].copy()

    def acceleration(self, t):
        return self._query(t)["acceleration"][0].copy()

    def orientation(self, t):
        x, y, z, w = self._query(t)["orientation"][0]
        return np.array([w, x, y, z])                      # the reference's Python API is (w, x, y, z)

    def angular_velocity(self, t):
        return self._query(t)["angular_velocity"][0].copy()

    def from_world(self, Xw, t):                           # trajectory.h:119-123
        q = self._query(t)
        return _quat_xyzw_to_rot(q["orientation"][0]).T @ (np.asarray(Xw, float) - q["position"][0])

    def to_world(self, Xt, t):                             # trajectory.h:125-128
        q = self._query(t)
        return _quat_xyzw_to_rot(q["orientation"][0]) @ np.asarray(Xt, float) + q["position"][0]

    @property
    def valid_time(self):
        return (self.min_time, self.max_time)

    @property
    def locked(self):
        return self._locked

    @locked.setter
    def locked(self, v):
        self._locked = bool(v)


class _Spline(_Trajectory):
    """spline_helpers.h:26-48 + spline_base.h:30-62 (segment meta of the owning entity: one segment = the whole spline)."""
    _width = 0

    def __init__(self, dt=1.0, t0=0.0):
        if not dt > 0:
            raise ValueError("dt must be positive")
        self._dt, self._t0 = float(dt), float(t0)
        self._cp = np.zeros((0, self._width))
        self._locked = False

    dt = property(lambda self: self._dt)
    t0 = property(lambda self: self._t0)

    def __len__(self):
        return len(self._cp)

    def _index(self, i):
        n = len(self._cp)
        if i >= n or i < -n:
            raise IndexError("index out of range")
        return i % n if n else i

    def __getitem__(self, i):
        return self._to_py(self._cp[self._index(i)])

    def __setitem__(self, i, cp):
        self._cp[self._index(i)] = self._from_py(cp)

    def append_knot(self, cp):
        self._cp = np.vstack([self._cp, self._from_py(cp)[None, :]])

    def extend_to(self, t, fill_value):                    # spline_base.h:355-359
        cp = self._from_py(fill_value)
        while len(self._cp) < 4 or self.max_time < t:
            self._cp = np.vstack([self._cp, cp[None, :]])

    def _check(self):
        if len(self._cp) < 4:
            raise ValueError("Spline had too few control points")      # std::range_error, spline_base.h:57-61

    @property
    def min_time(self):
        self._check()
        return self._t0

    @property
    def max_time(self):
        self._check()
        return self._t0 + (len(self._cp) - 3) * self._dt

    def clone(self):
        c = type(self)(self._dt, self._t0)
        c._cp = self._cp.copy()
        c._locked = self._locked
        return c

    @property
    def control_points(self):
        """Raw parameter blocks, (n, width) in the reference's storage order (read/write view)."""
        return self._cp


class UniformSE3SplineTrajectory(_Spline):
    """Control points are 4x4 matrices T = [R, p; 0, 1] (py_uniform_se3_spline_trajectory.cc:17-38); stored as
    [qx qy qz qw tx ty tz] (uniform_se3_spline_trajectory.h:27,57)."""
    _width = 7

    def __init__(self, dt=1.0, t0=0.0, compat_zero_dB=False):
        super().__init__(dt, t0)
        self.compat_zero_dB = bool(compat_zero_dB)

    def clone(self):
        c = super().clone()
        c.compat_zero_dB = self.compat_zero_dB
        return c

    @staticmethod
    def _from_py(T):
        T = np.asarray(T, float)
        if T.shape != (4, 4):
            raise ValueError("control point must be a 4x4 matrix")
        if abs(np.linalg.det(T[:3, :3]) - 1) >= _EPS_SOPHUS:
            raise ValueError("Rotation matrix determinant is not 1!")          # std::domain_error
        if ((T[3] - np.array([0, 0, 0, 1.0])) ** 2).sum() >= _EPS_SOPHUS:
            raise ValueError("Final row must be [0, 0, 0, 1]")
        return np.concatenate([_rot_to_quat_xyzw(T[:3, :3]), T[:3, 3]])

    @staticmethod
    def _to_py(cp):
        T = np.eye(4)
        T[:3, :3] = _quat_xyzw_to_rot(cp[:4])
        T[:3, 3] = cp[4:7]
        return T

    def _query(self, t):
        self._check()
        n, c = len(self._cp), self.compat_zero_dB
        p = self._cached_problem(("se3", self._dt, self._t0, n, c), lambda q: q.set_se3_spline(self._dt, self._t0, n, c))
        return p.traj_evaluate(self._cp, np.atleast_1d(np.asarray(t, float)))

    def evaluate(self, t):
        """(P, P', P'') as 4x4 matrices (py_uniform_se3_spline_trajectory.cc:53-60)."""
        self._check()
        n = len(self._cp)
        p = self._cached_problem(("se3", self._dt, self._t0, n, False), lambda q: q.set_se3_spline(self._dt, self._t0, n, False))
        m = p.se3_evaluate_matrices(self._cp, [float(t)])[0]
        return m[0], m[1], m[2]


class UniformR3SplineTrajectory(_Spline):
    _width = 3

    @staticmethod
    def _from_py(cp):
        cp = np.asarray(cp, float).reshape(-1)
        if cp.shape != (3,):
            raise ValueError("control point must be a 3-vector")
        return cp

    @staticmethod
    def _to_py(cp):
        return cp.copy()

    def _query(self, t):
        self._check()
        n = len(self._cp)
        ident = np.tile(np.array([0.0, 0, 0, 1]), (n, 1))           # uniform_r3_spline_trajectory.h:94-97: identity orientation
        p = self._cached_problem(("r3", self._dt, self._t0, n), lambda q: q.set_split_spline(self._dt, self._t0, n, self._dt, self._t0, n))
        return p.traj_evaluate((self._cp, ident), np.atleast_1d(np.asarray(t, float)))


class UniformSO3SplineTrajectory(_Spline):
    """Control points are unit quaternions (w, x, y, z) in Python, stored (x, y, z, w)
    (py_uniform_so3_spline_trajectory.cc:15-21, uniform_so3_spline_trajectory.h:19-27)."""
    _width = 4

    @staticmethod
    def _from_py(q):
        q = np.asarray(q, float).reshape(-1)
        if q.shape != (4,):
            raise ValueError("control point must be a quaternion (w, x, y, z)")
        if abs(np.linalg.norm(q) - 1) >= _EPS_UNIT:
            raise ValueError("Control point must be unit quaternion!")        # std::domain_error
        return np.array([q[1], q[2], q[3], q[0]])

    @staticmethod
    def _to_py(cp):
        return np.array([cp[3], cp[0], cp[1], cp[2]])

    def _query(self, t):
        self._check()
        n = len(self._cp)
        p = self._cached_problem(("so3", self._dt, self._t0, n), lambda q: q.set_split_spline(self._dt, self._t0, n, self._dt, self._t0, n))
        return p.traj_evaluate((np.zeros((n, 3)), self._cp), np.atleast_1d(np.asarray(t, float)))     # uniform_so3_spline_trajectory.h:52-58: zero position


class SplitTrajectory(_Trajectory):
    """split_trajectory.h:87-140."""

    def __init__(self, r3_dt=1.0, so3_dt=1.0, r3_t0=0.0, so3_t0=0.0):
        if isinstance(r3_dt, UniformR3SplineTrajectory) and isinstance(so3_dt, UniformSO3SplineTrajectory):
            self.R3_spline, self.SO3_spline = r3_dt, so3_dt
        else:
            self.R3_spline = UniformR3SplineTrajectory(r3_dt, r3_t0)
            self.SO3_spline = UniformSO3SplineTrajectory(so3_dt, so3_t0)

    @property
    def min_time(self):
        return max(self.R3_spline.min_time, self.SO3_spline.min_time)

    @property
    def max_time(self):
        return min(self.R3_spline.max_time, self.SO3_spline.max_time)

    @property
    def locked(self):
        if self.R3_spline.locked != self.SO3_spline.locked:
            raise RuntimeError("R3 and SO3 trajectories have different lock status!")
        return self.R3_spline.locked

    @locked.setter
    def locked(self, v):
        self.R3_spline.locked = v
        self.SO3_spline.locked = v

    def clone(self):
        return SplitTrajectory(self.R3_spline.clone(), self.SO3_spline.clone())

    def _query(self, t):
        self.R3_spline._check()
        self.SO3_spline._check()
        t = np.atleast_1d(np.asarray(t, float))
        if not ((self.min_time <= t) & (t < self.max_time)).all():
            raise ValueError(f"t={t} is out of range for the trajectory")
        r, s = self.R3_spline, self.SO3_spline
        key = ("split", r.dt, r.t0, len(r), s.dt, s.t0, len(s))
        p = self._cached_problem(key, lambda q: q.set_split_spline(r.dt, r.t0, len(r), s.dt, s.t0, len(s)))
        return p.traj_evaluate((r.control_points, s.control_points), t)


class _LoneSplineView(SplitTrajectory):
    """How a lone UniformR3SplineTrajectory / UniformSO3SplineTrajectory is evaluated: the reference's R3 view returns the identity
    orientation and zero angular velocity (uniform_r3_spline_trajectory.h:61-65), its SO3 view zero position / velocity / acceleration
    (uniform_so3_spline_trajectory.h:50-54) -- i.e. a split trajectory whose other half is a CONSTANT, LOCKED companion spline on the
    same knot grid.  The companion has no parameter blocks in the reference; it is never optimised nor counted here."""

    def __init__(self, spline):
        self.lone = spline
        self._sync()

    def _sync(self):
        sp = self.lone
        n = max(len(sp), 4)
        if isinstance(sp, UniformR3SplineTrajectory):
            comp = UniformSO3SplineTrajectory(sp.dt, sp.t0)
            comp._cp = np.tile(np.array([0.0, 0.0, 0.0, 1.0]), (n, 1))
            self.R3_spline, self.SO3_spline = sp, comp
        else:
            comp = UniformR3SplineTrajectory(sp.dt, sp.t0)
            comp._cp = np.zeros((n, 3))
            self.R3_spline, self.SO3_spline = comp, sp
        comp._locked, comp._companion = True, True
        self.companion = comp

    def refreshed(self):
        if len(self.companion) != max(len(self.lone), 4) or self.companion.dt != self.lone.dt or self.companion.t0 != self.lone.t0:
            self._sync()
        return self


def evaluable(traj):
    """The trajectory the CUDA path evaluates: SE3 and split trajectories as they are, a lone R3 / SO3 spline through _LoneSplineView."""
    if isinstance(traj, (UniformR3SplineTrajectory, UniformSO3SplineTrajectory)):
        v = getattr(traj, "_lone_view", None)
        if v is None:
            v = traj._lone_view = _LoneSplineView(traj)
        return v.refreshed()
    return traj


"""TrajectoryEstimator with the reference's Python surface (python/src/kontiki/py_trajectory_estimator.cc:59-79,
cpplib/include/kontiki/trajectory_estimator.h): add_measurement() collects measurements into batched groups, every
residual + Jacobian evaluation runs on the GPU through the C ABI, solve() is a Levenberg-Marquardt loop around it.

The reference delegates solve() to Ceres (SPARSE_SCHUR trust region).  Here the evaluation -- the hot path this repository
is about -- is the CUDA library; the linear algebra of the step is host-side sparse Cholesky (scipy), i.e. the control
plane stays on the host exactly as in the reference.  SURVEY.md section 8f-1 ("next"): normal equations on the device.
"""
import enum
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import _lib
from .measurements import (AccelerometerMeasurement, GyroscopeMeasurement, LiftingRsCameraMeasurement, NewtonRsCameraMeasurement, OrientationMeasurement, PositionMeasurement,
                           StaticRsCameraMeasurement, _problem_for)
from .trajectories import SplitTrajectory, UniformR3SplineTrajectory, UniformSE3SplineTrajectory, UniformSO3SplineTrajectory, evaluable


class CallbackReturnType(enum.Enum):       # py_ceres.cc:61-66
    Abort = 0
    Continue = 1
    TerminateSuccessfully = 2


class TerminationType(enum.Enum):          # py_ceres.cc:96-103
    Convergence = 0
    NoConvergence = 1
    Failure = 2
    UserSuccess = 3
    UserFailure = 4


class IterationSummary:                    # py_ceres.cc:68-93
    def __init__(self, **kw):
        self.iteration = 0
        self.step_is_valid = self.step_is_successful = True
        self.step_is_nonmonotonic = False
        self.cost = self.cost_change = self.gradient_max_norm = self.gradient_norm = self.step_norm = 0.0
        self.relative_decrease = self.trust_region_radius = self.eta = self.step_size = 0.0
        self.line_search_function_evaluations = self.line_search_gradient_evaluations = self.line_search_iterations = 0
        self.linear_solver_iterations = 0
        self.iteration_time_in_seconds = self.step_solver_time_in_seconds = self.cumulative_time_in_seconds = 0.0
        self.__dict__.update(kw)


class Summary:                             # py_ceres.cc:15-58 (the fields this solver can fill)
    def __init__(self):
        self.message = ""
        self.initial_cost = self.final_cost = self.fixed_cost = 0.0
        self.iterations = []
        self.termination_type = TerminationType.NoConvergence
        self.num_successful_steps = self.num_unsuccessful_steps = self.num_inner_iteration_steps = 0
        self.preprocessor_time_in_seconds = self.minimizer_time_in_seconds = self.postprocessor_time_in_seconds = 0.0
        self.total_time_in_seconds = self.linear_solver_time_in_seconds = 0.0
        self.residual_evaluation_time_in_seconds = self.jacobian_evaluation_time_in_seconds = 0.0
        self.num_parameter_blocks = self.num_parameters = self.num_effective_parameters = 0
        self.num_residual_blocks = self.num_residuals = 0
        self.num_parameter_blocks_reduced = self.num_parameters_reduced = self.num_effective_parameters_reduced = 0
        self.num_residual_blocks_reduced = self.num_residuals_reduced = 0
        self.is_constrained = False
        self.num_threads_given = self.num_threads_used = 1

    def IsSolutionUsable(self):
        return self.termination_type in (TerminationType.Convergence, TerminationType.NoConvergence, TerminationType.UserSuccess)

    def BriefReport(self):
        return (f"kontiki_b200 LM report: Iterations: {len(self.iterations)}, Initial cost: {self.initial_cost:e}, "
                f"Final cost: {self.final_cost:e}, Termination: {self.termination_type.name}")

    def FullReport(self):
        lines = [self.BriefReport(), f"Parameters {self.num_parameters} (reduced {self.num_parameters_reduced}), residuals {self.num_residuals}",
                 f"Time: residual eval {self.residual_evaluation_time_in_seconds:.4f} s, jacobian eval {self.jacobian_evaluation_time_in_seconds:.4f} s, "
                 f"linear solver {self.linear_solver_time_in_seconds:.4f} s, total {self.total_time_in_seconds:.4f} s", self.message]
        return "\n".join(lines)


def _se3_plus_jacobian(cp):
    """d Plus(T, delta)/d delta at 0 for Plus = T * exp([upsilon; omega]) (uniform_se3_spline_trajectory.h:25-48): (n, 7, 6)."""
    n = len(cp)
    P = np.zeros((n, 7, 6))
    x, y, z, w = cp[:, 0], cp[:, 1], cp[:, 2], cp[:, 3]
    # quaternion rows (x,y,z,w) x omega columns: 1/2 (w I + hat(v)) ; -1/2 v^T
    P[:, 0, 3], P[:, 0, 4], P[:, 0, 5] = 0.5 * w, -0.5 * z, 0.5 * y
    P[:, 1, 3], P[:, 1, 4], P[:, 1, 5] = 0.5 * z, 0.5 * w, -0.5 * x
    P[:, 2, 3], P[:, 2, 4], P[:, 2, 5] = -0.5 * y, 0.5 * x, 0.5 * w
    P[:, 3, 3], P[:, 3, 4], P[:, 3, 5] = -0.5 * x, -0.5 * y, -0.5 * z
    # translation rows x upsilon columns: R
    P[:, 4, 0] = 1 - 2 * (y * y + z * z); P[:, 4, 1] = 2 * (x * y - w * z); P[:, 4, 2] = 2 * (x * z + w * y)
    P[:, 5, 0] = 2 * (x * y + w * z); P[:, 5, 1] = 1 - 2 * (x * x + z * z); P[:, 5, 2] = 2 * (y * z - w * x)
    P[:, 6, 0] = 2 * (x * z - w * y); P[:, 6, 1] = 2 * (y * z + w * x); P[:, 6, 2] = 1 - 2 * (x * x + y * y)
    return P


def _se3_plus(cp, delta):
    """T * exp(delta), delta (n,6) = [upsilon; omega] (Sophus order)."""
    from .synthetic import _so3_exp, quat_to_rot, so3_exp_quat
    ups, om = delta[:, :3], delta[:, 3:]
    _, V = _so3_exp(om)
    R = quat_to_rot(cp[:, :4])
    t = cp[:, 4:7] + (R @ (V @ ups[..., None]))[..., 0]
    dq = so3_exp_quat(om)
    a, b = cp[:, :4], dq
    q = np.stack([a[:, 3] * b[:, 0] + a[:, 0] * b[:, 3] + a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1],
                  a[:, 3] * b[:, 1] + a[:, 1] * b[:, 3] + a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                  a[:, 3] * b[:, 2] + a[:, 2] * b[:, 3] + a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0],
                  a[:, 3] * b[:, 3] - a[:, 0] * b[:, 0] - a[:, 1] * b[:, 1] - a[:, 2] * b[:, 2]], 1)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.concatenate([q, t], 1)


def _quat_plus_jacobian(q):
    """ceres::EigenQuaternionParameterization: Plus(q, d) = q_d * q, q_d = (sin|d| d/|d|, cos|d|): d q/d d at 0, (n, 4, 3)."""
    n = len(q)
    P = np.zeros((n, 4, 3))
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    P[:, 0, 0], P[:, 0, 1], P[:, 0, 2] = w, z, -y          # vec = d w + d x v  => w I - hat(v)
    P[:, 1, 0], P[:, 1, 1], P[:, 1, 2] = -z, w, x
    P[:, 2, 0], P[:, 2, 1], P[:, 2, 2] = y, -x, w
    P[:, 3, 0], P[:, 3, 1], P[:, 3, 2] = -x, -y, -z
    return P


def _quat_plus(q, d):
    nrm = np.linalg.norm(d, axis=1, keepdims=True)
    k = np.where(nrm > 0, np.sin(nrm) / np.maximum(nrm, 1e-300), 1.0)
    a = np.concatenate([k * d, np.cos(nrm)], 1)
    b = q
    out = np.stack([a[:, 3] * b[:, 0] + a[:, 0] * b[:, 3] + a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1],
                    a[:, 3] * b[:, 1] + a[:, 1] * b[:, 3] + a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                    a[:, 3] * b[:, 2] + a[:, 2] * b[:, 3] + a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0],
                    a[:, 3] * b[:, 3] - a[:, 0] * b[:, 0] - a[:, 1] * b[:, 1] - a[:, 2] * b[:, 2]], 1)
    return out / np.linalg.norm(out, axis=1, keepdims=True)


class _NoSensor:
    """PositionMeasurement has no sensor: nothing to unlock."""
    relative_orientation_locked = relative_position_locked = time_offset_locked = True


_NO_SENSOR = _NoSensor()


class TrajectoryEstimator:
    """TrajectoryEstimator(trajectory) -- same surface as the reference; `device` selects the GPU."""

    def __init__(self, trajectory, device=0):
        # trajectory_estimator.h / py_trajectory_estimator.cc declare the estimator for all four trajectory classes
        if not isinstance(trajectory, (UniformSE3SplineTrajectory, SplitTrajectory, UniformR3SplineTrajectory, UniformSO3SplineTrajectory)):
            raise TypeError(f"No TrajectoryEstimator declared for {type(trajectory)}")
        self._trajectory, self._device = trajectory, device
        self._measurements = []
        self._callbacks = []
        self._problem = None
        self._built_for = None

    trajectory = property(lambda self: self._trajectory)
    # what the CUDA path evaluates: the trajectory itself, or a lone R3 / SO3 spline with its constant, locked companion (trajectories.evaluable)
    _tr = property(lambda self: evaluable(self._trajectory))

    def add_measurement(self, m):
        if not isinstance(m, (GyroscopeMeasurement, AccelerometerMeasurement, StaticRsCameraMeasurement, PositionMeasurement, OrientationMeasurement)):
            raise TypeError(f"unsupported measurement type {type(m).__name__}")
        # AddToEstimator checks the time span when the measurement is added (trajectory_estimator.h:97-122)
        tr = self._tr
        if isinstance(m, StaticRsCameraMeasurement):
            ref = m.observation.landmark.reference
            t1, t2 = sorted([ref.view.t0, m.observation.view.t0])
            if not m.camera.time_offset_locked:      # static_rscamera_measurement.h:153-157: the first span moves earlier, the second later
                t1, t2 = t1 - m.camera.max_time_offset, t2 + m.camera.max_time_offset
            spans = [(t1 - 1e-3, t1 + m.camera.readout + 1e-3), (t2 - 1e-3, t2 + m.camera.readout + 1e-3)]
        else:
            sensor = getattr(m, "imu", None)
            if sensor is not None and not sensor.time_offset_locked:      # gyroscope_measurement.h:88-91
                spans = [(m.t - sensor.max_time_offset, m.t + sensor.max_time_offset)]
            else:
                spans = [(m.t, m.t)]
        for a, b in spans:
            if a < tr.min_time or b >= tr.max_time:
                raise ValueError("Time span out of range for trajectory")
        self._measurements.append(m)
        self._problem = None

    def add_callback(self, callback, update_state=False):
        self._callbacks.append((callback, bool(update_state)))

    # ---- flattening ------------------------------------------------------------------------
    def _traj_key(self):
        tr = self._tr
        spl = [tr] if isinstance(tr, UniformSE3SplineTrajectory) else [tr.R3_spline, tr.SO3_spline]
        return tuple((s.dt, s.t0, len(s)) for s in spl)

    def _build(self):
        if self._problem is not None and self._built_for == self._traj_key():
            return
        if self._problem is not None:      # knots were appended (or the spline re-gridded) since the problem was flattened
            self._problem.close()
            self._problem = None
        self._built_for = self._traj_key()
        p, _ = _problem_for(self._tr)
        p.close()
        p = _lib.Problem(self._device)
        tr = self._tr
        if isinstance(tr, UniformSE3SplineTrajectory):
            p.set_se3_spline(tr.dt, tr.t0, len(tr), tr.compat_zero_dB)
        else:
            p.set_split_spline(tr.R3_spline.dt, tr.R3_spline.t0, len(tr.R3_spline), tr.SO3_spline.dt, tr.SO3_spline.t0, len(tr.SO3_spline))
        groups, self._landmarks, lm_index = {}, [], {}
        for i, m in enumerate(self._measurements):
            kind = type(m)
            sensor = m.camera if issubclass(kind, StaticRsCameraMeasurement) else (_NO_SENSOR if kind in (PositionMeasurement, OrientationMeasurement) else m.imu)
            groups.setdefault((kind, id(sensor)), (sensor, []))[1].append(i)
        self._groups = []
        for (kind, _), (sensor, rows) in groups.items():
            ms = [self._measurements[i] for i in rows]
            if issubclass(kind, StaticRsCameraMeasurement):
                lm = []
                for m in ms:
                    L = m.observation.landmark
                    if id(L) not in lm_index:
                        lm_index[id(L)] = len(self._landmarks)
                        self._landmarks.append(L)
                    lm.append(lm_index[id(L)])
                g = getattr(p, kind._add)(sensor._c_camera(), np.array([m.observation.uv for m in ms]), [m.observation.view.t0 for m in ms],
                                    np.array([m.observation.landmark.reference.uv for m in ms]), [m.observation.landmark.reference.view.t0 for m in ms],
                                    lm, [m.weight for m in ms], [m.huber_c for m in ms])
                self._groups.append(dict(g=g, kind="cam", rows=rows, lm=np.array(lm, np.int64), sensor=sensor, newton=kind is NewtonRsCameraMeasurement,
                                         lifting=kind is LiftingRsCameraMeasurement, ms=ms, huber=np.array([m.huber_c for m in ms], float)))
            elif kind is PositionMeasurement:
                g = p.add_position([m.t for m in ms], np.array([m.p for m in ms]))
                self._groups.append(dict(g=g, kind="pos", rows=rows, sensor=sensor, weight=np.ones(len(ms))))
            elif kind is OrientationMeasurement:
                g = p.add_orientation([m.t for m in ms], np.array([m._q_xyzw for m in ms]))
                self._groups.append(dict(g=g, kind="ori", rows=rows, sensor=sensor, weight=np.ones(len(ms))))
            else:
                fn = p.add_gyroscope if kind is GyroscopeMeasurement else p.add_accelerometer
                g = fn(sensor._c_sensor(), [m.t for m in ms], np.array([m._x for m in ms]), [m.weight for m in ms])
                self._groups.append(dict(g=g, kind="gyro" if kind is GyroscopeMeasurement else "accel", rows=rows, sensor=sensor,
                                         weight=np.array([m.weight for m in ms])))
        self._problem = p

    def _point(self):
        tr = self._tr
        knots = tr.control_points if isinstance(tr, UniformSE3SplineTrajectory) else (tr.R3_spline.control_points, tr.SO3_spline.control_points)
        rho = np.array([L.inverse_depth for L in self._landmarks]) if self._landmarks else None
        return knots, rho

    @staticmethod
    def _sensor_free(sensor):
        """Names of the sensor's unlocked parameter blocks, in the reference's block order (sensors.h:139-161, constant_bias_imu.h:106-118)."""
        free = []
        cam = hasattr(sensor, "camera_matrix")
        if cam and not sensor.relative_orientation_locked:
            free.append("q")                 # an IMU's relative pose is not applied by the reference (TODO.md:6): zero columns, left out
        if cam and not sensor.relative_position_locked:
            free.append("p")
        if not sensor.time_offset_locked:
            free.append("d")
        if not getattr(sensor, "accelerometer_bias_locked", True):
            free.append("ab")
        if not getattr(sensor, "gyroscope_bias_locked", True):
            free.append("gb")
        return free

    def _push_sensors(self):
        """The sensor parameters are part of the evaluation point (ktk_set_group_sensor / ktk_set_group_bias)."""
        for grp in self._groups:
            if grp.get("lifting"):                       # the row times are parameter blocks of the measurements (ktk_set_group_vt)
                self._problem.set_group_vt(grp["g"], [m.vt for m in grp["ms"]])
            sn = grp["sensor"]
            if self._sensor_free(sn):
                self._problem.set_group_sensor(grp["g"], sn._c_sensor())
            bias = getattr(sn, "gyroscope_bias" if grp["kind"] == "gyro" else "accelerometer_bias", None) if grp["kind"] in ("gyro", "accel") else None
            if bias is not None:
                self._problem.set_group_bias(grp["g"], bias)

    def evaluate(self, jacobians=True, robust=True):
        """One batched residual (+ Jacobian) evaluation at the current state; list of per-group outputs (C ABI layouts)."""
        self._build()
        self._push_sensors()
        knots, rho = self._point()
        flags = _lib.EVAL_RESIDUALS | (_lib.EVAL_JACOBIANS if jacobians else 0) | (_lib.EVAL_ROBUST if robust else 0)
        if jacobians and any(self._sensor_free(g["sensor"]) for g in self._groups):
            flags |= _lib.EVAL_SENSOR_JACOBIANS
        return self._problem.evaluate(knots, rho, flags)

    # ---- local (tangent) sparse Jacobian ----------------------------------------------------------------------------------
    def _columns(self):
        tr = self._tr
        if isinstance(tr, UniformSE3SplineTrajectory):
            blocks = [("se3", tr, 6)]
        else:
            blocks = [("r3", tr.R3_spline, 3), ("so3", tr.SO3_spline, 3)]
        off, layout = 0, {}
        for name, spl, dof in blocks:
            layout[name] = (off, dof, spl)
            off += dof * len(spl) if not spl.locked else 0
        layout["rho"] = (off, 1, None)
        free_lm = np.array([not L.locked for L in self._landmarks], bool) if self._landmarks else np.zeros(0, bool)
        self._lm_col = np.full(len(self._landmarks), -1, np.int64)
        self._lm_col[free_lm] = off + np.arange(free_lm.sum())
        off += int(free_lm.sum())
        # LiftingRs row times: one column per measurement (lifting_rscamera_measurement.h:200-205)
        for grp in getattr(self, "_groups", []):
            if grp.get("lifting"):
                grp["vt_col0"] = off
                off += len(grp["ms"])
        # unlocked sensor parameter blocks, one set of columns per distinct sensor object
        width = {"q": 3, "p": 3, "d": 1, "ab": 3, "gb": 3}
        self._sensor_cols = {}
        for grp in getattr(self, "_groups", []):
            sn = grp["sensor"]
            if id(sn) in self._sensor_cols or sn is _NO_SENSOR:
                continue
            cols = {}
            for name in self._sensor_free(sn):
                cols[name] = off
                off += width[name]
            self._sensor_cols[id(sn)] = (sn, cols)
        return layout, off

    def _sparse_system(self, outs):
        """r (stacked) and J in local coordinates as scipy CSR, from the packed device rows."""
        layout, ncols = self._columns()
        tr = self._tr
        rs, rows_i, cols_i, vals = [], [], [], []
        row0 = 0

        def add_blocks(Jb, first_knot, width_amb, name, nres):
            """Jb: (n, nk, nres, width_amb) ambient blocks at knots first_knot + k (nk = 4; the observation span of a Newton-RS row)."""
            off, dof, spl = layout[name]
            if spl.locked:
                return
            n = len(Jb)
            cp = spl.control_points
            if name == "se3":
                P = _se3_plus_jacobian(cp)
            elif name == "so3":
                P = _quat_plus_jacobian(cp)
            else:
                P = None
            for k in range(Jb.shape[1]):
                kn = np.minimum(first_knot + k, len(spl) - 1)      # blocks past a Newton-RS row's own span are zero
                Jl = Jb[:, k] if P is None else np.einsum("nra,nad->nrd", Jb[:, k], P[kn])
                r_idx = (row0 + nres * np.arange(n)[:, None, None] + np.arange(nres)[None, :, None]) + np.zeros((1, 1, dof), np.int64)
                c_idx = (off + dof * kn)[:, None, None] + np.arange(dof)[None, None, :] + np.zeros((1, nres, 1), np.int64)
                rows_i.append(r_idx.reshape(-1)); cols_i.append(c_idx.reshape(-1)); vals.append(Jl.reshape(-1))

        split = isinstance(tr, SplitTrajectory)
        for grp in self._groups:
            o = outs[grp["g"]]
            n = len(o["r"])
            if grp["kind"] == "cam" and split and (grp.get("lifting") or grp.get("newton")):
                # span rows on a split trajectory: [ref R3 4x(nres x 3) | ref SO3 4x(nres x 4) | obs R3 Wa x(..) | obs SO3 Wb x(..) | (vt nres) | rho nres]
                J = o["J"].reshape(n, -1)
                nres = 3 if grp.get("lifting") else 2
                Wa, Wb = self._problem.group_span_windows(grp["g"])
                o_rb, o_oa, o_ob, o_t = nres * 12, nres * 28, nres * (28 + 3 * Wa), nres * (28 + 3 * Wa + 4 * Wb)
                add_blocks(J[:, :o_rb].reshape(n, 4, nres, 3), o["i0"], 3, "r3", nres)
                add_blocks(J[:, o_rb:o_oa].reshape(n, 4, nres, 4), o["i0_c"], 4, "so3", nres)
                add_blocks(J[:, o_oa:o_ob].reshape(n, Wa, nres, 3), o["i0_b"], 3, "r3", nres)
                add_blocks(J[:, o_ob:o_t].reshape(n, Wb, nres, 4), o["i0_d"], 4, "so3", nres)
                r_idx = row0 + nres * np.arange(n)[:, None] + np.arange(nres)[None, :]
                if grp.get("lifting"):
                    rows_i.append(r_idx.reshape(-1)); cols_i.append(np.repeat(grp["vt_col0"] + np.arange(n), 3)); vals.append(J[:, o_t:o_t + 3].reshape(-1))
                    o_t += 3
                col = self._lm_col[grp["lm"]]
                free = col >= 0
                rows_i.append(r_idx[free].reshape(-1)); cols_i.append(np.repeat(col[free], nres)); vals.append(J[free, o_t:o_t + nres].reshape(-1))
            elif grp["kind"] == "cam" and grp.get("lifting"):
                J = o["J"].reshape(n, -1)
                nrow = J.shape[1]                                  # 90 + 21 W: [ref 4x(3x7) | obs W x(3x7) | vt 3 | rho 3]
                add_blocks(J[:, :84].reshape(n, 4, 3, 7), o["i0"], 7, "se3", 3)
                add_blocks(J[:, 84:nrow - 6].reshape(n, -1, 3, 7), o["i0_b"], 7, "se3", 3)
                r_idx = row0 + 3 * np.arange(n)[:, None] + np.arange(3)[None, :]
                rows_i.append(r_idx.reshape(-1)); cols_i.append(np.repeat(grp["vt_col0"] + np.arange(n), 3)); vals.append(J[:, nrow - 6:nrow - 3].reshape(-1))
                col = self._lm_col[grp["lm"]]
                free = col >= 0
                rows_i.append(r_idx[free].reshape(-1)); cols_i.append(np.repeat(col[free], 3)); vals.append(J[free, nrow - 3:nrow].reshape(-1))
                nres = 3
            elif grp["kind"] == "cam":
                J = o["J"].reshape(n, -1)
                nrow = J.shape[1]                                  # 114; Newton-RS 58 + 14 W
                if not split:
                    add_blocks(J[:, :56].reshape(n, 4, 2, 7), o["i0"], 7, "se3", 2)
                    add_blocks(J[:, 56:nrow - 2].reshape(n, -1, 2, 7), o["i0_b"], 7, "se3", 2)
                else:
                    add_blocks(J[:, 0:24].reshape(n, 4, 2, 3), o["i0"], 3, "r3", 2)
                    add_blocks(J[:, 24:56].reshape(n, 4, 2, 4), o["i0_c"], 4, "so3", 2)
                    add_blocks(J[:, 56:80].reshape(n, 4, 2, 3), o["i0_b"], 3, "r3", 2)
                    add_blocks(J[:, 80:112].reshape(n, 4, 2, 4), o["i0_d"], 4, "so3", 2)
                col = self._lm_col[grp["lm"]]
                free = col >= 0
                r_idx = (row0 + 2 * np.arange(n)[:, None] + np.arange(2)[None, :])[free]
                rows_i.append(r_idx.reshape(-1)); cols_i.append(np.repeat(col[free], 2)); vals.append(J[free, nrow - 2:nrow].reshape(-1))
                nres = 2
            elif grp["kind"] == "ori":
                nres = 1
                if not split:
                    add_blocks(o["J"].reshape(n, 4, 1, 7), o["i0"], 7, "se3", 1)
                else:
                    add_blocks(o["J"].reshape(n, 4, 1, 4), o["i0_c"], 4, "so3", 1)
            else:
                nres = 3
                if not split:
                    add_blocks(o["J"].reshape(n, 4, 3, 7), o["i0"], 7, "se3", 3)
                elif grp["kind"] == "gyro":
                    add_blocks(o["J"].reshape(n, 4, 3, 4), o["i0_c"], 4, "so3", 3)
                elif grp["kind"] == "pos":
                    add_blocks(o["J"].reshape(n, 4, 3, 3), o["i0"], 3, "r3", 3)
                else:
                    add_blocks(o["J"][:, :36].reshape(n, 4, 3, 3), o["i0"], 3, "r3", 3)
                    add_blocks(o["J"][:, 36:].reshape(n, 4, 3, 4), o["i0_c"], 4, "so3", 3)
            # sensor-block columns (KTK_EVAL_SENSOR_JACOBIANS)
            sn, scols = self._sensor_cols.get(id(grp["sensor"]), (None, {}))
            rr = row0 + nres * np.arange(n)[:, None] + np.arange(nres)[None, :]           # (n, nres) global residual rows
            if "d" in scols:
                Jd = o["Js"][:, 7 * nres:8 * nres] if grp["kind"] == "cam" else o["Js"][:, 0:3]
                rows_i.append(rr.reshape(-1)); cols_i.append(np.full(rr.size, scols["d"])); vals.append(Jd.reshape(-1))
            if grp["kind"] == "cam" and ("q" in scols or "p" in scols):      # Js = [q_ct (nres x 4) | p_ct (nres x 3) | time offset (nres)]
                if "q" in scols:
                    Pq = _quat_plus_jacobian(sn._q_ct[None, :])[0]                         # (4, 3)
                    Jq = o["Js"][:, 0:4 * nres].reshape(n, nres, 4) @ Pq
                    rows_i.append(np.repeat(rr.reshape(-1), 3)); cols_i.append(np.tile(scols["q"] + np.arange(3), rr.size)); vals.append(Jq.reshape(-1))
                if "p" in scols:
                    rows_i.append(np.repeat(rr.reshape(-1), 3)); cols_i.append(np.tile(scols["p"] + np.arange(3), rr.size)); vals.append(o["Js"][:, 4 * nres:7 * nres].reshape(-1))
            bname = {"gyro": "gb", "accel": "ab"}.get(grp["kind"])
            if bname in scols:                                                                # d r / d bias = -weight I
                rows_i.append(rr.reshape(-1)); cols_i.append(np.tile(scols[bname] + np.arange(3), n)); vals.append(np.repeat(-grp["weight"], 3))
            rs.append(o["r"].reshape(-1))
            row0 += nres * n
        r = np.concatenate(rs) if rs else np.zeros(0)
        if vals:
            J = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows_i), np.concatenate(cols_i))), shape=(row0, ncols))
        else:
            J = sp.csr_matrix((row0, ncols))
        return r, J, layout

    def _apply_step(self, delta, layout, rho_lower=0.0):
        for name in ("se3", "r3", "so3"):
            if name not in layout:
                continue
            off, dof, spl = layout[name]
            if spl.locked:
                continue
            d = delta[off:off + dof * len(spl)].reshape(len(spl), dof)
            cp = spl.control_points
            if name == "se3":
                cp[:] = _se3_plus(cp, d)
            elif name == "so3":
                cp[:] = _quat_plus(cp, d)
            else:
                cp[:] = cp + d
        for L, c in zip(self._landmarks, self._lm_col):
            if c >= 0:
                L.inverse_depth = max(rho_lower, L.inverse_depth + delta[c])      # lower bound 0, static_rscamera_measurement.h:178-181
        for grp in self._groups:
            if grp.get("lifting"):                                                # bounds [0, 1], lifting_rscamera_measurement.h:202-204
                for m, dv in zip(grp["ms"], delta[grp["vt_col0"]:grp["vt_col0"] + len(grp["ms"])]):
                    m.vt = float(np.clip(m.vt + dv, 0.0, 1.0))
        for sn, cols in self._sensor_cols.values():
            if "q" in cols:
                sn._q_ct = _quat_plus(sn._q_ct[None, :], delta[cols["q"]:cols["q"] + 3][None, :])[0]
            if "p" in cols:
                sn._p_ct = sn._p_ct + delta[cols["p"]:cols["p"] + 3]
            if "d" in cols:      # bounds +-max_time_offset (sensors.h:159-160)
                sn.time_offset = float(np.clip(sn.time_offset + delta[cols["d"]], -sn.max_time_offset, sn.max_time_offset))
            if "ab" in cols:
                sn.accelerometer_bias = sn.accelerometer_bias + delta[cols["ab"]:cols["ab"] + 3]
            if "gb" in cols:
                sn.gyroscope_bias = sn.gyroscope_bias + delta[cols["gb"]:cols["gb"] + 3]

    def _snapshot(self):
        tr = self._tr
        spl = [tr] if isinstance(tr, UniformSE3SplineTrajectory) else [tr.R3_spline, tr.SO3_spline]
        sens = [(sn, sn._q_ct.copy(), sn._p_ct.copy(), sn.time_offset, getattr(sn, "accelerometer_bias", None), getattr(sn, "gyroscope_bias", None))
                for sn, _ in getattr(self, "_sensor_cols", {}).values()]
        vts = [[m.vt for m in g["ms"]] for g in self._groups if g.get("lifting")]
        return [s.control_points.copy() for s in spl], [L.inverse_depth for L in self._landmarks], sens, vts

    def _restore(self, snap):
        tr = self._tr
        spl = [tr] if isinstance(tr, UniformSE3SplineTrajectory) else [tr.R3_spline, tr.SO3_spline]
        for s, c in zip(spl, snap[0]):
            s.control_points[:] = c
        for L, v in zip(self._landmarks, snap[1]):
            L.inverse_depth = v
        for sn, q, pp, d, ab, gb in snap[2]:
            sn._q_ct, sn._p_ct, sn.time_offset = q.copy(), pp.copy(), d
            if ab is not None:
                sn.accelerometer_bias, sn.gyroscope_bias = ab.copy(), gb.copy()
        for g, vt in zip([g for g in self._groups if g.get("lifting")], snap[3]):
            for m, v in zip(g["ms"], vt):
                m.vt = v

    @staticmethod
    def _cost(outs, groups, robust=True):
        """Ceres' cost 1/2 sum rho(s), s = |r|^2 per residual block.  The kernels return the CORRECTED residual r_c = sqrt(rho'(s)) r
        (ceres::internal::Corrector; alpha = 0 for HuberLoss since rho'' <= 0), so |r_c|^2 = a sqrt(s) in the linear region and
        rho(s) = 2 a sqrt(s) - a^2 = 2 |r_c|^2 - a^2 there (s > a^2  <=>  |r_c|^2 > a^2); rho(s) = s = |r_c|^2 in the quadratic region."""
        total = 0.0
        for g in groups:
            s = (outs[g["g"]]["r"] ** 2).sum(1)
            if robust and g["kind"] == "cam":
                a2 = g["huber"] ** 2
                s = np.where(s > a2, 2.0 * s - a2, s)
            total += float(s.sum())
        return 0.5 * total

    def solve(self, max_iterations=50, progress=True, num_threads=-1, linear_solver="auto"):
        """Levenberg-Marquardt on the GPU-evaluated residuals/Jacobians; returns a Summary (py_ceres.cc:15-58 fields).

        linear_solver: "host_cholesky" (rows copied to the host, sparse Cholesky with scipy -- the reference's split: Ceres
        does its linear algebra on the host too), "device_pcg" (rows stay on the GPU, matrix-free PCG, kontiki_b200/gn.py;
        multi-GPU aware), or "auto" (device_pcg above 20 000 measurements)."""
        auto = linear_solver == "auto"
        if auto:
            linear_solver = "device_pcg" if len(self._measurements) > 20000 else "host_cholesky"
        self._build()
        if any(g.get("newton") or g.get("lifting") for g in self._groups):
            if linear_solver == "device_pcg" and not auto:
                raise NotImplementedError("NewtonRs / LiftingRs camera rows are solved by the host_cholesky path only")
            linear_solver = "host_cholesky"
        if any(self._sensor_free(g["sensor"]) for g in self._groups):
            if linear_solver == "device_pcg":
                raise NotImplementedError("unlocked sensor parameters are optimised by the host_cholesky path only")
            linear_solver = "host_cholesky"
        if linear_solver == "device_pcg":
            return self._solve_device(max_iterations, progress)
        if linear_solver != "host_cholesky":
            raise ValueError("linear_solver must be 'auto', 'host_cholesky' or 'device_pcg'")
        t_start = time.perf_counter()
        s = Summary()
        self._build()
        layout, ncols = self._columns()
        tr = self._tr
        real = lambda spl: 0 if getattr(spl, "_companion", False) else len(spl)       # the companion of a lone spline has no parameter blocks
        n_knot_params = (7 * len(tr)) if isinstance(tr, UniformSE3SplineTrajectory) else (3 * real(tr.R3_spline) + 4 * real(tr.SO3_spline))
        n_sensors = sum(g["kind"] not in ("pos", "ori") for g in self._groups)        # one sensor per group (none for PositionMeasurement): q_ct(4) p_ct(3) time_offset(1), constant (sensors.h:135-165)
        n_vt = sum(len(g["ms"]) for g in self._groups if g.get("lifting"))
        s.num_parameters = n_knot_params + len(self._landmarks) + 8 * n_sensors + n_vt
        n_knot_blocks = len(tr) if isinstance(tr, UniformSE3SplineTrajectory) else real(tr.R3_spline) + real(tr.SO3_spline)
        s.num_parameter_blocks = n_knot_blocks + len(self._landmarks) + 3 * n_sensors + n_vt
        s.num_parameters_reduced = ncols + self._ambient_free(layout)      # ambient sizes of the non-constant blocks
        s.num_effective_parameters_reduced = ncols
        s.num_residual_blocks = s.num_residual_blocks_reduced = len(self._measurements)
        t0 = time.perf_counter()
        outs = self.evaluate(jacobians=True)
        s.jacobian_evaluation_time_in_seconds += time.perf_counter() - t0
        cost = self._cost(outs, self._groups)
        s.initial_cost = cost
        s.num_residuals = s.num_residuals_reduced = sum(outs[g["g"]]["r"].size for g in self._groups)
        radius, nu = 1e4, 2.0                               # Ceres defaults: initial_trust_region_radius 1e4
        term = TerminationType.NoConvergence
        if progress:
            print("iter      cost      cost_change  |gradient|   |step|    tr_ratio  tr_radius")
        if ncols == 0:
            s.final_cost, s.termination_type, s.message = cost, TerminationType.Convergence, "no free parameters"
            return s
        for it in range(max_iterations + 1):
            t_it = time.perf_counter()
            r, J, layout = self._sparse_system(outs)
            g = J.T @ r
            gmax = float(np.abs(g).max()) if g.size else 0.0
            isum = IterationSummary(iteration=it, cost=cost, gradient_max_norm=gmax, gradient_norm=float(np.linalg.norm(g)), trust_region_radius=radius)
            if it == 0:
                s.iterations.append(isum)
                if progress:
                    print(f"{it:4d}  {cost:12.6e}  {0.0:10.2e}  {gmax:10.2e}  {0.0:9.2e}  {0.0:9.2e}  {radius:9.2e}")
            if gmax < 1e-10:                                # Ceres gradient_tolerance
                term, s.message = TerminationType.Convergence, "Gradient tolerance reached"
                break
            if it == max_iterations:
                s.message = "Maximum number of iterations reached"
                break
            H = (J.T @ J).tocsc()
            D = H.diagonal()
            t_ls = time.perf_counter()
            try:
                delta = spla.spsolve(H + sp.diags(np.clip(D, 1e-6, 1e32) / radius), -g)
            except Exception as e:      # noqa: BLE001
                term, s.message = TerminationType.Failure, f"linear solver failed: {e}"
                break
            s.linear_solver_time_in_seconds += time.perf_counter() - t_ls
            snap = self._snapshot()
            self._apply_step(delta, layout)
            t0 = time.perf_counter()
            try:
                outs_new = self.evaluate(jacobians=True)
                cost_new = self._cost(outs_new, self._groups)
            except ValueError:
                cost_new, outs_new = np.inf, None
            s.jacobian_evaluation_time_in_seconds += time.perf_counter() - t0
            model = -float(delta @ (g + 0.5 * (H @ delta)))
            rho_ratio = (cost - cost_new) / model if model > 0 else -1.0
            step_norm = float(np.linalg.norm(delta))
            ok = np.isfinite(cost_new) and rho_ratio > 1e-3
            isum = IterationSummary(iteration=it + 1, cost=cost_new if ok else cost, cost_change=cost - cost_new, step_norm=step_norm, relative_decrease=rho_ratio,
                                    trust_region_radius=radius, step_is_successful=bool(ok), gradient_max_norm=gmax,
                                    iteration_time_in_seconds=time.perf_counter() - t_it, cumulative_time_in_seconds=time.perf_counter() - t_start)
            if ok:
                rel = (cost - cost_new) / max(cost, 1e-300)
                cost, outs = cost_new, outs_new
                radius = radius / max(1.0 / 3.0, 1.0 - (2.0 * rho_ratio - 1.0) ** 3)
                nu = 2.0
                s.num_successful_steps += 1
            else:
                self._restore(snap)
                radius, nu = radius / nu, 2 * nu
                s.num_unsuccessful_steps += 1
                rel = 1.0
            s.iterations.append(isum)
            if progress:
                print(f"{it + 1:4d}  {isum.cost:12.6e}  {isum.cost_change:10.2e}  {gmax:10.2e}  {step_norm:9.2e}  {rho_ratio:9.2e}  {radius:9.2e}")
            stop = None
            for cb, _ in self._callbacks:
                res = cb(isum)
                if res is not None and res is not CallbackReturnType.Continue:
                    stop = res
            if stop is CallbackReturnType.Abort:
                term, s.message = TerminationType.UserFailure, "User callback returned SOLVER_ABORT"
                break
            if stop is CallbackReturnType.TerminateSuccessfully:
                term, s.message = TerminationType.UserSuccess, "User callback returned SOLVER_TERMINATE_SUCCESSFULLY"
                break
            if ok and rel < 1e-6:                           # Ceres function_tolerance
                term, s.message = TerminationType.Convergence, "Function tolerance reached"
                break
            if radius < 1e-32:
                term, s.message = TerminationType.Failure, "Trust region radius collapsed"
                break
        s.final_cost = cost
        s.termination_type = term
        s.total_time_in_seconds = s.minimizer_time_in_seconds = time.perf_counter() - t_start
        return s

    # ---- Levenberg-Marquardt with the whole step on the device (kontiki_b200/gn.py, csrc/gn_device.cuh) ----------------------------------
    def _solve_device(self, max_iterations, progress, pcg_tol=1e-6, pcg_max_iter=300):
        """What ceres::Solve does with SPARSE_SCHUR (trajectory_estimator.h:38-64), on the GPU: rows stay in device memory, rho is eliminated
        there (implicit Schur complement), the reduced knot system is solved by block-Jacobi preconditioned CG whose scalars never leave the
        device, Plus() and the rho >= 0 bound are kernels.  The trust-region bookkeeping below needs three scalars per iteration."""
        from . import gn
        t_start = time.perf_counter()
        s = Summary()
        self._build()
        tr = self._tr
        split = isinstance(tr, SplitTrajectory)
        spl_a = tr.R3_spline if split else tr
        spl_b = tr.SO3_spline if split else None
        n_a, n_b, n_rho = len(spl_a), (len(spl_b) if split else 0), len(self._landmarks)
        lm_locked = np.array([L.locked for L in self._landmarks], np.uint8) if n_rho else None
        hubers = {grp["g"]: grp["huber"] for grp in self._groups if grp["kind"] == "cam"}
        sol = gn.DeviceSchurSolver(self._problem, split, n_a, n_b, n_rho, self._device, lm_locked, spl_a.locked, bool(split and spl_b.locked), hubers)
        da = 3 if split else 6
        n_free = (0 if spl_a.locked else da * n_a) + (0 if (not split or spl_b.locked) else 3 * n_b) + sum(not L.locked for L in self._landmarks)
        layout, ncols = self._columns()
        n_companion = sum(spl._width * len(spl) for spl in (spl_a, spl_b) if spl is not None and getattr(spl, "_companion", False))
        s.num_parameters = self._problem.num_parameters(n_rho) - n_companion + 8 * sum(g["kind"] not in ("pos", "ori") for g in self._groups)
        s.num_parameters_reduced = ncols + self._ambient_free(layout)
        s.num_effective_parameters_reduced = n_free
        s.num_residual_blocks = s.num_residual_blocks_reduced = len(self._measurements)
        s.num_residuals = s.num_residuals_reduced = sum(self._problem.group_size(g["g"]) * {"cam": 2, "ori": 1}.get(g["kind"], 3) for g in self._groups)

        def write_back():
            kf, rho = sol.point()
            if split:
                spl_a.control_points[:] = kf[:3 * n_a].reshape(n_a, 3)
                spl_b.control_points[:] = kf[3 * n_a:].reshape(n_b, 4)
            else:
                spl_a.control_points[:] = kf.reshape(n_a, 7)
            for L, v in zip(self._landmarks, rho):
                L.inverse_depth = float(v)

        kf0 = np.concatenate([spl_a.control_points.reshape(-1), spl_b.control_points.reshape(-1)]) if split else spl_a.control_points.reshape(-1)
        sol.set_point(kf0, np.array([L.inverse_depth for L in self._landmarks]))
        t0 = time.perf_counter()
        cost = sol.evaluate()
        s.jacobian_evaluation_time_in_seconds += time.perf_counter() - t0
        s.initial_cost = cost
        radius, nu, term = 1e4, 2.0, TerminationType.NoConvergence
        if progress:
            print("iter      cost      cost_change  |gradient|   |step|    tr_ratio  tr_radius  pcg")
        if n_free == 0:
            s.final_cost, s.termination_type, s.message = cost, TerminationType.Convergence, "no free parameters"
            return s
        for it in range(max_iterations + 1):
            t_it = time.perf_counter()
            gmax = float(sol.linearize(radius).item())
            if it == 0:
                s.iterations.append(IterationSummary(iteration=0, cost=cost, gradient_max_norm=gmax, trust_region_radius=radius))
                if progress:
                    print(f"{0:4d}  {cost:12.6e}  {0.0:10.2e}  {gmax:10.2e}  {0.0:9.2e}  {0.0:9.2e}  {radius:9.2e}")
            if gmax < 1e-10:
                term, s.message = TerminationType.Convergence, "Gradient tolerance reached"
                break
            if it == max_iterations:
                s.message = "Maximum number of iterations reached"
                break
            t_ls = time.perf_counter()
            n_cg, _ = sol.solve(radius, pcg_tol, pcg_max_iter)
            model, step_norm = sol.finish()
            s.linear_solver_time_in_seconds += time.perf_counter() - t_ls
            t0 = time.perf_counter()
            try:
                cost_new = sol.evaluate(at_new=True)
            except ValueError:
                cost_new = np.inf
            s.jacobian_evaluation_time_in_seconds += time.perf_counter() - t0
            rho_ratio = (cost - cost_new) / model if model > 0 else -1.0
            ok = np.isfinite(cost_new) and rho_ratio > 1e-3
            isum = IterationSummary(iteration=it + 1, cost=cost_new if ok else cost, cost_change=cost - cost_new, step_norm=step_norm, relative_decrease=rho_ratio,
                                    trust_region_radius=radius, step_is_successful=bool(ok), gradient_max_norm=gmax, linear_solver_iterations=n_cg,
                                    iteration_time_in_seconds=time.perf_counter() - t_it, cumulative_time_in_seconds=time.perf_counter() - t_start)
            if ok:
                rel = (cost - cost_new) / max(cost, 1e-300)
                cost = cost_new
                sol.accept()
                radius = radius / max(1.0 / 3.0, 1.0 - (2.0 * rho_ratio - 1.0) ** 3)
                nu = 2.0
                s.num_successful_steps += 1
            else:
                sol.evaluate()                               # the rows in device memory are the rejected point's: back to the current one
                radius, nu = radius / nu, 2 * nu
                s.num_unsuccessful_steps += 1
                rel = 1.0
            s.iterations.append(isum)
            if progress:
                print(f"{it + 1:4d}  {isum.cost:12.6e}  {isum.cost_change:10.2e}  {gmax:10.2e}  {step_norm:9.2e}  {rho_ratio:9.2e}  {radius:9.2e}  {n_cg}")
            stop = None
            for cb, update_state in self._callbacks:
                if update_state:
                    write_back()
                res = cb(isum)
                if res is not None and res is not CallbackReturnType.Continue:
                    stop = res
            if stop is CallbackReturnType.Abort:
                term, s.message = TerminationType.UserFailure, "User callback returned SOLVER_ABORT"
                break
            if stop is CallbackReturnType.TerminateSuccessfully:
                term, s.message = TerminationType.UserSuccess, "User callback returned SOLVER_TERMINATE_SUCCESSFULLY"
                break
            if ok and rel < 1e-6:
                term, s.message = TerminationType.Convergence, "Function tolerance reached"
                break
            if radius < 1e-32:
                term, s.message = TerminationType.Failure, "Trust region radius collapsed"
                break
        write_back()
        s.final_cost, s.termination_type = cost, term
        s.total_time_in_seconds = s.minimizer_time_in_seconds = time.perf_counter() - t_start
        self._problem.set_stream(0)
        return s

    def _ambient_free(self, layout):
        """Ceres' num_parameters_reduced counts AMBIENT sizes of the non-constant blocks (7 per SE3 knot, ...)."""
        extra = 0
        for name, amb in (("se3", 7), ("r3", 3), ("so3", 4)):
            if name in layout and not layout[name][2].locked:
                extra += (amb - layout[name][1]) * len(layout[name][2])
        return extra

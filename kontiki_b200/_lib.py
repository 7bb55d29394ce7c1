"""ctypes binding of the C ABI declared in include/kontiki_b200.h (libkontiki_b200.so, CUDA sm_100a).

There is no CPU path: if the library is missing, or no sm_100 device is present, every entry point raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

OK, ERANGE, ERUNTIME, EINVAL, ECUDA, EUNSUPPORTED = 0, -1, -2, -3, -4, -5
EVAL_RESIDUALS, EVAL_JACOBIANS, EVAL_ROBUST, EVAL_SENSOR_JACOBIANS, EVAL_LOCAL, EVAL_DEVICE_ORDER = 1, 2, 4, 8, 16, 32
GYROSCOPE, ACCELEROMETER, STATIC_RS, NEWTON_RS, POSITION, ORIENTATION, LIFTING_RS = 0, 1, 2, 3, 4, 5, 6
CAMERA_PINHOLE, CAMERA_ATAN = 0, 1
IMU_ROW, CAM_ROW = 84, 114


class Sensor(C.Structure):
    _fields_ = [("q_ct", C.c_double * 4), ("p_ct", C.c_double * 3), ("time_offset", C.c_double), ("max_time_offset", C.c_double),
                ("q_locked", C.c_int32), ("p_locked", C.c_int32), ("time_offset_locked", C.c_int32)]


class Camera(C.Structure):
    _fields_ = [("base", Sensor), ("rows", C.c_int32), ("cols", C.c_int32), ("readout", C.c_double), ("K", C.c_double * 9),
                ("model", C.c_int32), ("reserved", C.c_int32), ("wc", C.c_double * 2), ("gamma", C.c_double)]


PinholeCamera = Camera


class GroupOut(C.Structure):
    _fields_ = [("r", C.c_void_p), ("J", C.c_void_p), ("i0", C.c_void_p), ("i0_b", C.c_void_p), ("i0_c", C.c_void_p), ("i0_d", C.c_void_p), ("Js", C.c_void_p)]


EXPORTS = ["ktk_last_error", "ktk_problem_create", "ktk_problem_destroy", "ktk_set_stream", "ktk_set_se3_spline", "ktk_add_gyroscope",
           "ktk_add_accelerometer", "ktk_add_static_rs", "ktk_num_groups", "ktk_group_size", "ktk_group_kind", "ktk_evaluate",
           "ktk_evaluate_device", "ktk_synchronize", "ktk_launch_count", "ktk_host_alloc", "ktk_host_free", "ktk_get_structure",
           "ktk_expand_static_rs", "ktk_set_profiling", "ktk_read_profile", "ktk_set_split_spline", "ktk_group_row_size", "ktk_num_knot_doubles",
           "ktk_get_structure_so3", "ktk_traj_evaluate", "ktk_num_parameters", "ktk_j_apply", "ktk_jt_apply", "ktk_jtj_diagonal", "ktk_jtj_diagonal_local", "ktk_set_graphs", "ktk_set_group_sensor", "ktk_set_group_bias", "ktk_group_row_size_local", "ktk_se3_evaluate_matrices", "ktk_get_row_order", "ktk_add_newton_rs", "ktk_add_position", "ktk_add_orientation", "ktk_add_lifting_rs", "ktk_set_group_vt", "ktk_group_span_windows",
           "ktk_gn_prepare", "ktk_gn_cost", "ktk_gn_linearize_local", "ktk_gn_gradient_local", "ktk_gn_linearize_rhs", "ktk_gn_pcg_begin", "ktk_gn_product", "ktk_gn_pcg_update",
           "ktk_gn_pcg_status", "ktk_gn_finish_local", "ktk_gn_finish_mask", "ktk_gn_model_local", "ktk_gn_retract", "ktk_gn_buffer"]

_lib = None


class KontikiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def lib():
    """Loads the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        path = os.environ.get("KTK_LIB", _build.LIB_PATH)
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                              "kontiki_b200 has no CPU fallback.")
        L = C.CDLL(path)
        L.ktk_last_error.restype = C.c_char_p
        L.ktk_group_size.restype = C.c_int64
        L.ktk_launch_count.restype = C.c_int64
        L.ktk_host_alloc.restype = C.c_void_p
        L.ktk_host_alloc.argtypes = [C.c_int64]
        L.ktk_host_free.argtypes = [C.c_void_p]
        L.ktk_problem_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.ktk_problem_destroy.argtypes = [C.c_void_p]
        L.ktk_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.ktk_set_se3_spline.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int32, C.c_int32]
        L.ktk_add_gyroscope.argtypes = [C.c_void_p, C.POINTER(Sensor), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ktk_add_accelerometer.argtypes = L.ktk_add_gyroscope.argtypes
        L.ktk_add_position.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ktk_add_orientation.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.ktk_add_static_rs.argtypes = [C.c_void_p, C.POINTER(Camera), C.c_int64] + [C.c_void_p] * 7
        L.ktk_add_newton_rs.argtypes = L.ktk_add_static_rs.argtypes
        L.ktk_add_lifting_rs.argtypes = L.ktk_add_static_rs.argtypes
        L.ktk_set_group_vt.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.ktk_num_groups.argtypes = [C.c_void_p]
        L.ktk_group_size.argtypes = [C.c_void_p, C.c_int32]
        L.ktk_group_kind.argtypes = [C.c_void_p, C.c_int32]
        L.ktk_evaluate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.POINTER(GroupOut)]
        L.ktk_evaluate_device.argtypes = L.ktk_evaluate.argtypes
        L.ktk_synchronize.argtypes = [C.c_void_p]
        L.ktk_launch_count.argtypes = [C.c_void_p]
        L.ktk_get_structure.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.ktk_expand_static_rs.argtypes = [C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 5
        L.ktk_set_split_spline.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int32, C.c_double, C.c_double, C.c_int32]
        L.ktk_group_row_size.argtypes = [C.c_void_p, C.c_int32]
        L.ktk_group_row_size_local.argtypes = [C.c_void_p, C.c_int32]
        L.ktk_group_span_windows.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.ktk_num_knot_doubles.argtypes = [C.c_void_p]
        L.ktk_num_knot_doubles.restype = C.c_int64
        L.ktk_get_structure_so3.argtypes = L.ktk_get_structure.argtypes
        L.ktk_num_parameters.argtypes = [C.c_void_p, C.c_int64]
        L.ktk_num_parameters.restype = C.c_int64
        L.ktk_j_apply.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(GroupOut), C.c_void_p, C.POINTER(C.c_void_p)]
        L.ktk_jt_apply.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(GroupOut), C.POINTER(C.c_void_p), C.c_void_p]
        L.ktk_jtj_diagonal.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(GroupOut), C.c_void_p]
        L.ktk_get_row_order.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.ktk_jtj_diagonal_local.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(GroupOut), C.c_void_p, C.c_void_p, C.c_void_p]
        L.ktk_set_graphs.argtypes = [C.c_void_p, C.c_int32]
        L.ktk_se3_evaluate_matrices.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ktk_set_group_sensor.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Sensor)]
        L.ktk_set_group_bias.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.ktk_traj_evaluate.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ktk_set_profiling.argtypes = [C.c_void_p, C.c_int32]
        L.ktk_read_profile.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.ktk_gn_prepare.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(GroupOut), C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        for name in ("ktk_gn_cost", "ktk_gn_gradient_local", "ktk_gn_product", "ktk_gn_pcg_update", "ktk_gn_finish_local", "ktk_gn_finish_mask", "ktk_gn_model_local"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.ktk_gn_linearize_local.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ktk_gn_linearize_rhs.argtypes = [C.c_void_p, C.c_double]
        L.ktk_gn_pcg_begin.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int32]
        L.ktk_gn_pcg_status.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        L.ktk_gn_retract.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ktk_gn_buffer.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
        L.ktk_gn_buffer.restype = C.c_int64
        _lib = L
    return _lib


def check(code):
    """Maps C statuses to the exceptions pybind11 gives the reference's (std::range_error -> ValueError, ...)."""
    if code >= 0:
        return code
    msg = lib().ktk_last_error().decode()
    if code == ERANGE:
        raise ValueError(msg)
    if code == EINVAL:
        raise ValueError(msg)
    if code == EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise KontikiError(code, msg)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data


def make_sensor(q_ct=(0, 0, 0, 1), p_ct=(0, 0, 0), time_offset=0.0, max_time_offset=0.1, q_locked=True, p_locked=True, time_offset_locked=True):
    s = Sensor()
    s.q_ct[:] = [float(x) for x in q_ct]
    s.p_ct[:] = [float(x) for x in p_ct]
    s.time_offset, s.max_time_offset = float(time_offset), float(max_time_offset)
    s.q_locked, s.p_locked, s.time_offset_locked = int(q_locked), int(p_locked), int(time_offset_locked)
    return s


def make_camera(rows, cols, readout, K, wc=None, gamma=None, **sensor_kw):
    """PinholeCamera; AtanCamera when the distortion centre wc and gamma are given (sensors/atan_camera.h)."""
    c = Camera()
    c.base = make_sensor(**sensor_kw)
    c.rows, c.cols, c.readout = int(rows), int(cols), float(readout)
    c.K[:] = [float(x) for x in np.asarray(K, float).reshape(-1)]
    if gamma is not None:
        c.model = CAMERA_ATAN
        c.wc[:] = [float(wc[0]), float(wc[1])]
        c.gamma = float(gamma)
    return c


class Problem:
    """Thin object wrapper of ktk_problem (one CUDA device, one SE3 spline, any number of measurement groups)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(lib().ktk_problem_create(int(device), C.byref(self._h)))
        self.device = device
        self.n_knots = 0
        self.split = False

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().ktk_problem_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        check(lib().ktk_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def set_graphs(self, on):
        check(lib().ktk_set_graphs(self._h, int(bool(on))))

    def set_se3_spline(self, dt, t0, n_knots, compat_zero_dB=False):
        check(lib().ktk_set_se3_spline(self._h, float(dt), float(t0), int(n_knots), int(compat_zero_dB)))
        self.n_knots = int(n_knots)

    def set_split_spline(self, dt_r3, t0_r3, n_r3, dt_so3, t0_so3, n_so3):
        check(lib().ktk_set_split_spline(self._h, float(dt_r3), float(t0_r3), int(n_r3), float(dt_so3), float(t0_so3), int(n_so3)))
        self.split, self.n_r3, self.n_so3 = True, int(n_r3), int(n_so3)

    def group_row_size(self, g):
        return lib().ktk_group_row_size(self._h, g)

    def group_span_windows(self, g):
        """NewtonRs / LiftingRs groups: (W, 0) on SE3, (Wa, Wb) = widest observation span on the R3 / SO3 spline of a split trajectory."""
        wa, wb = C.c_int32(0), C.c_int32(0)
        check(lib().ktk_group_span_windows(self._h, int(g), C.byref(wa), C.byref(wb)))
        return wa.value, wb.value

    def _add_imu(self, fn, sensor, t, y, weight):
        t, y = _f64(t), _f64(y).reshape(-1, 3)
        if len(t) != len(y):
            raise ValueError("t and y differ in length")
        w = None if weight is None else _f64(weight)
        return check(fn(self._h, C.byref(sensor), len(t), _ptr(t), _ptr(y), _ptr(w)))

    def add_gyroscope(self, sensor, t, y, weight=None):
        return self._add_imu(lib().ktk_add_gyroscope, sensor, t, y, weight)

    def add_accelerometer(self, sensor, t, y, weight=None):
        return self._add_imu(lib().ktk_add_accelerometer, sensor, t, y, weight)

    def add_position(self, t, position, weight=None):
        t, y = _f64(t), _f64(position).reshape(-1, 3)
        if len(t) != len(y):
            raise ValueError("t and position differ in length")
        w = None if weight is None else _f64(weight)
        return check(lib().ktk_add_position(self._h, len(t), _ptr(t), _ptr(y), _ptr(w)))

    def add_orientation(self, t, q_xyzw):
        """OrientationMeasurement rows: q (n, 4) as (x, y, z, w); one residual per row (the angular distance)."""
        t, q = _f64(t), _f64(q_xyzw).reshape(-1, 4)
        if len(t) != len(q):
            raise ValueError("t and q differ in length")
        return check(lib().ktk_add_orientation(self._h, len(t), _ptr(t), _ptr(q)))

    def add_newton_rs(self, camera, obs_uv, obs_t0, ref_uv, ref_t0, lm_idx, weight=None, huber_c=None):
        return self.add_static_rs(camera, obs_uv, obs_t0, ref_uv, ref_t0, lm_idx, weight, huber_c, _fn=lib().ktk_add_newton_rs)

    def add_lifting_rs(self, camera, obs_uv, obs_t0, ref_uv, ref_t0, lm_idx, weight=None, huber_c=None):
        """LiftingRsCameraMeasurement rows (3 residuals; the row times start at obs_uv.y / rows, set_group_vt moves them)."""
        return self.add_static_rs(camera, obs_uv, obs_t0, ref_uv, ref_t0, lm_idx, weight, huber_c, _fn=lib().ktk_add_lifting_rs)

    def set_group_vt(self, g, vt):
        vt = _f64(vt).reshape(-1)
        if len(vt) != self.group_size(g):
            raise ValueError("vt must have one entry per measurement of the group")
        check(lib().ktk_set_group_vt(self._h, int(g), _ptr(vt)))

    def add_static_rs(self, camera, obs_uv, obs_t0, ref_uv, ref_t0, lm_idx, weight=None, huber_c=None, _fn=None):
        obs_uv, ref_uv = _f64(obs_uv).reshape(-1, 2), _f64(ref_uv).reshape(-1, 2)
        obs_t0, ref_t0 = _f64(obs_t0), _f64(ref_t0)
        lm = np.ascontiguousarray(lm_idx, np.int32)
        n = len(obs_t0)
        if not (len(obs_uv) == len(ref_uv) == len(ref_t0) == len(lm) == n):
            raise ValueError("static-RS arrays differ in length")
        w = None if weight is None else _f64(weight)
        h = None if huber_c is None else _f64(huber_c)
        return check((_fn or lib().ktk_add_static_rs)(self._h, C.byref(camera), n, _ptr(obs_uv), _ptr(obs_t0), _ptr(ref_uv), _ptr(ref_t0), _ptr(lm),
                                                      _ptr(w), _ptr(h)))

    @property
    def num_groups(self):
        return lib().ktk_num_groups(self._h)

    def group_size(self, g):
        return lib().ktk_group_size(self._h, g)

    def group_kind(self, g):
        return lib().ktk_group_kind(self._h, g)

    @property
    def launch_count(self):
        return lib().ktk_launch_count(self._h)

    def set_group_sensor(self, g, sensor):
        check(lib().ktk_set_group_sensor(self._h, int(g), C.byref(sensor)))

    def set_group_bias(self, g, bias):
        b = _f64(bias).reshape(3)
        check(lib().ktk_set_group_bias(self._h, int(g), _ptr(b)))

    def alloc_outputs(self, jacobians=True, sensor_jacobians=False, local=False):
        """Host (numpy) output arrays for every group, in the C ABI's packed layouts."""
        outs = []
        for g in range(self.num_groups):
            n, cam = self.group_size(g), self.group_kind(g) in (STATIC_RS, NEWTON_RS, LIFTING_RS)
            nres = {STATIC_RS: 2, NEWTON_RS: 2, ORIENTATION: 1}.get(self.group_kind(g), 3)
            o = dict(r=np.zeros((n, nres)), i0=np.full(n, -1, np.int32))
            if jacobians:
                if local:
                    o["J"] = np.zeros((n, lib().ktk_group_row_size_local(self._h, g)))
                else:
                    o["J"] = np.zeros((n, 4, nres, 7)) if (not cam and not self.split) else np.zeros((n, self.group_row_size(g)))
            if cam:
                o["i0_b"] = np.full(n, -1, np.int32)
            if self.split:
                o["i0_c"] = np.full(n, -1, np.int32)
                if cam:
                    o["i0_d"] = np.full(n, -1, np.int32)
            if sensor_jacobians:
                o["Js"] = np.zeros((n, (24 if self.group_kind(g) == LIFTING_RS else 16) if cam else 3))
            outs.append(o)
        return outs

    @staticmethod
    def _out_array(outs, getptr):
        arr = (GroupOut * max(len(outs), 1))()
        for i, o in enumerate(outs):
            arr[i].r = getptr(o.get("r"))
            arr[i].J = getptr(o.get("J"))
            arr[i].i0 = getptr(o.get("i0"))
            arr[i].i0_b = getptr(o.get("i0_b"))
            arr[i].i0_c = getptr(o.get("i0_c"))
            arr[i].i0_d = getptr(o.get("i0_d"))
            arr[i].Js = getptr(o.get("Js"))
        return arr

    def evaluate(self, knots, rho=None, flags=EVAL_RESIDUALS | EVAL_JACOBIANS, outs=None):
        """Host-buffer evaluation (ktk_evaluate).  knots: (n_knots, 7) [qx qy qz qw tx ty tz]; returns the list of group outputs."""
        if self.split:      # (r3 knots (n_r3,3), so3 knots (n_so3,4) x,y,z,w) -> [R3 | SO3], the order of split_trajectory.h:34-39
            r3, so3 = _f64(knots[0]), _f64(knots[1])
            if r3.shape != (self.n_r3, 3) or so3.shape != (self.n_so3, 4):
                raise ValueError(f"knots must be ((n_r3={self.n_r3}, 3), (n_so3={self.n_so3}, 4))")
            knots = np.concatenate([r3.reshape(-1), so3.reshape(-1)])
        else:
            knots = _f64(knots)
            if knots.shape != (self.n_knots, 7):
                raise ValueError(f"knots must have shape ({self.n_knots}, 7)")
        rho = None if rho is None else _f64(rho)
        if outs is None:
            outs = self.alloc_outputs(bool(flags & EVAL_JACOBIANS), bool(flags & EVAL_SENSOR_JACOBIANS), bool(flags & EVAL_LOCAL))
        arr = self._out_array(outs, _ptr)
        check(lib().ktk_evaluate(self._h, _ptr(knots), _ptr(rho), 0 if rho is None else len(rho), int(flags), arr))
        return outs

    def _flat_knots(self, knots):
        if self.split:
            r3, so3 = _f64(knots[0]), _f64(knots[1])
            if r3.shape != (self.n_r3, 3) or so3.shape != (self.n_so3, 4):
                raise ValueError(f"knots must be ((n_r3={self.n_r3}, 3), (n_so3={self.n_so3}, 4))")
            return np.concatenate([r3.reshape(-1), so3.reshape(-1)])
        knots = _f64(knots)
        if knots.shape != (self.n_knots, 7):
            raise ValueError(f"knots must have shape ({self.n_knots}, 7)")
        return knots.reshape(-1)

    def traj_evaluate(self, knots, t):
        """ktk_traj_evaluate: dict of position / velocity / acceleration / orientation (x,y,z,w) / angular_velocity at times t."""
        kf = self._flat_knots(knots)
        t = _f64(np.atleast_1d(t))
        out, st = np.zeros((len(t), 16)), np.zeros(len(t), np.int32)
        check(lib().ktk_traj_evaluate(self._h, _ptr(kf), len(t), _ptr(t), _ptr(out), _ptr(st)))
        return dict(position=out[:, 0:3], velocity=out[:, 3:6], acceleration=out[:, 6:9], orientation=out[:, 9:13], angular_velocity=out[:, 13:16])

    def se3_evaluate_matrices(self, knots, t):
        kf = self._flat_knots(knots)
        t = _f64(np.atleast_1d(t))
        out, st = np.zeros((len(t), 3, 4, 4)), np.zeros(len(t), np.int32)
        check(lib().ktk_se3_evaluate_matrices(self._h, _ptr(kf), len(t), _ptr(t), _ptr(out), _ptr(st)))
        return out

    def evaluate_flat(self, knots_flat, rho, flags, outs):
        """ktk_evaluate on an already flattened float64 knot array (the C ABI's layout); outs as from alloc_outputs()."""
        if knots_flat.dtype != np.float64 or knots_flat.size != lib().ktk_num_knot_doubles(self._h):
            raise ValueError("knots_flat must be float64 with ktk_num_knot_doubles() elements")
        arr = self._out_array(outs, _ptr)
        check(lib().ktk_evaluate(self._h, _ptr(knots_flat), _ptr(rho), 0 if rho is None else len(rho), int(flags), arr))
        return outs

    def evaluate_device(self, d_knots_ptr, d_rho_ptr, n_rho, flags, d_outs):
        """Device-buffer evaluation (ktk_evaluate_device): pointers are raw device addresses (e.g. torch data_ptr())."""
        arr = self._out_array(d_outs, lambda p: None if p is None else int(p))
        check(lib().ktk_evaluate_device(self._h, C.c_void_p(int(d_knots_ptr)), None if not d_rho_ptr else C.c_void_p(int(d_rho_ptr)), int(n_rho),
                                        int(flags), arr))

    def set_profiling(self, on):
        check(lib().ktk_set_profiling(self._h, int(bool(on))))

    def read_profile(self, g):
        ms, n = C.c_double(0), C.c_int64(0)
        check(lib().ktk_read_profile(self._h, g, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # ---- matrix-free Gauss-Newton products on device pointers (see kontiki_b200/gn.py) ----------------------------------
    def num_parameters(self, n_rho):
        return lib().ktk_num_parameters(self._h, int(n_rho))

    def _ptr_array(self, ptrs):
        arr = (C.c_void_p * max(len(ptrs), 1))()
        for i, q in enumerate(ptrs):
            arr[i] = None if not q else int(q)
        return arr

    def j_apply(self, d_outs, d_v, d_u, flags=0):
        check(lib().ktk_j_apply(self._h, int(flags), self._out_array(d_outs, lambda q: None if q is None else int(q)), C.c_void_p(int(d_v)), self._ptr_array(d_u)))

    def jt_apply(self, d_outs, d_u, d_y, flags=0):
        check(lib().ktk_jt_apply(self._h, int(flags), self._out_array(d_outs, lambda q: None if q is None else int(q)), self._ptr_array(d_u), C.c_void_p(int(d_y))))

    def jtj_diagonal(self, d_outs, d_y, flags=0):
        check(lib().ktk_jtj_diagonal(self._h, int(flags), self._out_array(d_outs, lambda q: None if q is None else int(q)), C.c_void_p(int(d_y))))

    def get_row_order(self, g):
        order = np.zeros(self.group_size(g), np.int32)
        check(lib().ktk_get_row_order(self._h, int(g), _ptr(order)))
        return order

    def jtj_diagonal_local(self, d_outs, d_Pa, d_Pb, d_y, flags=0):
        check(lib().ktk_jtj_diagonal_local(self._h, int(flags), self._out_array(d_outs, lambda q: None if q is None else int(q)), None if not d_Pa else C.c_void_p(int(d_Pa)),
                                           None if not d_Pb else C.c_void_p(int(d_Pb)), C.c_void_p(int(d_y))))

    def synchronize(self):
        check(lib().ktk_synchronize(self._h))

    # ---- Gauss-Newton step on the device (include/kontiki_b200.h "ktk_gn_*"; driven by kontiki_b200/gn.py) ---------------------------
    def gn_prepare(self, flags, d_outs, n_rho, lm_locked=None, lock_a=False, lock_b=False, hubers=None):
        arr = self._out_array(d_outs, lambda q: None if q is None else int(q))
        locked = None if lm_locked is None else np.ascontiguousarray(lm_locked, np.uint8)
        hub = (C.c_void_p * max(self.num_groups, 1))()
        keep = []
        for g in range(self.num_groups):
            h = None if hubers is None else hubers.get(g)
            if h is not None:
                h = np.ascontiguousarray(h, np.float64)
                keep.append(h)
                hub[g] = h.ctypes.data
        check(lib().ktk_gn_prepare(self._h, int(flags), arr, int(n_rho), None if locked is None else locked.ctypes.data, int(bool(lock_a)), int(bool(lock_b)), hub))

    def gn_call(self, name, *args):
        check(getattr(lib(), "ktk_gn_" + name)(self._h, *args))

    def gn_pcg_status(self):
        it, done, rel = C.c_int32(0), C.c_int32(0), C.c_double(0)
        check(lib().ktk_gn_pcg_status(self._h, C.byref(it), C.byref(done), C.byref(rel)))
        return it.value, bool(done.value), rel.value

    def gn_buffer(self, name):
        """(device pointer, number of doubles) of one of the solver's buffers."""
        ptr = C.c_void_p()
        n = lib().ktk_gn_buffer(self._h, name.encode(), C.byref(ptr))
        return (ptr.value or 0), int(n)

    def get_structure(self, g, cap=16):
        n = self.group_size(g)
        ids, nids = np.full((n, cap), -1, np.int32), np.zeros(n, np.int32)
        check(lib().ktk_get_structure(self._h, g, cap, _ptr(ids), _ptr(nids)))
        return ids, nids

    def get_structure_so3(self, g, cap=16):
        n = self.group_size(g)
        ids, nids = np.full((n, cap), -1, np.int32), np.zeros(n, np.int32)
        check(lib().ktk_get_structure_so3(self._h, g, cap, _ptr(ids), _ptr(nids)))
        return ids, nids

    def expand_static_rs(self, g, ids, J, i0_ref, i0_obs):
        n, cap = ids.shape
        out = np.zeros((n, cap, 3 if self.group_kind(g) == LIFTING_RS else 2, 7))
        J = _f64(J).reshape(n, self.group_row_size(g))
        check(lib().ktk_expand_static_rs(self._h, g, cap, _ptr(np.ascontiguousarray(ids, np.int32)), _ptr(J), _ptr(np.ascontiguousarray(i0_ref, np.int32)),
                                         _ptr(np.ascontiguousarray(i0_obs, np.int32)), _ptr(out)))
        return out

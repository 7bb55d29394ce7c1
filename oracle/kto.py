"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (kontiki_b200/) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_ANALYTIC_PATH = os.path.join(_HERE, "libanalytic_cpu.so")      # the optimised CPU variant of SURVEY.md 8(d), analytic_cpu.cpp

SE3, SPLIT, R3, SO3 = 0, 1, 2, 3
EvalPosition, EvalVelocity, EvalAcceleration, EvalOrientation, EvalAngularVelocity = 1, 2, 4, 8, 16
OK, RANGE_ERROR, RUNTIME_ERROR = 0, -1, -2


def build(force=False):
    """Compile the oracle with the committed Makefile (g++ only, no dependencies)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "kontiki_ref.hpp", "lie.hpp", "dual.hpp", "Makefile")]
    csrc = os.path.join(os.path.dirname(_HERE), "kontiki_b200", "csrc")
    srcs2 = [os.path.join(_HERE, "analytic_cpu.cpp"), os.path.join(_HERE, "Makefile")] + [os.path.join(csrc, f) for f in ("spline_math.cuh", "lie_math.cuh", "dualnum.cuh")]
    stale = lambda lib, deps: not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in deps)
    if force or stale(_LIB_PATH, srcs) or stale(_ANALYTIC_PATH, srcs2):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), env={**os.environ, "CXX": "g++"})
    return _LIB_PATH


class _Traj(C.Structure):
    _fields_ = [("kind", C.c_int), ("dt_a", C.c_double), ("t0_a", C.c_double), ("n_a", C.c_int), ("knots_a", C.c_void_p),
                ("dt_b", C.c_double), ("t0_b", C.c_double), ("n_b", C.c_int), ("knots_b", C.c_void_p),
                ("compat_zero_dB", C.c_int), ("locked", C.c_int)]


class _Sensor(C.Structure):
    _fields_ = [("q_ct", C.c_double * 4), ("p_ct", C.c_double * 3), ("time_offset", C.c_double), ("max_time_offset", C.c_double),
                ("q_locked", C.c_int), ("p_locked", C.c_int), ("d_locked", C.c_int), ("has_bias", C.c_int),
                ("abias", C.c_double * 3), ("gbias", C.c_double * 3), ("abias_locked", C.c_int), ("gbias_locked", C.c_int)]


class _Camera(C.Structure):
    _fields_ = [("readout", C.c_double), ("rows", C.c_int), ("cols", C.c_int), ("K", C.c_double * 9), ("model", C.c_int), ("method", C.c_int),
                ("wc", C.c_double * 2), ("gamma", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.kto_last_error.restype = C.c_char_p
        _lib.kto_huber_correct.restype = C.c_double
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Traj:
    """Trajectory description.  SE3 knots: (n,7) [qx qy qz qw tx ty tz]; R3: (n,3); SO3: (n,4) [x y z w]."""

    def __init__(self, kind, dt_a=1.0, t0_a=0.0, knots_a=None, dt_b=1.0, t0_b=0.0, knots_b=None, compat_zero_dB=False, locked=False):
        self.kind = kind
        self.dt_a, self.t0_a, self.dt_b, self.t0_b = dt_a, t0_a, dt_b, t0_b
        self.knots_a = None if knots_a is None else _f64(knots_a)
        self.knots_b = None if knots_b is None else _f64(knots_b)
        self.compat_zero_dB, self.locked = compat_zero_dB, locked

    @property
    def size_a(self):
        return 7 if self.kind == SE3 else 3

    def c(self):
        t = _Traj()
        t.kind = self.kind
        t.dt_a, t.t0_a = self.dt_a, self.t0_a
        t.n_a = 0 if self.knots_a is None else len(self.knots_a)
        t.knots_a = None if self.knots_a is None else self.knots_a.ctypes.data
        t.dt_b, t.t0_b = self.dt_b, self.t0_b
        t.n_b = 0 if self.knots_b is None else len(self.knots_b)
        t.knots_b = None if self.knots_b is None else self.knots_b.ctypes.data
        t.compat_zero_dB, t.locked = int(self.compat_zero_dB), int(self.locked)
        return t

    @property
    def min_time(self):
        if self.kind == SPLIT:
            return max(self.t0_a, self.t0_b)
        return self.t0_b if self.kind == SO3 else self.t0_a

    @property
    def max_time(self):
        ma = None if self.knots_a is None else self.t0_a + (len(self.knots_a) - 3) * self.dt_a
        mb = None if self.knots_b is None else self.t0_b + (len(self.knots_b) - 3) * self.dt_b
        if self.kind == SPLIT:
            return min(ma, mb)
        return mb if self.kind == SO3 else ma


class Sensor:
    def __init__(self, q_ct=(0, 0, 0, 1), p_ct=(0, 0, 0), time_offset=0.0, max_time_offset=0.1, q_locked=True, p_locked=True,
                 d_locked=True, abias=None, gbias=None, abias_locked=True, gbias_locked=True):
        self.q_ct, self.p_ct = np.asarray(q_ct, float), np.asarray(p_ct, float)   # q_ct is (x,y,z,w)
        self.time_offset, self.max_time_offset = time_offset, max_time_offset
        self.q_locked, self.p_locked, self.d_locked = q_locked, p_locked, d_locked
        self.has_bias = abias is not None
        self.abias = np.zeros(3) if abias is None else np.asarray(abias, float)
        self.gbias = np.zeros(3) if gbias is None else np.asarray(gbias, float)
        self.abias_locked, self.gbias_locked = abias_locked, gbias_locked

    def c(self):
        s = _Sensor()
        s.q_ct[:] = list(self.q_ct)
        s.p_ct[:] = list(self.p_ct)
        s.time_offset, s.max_time_offset = self.time_offset, self.max_time_offset
        s.q_locked, s.p_locked, s.d_locked = int(self.q_locked), int(self.p_locked), int(self.d_locked)
        s.has_bias = int(self.has_bias)
        s.abias[:] = list(self.abias)
        s.gbias[:] = list(self.gbias)
        s.abias_locked, s.gbias_locked = int(self.abias_locked), int(self.gbias_locked)
        return s


class Camera(Sensor):
    def __init__(self, rows, cols, readout, K=None, wc=None, gamma=None, method="static", **kw):
        """PinholeCamera, or AtanCamera when wc / gamma are given; method "static" | "newton" picks the measurement class."""
        super().__init__(**kw)
        self.rows, self.cols, self.readout = rows, cols, readout
        self.K = np.eye(3) if K is None else np.asarray(K, float)
        self.wc, self.gamma, self.method = wc, gamma, method

    def cmeta(self):
        m = _Camera()
        m.readout, m.rows, m.cols = self.readout, self.rows, self.cols
        m.K[:] = list(self.K.reshape(-1))
        m.model = 0 if self.gamma is None else 1
        m.method = {"static": 0, "newton": 1}[self.method]
        if self.gamma is not None:
            m.wc[:] = [float(self.wc[0]), float(self.wc[1])]
            m.gamma = float(self.gamma)
        return m


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def _check(code, raise_on_error):
    if code != 0 and raise_on_error:
        msg = lib().kto_last_error().decode()
        raise OracleError(code, msg)


def traj_evaluate(traj, t, flags, raise_on_error=True):
    t = _f64(np.atleast_1d(t))
    n = len(t)
    out = dict(position=np.zeros((n, 3)), velocity=np.zeros((n, 3)), acceleration=np.zeros((n, 3)), orientation=np.zeros((n, 4)),
               angular_velocity=np.zeros((n, 3)), status=np.zeros(n, np.int32))
    tc = traj.c()
    code = lib().kto_traj_evaluate(C.byref(tc), n, _p(t), int(flags), _p(out["position"]), _p(out["velocity"]), _p(out["acceleration"]),
                                   _p(out["orientation"]), _p(out["angular_velocity"]), _p(out["status"]))
    _check(code, raise_on_error)
    return out


def se3_evaluate_matrices(traj, t, raise_on_error=True):
    t = _f64(np.atleast_1d(t))
    n = len(t)
    P, Pp, Pb = np.zeros((n, 4, 4)), np.zeros((n, 4, 4)), np.zeros((n, 4, 4))
    st = np.zeros(n, np.int32)
    tc = traj.c()
    code = lib().kto_se3_evaluate_matrices(C.byref(tc), n, _p(t), _p(P), _p(Pp), _p(Pb), _p(st))
    _check(code, raise_on_error)
    return P, Pp, Pb


def imu_residuals(traj, imu, which, t, y, weight=None, jac_mode=2, nthreads=0, cap=None, raise_on_error=True):
    """which: 0 gyro / 1 accel / 2 position (PositionMeasurement; `imu` is inert) / 3 orientation (OrientationMeasurement: y = q (x,y,z,w),
    ONE residual per row).  Returns dict(r, ids_a, Ja, ids_b, Jb, Js, i0_a, i0_b, status, eval_seconds)."""
    nres = 1 if which == 3 else 3
    t, y = _f64(t), _f64(y).reshape(-1, 4 if which == 3 else 3)
    n = len(t)
    weight = np.ones(n) if weight is None else _f64(weight)
    if cap is None:
        cap = 4 if imu.d_locked else 4 + int(np.ceil(2 * imu.max_time_offset / min(traj.dt_a, traj.dt_b))) + 2
    has_a, has_b = traj.kind != SO3, traj.kind in (SPLIT, SO3)
    sa = traj.size_a
    r = np.zeros((n, nres))
    ids_a = np.full((n, cap), -1, np.int32) if has_a else None
    Ja = np.zeros((n, cap, nres, sa)) if (has_a and jac_mode) else None
    ids_b = np.full((n, cap), -1, np.int32) if has_b else None
    Jb = np.zeros((n, cap, nres, 4)) if (has_b and jac_mode) else None
    Js = np.zeros((n, 42)) if jac_mode else None
    i0_a, i0_b, st = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
    secs = C.c_double(0)
    tc, sc = traj.c(), imu.c()
    code = lib().kto_imu_residuals(C.byref(tc), C.byref(sc), int(which), n, _p(t), _p(y), _p(weight), int(jac_mode), int(nthreads), _p(r),
                                   cap, _p(ids_a), _p(Ja), cap, _p(ids_b), _p(Jb), _p(Js), _p(i0_a), _p(i0_b), _p(st), C.byref(secs))
    _check(code, raise_on_error)
    return dict(r=r, ids_a=ids_a, Ja=Ja, ids_b=ids_b, Jb=Jb, Js=Js, i0_a=i0_a, i0_b=i0_b, status=st, eval_seconds=secs.value)


def static_rs_residuals(traj, cam, obs_uv, obs_t0, ref_uv, ref_t0, lm_idx, rho, weight=None, lm_locked=None, jac_mode=2, nthreads=0,
                        cap=None, raise_on_error=True):
    obs_uv, ref_uv = _f64(obs_uv).reshape(-1, 2), _f64(ref_uv).reshape(-1, 2)
    obs_t0, ref_t0, rho = _f64(obs_t0), _f64(ref_t0), _f64(rho)
    lm_idx = np.ascontiguousarray(lm_idx, np.int32)
    n = len(obs_t0)
    weight = np.ones(n) if weight is None else _f64(weight)
    lml = None if lm_locked is None else np.ascontiguousarray(lm_locked, np.int8)
    if cap is None:
        span = cam.readout + 2e-3 + (0 if cam.d_locked else 2 * cam.max_time_offset)
        cap = 2 * (4 + int(np.ceil(span / min(traj.dt_a, traj.dt_b))) + 1)
    has_a, has_b = traj.kind != SO3, traj.kind in (SPLIT, SO3)
    sa = traj.size_a
    r = np.zeros((n, 2))
    ids_a = np.full((n, cap), -1, np.int32) if has_a else None
    Ja = np.zeros((n, cap, 2, sa)) if (has_a and jac_mode) else None
    ids_b = np.full((n, cap), -1, np.int32) if has_b else None
    Jb = np.zeros((n, cap, 2, 4)) if (has_b and jac_mode) else None
    Js = np.zeros((n, 16)) if jac_mode else None
    Jrho = np.zeros((n, 2)) if jac_mode else None
    i0 = [np.zeros(n, np.int32) for _ in range(4)]
    st = np.zeros(n, np.int32)
    secs = C.c_double(0)
    tc, sc, cm = traj.c(), cam.c(), cam.cmeta()
    code = lib().kto_static_rs_residuals(C.byref(tc), C.byref(sc), C.byref(cm), n, _p(obs_uv), _p(obs_t0), _p(ref_uv), _p(ref_t0), _p(lm_idx),
                                         _p(rho), _p(lml), _p(weight), int(jac_mode), int(nthreads), _p(r), cap, _p(ids_a), _p(Ja), cap,
                                         _p(ids_b), _p(Jb), _p(Js), _p(Jrho), _p(i0[0]), _p(i0[1]), _p(i0[2]), _p(i0[3]), _p(st), C.byref(secs))
    _check(code, raise_on_error)
    return dict(r=r, ids_a=ids_a, Ja=Ja, ids_b=ids_b, Jb=Jb, Js=Js, Jrho=Jrho, i0_ref_a=i0[0], i0_obs_a=i0[1], i0_ref_b=i0[2], i0_obs_b=i0[3],
                status=st, eval_seconds=secs.value)


def lifting_rs_residuals(traj, cam, obs_uv, obs_t0, ref_uv, ref_t0, lm_idx, rho, vt=None, weight=None, jac_mode=2, nthreads=0, cap=None,
                         raise_on_error=True):
    """LiftingRsCameraMeasurement rows (lifting_rscamera_measurement.h): 3 residuals; vt = current frame-normalised row times (default: the
    initial value obs_uv.y / rows, :68).  Returns dict(r (n,3), ids_a, Ja (n,cap,3,size), ids_b, Jb, Jvt (n,3), Jrho (n,3), i0_*...)."""
    obs_uv, ref_uv = _f64(obs_uv).reshape(-1, 2), _f64(ref_uv).reshape(-1, 2)
    obs_t0, ref_t0, rho = _f64(obs_t0), _f64(ref_t0), _f64(rho)
    lm_idx = np.ascontiguousarray(lm_idx, np.int32)
    n = len(obs_t0)
    vt = obs_uv[:, 1] / float(cam.rows) if vt is None else _f64(vt)
    vt = np.ascontiguousarray(vt, np.float64)
    weight = np.ones(n) if weight is None else _f64(weight)
    if cap is None:
        span = cam.readout + 2e-3 + (0 if cam.d_locked else 2 * cam.max_time_offset)
        cap = 2 * (4 + int(np.ceil(span / min(traj.dt_a, traj.dt_b))) + 1)
    has_a, has_b = traj.kind != SO3, traj.kind in (SPLIT, SO3)
    sa = traj.size_a
    r = np.zeros((n, 3))
    ids_a = np.full((n, cap), -1, np.int32) if has_a else None
    Ja = np.zeros((n, cap, 3, sa)) if (has_a and jac_mode) else None
    ids_b = np.full((n, cap), -1, np.int32) if has_b else None
    Jb = np.zeros((n, cap, 3, 4)) if (has_b and jac_mode) else None
    Jvt = np.zeros((n, 3)) if jac_mode else None
    Jrho = np.zeros((n, 3)) if jac_mode else None
    Js = np.zeros((n, 24)) if jac_mode else None      # sensor blocks: q_ct (3x4) | p_ct (3x3) | time offset (3)
    i0 = [np.zeros(n, np.int32) for _ in range(4)]
    st = np.zeros(n, np.int32)
    secs = C.c_double(0)
    tc, sc, cm = traj.c(), cam.c(), cam.cmeta()
    code = lib().kto_lifting_rs_residuals(C.byref(tc), C.byref(sc), C.byref(cm), n, _p(obs_uv), _p(obs_t0), _p(ref_uv), _p(ref_t0), _p(lm_idx),
                                          _p(rho), _p(vt), _p(weight), int(jac_mode), int(nthreads), _p(r), cap, _p(ids_a), _p(Ja), cap,
                                          _p(ids_b), _p(Jb), _p(Jvt), _p(Jrho), _p(i0[0]), _p(i0[1]), _p(i0[2]), _p(i0[3]), _p(st), C.byref(secs), _p(Js))
    _check(code, raise_on_error)
    return dict(r=r, ids_a=ids_a, Ja=Ja, ids_b=ids_b, Jb=Jb, Jvt=Jvt, Jrho=Jrho, Js=Js, vt=vt, i0_ref_a=i0[0], i0_obs_a=i0[1], i0_ref_b=i0[2], i0_obs_b=i0[3],
                status=st, eval_seconds=secs.value)


def huber_correct(a, r, J=None):
    """ceres::HuberLoss(a) + Corrector on one residual block; returns (rho, r_corrected, J_corrected)."""
    r = _f64(r).copy()
    J2 = None if J is None else _f64(J).copy()
    rho = lib().kto_huber_correct(C.c_double(a), len(r), 0 if J2 is None else J2.shape[1], _p(r), _p(J2))
    return rho, r, J2


def se3_plus(T, delta):
    T, delta = _f64(T), _f64(delta)
    out = np.zeros(7)
    lib().kto_se3_plus(_p(T), _p(delta), _p(out))
    return out


def spline_structure(dt, t0, spans, cap=64):
    spans = _f64(spans).reshape(-1, 2)
    ids = np.zeros(cap, np.int32)
    nids, nseg = C.c_int(0), C.c_int(0)
    seg_t0, seg_n = np.zeros(len(spans)), np.zeros(len(spans), np.int32)
    code = lib().kto_spline_structure(C.c_double(dt), C.c_double(t0), len(spans), _p(spans), cap, _p(ids), C.byref(nids), _p(seg_t0), _p(seg_n),
                                      C.byref(nseg))
    _check(code, True)
    return ids[:nids.value].copy(), seg_t0[:nseg.value].copy(), seg_n[:nseg.value].copy()


def num_threads():
    return lib().kto_num_threads()


def camera_project(cam, X, dX=None):
    """CameraView::EvaluateProjection(X, dX, derive=True): returns (y, dy)."""
    X = _f64(X); dX = np.zeros(3) if dX is None else _f64(dX)
    y, dy = np.zeros(2), np.zeros(2)
    cm = cam.cmeta()
    lib().kto_camera_project(C.byref(cm), _p(X), _p(dX), _p(y), _p(dy))
    return y, dy


def camera_unproject(cam, y):
    y = _f64(y); X = np.zeros(3)
    cm = cam.cmeta()
    lib().kto_camera_unproject(C.byref(cm), _p(y), _p(X))
    return X


_alib = None


def analytic_se3_evaluate(dt, t0, knots, gyro=None, accel=None, cam=None, nthreads=0):
    """The OPTIMISED CPU variant (oracle/analytic_cpu.cpp): the product's closed-form mathematics compiled for the host and
    looped over the rows with OpenMP.  gyro / accel: dicts t, y, weight; cam: dict K, readout, rows, obs_uv, obs_t0, ref_uv, ref_t0, lm_idx,
    rho, weight, huber_c (None = no loss), optional wc / gamma (AtanCamera).  Returns dict(seconds, bad, gyro=(r, J), accel=(r, J), cam=(r, J))."""
    global _alib
    if _alib is None:
        build()
        _alib = C.CDLL(_ANALYTIC_PATH)
        _alib.kta_se3_evaluate.restype = C.c_double
    knots = _f64(knots)
    out, keep = {}, []

    def imu_args(m):
        if m is None:
            return [0, None, None, None], (None, None)
        t, y = _f64(m["t"]), _f64(m["y"]).reshape(-1, 3)
        w = np.ones(len(t)) if m.get("weight") is None else _f64(m["weight"])
        keep.extend([t, y, w])
        return [len(t), _p(t), _p(y), _p(w)], (np.zeros((len(t), 3)), np.zeros((len(t), 84)))
    ga, (gr, gJ) = imu_args(gyro)
    aa, (ar, aJ) = imu_args(accel)
    if cam is None:
        ca, cr, cJ = [0, None, C.c_double(0.0), 0, 0, None, C.c_double(0.0)] + [None] * 8, None, None
    else:
        n = len(cam["obs_t0"])
        arrs = [_f64(cam["obs_uv"]).reshape(-1, 2), _f64(cam["obs_t0"]), _f64(cam["ref_uv"]).reshape(-1, 2), _f64(cam["ref_t0"]),
                np.ascontiguousarray(cam["lm_idx"], np.int32), _f64(cam["rho"]), np.ones(n) if cam.get("weight") is None else _f64(cam["weight"]),
                None if cam.get("huber_c") is None else _f64(cam["huber_c"])]
        K, wc = _f64(cam["K"]).reshape(-1), _f64(cam.get("wc", (0.0, 0.0)))
        keep.extend(arrs + [K, wc])
        ca = [n, _p(K), C.c_double(cam["readout"]), int(cam["rows"]), 0 if cam.get("gamma") is None else 1, _p(wc), C.c_double(cam.get("gamma") or 0.0)] + [_p(a) for a in arrs]
        cr, cJ = np.zeros((n, 2)), np.zeros((n, 114))
    bad = np.zeros(1, np.int32)
    secs = _alib.kta_se3_evaluate(C.c_double(t0), C.c_double(dt), len(knots), _p(knots), *ga, *aa, *ca, int(nthreads),
                                  _p(gr), _p(gJ), _p(ar), _p(aJ), _p(cr), _p(cJ), _p(bad))
    return dict(seconds=float(secs), bad=int(bad[0]), gyro=(gr, gJ), accel=(ar, aJ), cam=(cr, cJ))


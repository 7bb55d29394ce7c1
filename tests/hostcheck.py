"""Builds and binds tests/host_check.cpp: the product's __host__ __device__ mathematics compiled for the host.

TEST HARNESS ONLY -- lets the CPU-only suite (-m "not gpu") compare the exact formulas the CUDA kernels run with
the oracle.  The product package never loads this library.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "host_check.cpp")
_CSRC = os.path.join(os.path.dirname(_HERE), "kontiki_b200", "csrc")
_OUT = os.path.join(_HERE, "_build", "libhostcheck.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        deps = [_SRC] + [os.path.join(_CSRC, f) for f in ("spline_math.cuh", "split_math.cuh", "sensor_jac.cuh", "lie_math.cuh", "dualnum.cuh", "newton_math.cuh")]
        if not os.path.exists(_OUT) or any(os.path.getmtime(d) > os.path.getmtime(_OUT) for d in deps):
            os.makedirs(os.path.dirname(_OUT), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", _SRC, "-o", _OUT])
        _lib = C.CDLL(_OUT)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def prepass(knots7):
    knots7 = _f(knots7)
    n = len(knots7)
    k8, pairs = np.zeros((n, 8)), np.zeros((n, 104))
    lib().hc_prepass(_p(knots7), n, _p(k8), _p(pairs))
    return k8, pairs


def imu(which, knots7, dt, t0, t, y, w=None, compat=False, time_offset=0.0, max_time_offset=0.1, locked=True):
    k8, pairs = prepass(knots7)
    t, y = _f(t), _f(y).reshape(-1, 4 if which == 3 else 3)
    n = len(t)
    w = np.ones(n) if w is None else _f(w)
    r, J = (np.zeros((n, 1)), np.zeros((n, 4, 1, 7))) if which == 3 else (np.zeros((n, 3)), np.zeros((n, 4, 3, 7)))
    i0, st = np.zeros(n, np.int32), np.zeros(n, np.int32)
    lib().hc_imu(int(which), C.c_double(t0), C.c_double(dt), len(k8), int(compat), C.c_double(time_offset), C.c_double(max_time_offset),
                 int(locked), _p(k8), _p(pairs), n, _p(t), _p(y), _p(w), _p(r), _p(J), _p(i0), _p(st))
    return dict(r=r, J=J, i0=i0, status=st)


def kinv_cofactor(K):
    """pinhole_camera.h:63-67 inverts K per call (Eigen fixed 3x3 inverse = cofactors / determinant)."""
    a = np.asarray(K, float).reshape(3, 3)
    c = np.empty((3, 3))
    c[0, 0] = a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1]; c[0, 1] = a[0, 2] * a[2, 1] - a[0, 1] * a[2, 2]; c[0, 2] = a[0, 1] * a[1, 2] - a[0, 2] * a[1, 1]
    c[1, 0] = a[1, 2] * a[2, 0] - a[1, 0] * a[2, 2]; c[1, 1] = a[0, 0] * a[2, 2] - a[0, 2] * a[2, 0]; c[1, 2] = a[0, 2] * a[1, 0] - a[0, 0] * a[1, 2]
    c[2, 0] = a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0]; c[2, 1] = a[0, 1] * a[2, 0] - a[0, 0] * a[2, 1]; c[2, 2] = a[0, 0] * a[1, 1] - a[0, 1] * a[1, 0]
    det = a[0, 0] * c[0, 0] + a[0, 1] * c[1, 0] + a[0, 2] * c[2, 0]
    return c / det


def _set_camera_model(cam):
    """PinholeCamera unless the oracle-style camera carries AtanCamera parameters (wc, gamma)."""
    lib().hc_set_camera_model.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
    if getattr(cam, "gamma", None) is None:
        lib().hc_set_camera_model(0, 0.0, 0.0, 0.0)
    else:
        lib().hc_set_camera_model(1, float(cam.wc[0]), float(cam.wc[1]), float(cam.gamma))


def newton_window(t0, dt, readout, obs_t0):
    lib().hc_newton_window.argtypes = [C.c_double] * 4
    return max(int(lib().hc_newton_window(t0, dt, readout, float(t))) for t in np.atleast_1d(obs_t0))


def newton_rs(knots7, dt, t0, cam, obs_uv, obs_t0, ref_uv, ref_t0, lm_idx, rho, w=None, huber_c=None, fast=False):
    """NewtonRsCameraMeasurement rows in the packed layout [ref 4x(2x7) | obs W x(2x7) | rho 2]; W = widest observation span."""
    _set_camera_model(cam)
    k8, pairs = prepass(knots7)
    obs_uv, ref_uv = _f(obs_uv).reshape(-1, 2), _f(ref_uv).reshape(-1, 2)
    obs_t0, ref_t0, rho = _f(obs_t0), _f(ref_t0), _f(rho)
    lm_idx = np.ascontiguousarray(lm_idx, np.int32)
    n = len(obs_t0)
    w = np.ones(n) if w is None else _f(w)
    hc = None if huber_c is None else _f(huber_c)
    K = _f(cam.K).reshape(-1)
    Kinv = _f(kinv_cofactor(cam.K)).reshape(-1)
    W = newton_window(t0, dt, cam.readout, obs_t0)
    lib().hc_set_newton_fast(int(fast))      # fast: the rows as k_newton_rs_fast + k_newton_rs produce them (closed form for one / two evaluations)
    r, J = np.zeros((n, 2)), np.zeros((n, 58 + 14 * W))
    ir, kb, it, st = (np.zeros(n, np.int32) for _ in range(4))
    lib().hc_newton_rs(C.c_double(t0), C.c_double(dt), len(k8), _p(K), _p(Kinv), _p(_f(cam.q_ct)), _p(_f(cam.p_ct)), C.c_double(cam.time_offset),
                       C.c_double(cam.max_time_offset), int(cam.d_locked), C.c_double(cam.readout), int(cam.rows), _p(k8), _p(pairs), n,
                       _p(obs_uv), _p(obs_t0), _p(ref_uv), _p(ref_t0), _p(lm_idx), _p(rho), _p(w), _p(hc), int(W), _p(r), _p(J), _p(ir), _p(kb),
                       _p(it), _p(st))
    return dict(r=r, J=J, i0_ref=ir, i0_obs=kb, W=W, iterations=it, status=st)


def lifting_rs(knots7, dt, t0, cam, obs_uv, obs_t0, ref_uv, ref_t0, lm_idx, rho, vt=None, w=None, huber_c=None, analytic=True):
    """LiftingRsCameraMeasurement rows; J (n, 90 + 21 W) packed [ref 4x(3x7) | obs W x(3x7) | vt 3 | rho 3], kbase = first knot of the span.
    analytic=True: the closed-form rows (what k_lifting_rs runs); False: the forward-mode directions (k_lifting_rs_fwd)."""
    _set_camera_model(cam)
    lib().hc_set_lifting_analytic(int(analytic))
    k8, pairs = prepass(knots7)
    obs_uv, ref_uv = _f(obs_uv).reshape(-1, 2), _f(ref_uv).reshape(-1, 2)
    obs_t0, ref_t0, rho = _f(obs_t0), _f(ref_t0), _f(rho)
    lm_idx = np.ascontiguousarray(lm_idx, np.int32)
    n = len(obs_t0)
    vt = obs_uv[:, 1] / float(cam.rows) if vt is None else _f(vt)
    vt = np.ascontiguousarray(vt, np.float64)
    w = np.ones(n) if w is None else _f(w)
    hc = None if huber_c is None else _f(huber_c)
    K = _f(cam.K).reshape(-1)
    Kinv = _f(kinv_cofactor(cam.K)).reshape(-1)
    W = newton_window(t0, dt, cam.readout, obs_t0)
    r, J = np.zeros((n, 3)), np.zeros((n, 90 + 21 * W))
    ir, kb, st = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
    lib().hc_lifting_rs(C.c_double(t0), C.c_double(dt), len(k8), _p(K), _p(Kinv), _p(_f(cam.q_ct)), _p(_f(cam.p_ct)), C.c_double(cam.time_offset),
                        C.c_double(cam.max_time_offset), int(cam.d_locked), C.c_double(cam.readout), int(cam.rows), _p(k8), _p(pairs), n,
                        _p(obs_uv), _p(obs_t0), _p(ref_uv), _p(ref_t0), _p(lm_idx), _p(rho), _p(vt), _p(w), _p(hc), int(W), _p(r), _p(J), _p(ir), _p(kb),
                        _p(st))
    return dict(r=r, J=J, i0_ref=ir, i0_obs=kb, W=W, status=st, vt=vt)


def span_sensor(knots7, dt, t0, cam, obs_uv, obs_t0, ref_uv, ref_t0, lm_idx, rho, lifting=False, vt=None, w=None, huber_c=None):
    """Sensor-block columns of NewtonRs (Js (n, 16)) / LiftingRs (Js (n, 24)) rows: [q_ct (nres x 4) | p_ct (nres x 3) | time offset (nres)]."""
    _set_camera_model(cam)
    k8, pairs = prepass(knots7)
    obs_uv, ref_uv = _f(obs_uv).reshape(-1, 2), _f(ref_uv).reshape(-1, 2)
    obs_t0, ref_t0, rho = _f(obs_t0), _f(ref_t0), _f(rho)
    lm_idx = np.ascontiguousarray(lm_idx, np.int32)
    n = len(obs_t0)
    vt = np.ascontiguousarray(obs_uv[:, 1] / float(cam.rows) if vt is None else _f(vt), np.float64)
    w = np.ones(n) if w is None else _f(w)
    hc = None if huber_c is None else _f(huber_c)
    K = _f(cam.K).reshape(-1)
    Kinv = _f(kinv_cofactor(cam.K)).reshape(-1)
    W = newton_window(t0, dt, cam.readout, obs_t0)
    Js = np.zeros((n, 24 if lifting else 16))
    st = np.zeros(n, np.int32)
    lib().hc_span_sensor(int(bool(lifting)), C.c_double(t0), C.c_double(dt), len(k8), _p(K), _p(Kinv), _p(_f(cam.q_ct)), _p(_f(cam.p_ct)),
                         C.c_double(cam.time_offset), C.c_double(cam.max_time_offset), int(cam.d_locked), C.c_double(cam.readout), int(cam.rows),
                         _p(k8), _p(pairs), n, _p(obs_uv), _p(obs_t0), _p(ref_uv), _p(ref_t0), _p(lm_idx), _p(rho), _p(vt), _p(w), _p(hc), int(W),
                         _p(Js), _p(st))
    return dict(Js=Js, status=st, W=W)


def static_rs(knots7, dt, t0, cam, obs_uv, obs_t0, ref_uv, ref_t0, lm_idx, rho, w=None, huber_c=None):
    """cam: oracle.kto.Camera-like (K, q_ct, p_ct, time_offset, max_time_offset, d_locked, readout, rows)."""
    _set_camera_model(cam)
    k8, pairs = prepass(knots7)
    obs_uv, ref_uv = _f(obs_uv).reshape(-1, 2), _f(ref_uv).reshape(-1, 2)
    obs_t0, ref_t0, rho = _f(obs_t0), _f(ref_t0), _f(rho)
    lm_idx = np.ascontiguousarray(lm_idx, np.int32)
    n = len(obs_t0)
    w = np.ones(n) if w is None else _f(w)
    hc = None if huber_c is None else _f(huber_c)
    K = _f(cam.K).reshape(-1)
    Kinv = _f(kinv_cofactor(cam.K)).reshape(-1)
    r, J = np.zeros((n, 2)), np.zeros((n, 114))
    ir, io, st = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
    lib().hc_static_rs(C.c_double(t0), C.c_double(dt), len(k8), _p(K), _p(Kinv), _p(_f(cam.q_ct)), _p(_f(cam.p_ct)), C.c_double(cam.time_offset),
                       C.c_double(cam.max_time_offset), int(cam.d_locked), C.c_double(cam.readout), int(cam.rows), _p(k8), _p(pairs), n,
                       _p(obs_uv), _p(obs_t0), _p(ref_uv), _p(ref_t0), _p(lm_idx), _p(rho), _p(w), _p(hc), _p(r), _p(J), _p(ir), _p(io), _p(st))
    return dict(r=r, J=J, i0_ref=ir, i0_obs=io, status=st)


# ---- split (R3 + SO3) trajectory ----------------------------------------------------------------------------------------
def split_prepass(vecs3, quats):
    vecs3, quats = _f(vecs3), _f(quats)
    v4, pairs = np.zeros((len(vecs3), 4)), np.zeros((len(quats), 28))
    st = C.c_int(0)
    lib().hc_split_prepass(_p(vecs3), len(vecs3), _p(quats), len(quats), _p(v4), _p(pairs), C.byref(st))
    return v4, quats, pairs, st.value


def imu_split(which, vecs3, dt_r3, t0_r3, quats, dt_so3, t0_so3, t, y, w=None, time_offset=0.0, max_time_offset=0.1, locked=True):
    v4, q4, pairs, st0 = split_prepass(vecs3, quats)
    t, y = _f(t), _f(y).reshape(-1, 4 if which == 3 else 3)
    n = len(t)
    w = np.ones(n) if w is None else _f(w)
    r = np.zeros((n, 1 if which == 3 else 3))
    J = np.zeros((n, {0: 48, 1: 84, 2: 36, 3: 16}[which]))
    ia, ib, st = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
    lib().hc_imu_split(int(which), C.c_double(t0_r3), C.c_double(dt_r3), len(v4), C.c_double(t0_so3), C.c_double(dt_so3), len(q4), C.c_double(time_offset),
                       C.c_double(max_time_offset), int(locked), _p(v4), _p(q4), _p(pairs), n, _p(t), _p(y), _p(w), _p(r), _p(J), _p(ia), _p(ib), _p(st))
    return dict(r=r, J=J, i0_r3=ia, i0_so3=ib, status=st, prepass_status=st0)


def static_rs_split(vecs3, dt_r3, t0_r3, quats, dt_so3, t0_so3, cam, obs_uv, obs_t0, ref_uv, ref_t0, lm_idx, rho, w=None, huber_c=None):
    _set_camera_model(cam)
    v4, q4, pairs, st0 = split_prepass(vecs3, quats)
    obs_uv, ref_uv = _f(obs_uv).reshape(-1, 2), _f(ref_uv).reshape(-1, 2)
    obs_t0, ref_t0, rho = _f(obs_t0), _f(ref_t0), _f(rho)
    lm_idx = np.ascontiguousarray(lm_idx, np.int32)
    n = len(obs_t0)
    w = np.ones(n) if w is None else _f(w)
    hc = None if huber_c is None else _f(huber_c)
    K = _f(cam.K).reshape(-1)
    Kinv = _f(kinv_cofactor(cam.K)).reshape(-1)
    r, J = np.zeros((n, 2)), np.zeros((n, 114))
    idx, st = np.zeros((n, 4), np.int32), np.zeros(n, np.int32)
    lib().hc_static_rs_split(C.c_double(t0_r3), C.c_double(dt_r3), len(v4), C.c_double(t0_so3), C.c_double(dt_so3), len(q4), _p(K), _p(Kinv),
                             _p(_f(cam.q_ct)), _p(_f(cam.p_ct)), C.c_double(cam.time_offset), C.c_double(cam.max_time_offset), int(cam.d_locked),
                             C.c_double(cam.readout), int(cam.rows), _p(v4), _p(q4), _p(pairs), n, _p(obs_uv), _p(obs_t0), _p(ref_uv), _p(ref_t0),
                             _p(lm_idx), _p(rho), _p(w), _p(hc), _p(r), _p(J), _p(idx), _p(st))
    return dict(r=r, J=J, idx=idx, status=st)


def span_rs_split(vecs3, dt_r3, t0_r3, quats, dt_so3, t0_so3, cam, obs_uv, obs_t0, ref_uv, ref_t0, lm_idx, rho, lifting=False, vt=None, w=None, huber_c=None, sensor=False):
    """NewtonRs / LiftingRs rows on a split trajectory, packed [ref R3 4x(nres x 3) | ref SO3 4x(nres x 4) | obs R3 Wa x(..) | obs SO3 Wb x(..) | (vt) | rho];
    idx (n, 4) = ref R3 first knot, obs R3 span base, ref SO3 first knot, obs SO3 span base."""
    _set_camera_model(cam)
    v4, q4, pairs, st0 = split_prepass(vecs3, quats)
    obs_uv, ref_uv = _f(obs_uv).reshape(-1, 2), _f(ref_uv).reshape(-1, 2)
    obs_t0, ref_t0, rho = _f(obs_t0), _f(ref_t0), _f(rho)
    lm_idx = np.ascontiguousarray(lm_idx, np.int32)
    n = len(obs_t0)
    vt = np.ascontiguousarray(obs_uv[:, 1] / float(cam.rows) if vt is None else _f(vt), np.float64)
    w = np.ones(n) if w is None else _f(w)
    hc = None if huber_c is None else _f(huber_c)
    K = _f(cam.K).reshape(-1)
    Kinv = _f(kinv_cofactor(cam.K)).reshape(-1)
    lib().hc_span_window.argtypes = [C.c_double] * 4
    Wa = max(int(lib().hc_span_window(t0_r3, dt_r3, cam.readout, float(t))) for t in obs_t0)
    Wb = max(int(lib().hc_span_window(t0_so3, dt_so3, cam.readout, float(t))) for t in obs_t0)
    nres = 3 if lifting else 2
    row = int(lib().hc_span_split_row_len(int(bool(lifting)), Wa, Wb))
    r, J = np.zeros((n, nres)), np.zeros((n, row))
    idx, st = np.zeros((n, 4), np.int32), np.zeros(n, np.int32)
    lib().hc_span_rs_split(int(bool(lifting)), C.c_double(t0_r3), C.c_double(dt_r3), len(v4), C.c_double(t0_so3), C.c_double(dt_so3), len(q4), _p(K), _p(Kinv),
                           _p(_f(cam.q_ct)), _p(_f(cam.p_ct)), C.c_double(cam.time_offset), C.c_double(cam.max_time_offset), int(cam.d_locked),
                           C.c_double(cam.readout), int(cam.rows), _p(v4), _p(q4), _p(pairs), n, _p(obs_uv), _p(obs_t0), _p(ref_uv), _p(ref_t0),
                           _p(lm_idx), _p(rho), _p(vt), _p(w), _p(hc), Wa, Wb, _p(r), _p(J), _p(idx), _p(st))
    Js = None
    if sensor:
        Js, st2 = np.zeros((n, 8 * nres)), np.zeros(n, np.int32)
        lib().hc_span_sensor_split(int(bool(lifting)), C.c_double(t0_r3), C.c_double(dt_r3), len(v4), C.c_double(t0_so3), C.c_double(dt_so3), len(q4), _p(K), _p(Kinv),
                                   _p(_f(cam.q_ct)), _p(_f(cam.p_ct)), C.c_double(cam.time_offset), C.c_double(cam.max_time_offset), int(cam.d_locked),
                                   C.c_double(cam.readout), int(cam.rows), _p(v4), _p(q4), _p(pairs), n, _p(obs_uv), _p(obs_t0), _p(ref_uv), _p(ref_t0),
                                   _p(lm_idx), _p(rho), _p(vt), _p(w), _p(hc), Wa, Wb, _p(Js), _p(st2))
        assert (st2 == st).all()
    return dict(r=r, J=J, idx=idx, status=st, Wa=Wa, Wb=Wb, vt=vt, Js=Js)


def traj_eval_se3(knots7, dt, t0, t, compat=False):
    k8, pairs = prepass(knots7)
    t = _f(np.atleast_1d(t))
    out, st = np.zeros((len(t), 16)), np.zeros(len(t), np.int32)
    lib().hc_traj_eval_se3(C.c_double(t0), C.c_double(dt), len(k8), int(compat), _p(k8), _p(pairs), len(t), _p(t), _p(out), _p(st))
    return out, st


def traj_eval_split(vecs3, dt_r3, t0_r3, quats, dt_so3, t0_so3, t):
    v4, q4, pairs, _ = split_prepass(vecs3, quats)
    t = _f(np.atleast_1d(t))
    out, st = np.zeros((len(t), 16)), np.zeros(len(t), np.int32)
    lib().hc_traj_eval_split(C.c_double(t0_r3), C.c_double(dt_r3), len(v4), C.c_double(t0_so3), C.c_double(dt_so3), len(q4), _p(v4), _p(q4), _p(pairs),
                             len(t), _p(t), _p(out), _p(st))
    return out, st


# ---- sensor-block Jacobians ---------------------------------------------------------------------------------------------
def imu_time_offset_se3(which, knots7, dt, t0, t, w=None, compat=False, time_offset=0.0, max_time_offset=0.1, locked=True):
    k8, pairs = prepass(knots7)
    t = _f(t)
    n = len(t)
    w = np.ones(n) if w is None else _f(w)
    out, st = np.zeros((n, 3)), np.zeros(n, np.int32)
    lib().hc_imu_time_offset_se3(int(which), C.c_double(t0), C.c_double(dt), len(k8), int(compat), C.c_double(time_offset), C.c_double(max_time_offset),
                                 int(locked), _p(k8), _p(pairs), n, _p(t), _p(w), _p(out), _p(st))
    return out, st


def imu_time_offset_split(which, vecs3, dt_r3, t0_r3, quats, dt_so3, t0_so3, t, w=None, time_offset=0.0, max_time_offset=0.1, locked=True):
    v4, q4, pairs, _ = split_prepass(vecs3, quats)
    t = _f(t)
    n = len(t)
    w = np.ones(n) if w is None else _f(w)
    out, st = np.zeros((n, 3)), np.zeros(n, np.int32)
    lib().hc_imu_time_offset_split(int(which), C.c_double(t0_r3), C.c_double(dt_r3), len(v4), C.c_double(t0_so3), C.c_double(dt_so3), len(q4),
                                   C.c_double(time_offset), C.c_double(max_time_offset), int(locked), _p(v4), _p(q4), _p(pairs), n, _p(t), _p(w), _p(out), _p(st))
    return out, st


def static_rs_sensor_se3(knots7, dt, t0, cam, obs_uv, obs_t0, ref_uv, ref_t0, lm_idx, rho, w=None, huber_c=None):
    _set_camera_model(cam)
    k8, pairs = prepass(knots7)
    obs_uv, ref_uv = _f(obs_uv).reshape(-1, 2), _f(ref_uv).reshape(-1, 2)
    obs_t0, ref_t0, rho = _f(obs_t0), _f(ref_t0), _f(rho)
    lm_idx = np.ascontiguousarray(lm_idx, np.int32)
    n = len(obs_t0)
    w = np.ones(n) if w is None else _f(w)
    hc = None if huber_c is None else _f(huber_c)
    K = _f(cam.K).reshape(-1)
    Kinv = _f(kinv_cofactor(cam.K)).reshape(-1)
    out, st = np.zeros((n, 16)), np.zeros(n, np.int32)
    lib().hc_static_rs_sensor_se3(C.c_double(t0), C.c_double(dt), len(k8), _p(K), _p(Kinv), _p(_f(cam.q_ct)), _p(_f(cam.p_ct)), C.c_double(cam.time_offset),
                                  C.c_double(cam.max_time_offset), int(cam.d_locked), C.c_double(cam.readout), int(cam.rows), _p(k8), _p(pairs), n,
                                  _p(obs_uv), _p(obs_t0), _p(ref_uv), _p(ref_t0), _p(lm_idx), _p(rho), _p(w), _p(hc), _p(out), _p(st))
    return out, st


def static_rs_sensor_split(vecs3, dt_r3, t0_r3, quats, dt_so3, t0_so3, cam, obs_uv, obs_t0, ref_uv, ref_t0, lm_idx, rho, w=None, huber_c=None):
    _set_camera_model(cam)
    v4, q4, pairs, st0 = split_prepass(vecs3, quats)
    obs_uv, ref_uv = _f(obs_uv).reshape(-1, 2), _f(ref_uv).reshape(-1, 2)
    obs_t0, ref_t0, rho = _f(obs_t0), _f(ref_t0), _f(rho)
    lm_idx = np.ascontiguousarray(lm_idx, np.int32)
    n = len(obs_t0)
    w = np.ones(n) if w is None else _f(w)
    hc = None if huber_c is None else _f(huber_c)
    K = _f(cam.K).reshape(-1)
    Kinv = _f(kinv_cofactor(cam.K)).reshape(-1)
    out, st = np.zeros((n, 16)), np.zeros(n, np.int32)
    lib().hc_static_rs_sensor_split(C.c_double(t0_r3), C.c_double(dt_r3), len(v4), C.c_double(t0_so3), C.c_double(dt_so3), len(q4), _p(K), _p(Kinv), _p(_f(cam.q_ct)),
                                    _p(_f(cam.p_ct)), C.c_double(cam.time_offset), C.c_double(cam.max_time_offset), int(cam.d_locked), C.c_double(cam.readout),
                                    int(cam.rows), _p(v4), _p(q4), _p(pairs), n, _p(obs_uv), _p(obs_t0), _p(ref_uv), _p(ref_t0), _p(lm_idx), _p(rho), _p(w), _p(hc),
                                    _p(out), _p(st))
    return out, st


def se3_matrices(knots7, dt, t0, t):
    k8, pairs = prepass(knots7)
    t = _f(np.atleast_1d(t))
    out, st = np.zeros((len(t), 3, 4, 4)), np.zeros(len(t), np.int32)
    lib().hc_se3_matrices(C.c_double(t0), C.c_double(dt), len(k8), _p(k8), _p(pairs), len(t), _p(t), _p(out), _p(st))
    return out, st

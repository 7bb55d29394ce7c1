"""Regenerates tests/golden/mp_v1.npz: golden input/output vectors from the INDEPENDENT 60-digit transcription of the reference
(tests/mp_reference.py -- mpmath, Jacobians by central differences in 60-digit arithmetic; shares no code with oracle/ or the kernels).

Unlike oracle_v*.npz (frozen outputs of the oracle) these vectors do not come from anything the product or the oracle is built on, so
tests/test_mp_golden.py checks all three against them: the oracle, the product's mathematics compiled for the host, and (-m gpu) the CUDA
path through the C ABI.  Inputs: the reference's own fixture knots (python/tests/conftest.py:32-45, :52-67, :83-105) and seeded random
trajectories.  Takes about two minutes.

    python tests/golden/make_mp_golden.py        # rewrites mp_v1.npz; commit the result
"""
import os
import sys

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import fixtures_ref as fx  # noqa: E402
import mp_reference as mr  # noqa: E402
from test_oracle_independent import random_se3_knots, random_split  # noqa: E402


def fl(x):
    return np.array([float(v) for v in x])


def imu_rows(mt, which, t, y, w, wins):
    """r (n,3); dense Jacobians per spline: list of (n, n_knots, 3, width), non-zero only inside the active windows."""
    n = len(t)
    r = np.zeros((n, 3))
    Js = [np.zeros((n, len(arr), 3, len(arr[0]))) for arr in wins]
    i0s = np.zeros((n, len(wins)), np.int32)
    for k in range(n):
        fun = lambda: mr.imu_residual(mt, which, float(t[k]), y[k], float(w[k]))
        r[k] = fl(fun())
        e = mt.evaluate(mp.mpf(float(t[k])))
        for a, arr in enumerate(wins):
            i0 = e["i0"] if (a == 0) else e["i0_so3"]
            i0s[k, a] = i0
            width = len(arr[0])
            J = np.array(mr.jacobian(fun, [(arr, i0 + b, c) for b in range(4) for c in range(width)])).reshape(3, 4, width).transpose(1, 0, 2)
            Js[a][k, i0:i0 + 4] = J
    return r, Js, i0s


def camera_rows(mt, cam, obs_uv, obs_t0, ref_uv, ref_t0, rho, arrs, span_knots):
    n = len(rho)
    r = np.zeros((n, 2))
    Js = [np.zeros((n, len(arr), 2, len(arr[0]))) for arr in arrs]
    Jrho = np.zeros((n, 2))
    for k in range(n):
        box = [[mp.mpf(float(rho[k]))]]
        fun = lambda: mr.static_rs_residual(mt, cam, obs_uv[k], float(obs_t0[k]), ref_uv[k], float(ref_t0[k]), box[0][0])[0]
        rr, ir, io = mr.static_rs_residual(mt, cam, obs_uv[k], float(obs_t0[k]), ref_uv[k], float(ref_t0[k]), box[0][0])
        r[k] = fl(rr)
        Jrho[k] = np.array(mr.jacobian(fun, [(box, 0, 0)])).reshape(2)
        ids = sorted(set(range(ir, ir + 4)) | set(range(io, io + 4)))
        for a, arr in enumerate(arrs):
            width = len(arr[0])
            J = np.array(mr.jacobian(fun, [(arr, b, c) for b in ids for c in range(width)])).reshape(2, len(ids), width).transpose(1, 0, 2)
            Js[a][k, ids] = J
    return r, Js, Jrho


def build():
    out = {}
    rng = np.random.default_rng(20261018)
    # ---- IMU rows on SE3: the reference's fixture and a random walk
    for name, (knots, dt, t0) in (("fix", (fx.SE3_KNOTS, fx.SE3_DT, fx.SE3_T0)), ("rand", (random_se3_knots(9, 7), 0.37, -0.4))):
        n = 6
        t = np.sort(t0 + rng.uniform(0.02, len(knots) - 3.02, n) * dt)
        y, w = rng.uniform(-1, 1, (n, 3)), rng.uniform(0.5, 2, n)
        out[f"se3_{name}_knots"], out[f"se3_{name}_meta"] = np.asarray(knots, float), np.array([dt, t0])
        out[f"se3_{name}_t"], out[f"se3_{name}_y"], out[f"se3_{name}_w"] = t, y, w
        for which, tag, compat in ((0, "gyro", False), (1, "accel", False), (1, "accel_compat", True)):
            mt = mr.Trajectory("se3", dt, t0, knots=knots, compat_zero_dB=compat)
            r, Js, i0s = imu_rows(mt, which, t, y, w, [mt.knots])
            out[f"se3_{name}_{tag}_r"], out[f"se3_{name}_{tag}_J"], out[f"se3_{name}_{tag}_i0"] = r, Js[0], i0s[:, 0]
    # ---- IMU rows on the split fixture (R3 and SO3 fixtures side by side, different dt) and a random split trajectory
    for name, (r3, dta, t0a, so3, dtb, t0b) in (("fix", (fx.R3_KNOTS, fx.R3_DT, fx.R3_T0, fx.SO3_KNOTS, fx.SO3_DT, fx.SO3_T0)),
                                                ("rand", (*random_split(9, 21)[:1], 0.41, 0.3, random_split(9, 21)[1], 0.41, 0.3))):
        lo, hi = max(t0a, t0b), min(t0a + (len(r3) - 3) * dta, t0b + (len(so3) - 3) * dtb)
        n = 5
        t = np.sort(rng.uniform(lo + 1e-3, hi - 1e-3, n))
        y, w = rng.uniform(-1, 1, (n, 3)), rng.uniform(0.5, 2, n)
        out[f"split_{name}_r3"], out[f"split_{name}_so3"], out[f"split_{name}_meta"] = np.asarray(r3, float), np.asarray(so3, float), np.array([dta, t0a, dtb, t0b])
        out[f"split_{name}_t"], out[f"split_{name}_y"], out[f"split_{name}_w"] = t, y, w
        for which, tag in ((0, "gyro"), (1, "accel")):
            mt = SplitMp(r3, dta, t0a, so3, dtb, t0b)
            r, Js, i0s = imu_rows(mt, which, t, y, w, [mt.r3, mt.so3])
            out[f"split_{name}_{tag}_r"], out[f"split_{name}_{tag}_Ja"], out[f"split_{name}_{tag}_Jb"] = r, Js[0], Js[1]
            out[f"split_{name}_{tag}_i0a"], out[f"split_{name}_{tag}_i0b"] = i0s[:, 0], i0s[:, 1]
    # ---- static-RS rows, SE3 (relative pose + time offset set) and split
    K = np.array([[900., 0, 960], [0, 900., 540], [0, 0, 1]])
    q_ct = fx.so3_exp_xyzw(np.array([0.1, -0.2, 0.05]))
    p_ct, d = np.array([0.05, -0.02, 0.1]), 0.004
    knots, dt, t0 = random_se3_knots(14, 5, step=0.2), 0.05, 0.0
    n = 6
    lo, hi = t0 + 0.3 * dt, t0 + (len(knots) - 3.3) * dt - 0.03
    ref_t0 = rng.uniform(lo, hi, n)
    obs_t0 = np.clip(ref_t0 + rng.uniform(-2.5, 2.5, n) * dt, lo, hi)
    ref_uv, obs_uv = rng.uniform([100, 100], [1800, 1000], (n, 2)), rng.uniform([100, 100], [1800, 1000], (n, 2))
    rho = rng.uniform(0.05, 0.6, n)
    cam = dict(K=K, rows=1080, readout=0.026, q_ct=q_ct, p_ct=p_ct, time_offset=d)
    out["cam_K"], out["cam_meta"], out["cam_q_ct"], out["cam_p_ct"] = K, np.array([1080, 1920, 0.026, d]), q_ct, p_ct
    for k, v in (("obs_uv", obs_uv), ("obs_t0", obs_t0), ("ref_uv", ref_uv), ("ref_t0", ref_t0), ("rho", rho)):
        out["cam_" + k] = v
    mt = mr.Trajectory("se3", dt, t0, knots=knots)
    r, Js, Jrho = camera_rows(mt, cam, obs_uv, obs_t0, ref_uv, ref_t0, rho, [mt.knots], None)
    out["cam_se3_knots"], out["cam_se3_meta"], out["cam_se3_r"], out["cam_se3_J"], out["cam_se3_Jrho"] = knots, np.array([dt, t0]), r, Js[0], Jrho
    cam0 = dict(K=K, rows=1080, readout=0.026)
    mt = mr.Trajectory("split", dt, t0, r3=knots[:, 4:], so3=knots[:, :4])
    r, Js, Jrho = camera_rows(mt, cam0, obs_uv, obs_t0, ref_uv, ref_t0, rho, [mt.r3, mt.so3], None)
    out["cam_split_r"], out["cam_split_Ja"], out["cam_split_Jb"], out["cam_split_Jrho"] = r, Js[0], Js[1], Jrho
    return out


class SplitMp(mr.Trajectory):
    """Split trajectory whose two splines have their own (dt, t0), like the reference's split fixture (conftest.py:107-113)."""

    def __init__(self, r3, dta, t0a, so3, dtb, t0b):
        super().__init__("split", dta, t0a, r3=r3, so3=so3)
        self.dtb, self.t0b = mp.mpf(dtb), mp.mpf(t0b)

    def evaluate(self, t, acc=False):
        a, b = mr.r3_spline(self.r3, self.t0, self.dt, t), mr.so3_spline(self.so3, self.t0b, self.dtb, t)
        return dict(position=a["position"], velocity=a["velocity"], acceleration=a["acceleration"], orientation=b["orientation"],
                    angular_velocity=b["angular_velocity"], i0=a["i0"], i0_so3=b["i0"])


if __name__ == "__main__":
    path = os.path.join(HERE, "mp_v1.npz")
    np.savez_compressed(path, **build())
    print("wrote", path, os.path.getsize(path), "bytes")

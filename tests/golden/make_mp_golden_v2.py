"""Regenerates tests/golden/mp_v2.npz: golden vectors for the WIDENING rows from the independent 60-digit
transcription of the reference (tests/mp_reference.py): NewtonRsCameraMeasurement rows -- value of the Newton iteration on the row time and the
derivative THROUGH the iteration (central differences of the whole loop at 60 digits) -- with a PinholeCamera and an AtanCamera, and
LiftingRsCameraMeasurement rows (three residuals, row-time column).  Relative pose and time offset of the camera set.  mp_v1.npz (the hot path)
is not touched.  tests/test_mp_golden_v2.py checks the oracle, the product's mathematics compiled for the host (forward mode AND the closed
form the kernels run) and (-m gpu) the CUDA path against them.  Takes about two minutes.

    python tests/golden/make_mp_golden_v2.py        # rewrites mp_v2.npz; commit the result
"""
import os
import sys

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import mp_reference as mr  # noqa: E402
from test_oracle_independent import random_se3_knots, span_camera_case  # noqa: E402


def fl(x):
    return np.array([float(v) for v in x])


def span_rows(mt, cam, kind, obs_uv, obs_t0, ref_uv, ref_t0, rho, vt, dt, t0):
    """r (n, nres); dense J (n, n_knots, nres, 7): every knot the reference's residual block lists (the 4 reference knots and the knots of the
    observation span {t0_obs - 1e-3, t0_obs + readout + 1e-3}); Jrho (n, nres); Jvt (n, 3) for lifting rows; evaluations of the iteration."""
    n, nk, nres = len(rho), len(mt.knots), 2 if kind == "newton" else 3
    r, J, Jrho, Jvt, ev = np.zeros((n, nres)), np.zeros((n, nk, nres, 7)), np.zeros((n, nres)), np.zeros((n, 3)), np.zeros(n, np.int32)
    for k in range(n):
        box = [[mp.mpf(float(rho[k])), mp.mpf(float(vt[k]))]]
        if kind == "newton":
            full = lambda: mr.newton_rs_residual(mt, cam, obs_uv[k], float(obs_t0[k]), ref_uv[k], float(ref_t0[k]), box[0][0])
        else:
            full = lambda: mr.lifting_rs_residual(mt, cam, obs_uv[k], float(obs_t0[k]), ref_uv[k], float(ref_t0[k]), box[0][0], box[0][1])
        fun = lambda: full()[0]
        rr, ir, third = full()
        r[k] = fl(rr)
        ev[k] = third if kind == "newton" else 1
        t0o = float(obs_t0[k]) + cam["time_offset"]
        first = int(np.floor((t0o - 1e-3 - t0) / dt))
        last = int(np.floor((t0o + cam["readout"] + 1e-3 - t0) / dt)) + 3
        ids = sorted(set(range(ir, ir + 4)) | set(range(first, min(last, nk - 1) + 1)))
        J[k, ids] = np.array(mr.jacobian(fun, [(mt.knots, b, c) for b in ids for c in range(7)])).reshape(nres, len(ids), 7).transpose(1, 0, 2)
        rest = [(mt.knots, b, c) for b in range(nk) if b not in ids for c in (1, 4)]
        assert not rest or np.abs(np.array(mr.jacobian(fun, rest))).max() == 0.0
        Jrho[k] = np.array(mr.jacobian(fun, [(box, 0, 0)])).reshape(nres)
        if kind == "lifting":
            Jvt[k] = np.array(mr.jacobian(fun, [(box, 0, 1)])).reshape(3)
    return r, J, Jrho, Jvt, ev


def build():
    out = {}
    knots, dt, t0 = random_se3_knots(16, 9, step=0.2), 0.05, 0.0
    out["knots"], out["meta"] = knots, np.array([dt, t0])
    mt = mr.Trajectory("se3", dt, t0, knots=knots)
    for tag, atan in (("pin", False), ("atan", True)):
        cam, obs_uv, obs_t0, ref_uv, ref_t0, rho = span_camera_case(knots, dt, t0, 31 + atan, atan, 0.1, n=5)
        # observed rows from (nearly) exact to far off: the iteration stops after one, two and more evaluations
        obs_uv[:, 1] = np.clip(obs_uv[:, 1] + np.array([0.0, 6.0, -40.0, 250.0, -400.0]), 2.0, cam["rows"] - 3.0)
        for _ in range(2):      # row 0 observed where the landmark projects (to 0.05 px): the first Newton step is below half a row, ONE evaluation
            r0 = mr.newton_rs_residual(mt, cam, obs_uv[0], float(obs_t0[0]), ref_uv[0], float(ref_t0[0]), mp.mpf(float(rho[0])))[0]
            obs_uv[0] = obs_uv[0] - fl(r0)
        obs_uv[0] += np.array([0.05, -0.05])
        vt = np.clip(obs_uv[:, 1] / cam["rows"] + np.array([0.06, -0.04, 0.1, 0.0, -0.08]), 0.0, 1.0)
        if not atan:
            out["cam_K"], out["cam_meta"] = cam["K"], np.array([cam["rows"], 1920, cam["readout"], cam["time_offset"]])
            out["cam_q_ct"], out["cam_p_ct"] = np.array(cam["q_ct"]), np.array(cam["p_ct"])
        else:
            out["cam_wc"], out["cam_gamma"] = np.array(cam["wc"]), np.array(cam["gamma"])
        for key, v in (("obs_uv", obs_uv), ("obs_t0", obs_t0), ("ref_uv", ref_uv), ("ref_t0", ref_t0), ("rho", rho), ("vt", vt)):
            out[f"{tag}_{key}"] = v
        for kind in ("newton", "lifting"):
            r, J, Jrho, Jvt, ev = span_rows(mt, cam, kind, obs_uv, obs_t0, ref_uv, ref_t0, rho, vt, dt, t0)
            out[f"{tag}_{kind}_r"], out[f"{tag}_{kind}_J"], out[f"{tag}_{kind}_Jrho"] = r, J, Jrho
            if kind == "lifting":
                out[f"{tag}_{kind}_Jvt"] = Jvt
            else:
                out[f"{tag}_{kind}_evaluations"] = ev
            print(tag, kind, "rows", len(rho), "evaluations", ev)
    return out


if __name__ == "__main__":
    path = os.path.join(HERE, "mp_v2.npz")
    np.savez_compressed(path, **build())
    print("wrote", path, os.path.getsize(path), "bytes")

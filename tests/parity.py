"""Shared helpers of the parity tests: oracle calls in the packed layout of the C ABI and the error metric."""
import numpy as np

from oracle import kto

TOL = 1e-9   # BASELINE.json north_star: "within 1e-9 relative on fp64 residuals/Jacobians"


def rel_err(a, b):
    """max |a-b| / max|b| per measurement (block-norm relative error, SURVEY.md section 8d 'parity gate')."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    n = a.shape[0]
    d = np.abs(a - b).reshape(n, -1).max(1)
    s = np.maximum(np.abs(b).reshape(n, -1).max(1), 1e-300)
    return (d / s).max() if n else 0.0


def oracle_imu(traj, which, t, y, w=None, imu=None):
    """Oracle residuals + Jacobian in the packed layout [4 knots][3][7] at i0..i0+3 (locked IMU => exactly those blocks)."""
    imu = imu or kto.Sensor()
    res = kto.imu_residuals(traj, imu, which, t, y, w, jac_mode=2)
    assert (res["ids_a"][:, 0] == res["i0_a"]).all() and (res["ids_a"][:, 3] == res["i0_a"] + 3).all()
    return dict(r=res["r"], J=res["Ja"][:, :4], i0=res["i0_a"])


def scatter_cam(Jp, i0_ref, i0_obs, ids):
    """Packed camera row [ref 4x(2x7) | obs 4x(2x7) | rho 2] -> the reference's structural blocks (n, cap, 2, 7) for `ids`."""
    n, cap = ids.shape
    out = np.zeros((n, cap, 2, 7))
    Jp = np.asarray(Jp).reshape(n, 114)
    for i in range(n):
        pos = {int(k): j for j, k in enumerate(ids[i]) if k >= 0}
        for base, off in ((i0_ref[i], 0), (i0_obs[i], 56)):
            for k in range(4):
                out[i, pos[int(base) + k]] += Jp[i, off + 14 * k: off + 14 * (k + 1)].reshape(2, 7)
    return out, Jp[:, 112:114]

"""Shared helpers of the parity tests: oracle calls in the packed layout of the C ABI and the error metric."""
import numpy as np

from oracle import kto

TOL = 1e-9   # BASELINE.json north_star: "within 1e-9 relative on fp64 residuals/Jacobians"
# Camera residuals r = weight (uv - y_hat) are differences of pixel coordinates of size ~1e3, so the comparison is ABSOLUTE in pixels:
# 1e-9 px = 5e-13 of the pixel scale (gate what is achieved, ~1e-10 px; round 1 gated 1e-6 px).
CAM_R_TOL = 1e-9


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


def scatter_cam(Jp, i0_ref, i0_obs, ids, W=4):
    """Packed camera row [ref 4x(2x7) | obs W x(2x7) | rho 2] -> the reference's structural blocks (n, cap, 2, 7) for `ids`.
    W = 4 for static-RS rows; Newton-RS rows carry the whole observation span (blocks outside `ids` must be zero)."""
    n, cap = ids.shape
    out = np.zeros((n, cap, 2, 7))
    row = 58 + 14 * W
    Jp = np.asarray(Jp).reshape(n, row)
    for i in range(n):
        pos = {int(k): j for j, k in enumerate(ids[i]) if k >= 0}
        for base, off, nk in ((i0_ref[i], 0, 4), (i0_obs[i], 56, W)):
            for k in range(nk):
                blk = Jp[i, off + 14 * k: off + 14 * (k + 1)].reshape(2, 7)
                if int(base) + k in pos:
                    out[i, pos[int(base) + k]] += blk
                else:
                    assert not blk.any(), "non-zero block outside the residual's structural knots"
    return out, Jp[:, row - 2:row]


def scatter_lifting(Jp, i0_ref, i0_obs, ids, W):
    """Packed lifting row [ref 4x(3x7) | obs W x(3x7) | vt 3 | rho 3] -> (structural blocks (n, cap, 3, 7), Jvt (n, 3), Jrho (n, 3))."""
    n, cap = ids.shape
    out = np.zeros((n, cap, 3, 7))
    row = 90 + 21 * W
    Jp = np.asarray(Jp).reshape(n, row)
    for i in range(n):
        pos = {int(k): j for j, k in enumerate(ids[i]) if k >= 0}
        for base, off, nk in ((i0_ref[i], 0, 4), (i0_obs[i], 84, W)):
            for k in range(nk):
                blk = Jp[i, off + 21 * k: off + 21 * (k + 1)].reshape(3, 7)
                if int(base) + k in pos:
                    out[i, pos[int(base) + k]] += blk
                else:
                    assert not blk.any(), "non-zero block outside the residual's structural knots"
    return out, Jp[:, row - 6:row - 3], Jp[:, row - 3:row]



def scatter_span_split(Jp, idx, ids_a, ids_b, Wa, Wb, nres):
    """Packed span row on a split trajectory [ref R3 4x(nres x 3) | ref SO3 4x(nres x 4) | obs R3 Wa x(..) | obs SO3 Wb x(..) | tail] ->
    (R3 blocks (n, cap, nres, 3), SO3 blocks (n, cap, nres, 4), tail (n, tail)); idx (n, 4) = ref R3, obs R3 base, ref SO3, obs SO3 base."""
    n, cap_a = ids_a.shape
    cap_b = ids_b.shape[1]
    Ja, Jb = np.zeros((n, cap_a, nres, 3)), np.zeros((n, cap_b, nres, 4))
    Jp = np.asarray(Jp).reshape(n, -1)
    o_ra, o_rb, o_oa, o_ob, o_t = 0, nres * 12, nres * 28, nres * (28 + 3 * Wa), nres * (28 + 3 * Wa + 4 * Wb)
    for i in range(n):
        pa = {int(k): j for j, k in enumerate(ids_a[i]) if k >= 0}
        pb = {int(k): j for j, k in enumerate(ids_b[i]) if k >= 0}
        for base, off, nk, w, pos, out in ((idx[i, 0], o_ra, 4, 3, pa, Ja), (idx[i, 1], o_oa, Wa, 3, pa, Ja), (idx[i, 2], o_rb, 4, 4, pb, Jb), (idx[i, 3], o_ob, Wb, 4, pb, Jb)):
            for k in range(nk):
                blk = Jp[i, off + nres * w * k: off + nres * w * (k + 1)].reshape(nres, w)
                if int(base) + k in pos:
                    out[i, pos[int(base) + k]] += blk
                else:
                    assert not blk.any(), "non-zero block outside the residual's structural knots"
    return Ja, Jb, Jp[:, o_t:]

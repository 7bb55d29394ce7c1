"""Synthetic problems of BASELINE.json's configs: seeded numpy generators, no files.

Plain numpy on the host; used by bench.py, __graft_entry__.smoke() and the tests to build *inputs*.  Nothing here is on
the evaluation path (the vectorised pose evaluation below only places landmarks so that reprojection residuals are
realistic -- pixel-level noise plus a few outliers -- and is independent of both the CUDA path and the oracle).
"""
import numpy as np

IMAGE_ROWS, IMAGE_COLS, CAMERA_READOUT = 1080, 1920, 0.026      # reference python/tests/fixtures/camera_fixtures.py:8-10
CAMERA_K = np.array([[900.0, 0.0, 960.0], [0.0, 900.0, 540.0], [0.0, 0.0, 1.0]])


def _hat(w):
    z = np.zeros(w.shape[:-1])
    return np.stack([np.stack([z, -w[..., 2], w[..., 1]], -1), np.stack([w[..., 2], z, -w[..., 0]], -1), np.stack([-w[..., 1], w[..., 0], z], -1)], -2)


def so3_exp_quat(w):
    """(n,3) rotation vectors -> (n,4) quaternions (x,y,z,w)."""
    w = np.atleast_2d(w)
    th = np.linalg.norm(w, axis=-1, keepdims=True)
    k = np.where(th > 1e-8, np.sin(th / 2) / np.maximum(th, 1e-300), 0.5 - th ** 2 / 48)
    return np.concatenate([k * w, np.cos(th / 2)], -1)


def quat_to_rot(q):
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    return np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
                     np.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
                     np.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)], -2)


def _so3_exp(w):
    th = np.linalg.norm(w, axis=-1)[..., None, None]
    W = _hat(w)
    a = np.where(th > 1e-6, np.sin(th) / np.maximum(th, 1e-300), 1 - th ** 2 / 6)
    b = np.where(th > 1e-6, (1 - np.cos(th)) / np.maximum(th, 1e-300) ** 2, 0.5 - th ** 2 / 24)
    c = np.where(th > 1e-6, (th - np.sin(th)) / np.maximum(th, 1e-300) ** 3, 1 / 6 - th ** 2 / 120)
    I = np.eye(3)
    return I + a * W + b * W @ W, I + b * W + c * W @ W


def _so3_log(R):
    c = np.clip((np.trace(R, axis1=-2, axis2=-1) - 1) / 2, -1, 1)
    th = np.arccos(c)
    v = np.stack([R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]], -1)
    k = np.where(th > 1e-6, th / (2 * np.maximum(np.sin(th), 1e-300)), 0.5 + th ** 2 / 12)
    return k[..., None] * v


def smooth_se3_knots(n, dt, seed=1001, noise=1e-3):
    """SE3 control points of a smooth motion + small noise, (n,7) [qx qy qz qw tx ty tz]."""
    rng = np.random.default_rng(seed)
    tau = np.arange(n) * dt
    p = np.stack([5 * np.sin(.31 * tau), 5 * np.cos(.17 * tau), 1 + .5 * np.sin(.53 * tau)], 1)
    w = np.stack([.6 * np.sin(.23 * tau), .4 * np.sin(.41 * tau + 1), .8 * np.sin(.13 * tau + 2)], 1)
    w = w + rng.normal(0, noise, w.shape)
    p = p + rng.normal(0, noise, p.shape)
    q = so3_exp_quat(w)
    for i in range(1, n):
        if np.dot(q[i - 1], q[i]) < 0:
            q[i] = -q[i]
    return np.concatenate([q, p], 1)


def pose_eval(knots, dt, t0, t):
    """Cumulative SE(3) B-spline pose (R (m,3,3), p (m,3)) at times t -- data generation only."""
    t = np.asarray(t, float)
    s = (t - t0) / dt
    i0 = np.floor(s).astype(int)
    u = s - i0
    B = np.stack([(5 + 3 * u - 3 * u ** 2 + u ** 3) / 6, (1 + 3 * u + 3 * u ** 2 - 2 * u ** 3) / 6, u ** 3 / 6], -1)
    R = quat_to_rot(knots[i0, :4])
    p = knots[i0, 4:7].copy()
    for j in range(3):
        Ra, Rb = quat_to_rot(knots[i0 + j, :4]), quat_to_rot(knots[i0 + j + 1, :4])
        ta, tb = knots[i0 + j, 4:7], knots[i0 + j + 1, 4:7]
        dR = np.swapaxes(Ra, -1, -2) @ Rb
        dtv = (np.swapaxes(Ra, -1, -2) @ (tb - ta)[..., None])[..., 0]
        phi = _so3_log(dR)
        _, V = _so3_exp(phi)
        ups = np.linalg.solve(V, dtv[..., None])[..., 0]
        E, Vb = _so3_exp(B[:, j, None] * phi)
        a = (Vb @ (B[:, j, None] * ups)[..., None])[..., 0]
        p = p + (R @ a[..., None])[..., 0]
        R = R @ E
    return R, p


def valid_time(n_knots, dt, t0=0.0):
    return t0, t0 + (n_knots - 3) * dt


def make_imu(n, n_knots, dt, t0=0.0, seed=1, accel=False):
    """n gyroscope (or accelerometer) samples at uniform random times in the valid span (reference tests: conftest.py:169-179)."""
    rng = np.random.default_rng(seed)
    lo, hi = valid_time(n_knots, dt, t0)
    t = rng.uniform(lo, hi - 1e-9 * (hi - lo), n)
    y = rng.uniform(-1, 1, (n, 3))
    if accel:
        y[:, 2] += 9.8
    return dict(t=t, y=y, weight=np.ones(n))


def make_static_rs(knots, dt, n_landmarks, obs_per_landmark=10, t0=0.0, seed=4, noise_px=0.5, fps=30.0,
                   rows=IMAGE_ROWS, cols=IMAGE_COLS, readout=CAMERA_READOUT, K=CAMERA_K):
    """Landmarks with a reference observation and `obs_per_landmark` rolling-shutter observations in the following views
    (reference generator: python/tests/fixtures/sfm_fixtures.py:34-84; here vectorised, 2 fixed-point iterations on the row)."""
    rng = np.random.default_rng(seed)
    n_knots = len(knots)
    lo, hi = valid_time(n_knots, dt, t0)
    margin = 2e-3
    n_views = int(np.floor((hi - lo - readout - 2 * margin) * fps))
    view_t0 = lo + margin + np.arange(n_views) / fps
    row_delta = readout / rows
    Kinv = np.linalg.inv(K)
    ref_view = np.empty(n_landmarks, int)
    ref_uv = np.empty((n_landmarks, 2))
    rho = np.empty(n_landmarks)
    obs_uv = np.empty((n_landmarks, obs_per_landmark, 2))
    todo = np.arange(n_landmarks)
    for _ in range(200):
        if len(todo) == 0:
            break
        m = len(todo)
        rv = rng.integers(0, n_views - obs_per_landmark, m)
        uv = np.stack([rng.uniform(0, cols, m), rng.uniform(0, rows, m)], 1)
        z = rng.uniform(0.5, 100.0, m)
        Xc = z[:, None] * (np.concatenate([uv, np.ones((m, 1))], 1) @ Kinv.T)
        R, p = pose_eval(knots, dt, t0, view_t0[rv] + uv[:, 1] * row_delta)
        Xw = (R @ Xc[..., None])[..., 0] + p
        ok = np.ones(m, bool)
        ouv = np.empty((m, obs_per_landmark, 2))
        for k in range(obs_per_landmark):
            tv = view_t0[rv + 1 + k]
            v = np.full(m, rows / 2.0)
            for _it in range(3):
                R, p = pose_eval(knots, dt, t0, tv + v * row_delta)
                Xo = (np.swapaxes(R, -1, -2) @ (Xw - p)[..., None])[..., 0]
                pr = Xo @ K.T
                zz = np.where(pr[:, 2] > 1e-6, pr[:, 2], 1.0)
                y = pr[:, :2] / zz[:, None]
                v = np.clip(y[:, 1], 0, rows - 1e-6)
            y = y + rng.normal(0, noise_px, y.shape)
            ok &= (pr[:, 2] > 1e-2) & (y[:, 0] >= 0) & (y[:, 0] < cols) & (y[:, 1] >= 0) & (y[:, 1] < rows)
            ouv[:, k] = y
        good = todo[ok]
        ref_view[good], ref_uv[good], rho[good], obs_uv[good] = rv[ok], uv[ok], 1.0 / z[ok], ouv[ok]
        todo = todo[~ok]
    if len(todo):
        raise RuntimeError("could not place all landmarks in view")
    lm_idx = np.repeat(np.arange(n_landmarks, dtype=np.int32), obs_per_landmark)
    obs_view = (ref_view[:, None] + 1 + np.arange(obs_per_landmark)[None, :]).reshape(-1)
    n = len(lm_idx)
    return dict(obs_uv=obs_uv.reshape(-1, 2), obs_t0=view_t0[obs_view], ref_uv=ref_uv[lm_idx], ref_t0=view_t0[ref_view][lm_idx], lm_idx=lm_idx,
                rho=rho, weight=np.ones(n), huber_c=np.full(n, 5.0), rows=rows, cols=cols, readout=readout, K=K)


CONFIGS = {
    # name: (n_knots, dt, n_gyro, n_accel, n_landmarks, obs_per_landmark)
    "C1": (200, 0.1, 5_000, 0, 0, 0),
    "C2": (2_000, 0.05, 100_000, 100_000, 0, 0),
    "C3": (5_000, 0.02, 0, 0, 50_000, 10),
    "C4": (5_000, 0.02, 100_000, 100_000, 50_000, 10),
    "H1": (5_000, 0.02, 50_000, 50_000, 50_000, 10),    # north-star headline: 100k IMU + 500k reprojection
    "C5": (10_000, 0.01, 250_000, 250_000, 25_000, 20),  # split R3 + SO3 trajectory, 1M mixed measurements
}
SPLIT_CONFIGS = {"C5"}


def make_config(name, scale=1.0):
    """Returns dict(knots, dt, t0, gyro, accel, cam) for one of BASELINE.json's SE3 configs (scale < 1 shrinks the counts)."""
    n_knots, dt, ng, na, nl, opl = CONFIGS[name]
    ng, na, nl = int(ng * scale), int(na * scale), int(nl * scale)
    knots = smooth_se3_knots(n_knots, dt)
    out = dict(name=name, knots=knots, dt=dt, t0=0.0, gyro=None, accel=None, cam=None, split=name in SPLIT_CONFIGS)
    if out["split"]:      # R3 knots from the positions, SO3 knots from the (sign-continuous) orientations of the same smooth motion
        out["r3"], out["so3"] = knots[:, 4:7].copy(), knots[:, 0:4].copy()
    seeds = {"C1": (1, 0), "C2": (2, 3)}.get(name, (5, 6))
    if ng:
        out["gyro"] = make_imu(ng, n_knots, dt, seed=seeds[0])
    if na:
        out["accel"] = make_imu(na, n_knots, dt, seed=seeds[1], accel=True)
    if nl:
        out["cam"] = make_static_rs(knots, dt, nl, opl)
    return out


def algorithmic_bytes(cfg):
    """SURVEY.md section 8d contract figures: SE3 740 B per IMU row, 1012 B per static-RS row; split 452 / 744 / 1020 B."""
    ng = len(cfg["gyro"]["t"]) if cfg["gyro"] else 0
    na = len(cfg["accel"]["t"]) if cfg["accel"] else 0
    n_cam = len(cfg["cam"]["lm_idx"]) if cfg["cam"] else 0
    if cfg.get("split"):
        return 452 * ng + 744 * na + 1020 * n_cam
    return 740 * (ng + na) + 1012 * n_cam


def num_measurements(cfg):
    return sum(len(cfg[k]["t"]) for k in ("gyro", "accel") if cfg[k]) + (len(cfg["cam"]["lm_idx"]) if cfg["cam"] else 0)

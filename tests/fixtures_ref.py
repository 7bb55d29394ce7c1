"""Fixed inputs of the reference's own test-suite, restated as plain arrays.

Sources (read-only reference, /root/reference/python/tests/):
  conftest.py:32-45   R3 control points, dt=2.3, t0=1.22
  conftest.py:52-67   SO3 constant-rate control points, dt=0.6, t0=1.22, 10 deg/s about (1,0,1)/sqrt2
  conftest.py:83-105  SE3 control points, dt=2.3, t0=1.22
  fixtures/camera_fixtures.py:8-16  camera constants
Quaternions here are stored (x, y, z, w) (Eigen coefficient order); the reference's Python API uses (w, x, y, z).
"""
import numpy as np

IMAGE_ROWS, IMAGE_COLS, CAMERA_READOUT = 1080, 1920, 0.026

R3_DT, R3_T0 = 2.3, 1.22
R3_KNOTS = np.array([[1, 1, 2], [1, 2, 1.4], [1, 4, 0], [-2, 2, 2], [-3, -2, 1], [-4, -2, 0], [-1, 2, 0], [-2, -1.5, 1.2]], float)

SO3_DT, SO3_T0 = 0.6, 1.22
SO3_RATE = np.deg2rad(10)
SO3_AXIS = np.array([1., 0, 1]) / np.sqrt(2)


def so3_knots_wxyz():
    N = int(np.ceil(5. / SO3_DT)) + 3
    times = SO3_T0 + np.arange(-3, N - 3) * SO3_DT
    out = []
    for t in times:
        theta = SO3_RATE * t
        q = np.empty(4)
        q[0] = np.cos(theta / 2)
        q[1:] = np.sin(theta / 2) * SO3_AXIS
        out.append(q)
    return np.array(out)


def wxyz_to_xyzw(q):
    q = np.asarray(q, float)
    return np.concatenate([q[..., 1:], q[..., :1]], axis=-1)


def xyzw_to_wxyz(q):
    q = np.asarray(q, float)
    return np.concatenate([q[..., 3:], q[..., :3]], axis=-1)


SO3_KNOTS = wxyz_to_xyzw(so3_knots_wxyz())

SE3_DT, SE3_T0 = 2.3, 1.22
_SE3_CP = [([1, 0, 2, 3], [1, 4, 6]), ([3, 1, 2, 3], [-1, 2, 3]), ([1, 0, 1, 3], [2, 3, 2]),
           ([2, 1, 4, 1], [1, 4, 7]), ([1, 0, 2, 3], [1, 4, 6]), ([1, 1, 3, 1], [2, -1, 2])]


def quat_to_rotation_matrix(q_wxyz):
    """python/kontiki/rotations.py:3-27 convention (w first)."""
    w, x, y, z = q_wxyz
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def se3_knots():
    """(n,7) [qx qy qz qw tx ty tz]; the reference builds the knot from a 4x4 matrix via Sophus::SE3d(Matrix4d)."""
    out = []
    for q, p in _SE3_CP:
        q = np.array(q, float) / np.linalg.norm(q)
        out.append(np.concatenate([wxyz_to_xyzw(q), np.array(p, float)]))
    return np.array(out)


SE3_KNOTS = se3_knots()


def qmul_xyzw(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def qconj_xyzw(q):
    return np.array([-q[0], -q[1], -q[2], q[3]])


def rot_xyzw(q):
    return quat_to_rotation_matrix(xyzw_to_wxyz(q))


def so3_exp_xyzw(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.array([0.5 * w[0], 0.5 * w[1], 0.5 * w[2], 1.0])
    s = np.sin(th / 2) / th
    return np.array([s * w[0], s * w[1], s * w[2], np.cos(th / 2)])


def smooth_se3_knots(n, dt, seed=1001, noise=1e-3):
    """SURVEY.md section 8d synthetic SE3 trajectory: smooth motion + small tangent noise."""
    rng = np.random.default_rng(seed)
    tau = np.arange(n) * dt
    p = np.stack([5 * np.sin(.31 * tau), 5 * np.cos(.17 * tau), 1 + .5 * np.sin(.53 * tau)], 1)
    w = np.stack([.6 * np.sin(.23 * tau), .4 * np.sin(.41 * tau + 1), .8 * np.sin(.13 * tau + 2)], 1)
    w = w + rng.normal(0, noise, w.shape)
    p = p + rng.normal(0, noise, p.shape)
    q = np.array([so3_exp_xyzw(wi) for wi in w])
    for i in range(1, n):            # sign continuity
        if np.dot(q[i - 1], q[i]) < 0:
            q[i] = -q[i]
    return np.concatenate([q, p], 1)


# fixtures/camera_fixtures.py:12-16 (AtanCamera fixture)
ATAN_K = np.array([[853.12703455, 0., 988.06311256],
                   [0., 873.54956631, 525.71056312],
                   [0., 0., 1.]])
ATAN_WC = np.array([0.0029110778971412417, 0.0004189670467132041])
ATAN_GAMMA = 0.8894355177968156

"""INDEPENDENT high-precision pin of the oracle ("cross-checked by ... mpmath at 50 digits").

TEST INFRASTRUCTURE ONLY.  This file is a transcription, in mpmath arithmetic (60 significant digits), of the reference's
headers for the hot path, written from /root/reference/cpplib/include/kontiki/** and from the published formulas of the
un-vendored dependencies (Sophus @00f3fd91 SE3/SO3 exp / log / inverse / operator*, Eigen quaternion product,
`q * v`, toRotationMatrix) -- NOT from oracle/*.hpp and not from kontiki_b200/csrc.  It shares no
code with either: group elements are handled the way Sophus stores them (quaternion + translation), and, as a second
route that shares nothing even with that, SE3 poses are also evaluated with 4x4 matrix exponentials / logarithms
(mp.expm / mp.logm) which know nothing about Rodrigues-type closed forms.

Jacobians here are central differences of the residual in 60-digit arithmetic (h = 1e-25: truncation ~1e-50, rounding
~1e-35), with respect to the 7 (SE3) / 4 (SO3) / 3 (R3) stored doubles of every knot and to rho -- i.e. exactly what
ceres::Jet differentiates in the reference, including the direction along a knot's quaternion (the reference formulas
use the raw mapped quaternion in P0.matrix() and in Eigen's q*v polynomial).

Reference sites transcribed (relative to cpplib/include/kontiki/):
  trajectories/spline_base.h:18-28, 148-152            basis matrices, index / interpolation amount
  trajectories/uniform_se3_spline_trajectory.h:81-194   SE3 spline, position / velocity / acceleration / orientation / omega
  trajectories/uniform_so3_spline_trajectory.h:46-125   SO3 spline;  math/quaternion_math.h:16-95  logq / expq / angular_velocity
  trajectories/uniform_r3_spline_trajectory.h:34-101    R3 spline;   trajectories/split_trajectory.h:41-58
  sensors/imu.h:47-59, constants.h:13,24                gyroscope / accelerometer model
  sensors/pinhole_camera.h:47-67, measurements/static_rscamera_measurement.h:21-55, 89-94
  measurements/gyroscope_measurement.h:36-38, accelerometer_measurement.h:37-39
Widening rows, same rules:
  sensors/atan_camera.h:54-103 (EvaluateProjection / Unproject), sensors/pinhole_camera.h:47-61 (dy)
  measurements/newton_rscamera_measurement.h:23-120 (reproject_newton: the Newton iteration on the row time, df from the hand-written first
    derivatives, math/quaternion_math.h:103-115 dq_from_angular_velocity / vector_sandwich), :150-155 (Error)
  measurements/lifting_rscamera_measurement.h:21-56 (reproject_lifting), :98-112 (Error, three residuals)
  measurements/position_measurement.h:22-31, orientation_measurement.h:24-32 (Eigen 3.3 angularDistance = 2 atan2(|vec(d)|, |d.w|), d = q conj(qhat))
"""
import mpmath as mp

mp.mp.dps = 60
ZERO, ONE, TWO, HALF = mp.mpf(0), mp.mpf(1), mp.mpf(2), mp.mpf(1) / 2
SOPHUS_EPS = mp.mpf("1e-10")          # Sophus::Constants<double>::epsilon()
GRAVITY = [ZERO, ZERO, mp.mpf("-9.80665")]   # constants.h:13,24

# spline_base.h:18-28 (row vector times matrix)
M_CUMUL = [[mp.mpf(v) / 6 for v in row] for row in ([6, 5, 1, 0], [0, 3, 3, 0], [0, -3, 3, 0], [0, 1, -2, 1])]
M_PLAIN = [[mp.mpf(v) / 6 for v in row] for row in ([1, 4, 1, 0], [-3, 0, 3, 0], [3, -6, 3, 0], [-1, 3, -3, 1])]


def mpv(x):
    return [mp.mpf(float(v)) if not isinstance(v, mp.mpf) else v for v in x]


# ---- Eigen quaternions, storage (x, y, z, w) ------------------------------------------------------------------------
def q_mul(a, b):      # Eigen::Quaternion operator* (Hamilton product, no normalisation)
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return [aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx,
            aw * bw - ax * bx - ay * by - az * bz]


def q_conj(q):
    return [-q[0], -q[1], -q[2], q[3]]


def cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def q_rot(q, v):      # Eigen `q * v` (_transformVector): v + w (2 u x v) + u x (2 u x v), a polynomial in q -- no normalisation
    u = q[:3]
    uv = [TWO * c for c in cross(u, v)]
    c2 = cross(u, uv)
    return [v[i] + q[3] * uv[i] + c2[i] for i in range(3)]


def q_to_R(q):        # Eigen toRotationMatrix (polynomial in q)
    x, y, z, w = q
    tx, ty, tz = TWO * x, TWO * y, TWO * z
    twx, twy, twz, txx, txy, txz, tyy, tyz, tzz = tx * w, ty * w, tz * w, tx * x, ty * x, tz * x, ty * y, tz * y, tz * z
    return mp.matrix([[ONE - (tyy + tzz), txy - twz, txz + twy], [txy + twz, ONE - (txx + tzz), tyz - twx], [txz - twy, tyz + twx, ONE - (txx + tyy)]])


def q_normalized(q):
    n = mp.sqrt(sum(c * c for c in q))
    return [c / n for c in q]


# ---- Sophus SO3 / SE3 as published -------------------------------------------------------------
# An SE3 is (q, t).  SO3(quaternion) normalises; SO3 * SO3 is the Hamilton product handed to that constructor.
def so3_exp(w):
    th2 = sum(c * c for c in w)
    th = mp.sqrt(th2)
    if th < SOPHUS_EPS:
        th4 = th2 * th2
        imag = HALF - th2 / 48 + th4 / 3840
        real = ONE - th2 / 8 + th4 / 384
    else:
        imag = mp.sin(th / 2) / th
        real = mp.cos(th / 2)
    return [imag * w[0], imag * w[1], imag * w[2], real], th


def so3_log(q):
    n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2]
    n = mp.sqrt(n2)
    w = q[3]
    if n < SOPHUS_EPS:
        f = TWO / w - TWO * n2 / (w * w * w)
    elif abs(w) < SOPHUS_EPS:
        f = (mp.pi if w > 0 else -mp.pi) / n
    else:
        f = TWO * mp.atan(n / w) / n
    return [f * q[0], f * q[1], f * q[2]]


def hat3(w):
    return mp.matrix([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])


def se3_exp(xi):      # xi = [upsilon; omega];  t = V upsilon, V = I + (1-cos th)/th^2 W + (th - sin th)/th^3 W^2
    ups, om = xi[:3], xi[3:]
    q, th = so3_exp(om)
    W = hat3(om)
    if th < SOPHUS_EPS:
        V = q_to_R(q)
    else:
        V = mp.eye(3) + (ONE - mp.cos(th)) / (th * th) * W + (th - mp.sin(th)) / (th ** 3) * (W * W)
    t = V * mp.matrix(ups)
    return q, [t[0], t[1], t[2]]


def se3_log(q, t):
    om = so3_log(q)
    th = mp.sqrt(sum(c * c for c in om))
    W = hat3(om)
    if abs(th) < SOPHUS_EPS:
        Vinv = mp.eye(3) - HALF * W + (ONE / 12) * (W * W)
    else:
        Vinv = mp.eye(3) - HALF * W + (ONE - th * mp.cos(th / 2) / (TWO * mp.sin(th / 2))) / (th * th) * (W * W)
    u = Vinv * mp.matrix(t)
    return [u[0], u[1], u[2], om[0], om[1], om[2]]


def se3_inverse(q, t):
    qi = q_normalized(q_conj(q))
    return qi, q_rot(qi, [-c for c in t])


def se3_mul(qa, ta, qb, tb):
    r = q_rot(qa, tb)
    return q_normalized(q_mul(qa, qb)), [ta[i] + r[i] for i in range(3)]


def se3_matrix(q, t):
    R = q_to_R(q)
    M = mp.eye(4)
    for i in range(3):
        for j in range(3):
            M[i, j] = R[i, j]
        M[i, 3] = t[i]
    return M


def se3_hat(xi):
    M = mp.zeros(4)
    W = hat3(xi[3:])
    for i in range(3):
        for j in range(3):
            M[i, j] = W[i, j]
        M[i, 3] = xi[i]
    return M


# ---- splines ------------------------------------------------------------------------------------------------------------
def index_and_u(t, t0, dt, n):
    """spline_base.h:148-152 on the whole spline; the per-residual segment view shifts the origin by an exact multiple of dt, which at 60
    digits is the same number.  (Bit-exact index arithmetic in doubles is the oracle's own business and is pinned elsewhere.)"""
    s = (t - t0) / dt
    i0 = int(mp.floor(s))
    if n < 4 or i0 < 0 or i0 > n - 4:
        raise ValueError("out of range")
    return i0, s - i0


def basis(u, dt, M):
    U = [ONE, u, u * u, u ** 3]
    dU = [ZERO, ONE / dt, TWO * u / dt, 3 * u * u / dt]
    d2U = [ZERO, ZERO, TWO / (dt * dt), 6 * u / (dt * dt)]
    mul = lambda r: [sum(r[k] * M[k][j] for k in range(4)) for j in range(4)]
    return mul(U), mul(dU), mul(d2U)


def se3_spline(knots, t0, dt, t, compat_zero_dB=False, want_acc=False):
    """uniform_se3_spline_trajectory.h:81-194.  knots: list of 7-lists [qx qy qz qw tx ty tz] (mpf).  Returns dict position, velocity,
    acceleration, orientation (x,y,z,w), angular_velocity.  compat_zero_dB reproduces the Jet path of the accelerometer flags (dB unassigned = 0)."""
    i0, u = index_and_u(t, t0, dt, len(knots))
    B, dB, d2B = basis(u, dt, M_CUMUL)
    if compat_zero_dB:
        dB = [ZERO] * 4
    Pq, Pt = list(knots[i0][:4]), list(knots[i0][4:])
    A, Ap, Ab = [], [], []
    for j in range(1, 4):
        a, b = knots[i0 + j - 1], knots[i0 + j]
        iq, it = se3_inverse(a[:4], a[4:])
        dq, dtt = se3_mul(iq, it, b[:4], b[4:])
        om = se3_log(dq, dtt)
        Oh = se3_hat(om)
        Aq, At = se3_exp([B[j] * c for c in om])
        Pq, Pt = se3_mul(Pq, Pt, Aq, At)
        Am = se3_matrix(Aq, At)
        A.append(Am)
        Ajp = Am * Oh * dB[j]
        Ap.append(Ajp)
        Ab.append(Ajp * Oh * dB[j] + Am * Oh * d2B[j])
    P0 = se3_matrix(knots[i0][:4], knots[i0][4:])          # raw mapped knot: rotationMatrix() is the polynomial in q
    M1 = Ap[0] * A[1] * A[2] + A[0] * Ap[1] * A[2] + A[0] * A[1] * Ap[2]
    Pp = P0 * M1
    out = dict(position=Pt, orientation=Pq, velocity=[Pp[0, 3], Pp[1, 3], Pp[2, 3]], i0=i0)
    Rt = q_to_R(Pq).T
    Wh = Pp[0:3, 0:3] * Rt
    out["angular_velocity"] = [HALF * (Wh[2, 1] - Wh[1, 2]), HALF * (Wh[0, 2] - Wh[2, 0]), HALF * (Wh[1, 0] - Wh[0, 1])]
    if want_acc:
        M2 = (Ab[0] * A[1] * A[2] + A[0] * Ab[1] * A[2] + A[0] * A[1] * Ab[2] + TWO * Ap[0] * Ap[1] * A[2] + TWO * Ap[0] * A[1] * Ap[2]
              + TWO * A[0] * Ap[1] * Ap[2])
        Pb = P0 * M2
        out["acceleration"] = [Pb[0, 3], Pb[1, 3], Pb[2, 3]]
    return out


def se3_pose_expm(knots, t0, dt, t):
    """The second, closed-form-free route: P(t) = P0 prod_j expm(B_j logm(P_{j-1}^-1 P_j)) with 4x4 matrix functions only (unit knots)."""
    i0, u = index_and_u(t, t0, dt, len(knots))
    B, _, _ = basis(u, dt, M_CUMUL)
    mats = [se3_matrix(q_normalized(k[:4]), k[4:]) for k in knots[i0:i0 + 4]]
    P = mats[0]
    for j in range(1, 4):
        L = mp.logm(mp.inverse(mats[j - 1]) * mats[j])
        P = P * mp.expm(B[j] * L)
    return P


def logq(q):          # quaternion_math.h:16-59 (half-angle vector)
    v2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2]
    k = mp.atan2(mp.sqrt(v2), q[3]) / mp.sqrt(v2) if v2 > mp.mpf("1e-16") else ONE
    return [k * q[0], k * q[1], k * q[2], ZERO]


def expq(q):          # quaternion_math.h:62-89
    v2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2]
    ea = mp.exp(q[3])
    if v2 > mp.mpf("1e-16"):
        vn = mp.sqrt(v2)
        ka, kv = ea * mp.cos(vn), ea * mp.sin(vn) / vn
    else:
        ka = kv = ea
    return [kv * q[0], kv * q[1], kv * q[2], ka]


def so3_spline(quats, t0, dt, t):
    """uniform_so3_spline_trajectory.h:46-125; Eigen's quaternion products do not renormalise."""
    i0, u = index_and_u(t, t0, dt, len(quats))
    B, dB, _ = basis(u, dt, M_CUMUL)
    q = list(quats[i0])
    parts = [[ZERO, ZERO, ZERO, ONE] for _ in range(3)]
    for j in range(1, 4):
        om = logq(q_mul(q_conj(quats[i0 + j - 1]), quats[i0 + j]))
        e = expq([c * B[j] for c in om])
        q = q_mul(q, e)
        for m in range(3):
            if m == j - 1:
                parts[m] = q_mul(parts[m], [c * dB[j] for c in om])
            parts[m] = q_mul(parts[m], e)
    dq = q_mul(quats[i0], [parts[0][c] + parts[1][c] + parts[2][c] for c in range(4)])
    w = q_mul(dq, q_conj(q))
    return dict(orientation=q, angular_velocity=[TWO * w[0], TWO * w[1], TWO * w[2]], i0=i0)


def r3_spline(vecs, t0, dt, t):
    """uniform_r3_spline_trajectory.h:34-101"""
    i0, u = index_and_u(t, t0, dt, len(vecs))
    Bp, Bv, Ba = basis(u, dt, M_PLAIN)
    comb = lambda Bk: [sum(Bk[k] * vecs[i0 + k][c] for k in range(4)) for c in range(3)]
    return dict(position=comb(Bp), velocity=comb(Bv), acceleration=comb(Ba), i0=i0)


class Trajectory:
    """kind 'se3' (knots n x 7) or 'split' (r3 n x 3, so3 n x 4); plain Python lists of mpf so that a knot entry can be nudged."""

    def __init__(self, kind, dt, t0, knots=None, r3=None, so3=None, compat_zero_dB=False):
        self.kind, self.dt, self.t0, self.compat = kind, mp.mpf(dt), mp.mpf(t0), compat_zero_dB
        self.knots = None if knots is None else [mpv(k) for k in knots]
        self.r3 = None if r3 is None else [mpv(k) for k in r3]
        self.so3 = None if so3 is None else [mpv(k) for k in so3]

    def evaluate(self, t, acc=False):
        if self.kind == "se3":
            return se3_spline(self.knots, self.t0, self.dt, t, compat_zero_dB=self.compat and acc, want_acc=acc)
        a, b = r3_spline(self.r3, self.t0, self.dt, t), so3_spline(self.so3, self.t0, self.dt, t)      # split_trajectory.h:41-58
        return dict(position=a["position"], velocity=a["velocity"], acceleration=a["acceleration"], orientation=b["orientation"],
                    angular_velocity=b["angular_velocity"], i0=a["i0"], i0_so3=b["i0"])

    def params(self):
        """(list-of-lists, row, col) handles of every stored knot double, in the order [SE3 knots] or [R3 knots | SO3 knots]."""
        out = []
        for arr in ([self.knots] if self.kind == "se3" else [self.r3, self.so3]):
            for i, k in enumerate(arr):
                out.extend((arr, i, c) for c in range(len(k)))
        return out


# ---- sensors and measurements ---------------------------------------------------------------------------------------------
def gyroscope(traj, t, time_offset=0):          # imu.h:47-52
    e = traj.evaluate(t + time_offset)
    return q_rot(q_conj(e["orientation"]), e["angular_velocity"])


def accelerometer(traj, t, time_offset=0):      # imu.h:55-59
    e = traj.evaluate(t + time_offset, acc=True)
    return q_rot(q_conj(e["orientation"]), [e["acceleration"][i] + GRAVITY[i] for i in range(3)])


def imu_residual(traj, which, t, y, weight=1, time_offset=0):      # gyroscope_measurement.h:36-38 / accelerometer_measurement.h:37-39
    m = gyroscope(traj, mp.mpf(t), time_offset) if which == 0 else accelerometer(traj, mp.mpf(t), time_offset)
    return [mp.mpf(weight) * (mp.mpf(float(y[i])) - m[i]) for i in range(3)]


def pinhole_project(K, X):                       # pinhole_camera.h:47-51
    p = K * mp.matrix(X)
    return [p[0] / p[2], p[1] / p[2]]


def pinhole_unproject(K, y):                     # pinhole_camera.h:63-67 (mp.inverse: Gaussian elimination, not the cofactor formula)
    x = mp.inverse(K) * mp.matrix([y[0], y[1], ONE])
    return [x[0], x[1], x[2]]


def static_rs_residual(traj, cam, obs_uv, obs_t0, ref_uv, ref_t0, rho, weight=1):
    """static_rscamera_measurement.h:21-55, 89-94.  cam: dict K (3x3), rows, readout, q_ct (x,y,z,w), p_ct, time_offset."""
    K = mp.matrix([[mp.mpf(float(v)) for v in row] for row in cam["K"]])
    q_ct, p_ct = mpv(cam.get("q_ct", (0, 0, 0, 1))), mpv(cam.get("p_ct", (0, 0, 0)))
    d = mp.mpf(cam.get("time_offset", 0.0))
    row_delta = mp.mpf(cam["readout"]) / mp.mpf(cam["rows"])
    ouv, ruv = mpv(obs_uv), mpv(ref_uv)
    t_ref = mp.mpf(ref_t0) + d + ruv[1] * row_delta
    t_obs = mp.mpf(obs_t0) + d + ouv[1] * row_delta
    er, eo = traj.evaluate(t_ref), traj.evaluate(t_obs)
    yh = pinhole_unproject(K, ruv)
    X_ref = q_rot(q_conj(q_ct), [yh[i] - rho * p_ct[i] for i in range(3)])
    Xr = q_rot(er["orientation"], X_ref)
    X = [Xr[i] + er["position"][i] * rho for i in range(3)]
    X_obs = q_rot(q_conj(eo["orientation"]), [X[i] - rho * eo["position"][i] for i in range(3)])
    Xc0 = q_rot(q_ct, X_obs)
    X_cam = [Xc0[i] + p_ct[i] * rho for i in range(3)]
    y = pinhole_project(K, X_cam)
    return [mp.mpf(weight) * (ouv[i] - y[i]) for i in range(2)], er["i0"], eo["i0"]


# ---- widening rows: AtanCamera, NewtonRs / LiftingRs camera measurements, Position / Orientation measurements -------------------------------
def _cam_K(cam):
    return mp.matrix([[mp.mpf(float(v)) for v in row] for row in cam["K"]])


def camera_project(cam, X, dX=None):
    """EvaluateProjection(X, dX, derive): pinhole_camera.h:47-61, atan_camera.h:54-90 (cam["model"] == "atan": wc, gamma).  Returns (y, dy or None)."""
    K = _cam_K(cam)
    if cam.get("model", "pinhole") != "atan":
        p = K * mp.matrix(X)
        y = [p[0] / p[2], p[1] / p[2]]
        if dX is None:
            return y, None
        dp = K * mp.matrix(dX)
        den = p[2] * p[2] + mp.mpf("1e-32")
        return y, [(dp[0] * p[2] - p[0] * dp[2]) / den, (dp[1] * p[2] - p[1] * dp[2]) / den]
    eps, gamma, wc = mp.mpf("1e-32"), mp.mpf(cam["gamma"]), mpv(cam["wc"])
    A = [X[0] / (X[2] + eps), X[1] / (X[2] + eps)]
    L = [A[0] - wc[0], A[1] - wc[1]]
    r = mp.sqrt(L[0] * L[0] + L[1] * L[1] + eps)
    f = mp.atan(r * gamma) / gamma
    g = [L[0] / r, L[1] / r]
    Y = mp.matrix([wc[0] + f * g[0], wc[1] + f * g[1], ONE])
    yy = K * Y
    y = [yy[0], yy[1]]                                     # "Normalization not needed since Y(2) == 1"
    if dX is None:
        return y, None
    dx = (dX[0] * X[2] - X[0] * dX[2]) / (X[2] * X[2] + eps)
    dyy = (dX[1] * X[2] - X[1] * dX[2]) / (X[2] * X[2] + eps)
    common = g[0] * dx + g[1] * dyy
    df = common / (ONE + gamma * gamma * r * r)
    du = f * ((dx * r - L[0] * common) / (r * r)) + df * g[0]
    dv = f * ((dyy * r - L[1] * common) / (r * r)) + df * g[1]
    d = K * mp.matrix([du, dv, ZERO])
    return y, [d[0], d[1]]


def camera_unproject(cam, y):
    """pinhole_camera.h:63-67 / atan_camera.h:92-103"""
    x = mp.inverse(_cam_K(cam)) * mp.matrix([y[0], y[1], ONE])
    if cam.get("model", "pinhole") != "atan":
        return [x[0], x[1], x[2]]
    eps, gamma, wc = mp.mpf("1e-32"), mp.mpf(cam["gamma"]), mpv(cam["wc"])
    L = [x[0] - wc[0], x[1] - wc[1]]
    r = mp.sqrt(L[0] * L[0] + L[1] * L[1] + eps)
    f = mp.tan(r * gamma) / gamma
    return [wc[0] + f * L[0] / r, wc[1] + f * L[1] / r, ONE]


def _landmark(traj, cam, ref_uv, ref_t0, rho):
    """X = R(t_ref) q_ct^-1 (unproject(ref) - rho p_ct) + rho p(t_ref): the lines the three camera measurements share."""
    q_ct, p_ct = mpv(cam.get("q_ct", (0, 0, 0, 1))), mpv(cam.get("p_ct", (0, 0, 0)))
    d = mp.mpf(cam.get("time_offset", 0.0))
    row_delta = mp.mpf(cam["readout"]) / mp.mpf(cam["rows"])
    ruv = mpv(ref_uv)
    er = traj.evaluate(mp.mpf(ref_t0) + d + ruv[1] * row_delta)
    yh = camera_unproject(cam, ruv)
    Xr = q_rot(er["orientation"], q_rot(q_conj(q_ct), [yh[i] - rho * p_ct[i] for i in range(3)]))
    return [Xr[i] + er["position"][i] * rho for i in range(3)], er["i0"], q_ct, p_ct, d, row_delta


def _sandwich(qa, x, qb):      # quaternion_math.h:108-115: vec(qa * (x, 0) * qb)
    return q_mul(q_mul(qa, [x[0], x[1], x[2], ZERO]), qb)[:3]


def newton_rs_residual(traj, cam, obs_uv, obs_t0, ref_uv, ref_t0, rho, weight=1, max_iterations=5):
    """newton_rscamera_measurement.h:23-120, :150-155.  Returns (r (2), i0_ref, number of evaluations of the iteration's body)."""
    X, i0_ref, q_ct, p_ct, d, row_delta = _landmark(traj, cam, ref_uv, ref_t0, rho)
    ouv = mpv(obs_uv)
    rows, readout = mp.mpf(cam["rows"]), mp.mpf(cam["readout"])
    t0_obs = mp.mpf(obs_t0) + d
    t_obs = t0_obs + ouv[1] * row_delta
    max_dt2 = (HALF * readout / rows) ** 2
    lo, hi = t0_obs, t0_obs + readout
    y, n_eval = None, 0
    for _ in range(max_iterations):
        e = traj.evaluate(t_obs)
        n_eval += 1
        p, dp, q, w = e["position"], e["velocity"], e["orientation"], e["angular_velocity"]
        dq = [HALF * v for v in q_mul([w[0], w[1], w[2], ZERO], q)]          # quaternion_math.h:103-106
        dq_inv, q_inv = q_conj(dq), q_conj(q)
        s = [X[i] - rho * p[i] for i in range(3)]
        ds = [-rho * dp[i] for i in range(3)]
        X_obs = q_rot(q_inv, s)
        Xc0 = q_rot(q_ct, X_obs)
        X_cam = [Xc0[i] + rho * p_ct[i] for i in range(3)]
        a1, a2, a3 = _sandwich(dq_inv, s, q), _sandwich(q_inv, ds, q), _sandwich(q_inv, s, dq)
        dX_obs = [a1[i] + a2[i] + a3[i] for i in range(3)]
        dXc0 = q_rot(q_ct, dX_obs)
        dX_cam = [dXc0[i] + rho * p_ct[i] for i in range(3)]               # sic (:92)
        y, dy = camera_project(cam, X_cam, dX_cam)
        f = y[1] - rows * (t_obs - t0_obs) / readout
        df = dy[1] - rows / readout
        dt = f / df
        t_obs = t_obs - dt
        if dt * dt < max_dt2:
            break
        if t_obs < lo:
            t_obs = lo
        elif t_obs > hi:
            t_obs = hi
    return [mp.mpf(weight) * (ouv[i] - y[i]) for i in range(2)], i0_ref, n_eval


def lifting_rs_residual(traj, cam, obs_uv, obs_t0, ref_uv, ref_t0, rho, vt, weight=1):
    """lifting_rscamera_measurement.h:21-56, :98-112; vt_orig = obs.v / rows (:68).  Returns (r (3), i0_ref, i0 of the lifted evaluation)."""
    X, i0_ref, q_ct, p_ct, d, _ = _landmark(traj, cam, ref_uv, ref_t0, rho)
    ouv = mpv(obs_uv)
    rows = mp.mpf(cam["rows"])
    eo = traj.evaluate(mp.mpf(obs_t0) + d + vt * mp.mpf(cam["readout"]))
    X_obs = q_rot(q_conj(eo["orientation"]), [X[i] - rho * eo["position"][i] for i in range(3)])
    Xc0 = q_rot(q_ct, X_obs)
    y, _ = camera_project(cam, [Xc0[i] + p_ct[i] * rho for i in range(3)])
    e = [ouv[0] - y[0], ouv[1] - y[1], rows * (vt - ouv[1] / rows)]
    return [mp.mpf(weight) * v for v in e], i0_ref, eo["i0"]


def position_residual(traj, t, p):                 # position_measurement.h:22-31
    e = traj.evaluate(mp.mpf(t))
    return [mp.mpf(float(p[i])) - e["position"][i] for i in range(3)]


def orientation_residual(traj, t, q):              # orientation_measurement.h:24-32; Eigen 3.3 Quaternion::angularDistance
    d = q_mul(mpv(q), q_conj(traj.evaluate(mp.mpf(t))["orientation"]))
    return [TWO * mp.atan2(mp.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), abs(d[3]))]


def jacobian(fun, handles, h=mp.mpf("1e-25")):
    """Central differences of fun() (list of mpf) with respect to the stored doubles named by `handles` ((container, row, col))."""
    cols = []
    for arr, i, c in handles:
        keep = arr[i][c]
        arr[i][c] = keep + h
        fp = fun()
        arr[i][c] = keep - h
        fm = fun()
        arr[i][c] = keep
        cols.append([(a - b) / (2 * h) for a, b in zip(fp, fm)])
    return [[float(cols[j][r]) for j in range(len(cols))] for r in range(len(cols[0]))]

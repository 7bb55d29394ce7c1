// kontiki_b200 -- NewtonRsCameraMeasurement (measurements/newton_rscamera_measurement.h:23-120) on the SE3 spline.
//
// The reference's Jacobian of this measurement is the forward-mode (ceres::Jet) derivative THROUGH the Newton iteration on
// the row time: t_{k+1} = t_k - f_k / df_k carries a dual part, and df_k comes from hand-written first-derivative formulas
// (:76-96) whose own dual part needs second-order quantities of the spline.  There is no closed form worth having, so
// this path differentiates exactly as the reference does -- one forward-mode direction per thread (dualnum.cuh D1) --
// but on the hoisted structure of the rest of the library:
//   * the knot-pair logs omega_j and their 6x14 ambient Jacobians D_j come from the K0 prepass (no log in the loop):
//     direction "component c of knot k" seeds  d omega_p = Da_p[:, c] (k = p-1) / Db_p[:, c] (k = p)  and knot i0 itself;
//   * the reference side X(t_ref) is evaluated once per landmark (k_landmark_ref record) and enters as a seed
//     dX = dX/dknot[:, c]  (reference-window directions) or dX/drho.
// Operation order follows uniform_se3_spline_trajectory.h:101-194 where it decides which derivative a non-unit knot
// quaternion sees: the raw knot enters through P0.matrix() (:178-190) and the first translation
// step q0 * a1 (Sophus SE3::operator*=), every composed rotation is re-normalised.
#pragma once
#include "spline_math.cuh"
#include "split_math.cuh"      // se3_body_twist (analytic lifting rows)

namespace kb {

template <class T> struct TV3 { T x, y, z; };
template <class T> struct TQ { T x, y, z, w; };
template <class T> struct TM3 { T a[9]; };

template <class T> KB_HD TV3<T> tv3(T x, T y, T z) { TV3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <class T> KB_HD TV3<T> operator+(const TV3<T>& a, const TV3<T>& b) { return tv3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class T> KB_HD TV3<T> operator-(const TV3<T>& a, const TV3<T>& b) { return tv3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class T> KB_HD TV3<T> operator*(const T& s, const TV3<T>& a) { return tv3<T>(s * a.x, s * a.y, s * a.z); }
template <class T> KB_HD T tdot(const TV3<T>& a, const TV3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> KB_HD TV3<T> tcross(const TV3<T>& a, const TV3<T>& b) { return tv3<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
template <class T> KB_HD TM3<T> operator*(const TM3<T>& A, const TM3<T>& B) {
  TM3<T> r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.a[3 * i + j] = A.a[3 * i] * B.a[j] + A.a[3 * i + 1] * B.a[3 + j] + A.a[3 * i + 2] * B.a[6 + j];
  return r; }
template <class T> KB_HD TM3<T> operator+(const TM3<T>& A, const TM3<T>& B) { TM3<T> r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.a[i] = A.a[i] + B.a[i]; return r; }
template <class T> KB_HD TV3<T> operator*(const TM3<T>& A, const TV3<T>& v) {
  return tv3<T>(A.a[0] * v.x + A.a[1] * v.y + A.a[2] * v.z, A.a[3] * v.x + A.a[4] * v.y + A.a[5] * v.z, A.a[6] * v.x + A.a[7] * v.y + A.a[8] * v.z); }
// A * B^T
template <class T> KB_HD TM3<T> tmul_nt(const TM3<T>& A, const TM3<T>& B) {
  TM3<T> r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.a[3 * i + j] = A.a[3 * i] * B.a[3 * j] + A.a[3 * i + 1] * B.a[3 * j + 1] + A.a[3 * i + 2] * B.a[3 * j + 2];
  return r; }
// s * A * hat(v): every row of A crossed with v
template <class T> KB_HD TM3<T> tmul_hat(const T& s, const TM3<T>& A, const TV3<T>& v) {
  TM3<T> r;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const T ax = A.a[3 * i], ay = A.a[3 * i + 1], az = A.a[3 * i + 2];
    r.a[3 * i] = s * (ay * v.z - az * v.y); r.a[3 * i + 1] = s * (az * v.x - ax * v.z); r.a[3 * i + 2] = s * (ax * v.y - ay * v.x); }
  return r; }

// ---- Eigen::Quaternion semantics (x, y, z, w) -------------------------------------------------------------------------------
template <class T> KB_HD TQ<T> tqmul(const TQ<T>& a, const TQ<T>& b) {          // Hamilton product, no renormalisation
  TQ<T> r;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  return r; }
template <class T> KB_HD TQ<T> tqconj(const TQ<T>& q) { TQ<T> r; r.x = -q.x; r.y = -q.y; r.z = -q.z; r.w = q.w; return r; }
template <class T> KB_HD TQ<T> tqnormalized(const TQ<T>& q) {
  const T inv = T(1.0) / t_sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  TQ<T> r; r.x = q.x * inv; r.y = q.y * inv; r.z = q.z * inv; r.w = q.w * inv; return r; }
// Eigen's q * v: v + w (2 u x v) + u x (2 u x v) -- a polynomial in q, not scale invariant
template <class T> KB_HD TV3<T> tqrot(const TQ<T>& q, const TV3<T>& v) {
  const TV3<T> u = tv3<T>(q.x, q.y, q.z);
  TV3<T> uv = tcross(u, v); uv = uv + uv;
  return v + q.w * uv + tcross(u, uv); }
template <class T> KB_HD TM3<T> tqmat(const TQ<T>& q) {                        // Eigen toRotationMatrix (same polynomial)
  const T tx = T(2.0) * q.x, ty = T(2.0) * q.y, tz = T(2.0) * q.z;
  const T twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  TM3<T> r;
  r.a[0] = T(1.0) - (tyy + tzz); r.a[1] = txy - twz; r.a[2] = txz + twy;
  r.a[3] = txy + twz; r.a[4] = T(1.0) - (txx + tzz); r.a[5] = tyz - twx;
  r.a[6] = txz - twy; r.a[7] = tyz + twx; r.a[8] = T(1.0) - (txx + tyy);
  return r; }

// A_j = exp(B omega): rotation quaternion, its matrix E, translation a = V (B upsilon).  Coefficients as entire functions of
// x = theta^2 (series below 1e-2: the closed forms cancel), so that B -> 0 needs no special case.
template <class T> struct TExp { TQ<T> q; TM3<T> E; TV3<T> a; };
template <class T> KB_HD TExp<T> se3_exp_t(const T* om, const T& B) {
  const TV3<T> th = tv3<T>(B * om[3], B * om[4], B * om[5]), up = tv3<T>(B * om[0], B * om[1], B * om[2]);
  const T x = tdot(th, th);
  T sh, ch, cb, cc;        // sin(theta/2)/theta, cos(theta/2), (1 - cos theta)/theta^2, (theta - sin theta)/theta^3
  if (value(x) < KB_SMALL_X) {
    const T h = T(0.25) * x;
    sh = T(0.5) * (T(1.0) + h * (T(-1.0 / 6.0) + h * (T(1.0 / 120.0) + h * (T(-1.0 / 5040.0) + h * T(1.0 / 362880.0)))));
    ch = T(1.0) + h * (T(-0.5) + h * (T(1.0 / 24.0) + h * (T(-1.0 / 720.0) + h * (T(1.0 / 40320.0) + h * T(-1.0 / 3628800.0)))));
    cb = T(0.5) + x * (T(-1.0 / 24.0) + x * (T(1.0 / 720.0) + x * (T(-1.0 / 40320.0) + x * (T(1.0 / 3628800.0) + x * T(-1.0 / 479001600.0)))));
    cc = T(1.0 / 6.0) + x * (T(-1.0 / 120.0) + x * (T(1.0 / 5040.0) + x * (T(-1.0 / 362880.0) + x * (T(1.0 / 39916800.0) + x * T(-1.0 / 6227020800.0)))));
  } else {
    const T theta = t_sqrt(x), half = T(0.5) * theta;
    sh = t_sin(half) / theta; ch = t_cos(half);
    cb = (T(1.0) - t_cos(theta)) / x; cc = (theta - t_sin(theta)) / (x * theta);
  }
  TExp<T> e;
  e.q.x = sh * th.x; e.q.y = sh * th.y; e.q.z = sh * th.z; e.q.w = ch;
  e.E = tqmat(e.q);
  // a = V up,  V = I + cb hat(th) + cc hat(th)^2
  const TV3<T> c1 = tcross(th, up), c2 = tcross(th, c1);
  e.a = up + cb * c1 + cc * c2;
  return e;
}

// position, velocity, orientation, world angular velocity of the cumulative spline at basis (B, dB):
// uniform_se3_spline_trajectory.h:81-99 (outputs) and :101-194 (P, P').
template <class T> struct TEval { TV3<T> p, v, w; TQ<T> q; };
template <class T> KB_HD TEval<T> se3_eval_t(const T* k0, const T* om1, const T* om2, const T* om3, const T* B, const T* dB) {
  const T* om[3] = {om1, om2, om3};
  TExp<T> e[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) e[j] = se3_exp_t<T>(om[j], B[j]);
  TEval<T> r;
  // P = P0 A1 A2 A3 with Sophus' operator*= : t += so3 * a (the raw knot for j = 1), so3 = normalised product
  TQ<T> q; q.x = k0[0]; q.y = k0[1]; q.z = k0[2]; q.w = k0[3];
  const TQ<T> q0 = q;
  TV3<T> t = tv3<T>(k0[4], k0[5], k0[6]);
#pragma unroll
  for (int j = 0; j < 3; ++j) { t = t + tqrot(q, e[j].a); q = tqnormalized(tqmul(q, e[j].q)); }
  r.p = t; r.q = q;
  // P' = P0.matrix() (A1' A2 A3 + A1 A2' A3 + A1 A2 A3'),  A_j' = A_j hat(omega_j) dB_j = dB_j [[E_j hat(phi_j), E_j upsilon_j], [0, 0]]
  TM3<T> Ed[3]; TV3<T> ad[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    Ed[j] = tmul_hat(dB[j], e[j].E, tv3<T>(om[j][3], om[j][4], om[j][5]));
    ad[j] = dB[j] * (e[j].E * tv3<T>(om[j][0], om[j][1], om[j][2]));
  }
  const TM3<T> E23 = e[1].E * e[2].E, E12 = e[0].E * e[1].E;
  const TV3<T> c2 = e[1].E * e[2].a + e[1].a;                                  // translation of A2 A3
  const TM3<T> MR = Ed[0] * E23 + e[0].E * (Ed[1] * e[2].E) + E12 * Ed[2];
  const TV3<T> Mt = (Ed[0] * c2 + ad[0]) + e[0].E * (Ed[1] * e[2].a + ad[1]) + E12 * ad[2];
  const TM3<T> R0 = tqmat(q0);                                                 // raw knot: P0.matrix()
  r.v = R0 * Mt;
  const TM3<T> W = tmul_nt(R0 * MR, tqmat(q));                                 // P'[0:3,0:3] * R(P)^T  (:93-96)
  r.w = tv3<T>(T(0.5) * (W.a[7] - W.a[5]), T(0.5) * (W.a[2] - W.a[6]), T(0.5) * (W.a[3] - W.a[1]));
  return r;
}

// CameraView::EvaluateProjection(X, dX, derive = true): y and the hand-written dy (pinhole_camera.h:47-61, atan_camera.h:54-90)
template <class T> KB_HD void camera_project_t(const CameraConst& cam, const TV3<T>& X, const TV3<T>& dX, T* y, T* dy) {
  const double* K = cam.K;
  if (cam.model == 0) {
    const TV3<T> p = tv3<T>(T(K[0]) * X.x + T(K[1]) * X.y + T(K[2]) * X.z, T(K[3]) * X.x + T(K[4]) * X.y + T(K[5]) * X.z, T(K[6]) * X.x + T(K[7]) * X.y + T(K[8]) * X.z);
    const TV3<T> dp = tv3<T>(T(K[0]) * dX.x + T(K[1]) * dX.y + T(K[2]) * dX.z, T(K[3]) * dX.x + T(K[4]) * dX.y + T(K[5]) * dX.z, T(K[6]) * dX.x + T(K[7]) * dX.y + T(K[8]) * dX.z);
    y[0] = p.x / p.z; y[1] = p.y / p.z;
    const T z2 = p.z * p.z + T(1e-32);
    dy[0] = (dp.x * p.z - p.x * dp.z) / z2;
    dy[1] = (dp.y * p.z - p.y * dp.z) / z2;
    return;
  }
  const T eps = T(1e-32), gamma = T(cam.gamma), wc0 = T(cam.wc[0]), wc1 = T(cam.wc[1]);
  const T A0 = X.x / (X.z + eps), A1 = X.y / (X.z + eps);
  const T L0 = A0 - wc0, L1 = A1 - wc1;
  const T r = t_sqrt((L0 * L0 + L1 * L1) + eps);
  const T f = t_atan(r * gamma) / gamma;
  const T g0 = L0 / r, g1 = L1 / r;
  const T Y0 = wc0 + f * g0, Y1 = wc1 + f * g1;
  y[0] = T(K[0]) * Y0 + T(K[1]) * Y1 + T(K[2]);
  y[1] = T(K[3]) * Y0 + T(K[4]) * Y1 + T(K[5]);
  const T dx = (dX.x * X.z - X.x * dX.z) / (X.z * X.z + eps);
  const T dyy = (dX.y * X.z - X.y * dX.z) / (X.z * X.z + eps);
  const T common = g0 * dx + g1 * dyy;
  const T df = common / (T(1.0) + gamma * gamma * r * r);
  const T dgu = (dx * r - L0 * common) / (r * r);
  const T du = f * dgu + df * g0;
  const T dgv = (dyy * r - L1 * common) / (r * r);
  const T dv = f * dgv + df * g1;
  dy[0] = T(K[0]) * du + T(K[1]) * dv;
  dy[1] = T(K[3]) * du + T(K[4]) * dv;
}

// One direction of a Newton-RS row.  `dir` selects the seed (see the packed layout in include/kontiki_b200.h):
//   [0, 28)            reference-window knot dir/7, component dir%7   -> dX = column of the landmark record's dX/dknots
//   [28, 28 + 7 W)     observation-window knot kbase + (dir-28)/7, component (dir-28)%7
//   28 + 7 W           rho
//   anything else      values only
// Returns 0 or kStatusRange; y = projection y_out (2), dy = its derivative along the seed.
struct NewtonRow { double y[2], dy[2]; int iterations; double t_last; int clamped_last; };      // t_last: row time of the LAST evaluation (the one y comes from)
// The seeds of one direction: the landmark X and rho (reference side, from the record), the observation knot / component that carries the unit
// derivative, the camera's relative pose (plain values here).
struct NewtonSeeds { TV3<D1> X; D1 rho; int ok, oc; TQ<D1> qct; TV3<D1> pct; };
KB_HD NewtonSeeds newton_seeds(const CameraConst& cam, const double* rec, int kbase, int W, int dir) {
  typedef D1 T;
  NewtonSeeds sd;
  const int rho_dir = 28 + 7 * W;
  // record: X(3) | dX/drho(3) | rho | i0_ref | dX/dknots [4][3][7]
  sd.X = tv3<T>(T(rec[0]), T(rec[1]), T(rec[2]));
  sd.rho = T(rec[6]);
  if (dir >= 0 && dir < 28) { const double* d = rec + kRefDOff + 21 * (dir / 7) + dir % 7; sd.X.x.d = d[0]; sd.X.y.d = d[7]; sd.X.z.d = d[14]; }
  else if (dir == rho_dir) { sd.X.x.d = rec[3]; sd.X.y.d = rec[4]; sd.X.z.d = rec[5]; sd.rho.d = 1.0; }
  sd.ok = (dir >= 28 && dir < rho_dir) ? kbase + (dir - 28) / 7 : -1;      // observation-window knot of this direction
  sd.oc = (dir - 28) % 7;
  sd.qct.x = T(cam.q_ct[0]); sd.qct.y = T(cam.q_ct[1]); sd.qct.z = T(cam.q_ct[2]); sd.qct.w = T(cam.q_ct[3]);
  sd.pct = tv3<T>(T(cam.p_ct[0]), T(cam.p_ct[1]), T(cam.p_ct[2]));
  return sd;
}
// ONE evaluation of the iteration's body (newton_rscamera_measurement.h:62-103) at row time t_obs: projection y, f and df on the dual number.
struct NewtonEval { D1 y[2], f, df; };
KB_HD int newton_rs_eval(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, int nseg, const Segment& s0, const Segment& s1,
                         const NewtonSeeds& sd, int kbase, int W, double t0_obs, D1 t_obs, NewtonEval& e) {
  typedef D1 T;
  const double rows = (double)cam.rows;
  // trajectory.Evaluate(t_obs): SplineView::Evaluate + CalculateIndexAndInterpolationAmount (floor drops the derivative)
  int i0; double u0;
  if (locate_in_segments(nseg, s0, s1, t_obs.a, sp.t0, sp.dt, i0, u0) < 0) return kStatusRange;
  if (i0 < kbase || i0 + 4 > kbase + W) return kStatusRange;               // outside the observation span's knots
  const T u = T(u0, t_obs.d / sp.dt);
  const T u2 = u * u, u3 = u2 * u;
  const double di = 1.0 / sp.dt;
  T B[3], dB[3];
  B[0] = (T(5.0) + T(3.0) * u - T(3.0) * u2 + u3) * T(1.0 / 6.0);
  B[1] = (T(1.0) + T(3.0) * u + T(3.0) * u2 - T(2.0) * u3) * T(1.0 / 6.0);
  B[2] = u3 * T(1.0 / 6.0);
  dB[0] = T(di) * (T(3.0) - T(6.0) * u + T(3.0) * u2) * T(1.0 / 6.0);
  dB[1] = T(di) * (T(3.0) + T(6.0) * u - T(6.0) * u2) * T(1.0 / 6.0);
  dB[2] = T(di) * (T(3.0) * u2) * T(1.0 / 6.0);
  T k0[7], om[3][6];
  const double* kn = knots + (size_t)i0 * kKnotStride;
#pragma unroll
  for (int c = 0; c < 7; ++c) k0[c] = T(kn[c], (sd.ok == i0 && sd.oc == c) ? 1.0 : 0.0);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int p = i0 + 1 + j;                                              // pair (knot p-1, knot p)
    const double* pr = pairs + (size_t)p * kPairStride;
    const double* D = sd.ok == p - 1 ? pr + kPairDOff : (sd.ok == p ? pr + kPairDOff + kPairSide : nullptr);
#pragma unroll
    for (int m = 0; m < 6; ++m) om[j][m] = T(pr[m], D ? D[m * 8 + sd.oc] : 0.0);
  }
  const TEval<T> ev = se3_eval_t<T>(k0, om[0], om[1], om[2], B, dB);
  // :66-96
  TQ<T> wq; wq.x = ev.w.x; wq.y = ev.w.y; wq.z = ev.w.z; wq.w = T(0.0);
  TQ<T> dq = tqmul(wq, ev.q); dq.x = T(0.5) * dq.x; dq.y = T(0.5) * dq.y; dq.z = T(0.5) * dq.z; dq.w = T(0.5) * dq.w;
  const TQ<T> dq_inv = tqconj(dq), q_inv = tqconj(ev.q);
  const TV3<T> s = sd.X - sd.rho * ev.p;
  const TV3<T> ds = (T(0.0) - sd.rho) * ev.v;
  const TV3<T> X_obs = tqrot(q_inv, s);
  const TV3<T> X_cam = tqrot(sd.qct, X_obs) + sd.rho * sd.pct;
  TQ<T> sq; sq.x = s.x; sq.y = s.y; sq.z = s.z; sq.w = T(0.0);
  TQ<T> dsq; dsq.x = ds.x; dsq.y = ds.y; dsq.z = ds.z; dsq.w = T(0.0);
  const TQ<T> a1 = tqmul(tqmul(dq_inv, sq), ev.q), a2 = tqmul(tqmul(q_inv, dsq), ev.q), a3 = tqmul(tqmul(q_inv, sq), dq);
  const TV3<T> dX_obs = tv3<T>(a1.x + a2.x + a3.x, a1.y + a2.y + a3.y, a1.z + a2.z + a3.z);
  const TV3<T> dX_cam = tqrot(sd.qct, dX_obs) + sd.rho * sd.pct;                    // sic (:92)
  T dy[2];
  camera_project_t<T>(cam, X_cam, dX_cam, e.y, dy);
  // :101-104
  e.f = e.y[1] - (T(rows) * (t_obs - T(t0_obs)) / T(cam.readout));
  e.df = dy[1] - T(rows / cam.readout);
  return 0;
}
KB_HD int newton_rs_direction(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* rec,
                              const double* obs_uv, double obs_t0, double ref_t0, int kbase, int W, int dir, NewtonRow& out) {
  typedef D1 T;
  Segment s0, s1;
  const int nseg = static_rs_segments(sp, cam, ref_t0, obs_t0, s0, s1);
  if (nseg == 0) return kStatusRange;
  const NewtonSeeds sd = newton_seeds(cam, rec, kbase, W, dir);
  // newton_rscamera_measurement.h:37-58
  const double rows = (double)cam.rows;
  const double t0_obs = add_rn(obs_t0, cam.time_offset);
  T t_obs = T(static_rs_time(cam, obs_t0, obs_uv[1]));                      // not FMA-contracted: it fixes the first knot index
  const double max_dt = 0.5 * cam.readout / rows, max_dt2 = max_dt * max_dt;
  const double min_bound = t0_obs, max_bound = add_rn(t0_obs, cam.readout);
  NewtonEval e; e.y[0] = T(0.0); e.y[1] = T(0.0);
  out.iterations = 0; out.clamped_last = 0; out.t_last = t_obs.a;
  int clamped = 0;
  for (int iter = 0; iter < 5; ++iter) {
    out.t_last = t_obs.a; out.clamped_last = clamped;
    const int st = newton_rs_eval(sp, cam, knots, pairs, nseg, s0, s1, sd, kbase, W, t0_obs, t_obs, e);
    if (st != 0) return st;
    // :105-117
    const T dt = e.f / e.df;
    t_obs = t_obs - dt;
    out.iterations = iter + 1;
    if (dt.a * dt.a < max_dt2) break;
    clamped = 0;
    if (t_obs.a < min_bound) { t_obs = T(min_bound); clamped = 1; }
    else if (t_obs.a > max_bound) { t_obs = T(max_bound); clamped = 1; }
  }
  out.y[0] = e.y[0].a; out.y[1] = e.y[1].a; out.dy[0] = e.y[0].d; out.dy[1] = e.y[1].d;
  return 0;
}
// Rows whose iteration stops after its SECOND evaluation (nearly all rows that iterate at all: the first step lands within a fraction of a
// row) do not need forward mode through both evaluations: y_out = pi(theta, t_1(theta)) with t_1 = t_0 - f/df at t_0, so
//   d y_out / d theta = d pi / d theta at t_1 (the static row at t_1, closed form)  +  pi'(t_1) * d t_1 / d theta,
// and only d t_1 / d theta = -d(f/df)/d theta at t_0 needs the dual number: ONE evaluation per direction instead of two.
// (a) d(f/df)/d theta along `dir` at the initial row time:
KB_HD int newton_rs_first_step_d(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* rec,
                                 const double* obs_uv, double obs_t0, double ref_t0, int kbase, int W, int dir, double& dtd) {
  Segment s0, s1;
  const int nseg = static_rs_segments(sp, cam, ref_t0, obs_t0, s0, s1);
  if (nseg == 0) return kStatusRange;
  const NewtonSeeds sd = newton_seeds(cam, rec, kbase, W, dir);
  NewtonEval e;
  const int st = newton_rs_eval(sp, cam, knots, pairs, nseg, s0, s1, sd, kbase, W, add_rn(obs_t0, cam.time_offset), D1(static_rs_time(cam, obs_t0, obs_uv[1])), e);
  if (st != 0) return st;
  dtd = (e.f / e.df).d;
  return 0;
}
// (b) pi'(t): the TRUE time derivative of the projection at row time t (not the hand-written dy of :76-96, which carries the rho p_ct slip)
KB_HD int newton_rs_time_derivative(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* rec,
                                    double obs_t0, double ref_t0, int kbase, int W, double t, double* pd) {
  Segment s0, s1;
  const int nseg = static_rs_segments(sp, cam, ref_t0, obs_t0, s0, s1);
  if (nseg == 0) return kStatusRange;
  const NewtonSeeds sd = newton_seeds(cam, rec, kbase, W, -1);
  NewtonEval e;
  const int st = newton_rs_eval(sp, cam, knots, pairs, nseg, s0, s1, sd, kbase, W, add_rn(obs_t0, cam.time_offset), D1(t, 1.0), e);
  if (st != 0) return st;
  pd[0] = e.y[0].d; pd[1] = e.y[1].d;
  return 0;
}
// (c) first knot and interpolation amount at an arbitrary row time inside the residual's spans (what the closed-form row at t_1 is built on)
KB_HD bool newton_locate(const SplineConst& sp, const CameraConst& cam, double obs_t0, double ref_t0, double t, int& io, double& uo) {
  Segment s0, s1;
  const int nseg = static_rs_segments(sp, cam, ref_t0, obs_t0, s0, s1);
  return nseg != 0 && locate_in_segments(nseg, s0, s1, t, sp.t0, sp.dt, io, uo) >= 0;
}

// Knots of the observation span {t0_obs - 1e-3, t0_obs + readout + 1e-3} of the residual (newton_rscamera_measurement.h:210-236,
// spline_base.h:371-377), time offset locked: first knot and count.  Every row time the iteration visits lies inside.
KB_HD int newton_obs_window_base(const SplineConst& sp, const CameraConst& cam, double obs_t0) { return knot_floor(sub_rn(obs_t0, 1e-3), sp.t0, sp.dt); }
KB_HD int newton_obs_window_size(const SplineConst& sp, const CameraConst& cam, double obs_t0) {
  return knot_floor(add_rn(add_rn(obs_t0, cam.readout), 1e-3), sp.t0, sp.dt) + 4 - newton_obs_window_base(sp, cam, obs_t0); }

// Where direction `dir` lands in the packed row [ref 4x(2x7) | obs W x(2x7) | rho 2] (residual row 0; row 1 is +7, rho +1)
KB_HD int newton_dir_offset(int dir, int W, int& stride) {
  stride = 7;
  if (dir < 28) return 14 * (dir / 7) + dir % 7;
  if (dir < 28 + 7 * W) { const int d = dir - 28; return 56 + 14 * (d / 7) + d % 7; }
  stride = 1;
  return 56 + 14 * W;
}
// r = weight (uv_obs - y) (:150-155), column j = -weight dy, then ceres::HuberLoss + Corrector as Ceres applies them after Evaluate
KB_HD void newton_rs_finish(const NewtonRow& o, const double* obs_uv, double weight, double huber_c, double* r, double* j) {
  const double r0 = weight * (obs_uv[0] - o.y[0]), r1 = weight * (obs_uv[1] - o.y[1]);
  double j0 = -weight * o.dy[0], j1 = -weight * o.dy[1], rs = 1.0;
  if (huber_c > 0.0) {
    const HuberScale h = huber_scale(huber_c, r0 * r0 + r1 * r1);
    const double rj = r0 * j0 + r1 * j1;
    j0 = h.sqrt_rho1 * (j0 - h.alpha_sq_norm * r0 * rj); j1 = h.sqrt_rho1 * (j1 - h.alpha_sq_norm * r1 * rj);
    rs = h.residual_scaling;
  }
  r[0] = rs * r0; r[1] = rs * r1; j[0] = j0; j[1] = j1;
}
// A whole row, direction after direction (host check; the kernel runs one direction per thread).
KB_HD int newton_rs_row(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* rec,
                        const double* obs_uv, double obs_t0, double ref_t0, int kbase, int W, double weight, double huber_c,
                        double* r, double* J, int* iterations) {
  const int ndir = 29 + 7 * W;
  for (int dir = 0; dir < ndir; ++dir) {
    NewtonRow o;
    const int st = newton_rs_direction(sp, cam, knots, pairs, rec, obs_uv, obs_t0, ref_t0, kbase, W, dir, o);
    if (st != 0) return st;
    double j[2]; int stride;
    newton_rs_finish(o, obs_uv, weight, huber_c, r, j);
    const int off = newton_dir_offset(dir, W, stride);
    J[off] = j[0]; J[off + stride] = j[1];
    *iterations = o.iterations;
  }
  return 0;
}

// The closed-form part of a Newton-RS row (what k_newton_rs_fast does per row).  `row`: 114 staged doubles [Jref 56 | Jobs 56 | rho 2] (the landmark
// record is copied to row + 22 and rewritten in place like in k_static_rs).  Returns
//    0  the iteration stopped after ONE evaluation: the row is the static row at the observed row time, complete;
//    2  it stopped after TWO: `row` holds the static row at t_1 and aux = {y(t_1) (2), pi'(t_1) (2)}; every column still needs
//       + finish(pi'(t_1) * d t_1 / d theta)  with  d t_1 / d theta = -newton_rs_first_step_d  (zero when t_1 was clamped to the readout interval);
//   -1  anything else: the row goes through forward mode (newton_rs_direction per direction).
// rel = position of the row's four active observation knots inside the span (first active knot - kbase).
KB_HD int newton_rs_row_closed(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* rec, const double* ouv,
                               double obs_t0, double ref_t0, int kbase, int W, double weight, double huber_c, bool allow_two, double* row, double* r,
                               int* ir_out, int* rel, double* aux) {
  NewtonRow o;
  if ((int)rec[7] < 0 || newton_rs_direction(sp, cam, knots, pairs, rec, ouv, obs_t0, ref_t0, kbase, W, -1, o) != 0) return -1;
  if (!(o.iterations == 1 || (allow_two && o.iterations == 2))) return -1;
  ObsForward f; f.status = kStatusRange; f.io = -1;
  double uo;
  const bool ok = o.iterations == 1 ? static_rs_row_locate_u(sp, cam, ouv, obs_t0, ref_t0, f.io, uo) : newton_locate(sp, cam, obs_t0, ref_t0, o.t_last, f.io, uo);
  if (!ok || f.io < kbase || f.io + 4 > kbase + W) return -1;
  f.status = 0; f.bo = cumulative_basis(uo, sp.dt);
  static_rs_row_pose(knots, pairs, f);
  double jrho[2];
  int ir = -1, io = -1;
  ObsAdjoint adj;
  for (int c = 0; c < kRefStride; ++c) row[22 + c] = rec[c];
  if (static_rs_row_ref_half(cam, f, row + 22, ouv, weight, huber_c, r, row, jrho, &ir, &io, adj) != 0) return -1;
  static_rs_row_obs_half(knots, pairs, f, adj, row + 56);
  row[112] = jrho[0]; row[113] = jrho[1];
  *ir_out = ir; *rel = io - kbase;
  if (o.iterations == 1) return 0;
  aux[0] = o.y[0]; aux[1] = o.y[1]; aux[2] = 0.0; aux[3] = 0.0;
  if (!o.clamped_last && newton_rs_time_derivative(sp, cam, knots, pairs, rec, obs_t0, ref_t0, kbase, W, o.t_last, aux + 2) != 0) return -1;
  return 2;
}
// ... and the correction of such a row (mode 2) with 32 dual evaluations per row instead of 29 + 7 W (what k_newton_rs_two_w runs, one warp per row).  f/df at the initial row
// time sees the reference window and rho only THROUGH the landmark X (and rho directly), so its derivative along the 28 reference directions is the
// chain rule g_X . dX/dknot[:, c] with g_X from three evaluations seeded on the components of X, along rho g_X . dX/drho + g_rho; and of the 7 W
// observation directions only the four knots active at the initial row time have a derivative at all.  Lane layout:
//   [0, 28) component lane % 7 of active observation knot io + lane / 7 | 28, 29, 30: X.x, X.y, X.z | 31: rho alone.
KB_HD int newton_rs_first_step_lane(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* rec,
                                    const double* obs_uv, double obs_t0, double ref_t0, int kbase, int W, int lane, int* io_out, double& dtd) {
  Segment s0, s1;
  const int nseg = static_rs_segments(sp, cam, ref_t0, obs_t0, s0, s1);
  if (nseg == 0) return kStatusRange;
  const double t = static_rs_time(cam, obs_t0, obs_uv[1]);
  int io; double uo;
  if (locate_in_segments(nseg, s0, s1, t, sp.t0, sp.dt, io, uo) < 0) return kStatusRange;
  *io_out = io;
  NewtonSeeds sd = newton_seeds(cam, rec, kbase, W, -1);
  if (lane < 28) { sd.ok = io + lane / 7; sd.oc = lane % 7; }
  else if (lane == 28) sd.X.x.d = 1.0;
  else if (lane == 29) sd.X.y.d = 1.0;
  else if (lane == 30) sd.X.z.d = 1.0;
  else sd.rho.d = 1.0;
  NewtonEval e;
  const int st = newton_rs_eval(sp, cam, knots, pairs, nseg, s0, s1, sd, kbase, W, add_rn(obs_t0, cam.time_offset), D1(t), e);
  if (st != 0) return st;
  dtd = (e.f / e.df).d;
  return 0;
}
// d(f/df) along a reference-window direction (c < 28) or rho (c == 28) from the gradient g = {g_X (3), g_rho} and the landmark record
KB_HD double newton_rs_ref_chain(const double* rec, const double* g, int c) {
  if (c < 28) { const double* d = rec + kRefDOff + 21 * (c / 7) + c % 7; return g[0] * d[0] + g[1] * d[7] + g[2] * d[14]; }
  return g[0] * rec[3] + g[1] * rec[4] + g[2] * rec[5] + g[3];
}
// finish(pi'(t_1) * d t_1 / d theta) for one column, d t_1 = -d(f/df)
KB_HD void newton_rs_two_step_finish(const double* ouv, double weight, double huber_c, const double* aux, double dtd, double* j) {
  NewtonRow oo; oo.y[0] = aux[0]; oo.y[1] = aux[1]; oo.dy[0] = -aux[2] * dtd; oo.dy[1] = -aux[3] * dtd; oo.iterations = 2;
  double r[2];
  newton_rs_finish(oo, ouv, weight, huber_c, r, j);
}
// Host form of the warp's work: all 32 lanes, then every column that has a correction.
KB_HD int newton_rs_two_step_row_lanes(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* rec, const double* ouv,
                                       double obs_t0, double ref_t0, int kbase, int W, double weight, double huber_c, const double* aux, double* J) {
  if (aux[2] == 0.0 && aux[3] == 0.0) return 0;
  double dtd[32];
  int io = -1;
  for (int lane = 0; lane < 32; ++lane) {
    const int st = newton_rs_first_step_lane(sp, cam, knots, pairs, rec, ouv, obs_t0, ref_t0, kbase, W, lane, &io, dtd[lane]);
    if (st != 0) return st;
  }
  for (int lane = 0; lane < 29; ++lane) {
    double j[2]; int stride;
    if (lane < 28) {
      newton_rs_two_step_finish(ouv, weight, huber_c, aux, dtd[lane], j);
      const int off = newton_dir_offset(28 + 7 * (io - kbase) + lane, W, stride);
      J[off] += j[0]; J[off + stride] += j[1];
    }
    newton_rs_two_step_finish(ouv, weight, huber_c, aux, newton_rs_ref_chain(rec, dtd + 28, lane), j);
    const int off = newton_dir_offset(lane < 28 ? lane : 28 + 7 * W, W, stride);
    J[off] += j[0]; J[off + stride] += j[1];
  }
  return 0;
}
// A whole row the way the two kernels produce it (host check): closed form where the iteration stops after one or two evaluations.
KB_HD int newton_rs_row_fast(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* rec, const double* obs_uv,
                             double obs_t0, double ref_t0, int kbase, int W, double weight, double huber_c, double* r, double* J, int* iterations) {
  double row[114], aux[4];
  int ir = -1, rel = 0;
  const int mode = newton_rs_row_closed(sp, cam, knots, pairs, rec, obs_uv, obs_t0, ref_t0, kbase, W, weight, huber_c, true, row, r, &ir, &rel, aux);
  if (mode < 0) return newton_rs_row(sp, cam, knots, pairs, rec, obs_uv, obs_t0, ref_t0, kbase, W, weight, huber_c, r, J, iterations);
  const int row_len = 58 + 14 * W;
  for (int c = 0; c < row_len; ++c) J[c] = 0.0;
  for (int c = 0; c < 56; ++c) { J[c] = row[c]; J[56 + 14 * rel + c] = row[56 + c]; }
  J[56 + 14 * W] = row[112]; J[57 + 14 * W] = row[113];
  *iterations = mode == 0 ? 1 : 2;
  if (mode == 2) return newton_rs_two_step_row_lanes(sp, cam, knots, pairs, rec, obs_uv, obs_t0, ref_t0, kbase, W, weight, huber_c, aux, J);
  return 0;
}

// ---- NewtonRs rows in CLOSED FORM for any number of evaluations (reverse mode) ----------------------------------------------------------------
// With h(theta, t) = f/df the iteration is t_{k+1} = t_k - h(theta, t_k) and y_out = pi(theta, t_last(theta)), so
//   d y_out / d theta = d pi / d theta |t_last   (the static row at t_last, closed form)   +   pi'(t_last) * D,      D = d t_last / d theta,
//   D_0 = 0,   D_{k+1} = D_k (1 - dh/dt |t_k) - grad_theta h |t_k,   D_{k+1} = 0 when t_{k+1} was clamped to the readout interval
// (the Jets of the reference carry exactly this through `t_obs = t_obs - dt`; a clamped time is a constant).  grad_theta h is ONE reverse sweep:
// on unit quaternions the body of the iteration (newton_rscamera_measurement.h:62-103) is
//   X_obs = R^T (X - rho p),   dX_obs = X_obs x w_b - rho v_b   (the three quaternion products of :76-90 with dq = w q / 2, ds = -rho v),
//   X_cam = R_ct X_obs + rho p_ct,   dX_cam = R_ct dX_obs + rho p_ct (sic, :92),   (y, dy) = projection,   f = y_v - rows (t - t0)/readout,  df = dy_v - rows/readout,
// a scalar function of the pose (R, p), the BODY twist (v_b, w_b), the landmark X and rho.  Its adjoints with respect to the pose go through
// pose_backward<1> (the static row's sweep), those with respect to the body twist through twist_backward<1> (the accelerometer's recursion
// s_j = Ad(A_j^-1) s_{j-1} + dB_j omega_j without its second-derivative part), the reference window and rho follow from grad_X through the landmark
// record.  The radial component along the raw quaternion of knot i0 (the Jacobian is ambient) has three sources: p (inside
// pose_backward), v = R(q0 raw) M_t => dv/ds = 2 (v - R0^T v), and w = vee(R(q0 raw) M_R R^T) => dw/ds = 2 w - vee2(R0^T hat(w)) (as in gyro_se3).
// Second derivatives of the camera model: the gradient of h with respect to (X_cam, dX_cam) is taken by six dual evaluations of the PROJECTION alone.
template <int N>
KB_HD void twist_backward(const double* p1, const double* p2, const double* p3, const Basis& bs, const Mr<N>& Gv, const Mr<N>& Gw, double scale, double* J) {
  // one exp part alive at a time (the reverse order needs A3 first): A2 is rebuilt for its level, like in the accelerometer's sweep
  V3 y2u, y2w, y3u, y3w;
  ExpPart e;
  {
    const V3 u1 = v3(p1[0], p1[1], p1[2]), f1 = v3(p1[3], p1[4], p1[5]);
    const V3 u2 = v3(p2[0], p2[1], p2[2]), f2 = v3(p2[3], p2[4], p2[5]);
    const V3 s1u = bs.dB[0] * u1, s1w = bs.dB[0] * f1;
    exp_part(p2, bs.B[1], true, false, e);
    y2w = mul_t(e.E, s1w); y2u = mul_t(e.E, s1u - cross(e.a, s1w));
    const V3 s2u = y2u + bs.dB[1] * u2, s2w = y2w + bs.dB[1] * f2;
    exp_part(p3, bs.B[2], true, true, e);
    y3w = mul_t(e.E, s2w); y3u = mul_t(e.E, s2u - cross(e.a, s2w));
  }
  G6<N> gs, gj;
  gs.U = Gv; gs.W = Gw;
  gj = gadd(gscale(bs.dB[2], gs), mul_Jr6(mul_ad(gs, y3u, y3w), e, bs.B[2]));
  contract_pair<N, true>(J + 3 * N * 7, gj, p3 + kPairDOff + kPairSide, scale);
  contract_pair<N, true>(J + 2 * N * 7, gj, p3 + kPairDOff, scale);
  gs = mul_Adinv(gs, e.E, e.a);
  exp_part(p2, bs.B[1], true, true, e);
  gj = gadd(gscale(bs.dB[1], gs), mul_Jr6(mul_ad(gs, y2u, y2w), e, bs.B[1]));
  contract_pair<N, true>(J + 2 * N * 7, gj, p2 + kPairDOff + kPairSide, scale);
  contract_pair<N, true>(J + 1 * N * 7, gj, p2 + kPairDOff, scale);
  gs = mul_Adinv(gs, e.E, e.a);
  gj = gscale(bs.dB[0], gs);
  contract_pair<N, true>(J + 1 * N * 7, gj, p1 + kPairDOff + kPairSide, scale);
  contract_pair<N, true>(J + 0 * N * 7, gj, p1 + kPairDOff, scale);
}
// One evaluation of the iteration's body at row time t, plain doubles, on the hoisted structure.
struct NewtonPoint { int io; double uo; Basis bs; Pose P; V3 vb, wb, Xobs, dXobs, Xc, dXc; double y[2], dy[2], f, df, frow; };
KB_HD int newton_point(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* rec, int nseg, const Segment& s0,
                       const Segment& s1, int kbase, int W, double t0_obs, double t, NewtonPoint& e) {
  if (locate_in_segments(nseg, s0, s1, t, sp.t0, sp.dt, e.io, e.uo) < 0) return kStatusRange;
  if (e.io < kbase || e.io + 4 > kbase + W) return kStatusRange;
  e.bs = cumulative_basis(e.uo, sp.dt);
  const double* k0 = knots + (size_t)e.io * kKnotStride;
  const double* p1 = pairs + (size_t)(e.io + 1) * kPairStride; const double* p2 = p1 + kPairStride; const double* p3 = p2 + kPairStride;
  pose_forward(k0, p1, p2, p3, e.bs, e.P);
  KB_SEQ();
  V3 dvb;
  se3_body_twist(p1, p2, p3, e.bs, e.vb, e.wb, dvb);
  KB_SEQ();
  const V3 X = v3(rec[0], rec[1], rec[2]);
  const double rho = rec[6];
  const M3 Rct = load_m3(cam.Rct);
  const V3 pct = v3(cam.p_ct[0], cam.p_ct[1], cam.p_ct[2]);
  e.Xobs = mul_t(e.P.R, X - rho * e.P.p);
  e.dXobs = cross(e.Xobs, e.wb) - rho * e.vb;
  e.Xc = Rct * e.Xobs + rho * pct;
  e.dXc = Rct * e.dXobs + rho * pct;                                               // sic (:92)
  camera_project_t<double>(cam, tv3<double>(e.Xc.x, e.Xc.y, e.Xc.z), tv3<double>(e.dXc.x, e.dXc.y, e.dXc.z), e.y, e.dy);
  const double rows = (double)cam.rows;
  e.frow = rows * (t - t0_obs) / cam.readout;
  e.f = e.y[1] - e.frow;
  e.df = e.dy[1] - rows / cam.readout;
  return 0;
}
// grad h at the point e: gobs [4][7] with respect to the four active observation knots, g = {grad_X (3), d h / d rho at fixed X}
KB_HD void newton_h_gradient(const CameraConst& cam, const double* knots, const double* pairs, const double* rec, const NewtonPoint& e, double* gobs, double* g) {
  const double rows = (double)cam.rows, rho = rec[6];
  double a[6];
#pragma unroll 1
  for (int k = 0; k < 6; ++k) {
    TV3<D1> Xd = tv3<D1>(D1(e.Xc.x, k == 0 ? 1.0 : 0.0), D1(e.Xc.y, k == 1 ? 1.0 : 0.0), D1(e.Xc.z, k == 2 ? 1.0 : 0.0));
    TV3<D1> dXd = tv3<D1>(D1(e.dXc.x, k == 3 ? 1.0 : 0.0), D1(e.dXc.y, k == 4 ? 1.0 : 0.0), D1(e.dXc.z, k == 5 ? 1.0 : 0.0));
    D1 y[2], dy[2];
    camera_project_t<D1>(cam, Xd, dXd, y, dy);
    a[k] = ((y[1] - D1(e.frow)) / (dy[1] - D1(rows / cam.readout))).d;
  }
  const M3 Rct = load_m3(cam.Rct);
  const V3 pct = v3(cam.p_ct[0], cam.p_ct[1], cam.p_ct[2]);
  const V3 aXc = v3(a[0], a[1], a[2]), adXc = v3(a[3], a[4], a[5]);
  const V3 adXobs = mul_t(Rct, adXc);
  const V3 aXobs = mul_t(Rct, aXc) + cross(e.wb, adXobs);
  const V3 awb = cross(adXobs, e.Xobs), avb = (-rho) * adXobs;
  const V3 gX = e.P.R * aXobs;
  g[0] = gX.x; g[1] = gX.y; g[2] = gX.z;
  g[3] = -dot(aXobs, mul_t(e.P.R, e.P.p)) - dot(adXobs, e.vb) + dot(aXc + adXc, pct);
  const double* k0 = knots + (size_t)e.io * kKnotStride;
  const double* p1 = pairs + (size_t)(e.io + 1) * kPairStride; const double* p2 = p1 + kPairStride; const double* p3 = p2 + kPairStride;
  // radial terms of the world velocity and angular velocity (raw R(q0) on the left of P'): one scalar, formed BEFORE the sweeps so that nothing of
  // the forward pass but the basis stays live across them
  double grad[1];
  {
    const M3 R0 = quat_to_rot(k0[0], k0[1], k0[2], k0[3]);
    const V3 v = e.P.R * e.vb, w = e.P.R * e.wb, av = e.P.R * avb, aw = e.P.R * awb;
    const M3 Mm = mul_tn(R0, hat(w));
    const V3 dw = 2.0 * w - v3(Mm.a[7] - Mm.a[5], Mm.a[2] - Mm.a[6], Mm.a[3] - Mm.a[1]);
    grad[0] = dot(av, 2.0 * (v - mul_t(R0, v))) + dot(aw, dw);
  }
  Mr<1> Gp, GpR, Gth, Gv, Gw;
  Gp.a[0] = -rho * gX.x; Gp.a[1] = -rho * gX.y; Gp.a[2] = -rho * gX.z;
  GpR.a[0] = -rho * aXobs.x; GpR.a[1] = -rho * aXobs.y; GpR.a[2] = -rho * aXobs.z;
  { const V3 th = cross(aXobs, e.Xobs); Gth.a[0] = th.x; Gth.a[1] = th.y; Gth.a[2] = th.z; }      // Go hat(Xobs): the row crossed with Xobs
  Gv.a[0] = avb.x; Gv.a[1] = avb.y; Gv.a[2] = avb.z;
  Gw.a[0] = awb.x; Gw.a[1] = awb.y; Gw.a[2] = awb.z;
  const Basis bs = e.bs;
  // The two sweeps are independent until they meet in gobs, and left alone the compiler interleaves them for instruction-level parallelism: 168 registers
  // and 2.7 KB of spill per thread.  KB_SEQ keeps them one after the other: 254 registers, no spill.
  KB_SEQ();
  pose_backward<1>(k0, p1, p2, p3, bs, Gp, GpR, Gth, 1.0, gobs);
  KB_SEQ();
  twist_backward<1>(p1, p2, p3, bs, Gv, Gw, 1.0, gobs);
  KB_SEQ();
  Mr<1> zero; zero.a[0] = zero.a[1] = zero.a[2] = 0.0;
  add_q0_block<1>(gobs, zero, grad, k0, 1.0);
}
// dh/dt at fixed theta: one time-seeded dual evaluation.  Out of line on the device: only rows with three or more evaluations come here, and the
// dual-number evaluation must not set the register budget of the closed-form kernels.
KB_COLD int newton_h_time_derivative(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* rec,
                                     double obs_t0, double ref_t0, int kbase, int W, double t, double* ht) {
  Segment s0, s1;
  const int nseg = static_rs_segments(sp, cam, ref_t0, obs_t0, s0, s1);
  if (nseg == 0) return kStatusRange;
  const NewtonSeeds sd = newton_seeds(cam, rec, kbase, W, -1);
  NewtonEval ev;
  const int st = newton_rs_eval(sp, cam, knots, pairs, nseg, s0, s1, sd, kbase, W, add_rn(obs_t0, cam.time_offset), D1(t, 1.0), ev);
  if (st != 0) return st;
  *ht = (ev.f / ev.df).d;
  return 0;
}
// The iteration on values, and (D != nullptr) D = d t_last / d theta in the span layout [ref 4 x 7 | obs W x 7 | rho] along the way.
struct NewtonIter { int iterations, clamped_last, io; double t_last, uo, y[2], pd[2]; };
KB_HD int newton_rs_iterate(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* rec, const double* obs_uv,
                            double obs_t0, double ref_t0, int kbase, int W, NewtonIter& it, double* D) {
  Segment s0, s1;
  const int nseg = static_rs_segments(sp, cam, ref_t0, obs_t0, s0, s1);
  if (nseg == 0) return kStatusRange;
  const double rows = (double)cam.rows;
  const double t0_obs = add_rn(obs_t0, cam.time_offset);
  double t = static_rs_time(cam, obs_t0, obs_uv[1]);
  const double max_dt = 0.5 * cam.readout / rows, max_dt2 = max_dt * max_dt;
  const double min_bound = t0_obs, max_bound = add_rn(t0_obs, cam.readout);
  const int nD = 29 + 7 * W;
  bool Dzero = true;
  if (D) for (int c = 0; c < nD; ++c) D[c] = 0.0;
  int clamped = 0;
  it.iterations = 0; it.pd[0] = it.pd[1] = 0.0;
#pragma unroll 1
  for (int iter = 0; iter < 5; ++iter) {
    it.t_last = t; it.clamped_last = clamped;
    NewtonPoint e;
    const int st = newton_point(sp, cam, knots, pairs, rec, nseg, s0, s1, kbase, W, t0_obs, t, e);
    if (st != 0) return st;
    const double h = e.f / e.df;
    it.iterations = iter + 1;
    if (h * h < max_dt2 || iter == 4) {
      it.y[0] = e.y[0]; it.y[1] = e.y[1]; it.io = e.io; it.uo = e.uo;
      if (iter > 0 && !clamped) {                       // pi'(t_last): the TRUE time derivative (X_cam' = R_ct dX_obs, without the rho p_ct slip of :92)
        const V3 dX = load_m3(cam.Rct) * e.dXobs;
        D1 y[2], dy[2];
        const TV3<D1> Xd = tv3<D1>(D1(e.Xc.x, dX.x), D1(e.Xc.y, dX.y), D1(e.Xc.z, dX.z)), z = tv3<D1>(D1(0.0), D1(0.0), D1(0.0));
        camera_project_t<D1>(cam, Xd, z, y, dy);
        it.pd[0] = y[0].d; it.pd[1] = y[1].d;
      }
      break;
    }
    if (D) {
      if (!Dzero) {                                     // dh/dt at fixed theta (rows with three or more evaluations only)
        double ht;
        const int st2 = newton_h_time_derivative(sp, cam, knots, pairs, rec, obs_t0, ref_t0, kbase, W, t, &ht);
        if (st2 != 0) return st2;
        const double keep = 1.0 - ht;
        for (int c = 0; c < nD; ++c) D[c] *= keep;
      }
      double gobs[28], g[4];
      KB_SEQ();
      newton_h_gradient(cam, knots, pairs, rec, e, gobs, g);
      KB_SEQ();
#pragma unroll 1
      for (int c = 0; c < 29; ++c) D[c < 28 ? c : nD - 1] -= newton_rs_ref_chain(rec, g, c);
      double* Do = D + 28 + 7 * (e.io - kbase);
#pragma unroll
      for (int c = 0; c < 28; ++c) Do[c] -= gobs[c];
      Dzero = false;
    }
    t = t - h;
    clamped = 0;
    if (t < min_bound) { t = min_bound; clamped = 1; }
    else if (t > max_bound) { t = max_bound; clamped = 1; }
    if (clamped && D && !Dzero) { for (int c = 0; c < nD; ++c) D[c] = 0.0; Dzero = true; }
  }
  return 0;
}
// What finish() makes of a unit of d t_last / d theta: the two entries every column's D is multiplied with (the corrector is linear in the column)
KB_HD void newton_rs_shift_column(const NewtonIter& it, const double* ouv, double weight, double huber_c, double* jfin) {
  NewtonRow oo; oo.y[0] = it.y[0]; oo.y[1] = it.y[1]; oo.dy[0] = it.pd[0]; oo.dy[1] = it.pd[1]; oo.iterations = it.iterations;
  double r[2];
  newton_rs_finish(oo, ouv, weight, huber_c, r, jfin);
}
// The closed-form part of a row for ANY number of evaluations (k_newton_rs_fast with KTK_NEWTON_FAST >= 4): `row` = the static row at t_last, staged
// like in newton_rs_row_closed.  Returns 0: complete (one evaluation, or t_last clamped: D has no effect); 2: aux = {y(t_last), pi'(t_last)} and the row
// still needs + jfin (x) D (k_newton_rs_rev); -1: forward mode.
KB_HD int newton_rs_row_closed_any(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* rec, const double* ouv,
                                   double obs_t0, double ref_t0, int kbase, int W, double weight, double huber_c, double* row, double* r,
                                   int* ir_out, int* rel, double* aux) {
  if ((int)rec[7] < 0) return -1;
  NewtonIter it;
  if (newton_rs_iterate(sp, cam, knots, pairs, rec, ouv, obs_t0, ref_t0, kbase, W, it, nullptr) != 0) return -1;
  ObsForward f; f.status = 0; f.io = it.io; f.bo = cumulative_basis(it.uo, sp.dt);
  static_rs_row_pose(knots, pairs, f);
  double jrho[2];
  int ir = -1, io = -1;
  ObsAdjoint adj;
  for (int c = 0; c < kRefStride; ++c) row[22 + c] = rec[c];
  if (static_rs_row_ref_half(cam, f, row + 22, ouv, weight, huber_c, r, row, jrho, &ir, &io, adj) != 0) return -1;
  static_rs_row_obs_half(knots, pairs, f, adj, row + 56);
  row[112] = jrho[0]; row[113] = jrho[1];
  *ir_out = ir; *rel = io - kbase;
  if (it.pd[0] == 0.0 && it.pd[1] == 0.0) return 0;
  aux[0] = it.y[0]; aux[1] = it.y[1]; aux[2] = it.pd[0]; aux[3] = it.pd[1];
  return 2;
}
// A whole row in closed form (host check of what k_newton_rs_fast + k_newton_rs_rev produce): static row at t_last + jfin (x) D.
KB_HD int newton_rs_row_reverse(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* rec, const double* obs_uv,
                                double obs_t0, double ref_t0, int kbase, int W, double weight, double huber_c, double* r, double* J, int* iterations) {
  double D[29 + 7 * 16];
  if (W > 16 || (int)rec[7] < 0) return newton_rs_row(sp, cam, knots, pairs, rec, obs_uv, obs_t0, ref_t0, kbase, W, weight, huber_c, r, J, iterations);
  NewtonIter it;
  const int st = newton_rs_iterate(sp, cam, knots, pairs, rec, obs_uv, obs_t0, ref_t0, kbase, W, it, D);
  if (st != 0) return st;
  ObsForward f; f.status = 0; f.io = it.io; f.bo = cumulative_basis(it.uo, sp.dt);
  static_rs_row_pose(knots, pairs, f);
  double row[114], jrho[2];
  int ir = -1, io = -1;
  ObsAdjoint adj;
  for (int c = 0; c < kRefStride; ++c) row[22 + c] = rec[c];
  if (static_rs_row_ref_half(cam, f, row + 22, obs_uv, weight, huber_c, r, row, jrho, &ir, &io, adj) != 0) return kStatusRange;
  static_rs_row_obs_half(knots, pairs, f, adj, row + 56);
  const int row_len = 58 + 14 * W, rel = it.io - kbase;
  for (int c = 0; c < row_len; ++c) J[c] = 0.0;
  for (int c = 0; c < 56; ++c) { J[c] = row[c]; J[56 + 14 * rel + c] = row[56 + c]; }
  J[56 + 14 * W] = jrho[0]; J[57 + 14 * W] = jrho[1];
  *iterations = it.iterations;
  if (it.pd[0] != 0.0 || it.pd[1] != 0.0) {
    double jfin[2];
    newton_rs_shift_column(it, obs_uv, weight, huber_c, jfin);
    for (int k = 0; k < 4 + W; ++k)
      for (int c = 0; c < 7; ++c) { J[14 * k + c] += jfin[0] * D[7 * k + c]; J[14 * k + 7 + c] += jfin[1] * D[7 * k + c]; }
    J[56 + 14 * W] += jfin[0] * D[28 + 7 * W]; J[57 + 14 * W] += jfin[1] * D[28 + 7 * W];
  }
  return 0;
}

// ---- LiftingRsCameraMeasurement (measurements/lifting_rscamera_measurement.h:21-56, :98-118) ---------------------------------------------
// The static projection with the observation evaluated at the LIFTED time t_obs = t0_obs + time_offset + vt * readout, vt in [0, 1] a
// parameter block of the measurement; 3 residuals weight * [uv - y ; rows (vt - vt_orig)].  Same forward-mode machinery as the Newton
// rows (one direction at a time, hoisted pair logs and landmark record), no iteration.  The product kernel runs the CLOSED FORM further down
// (lifting_rs_row_analytic); this forward-mode version is the independent cross-check of the host-compiled test harness.  Directions:
//   [0, 28) reference-window knots | [28, 28 + 7 W) observation-span knots | 28 + 7 W: vt | 29 + 7 W: rho | anything else: values only
// Packed row (include/kontiki_b200.h): [ref 4 x (3x7) | obs W x (3x7) | d r/d vt (3) | d r/d rho (3)] = 90 + 21 W doubles.
struct LiftingRow { double y[2], dy[2], dvt; };
KB_HD int lifting_rs_direction(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* rec,
                               double obs_t0, double ref_t0, double vt, int kbase, int W, int dir, LiftingRow& out) {
  typedef D1 T;
  Segment s0, s1;
  const int nseg = static_rs_segments(sp, cam, ref_t0, obs_t0, s0, s1);
  if (nseg == 0) return kStatusRange;
  const int vt_dir = 28 + 7 * W, rho_dir = vt_dir + 1;
  TV3<T> X = tv3<T>(T(rec[0]), T(rec[1]), T(rec[2]));
  T rho = T(rec[6]);
  if (dir >= 0 && dir < 28) { const double* d = rec + kRefDOff + 21 * (dir / 7) + dir % 7; X.x.d = d[0]; X.y.d = d[7]; X.z.d = d[14]; }
  else if (dir == rho_dir) { X.x.d = rec[3]; X.y.d = rec[4]; X.z.d = rec[5]; rho.d = 1.0; }
  const int ok = (dir >= 28 && dir < vt_dir) ? kbase + (dir - 28) / 7 : -1;
  const int oc = (dir - 28) % 7;
  TQ<T> qct; qct.x = T(cam.q_ct[0]); qct.y = T(cam.q_ct[1]); qct.z = T(cam.q_ct[2]); qct.w = T(cam.q_ct[3]);
  const TV3<T> pct = tv3<T>(T(cam.p_ct[0]), T(cam.p_ct[1]), T(cam.p_ct[2]));
  // t_obs = t0_obs + time_offset + vt * readout (:34), left to right, never FMA-contracted: it fixes the first knot index
  const T t_obs = T(add_rn(add_rn(obs_t0, cam.time_offset), mul_rn(vt, cam.readout)), dir == vt_dir ? cam.readout : 0.0);
  out.dvt = dir == vt_dir ? 1.0 : 0.0;
  int i0; double u0;
  if (locate_in_segments(nseg, s0, s1, t_obs.a, sp.t0, sp.dt, i0, u0) < 0) return kStatusRange;
  if (i0 < kbase || i0 + 4 > kbase + W) return kStatusRange;
  const T u = T(u0, t_obs.d / sp.dt);
  const T u2 = u * u, u3 = u2 * u;
  const double di = 1.0 / sp.dt;
  T B[3], dB[3];
  B[0] = (T(5.0) + T(3.0) * u - T(3.0) * u2 + u3) * T(1.0 / 6.0);
  B[1] = (T(1.0) + T(3.0) * u + T(3.0) * u2 - T(2.0) * u3) * T(1.0 / 6.0);
  B[2] = u3 * T(1.0 / 6.0);
  dB[0] = T(di) * (T(3.0) - T(6.0) * u + T(3.0) * u2) * T(1.0 / 6.0);
  dB[1] = T(di) * (T(3.0) + T(6.0) * u - T(6.0) * u2) * T(1.0 / 6.0);
  dB[2] = T(di) * (T(3.0) * u2) * T(1.0 / 6.0);
  T k0[7], om[3][6];
  const double* kn = knots + (size_t)i0 * kKnotStride;
#pragma unroll
  for (int c = 0; c < 7; ++c) k0[c] = T(kn[c], (ok == i0 && oc == c) ? 1.0 : 0.0);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int p = i0 + 1 + j;
    const double* pr = pairs + (size_t)p * kPairStride;
    const double* D = ok == p - 1 ? pr + kPairDOff : (ok == p ? pr + kPairDOff + kPairSide : nullptr);
#pragma unroll
    for (int m = 0; m < 6; ++m) om[j][m] = T(pr[m], D ? D[m * 8 + oc] : 0.0);
  }
  const TEval<T> ev = se3_eval_t<T>(k0, om[0], om[1], om[2], B, dB);
  // :43-54
  const TV3<T> X_obs = tqrot(tqconj(ev.q), X - rho * ev.p);
  const TV3<T> X_cam = tqrot(qct, X_obs) + rho * pct;
  T y[2], dy[2];
  camera_project_t<T>(cam, X_cam, tv3<T>(T(0.0), T(0.0), T(0.0)), y, dy);
  out.y[0] = y[0].a; out.y[1] = y[1].a; out.dy[0] = y[0].d; out.dy[1] = y[1].d;
  return 0;
}
KB_HD int lifting_dir_offset(int dir, int W) {                 // residual row 0 of the direction's column; rows 1, 2 are +7 (+1 for vt / rho)
  if (dir < 28) return 21 * (dir / 7) + dir % 7;
  if (dir < 28 + 7 * W) { const int d = dir - 28; return 84 + 21 * (d / 7) + d % 7; }
  return 84 + 21 * W + 3 * (dir - 28 - 7 * W);
}
// r = weight [uv - y ; rows (vt - vt_orig)] (:105-116), column j = weight [-dy ; rows dvt], then ceres::HuberLoss + Corrector (3 residuals)
KB_HD void lifting_rs_finish(const LiftingRow& o, const CameraConst& cam, const double* obs_uv, double vt, double weight, double huber_c,
                             double* r, double* j) {
  const double rows = (double)cam.rows;
  const double r0 = weight * (obs_uv[0] - o.y[0]), r1 = weight * (obs_uv[1] - o.y[1]), r2 = weight * (rows * (vt - obs_uv[1] / rows));
  double j0 = -weight * o.dy[0], j1 = -weight * o.dy[1], j2 = weight * (rows * o.dvt), rs = 1.0;
  if (huber_c > 0.0) {
    const HuberScale h = huber_scale(huber_c, r0 * r0 + r1 * r1 + r2 * r2);
    const double rj = r0 * j0 + r1 * j1 + r2 * j2;
    j0 = h.sqrt_rho1 * (j0 - h.alpha_sq_norm * r0 * rj); j1 = h.sqrt_rho1 * (j1 - h.alpha_sq_norm * r1 * rj); j2 = h.sqrt_rho1 * (j2 - h.alpha_sq_norm * r2 * rj);
    rs = h.residual_scaling;
  }
  r[0] = rs * r0; r[1] = rs * r1; r[2] = rs * r2; j[0] = j0; j[1] = j1; j[2] = j2;
}
// A whole row, direction after direction (host check; the kernel runs one direction per thread).
KB_HD int lifting_rs_row(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* rec,
                         const double* obs_uv, double obs_t0, double ref_t0, double vt, int kbase, int W, double weight, double huber_c,
                         double* r, double* J) {
  const int ndir = 30 + 7 * W;
  for (int dir = 0; dir < ndir; ++dir) {
    LiftingRow o;
    const int st = lifting_rs_direction(sp, cam, knots, pairs, rec, obs_t0, ref_t0, vt, kbase, W, dir, o);
    if (st != 0) return st;
    double j[3];
    lifting_rs_finish(o, cam, obs_uv, vt, weight, huber_c, r, j);
    const int off = lifting_dir_offset(dir, W);
    const int stride = dir < 28 + 7 * W ? 7 : 1;
    J[off] = j[0]; J[off + stride] = j[1]; J[off + 2 * stride] = j[2];
  }
  return 0;
}

// ---- sensor-block columns of NewtonRs / LiftingRs rows (KTK_EVAL_SENSOR_JACOBIANS; sensors.h:135-165) ----------------------------------------
// Direction c of the camera's own parameter blocks: 0..3 relative orientation q_ct (ambient, x y z w), 4..6 relative position p_ct.  Both
// sides of the row depend on them -- the landmark X(t_ref) = q_r (conj(q_ct) (yh - rho p_ct)) + rho p_r (newton_rscamera_measurement.h:50-53,
// lifting_rscamera_measurement.h:36-41) and the projection q_ct X_obs + rho p_ct -- so the reference side is evaluated here on the dual number
// instead of being read from the landmark record.  Eigen's q * v is a polynomial in the RAW q_ct (tqrot), which is what the derivative along a
// non-unit direction sees.  The time offset of these two measurements stays locked (unlocked: KTK_EUNSUPPORTED; the reference moves both spans).
template <class T> struct SensorSeed { TQ<T> qct; TV3<T> pct; };
KB_HD SensorSeed<D1> camera_sensor_seed(const CameraConst& cam, int c) {
  SensorSeed<D1> s;
  s.qct.x = D1(cam.q_ct[0], c == 0 ? 1.0 : 0.0); s.qct.y = D1(cam.q_ct[1], c == 1 ? 1.0 : 0.0);
  s.qct.z = D1(cam.q_ct[2], c == 2 ? 1.0 : 0.0); s.qct.w = D1(cam.q_ct[3], c == 3 ? 1.0 : 0.0);
  s.pct = tv3<D1>(D1(cam.p_ct[0], c == 4 ? 1.0 : 0.0), D1(cam.p_ct[1], c == 5 ? 1.0 : 0.0), D1(cam.p_ct[2], c == 6 ? 1.0 : 0.0));
  return s;
}
// plain-value spline evaluation at (i0, u0) lifted to T (no knot seeds)
template <class T> KB_HD TEval<T> se3_eval_plain_t(const SplineConst& sp, const double* knots, const double* pairs, int i0, double u0) {
  const T u = T(u0);
  const T u2 = u * u, u3 = u2 * u;
  const double di = 1.0 / sp.dt;
  T B[3], dB[3];
  B[0] = (T(5.0) + T(3.0) * u - T(3.0) * u2 + u3) * T(1.0 / 6.0);
  B[1] = (T(1.0) + T(3.0) * u + T(3.0) * u2 - T(2.0) * u3) * T(1.0 / 6.0);
  B[2] = u3 * T(1.0 / 6.0);
  dB[0] = T(di) * (T(3.0) - T(6.0) * u + T(3.0) * u2) * T(1.0 / 6.0);
  dB[1] = T(di) * (T(3.0) + T(6.0) * u - T(6.0) * u2) * T(1.0 / 6.0);
  dB[2] = T(di) * (T(3.0) * u2) * T(1.0 / 6.0);
  T k0[7], om[3][6];
  const double* kn = knots + (size_t)i0 * kKnotStride;
#pragma unroll
  for (int c = 0; c < 7; ++c) k0[c] = T(kn[c]);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double* pr = pairs + (size_t)(i0 + 1 + j) * kPairStride;
#pragma unroll
    for (int m = 0; m < 6; ++m) om[j][m] = T(pr[m]);
  }
  return se3_eval_t<T>(k0, om[0], om[1], om[2], B, dB);
}
// X(t_ref) on the dual number: the reference pose in plain doubles, the camera's relative pose seeded
KB_HD int sensor_ref_point(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, int nseg, const Segment& s0,
                           const Segment& s1, const double* ref_uv, double ref_t0, double rho, const SensorSeed<D1>& sd, TV3<D1>& X) {
  typedef D1 T;
  int i0; double u0;
  if (locate_in_segments(nseg, s0, s1, static_rs_time(cam, ref_t0, ref_uv[1]), sp.t0, sp.dt, i0, u0) < 0) return kStatusRange;
  if (i0 < 0 || i0 + 3 >= sp.n_knots) return kStatusRange;
  const TEval<T> er = se3_eval_plain_t<T>(sp, knots, pairs, i0, u0);
  const V3 yh = camera_unproject(cam, ref_uv[0], ref_uv[1]);
  const TV3<T> Xref = tqrot(tqconj(sd.qct), tv3<T>(T(yh.x), T(yh.y), T(yh.z)) - T(rho) * sd.pct);
  X = tqrot(er.q, Xref) + T(rho) * er.p;
  return 0;
}
// NewtonRs: column c of the sensor blocks; same iteration as newton_rs_direction, knots unseeded
KB_HD int newton_rs_sensor_direction(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* ref_uv,
                                     double ref_t0, double rho_v, const double* obs_uv, double obs_t0, int kbase, int W, int c, NewtonRow& out) {
  typedef D1 T;
  Segment s0, s1;
  const int nseg = static_rs_segments(sp, cam, ref_t0, obs_t0, s0, s1);
  if (nseg == 0) return kStatusRange;
  const SensorSeed<T> sd = camera_sensor_seed(cam, c);
  TV3<T> X;
  if (sensor_ref_point(sp, cam, knots, pairs, nseg, s0, s1, ref_uv, ref_t0, rho_v, sd, X) != 0) return kStatusRange;
  const T rho = T(rho_v);
  const double rows = (double)cam.rows;
  const double t0_obs = add_rn(obs_t0, cam.time_offset);
  T t_obs = T(static_rs_time(cam, obs_t0, obs_uv[1]));
  const double max_dt = 0.5 * cam.readout / rows, max_dt2 = max_dt * max_dt;
  const double min_bound = t0_obs, max_bound = add_rn(t0_obs, cam.readout);
  T y[2] = {T(0.0), T(0.0)};
  out.iterations = 0;
  for (int iter = 0; iter < 5; ++iter) {
    int i0; double u0;
    if (locate_in_segments(nseg, s0, s1, t_obs.a, sp.t0, sp.dt, i0, u0) < 0) return kStatusRange;
    if (i0 < kbase || i0 + 4 > kbase + W) return kStatusRange;
    // the row time carries a derivative from the second iteration on: u = (u0, dt_obs / dt)
    const T u = T(u0, t_obs.d / sp.dt);
    const T u2 = u * u, u3 = u2 * u;
    const double di = 1.0 / sp.dt;
    T B[3], dB[3];
    B[0] = (T(5.0) + T(3.0) * u - T(3.0) * u2 + u3) * T(1.0 / 6.0);
    B[1] = (T(1.0) + T(3.0) * u + T(3.0) * u2 - T(2.0) * u3) * T(1.0 / 6.0);
    B[2] = u3 * T(1.0 / 6.0);
    dB[0] = T(di) * (T(3.0) - T(6.0) * u + T(3.0) * u2) * T(1.0 / 6.0);
    dB[1] = T(di) * (T(3.0) + T(6.0) * u - T(6.0) * u2) * T(1.0 / 6.0);
    dB[2] = T(di) * (T(3.0) * u2) * T(1.0 / 6.0);
    T k0[7], om[3][6];
    const double* kn = knots + (size_t)i0 * kKnotStride;
#pragma unroll
    for (int cc = 0; cc < 7; ++cc) k0[cc] = T(kn[cc]);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double* pr = pairs + (size_t)(i0 + 1 + j) * kPairStride;
#pragma unroll
      for (int m = 0; m < 6; ++m) om[j][m] = T(pr[m]);
    }
    const TEval<T> ev = se3_eval_t<T>(k0, om[0], om[1], om[2], B, dB);
    TQ<T> wq; wq.x = ev.w.x; wq.y = ev.w.y; wq.z = ev.w.z; wq.w = T(0.0);
    TQ<T> dq = tqmul(wq, ev.q); dq.x = T(0.5) * dq.x; dq.y = T(0.5) * dq.y; dq.z = T(0.5) * dq.z; dq.w = T(0.5) * dq.w;
    const TQ<T> dq_inv = tqconj(dq), q_inv = tqconj(ev.q);
    const TV3<T> s = X - rho * ev.p;
    const TV3<T> ds = (T(0.0) - rho) * ev.v;
    const TV3<T> X_obs = tqrot(q_inv, s);
    const TV3<T> X_cam = tqrot(sd.qct, X_obs) + rho * sd.pct;
    TQ<T> sq; sq.x = s.x; sq.y = s.y; sq.z = s.z; sq.w = T(0.0);
    TQ<T> dsq; dsq.x = ds.x; dsq.y = ds.y; dsq.z = ds.z; dsq.w = T(0.0);
    const TQ<T> a1 = tqmul(tqmul(dq_inv, sq), ev.q), a2 = tqmul(tqmul(q_inv, dsq), ev.q), a3 = tqmul(tqmul(q_inv, sq), dq);
    const TV3<T> dX_obs = tv3<T>(a1.x + a2.x + a3.x, a1.y + a2.y + a3.y, a1.z + a2.z + a3.z);
    const TV3<T> dX_cam = tqrot(sd.qct, dX_obs) + rho * sd.pct;                 // sic (:92)
    T dy[2];
    camera_project_t<T>(cam, X_cam, dX_cam, y, dy);
    const T f = y[1] - (T(rows) * (t_obs - T(t0_obs)) / T(cam.readout));
    const T df = dy[1] - T(rows / cam.readout);
    const T dt = f / df;
    t_obs = t_obs - dt;
    out.iterations = iter + 1;
    if (dt.a * dt.a < max_dt2) break;
    if (t_obs.a < min_bound) t_obs = T(min_bound);
    else if (t_obs.a > max_bound) t_obs = T(max_bound);
  }
  out.y[0] = y[0].a; out.y[1] = y[1].a; out.dy[0] = y[0].d; out.dy[1] = y[1].d;
  return 0;
}
// LiftingRs: column c of the sensor blocks (the third residual does not depend on the camera pose: dvt = 0)
KB_HD int lifting_rs_sensor_direction(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* ref_uv,
                                      double ref_t0, double rho_v, double obs_t0, double vt, int kbase, int W, int c, LiftingRow& out) {
  typedef D1 T;
  Segment s0, s1;
  const int nseg = static_rs_segments(sp, cam, ref_t0, obs_t0, s0, s1);
  if (nseg == 0) return kStatusRange;
  const SensorSeed<T> sd = camera_sensor_seed(cam, c);
  TV3<T> X;
  if (sensor_ref_point(sp, cam, knots, pairs, nseg, s0, s1, ref_uv, ref_t0, rho_v, sd, X) != 0) return kStatusRange;
  const T rho = T(rho_v);
  const double t_obs = add_rn(add_rn(obs_t0, cam.time_offset), mul_rn(vt, cam.readout));
  int i0; double u0;
  if (locate_in_segments(nseg, s0, s1, t_obs, sp.t0, sp.dt, i0, u0) < 0) return kStatusRange;
  if (i0 < kbase || i0 + 4 > kbase + W) return kStatusRange;
  const TEval<T> ev = se3_eval_plain_t<T>(sp, knots, pairs, i0, u0);
  const TV3<T> X_obs = tqrot(tqconj(ev.q), X - rho * ev.p);
  const TV3<T> X_cam = tqrot(sd.qct, X_obs) + rho * sd.pct;
  T y[2], dy[2];
  camera_project_t<T>(cam, X_cam, tv3<T>(T(0.0), T(0.0), T(0.0)), y, dy);
  out.y[0] = y[0].a; out.y[1] = y[1].a; out.dy[0] = y[0].d; out.dy[1] = y[1].d; out.dvt = 0.0;
  return 0;
}

// Column c (0..6) of a row's sensor blocks, Huber-corrected like its knot columns, into Js_row = [q_ct (nres x 4) | p_ct (nres x 3) | time offset (nres)].
KB_HD int span_sensor_column(bool lifting, const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* ref_uv,
                             double ref_t0, double rho, const double* obs_uv, double obs_t0, double vt, int kbase, int W, double weight, double huber_c,
                             int c, double* Js_row) {
  const int nres = lifting ? 3 : 2;
  double r[3], j[3];
  if (lifting) {
    LiftingRow o;
    const int st = lifting_rs_sensor_direction(sp, cam, knots, pairs, ref_uv, ref_t0, rho, obs_t0, vt, kbase, W, c, o);
    if (st != 0) return st;
    lifting_rs_finish(o, cam, obs_uv, vt, weight, huber_c, r, j);
  } else {
    NewtonRow o;
    const int st = newton_rs_sensor_direction(sp, cam, knots, pairs, ref_uv, ref_t0, rho, obs_uv, obs_t0, kbase, W, c, o);
    if (st != 0) return st;
    newton_rs_finish(o, obs_uv, weight, huber_c, r, j);
  }
  for (int rr = 0; rr < nres; ++rr) {
    if (c < 4) Js_row[4 * rr + c] = j[rr]; else Js_row[4 * nres + 3 * rr + (c - 4)] = j[rr];
    if (c == 0) Js_row[7 * nres + rr] = 0.0;      // time offset: locked for these measurements
  }
  return 0;
}

// ---- LiftingRs rows in CLOSED FORM (one thread per row, the static kernel's machinery with three residual rows) ------------------------------
// Everything but the d/d vt column is the static row evaluated at the lifted time: with the corrector C (3 x 3, identity without the loss)
//   d r / d Xc = C [:, 0:2] (-weight) d y / d Xc          (3 x 3; the timing residual does not see the point)
// runs through the same reference-window product and the same reverse sweep (pose_backward<3>).  The row-time column uses the body twist
// (v_b, w_b) of the spline at t_obs:  d Xobs / d t = -w_b x Xobs - rho v_b   (Xobs = R^T (X - rho p)),  so
//   d r / d vt = (d r / d Xc) R_ct (d Xobs / d t) readout + C [:, 2] weight rows.
// Outputs: r (3), Jref [4][3][7] (84), Jobs [4][3][7] (84: the ACTIVE window, first knot *i0_obs), Jvt (3), Jrho (3).
// Two parts: the first consumes the landmark record (residual, reference-window blocks, the two tail columns as values), the second is the reverse sweep
// into the observation blocks and needs none of it (a kernel may reuse the record's staging space in between; a sequence point keeps them apart).
struct LiftingMid { Mr<3> Gp, GpR, Gth; double jvt[3], jrho[3]; int io; double uo; };
KB_HD int lifting_rs_row_first(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* rec,
                               const double* obs_uv, double obs_t0, double ref_t0, double vt, double weight, double huber_c,
                               double* r, double* Jref, int* i0_ref, LiftingMid& mid) {
  double* Jvt = mid.jvt; double* Jrho = mid.jrho;
  int i0_obs_v; int* i0_obs = &i0_obs_v;
  Segment s0, s1;
  const int nseg = static_rs_segments(sp, cam, ref_t0, obs_t0, s0, s1);
  if (nseg == 0) return kStatusRange;
  const double t_obs = add_rn(add_rn(obs_t0, cam.time_offset), mul_rn(vt, cam.readout));
  int io; double uo;
  if (locate_in_segments(nseg, s0, s1, t_obs, sp.t0, sp.dt, io, uo) < 0) return kStatusRange;
  if (io < 0 || io + 3 >= sp.n_knots) return kStatusRange;
  *i0_ref = (int)rec[7]; *i0_obs = io;
  if (*i0_ref < 0) return kStatusRange;
  const Basis bs = cumulative_basis(uo, sp.dt);
  const double* k0 = knots + (size_t)io * kKnotStride;
  const double* p1 = pairs + (size_t)(io + 1) * kPairStride; const double* p2 = p1 + kPairStride; const double* p3 = p2 + kPairStride;
  Pose P;
  pose_forward(k0, p1, p2, p3, bs, P);
  KB_SEQ();
  V3 vb, wb, dvb;
  se3_body_twist(p1, p2, p3, bs, vb, wb, dvb);
  KB_SEQ();
  const V3 X = v3(rec[0], rec[1], rec[2]), dXr = v3(rec[3], rec[4], rec[5]);
  const double rho = rec[6];
  const M3 Rct = load_m3(cam.Rct);
  const V3 pct = v3(cam.p_ct[0], cam.p_ct[1], cam.p_ct[2]);
  const V3 Xobs = mul_t(P.R, X - rho * P.p);
  const V3 Xc = Rct * Xobs + rho * pct;
  double y0, y1;
  Mr<2> Jp0;
  camera_project_jac(cam, Xc, y0, y1, Jp0);
  const double rows = (double)cam.rows;
  const double r0 = weight * (obs_uv[0] - y0), r1 = weight * (obs_uv[1] - y1), r2 = weight * (rows * (vt - obs_uv[1] / rows));
  double C[9] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0}, rs = 1.0;
  if (huber_c > 0.0) {
    const HuberScale h = huber_scale(huber_c, r0 * r0 + r1 * r1 + r2 * r2);
    const double rr[3] = {r0, r1, r2};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) C[3 * i + j] = h.sqrt_rho1 * ((i == j ? 1.0 : 0.0) - h.alpha_sq_norm * rr[i] * rr[j]);
    rs = h.residual_scaling;
  }
  r[0] = rs * r0; r[1] = rs * r1; r[2] = rs * r2;
  Mr<3> Jp;                                       // d r / d Xc, corrected
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int c = 0; c < 3; ++c) Jp.a[3 * i + c] = -weight * (C[3 * i] * Jp0.a[c] + C[3 * i + 1] * Jp0.a[3 + c]);
  const Mr<3> Go = rmul(Jp, Rct);                 // d r / d Xobs
  const Mr<3> GX = rmul_nt(Go, P.R);              // d r / d X
  const V3 dXc = Rct * mul_t(P.R, dXr - P.p) + pct;
  const V3 dXobs_dt = (-1.0) * cross(wb, Xobs) - rho * vb;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    Jrho[i] = Jp.a[3 * i] * dXc.x + Jp.a[3 * i + 1] * dXc.y + Jp.a[3 * i + 2] * dXc.z;
    Jvt[i] = cam.readout * (Go.a[3 * i] * dXobs_dt.x + Go.a[3 * i + 1] * dXobs_dt.y + Go.a[3 * i + 2] * dXobs_dt.z) + C[3 * i + 2] * (weight * rows);
  }
  KB_SEQ();
  // reference-window blocks: GX (3x3) * dX/dknot_k (3x7)
  const double* dXk = rec + kRefDOff;
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int c = 0; c < 7; ++c)
        Jref[21 * k + 7 * i + c] = GX.a[3 * i] * dXk[21 * k + c] + GX.a[3 * i + 1] * dXk[21 * k + 7 + c] + GX.a[3 * i + 2] * dXk[21 * k + 14 + c];
  mid.Gp = rscale(-rho, GX); mid.GpR = rscale(-rho, Go); mid.Gth = rmul_hat(Go, Xobs); mid.io = io; mid.uo = uo;
  return 0;
}
KB_HD void lifting_rs_row_second(const SplineConst& sp, const double* knots, const double* pairs, const LiftingMid& mid, double* Jobs) {
  const Basis bs = cumulative_basis(mid.uo, sp.dt);
  const double* k0 = knots + (size_t)mid.io * kKnotStride;
  const double* p1 = pairs + (size_t)(mid.io + 1) * kPairStride; const double* p2 = p1 + kPairStride; const double* p3 = p2 + kPairStride;
  pose_backward<3>(k0, p1, p2, p3, bs, mid.Gp, mid.GpR, mid.Gth, 1.0, Jobs);
}
KB_HD int lifting_rs_row_analytic(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* rec,
                                  const double* obs_uv, double obs_t0, double ref_t0, double vt, double weight, double huber_c,
                                  double* r, double* Jref, double* Jobs, double* Jvt, double* Jrho, int* i0_ref, int* i0_obs) {
  LiftingMid mid;
  const int st = lifting_rs_row_first(sp, cam, knots, pairs, rec, obs_uv, obs_t0, ref_t0, vt, weight, huber_c, r, Jref, i0_ref, mid);
  if (st != 0) return st;
  *i0_obs = mid.io;
#pragma unroll
  for (int c = 0; c < 3; ++c) { Jvt[c] = mid.jvt[c]; Jrho[c] = mid.jrho[c]; }
  KB_SEQ();
  lifting_rs_row_second(sp, knots, pairs, mid, Jobs);
  return 0;
}
// ... packed into the C ABI's row [ref 4 x (3x7) | obs W x (3x7) | vt 3 | rho 3] (the blocks of the span outside the active window are zero)
KB_HD int lifting_rs_row_packed(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* rec,
                                const double* obs_uv, double obs_t0, double ref_t0, double vt, int kbase, int W, double weight, double huber_c,
                                double* r, double* J) {
  double Jref[84], Jobs[84], Jvt[3], Jrho[3];
  int ir, io;
  const int st = lifting_rs_row_analytic(sp, cam, knots, pairs, rec, obs_uv, obs_t0, ref_t0, vt, weight, huber_c, r, Jref, Jobs, Jvt, Jrho, &ir, &io);
  if (st != 0) return st;
  if (io < kbase || io + 4 > kbase + W) return kStatusRange;
  for (int c = 0; c < 84; ++c) J[c] = Jref[c];
  for (int c = 0; c < 21 * W; ++c) J[84 + c] = 0.0;
  for (int c = 0; c < 84; ++c) J[84 + 21 * (io - kbase) + c] = Jobs[c];
  for (int c = 0; c < 3; ++c) { J[84 + 21 * W + c] = Jvt[c]; J[87 + 21 * W + c] = Jrho[c]; }
  return 0;
}

// =====================================================================================================================================
// NewtonRs / LiftingRs rows on a SPLIT trajectory (UniformR3 + UniformSO3 splines; split_trajectory.h:41-58): forward mode, one direction at a
// time, like the SE3 rows above.  The R3 spline is linear in its knots (uniform_r3_spline_trajectory.h:34-101); the SO3 spline is evaluated
// exactly as the reference does (uniform_so3_spline_trajectory.h:46-125: q = q0 prod expq(B_j omega_j), dq by the product rule, angular
// velocity 2 (dq conj(q)).vec) on the hoisted half-angle vectors omega_j = logq(conj(q_{j-1}) q_j) of the SO3 prepass, whose 3 x 8 ambient
// Jacobians seed the knot directions.  Packed rows (nres = 2 Newton, 3 Lifting):
//   [ref R3 4 x (nres x 3) | ref SO3 4 x (nres x 4) | obs R3 Wa x (nres x 3) | obs SO3 Wb x (nres x 4) | (vt nres) | rho nres]
// Directions, in that order: 12 + 16 + 3 Wa + 4 Wb (+ vt) + rho.
// =====================================================================================================================================
template <class T> KB_HD TQ<T> expq_t(const TV3<T>& v) {      // quaternion_math.h:61-86 with a zero scalar part (exp(0) = 1)
  const T v2 = tdot(v, v);
  T ka(1.0), kv(1.0);
  if (value(v2) > kEpsLogq) { const T vn = t_sqrt(v2); ka = t_cos(vn); kv = t_sin(vn) / vn; }
  TQ<T> q; q.x = kv * v.x; q.y = kv * v.y; q.z = kv * v.z; q.w = ka;
  return q;
}
struct SplitSeed { int spline, knot, comp; };      // spline: 0 none, 1 R3, 2 SO3
// position / velocity of the R3 spline and orientation / world angular velocity of the SO3 spline at (ia, ua) / (ib, ub); td = d t / d seed
template <class T> KB_HD TEval<T> split_eval_t(const SplitConst& sp, const double* vecs, const double* quats, const double* so3pairs, int ia, double ua0,
                                               int ib, double ub0, double td, const SplitSeed& sd) {
  TEval<T> r;
  {      // R3: p = sum (U M)_k c_k, v = sum (dU M)_k c_k
    const T u = T(ua0, td / sp.dt_r3), u2 = u * u, u3 = u2 * u;
    const T di = T(1.0 / sp.dt_r3), s6 = T(1.0 / 6.0);
    T Bp[4], Bv[4];
    Bp[0] = (T(1.0) - T(3.0) * u + T(3.0) * u2 - u3) * s6; Bp[1] = (T(4.0) - T(6.0) * u2 + T(3.0) * u3) * s6;
    Bp[2] = (T(1.0) + T(3.0) * u + T(3.0) * u2 - T(3.0) * u3) * s6; Bp[3] = u3 * s6;
    Bv[0] = di * (T(-3.0) + T(6.0) * u - T(3.0) * u2) * s6; Bv[1] = di * (T(-12.0) * u + T(9.0) * u2) * s6;
    Bv[2] = di * (T(3.0) + T(6.0) * u - T(9.0) * u2) * s6; Bv[3] = di * (T(3.0) * u2) * s6;
    r.p = tv3<T>(T(0.0), T(0.0), T(0.0)); r.v = r.p;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double* c = vecs + (size_t)(ia + k) * kVecStride;
      const bool hit = sd.spline == 1 && sd.knot == ia + k;
      const TV3<T> cp = tv3<T>(T(c[0], hit && sd.comp == 0 ? 1.0 : 0.0), T(c[1], hit && sd.comp == 1 ? 1.0 : 0.0), T(c[2], hit && sd.comp == 2 ? 1.0 : 0.0));
      r.p = r.p + Bp[k] * cp; r.v = r.v + Bv[k] * cp;
    }
  }
  {      // SO3
    const T u = T(ub0, td / sp.dt_so3), u2 = u * u, u3 = u2 * u;
    const T di = T(1.0 / sp.dt_so3), s6 = T(1.0 / 6.0);
    T B[3], dB[3];
    B[0] = (T(5.0) + T(3.0) * u - T(3.0) * u2 + u3) * s6; B[1] = (T(1.0) + T(3.0) * u + T(3.0) * u2 - T(2.0) * u3) * s6; B[2] = u3 * s6;
    dB[0] = di * (T(3.0) - T(6.0) * u + T(3.0) * u2) * s6; dB[1] = di * (T(3.0) + T(6.0) * u - T(6.0) * u2) * s6; dB[2] = di * (T(3.0) * u2) * s6;
    const double* q0d = quats + (size_t)ib * kQuatStride;
    const bool hit0 = sd.spline == 2 && sd.knot == ib;
    TQ<T> q0; q0.x = T(q0d[0], hit0 && sd.comp == 0 ? 1.0 : 0.0); q0.y = T(q0d[1], hit0 && sd.comp == 1 ? 1.0 : 0.0);
    q0.z = T(q0d[2], hit0 && sd.comp == 2 ? 1.0 : 0.0); q0.w = T(q0d[3], hit0 && sd.comp == 3 ? 1.0 : 0.0);
    TQ<T> q = q0, parts[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) { parts[m].x = T(0.0); parts[m].y = T(0.0); parts[m].z = T(0.0); parts[m].w = T(1.0); }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int pidx = ib + 1 + j;                                         // pair (knot pidx - 1, knot pidx)
      const double* pr = so3pairs + (size_t)pidx * kSo3PairStride;
      const double* D = sd.spline != 2 ? nullptr : (sd.knot == pidx - 1 ? pr + kSo3PairDOff : (sd.knot == pidx ? pr + kSo3PairDOff + kSo3PairSide : nullptr));
      const TV3<T> om = tv3<T>(T(pr[0], D ? D[0 * 4 + sd.comp] : 0.0), T(pr[1], D ? D[1 * 4 + sd.comp] : 0.0), T(pr[2], D ? D[2 * 4 + sd.comp] : 0.0));
      const TQ<T> e = expq_t<T>(B[j] * om);
      q = tqmul(q, e);
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        if (m == j) { TQ<T> w; w.x = om.x * dB[j]; w.y = om.y * dB[j]; w.z = om.z * dB[j]; w.w = T(0.0); parts[m] = tqmul(parts[m], w); }
        parts[m] = tqmul(parts[m], e);
      }
    }
    TQ<T> sum; sum.x = parts[0].x + parts[1].x + parts[2].x; sum.y = parts[0].y + parts[1].y + parts[2].y;
    sum.z = parts[0].z + parts[1].z + parts[2].z; sum.w = parts[0].w + parts[1].w + parts[2].w;
    const TQ<T> dq = tqmul(q0, sum);
    const TQ<T> wq = tqmul(dq, tqconj(q));
    r.q = q; r.w = tv3<T>(T(2.0) * wq.x, T(2.0) * wq.y, T(2.0) * wq.z);
  }
  return r;
}
// observation span of one spline (newton_rscamera_measurement.h:210-236 with the time offset locked): first knot and number of knots
KB_HD int span_window_base(double t0, double dt, double obs_t0) { return knot_floor(sub_rn(obs_t0, 1e-3), t0, dt); }
KB_HD int span_window_size(double t0, double dt, double readout, double obs_t0) { return knot_floor(add_rn(add_rn(obs_t0, readout), 1e-3), t0, dt) + 4 - span_window_base(t0, dt, obs_t0); }
KB_HD int span_split_ndir(bool lifting, int Wa, int Wb) { return 28 + 3 * Wa + 4 * Wb + (lifting ? 2 : 1); }
KB_HD int span_split_row_len(bool lifting, int Wa, int Wb) { return (lifting ? 3 : 2) * (28 + 3 * Wa + 4 * Wb) + (lifting ? 6 : 2); }
// where direction `dir` lands in the packed row (residual row 0; the next residual rows are `stride` further on)
KB_HD int span_split_dir_offset(bool lifting, int Wa, int Wb, int dir, int& stride) {
  const int nres = lifting ? 3 : 2;
  if (dir < 12) { stride = 3; return nres * 3 * (dir / 3) + dir % 3; }
  if (dir < 28) { const int d = dir - 12; stride = 4; return nres * 12 + nres * 4 * (d / 4) + d % 4; }
  if (dir < 28 + 3 * Wa) { const int d = dir - 28; stride = 3; return nres * 28 + nres * 3 * (d / 3) + d % 3; }
  if (dir < 28 + 3 * Wa + 4 * Wb) { const int d = dir - 28 - 3 * Wa; stride = 4; return nres * (28 + 3 * Wa) + nres * 4 * (d / 4) + d % 4; }
  stride = 1;
  return nres * (28 + 3 * Wa + 4 * Wb) + nres * (dir - 28 - 3 * Wa - 4 * Wb);
}
// One direction of a NewtonRs (lifting = false) / LiftingRs row on a split trajectory.  rec: landmark record of k_landmark_ref_split.
struct SpanSplitRow { double y[2], dy[2], dvt; int iterations; };
// sensor_c >= 0 (with dir < 0, rec == nullptr): column sensor_c of the camera's relative pose (0..3 q_ct, 4..6 p_ct) -- the landmark X(t_ref) is then
// evaluated here on the dual number from (ref_uv, rho_v), because it depends on the camera pose too.
KB_HD int span_split_direction(bool lifting, const SplitConst& sp, const CameraConst& cam, const double* vecs, const double* quats, const double* so3pairs,
                               const double* rec, const double* obs_uv, double obs_t0, double ref_t0, double vt, int ka, int Wa, int kb, int Wb,
                               int dir, SpanSplitRow& out, int sensor_c = -1, const double* ref_uv = nullptr, double rho_v = 0.0) {
  typedef D1 T;
  Segment a0, a1, b0, b1;
  const int nsa = static_rs_segments_split(sp, cam, ref_t0, obs_t0, sp.t0_r3, sp.dt_r3, a0, a1);
  const int nsb = static_rs_segments_split(sp, cam, ref_t0, obs_t0, sp.t0_so3, sp.dt_so3, b0, b1);
  if (nsa == 0 || nsb == 0) return kStatusRange;
  const int d_obs_a = 28, d_obs_b = 28 + 3 * Wa, d_tail = 28 + 3 * Wa + 4 * Wb;
  const int vt_dir = lifting ? d_tail : -1, rho_dir = lifting ? d_tail + 1 : d_tail;
  SplitSeed sd; sd.spline = 0; sd.knot = -1; sd.comp = 0;
  TQ<T> qct; qct.x = T(cam.q_ct[0]); qct.y = T(cam.q_ct[1]); qct.z = T(cam.q_ct[2]); qct.w = T(cam.q_ct[3]);
  TV3<T> pct = tv3<T>(T(cam.p_ct[0]), T(cam.p_ct[1]), T(cam.p_ct[2]));
  TV3<T> X;
  T rho;
  if (sensor_c >= 0) {
    const SensorSeed<T> ss = camera_sensor_seed(cam, sensor_c);
    qct = ss.qct; pct = ss.pct; rho = T(rho_v);
    int ia, ib; double ua, ub;
    const double tr = static_rs_time(cam, ref_t0, ref_uv[1]);
    if (locate_in_segments(nsa, a0, a1, tr, sp.t0_r3, sp.dt_r3, ia, ua) < 0 || locate_in_segments(nsb, b0, b1, tr, sp.t0_so3, sp.dt_so3, ib, ub) < 0) return kStatusRange;
    if (ia < 0 || ia + 3 >= sp.n_r3 || ib < 0 || ib + 3 >= sp.n_so3) return kStatusRange;
    const TEval<T> er = split_eval_t<T>(sp, vecs, quats, so3pairs, ia, ua, ib, ub, 0.0, sd);
    const V3 yh = camera_unproject(cam, ref_uv[0], ref_uv[1]);
    X = tqrot(er.q, tqrot(tqconj(qct), tv3<T>(T(yh.x), T(yh.y), T(yh.z)) - rho * pct)) + rho * er.p;
  } else {
    X = tv3<T>(T(rec[0]), T(rec[1]), T(rec[2]));
    rho = T(rec[6]);
  }
  if (sensor_c >= 0) { /* knots and rho unseeded */ }
  else if (dir >= 0 && dir < 12) { const double v = rec[kRefSplitBp + dir / 3]; const int c = dir % 3; if (c == 0) X.x.d = v; else if (c == 1) X.y.d = v; else X.z.d = v; }
  else if (dir >= 12 && dir < 28) { const double* d = rec + kRefSplitDq + 12 * ((dir - 12) / 4) + (dir - 12) % 4; X.x.d = d[0]; X.y.d = d[4]; X.z.d = d[8]; }
  else if (dir >= d_obs_a && dir < d_obs_b) { sd.spline = 1; sd.knot = ka + (dir - d_obs_a) / 3; sd.comp = (dir - d_obs_a) % 3; }
  else if (dir >= d_obs_b && dir < d_tail) { sd.spline = 2; sd.knot = kb + (dir - d_obs_b) / 4; sd.comp = (dir - d_obs_b) % 4; }
  else if (dir == rho_dir) { X.x.d = rec[3]; X.y.d = rec[4]; X.z.d = rec[5]; rho.d = 1.0; }
  out.iterations = 0; out.dvt = 0.0;
  if (lifting) {
    const T t_obs = T(add_rn(add_rn(obs_t0, cam.time_offset), mul_rn(vt, cam.readout)), dir == vt_dir ? cam.readout : 0.0);
    out.dvt = dir == vt_dir ? 1.0 : 0.0;
    int ia, ib; double ua, ub;
    if (locate_in_segments(nsa, a0, a1, t_obs.a, sp.t0_r3, sp.dt_r3, ia, ua) < 0 || locate_in_segments(nsb, b0, b1, t_obs.a, sp.t0_so3, sp.dt_so3, ib, ub) < 0) return kStatusRange;
    if (ia < ka || ia + 4 > ka + Wa || ib < kb || ib + 4 > kb + Wb) return kStatusRange;
    const TEval<T> ev = split_eval_t<T>(sp, vecs, quats, so3pairs, ia, ua, ib, ub, t_obs.d, sd);
    const TV3<T> X_obs = tqrot(tqconj(ev.q), X - rho * ev.p);
    const TV3<T> X_cam = tqrot(qct, X_obs) + rho * pct;
    T y[2], dy[2];
    camera_project_t<T>(cam, X_cam, tv3<T>(T(0.0), T(0.0), T(0.0)), y, dy);
    out.y[0] = y[0].a; out.y[1] = y[1].a; out.dy[0] = y[0].d; out.dy[1] = y[1].d;
    return 0;
  }
  const double rows = (double)cam.rows;
  const double t0_obs = add_rn(obs_t0, cam.time_offset);
  T t_obs = T(static_rs_time(cam, obs_t0, obs_uv[1]));
  const double max_dt = 0.5 * cam.readout / rows, max_dt2 = max_dt * max_dt;
  const double min_bound = t0_obs, max_bound = add_rn(t0_obs, cam.readout);
  T y[2] = {T(0.0), T(0.0)};
  for (int iter = 0; iter < 5; ++iter) {
    int ia, ib; double ua, ub;
    if (locate_in_segments(nsa, a0, a1, t_obs.a, sp.t0_r3, sp.dt_r3, ia, ua) < 0 || locate_in_segments(nsb, b0, b1, t_obs.a, sp.t0_so3, sp.dt_so3, ib, ub) < 0) return kStatusRange;
    if (ia < ka || ia + 4 > ka + Wa || ib < kb || ib + 4 > kb + Wb) return kStatusRange;
    const TEval<T> ev = split_eval_t<T>(sp, vecs, quats, so3pairs, ia, ua, ib, ub, t_obs.d, sd);
    TQ<T> wq; wq.x = ev.w.x; wq.y = ev.w.y; wq.z = ev.w.z; wq.w = T(0.0);
    TQ<T> dq = tqmul(wq, ev.q); dq.x = T(0.5) * dq.x; dq.y = T(0.5) * dq.y; dq.z = T(0.5) * dq.z; dq.w = T(0.5) * dq.w;
    const TQ<T> dq_inv = tqconj(dq), q_inv = tqconj(ev.q);
    const TV3<T> s = X - rho * ev.p;
    const TV3<T> ds = (T(0.0) - rho) * ev.v;
    const TV3<T> X_obs = tqrot(q_inv, s);
    const TV3<T> X_cam = tqrot(qct, X_obs) + rho * pct;
    TQ<T> sq; sq.x = s.x; sq.y = s.y; sq.z = s.z; sq.w = T(0.0);
    TQ<T> dsq; dsq.x = ds.x; dsq.y = ds.y; dsq.z = ds.z; dsq.w = T(0.0);
    const TQ<T> a1q = tqmul(tqmul(dq_inv, sq), ev.q), a2q = tqmul(tqmul(q_inv, dsq), ev.q), a3q = tqmul(tqmul(q_inv, sq), dq);
    const TV3<T> dX_obs = tv3<T>(a1q.x + a2q.x + a3q.x, a1q.y + a2q.y + a3q.y, a1q.z + a2q.z + a3q.z);
    const TV3<T> dX_cam = tqrot(qct, dX_obs) + rho * pct;                    // sic (newton_rscamera_measurement.h:92)
    T dy[2];
    camera_project_t<T>(cam, X_cam, dX_cam, y, dy);
    const T f = y[1] - (T(rows) * (t_obs - T(t0_obs)) / T(cam.readout));
    const T df = dy[1] - T(rows / cam.readout);
    const T dt = f / df;
    t_obs = t_obs - dt;
    out.iterations = iter + 1;
    if (dt.a * dt.a < max_dt2) break;
    if (t_obs.a < min_bound) t_obs = T(min_bound);
    else if (t_obs.a > max_bound) t_obs = T(max_bound);
  }
  out.y[0] = y[0].a; out.y[1] = y[1].a; out.dy[0] = y[0].d; out.dy[1] = y[1].d;
  return 0;
}
// residual and one Jacobian column of a split span row (Huber-corrected like the SE3 rows), written at its place in the packed row
KB_HD int span_split_column(bool lifting, const SplitConst& sp, const CameraConst& cam, const double* vecs, const double* quats, const double* so3pairs,
                            const double* rec, const double* obs_uv, double obs_t0, double ref_t0, double vt, int ka, int Wa, int kb, int Wb,
                            double weight, double huber_c, int dir, double* r, double* Jrow) {
  SpanSplitRow o;
  const int st = span_split_direction(lifting, sp, cam, vecs, quats, so3pairs, rec, obs_uv, obs_t0, ref_t0, vt, ka, Wa, kb, Wb, dir, o);
  if (st != 0) return st;
  double j[3];
  if (lifting) {
    LiftingRow lo; lo.y[0] = o.y[0]; lo.y[1] = o.y[1]; lo.dy[0] = o.dy[0]; lo.dy[1] = o.dy[1]; lo.dvt = o.dvt;
    lifting_rs_finish(lo, cam, obs_uv, vt, weight, huber_c, r, j);
  } else {
    NewtonRow no; no.y[0] = o.y[0]; no.y[1] = o.y[1]; no.dy[0] = o.dy[0]; no.dy[1] = o.dy[1]; no.iterations = o.iterations;
    newton_rs_finish(no, obs_uv, weight, huber_c, r, j);
  }
  if (Jrow && dir >= 0) {
    int stride;
    const int off = span_split_dir_offset(lifting, Wa, Wb, dir, stride);
    for (int rr = 0; rr < (lifting ? 3 : 2); ++rr) Jrow[off + rr * stride] = j[rr];
  }
  return 0;
}

// Column c (0..6) of the sensor blocks of a split span row into Js_row = [q_ct (nres x 4) | p_ct (nres x 3) | time offset (nres, zero)]
KB_HD int span_split_sensor_column(bool lifting, const SplitConst& sp, const CameraConst& cam, const double* vecs, const double* quats, const double* so3pairs,
                                   const double* ref_uv, double ref_t0, double rho, const double* obs_uv, double obs_t0, double vt, int ka, int Wa, int kb,
                                   int Wb, double weight, double huber_c, int c, double* Js_row) {
  const int nres = lifting ? 3 : 2;
  SpanSplitRow o;
  const int st = span_split_direction(lifting, sp, cam, vecs, quats, so3pairs, nullptr, obs_uv, obs_t0, ref_t0, vt, ka, Wa, kb, Wb, -1, o, c, ref_uv, rho);
  if (st != 0) return st;
  double r[3], j[3];
  if (lifting) {
    LiftingRow lo; lo.y[0] = o.y[0]; lo.y[1] = o.y[1]; lo.dy[0] = o.dy[0]; lo.dy[1] = o.dy[1]; lo.dvt = 0.0;
    lifting_rs_finish(lo, cam, obs_uv, vt, weight, huber_c, r, j);
  } else {
    NewtonRow no; no.y[0] = o.y[0]; no.y[1] = o.y[1]; no.dy[0] = o.dy[0]; no.dy[1] = o.dy[1]; no.iterations = o.iterations;
    newton_rs_finish(no, obs_uv, weight, huber_c, r, j);
  }
  for (int rr = 0; rr < nres; ++rr) {
    if (c < 4) Js_row[4 * rr + c] = j[rr]; else Js_row[4 * nres + 3 * rr + (c - 4)] = j[rr];
    if (c == 0) Js_row[7 * nres + rr] = 0.0;
  }
  return 0;
}

}  // namespace kb

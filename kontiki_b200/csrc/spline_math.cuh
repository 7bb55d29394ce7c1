// kontiki_b200 -- per-measurement mathematics of the hot path, written once as __host__ __device__ inline
// functions: the CUDA kernels (kernels.cu) call them per thread; tests/host_check.cpp compiles the very same
// text for the host so that the formulas can be checked against the CPU oracle in a container without a GPU.
// The shipped library never runs them on the host.
//
// What the reference computes (all citations relative to /root/reference/cpplib/include/kontiki/):
//   trajectories/uniform_se3_spline_trajectory.h:101-194   cumulative SE(3) B-spline  P, P', P''
//   trajectories/uniform_se3_spline_trajectory.h:81-99     position / velocity / acceleration / orientation / omega_world
//   sensors/imu.h:47-59                                    gyroscope = q^-1 w_world,  accelerometer = q^-1 (a + g)
//   measurements/gyroscope_measurement.h:36-38, accelerometer_measurement.h:37-39      r = weight (y - y_hat)
//   measurements/static_rscamera_measurement.h:21-55,89-94  static rolling-shutter reprojection
//   sensors/pinhole_camera.h:47-67                          K X / unproject
// and differentiates with ceres::Jet over the 7 ambient doubles of every knot.  Here the same quantities are
// computed in closed form on the group (body twists, Ad/ad, right Jacobians of SE(3)); the Jacobian with respect to
// the knots is assembled by the chain rule through the hoisted knot-pair logs
//     omega_j = log(P_{i0+j-1}^-1 P_{i0+j}),   D_j = d omega_j / d(knot_{i0+j-1}, knot_{i0+j})   (6 x 14, ambient)
// which the prepass K0 evaluates once per evaluation point (pair_log below), plus the direct dependence on knot i0.
//
// Ambient-coordinate note.  The first knot P0 enters the reference formulas directly (P0.matrix(), P0 * A1), and the
// Eigen/Sophus polynomials R(q), q*v are not scale invariant, so d/d(q0) has a component along q0 itself.  For a unit
// quaternion the ambient 4-vector derivative splits exactly into  J_tangent * dtheta/dq + J_radial * q^T  with
// dtheta/dq = 2 [w I - hat(v) | -v]  (right perturbation q (x) exp(dtheta/2)) and  d/ds R((1+s) q)|0 = 2 (R - I).
// J_radial is derived per measurement type below ("radial" comments).
#pragma once
#include "dualnum.cuh"
#include "lie_math.cuh"

namespace kb {

// ---- device data layout ------------------------------------------------------------------------------------------
// knot record  : 8 doubles  [qx qy qz qw tx ty tz pad]            (64 B, any window start is 16-B aligned for TMA)
// pair record  : 104 doubles [omega(6) pad(2) Da(6x8) Db(6x8)]     (832 B); slot p holds the pair (knot p-1, knot p);
//                Da / Db = d omega / d knot p-1 / d knot p (6 x 7, rows padded to 8 so that every row is 16-B aligned)
constexpr int kKnotStride = 8;
constexpr int kPairStride = 104;
constexpr int kPairDOff = 8;
constexpr int kPairSide = 48;      // doubles between Da and Db
struct alignas(16) Dbl2 { double x, y; };
constexpr double kGravity = 9.80665;   // constants.h:13,24
constexpr double kSophusEps = 1e-10;   // Sophus::Constants<double>::epsilon()

// Index arithmetic must not be FMA-contracted: spline_base.h:148-152 and :380 round after
// every operation.  On the device the _rn intrinsics are never fused; the host check is built with -ffp-contract=off.
#if defined(__CUDA_ARCH__)
KB_HD double mul_rn(double a, double b) { return __dmul_rn(a, b); }
KB_HD double add_rn(double a, double b) { return __dadd_rn(a, b); }
KB_HD double sub_rn(double a, double b) { return __dsub_rn(a, b); }
KB_HD double div_rn(double a, double b) { return __ddiv_rn(a, b); }
#else
KB_HD double mul_rn(double a, double b) { return a * b; }
KB_HD double add_rn(double a, double b) { return a + b; }
KB_HD double sub_rn(double a, double b) { return a - b; }
KB_HD double div_rn(double a, double b) { return a / b; }
#endif

// spline_base.h:148-152: i = floor((t - t0) / dt) on the whole spline (used by the structure rule, :371-377)
KB_HD int knot_floor(double t, double t0, double dt) { return (int)floor(div_rn(sub_rn(t, t0), dt)); }

// One segment of the per-residual spline view: first knot `start`, `n` knots (spline_base.h:30-62, :378-383).
struct Segment { int start, n; };

// SplineEntity::AddToProblem for ONE span (IMU with locked time offset gives {t,t}: exactly 4 knots).
KB_HD void segments_one_span(double ta, double tb, double t0, double dt, Segment& s0) {
  const int i1 = knot_floor(ta, t0, dt), i2 = knot_floor(tb, t0, dt);
  s0.start = i1; s0.n = i2 + 4 - i1;
}
// ... and for two ordered spans (static RS camera, static_rscamera_measurement.h:137-166): the second span either
// opens a new segment or extends the first (spline_base.h:371-403).  Returns the number of segments.
KB_HD int segments_two_spans(double a1, double b1, double a2, double b2, double t0, double dt, Segment& s0, Segment& s1) {
  segments_one_span(a1, b1, t0, dt, s0);
  const int end0 = s0.start + s0.n - 1;
  int j1 = knot_floor(a2, t0, dt);
  const int j2 = knot_floor(b2, t0, dt);
  if (j1 > end0) { s1.start = j1; s1.n = j2 + 4 - j1; return 2; }
  j1 = end0 + 1;
  if (j2 + 4 > j1) s0.n += j2 + 4 - j1;
  s1.start = 0; s1.n = 0;
  return 1;
}
// SplineView::Evaluate + CalculateIndexAndInterpolationAmount inside one segment (spline_base.h:188-202, :148-152,
// uniform_se3_spline_trajectory.h:121-127).  The segment origin is t0 + dt*start rounded as in spline_base.h:380.
// Returns false where the reference throws std::range_error.
KB_HD bool segment_locate(const Segment& s, double t, double t0, double dt, int& i0, double& u) {
  if (s.n < 4) return false;
  const double tseg = add_rn(t0, mul_rn(dt, (double)s.start));
  const double tmax = add_rn(tseg, mul_rn((double)(s.n - 3), dt));
  if (!(t >= tseg && t < tmax)) return false;
  const double sc = div_rn(sub_rn(t, tseg), dt);
  const int il = (int)floor(sc);
  if (il < 0 || il > s.n - 4) return false;
  u = sub_rn(sc, (double)il);
  i0 = s.start + il;
  return true;
}

// Cumulative cubic B-spline basis and its time derivatives (spline_base.h:18-28 M_cumul,
// uniform_se3_spline_trajectory.h:129-146); index j-1 for j = 1..3 (B_0 == 1 is not stored).
struct Basis { double B[3], dB[3], d2B[3]; };
KB_HD Basis cumulative_basis(double u, double dt) {
  Basis b;
  const double u2 = u * u, u3 = u2 * u, di = 1.0 / dt, di2 = di * di;
  b.B[0] = (5.0 + 3.0 * u - 3.0 * u2 + u3) * (1.0 / 6.0);
  b.B[1] = (1.0 + 3.0 * u + 3.0 * u2 - 2.0 * u3) * (1.0 / 6.0);
  b.B[2] = u3 * (1.0 / 6.0);
  b.dB[0] = di * (3.0 - 6.0 * u + 3.0 * u2) * (1.0 / 6.0);
  b.dB[1] = di * (3.0 + 6.0 * u - 6.0 * u2) * (1.0 / 6.0);
  b.dB[2] = di * (3.0 * u2) * (1.0 / 6.0);
  b.d2B[0] = di2 * (u - 1.0);
  b.d2B[1] = di2 * (1.0 - 2.0 * u);
  b.d2B[2] = di2 * u;
  return b;
}

// =================================================================================================================
// K0: knot-pair log map, templated on the scalar so that one dual direction gives one column of D.
// Arithmetic follows what Sophus does at the reference's call site  ((Pa.inverse() * Pb).log(),
// uniform_se3_spline_trajectory.h:159-162): inverse() and operator* re-normalise the quaternion, q*v is Eigen's
// polynomial, SO3::log uses atan(n/w).  (Sophus is un-vendored; same assumptions as SURVEY.md Appendix B.)
// =================================================================================================================
template <class T> KB_HD void q_normalize(T& x, T& y, T& z, T& w) {
  const T inv = T(1.0) / t_sqrt(x * x + y * y + z * z + w * w);
  x = x * inv; y = y * inv; z = z * inv; w = w * inv; }
template <class T> KB_HD void q_rotate(T qx, T qy, T qz, T qw, T vx, T vy, T vz, T& ox, T& oy, T& oz) {
  T ux = qy * vz - qz * vy, uy = qz * vx - qx * vz, uz = qx * vy - qy * vx;
  ux = ux + ux; uy = uy + uy; uz = uz + uz;
  ox = vx + qw * ux + (qy * uz - qz * uy);
  oy = vy + qw * uy + (qz * ux - qx * uz);
  oz = vz + qw * uz + (qx * uy - qy * ux); }

template <class T> KB_HD void pair_log(const T* a, const T* b, T* om) {
  // Pa^-1 = (conj(qa) normalised, R^-1 (-ta))
  T ix = -a[0], iy = -a[1], iz = -a[2], iw = a[3];
  q_normalize(ix, iy, iz, iw);
  T itx, ity, itz;
  q_rotate(ix, iy, iz, iw, -a[4], -a[5], -a[6], itx, ity, itz);
  // Pa^-1 * Pb
  T qx = iw * b[0] + ix * b[3] + iy * b[2] - iz * b[1];
  T qy = iw * b[1] + iy * b[3] + iz * b[0] - ix * b[2];
  T qz = iw * b[2] + iz * b[3] + ix * b[1] - iy * b[0];
  T qw = iw * b[3] - ix * b[0] - iy * b[1] - iz * b[2];
  q_normalize(qx, qy, qz, qw);
  T rx, ry, rz;
  q_rotate(ix, iy, iz, iw, b[4], b[5], b[6], rx, ry, rz);
  const T tx = itx + rx, ty = ity + ry, tz = itz + rz;
  // SO3::log
  const T sqn = qx * qx + qy * qy + qz * qz;
  const double nv = sqrt(value(sqn));
  T f, theta;
  if (nv < kSophusEps) {
    f = T(2.0) / qw - T(2.0) * sqn / (qw * qw * qw);
    theta = T(value(f) * nv);
  } else {
    const T n = t_sqrt(sqn);
    if (fabs(value(qw)) < kSophusEps) f = (value(qw) > 0.0 ? T(3.14159265358979323846) : T(-3.14159265358979323846)) / n;
    else f = T(2.0) * t_atan(n / qw) / n;
    theta = f * n;
  }
  const T px = f * qx, py = f * qy, pz = f * qz;
  // SE3::log: upsilon = V^-1 t,  V^-1 = I - 1/2 hat(phi) + k hat(phi)^2
  T k;
  if (fabs(value(theta)) < kSophusEps) k = T(1.0 / 12.0);
  else { const T h = T(0.5) * theta; k = (T(1.0) - theta * t_cos(h) / (T(2.0) * t_sin(h))) / (theta * theta); }
  const T c1x = py * tz - pz * ty, c1y = pz * tx - px * tz, c1z = px * ty - py * tx;          // phi x t
  const T c2x = py * c1z - pz * c1y, c2y = pz * c1x - px * c1z, c2z = px * c1y - py * c1x;    // phi x (phi x t)
  om[0] = tx - T(0.5) * c1x + k * c2x;
  om[1] = ty - T(0.5) * c1y + k * c2y;
  om[2] = tz - T(0.5) * c1z + k * c2z;
  om[3] = px; om[4] = py; om[5] = pz;
}

// One (pair, direction) work item of K0.  dir in [0,14): derivative with respect to ambient scalar `dir` of
// [knot p-1 (7), knot p (7)]; dir == 14: values only.  knots: kKnotStride records; out: pair record `p`.
// KSTRIDE: doubles between knots in `knots` (kKnotStride records, or 7 for the caller's n x 7 array: the fused pack + prepass kernel).
template <int KSTRIDE = kKnotStride>
KB_HD void pair_prepass_item(const double* knots, int p, int dir, double* pairs) {
  const double* ka = knots + (size_t)(p - 1) * KSTRIDE;
  const double* kb_ = knots + (size_t)p * KSTRIDE;
  double* rec = pairs + (size_t)p * kPairStride;
  if (dir >= 14) {
    double om[6];
    pair_log<double>(ka, kb_, om);
    for (int i = 0; i < 6; ++i) rec[i] = om[i];
    rec[6] = 0.0; rec[7] = 0.0;
    return;
  }
  D1 a[7], b[7], om[6];
  for (int i = 0; i < 7; ++i) { a[i] = D1(ka[i], dir == i ? 1.0 : 0.0); b[i] = D1(kb_[i], dir == 7 + i ? 1.0 : 0.0); }
  pair_log<D1>(a, b, om);
  double* D = rec + kPairDOff + (dir >= 7 ? kPairSide : 0);
  const int c = dir >= 7 ? dir - 7 : dir;
  for (int i = 0; i < 6; ++i) D[i * 8 + c] = om[i].d;
  if (c == 6) for (int i = 0; i < 6; ++i) D[i * 8 + 7] = 0.0;
}

// =================================================================================================================
// Per-measurement pieces
// =================================================================================================================
// A_j = exp(B_j omega_j) = (E, a):  theta = B phi, rho = B upsilon,  E = Exp(theta),  V = J_l(theta),  a = V rho.
// J_r(theta) = V^T.
struct ExpPart { M3 E, V; V3 th, rho, a; double x; AngleCoefs c; };
KB_HD void exp_part(const double* om, double B, bool need_trans, bool need_q, ExpPart& e) {
  e.th = v3(B * om[3], B * om[4], B * om[5]);
  e.x = dot(e.th, e.th);
  e.c = angle_coefs(e.x, need_q);
  const M3 tt = outer(e.th, e.th), h = hat(e.th);
  const double de = 1.0 - e.c.cb * e.x, dv = 1.0 - e.c.cc * e.x;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    e.E.a[i] = e.c.sa * h.a[i] + e.c.cb * tt.a[i];
    e.V.a[i] = e.c.cb * h.a[i] + e.c.cc * tt.a[i];
  }
  e.E.a[0] += de; e.E.a[4] += de; e.E.a[8] += de;
  e.V.a[0] += dv; e.V.a[4] += dv; e.V.a[8] += dv;
  if (need_trans) { e.rho = v3(B * om[0], B * om[1], B * om[2]); e.a = e.V * e.rho; }
  else { e.rho = v3(0, 0, 0); e.a = v3(0, 0, 0); }
}

// An N x 6 adjoint block  [U | W]  acting on a twist [upsilon; phi]  (N = rows of the residual: 3 IMU, 2 camera).
template <int N> struct G6 { Mr<N> U, W; };
// G * ad(x),  ad([xu; xw]) = [[hat xw, hat xu], [0, hat xw]]
template <int N> KB_HD G6<N> mul_ad(const G6<N>& g, V3 xu, V3 xw) {
  G6<N> r; r.U = rmul_hat(g.U, xw); r.W = radd(rmul_hat(g.U, xu), rmul_hat(g.W, xw)); return r; }
// G * Ad(A^-1),  Ad(A^-1) = [[E^T, -E^T hat(a)], [0, E^T]]
template <int N> KB_HD G6<N> mul_Adinv(const G6<N>& g, const M3& E, V3 a) {
  G6<N> r; r.U = rmul_nt(g.U, E); r.W = rsub(rmul_nt(g.W, E), rmul_hat(r.U, a)); return r; }
// G * B * Jr6(B omega),  Jr6 = [[Jr, Qr], [0, Jr]],  Jr = V^T
template <int N> KB_HD G6<N> mul_Jr6(const G6<N>& g, const ExpPart& e, double B) {
  const M3 Qr = se3_q_block(e.rho, e.th, e.c, e.x, -1.0);
  G6<N> r; r.U = rscale(B, rmul_nt(g.U, e.V)); r.W = rscale(B, radd(rmul(g.U, Qr), rmul_nt(g.W, e.V))); return r; }
template <int N> KB_HD G6<N> gadd(const G6<N>& a, const G6<N>& b) { G6<N> r; r.U = radd(a.U, b.U); r.W = radd(a.W, b.W); return r; }
template <int N> KB_HD G6<N> gscale(double s, const G6<N>& a) { G6<N> r; r.U = rscale(s, a.U); r.W = rscale(s, a.W); return r; }

// J_block(N x 7) (+)= scale * [G.U | G.W](N x 6) * D,   D = one side (6 x 8 padded) of a pair record, 16-B aligned.
// The rotation part of log(Pa^-1 Pb) does not depend on the translations: rows 3..5 of D are zero in columns 4..6 and are skipped.
template <int N, bool ACC>
KB_HD void contract_pair(double* J, const G6<N>& g, const double* D, double scale) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {            // columns 0..3, then 4..6 (+ pad): N x 4 accumulators at a time
    double acc[N][4];
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
#pragma unroll
    for (int m = 0; m < (h == 0 ? 6 : 3); ++m) {       // d phi / d t == 0: the translation columns only see the upsilon rows
      const Dbl2* d2 = reinterpret_cast<const Dbl2*>(D + m * 8 + 4 * h);
      const Dbl2 v0 = d2[0], v1 = d2[1];
      const double d[4] = {v0.x, v0.y, v1.x, v1.y};
#pragma unroll
      for (int r = 0; r < N; ++r) {
        const double gm = m < 3 ? g.U.a[3 * r + m] : g.W.a[3 * r + m - 3];
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] += gm * d[c];
      }
    }
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
      for (int c = 0; c < (h == 0 ? 4 : 3); ++c) {
        if (ACC) J[r * 7 + 4 * h + c] += scale * acc[r][c]; else J[r * 7 + 4 * h + c] = scale * acc[r][c];
      }
  }
}
// rotation-only version: G (N x 3) * D[3:6, 0:4]; translation columns of the phi rows are identically zero
template <int N, bool ACC>
KB_HD void contract_pair_rot(double* J, const Mr<N>& G, const double* D, double scale) {
  double acc[N][4];
#pragma unroll
  for (int r = 0; r < N; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    const Dbl2* d2 = reinterpret_cast<const Dbl2*>(D + (3 + m) * 8);
    const Dbl2 v0 = d2[0], v1 = d2[1];
    const double d[4] = {v0.x, v0.y, v1.x, v1.y};
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] += G.a[3 * r + m] * d[c];
  }
#pragma unroll
  for (int r = 0; r < N; ++r) {
#pragma unroll
    for (int c = 0; c < 4; ++c) { if (ACC) J[r * 7 + c] += scale * acc[r][c]; else J[r * 7 + c] = scale * acc[r][c]; }
    if (!ACC) { J[r * 7 + 4] = 0.0; J[r * 7 + 5] = 0.0; J[r * 7 + 6] = 0.0; }
  }
}
// quaternion block of knot i0: J[r, 0:4] += scale * (g_theta[r] * dtheta/dq + g_rad[r] * q^T)
template <int N>
KB_HD void add_q0_block(double* J, const Mr<N>& Gth, const double* grad, const double* q, double scale) {
  const V3 v = v3(q[0], q[1], q[2]); const double w = q[3];
#pragma unroll
  for (int r = 0; r < N; ++r) {
    const V3 g = rrow(Gth, r);
    const V3 t = 2.0 * (w * g - cross(g, v));
    J[r * 7 + 0] += scale * (t.x + grad[r] * q[0]);
    J[r * 7 + 1] += scale * (t.y + grad[r] * q[1]);
    J[r * 7 + 2] += scale * (t.z + grad[r] * q[2]);
    J[r * 7 + 3] += scale * (-2.0 * dot(g, v) + grad[r] * q[3]);
  }
}

// ---- gyroscope on SE3 (imu.h:47-52, gyroscope_measurement.h:36-38) -----------------------------------------------
// Body angular velocity of the cumulative spline:  w_b = dB3 phi3 + E3^T (dB2 phi2 + E2^T dB1 phi1), which equals
// q^-1 * vee(R' R^T) of uniform_se3_spline_trajectory.h:93-96.  J: [4 knots][3][7].
KB_HD void gyro_se3(const double* knot0, const double* p1, const double* p2, const double* p3, const Basis& bs, double weight,
                    const double* y, double* r, double* J) {
  ExpPart e1, e2, e3;
  exp_part(p1, bs.B[0], false, false, e1);
  exp_part(p2, bs.B[1], false, false, e2);
  exp_part(p3, bs.B[2], false, false, e3);
  const V3 f1 = v3(p1[3], p1[4], p1[5]), f2 = v3(p2[3], p2[4], p2[5]), f3 = v3(p3[3], p3[4], p3[5]);
  const V3 y2 = mul_t(e2.E, bs.dB[0] * f1);
  const V3 s2 = y2 + bs.dB[1] * f2;
  const V3 y3 = mul_t(e3.E, s2);
  const V3 wb = y3 + bs.dB[2] * f3;
  r[0] = weight * (y[0] - wb.x); r[1] = weight * (y[1] - wb.y); r[2] = weight * (y[2] - wb.z);
  // d w_b / d phi_j
  M3 G3 = bs.B[2] * hat_mul(y3, transpose(e3.V)); G3.a[0] += bs.dB[2]; G3.a[4] += bs.dB[2]; G3.a[8] += bs.dB[2];
  M3 T2 = bs.B[1] * hat_mul(y2, transpose(e2.V)); T2.a[0] += bs.dB[1]; T2.a[4] += bs.dB[1]; T2.a[8] += bs.dB[1];
  const M3 G2 = mul_tn(e3.E, T2);
  const M3 G1 = bs.dB[0] * mul_tn(e3.E, transpose(e2.E));
  const double sc = -weight;
  contract_pair_rot<3, false>(J + 0, G1, p1 + kPairDOff, sc);
  contract_pair_rot<3, false>(J + 21, G1, p1 + kPairDOff + kPairSide, sc);
  contract_pair_rot<3, true>(J + 21, G2, p2 + kPairDOff, sc);
  contract_pair_rot<3, false>(J + 42, G2, p2 + kPairDOff + kPairSide, sc);
  contract_pair_rot<3, true>(J + 42, G3, p3 + kPairDOff, sc);
  contract_pair_rot<3, false>(J + 63, G3, p3 + kPairDOff + kPairSide, sc);
  // radial: P' = R(q0 raw) * M1 (uniform_se3_spline_trajectory.h:178-183) => d(R' R^T)/ds = 2 (I - R0^T) hat(w_world)
  const M3 R0 = quat_to_rot(knot0[0], knot0[1], knot0[2], knot0[3]);
  const M3 R = R0 * (e1.E * (e2.E * e3.E));
  const V3 ww = R * wb;
  const M3 Mm = mul_tn(R0, hat(ww));
  const V3 dw = 2.0 * ww - v3(Mm.a[7] - Mm.a[5], Mm.a[2] - Mm.a[6], Mm.a[3] - Mm.a[1]);
  const V3 rad = mul_t(R, dw);
  const double radv[3] = {rad.x, rad.y, rad.z};
  add_q0_block<3>(J, m3_zero(), radv, knot0, sc);
}

// ---- accelerometer on SE3 (imu.h:55-59, accelerometer_measurement.h:37-39) ---------------------------------------
// With body twist xi_b (P' = P xi_b^) and its time derivative:  R^T p'' = w_b x v_b + v_b', so
//   accel = w_b x v_b + v_b' + R^T g,    g = (0, 0, -9.80665).
// Recursion (s_j = Ad(A_j^-1) s_{j-1} + dB_j omega_j):
//   s1 = dB1 w1, s1' = d2B1 w1;  y_j = Ad(A_j^-1) s_{j-1},  z_j = Ad(A_j^-1) s'_{j-1},
//   s_j = y_j + dB_j w_j,  s'_j = z_j - dB_j ad(w_j) y_j + d2B_j w_j.
// compat_zero_dB: the reference's Jet path leaves dB == 0 for the accelerometer's flags;
// passing a Basis with dB = 0 reproduces it (P'' = P0 sum A''_j ... with A'_j = 0).
// Reverse sweep carries 3x6 adjoints.  J: [4 knots][3][7].
// PARK: the body twists of the forward pass that the reverse sweep needs again (y_j, z_j: 24 doubles) are parked in `scratch` (the kernels
// pass spare shared memory behind the row) instead of staying live in registers across the whole sweep; false: a local array (host harness).
template <bool PARK = false>
KB_HD void accel_se3(const double* knot0, const double* p1, const double* p2, const double* p3, const Basis& bs, double weight,
                     const double* y, double* r, double* J, double* scratch = nullptr) {
  double local_park[PARK ? 1 : 24];
  double* park = PARK ? scratch : local_park;
  const V3 u1 = v3(p1[0], p1[1], p1[2]), f1 = v3(p1[3], p1[4], p1[5]);
  const V3 u2 = v3(p2[0], p2[1], p2[2]), f2 = v3(p2[3], p2[4], p2[5]);
  const V3 u3 = v3(p3[0], p3[1], p3[2]), f3 = v3(p3[3], p3[4], p3[5]);
  ExpPart e;
  // forward (exp parts are rebuilt in the reverse sweep instead of being kept alive)
  const V3 s1u = bs.dB[0] * u1, s1w = bs.dB[0] * f1, d1u = bs.d2B[0] * u1, d1w = bs.d2B[0] * f1;
  exp_part(p2, bs.B[1], true, false, e);
  const M3 E2 = e.E;
  const V3 y2w = mul_t(e.E, s1w), y2u = mul_t(e.E, s1u - cross(e.a, s1w));
  const V3 z2w = mul_t(e.E, d1w), z2u = mul_t(e.E, d1u - cross(e.a, d1w));
  const V3 s2u = y2u + bs.dB[1] * u2, s2w = y2w + bs.dB[1] * f2;
  const V3 d2u = z2u - bs.dB[1] * (cross(f2, y2u) + cross(u2, y2w)) + bs.d2B[1] * u2;
  const V3 d2w = z2w - bs.dB[1] * cross(f2, y2w) + bs.d2B[1] * f2;
  exp_part(p3, bs.B[2], true, false, e);
  {
    park[0] = y2u.x; park[1] = y2u.y; park[2] = y2u.z; park[3] = y2w.x; park[4] = y2w.y; park[5] = y2w.z;
    park[6] = z2u.x; park[7] = z2u.y; park[8] = z2u.z; park[9] = z2w.x; park[10] = z2w.y; park[11] = z2w.z;
  }
  V3 y3w = mul_t(e.E, s2w), y3u = mul_t(e.E, s2u - cross(e.a, s2w));
  V3 z3w = mul_t(e.E, d2w), z3u = mul_t(e.E, d2u - cross(e.a, d2w));
  const V3 vb = y3u + bs.dB[2] * u3, wb = y3w + bs.dB[2] * f3;
  const V3 dvb = z3u - bs.dB[2] * (cross(f3, y3u) + cross(u3, y3w)) + bs.d2B[2] * u3;
  {
    park[12] = y3u.x; park[13] = y3u.y; park[14] = y3u.z; park[15] = y3w.x; park[16] = y3w.y; park[17] = y3w.z;
    park[18] = z3u.x; park[19] = z3u.y; park[20] = z3u.z; park[21] = z3w.x; park[22] = z3w.y; park[23] = z3w.z;
  }
  const V3 fb = cross(wb, vb) + dvb;
  M3 T = E2 * e.E;                                        // E2 E3
  exp_part(p1, bs.B[0], false, false, e);
  T = e.E * T;                                            // E1 E2 E3
  const M3 R0 = quat_to_rot(knot0[0], knot0[1], knot0[2], knot0[3]);
  const V3 gb = mul_t(T, mul_t(R0, v3(0.0, 0.0, -kGravity)));     // R^T g
  const V3 acc = fb + gb;
  r[0] = weight * (y[0] - acc.x); r[1] = weight * (y[1] - acc.y); r[2] = weight * (y[2] - acc.z);
  // knot i0 directly: tangent through R^T g (d(R^T g) = hat(R^T g) dtheta_body, dtheta_body = T0^T d),
  // radial through P'' = R(q0 raw) M2 (uniform_se3_spline_trajectory.h:187-190): 2 (f_b - R^T T0 f_b)
  const double sc = -weight;
  const M3 hgb = hat(gb);
  const V3 Tf = T * fb;
  const V3 rad = 2.0 * (fb - mul_t(T, mul_t(R0, Tf)));
  // reverse: 3x6 adjoints of (s_j, s_j'), from the last factor to the first; T = E_{j+1}..E_3 rebuilt on the way.
  // Register pressure: five 3x6 blocks (gs, gd, gy, gj, t) plus an exp part were live across the Q block and the two contractions of a level
  // (988 B of spill per thread in round 1).  The adjoints the NEXT level needs are formed as soon as their inputs exist and PARKED in the
  // part of the output row that is not written yet (J[0..36) during level 3, J[0..18) during level 2 -- the row is built back to front), so
  // that only gj and t are live across the expensive part.
  KB_SEQ();
  G6<3> gs, gd, gj, t;
  gs.U = hat(wb); gs.W = (-1.0) * hat(vb);
  gd.U = m3_identity(); gd.W = m3_zero();
  {
    exp_part(p3, bs.B[2], true, true, e);
    y3u = v3(park[12], park[13], park[14]); y3w = v3(park[15], park[16], park[17]); z3u = v3(park[18], park[19], park[20]); z3w = v3(park[21], park[22], park[23]);
    const G6<3> gy = gadd(gs, gscale(-bs.dB[2], mul_ad(gd, u3, f3)));
    gj = gadd(gadd(gscale(bs.dB[2], gs), gscale(bs.d2B[2], gd)), gscale(bs.dB[2], mul_ad(gd, y3u, y3w)));
    t = gadd(mul_ad(gy, y3u, y3w), mul_ad(gd, z3u, z3w));
    t.W = t.W + hgb;                 // F3^T F3 = I
    gs = mul_Adinv(gy, e.E, e.a);
    gd = mul_Adinv(gd, e.E, e.a);
#pragma unroll
    for (int i = 0; i < 9; ++i) { J[i] = gs.U.a[i]; J[9 + i] = gs.W.a[i]; J[18 + i] = gd.U.a[i]; J[27 + i] = gd.W.a[i]; }
    gj = gadd(gj, mul_Jr6(t, e, bs.B[2]));
    contract_pair<3, false>(J + 63, gj, p3 + kPairDOff + kPairSide, sc);
    contract_pair<3, false>(J + 42, gj, p3 + kPairDOff, sc);
    T = e.E;
  }
  KB_SEQ();
  {
    exp_part(p2, bs.B[1], true, true, e);
#pragma unroll
    for (int i = 0; i < 9; ++i) { gs.U.a[i] = J[i]; gs.W.a[i] = J[9 + i]; gd.U.a[i] = J[18 + i]; gd.W.a[i] = J[27 + i]; }
    const V3 y2u = v3(park[0], park[1], park[2]), y2w = v3(park[3], park[4], park[5]), z2u = v3(park[6], park[7], park[8]), z2w = v3(park[9], park[10], park[11]);
    const G6<3> gy = gadd(gs, gscale(-bs.dB[1], mul_ad(gd, u2, f2)));
    gj = gadd(gadd(gscale(bs.dB[1], gs), gscale(bs.d2B[1], gd)), gscale(bs.dB[1], mul_ad(gd, y2u, y2w)));
    t = gadd(mul_ad(gy, y2u, y2w), mul_ad(gd, z2u, z2w));
    t.W = t.W + mul_nt(hgb, T);      // F3^T F2 = E3^T
    {                                // level 1 only needs dB1 gs' + d2B1 gd': form it now, park 18 doubles
      const G6<3> g1 = gadd(gscale(bs.dB[0], mul_Adinv(gy, e.E, e.a)), gscale(bs.d2B[0], mul_Adinv(gd, e.E, e.a)));
#pragma unroll
      for (int i = 0; i < 9; ++i) { J[i] = g1.U.a[i]; J[9 + i] = g1.W.a[i]; }
    }
    gj = gadd(gj, mul_Jr6(t, e, bs.B[1]));
    contract_pair<3, true>(J + 42, gj, p2 + kPairDOff + kPairSide, sc);
    contract_pair<3, false>(J + 21, gj, p2 + kPairDOff, sc);
    T = e.E * T;
  }
  KB_SEQ();
  {
    exp_part(p1, bs.B[0], false, false, e);
#pragma unroll
    for (int i = 0; i < 9; ++i) { gj.U.a[i] = J[i]; gj.W.a[i] = J[9 + i]; }
    // rotation of A1 only enters through R^T g:  F3^T F1 = (E2 E3)^T
    gj.W = gj.W + bs.B[0] * mul_nt(mul_nt(hgb, T), e.V);
    contract_pair<3, true>(J + 21, gj, p1 + kPairDOff + kPairSide, sc);
    contract_pair<3, false>(J + 0, gj, p1 + kPairDOff, sc);
    T = e.E * T;
  }
  const double radv[3] = {rad.x, rad.y, rad.z};
  add_q0_block<3>(J, mul_nt(hgb, T), radv, knot0, sc);
}

// ---- pose (position + orientation) of the cumulative spline and its reverse sweep ---------------------------------
//   R = R0 E1 E2 E3,  p = t0 + R0 c1,  c1 = a1 + E1 c2,  c2 = a2 + E2 a3
// Both sweeps run from the last factor to the first and carry only the body-frame tail T = E_{j+1} ... E_3 and
// c = c_{j+1}; the exp parts are rebuilt per level in the reverse sweep, which is far cheaper than keeping
// 3 x 33 doubles (plus the partial products) alive in registers between the sweeps.
struct Pose { M3 R; V3 p; };
KB_HD void pose_forward(const double* knot0, const double* p1, const double* p2, const double* p3, const Basis& bs, Pose& P) {
  ExpPart e;
  exp_part(p3, bs.B[2], true, false, e);
  M3 T = e.E; V3 c = e.a;
  exp_part(p2, bs.B[1], true, false, e);
  c = e.a + e.E * c; T = e.E * T;
  exp_part(p1, bs.B[0], true, false, e);
  c = e.a + e.E * c; T = e.E * T;
  const M3 R0 = quat_to_rot(knot0[0], knot0[1], knot0[2], knot0[3]);
  P.R = R0 * T;
  P.p = v3(knot0[4], knot0[5], knot0[6]) + R0 * c;
}
// Given N row-adjoints Gp with respect to p (world, additive; GpR = Gp * R is passed too because the callers have it
// for free) and Gth to a body-frame rotation perturbation R <- R Exp(d), writes scale * d(row)/d(knots i0..i0+3)
// into J ([4][N][7]).  With F_j = R0 E1..E_j = R T_j^T:
//   eps_j = B_j Jr6(B_j w_j) d(w_j):  d/d(eps_rho_j) = Gp F_j = GpR T_j^T,  d/d(eps_theta_j) = Gth T_j^T - (GpR T_j^T) hat(c_{j+1})
// PARK_GP: Gp is not needed until the very end (knot i0's translation columns and its radial term), and N x 3 doubles held across the
// three levels of the sweep are spilled in the 255-register kernels -- a local-memory reload there is an L2 round trip.  The caller parks
// Gp in the translation columns of block 0 of J (J[7 r + 4 + c], in shared memory; nothing writes them before the last contraction) and
// the sweep reads it back where it needs it.
template <int N, bool PARK_GP = false>
KB_HD void pose_backward(const double* knot0, const double* p1, const double* p2, const double* p3, const Basis& bs,
                         const Mr<N>& Gp_in, const Mr<N>& GpR_in, const Mr<N>& Gth, double scale, double* J) {
  static_assert(!PARK_GP || N == 2, "parking layout: block 0 of a 2-row Jacobian");
  // parked (N == 2): Gp at J[4..6], J[11..13];  GpR at J[0..2], J[7..9]
  auto load_GpR = [&]() -> Mr<N> {
    if (!PARK_GP) return GpR_in;
    Mr<N> m;
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) m.a[3 * r + cc] = J[r * 7 + cc];
    return m;
  };
  G6<N> g, t;
  ExpPart e;
  exp_part(p3, bs.B[2], true, true, e);
  t.U = load_GpR(); t.W = Gth;
  g = mul_Jr6(t, e, bs.B[2]);
  contract_pair<N, false>(J + 3 * N * 7, g, p3 + kPairDOff + kPairSide, scale);
  contract_pair<N, false>(J + 2 * N * 7, g, p3 + kPairDOff, scale);
  M3 T = e.E; V3 c = e.a;
  exp_part(p2, bs.B[1], true, true, e);
  t.U = rmul_nt(load_GpR(), T); t.W = rsub(rmul_nt(Gth, T), rmul_hat(t.U, c));
  g = mul_Jr6(t, e, bs.B[1]);
  contract_pair<N, true>(J + 2 * N * 7, g, p2 + kPairDOff + kPairSide, scale);
  contract_pair<N, false>(J + 1 * N * 7, g, p2 + kPairDOff, scale);
  c = e.a + e.E * c; T = e.E * T;
  exp_part(p1, bs.B[0], true, true, e);
  t.U = rmul_nt(load_GpR(), T); t.W = rsub(rmul_nt(Gth, T), rmul_hat(t.U, c));
  g = mul_Jr6(t, e, bs.B[0]);
  contract_pair<N, true>(J + 1 * N * 7, g, p1 + kPairDOff + kPairSide, scale);
  c = e.a + e.E * c; T = e.E * T;
  // knot i0 directly: t0 additive; R0 <- R0 Exp(d): dp = -R0 hat(c1) d, dtheta_body = T0^T d;
  // radial: P.t = t0 + q0 * a1 + ... with Eigen's polynomial q*v  =>  dp/ds = 2 (R0 - I) a1
  const Mr<N> GpR0 = rmul_nt(load_GpR(), T);
  const Mr<N> Gth0 = rsub(rmul_nt(Gth, T), rmul_hat(GpR0, c));
  Mr<N> Gp;
#pragma unroll
  for (int r = 0; r < N; ++r)
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) Gp.a[3 * r + cc] = PARK_GP ? J[r * 7 + 4 + cc] : Gp_in.a[3 * r + cc];
  contract_pair<N, false>(J + 0 * N * 7, g, p1 + kPairDOff, scale);      // overwrites the parking space
  const M3 R0 = quat_to_rot(knot0[0], knot0[1], knot0[2], knot0[3]);
  const V3 dps = 2.0 * (R0 * e.a - e.a);
  double grad[N];
#pragma unroll
  for (int r = 0; r < N; ++r) grad[r] = dot(rrow(Gp, r), dps);
  add_q0_block<N>(J, Gth0, grad, knot0, scale);
#pragma unroll
  for (int r = 0; r < N; ++r) { J[r * 7 + 4] += scale * Gp.a[3 * r]; J[r * 7 + 5] += scale * Gp.a[3 * r + 1]; J[r * 7 + 6] += scale * Gp.a[3 * r + 2]; }
}

// ceres::HuberLoss(a) + ceres::internal::Corrector (un-vendored Ceres 1.x): scale factors for one
// residual block with squared norm s.  Ceres applies them to r and J after Evaluate.
// Cold branches (rarely taken, heavy in libm calls) are kept OUT of the straight-line code of the row kernels: inlining them costs
// instruction-cache footprint and branch-merge moves in kernels that are bound by issue slots (profiles/README.md, r1i).
#if defined(__CUDACC__)
#define KB_COLD __host__ __device__ __noinline__
#else
#define KB_COLD inline
#endif
struct HuberScale { double sqrt_rho1, residual_scaling, alpha_sq_norm, rho0; };
// The linear region (|r| > a) is out of line on the device (KB_COLD): two square roots and three divisions with their slow-path
// subroutines do not belong in the straight-line code of every row; in the quadratic region rho' = 1, rho'' = 0 and every factor is
// exactly 1 or 0.
// One factor per call, returned in registers (which: 0 sqrt_rho1, 1 residual_scaling, 2 alpha_sq_norm, 3 rho0): see angle_coef_large.
KB_COLD double huber_linear_region(double a, double s, int which) {
  const double b = a * a;
  const double rr = sqrt(s);
  if (which == 3) return 2.0 * a * rr - b;
  const double rho1 = fmax(2.2250738585072014e-308, a / rr), rho2 = -rho1 / (2.0 * s);
  const double sqrt_rho1 = sqrt(rho1);
  if (which == 0) return sqrt_rho1;
  if (s == 0.0 || rho2 <= 0.0) return which == 1 ? sqrt_rho1 : 0.0;
  const double Dd = 1.0 + 2.0 * s * rho2 / rho1; const double alpha = 1.0 - sqrt(Dd);
  return which == 1 ? sqrt_rho1 / (1.0 - alpha) : alpha / s;
}
KB_HD HuberScale huber_scale(double a, double s) {
  HuberScale h; h.rho0 = s; h.sqrt_rho1 = 1.0; h.residual_scaling = 1.0; h.alpha_sq_norm = 0.0;
  if (s > a * a) {
    h.sqrt_rho1 = huber_linear_region(a, s, 0); h.residual_scaling = huber_linear_region(a, s, 1);
    h.alpha_sq_norm = huber_linear_region(a, s, 2); h.rho0 = huber_linear_region(a, s, 3);
  }
  return h;
}

// ---- static rolling-shutter camera on SE3 (static_rscamera_measurement.h:21-55, :89-94) ---------------------------
struct CameraConst {
  double K[9], Kinv[9];     // pinhole_camera.h:25 (meta, not optimised); Kinv by cofactors once instead of per call (:63-67)
  double q_ct[4], p_ct[3];  // sensors.h:36-57 relative pose (x,y,z,w)
  double Rct[9];            // R(q_ct) (Eigen toRotationMatrix), once per group instead of once per row; camera_set_pose() fills it
  double time_offset, row_delta;   // row_delta = readout / rows (static_rscamera_measurement.h:30)
  double readout, max_time_offset;
  int time_offset_locked;
  int rows;                 // image rows (camera.h:24-28)
  int model;                // 0 PinholeCamera (pinhole_camera.h), 1 AtanCamera (atan_camera.h)
  double wc[2], gamma;      // AtanCamera distortion centre and parameter (atan_camera.h:20-22)
};
KB_HD M3 load_m3(const double* a) { M3 m;
#pragma unroll
  for (int i = 0; i < 9; ++i) m.a[i] = a[i]; return m; }
KB_HD void camera_set_pose(CameraConst& c, const double* q_ct, const double* p_ct) {
  for (int i = 0; i < 4; ++i) c.q_ct[i] = q_ct[i];
  for (int i = 0; i < 3; ++i) c.p_ct[i] = p_ct[i];
  const M3 R = quat_to_rot(q_ct[0], q_ct[1], q_ct[2], q_ct[3]);
  for (int i = 0; i < 9; ++i) c.Rct[i] = R.a[i];
}

// CameraView::Unproject: pinhole K^-1 [u v 1] (pinhole_camera.h:63-67); atan additionally undoes the distortion
// L = phn.xy - wc, r = sqrt(|L|^2 + eps), f = tan(r gamma) / gamma, Y = [wc + f L / r, 1] (atan_camera.h:92-103).
// The AtanCamera branches are KB_COLD with scalar arguments (a struct by reference would pin the camera constants to local memory).
KB_COLD void camera_unproject_atan(double wc0, double wc1, double gamma, double px, double py, double* out) {
  const double L0 = px - wc0, L1 = py - wc1;
  const double r = sqrt((L0 * L0 + L1 * L1) + 1e-32);
  const double f = tan(r * gamma) / gamma;
  out[0] = wc0 + f * L0 / r; out[1] = wc1 + f * L1 / r;
}
KB_HD V3 camera_unproject(const CameraConst& cam, double u, double v) {
  const V3 ph = load_m3(cam.Kinv) * v3(u, v, 1.0);
  if (cam.model == 0) return ph;
  double o[2];
  camera_unproject_atan(cam.wc[0], cam.wc[1], cam.gamma, ph.x, ph.y, o);
  return v3(o[0], o[1], 1.0);
}
// CameraView::Project (sensors/camera.h:59-63) and its 2 x 3 Jacobian d y / d X.
//   pinhole (pinhole_camera.h:47-51): p = K X, y = p.xy / p.z
//   atan (atan_camera.h:54-75): A = X.xy / (X.z + eps), L = A - wc, r = sqrt(|L|^2 + eps), f = atan(r gamma) / gamma,
//     g = L / r, Y = [wc + f g, 1], y = (K Y).xy;   dY/dA = f' g g^T + (f / r)(I - g g^T), f' = 1 / (1 + gamma^2 r^2)
// out: y0 y1 | J (2 x 3 row-major)
KB_COLD void camera_project_jac_atan(double k0, double k1, double k2, double k3, double k4, double k5, double wc0, double wc1, double gamma,
                                     double Xx, double Xy, double Xz, double* out) {
  const double eps = 1e-32;
  const double iz = 1.0 / (Xz + eps);
  const double A0 = Xx * iz, A1 = Xy * iz;
  const double L0 = A0 - wc0, L1 = A1 - wc1;
  const double r = sqrt((L0 * L0 + L1 * L1) + eps), ir = 1.0 / r;
  const double f = atan(r * gamma) / gamma;
  const double g0 = L0 * ir, g1 = L1 * ir;
  const double Y0 = wc0 + f * g0, Y1 = wc1 + f * g1;
  out[0] = k0 * Y0 + k1 * Y1 + k2;
  out[1] = k3 * Y0 + k4 * Y1 + k5;
  const double fr = f * ir, dd = 1.0 / (1.0 + gamma * gamma * r * r) - fr;
  const double m00 = fr + dd * g0 * g0, m01 = dd * g0 * g1, m11 = fr + dd * g1 * g1;
  // dA/dX = iz [I2 | -A]
  const double d0[3] = {iz * m00, iz * m01, -iz * (m00 * A0 + m01 * A1)};
  const double d1[3] = {iz * m01, iz * m11, -iz * (m01 * A0 + m11 * A1)};
  for (int c = 0; c < 3; ++c) { out[2 + c] = k0 * d0[c] + k1 * d1[c]; out[5 + c] = k3 * d0[c] + k4 * d1[c]; }
}
KB_HD void camera_project_jac(const CameraConst& cam, V3 X, double& y0, double& y1, Mr<2>& J) {
  if (cam.model == 0) {
    const M3 Km = load_m3(cam.K);
    const V3 pr = Km * X;
    const double iz = 1.0 / pr.z;
    y0 = pr.x * iz; y1 = pr.y * iz;
    J.a[0] = iz * (Km.a[0] - y0 * Km.a[6]); J.a[1] = iz * (Km.a[1] - y0 * Km.a[7]); J.a[2] = iz * (Km.a[2] - y0 * Km.a[8]);
    J.a[3] = iz * (Km.a[3] - y1 * Km.a[6]); J.a[4] = iz * (Km.a[4] - y1 * Km.a[7]); J.a[5] = iz * (Km.a[5] - y1 * Km.a[8]);
    return;
  }
  double o[8];
  camera_project_jac_atan(cam.K[0], cam.K[1], cam.K[2], cam.K[3], cam.K[4], cam.K[5], cam.wc[0], cam.wc[1], cam.gamma, X.x, X.y, X.z, o);
  y0 = o[0]; y1 = o[1];
#pragma unroll
  for (int c = 0; c < 6; ++c) J.a[c] = o[2 + c];
}

// Reference side, ONCE PER LANDMARK REFERENCE (hoisted: the reference re-evaluates it for every observation).
// All observations of a landmark share the reference observation (landmark.h:19-54), hence
//   X = R_r R_ct^T (K^-1 [u v 1] - rho p_ct) + rho p_r            (static_rscamera_measurement.h:43-46)
// and its derivatives with respect to the 4 reference-window knots and to rho.
// record (kRefStride doubles): X(3) | dX/drho(3) | rho | i0_ref (as double) | dX/dknots [4][3][7]
constexpr int kRefStride = 92;
constexpr int kRefDOff = 8;
KB_HD void landmark_ref_se3(const CameraConst& cam, const double* k0, const double* p1, const double* p2, const double* p3, const Basis& bs,
                            const double* ref_uv, double rho, int i0, double* rec) {
  Pose P;
  pose_forward(k0, p1, p2, p3, bs, P);
#ifdef KTK_RCT_INLINE
  const M3 Rct = quat_to_rot(cam.q_ct[0], cam.q_ct[1], cam.q_ct[2], cam.q_ct[3]);
#else
  const M3 Rct = load_m3(cam.Rct);
#endif
  const V3 pct = v3(cam.p_ct[0], cam.p_ct[1], cam.p_ct[2]);
  const V3 yh = camera_unproject(cam, ref_uv[0], ref_uv[1]);
  const V3 Xref = mul_t(Rct, yh - rho * pct);
  const V3 X = P.R * Xref + rho * P.p;
  const V3 dXr = P.p - P.R * mul_t(Rct, pct);
  rec[0] = X.x; rec[1] = X.y; rec[2] = X.z; rec[3] = dXr.x; rec[4] = dXr.y; rec[5] = dXr.z; rec[6] = rho; rec[7] = (double)i0;
  // X = R_r Xref + rho p_r:  dX/dp_r = rho I,  dX/dtheta_r = -R_r hat(Xref)
  pose_backward<3>(k0, p1, p2, p3, bs, rho * m3_identity(), rho * P.R, (-1.0) * mul_hat(P.R, Xref), 1.0, rec + kRefDOff);
}

// Observation side, per measurement.  `ref` is the landmark record above; its dX/dknots part may alias J (the row
// buffer) at J + kRefInRow: the reference-window blocks are produced front to back, each read before it is overwritten.
// J: [ref window: 4 knots][2][7] (56) | [obs window: 4 knots][2][7] (56);  Jrho: d r / d rho (2) -- the last two doubles of
// the packed 114-double row, kept separate so that the kernels stage 112 doubles per row (8 warps of rows per SM).
constexpr int kRefInRow = 0;       // the record fills the row buffer: block k is read at 8 + 21 k (into registers) and written at 14 k
// (the observation pose P is evaluated by the caller first: it does not need the landmark record, so the kernels
//  overlap the record gather with it)
// The adjoints of the observation pose that the second half needs (18 doubles carried across the first scatter).
struct ObsAdjoint { Mr<2> Gp, GpR, Gth; };
// Projection part of the first half: projection, residual (Huber-corrected), d r / d rho, the adjoints of the observation pose and
// GX = d r / d X.  `hdr` = the first 8 doubles of the landmark record (X, dX/drho, rho, i0_ref).
KB_HD void static_rs_project(const CameraConst& cam, const Pose& P, const double* hdr, const double* obs_uv, double weight, double huber_c,
                             double* r, double* Jrho, int* i0_ref, ObsAdjoint& adj, Mr<2>& GX) {
  *i0_ref = (int)hdr[7];
  const V3 X = v3(hdr[0], hdr[1], hdr[2]), dXr = v3(hdr[3], hdr[4], hdr[5]);
  const double rho = hdr[6];
#ifdef KTK_RCT_INLINE
  const M3 Rct = quat_to_rot(cam.q_ct[0], cam.q_ct[1], cam.q_ct[2], cam.q_ct[3]);
#else
  const M3 Rct = load_m3(cam.Rct);
#endif
  const V3 pct = v3(cam.p_ct[0], cam.p_ct[1], cam.p_ct[2]);
  const V3 Xobs = mul_t(P.R, X - rho * P.p);                 // static_rscamera_measurement.h:49
  const V3 Xc = Rct * Xobs + rho * pct;                      // :52
  double y0, y1;
  Mr<2> Jp0;                                                 // d y / d Xc
  camera_project_jac(cam, Xc, y0, y1, Jp0);                  // pinhole_camera.h:47-51 / atan_camera.h:54-75
  double r0 = weight * (obs_uv[0] - y0), r1 = weight * (obs_uv[1] - y1);
  // ceres::HuberLoss + Corrector folded into the row scale: J <- sqrt(rho') (J - alpha/|r|^2 r r^T J), r <- r * scaling
  // (2x2 matrix C applied to the two rows; identity when the loss is off or in its quadratic region)
  double c00 = 1.0, c01 = 0.0, c10 = 0.0, c11 = 1.0, rs = 1.0;
  if (huber_c > 0.0) {
    const HuberScale h = huber_scale(huber_c, r0 * r0 + r1 * r1);
    c00 = h.sqrt_rho1 * (1.0 - h.alpha_sq_norm * r0 * r0); c01 = -h.sqrt_rho1 * h.alpha_sq_norm * r0 * r1;
    c10 = c01; c11 = h.sqrt_rho1 * (1.0 - h.alpha_sq_norm * r1 * r1);
    rs = h.residual_scaling;
  }
  r[0] = rs * r0; r[1] = rs * r1;
  // d r / d Xc = -weight * C * d y / d Xc   (2 x 3)
  Mr<2> Jp;
#pragma unroll
  for (int c = 0; c < 3; ++c) { Jp.a[c] = -weight * (c00 * Jp0.a[c] + c01 * Jp0.a[3 + c]); Jp.a[3 + c] = -weight * (c10 * Jp0.a[c] + c11 * Jp0.a[3 + c]); }
  const Mr<2> Go = rmul(Jp, Rct);               // d r / d Xobs
  GX = rmul_nt(Go, P.R);                        // d r / d X
  // inverse depth: dXc/drho = R_ct R_o^T (dX/drho - p_o) + p_ct
  const V3 dXc = Rct * mul_t(P.R, dXr - P.p) + pct;
  Jrho[0] = Jp.a[0] * dXc.x + Jp.a[1] * dXc.y + Jp.a[2] * dXc.z; Jrho[1] = Jp.a[3] * dXc.x + Jp.a[4] * dXc.y + Jp.a[5] * dXc.z;
  // observation pose: Xobs = R_o^T (X - rho p_o):  d/dp_o = -rho GX,  d/dtheta_o = Go hat(Xobs)
  adj.Gp = rscale(-rho, GX); adj.GpR = rscale(-rho, Go); adj.Gth = rmul_hat(Go, Xobs);
}
// One reference-window block: GX (2x3) * dX/dknot_k (3x7) -> out (2x7).  `blk` is in registers, so `out` may alias the record.
KB_HD void static_rs_ref_block(const Mr<2>& GX, const double* blk, double* out) {
#pragma unroll
  for (int rr = 0; rr < 2; ++rr)
#pragma unroll
    for (int c = 0; c < 7; ++c) out[7 * rr + c] = GX.a[3 * rr] * blk[c] + GX.a[3 * rr + 1] * blk[7 + c] + GX.a[3 * rr + 2] * blk[14 + c];
}
// First half: projection, residual, reference-window blocks Jref[56] (may alias `ref`, see kRefInRow) and d r / d rho.
KB_HD void static_rs_obs_ref_half(const CameraConst& cam, const Pose& P, const double* ref, const double* obs_uv, double weight, double huber_c,
                                  double* r, double* Jref, double* Jrho, int* i0_ref, ObsAdjoint& adj) {
  Mr<2> GX;
  static_rs_project(cam, P, ref, obs_uv, weight, huber_c, r, Jrho, i0_ref, adj, GX);
  // reference-window blocks: GX (2x3) * dX/dknot_k (3x7), in place (see kRefInRow)
  const double* dXk = ref + kRefDOff;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double blk[21];
#pragma unroll
    for (int i = 0; i < 21; ++i) blk[i] = dXk[21 * k + i];
    static_rs_ref_block(GX, blk, Jref + 14 * k);
  }
}

// ---- local (tangent) coordinates -----------------------------------------------------------------------------------------
// Ceres applies the knots' LocalParameterization to the ambient blocks after Evaluate; KTK_EVAL_LOCAL does it in the kernel.
// SE3 knot (uniform_se3_spline_trajectory.h:17-49): Plus(T, [upsilon; omega]) = T exp(.), so with q = (v, w):
//   J_local[:, 0:3] = J[:, 4:7] R(q),   J_local[:, 3:6] = J[:, 0:4] * 1/2 [[w I + hat(v)], [-v^T]]
// Repacks `nblk` consecutive (N x 7) blocks of knots i0.. into (N x 6) blocks in place, front to back.
template <int N>
KB_HD void localize_se3_blocks(double* blocks, int nblk, const double* knot0) {
  for (int k = 0; k < nblk; ++k) {
    const double* q = knot0 + (size_t)k * kKnotStride;
    const M3 R = quat_to_rot(q[0], q[1], q[2], q[3]);
    double in[N * 7];
#pragma unroll
    for (int i = 0; i < N * 7; ++i) in[i] = blocks[k * N * 7 + i];
#pragma unroll
    for (int r = 0; r < N; ++r) {
      const double* j = in + 7 * r;
      double* o = blocks + k * N * 6 + 6 * r;
      o[0] = j[4] * R.a[0] + j[5] * R.a[3] + j[6] * R.a[6];
      o[1] = j[4] * R.a[1] + j[5] * R.a[4] + j[6] * R.a[7];
      o[2] = j[4] * R.a[2] + j[5] * R.a[5] + j[6] * R.a[8];
      o[3] = 0.5 * (j[0] * q[3] + j[1] * q[2] - j[2] * q[1] - j[3] * q[0]);
      o[4] = 0.5 * (-j[0] * q[2] + j[1] * q[3] + j[2] * q[0] - j[3] * q[1]);
      o[5] = 0.5 * (j[0] * q[1] - j[1] * q[0] + j[2] * q[3] - j[3] * q[2]);
    }
  }
}
// SO3 knot (ceres::EigenQuaternionParameterization, uniform_so3_spline_trajectory.h:21): Plus(q, d) = q_d q, d the half-angle
// vector:  J_local = J [[w I - hat(v)], [-v^T]].   (N x 4) -> (N x 3) in place.
template <int N>
KB_HD void localize_so3_blocks(double* blocks, int nblk, const double* quat0) {
  for (int k = 0; k < nblk; ++k) {
    const double* q = quat0 + (size_t)k * 4;
    double in[N * 4];
#pragma unroll
    for (int i = 0; i < N * 4; ++i) in[i] = blocks[k * N * 4 + i];
#pragma unroll
    for (int r = 0; r < N; ++r) {
      const double* j = in + 4 * r;
      double* o = blocks + k * N * 3 + 3 * r;
      o[0] = j[0] * q[3] - j[1] * q[2] + j[2] * q[1] - j[3] * q[0];
      o[1] = j[0] * q[2] + j[1] * q[3] - j[2] * q[0] - j[3] * q[1];
      o[2] = -j[0] * q[1] + j[1] * q[0] + j[2] * q[3] - j[3] * q[2];
    }
  }
}

// =================================================================================================================
// Row drivers: everything one measurement does, from its record to its residual / packed Jacobian row.
// `knots` / `pairs` are indexable by GLOBAL knot index.
// Return 0, or a negative ktk status where the reference would have thrown (trajectory_estimator.h:97-122,
// spline_base.h:196-201, uniform_se3_spline_trajectory.h:121-127).
// =================================================================================================================
struct SplineConst { double t0, dt; int n_knots; int compat_zero_dB; };
constexpr int kStatusRange = -1;      // std::range_error in the reference

KB_HD double spline_max_time(const SplineConst& sp) { return add_rn(sp.t0, mul_rn((double)(sp.n_knots - 3), sp.dt)); }

struct ImuConst { double time_offset, max_time_offset; int time_offset_locked; double bias[3]; };   // bias: ConstantBiasImu (constant_bias_imu.h:52-61), 0 for BasicImu

// PositionMeasurement on SE3 (measurements/position_measurement.h:24-31): r = weight (p_meas - position(t)); J: [4 knots][3][7]
KB_HD void position_se3(const double* knot0, const double* p1, const double* p2, const double* p3, const Basis& bs, double weight,
                        const double* y, double* r, double* J) {
  Pose P;
  pose_forward(knot0, p1, p2, p3, bs, P);
  r[0] = weight * (y[0] - P.p.x); r[1] = weight * (y[1] - P.p.y); r[2] = weight * (y[2] - P.p.z);
  pose_backward<3>(knot0, p1, p2, p3, bs, m3_identity(), P.R, m3_zero(), -weight, J);
}

// OrientationMeasurement (measurements/orientation_measurement.h:27-31): ONE residual, Eigen 3.3's q.angularDistance(q_hat) =
// 2 atan2(|vec(d)|, |d.w|), d = q conj(q_hat): the rotation angle theta in [0, pi] of E = R(q) R_hat^T, evaluated as
// atan2(|s|, c) with s = vee(E - E^T)/2 = sin(theta) n, c = (tr E - 1)/2.  A body-frame perturbation R_hat <- R_hat Exp(d) turns E into
// E Exp(-R_hat d), and d theta = n . eps for E Exp(eps) (n is an eigenvector of Jr^-1 with eigenvalue 1), so G_theta = -(R_hat^T n)^T.
// theta = 0 or pi has no derivative (the reference's Jets give 0/0 there as well).  Returns theta, fills G_theta (1 x 3).
KB_HD double orientation_angle(const double* q_meas /* x y z w */, const M3& Rhat, Mr<1>& Gth) {
  const double inv = 1.0 / sqrt(q_meas[0] * q_meas[0] + q_meas[1] * q_meas[1] + q_meas[2] * q_meas[2] + q_meas[3] * q_meas[3]);
  const M3 E = mul_nt(quat_to_rot(q_meas[0] * inv, q_meas[1] * inv, q_meas[2] * inv, q_meas[3] * inv), Rhat);
  const V3 sv = v3(0.5 * (E.a[7] - E.a[5]), 0.5 * (E.a[2] - E.a[6]), 0.5 * (E.a[3] - E.a[1]));
  const double sn = sqrt(dot(sv, sv)), c = 0.5 * (trace(E) - 1.0);
  const V3 n = (1.0 / sn) * sv;
  const V3 g = mul_t(Rhat, n);
  Gth.a[0] = -g.x; Gth.a[1] = -g.y; Gth.a[2] = -g.z;
  return atan2(sn, c);
}
KB_HD void orientation_se3(const double* knot0, const double* p1, const double* p2, const double* p3, const Basis& bs, const double* q_meas,
                           double* r, double* J /* [4 knots][1][7] */) {
  Pose P;
  pose_forward(knot0, p1, p2, p3, bs, P);
  Mr<1> Gth, zero;
  zero.a[0] = zero.a[1] = zero.a[2] = 0.0;
  r[0] = orientation_angle(q_meas, P.R, Gth);
  pose_backward<1>(knot0, p1, p2, p3, bs, zero, zero, Gth, 1.0, J);
}

// gyroscope (which = 0) / accelerometer (which = 1) / position (which = 2) / orientation (which = 3: y = q (x,y,z,w), r[1], J [4][1][7]);
// gyroscope_measurement.h:75-105 builds the span, :58-68 evaluates.
template <bool PARK = false>
KB_HD int imu_row(int which, const SplineConst& sp, const ImuConst& imu, const double* knots, const double* pairs,
                  double t, const double* y, double weight, double* r, double* J, int* i0_out, double* scratch = nullptr) {
  double ta = t, tb = t;
  if (!imu.time_offset_locked) { ta = sub_rn(t, imu.max_time_offset); tb = add_rn(t, imu.max_time_offset); }
  if (sp.n_knots < 4 || !(ta >= sp.t0) || !(tb < spline_max_time(sp))) return kStatusRange;
  Segment seg; segments_one_span(ta, tb, sp.t0, sp.dt, seg);
  int i0; double u;
  if (!segment_locate(seg, add_rn(t, imu.time_offset), sp.t0, sp.dt, i0, u)) return kStatusRange;
  Basis bs = cumulative_basis(u, sp.dt);
  const double* k0 = knots + (size_t)i0 * kKnotStride;
  const double* p1 = pairs + (size_t)(i0 + 1) * kPairStride;
  if (which == 0) gyro_se3(k0, p1, p1 + kPairStride, p1 + 2 * kPairStride, bs, weight, y, r, J);
  else if (which == 2) position_se3(k0, p1, p1 + kPairStride, p1 + 2 * kPairStride, bs, weight, y, r, J);
  else if (which == 3) orientation_se3(k0, p1, p1 + kPairStride, p1 + 2 * kPairStride, bs, y, r, J);
  else {
    if (sp.compat_zero_dB) { bs.dB[0] = 0.0; bs.dB[1] = 0.0; bs.dB[2] = 0.0; }
    accel_se3<PARK>(k0, p1, p1 + kPairStride, p1 + 2 * kPairStride, bs, weight, y, r, J, scratch);
  }
  *i0_out = i0;
  return 0;
}

// The two spans of a static-RS residual (static_rscamera_measurement.h:137-166) and their segments
// (spline_base.h:371-403).  Returns the number of segments, 0 where CheckTimeSpans (trajectory_estimator.h:97-122) throws.
KB_HD int static_rs_segments(const SplineConst& sp, const CameraConst& cam, double ref_t0, double obs_t0, Segment& s0, Segment& s1) {
  double t1, t2;
  if (ref_t0 <= obs_t0) { t1 = ref_t0; t2 = obs_t0; } else { t1 = obs_t0; t2 = ref_t0; }
  if (!cam.time_offset_locked) { t1 = sub_rn(t1, cam.max_time_offset); t2 = add_rn(t2, cam.max_time_offset); }
  const double margin = 1e-3;
  const double a1 = sub_rn(t1, margin), b1 = add_rn(add_rn(t1, cam.readout), margin);
  const double a2 = sub_rn(t2, margin), b2 = add_rn(add_rn(t2, cam.readout), margin);
  const double tmax = spline_max_time(sp);
  if (sp.n_knots < 4 || !(a1 >= sp.t0) || !(b1 < tmax) || !(a2 >= sp.t0) || !(b2 < tmax) || a1 > b1 || a2 > b2 || a2 < a1) return 0;
  return segments_two_spans(a1, b1, a2, b2, sp.t0, sp.dt, s0, s1);
}
// Evaluation time of an observation: t0_view + time_offset + v * row_delta (static_rscamera_measurement.h:30-33)
KB_HD double static_rs_time(const CameraConst& cam, double view_t0, double v) { return add_rn(add_rn(view_t0, cam.time_offset), mul_rn(v, cam.row_delta)); }
// SplineView::Evaluate over the (at most two) segments: first segment that holds t (spline_base.h:188-202).
// Returns the index of that segment (0/1) or -1.
KB_HD int locate_in_segments(int nseg, const Segment& s0, const Segment& s1, double t, double t0, double dt, int& i0, double& u) {
  if (segment_locate(s0, t, t0, dt, i0, u)) return 0;
  if (nseg == 2 && segment_locate(s1, t, t0, dt, i0, u)) return 1;
  return -1;
}

// One landmark-reference record: the spline evaluation at the reference observation's time inside the segment whose
// first knot is `seg_start` with `seg_n` knots (found on the host from the residual's two spans; identical for all
// observations of a landmark whose reference view is the earlier one).
KB_HD int landmark_ref_row(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* ref_uv,
                           double ref_t0, int seg_start, int seg_n, double rho, double* rec) {
  Segment s; s.start = seg_start; s.n = seg_n;
  int i0; double u;
  if (!segment_locate(s, static_rs_time(cam, ref_t0, ref_uv[1]), sp.t0, sp.dt, i0, u)) return kStatusRange;
  if (i0 < 0 || i0 + 3 >= sp.n_knots) return kStatusRange;
  const Basis bs = cumulative_basis(u, sp.dt);
  const double* p1 = pairs + (size_t)(i0 + 1) * kPairStride;
  landmark_ref_se3(cam, knots + (size_t)i0 * kKnotStride, p1, p1 + kPairStride, p1 + 2 * kPairStride, bs, ref_uv, rho, i0, rec);
  return 0;
}

// static RS camera row; static_rscamera_measurement.h:130-198 builds the two spans, :112-123 / :21-55 evaluate.
// huber_c > 0 applies ceres::HuberLoss + Corrector to (r, J) as Ceres does after Evaluate.
// Part 1 (no landmark record needed): spans, segment lookup and basis, then the observation pose.  The two halves are
// separate so that a kernel can choose where `pairs` points (global table or a staged window) once i0 is known.
struct ObsForward { int status, io; Basis bo; Pose P; };
KB_HD void static_rs_row_locate(const SplineConst& sp, const CameraConst& cam, const double* obs_uv, double obs_t0, double ref_t0, ObsForward& f) {
  Segment s0, s1;
  const int nseg = static_rs_segments(sp, cam, ref_t0, obs_t0, s0, s1);
  double uo = 0.0;
  f.io = -1;
  if (nseg == 0 || locate_in_segments(nseg, s0, s1, static_rs_time(cam, obs_t0, obs_uv[1]), sp.t0, sp.dt, f.io, uo) < 0) { f.status = kStatusRange; f.io = -1; return; }
  f.status = 0;
  f.bo = cumulative_basis(uo, sp.dt);
}
// The same lookup for the host: first knot and interpolation amount of the observation evaluation.  Neither depends on the
// evaluation point (knots, rho), so upload_group() runs it ONCE per row and the kernels read (io, u) instead of redoing the two
// spans, the segment rule and their six fp64 divisions in every evaluation.  Returns false where the reference throws.
KB_HD bool static_rs_row_locate_u(const SplineConst& sp, const CameraConst& cam, const double* obs_uv, double obs_t0, double ref_t0, int& io, double& uo) {
  Segment s0, s1;
  const int nseg = static_rs_segments(sp, cam, ref_t0, obs_t0, s0, s1);
  io = -1; uo = 0.0;
  if (nseg == 0 || locate_in_segments(nseg, s0, s1, static_rs_time(cam, obs_t0, obs_uv[1]), sp.t0, sp.dt, io, uo) < 0) { io = -1; return false; }
  return true;
}
KB_HD void static_rs_row_pose(const double* knots, const double* pairs, ObsForward& f) {
  if (f.status != 0) return;
  const double* po1 = pairs + (size_t)(f.io + 1) * kPairStride;
  pose_forward(knots + (size_t)f.io * kKnotStride, po1, po1 + kPairStride, po1 + 2 * kPairStride, f.bo, f.P);
}
// Part 2: projection, residual, reference-window half of the row (ref = landmark record, may alias Jref).
KB_HD int static_rs_row_ref_half(const CameraConst& cam, const ObsForward& f, const double* ref, const double* obs_uv, double weight, double huber_c,
                                 double* r, double* Jref, double* Jrho, int* i0_ref_out, int* i0_obs_out, ObsAdjoint& adj) {
  if (f.status != 0) return f.status;
  int ir;
  static_rs_obs_ref_half(cam, f.P, ref, obs_uv, weight, huber_c, r, Jref, Jrho, &ir, adj);
  if (ir < 0) return kStatusRange;                 // the landmark record itself was out of range
  *i0_ref_out = ir; *i0_obs_out = f.io;
  return 0;
}
// Part 3: observation-window blocks Jobs[56] from the adjoints.
KB_HD void static_rs_row_obs_half(const double* knots, const double* pairs, const ObsForward& f, const ObsAdjoint& adj, double* Jobs) {
  const double* po1 = pairs + (size_t)(f.io + 1) * kPairStride;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    Jobs[r * 7 + 4] = adj.Gp.a[3 * r]; Jobs[r * 7 + 5] = adj.Gp.a[3 * r + 1]; Jobs[r * 7 + 6] = adj.Gp.a[3 * r + 2];
    Jobs[r * 7 + 0] = adj.GpR.a[3 * r]; Jobs[r * 7 + 1] = adj.GpR.a[3 * r + 1]; Jobs[r * 7 + 2] = adj.GpR.a[3 * r + 2];
  }
  pose_backward<2, true>(knots + (size_t)f.io * kKnotStride, po1, po1 + kPairStride, po1 + 2 * kPairStride, f.bo, adj.Gp, adj.GpR, adj.Gth, 1.0, Jobs);
}

}  // namespace kb

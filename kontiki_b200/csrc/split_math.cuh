// kontiki_b200 -- per-measurement mathematics for the SPLIT trajectory (UniformR3SplineTrajectory for the position +
// UniformSO3SplineTrajectory for the orientation), same structure as spline_math.cuh: __host__ __device__ inline, run
// by the CUDA kernels on the device and by tests/host_check.cpp on the host (test harness only).
//
// Reference (relative to /root/reference/cpplib/include/kontiki/):
//   trajectories/split_trajectory.h:34-66,117-123      parameter slices [R3 | SO3], flag dispatch, min/max time
//   trajectories/uniform_r3_spline_trajectory.h:34-101 standard cubic B-spline  p, v, a  (basis M, spline_base.h:18-28)
//   trajectories/uniform_so3_spline_trajectory.h:46-125 cumulative quaternion spline + dq, w_world = 2 (dq q^-1).vec
//   math/quaternion_math.h:16-95                       logq (atan2, half-angle vector), expq, angular_velocity
//
// Quaternion-norm ("radial") behaviour of the reference's arithmetic, which fixes the ambient 4-vector Jacobians:
//   * logq(conj(qa) qb) is scale invariant (atan2(|v|, w) and v/|v|), so the three relative rotations -- and with them
//     e_j = expq(B_j logq(..)) -- do not depend on the norm of any knot; their ambient derivatives come out of the same
//     single-direction dual number prepass as for SE3 (so3_pair_log below);
//   * q = q0 e1 e2 e3 and dq = q0 (sum ...) are LINEAR in q0 (Eigen's product does not renormalise,
//     uniform_so3_spline_trajectory.h:96-121), and q*v / R(q) are Eigen's polynomials:
//       gyro(s)  = (1+s)^2 w_w + (1+s)^4 (R^T - I) w_w      =>  d/ds = 4 w_b - 2 R w_b
//       R((1+s)q) = I + (1+s)^2 (R - I)                       =>  d/ds = 2 (R - I)      (accelerometer, camera)
#pragma once
#include "spline_math.cuh"

namespace kb {

// SO3 knot record: 4 doubles [x y z w] (already 32 B); R3 knot record: 4 doubles [x y z pad]
// SO3 pair record: 28 doubles [phi(3) pad(1) Da(3x4) Db(3x4)] (224 B); phi = logq(conj(q_{p-1}) q_p).vec, the HALF-angle vector
constexpr int kQuatStride = 4;
constexpr int kVecStride = 4;
constexpr int kSo3PairStride = 28;
constexpr int kSo3PairDOff = 4;
constexpr int kSo3PairSide = 12;
constexpr double kEpsLogq = 1e-16;       // quaternion_math.h:10 (on the squared norm)
constexpr double kEpsUnit = 1e-5;        // quaternion_math.h:11
constexpr int kStatusRuntime = -2;       // std::runtime_error in the reference (logq on a non-unit quaternion)

KB_HD double t_atan2(double y, double x) { return atan2(y, x); }
KB_HD D1 t_atan2(D1 y, D1 x) { const double inv = 1.0 / (x.a * x.a + y.a * y.a); return D1(atan2(y.a, x.a), inv * (x.a * y.d - y.a * x.d)); }

// logq(conj(qa) * qb).vec  (uniform_so3_spline_trajectory.h:97-98, quaternion_math.h:16-59).  Returns false where the
// reference throws ("logq: Only implemented for unit quaternions").
template <class T> KB_HD bool so3_pair_log(const T* a, const T* b, T* phi) {
  const T cx = -a[0], cy = -a[1], cz = -a[2], cw = a[3];
  const T qx = cw * b[0] + cx * b[3] + cy * b[2] - cz * b[1];
  const T qy = cw * b[1] + cy * b[3] + cz * b[0] - cx * b[2];
  const T qz = cw * b[2] + cz * b[3] + cx * b[1] - cy * b[0];
  const T qw = cw * b[3] - cx * b[0] - cy * b[1] - cz * b[2];
  const T v2 = qx * qx + qy * qy + qz * qz;
  const double qn = sqrt(value(v2) + value(qw) * value(qw));
  if (fabs(qn - 1.0) > kEpsUnit) return false;
  T k(1.0);
  if (value(v2) > kEpsLogq) { const T vn = t_sqrt(v2); k = t_atan2(vn, qw) / vn; }
  phi[0] = qx * k; phi[1] = qy * k; phi[2] = qz * k;
  return true;
}
// One (pair, direction) item of the SO3 prepass: dir in [0,8) = ambient scalar of [q_{p-1} (4), q_p (4)], dir == 8: value.
KB_HD int so3_pair_prepass_item(const double* quats, int p, int dir, double* pairs) {
  const double* qa = quats + (size_t)(p - 1) * kQuatStride;
  const double* qb = quats + (size_t)p * kQuatStride;
  double* rec = pairs + (size_t)p * kSo3PairStride;
  if (dir >= 8) {
    double phi[3];
    if (!so3_pair_log<double>(qa, qb, phi)) { rec[0] = rec[1] = rec[2] = 0.0; rec[3] = 0.0; return kStatusRuntime; }
    rec[0] = phi[0]; rec[1] = phi[1]; rec[2] = phi[2]; rec[3] = 0.0;
    return 0;
  }
  D1 a[4], b[4], phi[3];
  for (int i = 0; i < 4; ++i) { a[i] = D1(qa[i], dir == i ? 1.0 : 0.0); b[i] = D1(qb[i], dir == 4 + i ? 1.0 : 0.0); }
  const bool ok = so3_pair_log<D1>(a, b, phi);
  double* D = rec + kSo3PairDOff + (dir >= 4 ? kSo3PairSide : 0);
  const int c = dir & 3;
  for (int i = 0; i < 3; ++i) D[i * 4 + c] = ok ? phi[i].d : 0.0;
  return 0;
}

// Standard (non-cumulative) cubic B-spline basis of the R3 spline and its time derivatives
// (uniform_r3_spline_trajectory.h:57-78, spline_base.h:18-22 M).
struct BasisR3 { double Bp[4], Bv[4], Ba[4]; };
KB_HD BasisR3 r3_basis(double u, double dt) {
  BasisR3 b;
  const double u2 = u * u, u3 = u2 * u, di = 1.0 / dt, di2 = di * di, s = 1.0 / 6.0;
  b.Bp[0] = (1.0 - 3.0 * u + 3.0 * u2 - u3) * s; b.Bp[1] = (4.0 - 6.0 * u2 + 3.0 * u3) * s;
  b.Bp[2] = (1.0 + 3.0 * u + 3.0 * u2 - 3.0 * u3) * s; b.Bp[3] = u3 * s;
  b.Bv[0] = di * (-3.0 + 6.0 * u - 3.0 * u2) * s; b.Bv[1] = di * (-12.0 * u + 9.0 * u2) * s;
  b.Bv[2] = di * (3.0 + 6.0 * u - 9.0 * u2) * s; b.Bv[3] = di * (3.0 * u2) * s;
  b.Ba[0] = di2 * (1.0 - u); b.Ba[1] = di2 * (3.0 * u - 2.0); b.Ba[2] = di2 * (1.0 - 3.0 * u); b.Ba[3] = di2 * u;
  return b;
}
KB_HD V3 r3_combine(const double* c0, const double* w) {
  return v3(w[0] * c0[0] + w[1] * c0[kVecStride] + w[2] * c0[2 * kVecStride] + w[3] * c0[3 * kVecStride],
            w[0] * c0[1] + w[1] * c0[kVecStride + 1] + w[2] * c0[2 * kVecStride + 1] + w[3] * c0[3 * kVecStride + 1],
            w[0] * c0[2] + w[1] * c0[kVecStride + 2] + w[2] * c0[2 * kVecStride + 2] + w[3] * c0[3 * kVecStride + 2]);
}

// Rotation part of exp(B * 2 phi): the SO3 spline stores half-angle vectors, the rotation vector is 2 phi.
KB_HD void so3_exp_part(const double* pr, double B, ExpPart& e) {
  const double om[6] = {0.0, 0.0, 0.0, 2.0 * pr[0], 2.0 * pr[1], 2.0 * pr[2]};
  exp_part(om, B, false, false, e);
}
// R = R(q0) E1 E2 E3
KB_HD M3 so3_forward(const double* q0, const double* p1, const Basis& bs) {
  ExpPart e;
  so3_exp_part(p1 + 2 * kSo3PairStride, bs.B[2], e);
  M3 T = e.E;
  so3_exp_part(p1 + kSo3PairStride, bs.B[1], e);
  T = e.E * T;
  so3_exp_part(p1, bs.B[0], e);
  T = e.E * T;
  return quat_to_rot(q0[0], q0[1], q0[2], q0[3]) * T;
}

// J_block(N x 4) (+)= scale * G (N x 3) * D (3 x 4 side of an SO3 pair record)
template <int N, bool ACC>
KB_HD void contract_so3(double* J, const Mr<N>& G, const double* D, double scale) {
#pragma unroll
  for (int r = 0; r < N; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double s = G.a[3 * r] * D[c] + G.a[3 * r + 1] * D[4 + c] + G.a[3 * r + 2] * D[8 + c];
      if (ACC) J[r * 4 + c] += scale * s; else J[r * 4 + c] = scale * s;
    }
}
// quaternion block of knot i0 (N x 4): J += scale * (Gth0 * dtheta/dq + grad * q^T)
template <int N>
KB_HD void add_q0_block4(double* J, const Mr<N>& Gth, const double* grad, const double* q, double scale) {
  const V3 v = v3(q[0], q[1], q[2]); const double w = q[3];
#pragma unroll
  for (int r = 0; r < N; ++r) {
    const V3 g = rrow(Gth, r);
    const V3 t = 2.0 * (w * g - cross(g, v));
    J[r * 4 + 0] += scale * (t.x + grad[r] * q[0]);
    J[r * 4 + 1] += scale * (t.y + grad[r] * q[1]);
    J[r * 4 + 2] += scale * (t.z + grad[r] * q[2]);
    J[r * 4 + 3] += scale * (-2.0 * dot(g, v) + grad[r] * q[3]);
  }
}
// Reverse sweep of the orientation: given N row-adjoints Gth w.r.t. a body-frame perturbation R <- R Exp(d) and the radial
// adjoint grad (d row / ds for q0 <- (1+s) q0), writes scale * d(row)/d(4 quaternion knots) into J ([4][N][4]).
// The factor 2 converts d/d(rotation vector) into d/d(half-angle vector phi).
template <int N>
KB_HD void so3_backward(const double* q0, const double* p1, const Basis& bs, const Mr<N>& Gth, const double* grad, double scale, double* J) {
  ExpPart e;
  const double* p2 = p1 + kSo3PairStride; const double* p3 = p2 + kSo3PairStride;
  so3_exp_part(p3, bs.B[2], e);
  Mr<N> G = rscale(2.0 * bs.B[2], rmul_nt(Gth, e.V));
  contract_so3<N, false>(J + 3 * N * 4, G, p3 + kSo3PairDOff + kSo3PairSide, scale);
  contract_so3<N, false>(J + 2 * N * 4, G, p3 + kSo3PairDOff, scale);
  M3 T = e.E;
  so3_exp_part(p2, bs.B[1], e);
  G = rscale(2.0 * bs.B[1], rmul_nt(rmul_nt(Gth, T), e.V));
  contract_so3<N, true>(J + 2 * N * 4, G, p2 + kSo3PairDOff + kSo3PairSide, scale);
  contract_so3<N, false>(J + 1 * N * 4, G, p2 + kSo3PairDOff, scale);
  T = e.E * T;
  so3_exp_part(p1, bs.B[0], e);
  G = rscale(2.0 * bs.B[0], rmul_nt(rmul_nt(Gth, T), e.V));
  contract_so3<N, true>(J + 1 * N * 4, G, p1 + kSo3PairDOff + kSo3PairSide, scale);
  contract_so3<N, false>(J + 0 * N * 4, G, p1 + kSo3PairDOff, scale);
  T = e.E * T;
  add_q0_block4<N>(J, rmul_nt(Gth, T), grad, q0, scale);
}

// ---- gyroscope on Split (imu.h:47-52; only the SO3 spline is evaluated, split_trajectory.h:48-57) -------------------
// J: [4 SO3 knots][3][4]
KB_HD void gyro_split(const double* q0, const double* p1, const Basis& bs, double weight, const double* y, double* r, double* J) {
  const double* p2 = p1 + kSo3PairStride; const double* p3 = p2 + kSo3PairStride;
  ExpPart e1, e2, e3;
  so3_exp_part(p1, bs.B[0], e1); so3_exp_part(p2, bs.B[1], e2); so3_exp_part(p3, bs.B[2], e3);
  const V3 f1 = v3(2.0 * p1[0], 2.0 * p1[1], 2.0 * p1[2]), f2 = v3(2.0 * p2[0], 2.0 * p2[1], 2.0 * p2[2]), f3 = v3(2.0 * p3[0], 2.0 * p3[1], 2.0 * p3[2]);
  const V3 y2 = mul_t(e2.E, bs.dB[0] * f1);
  const V3 s2 = y2 + bs.dB[1] * f2;
  const V3 y3 = mul_t(e3.E, s2);
  const V3 wb = y3 + bs.dB[2] * f3;
  r[0] = weight * (y[0] - wb.x); r[1] = weight * (y[1] - wb.y); r[2] = weight * (y[2] - wb.z);
  M3 G3 = bs.B[2] * hat_mul(y3, transpose(e3.V)); G3.a[0] += bs.dB[2]; G3.a[4] += bs.dB[2]; G3.a[8] += bs.dB[2];
  M3 T2 = bs.B[1] * hat_mul(y2, transpose(e2.V)); T2.a[0] += bs.dB[1]; T2.a[4] += bs.dB[1]; T2.a[8] += bs.dB[1];
  const M3 G2 = mul_tn(e3.E, T2);
  const M3 G1 = bs.dB[0] * mul_tn(e3.E, transpose(e2.E));
  const double sc = -2.0 * weight;           // d/d(phi) = 2 d/d(rotation vector)
  contract_so3<3, false>(J + 0, G1, p1 + kSo3PairDOff, sc);
  contract_so3<3, false>(J + 12, G1, p1 + kSo3PairDOff + kSo3PairSide, sc);
  contract_so3<3, true>(J + 12, G2, p2 + kSo3PairDOff, sc);
  contract_so3<3, false>(J + 24, G2, p2 + kSo3PairDOff + kSo3PairSide, sc);
  contract_so3<3, true>(J + 24, G3, p3 + kSo3PairDOff, sc);
  contract_so3<3, false>(J + 36, G3, p3 + kSo3PairDOff + kSo3PairSide, sc);
  // radial (see header): 4 w_b - 2 R w_b
  const M3 R = quat_to_rot(q0[0], q0[1], q0[2], q0[3]) * (e1.E * (e2.E * e3.E));
  const V3 rad = 4.0 * wb - 2.0 * (R * wb);
  const double radv[3] = {rad.x, rad.y, rad.z};
  add_q0_block4<3>(J, m3_zero(), radv, q0, -weight);
}

// ---- accelerometer on Split (imu.h:55-59): accel = q^-1 (a_world + g), a_world from the R3 spline -------------------
// J: [4 R3 knots][3][3] (36) | [4 SO3 knots][3][4] (48)
KB_HD void accel_split(const double* c0, const BasisR3& br, const double* q0, const double* p1, const Basis& bs, double weight,
                       const double* y, double* r, double* J) {
  const M3 R = so3_forward(q0, p1, bs);
  const V3 aw = r3_combine(c0, br.Ba) + v3(0.0, 0.0, -kGravity);
  const V3 ab = mul_t(R, aw);
  r[0] = weight * (y[0] - ab.x); r[1] = weight * (y[1] - ab.y); r[2] = weight * (y[2] - ab.z);
  const double sc = -weight;
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) J[9 * k + 3 * i + j] = sc * br.Ba[k] * R.a[3 * j + i];     // Ba_k R^T
  const V3 rad = 2.0 * (ab - aw);
  const double radv[3] = {rad.x, rad.y, rad.z};
  so3_backward<3>(q0, p1, bs, hat(ab), radv, sc, J + 36);
}

// ---- static RS camera on Split -----------------------------------------------------------------------------------------
// Landmark-reference record (kRefSplitStride doubles):
//   X(3) | dX/drho(3) | rho | i0_ref_r3 | i0_ref_so3 | pad | rho*Bp_ref[4] | pad(2) | dX/dq [4 SO3 knots][3][4]
constexpr int kRefSplitStride = 64;
constexpr int kRefSplitBp = 10;
constexpr int kRefSplitDq = 16;
KB_HD void landmark_ref_split(const CameraConst& cam, const double* c0, const BasisR3& br, int i0_r3, const double* q0, const double* p1,
                              const Basis& bs, int i0_so3, const double* ref_uv, double rho, double* rec) {
  const M3 R = so3_forward(q0, p1, bs);
  const V3 p = r3_combine(c0, br.Bp);
  const M3 Rct = load_m3(cam.Rct);
  const V3 pct = v3(cam.p_ct[0], cam.p_ct[1], cam.p_ct[2]);
  const V3 yh = camera_unproject(cam, ref_uv[0], ref_uv[1]);
  const V3 Xref = mul_t(Rct, yh - rho * pct);
  const V3 RX = R * Xref;
  const V3 X = RX + rho * p;
  const V3 dXr = p - R * mul_t(Rct, pct);
  rec[0] = X.x; rec[1] = X.y; rec[2] = X.z; rec[3] = dXr.x; rec[4] = dXr.y; rec[5] = dXr.z; rec[6] = rho;
  rec[7] = (double)i0_r3; rec[8] = (double)i0_so3; rec[9] = 0.0;
  for (int k = 0; k < 4; ++k) rec[kRefSplitBp + k] = rho * br.Bp[k];
  rec[14] = 0.0; rec[15] = 0.0;
  // X = R Xref + rho p:  dX/dtheta_body = -R hat(Xref);  radial: 2 (R - I) Xref
  const V3 rad = 2.0 * (RX - Xref);
  const double radv[3] = {rad.x, rad.y, rad.z};
  so3_backward<3>(q0, p1, bs, (-1.0) * mul_hat(R, Xref), radv, 1.0, rec + kRefSplitDq);
}
// Observation side.  J: [ref R3 4x(2x3)] (24) | [ref SO3 4x(2x4)] (32) | [obs R3] (24) | [obs SO3] (32);  Jrho: d r/d rho (2)
// `ref` may alias J + kRefSplitInRow (the kernels gather the record into the row buffer): its fields are consumed
// front to back before the positions they occupy are written.
constexpr int kRefSplitInRow = 48;      // 48 + 64 = 112 = staged row length
KB_HD void static_rs_obs_split(const CameraConst& cam, const double* c0, const BasisR3& br, const double* q0, const double* p1, const Basis& bs,
                               const M3& R, const double* ref, const double* obs_uv, double weight, double huber_c, double* r, double* J,
                               double* Jrho, int* i0_ref_r3, int* i0_ref_so3) {
  *i0_ref_r3 = (int)ref[7]; *i0_ref_so3 = (int)ref[8];
  const V3 X = v3(ref[0], ref[1], ref[2]), dXr = v3(ref[3], ref[4], ref[5]);
  const double rho = ref[6];
  const double rbp[4] = {ref[kRefSplitBp], ref[kRefSplitBp + 1], ref[kRefSplitBp + 2], ref[kRefSplitBp + 3]};
  const V3 p = r3_combine(c0, br.Bp);
  const M3 Rct = load_m3(cam.Rct);
  const V3 pct = v3(cam.p_ct[0], cam.p_ct[1], cam.p_ct[2]);
  const V3 Xw = X - rho * p;
  const V3 Xobs = mul_t(R, Xw);
  const V3 Xc = Rct * Xobs + rho * pct;
  double y0, y1;
  Mr<2> Jp0;
  camera_project_jac(cam, Xc, y0, y1, Jp0);
  double r0 = weight * (obs_uv[0] - y0), r1 = weight * (obs_uv[1] - y1);
  double c00 = 1.0, c01 = 0.0, c10 = 0.0, c11 = 1.0, rs = 1.0;
  if (huber_c > 0.0) {
    const HuberScale h = huber_scale(huber_c, r0 * r0 + r1 * r1);
    c00 = h.sqrt_rho1 * (1.0 - h.alpha_sq_norm * r0 * r0); c01 = -h.sqrt_rho1 * h.alpha_sq_norm * r0 * r1;
    c10 = c01; c11 = h.sqrt_rho1 * (1.0 - h.alpha_sq_norm * r1 * r1);
    rs = h.residual_scaling;
  }
  r[0] = rs * r0; r[1] = rs * r1;
  Mr<2> Jp;
#pragma unroll
  for (int c = 0; c < 3; ++c) { Jp.a[c] = -weight * (c00 * Jp0.a[c] + c01 * Jp0.a[3 + c]); Jp.a[3 + c] = -weight * (c10 * Jp0.a[c] + c11 * Jp0.a[3 + c]); }
  const Mr<2> Go = rmul(Jp, Rct);               // d r / d Xobs
  const Mr<2> GX = rmul_nt(Go, R);              // d r / d X
  // reference SO3 window: GX (2x3) * dX/dq_k (3x4), in place: block k read at 48+16+12k, written at 24+8k
  const double* dq = ref + kRefSplitDq;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double blk[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) blk[i] = dq[12 * k + i];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
      for (int c = 0; c < 4; ++c) J[24 + 8 * k + 4 * rr + c] = GX.a[3 * rr] * blk[c] + GX.a[3 * rr + 1] * blk[4 + c] + GX.a[3 * rr + 2] * blk[8 + c];
  }
  // reference R3 window: GX * rho Bp_k ; observation R3 window: -rho Bp_k(obs) GX
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < 6; ++i) { J[6 * k + i] = rbp[k] * GX.a[i]; J[56 + 6 * k + i] = -rho * br.Bp[k] * GX.a[i]; }
  // inverse depth
  const V3 dXc = Rct * mul_t(R, dXr - p) + pct;
  const double jr0 = Jp.a[0] * dXc.x + Jp.a[1] * dXc.y + Jp.a[2] * dXc.z, jr1 = Jp.a[3] * dXc.x + Jp.a[4] * dXc.y + Jp.a[5] * dXc.z;
  // observation orientation: Xobs = R^T Xw: d/dtheta_body = Go hat(Xobs); radial: 2 (R^T - I) Xw = 2 (Xobs - Xw)
  const V3 radx = 2.0 * (Xobs - Xw);
  const double radv[2] = {Go.a[0] * radx.x + Go.a[1] * radx.y + Go.a[2] * radx.z, Go.a[3] * radx.x + Go.a[4] * radx.y + Go.a[5] * radx.z};
  so3_backward<2>(q0, p1, bs, rmul_hat(Go, Xobs), radv, 1.0, J + 80);
  Jrho[0] = jr0; Jrho[1] = jr1;
}

// =================================================================================================================
// Row drivers
// =================================================================================================================
struct SplitConst { double t0_r3, dt_r3; int n_r3; double t0_so3, dt_so3; int n_so3; };
KB_HD double split_min_time(const SplitConst& sp) { return sp.t0_r3 > sp.t0_so3 ? sp.t0_r3 : sp.t0_so3; }     // split_trajectory.h:60-62
KB_HD double split_max_time(const SplitConst& sp) {                                                                 // :64-66
  const double a = add_rn(sp.t0_r3, mul_rn((double)(sp.n_r3 - 3), sp.dt_r3)), b = add_rn(sp.t0_so3, mul_rn((double)(sp.n_so3 - 3), sp.dt_so3));
  return a < b ? a : b;
}

// which: 0 gyroscope (J 48 doubles, i0 = SO3), 1 accelerometer (J 84 doubles; i0_r3, i0_so3), 2 position (J [4 R3 knots][3][3] = 36),
// 3 orientation (y = q (x,y,z,w), ONE residual, J [4 SO3 knots][1][4] = 16)
KB_HD int imu_row_split(int which, const SplitConst& sp, const ImuConst& imu, const double* vecs, const double* quats, const double* pairs,
                        double t, const double* y, double weight, double* r, double* J, int* i0_r3, int* i0_so3) {
  double ta = t, tb = t;
  if (!imu.time_offset_locked) { ta = sub_rn(t, imu.max_time_offset); tb = add_rn(t, imu.max_time_offset); }
  if (sp.n_r3 < 4 || sp.n_so3 < 4 || !(ta >= split_min_time(sp)) || !(tb < split_max_time(sp))) return kStatusRange;
  const double te = add_rn(t, imu.time_offset);
  Segment seg; int ib; double ub;
  if (which == 2) {      // position_measurement.h: only the R3 spline is evaluated; the SO3 segment is structure only
    int ia; double ua;
    segments_one_span(ta, tb, sp.t0_so3, sp.dt_so3, seg);
    *i0_so3 = seg.start;
    segments_one_span(ta, tb, sp.t0_r3, sp.dt_r3, seg);
    if (!segment_locate(seg, te, sp.t0_r3, sp.dt_r3, ia, ua)) return kStatusRange;
    *i0_r3 = ia;
    const BasisR3 br = r3_basis(ua, sp.dt_r3);
    const V3 pos = r3_combine(vecs + (size_t)ia * kVecStride, br.Bp);
    r[0] = weight * (y[0] - pos.x); r[1] = weight * (y[1] - pos.y); r[2] = weight * (y[2] - pos.z);
    for (int k = 0; k < 4; ++k)
      for (int a = 0; a < 3; ++a)
        for (int c = 0; c < 3; ++c) J[9 * k + 3 * a + c] = a == c ? -weight * br.Bp[k] : 0.0;
    return 0;
  }
  segments_one_span(ta, tb, sp.t0_so3, sp.dt_so3, seg);
  if (!segment_locate(seg, te, sp.t0_so3, sp.dt_so3, ib, ub)) return kStatusRange;
  const Basis bs = cumulative_basis(ub, sp.dt_so3);
  const double* q0 = quats + (size_t)ib * kQuatStride;
  const double* p1 = pairs + (size_t)(ib + 1) * kSo3PairStride;
  *i0_so3 = ib;
  if (which == 3) {      // orientation_measurement.h: only the SO3 spline is evaluated (J [4 SO3 knots][1][4]); angularDistance is scale invariant
    segments_one_span(ta, tb, sp.t0_r3, sp.dt_r3, seg);
    *i0_r3 = seg.start;
    const M3 R = so3_forward(q0, p1, bs);
    Mr<1> Gth;
    r[0] = orientation_angle(y, R, Gth);
    const double grad[1] = {0.0};
    so3_backward<1>(q0, p1, bs, Gth, grad, 1.0, J);
    return 0;
  }
  if (which == 0) {
    // the R3 segment is part of the residual's structure (split_trajectory.h:117-123) and must exist, but is not evaluated
    segments_one_span(ta, tb, sp.t0_r3, sp.dt_r3, seg);
    *i0_r3 = seg.start;
    gyro_split(q0, p1, bs, weight, y, r, J);
    return 0;
  }
  int ia; double ua;
  segments_one_span(ta, tb, sp.t0_r3, sp.dt_r3, seg);
  if (!segment_locate(seg, te, sp.t0_r3, sp.dt_r3, ia, ua)) return kStatusRange;
  *i0_r3 = ia;
  accel_split(vecs + (size_t)ia * kVecStride, r3_basis(ua, sp.dt_r3), q0, p1, bs, weight, y, r, J);
  return 0;
}

// spans of a static-RS residual against the split trajectory's valid time, then the segments of ONE of its two splines
KB_HD int static_rs_segments_split(const SplitConst& sp, const CameraConst& cam, double ref_t0, double obs_t0, double t0, double dt, Segment& s0, Segment& s1) {
  double t1, t2;
  if (ref_t0 <= obs_t0) { t1 = ref_t0; t2 = obs_t0; } else { t1 = obs_t0; t2 = ref_t0; }
  if (!cam.time_offset_locked) { t1 = sub_rn(t1, cam.max_time_offset); t2 = add_rn(t2, cam.max_time_offset); }
  const double margin = 1e-3;
  const double a1 = sub_rn(t1, margin), b1 = add_rn(add_rn(t1, cam.readout), margin);
  const double a2 = sub_rn(t2, margin), b2 = add_rn(add_rn(t2, cam.readout), margin);
  const double tmin = split_min_time(sp), tmax = split_max_time(sp);
  if (sp.n_r3 < 4 || sp.n_so3 < 4 || !(a1 >= tmin) || !(b1 < tmax) || !(a2 >= tmin) || !(b2 < tmax) || a1 > b1 || a2 > b2 || a2 < a1) return 0;
  return segments_two_spans(a1, b1, a2, b2, t0, dt, s0, s1);
}

KB_HD int landmark_ref_row_split(const SplitConst& sp, const CameraConst& cam, const double* vecs, const double* quats, const double* pairs,
                                 const double* ref_uv, double ref_t0, int r3_start, int r3_n, int so3_start, int so3_n, double rho, double* rec) {
  const double t = static_rs_time(cam, ref_t0, ref_uv[1]);
  Segment s; int ia, ib; double ua, ub;
  s.start = r3_start; s.n = r3_n;
  if (!segment_locate(s, t, sp.t0_r3, sp.dt_r3, ia, ua) || ia < 0 || ia + 3 >= sp.n_r3) return kStatusRange;
  s.start = so3_start; s.n = so3_n;
  if (!segment_locate(s, t, sp.t0_so3, sp.dt_so3, ib, ub) || ib < 0 || ib + 3 >= sp.n_so3) return kStatusRange;
  landmark_ref_split(cam, vecs + (size_t)ia * kVecStride, r3_basis(ua, sp.dt_r3), ia, quats + (size_t)ib * kQuatStride,
                     pairs + (size_t)(ib + 1) * kSo3PairStride, cumulative_basis(ub, sp.dt_so3), ib, ref_uv, rho, rec);
  return 0;
}

struct ObsForwardSplit { int status, ia, ib; Basis bs; BasisR3 br; M3 R; };
KB_HD void static_rs_row_forward_split(const SplitConst& sp, const CameraConst& cam, const double* quats, const double* pairs, const double* obs_uv,
                                       double obs_t0, double ref_t0, ObsForwardSplit& f) {
  f.status = kStatusRange; f.ia = -1; f.ib = -1;
  const double t = static_rs_time(cam, obs_t0, obs_uv[1]);
  Segment s0, s1; double ua = 0.0, ub = 0.0;
  int nseg = static_rs_segments_split(sp, cam, ref_t0, obs_t0, sp.t0_r3, sp.dt_r3, s0, s1);
  if (nseg == 0 || locate_in_segments(nseg, s0, s1, t, sp.t0_r3, sp.dt_r3, f.ia, ua) < 0) { f.ia = -1; return; }
  nseg = static_rs_segments_split(sp, cam, ref_t0, obs_t0, sp.t0_so3, sp.dt_so3, s0, s1);
  if (nseg == 0 || locate_in_segments(nseg, s0, s1, t, sp.t0_so3, sp.dt_so3, f.ib, ub) < 0) { f.ia = -1; f.ib = -1; return; }
  f.status = 0;
  f.br = r3_basis(ua, sp.dt_r3);
  f.bs = cumulative_basis(ub, sp.dt_so3);
  f.R = so3_forward(quats + (size_t)f.ib * kQuatStride, pairs + (size_t)(f.ib + 1) * kSo3PairStride, f.bs);
}
KB_HD int static_rs_row_finish_split(const CameraConst& cam, const double* vecs, const double* quats, const double* pairs, const ObsForwardSplit& f,
                                     const double* ref, const double* obs_uv, double weight, double huber_c, double* r, double* J, double* Jrho,
                                     int* idx /*[4]*/) {
  if (f.status != 0) return f.status;
  int ira, irb;
  static_rs_obs_split(cam, vecs + (size_t)f.ia * kVecStride, f.br, quats + (size_t)f.ib * kQuatStride, pairs + (size_t)(f.ib + 1) * kSo3PairStride,
                      f.bs, f.R, ref, obs_uv, weight, huber_c, r, J, Jrho, &ira, &irb);
  if (ira < 0) return kStatusRange;
  idx[0] = ira; idx[1] = f.ia; idx[2] = irb; idx[3] = f.ib;      // ref R3, obs R3, ref SO3, obs SO3
  return 0;
}


// =================================================================================================================
// Point queries: trajectory.position(t) / velocity / acceleration / orientation / angular_velocity on the WHOLE spline
// (python/src/kontiki/trajectories/trajectory_helper.h:12-34 -> trajectory.h:98-132; the owning entity is one segment).
// out[16] = position(3) | velocity(3) | acceleration(3) | orientation x,y,z,w (4) | angular velocity, world frame (3)
// =================================================================================================================
KB_HD void quat_mul(const double* a, const double* b, double* o) {
  const double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  const double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  const double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  const double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
// unit quaternion of Exp(theta), theta a rotation vector
KB_HD void quat_exp(V3 th, double* q) {
  const double x = dot(th, th);
  double k, c;
  if (x < 1e-8) { k = 0.5 - x / 48.0; c = 1.0 - x / 8.0 + x * x / 384.0; }
  else { const double a = sqrt(x); k = sin(0.5 * a) / a; c = cos(0.5 * a); }
  q[0] = k * th.x; q[1] = k * th.y; q[2] = k * th.z; q[3] = c;
}
// body twist of the cumulative SE3 spline and its time derivative (the forward half of accel_se3)
KB_HD void se3_body_twist(const double* p1, const double* p2, const double* p3, const Basis& bs, V3& vb, V3& wb, V3& dvb) {
  const V3 u1 = v3(p1[0], p1[1], p1[2]), f1 = v3(p1[3], p1[4], p1[5]);
  const V3 u2 = v3(p2[0], p2[1], p2[2]), f2 = v3(p2[3], p2[4], p2[5]);
  const V3 u3 = v3(p3[0], p3[1], p3[2]), f3 = v3(p3[3], p3[4], p3[5]);
  ExpPart e;
  const V3 s1u = bs.dB[0] * u1, s1w = bs.dB[0] * f1, d1u = bs.d2B[0] * u1, d1w = bs.d2B[0] * f1;
  exp_part(p2, bs.B[1], true, false, e);
  const V3 y2w = mul_t(e.E, s1w), y2u = mul_t(e.E, s1u - cross(e.a, s1w));
  const V3 z2w = mul_t(e.E, d1w), z2u = mul_t(e.E, d1u - cross(e.a, d1w));
  const V3 s2u = y2u + bs.dB[1] * u2, s2w = y2w + bs.dB[1] * f2;
  const V3 d2u = z2u - bs.dB[1] * (cross(f2, y2u) + cross(u2, y2w)) + bs.d2B[1] * u2;
  const V3 d2w = z2w - bs.dB[1] * cross(f2, y2w) + bs.d2B[1] * f2;
  exp_part(p3, bs.B[2], true, false, e);
  const V3 y3w = mul_t(e.E, s2w), y3u = mul_t(e.E, s2u - cross(e.a, s2w));
  const V3 z3u = mul_t(e.E, d2u - cross(e.a, d2w));
  vb = y3u + bs.dB[2] * u3; wb = y3w + bs.dB[2] * f3;
  dvb = z3u - bs.dB[2] * (cross(f3, y3u) + cross(u3, y3w)) + bs.d2B[2] * u3;
}
KB_HD int traj_eval_se3(const SplineConst& sp, const double* knots, const double* pairs, double t, double* out) {
  Segment s; s.start = 0; s.n = sp.n_knots;
  int i0; double u;
  if (!segment_locate(s, t, sp.t0, sp.dt, i0, u)) return kStatusRange;
  const Basis bs = cumulative_basis(u, sp.dt);
  const double* k0 = knots + (size_t)i0 * kKnotStride;
  const double* p1 = pairs + (size_t)(i0 + 1) * kPairStride; const double* p2 = p1 + kPairStride; const double* p3 = p2 + kPairStride;
  Pose P; pose_forward(k0, p1, p2, p3, bs, P);
  V3 vb, wb, dvb;
  se3_body_twist(p1, p2, p3, bs, vb, wb, dvb);
  const V3 v = P.R * vb, w = P.R * wb;
  V3 a;
  if (sp.compat_zero_dB) { Basis b0 = bs; b0.dB[0] = b0.dB[1] = b0.dB[2] = 0.0; V3 v0, w0, d0; se3_body_twist(p1, p2, p3, b0, v0, w0, d0); a = P.R * d0; }
  else a = P.R * (cross(wb, vb) + dvb);
  double q[4] = {k0[0], k0[1], k0[2], k0[3]}, e[4];
  quat_exp(bs.B[0] * v3(p1[3], p1[4], p1[5]), e); quat_mul(q, e, q);
  quat_exp(bs.B[1] * v3(p2[3], p2[4], p2[5]), e); quat_mul(q, e, q);
  quat_exp(bs.B[2] * v3(p3[3], p3[4], p3[5]), e); quat_mul(q, e, q);
  const double qn = 1.0 / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);     // Sophus renormalises
  out[0] = P.p.x; out[1] = P.p.y; out[2] = P.p.z; out[3] = v.x; out[4] = v.y; out[5] = v.z; out[6] = a.x; out[7] = a.y; out[8] = a.z;
  out[9] = q[0] * qn; out[10] = q[1] * qn; out[11] = q[2] * qn; out[12] = q[3] * qn; out[13] = w.x; out[14] = w.y; out[15] = w.z;
  return 0;
}
KB_HD int traj_eval_split(const SplitConst& sp, const double* vecs, const double* quats, const double* pairs, double t, double* out) {
  Segment s; int ia, ib; double ua, ub;
  s.start = 0; s.n = sp.n_r3;
  if (!segment_locate(s, t, sp.t0_r3, sp.dt_r3, ia, ua)) return kStatusRange;
  s.start = 0; s.n = sp.n_so3;
  if (!segment_locate(s, t, sp.t0_so3, sp.dt_so3, ib, ub)) return kStatusRange;
  const BasisR3 br = r3_basis(ua, sp.dt_r3);
  const Basis bs = cumulative_basis(ub, sp.dt_so3);
  const double* c0 = vecs + (size_t)ia * kVecStride;
  const V3 p = r3_combine(c0, br.Bp), v = r3_combine(c0, br.Bv), a = r3_combine(c0, br.Ba);
  const double* q0 = quats + (size_t)ib * kQuatStride;
  const double* p1 = pairs + (size_t)(ib + 1) * kSo3PairStride; const double* p2 = p1 + kSo3PairStride; const double* p3 = p2 + kSo3PairStride;
  ExpPart e2, e3;
  so3_exp_part(p2, bs.B[1], e2); so3_exp_part(p3, bs.B[2], e3);
  const V3 f1 = v3(2.0 * p1[0], 2.0 * p1[1], 2.0 * p1[2]), f2 = v3(2.0 * p2[0], 2.0 * p2[1], 2.0 * p2[2]), f3 = v3(2.0 * p3[0], 2.0 * p3[1], 2.0 * p3[2]);
  const V3 wb = mul_t(e3.E, mul_t(e2.E, bs.dB[0] * f1) + bs.dB[1] * f2) + bs.dB[2] * f3;
  const V3 w = so3_forward(q0, p1, bs) * wb;
  double q[4] = {q0[0], q0[1], q0[2], q0[3]}, e[4];
  quat_exp(bs.B[0] * f1, e); quat_mul(q, e, q);
  quat_exp(bs.B[1] * f2, e); quat_mul(q, e, q);
  quat_exp(bs.B[2] * f3, e); quat_mul(q, e, q);                                             // Eigen: no renormalisation
  out[0] = p.x; out[1] = p.y; out[2] = p.z; out[3] = v.x; out[4] = v.y; out[5] = v.z; out[6] = a.x; out[7] = a.y; out[8] = a.z;
  out[9] = q[0]; out[10] = q[1]; out[11] = q[2]; out[12] = q[3]; out[13] = w.x; out[14] = w.y; out[15] = w.z;
  return 0;
}

}  // namespace kb

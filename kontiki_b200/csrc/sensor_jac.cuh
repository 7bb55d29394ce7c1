// kontiki_b200 -- Jacobians with respect to the SENSOR parameter blocks: time offset, camera
// relative pose.  In the reference every residual also carries the sensor's blocks q_ct(4), p_ct(3), time_offset(1)
// [ConstantBiasImu: + accelerometer_bias(3), gyroscope_bias(3)] (sensors/sensors.h:135-165, constant_bias_imu.h:100-119);
// they are constant by default (sensors.h:93-95) and Ceres then passes jacobians[k] == NULL for them, which is why the hot
// kernels do not produce them.  When a block is unlocked these cold kernels add its columns; same __host__ __device__
// arrangement as spline_math.cuh (tests/host_check.cpp compiles this text for the host).
//
//   * time offset d: t_eval = t + d (imu.h:49, static_rscamera_measurement.h:32-33) and floor() drops the derivative
//     (spline_base.h:155-163), so d r / d d is the time derivative of the residual at fixed knots.
//   * IMU relative pose: not applied by the reference (TODO.md:6) => those columns are identically zero.
//   * ConstantBiasImu biases: r = w (y - (model + bias)) (constant_bias_imu.h:52-61) => d r / d bias = -w I, not computed here.
#pragma once
#include "split_math.cuh"

namespace kb {

struct Tw { V3 u, w; };      // twist [upsilon; phi]
KB_HD Tw tw_scale(double s, const double* om) { Tw t; t.u = v3(s * om[0], s * om[1], s * om[2]); t.w = v3(s * om[3], s * om[4], s * om[5]); return t; }
KB_HD Tw tw_add(Tw a, Tw b) { Tw t; t.u = a.u + b.u; t.w = a.w + b.w; return t; }
KB_HD Tw tw_axpy(double s, Tw a, Tw b) { Tw t; t.u = s * a.u + b.u; t.w = s * a.w + b.w; return t; }
KB_HD Tw tw_Adinv(const M3& E, V3 a, Tw x) { Tw t; t.w = mul_t(E, x.w); t.u = mul_t(E, x.u - cross(a, x.w)); return t; }
KB_HD Tw tw_ad(const double* om, Tw x) {
  const V3 ou = v3(om[0], om[1], om[2]), ow = v3(om[3], om[4], om[5]);
  Tw t; t.u = cross(ow, x.u) + cross(ou, x.w); t.w = cross(ow, x.w); return t; }

// Body twist s of the cumulative SE3 spline (P' = P s^) and its first two time derivatives.
//   y = Ad(A_j^-1) s, z = Ad(A_j^-1) s', q = Ad(A_j^-1) s'':
//   s <- y + dB w;  s' <- z - dB ad(w) y + d2B w;  s'' <- q - 2 dB ad(w) z - d2B ad(w) y + dB^2 ad(w)^2 y + d3B w
KB_HD void se3_body_twist3(const double* p1, const double* p2, const double* p3, const Basis& bs, double dt, Tw& s, Tw& ds, Tw& dds) {
  const double di3 = 1.0 / (dt * dt * dt);
  const double d3B[3] = {di3, -2.0 * di3, di3};
  s = tw_scale(bs.dB[0], p1); ds = tw_scale(bs.d2B[0], p1); dds = tw_scale(d3B[0], p1);
  const double* pj[2] = {p2, p3};
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    ExpPart e;
    exp_part(pj[j], bs.B[j + 1], true, false, e);
    const Tw y = tw_Adinv(e.E, e.a, s), z = tw_Adinv(e.E, e.a, ds), q = tw_Adinv(e.E, e.a, dds);
    const Tw a1 = tw_ad(pj[j], y), a2 = tw_ad(pj[j], z), a3 = tw_ad(pj[j], a1);
    const double dB = bs.dB[j + 1], d2B = bs.d2B[j + 1];
    s = tw_add(y, tw_scale(dB, pj[j]));
    ds = tw_add(tw_axpy(-dB, a1, z), tw_scale(d2B, pj[j]));
    dds = tw_add(tw_axpy(dB * dB, a3, tw_axpy(-d2B, a1, tw_axpy(-2.0 * dB, a2, q))), tw_scale(d3B[j + 1], pj[j]));
  }
}

// UniformSE3SplineTrajectory.evaluate(t) of the reference's Python API (py_uniform_se3_spline_trajectory.cc:53-60): the 4x4
// matrices P, P', P'' of EvaluateSpline with every flag set (uniform_se3_spline_trajectory.h:101-194):
//   P' = P s^,   P'' = P (s^ s^ + s'^)     with the body twist s = [v; w] and hat([v; w]) = [[hat w, v], [0, 0]].
// out[48] = P | P' | P'' row-major.
KB_HD int traj_eval_se3_matrices(const SplineConst& sp, const double* knots, const double* pairs, double t, double* out) {
  Segment sg; sg.start = 0; sg.n = sp.n_knots;
  int i0; double u;
  if (!segment_locate(sg, t, sp.t0, sp.dt, i0, u)) return kStatusRange;
  const Basis bs = cumulative_basis(u, sp.dt);
  const double* p1 = pairs + (size_t)(i0 + 1) * kPairStride; const double* p2 = p1 + kPairStride; const double* p3 = p2 + kPairStride;
  Pose P; pose_forward(knots + (size_t)i0 * kKnotStride, p1, p2, p3, bs, P);
  Tw s, ds, dds;
  se3_body_twist3(p1, p2, p3, bs, sp.dt, s, ds, dds);
  const M3 hw = hat(s.w);
  const M3 R1 = P.R * hw, R2 = P.R * (hw * hw + hat(ds.w));
  const V3 t1 = P.R * s.u, t2 = P.R * (cross(s.w, s.u) + ds.u);
  for (int i = 0; i < 48; ++i) out[i] = 0.0;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) { out[4 * i + j] = P.R.a[3 * i + j]; out[16 + 4 * i + j] = R1.a[3 * i + j]; out[32 + 4 * i + j] = R2.a[3 * i + j]; }
  }
  out[3] = P.p.x; out[7] = P.p.y; out[11] = P.p.z; out[15] = 1.0;
  out[16 + 3] = t1.x; out[16 + 7] = t1.y; out[16 + 11] = t1.z;
  out[32 + 3] = t2.x; out[32 + 7] = t2.y; out[32 + 11] = t2.z;
  return 0;
}

// d r / d time_offset of a gyroscope (which = 0) / accelerometer (which = 1) row on SE3; out[3].
// Returns a status like imu_row.  compat_zero_dB accelerometer rows are refused (the reference's Jet path differentiates
// an expression whose dB was never assigned; its time derivative is not the derivative of the intended accelerometer).
KB_HD int imu_time_offset_jac_se3(int which, const SplineConst& sp, const ImuConst& imu, const double* knots, const double* pairs, double t,
                                  double weight, double* out) {
  double ta = t, tb = t;
  if (!imu.time_offset_locked) { ta = sub_rn(t, imu.max_time_offset); tb = add_rn(t, imu.max_time_offset); }
  if (sp.n_knots < 4 || !(ta >= sp.t0) || !(tb < spline_max_time(sp))) return kStatusRange;
  Segment seg; segments_one_span(ta, tb, sp.t0, sp.dt, seg);
  int i0; double u;
  if (!segment_locate(seg, add_rn(t, imu.time_offset), sp.t0, sp.dt, i0, u)) return kStatusRange;
  const Basis bs = cumulative_basis(u, sp.dt);
  const double* p1 = pairs + (size_t)(i0 + 1) * kPairStride; const double* p2 = p1 + kPairStride; const double* p3 = p2 + kPairStride;
  Tw s, ds, dds;
  se3_body_twist3(p1, p2, p3, bs, sp.dt, s, ds, dds);
  V3 d;
  if (which == 0) d = ds.w;                                   // d w_b / dt
  else {
    if (sp.compat_zero_dB) return -5;
    Pose P; pose_forward(knots + (size_t)i0 * kKnotStride, p1, p2, p3, bs, P);
    const V3 gb = mul_t(P.R, v3(0.0, 0.0, -kGravity));
    d = cross(ds.w, s.u) + cross(s.w, ds.u) + dds.u - cross(s.w, gb);       // d/dt (w x v + v' + R^T g)
  }
  out[0] = -weight * d.x; out[1] = -weight * d.y; out[2] = -weight * d.z;
  return 0;
}

// ... on a split trajectory: gyro needs d w_b/dt of the SO3 spline, the accelerometer R^T (a + g) needs the jerk of the R3 spline.
KB_HD int imu_time_offset_jac_split(int which, const SplitConst& sp, const ImuConst& imu, const double* vecs, const double* quats, const double* pairs,
                                    double t, double weight, double* out) {
  double ta = t, tb = t;
  if (!imu.time_offset_locked) { ta = sub_rn(t, imu.max_time_offset); tb = add_rn(t, imu.max_time_offset); }
  if (sp.n_r3 < 4 || sp.n_so3 < 4 || !(ta >= split_min_time(sp)) || !(tb < split_max_time(sp))) return kStatusRange;
  const double te = add_rn(t, imu.time_offset);
  Segment seg; int ib; double ub;
  segments_one_span(ta, tb, sp.t0_so3, sp.dt_so3, seg);
  if (!segment_locate(seg, te, sp.t0_so3, sp.dt_so3, ib, ub)) return kStatusRange;
  const Basis bs = cumulative_basis(ub, sp.dt_so3);
  const double* q0 = quats + (size_t)ib * kQuatStride;
  const double* p1 = pairs + (size_t)(ib + 1) * kSo3PairStride; const double* p2 = p1 + kSo3PairStride; const double* p3 = p2 + kSo3PairStride;
  // rotation-only twist recursion with rotation vectors 2 phi
  const V3 f1 = v3(2.0 * p1[0], 2.0 * p1[1], 2.0 * p1[2]), f2 = v3(2.0 * p2[0], 2.0 * p2[1], 2.0 * p2[2]), f3 = v3(2.0 * p3[0], 2.0 * p3[1], 2.0 * p3[2]);
  ExpPart e2, e3;
  so3_exp_part(p2, bs.B[1], e2); so3_exp_part(p3, bs.B[2], e3);
  const V3 y2 = mul_t(e2.E, bs.dB[0] * f1), z2 = mul_t(e2.E, bs.d2B[0] * f1);
  const V3 s2 = y2 + bs.dB[1] * f2, d2 = z2 - bs.dB[1] * cross(f2, y2) + bs.d2B[1] * f2;
  const V3 y3 = mul_t(e3.E, s2), z3 = mul_t(e3.E, d2);
  const V3 wb = y3 + bs.dB[2] * f3, dwb = z3 - bs.dB[2] * cross(f3, y3) + bs.d2B[2] * f3;
  V3 d;
  if (which == 0) d = dwb;
  else {
    int ia; double ua;
    segments_one_span(ta, tb, sp.t0_r3, sp.dt_r3, seg);
    if (!segment_locate(seg, te, sp.t0_r3, sp.dt_r3, ia, ua)) return kStatusRange;
    const BasisR3 br = r3_basis(ua, sp.dt_r3);
    const double di3 = 1.0 / (sp.dt_r3 * sp.dt_r3 * sp.dt_r3);
    const double Bj[4] = {-di3, 3.0 * di3, -3.0 * di3, di3};             // third derivative of the R3 basis (constant on an interval)
    const double* c0 = vecs + (size_t)ia * kVecStride;
    const M3 R = so3_forward(q0, p1, bs);
    const V3 ab = mul_t(R, r3_combine(c0, br.Ba) + v3(0.0, 0.0, -kGravity));
    d = mul_t(R, r3_combine(c0, Bj)) - cross(wb, ab);                   // d/dt R^T (a + g) = R^T a' - w_b x R^T (a + g)
  }
  out[0] = -weight * d.x; out[1] = -weight * d.y; out[2] = -weight * d.z;
  return 0;
}

// Sensor-block Jacobians of a static-RS camera row on SE3: out[16] = d r/d q_ct (2x4) | d r/d p_ct (2x3) | d r/d time_offset (2x1) | pad(2)
// (the layout of Ceres' jacobians[] for the three camera blocks, row-major; huber_c > 0 applies the same Corrector as the knot blocks).
KB_HD int static_rs_sensor_jac_se3(const SplineConst& sp, const CameraConst& cam, const double* knots, const double* pairs, const double* obs_uv,
                                   double obs_t0, const double* ref_uv, double ref_t0, double rho, double weight, double huber_c, double* out) {
  Segment s0, s1;
  const int nseg = static_rs_segments(sp, cam, ref_t0, obs_t0, s0, s1);
  if (nseg == 0) return kStatusRange;
  int ir, io; double ur, uo;
  if (locate_in_segments(nseg, s0, s1, static_rs_time(cam, ref_t0, ref_uv[1]), sp.t0, sp.dt, ir, ur) < 0) return kStatusRange;
  if (locate_in_segments(nseg, s0, s1, static_rs_time(cam, obs_t0, obs_uv[1]), sp.t0, sp.dt, io, uo) < 0) return kStatusRange;
  const Basis br = cumulative_basis(ur, sp.dt), bo = cumulative_basis(uo, sp.dt);
  const double* r1 = pairs + (size_t)(ir + 1) * kPairStride; const double* o1 = pairs + (size_t)(io + 1) * kPairStride;
  Pose Pr, Po;
  pose_forward(knots + (size_t)ir * kKnotStride, r1, r1 + kPairStride, r1 + 2 * kPairStride, br, Pr);
  pose_forward(knots + (size_t)io * kKnotStride, o1, o1 + kPairStride, o1 + 2 * kPairStride, bo, Po);
  Tw sr, so, t1, t2;
  se3_body_twist3(r1, r1 + kPairStride, r1 + 2 * kPairStride, br, sp.dt, sr, t1, t2);
  se3_body_twist3(o1, o1 + kPairStride, o1 + 2 * kPairStride, bo, sp.dt, so, t1, t2);
  const M3 Rct = quat_to_rot(cam.q_ct[0], cam.q_ct[1], cam.q_ct[2], cam.q_ct[3]);
  const V3 pct = v3(cam.p_ct[0], cam.p_ct[1], cam.p_ct[2]);
  const V3 yv = camera_unproject(cam, ref_uv[0], ref_uv[1]) - rho * pct;
  const V3 Xref = mul_t(Rct, yv);
  const V3 X = Pr.R * Xref + rho * Pr.p;
  const V3 Xw = X - rho * Po.p;
  const V3 Xobs = mul_t(Po.R, Xw);
  const V3 RX = Rct * Xobs;
  double y0, y1;
  Mr<2> Jp0, Gc;
  camera_project_jac(cam, RX + rho * pct, y0, y1, Jp0);
  const double r0 = weight * (obs_uv[0] - y0), rr1 = weight * (obs_uv[1] - y1);
  double c00 = 1.0, c01 = 0.0, c10 = 0.0, c11 = 1.0;
  if (huber_c > 0.0) {
    const HuberScale h = huber_scale(huber_c, r0 * r0 + rr1 * rr1);
    c00 = h.sqrt_rho1 * (1.0 - h.alpha_sq_norm * r0 * r0); c01 = -h.sqrt_rho1 * h.alpha_sq_norm * r0 * rr1;
    c10 = c01; c11 = h.sqrt_rho1 * (1.0 - h.alpha_sq_norm * rr1 * rr1);
  }
  for (int c = 0; c < 3; ++c) { Gc.a[c] = -weight * (c00 * Jp0.a[c] + c01 * Jp0.a[3 + c]); Gc.a[3 + c] = -weight * (c10 * Jp0.a[c] + c11 * Jp0.a[3 + c]); }
  const Mr<2> Go = rmul(Gc, Rct), GX = rmul_nt(Go, Po.R), GXR = rmul(GX, Pr.R);
  // time offset: both evaluation times move with d
  const V3 dXdt = Pr.R * (cross(sr.w, Xref) + rho * sr.u);
  const V3 dpo = Po.R * so.u;
  const Mr<2> Gth = rmul_hat(Go, Xobs);
  // relative pose.  tangent (R_ct <- R_ct Exp(d)): dXc = -R_ct hat(Xobs) d + R_ct R_o^T R_r hat(Xref) d
  const Mr<2> Jth = rsub(rmul_hat(GXR, Xref), rmul_hat(Go, Xobs));
  // radial (Eigen's polynomial q*v): d/ds q_ct*Xobs = 2 (R_ct Xobs - Xobs),  d/ds conj(q_ct)*y = 2 (Xref - y)
  const V3 radc = 2.0 * (RX - Xobs), radr = 2.0 * (Xref - yv);
  const V3 v = v3(cam.q_ct[0], cam.q_ct[1], cam.q_ct[2]); const double w = cam.q_ct[3];
  const Mr<2> GXRRct = rmul_nt(GXR, Rct);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const V3 g = rrow(Jth, r);
    const V3 tq = 2.0 * (w * g - cross(g, v));
    const double rad = dot(rrow(Gc, r), radc) + dot(rrow(GXR, r), radr);
    out[r * 4 + 0] = tq.x + rad * cam.q_ct[0]; out[r * 4 + 1] = tq.y + rad * cam.q_ct[1]; out[r * 4 + 2] = tq.z + rad * cam.q_ct[2];
    out[r * 4 + 3] = -2.0 * dot(g, v) + rad * cam.q_ct[3];
    // d Xc / d p_ct = rho (I - R_ct R_o^T R_r R_ct^T)
    out[8 + r * 3 + 0] = rho * (Gc.a[3 * r] - GXRRct.a[3 * r]); out[8 + r * 3 + 1] = rho * (Gc.a[3 * r + 1] - GXRRct.a[3 * r + 1]);
    out[8 + r * 3 + 2] = rho * (Gc.a[3 * r + 2] - GXRRct.a[3 * r + 2]);
    out[14 + r] = dot(rrow(GX, r), dXdt) - rho * dot(rrow(GX, r), dpo) + dot(rrow(Gth, r), so.w);
  }
  return 0;
}

// Pose, body angular velocity and WORLD linear velocity of a split trajectory at located (ia, ua) / (ib, ub):
// R from the SO3 spline (uniform_so3_spline_trajectory.h:96-104), w_b from its derivative (:108-121, rotation vectors 2 phi), p and p' from
// the R3 spline (uniform_r3_spline_trajectory.h:62-92).
KB_HD void split_pose_twist(const SplitConst& sp, const double* vecs, const double* quats, const double* pairs, int ia, double ua, int ib, double ub,
                            M3& R, V3& p, V3& wb, V3& vw) {
  const Basis bs = cumulative_basis(ub, sp.dt_so3);
  const BasisR3 br = r3_basis(ua, sp.dt_r3);
  const double* q0 = quats + (size_t)ib * kQuatStride;
  const double* p1 = pairs + (size_t)(ib + 1) * kSo3PairStride; const double* p2 = p1 + kSo3PairStride; const double* p3 = p2 + kSo3PairStride;
  const V3 f1 = v3(2.0 * p1[0], 2.0 * p1[1], 2.0 * p1[2]), f2 = v3(2.0 * p2[0], 2.0 * p2[1], 2.0 * p2[2]), f3 = v3(2.0 * p3[0], 2.0 * p3[1], 2.0 * p3[2]);
  ExpPart e2, e3;
  so3_exp_part(p2, bs.B[1], e2); so3_exp_part(p3, bs.B[2], e3);
  const V3 s2 = mul_t(e2.E, bs.dB[0] * f1) + bs.dB[1] * f2;
  wb = mul_t(e3.E, s2) + bs.dB[2] * f3;
  R = so3_forward(q0, p1, bs);
  const double* c0 = vecs + (size_t)ia * kVecStride;
  p = r3_combine(c0, br.Bp); vw = r3_combine(c0, br.Bv);
}

// Sensor-block Jacobians of a static-RS camera row on a SPLIT trajectory (sensors.h:135-165 blocks, split_trajectory.h:41-58 evaluation):
// same output layout and the same chain as static_rs_sensor_jac_se3; the pose derivative is (w_b, world velocity) of the two splines.
KB_HD int static_rs_sensor_jac_split(const SplitConst& sp, const CameraConst& cam, const double* vecs, const double* quats, const double* pairs,
                                     const double* obs_uv, double obs_t0, const double* ref_uv, double ref_t0, double rho, double weight, double huber_c, double* out) {
  Segment s0, s1;
  int ira, irb, ioa, iob; double ura, urb, uoa, uob;
  const double tr = static_rs_time(cam, ref_t0, ref_uv[1]), to = static_rs_time(cam, obs_t0, obs_uv[1]);
  int nseg = static_rs_segments_split(sp, cam, ref_t0, obs_t0, sp.t0_r3, sp.dt_r3, s0, s1);
  if (nseg == 0 || locate_in_segments(nseg, s0, s1, tr, sp.t0_r3, sp.dt_r3, ira, ura) < 0 || locate_in_segments(nseg, s0, s1, to, sp.t0_r3, sp.dt_r3, ioa, uoa) < 0) return kStatusRange;
  nseg = static_rs_segments_split(sp, cam, ref_t0, obs_t0, sp.t0_so3, sp.dt_so3, s0, s1);
  if (nseg == 0 || locate_in_segments(nseg, s0, s1, tr, sp.t0_so3, sp.dt_so3, irb, urb) < 0 || locate_in_segments(nseg, s0, s1, to, sp.t0_so3, sp.dt_so3, iob, uob) < 0) return kStatusRange;
  M3 Rr, Ro; V3 pr, po, wr, wo, vr, vo;
  split_pose_twist(sp, vecs, quats, pairs, ira, ura, irb, urb, Rr, pr, wr, vr);
  split_pose_twist(sp, vecs, quats, pairs, ioa, uoa, iob, uob, Ro, po, wo, vo);
  const M3 Rct = quat_to_rot(cam.q_ct[0], cam.q_ct[1], cam.q_ct[2], cam.q_ct[3]);
  const V3 pct = v3(cam.p_ct[0], cam.p_ct[1], cam.p_ct[2]);
  const V3 yv = camera_unproject(cam, ref_uv[0], ref_uv[1]) - rho * pct;
  const V3 Xref = mul_t(Rct, yv);
  const V3 X = Rr * Xref + rho * pr;
  const V3 Xobs = mul_t(Ro, X - rho * po);
  const V3 RX = Rct * Xobs;
  double y0, y1;
  Mr<2> Jp0, Gc;
  camera_project_jac(cam, RX + rho * pct, y0, y1, Jp0);
  const double r0 = weight * (obs_uv[0] - y0), rr1 = weight * (obs_uv[1] - y1);
  double c00 = 1.0, c01 = 0.0, c10 = 0.0, c11 = 1.0;
  if (huber_c > 0.0) {
    const HuberScale h = huber_scale(huber_c, r0 * r0 + rr1 * rr1);
    c00 = h.sqrt_rho1 * (1.0 - h.alpha_sq_norm * r0 * r0); c01 = -h.sqrt_rho1 * h.alpha_sq_norm * r0 * rr1;
    c10 = c01; c11 = h.sqrt_rho1 * (1.0 - h.alpha_sq_norm * rr1 * rr1);
  }
  for (int c = 0; c < 3; ++c) { Gc.a[c] = -weight * (c00 * Jp0.a[c] + c01 * Jp0.a[3 + c]); Gc.a[3 + c] = -weight * (c10 * Jp0.a[c] + c11 * Jp0.a[3 + c]); }
  const Mr<2> Go = rmul(Gc, Rct), GX = rmul_nt(Go, Ro), GXR = rmul(GX, Rr);
  const V3 dXdt = Rr * cross(wr, Xref) + rho * vr;      // d/dt (R_r Xref + rho p_r)
  const Mr<2> Gth = rmul_hat(Go, Xobs);
  const Mr<2> Jth = rsub(rmul_hat(GXR, Xref), rmul_hat(Go, Xobs));
  const V3 radc = 2.0 * (RX - Xobs), radr = 2.0 * (Xref - yv);
  const V3 v = v3(cam.q_ct[0], cam.q_ct[1], cam.q_ct[2]); const double w = cam.q_ct[3];
  const Mr<2> GXRRct = rmul_nt(GXR, Rct);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const V3 g = rrow(Jth, r);
    const V3 tq = 2.0 * (w * g - cross(g, v));
    const double rad = dot(rrow(Gc, r), radc) + dot(rrow(GXR, r), radr);
    out[r * 4 + 0] = tq.x + rad * cam.q_ct[0]; out[r * 4 + 1] = tq.y + rad * cam.q_ct[1]; out[r * 4 + 2] = tq.z + rad * cam.q_ct[2];
    out[r * 4 + 3] = -2.0 * dot(g, v) + rad * cam.q_ct[3];
    out[8 + r * 3 + 0] = rho * (Gc.a[3 * r] - GXRRct.a[3 * r]); out[8 + r * 3 + 1] = rho * (Gc.a[3 * r + 1] - GXRRct.a[3 * r + 1]);
    out[8 + r * 3 + 2] = rho * (Gc.a[3 * r + 2] - GXRRct.a[3 * r + 2]);
    out[14 + r] = dot(rrow(GX, r), dXdt) - rho * dot(rrow(GX, r), vo) + dot(rrow(Gth, r), wo);
  }
  return 0;
}

}  // namespace kb

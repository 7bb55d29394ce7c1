// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
//
// CPU restatement of the hovren/kontiki residual/Jacobian hot path, templated on the scalar
// exactly like the reference (T = double for residuals, T = Dual<4> for the Jacobian passes of
// ceres::DynamicAutoDiffCostFunction).  Every function cites the reference file:line it follows;
// paths are relative to /root/reference/cpplib/include/kontiki/.
//
// PARITY STATUS.  The reference holds no golden vectors for this path and cannot be built here (Ceres 1.x,
// Sophus, Eigen absent), so this restatement is pinned from two independent sides, values AND Jacobians:
//   * the reference's own property tests, restated on its own fixture knots (tests/test_oracle_pinning.py, tests/fixtures_ref.py);
//   * a 60-digit mpmath transcription of the reference headers written from the published Sophus / Eigen formulas, sharing no code
//     with this file (tests/mp_reference.py): values directly, Jacobians by central differences of the 60-digit residual, SE3 poses
//     also by 4x4 expm / logm (tests/test_oracle_independent.py; committed vectors tests/golden/mp_v1.npz, tests/test_mp_golden.py).
//     Covered: SE3 / split trajectories x gyroscope / accelerometer / static-RS rows (the hot path) and the widening rows -- AtanCamera,
//     NewtonRs (value of the iteration and the derivative through it), LiftingRs (incl. the row-time column), Position / Orientation.
// What stays an assumption is the semantics of the un-vendored dependencies themselves.
#pragma once
#include <memory>
#include <sstream>
#include <stdexcept>
#include <vector>

#include "lie.hpp"

namespace kto {

// trajectories/trajectory.h:17-23
enum EvaluationFlags { EvalPosition = 1, EvalVelocity = 2, EvalAcceleration = 4, EvalOrientation = 8, EvalAngularVelocity = 16 };

// trajectories/trajectory.h:26-80
template <class T>
struct TrajectoryEvaluation {
  explicit TrajectoryEvaluation(int f) : flags(f) {}
  int flags;
  Vec3<T> position, velocity, acceleration, angular_velocity;
  Quat<T> orientation;
  bool Position() const { return flags & EvalPosition; }
  bool Velocity() const { return flags & EvalVelocity; }
  bool Acceleration() const { return flags & EvalAcceleration; }
  bool Orientation() const { return flags & EvalOrientation; }
  bool AngularVelocity() const { return flags & EvalAngularVelocity; }
  int FlagsLinear() const { return flags & (EvalPosition | EvalVelocity | EvalAcceleration); }
  int FlagsRotation() const { return flags & (EvalOrientation | EvalAngularVelocity); }
};
template <class T> using Result = std::unique_ptr<TrajectoryEvaluation<T>>;

// trajectories/spline_base.h:18-28 (row-vector * matrix convention: B = U^T * M)
static const double kM[4][4] = {{1. / 6., 4. / 6., 1. / 6., 0}, {-3. / 6., 0, 3. / 6., 0}, {3. / 6., -6. / 6, 3. / 6., 0}, {-1. / 6., 3. / 6., -3. / 6., 1. / 6.}};
static const double kMcumul[4][4] = {{6. / 6., 5. / 6., 1. / 6., 0}, {0. / 6., 3. / 6., 3. / 6., 0}, {0. / 6., -3. / 6., 3. / 6., 0}, {0. / 6., 1. / 6., -2. / 6., 1. / 6.}};
template <class T> inline void basis_mul(const T U[4], const double M[4][4], T B[4]) {
  for (int j = 0; j < 4; ++j) { T s = U[0] * T(M[0][j]); for (int i = 1; i < 4; ++i) s += U[i] * T(M[i][j]); B[j] = s; } }

// trajectories/spline_base.h:30-62
struct SplineSegmentMeta {
  double t0 = 0.0, dt = 1.0; size_t n = 0;
  void Validate() const { if (n < 4) throw std::range_error("Spline had too few control points"); }
  double MinTime() const { Validate(); return t0; }
  double MaxTime() const { Validate(); return t0 + (n - 3) * dt; }
};
// trajectories/spline_base.h:64-93
struct SplineMeta {
  std::vector<SplineSegmentMeta> segments;
  size_t NumParameters() const { size_t n = 0; for (auto& s : segments) n += s.n; return n; }
};

enum SplineKind { kSE3 = 0, kSO3 = 1, kR3 = 2 };

struct EvalOptions { bool compat_zero_dB = false; };

// ---- one segment view (trajectories/spline_base.h:108-166) ---------------------------------------
template <class T>
struct SegmentView {
  SplineSegmentMeta meta; T const* const* params;  // params[i] -> knot i of this segment
  SplineKind kind; EvalOptions opt;
  // spline_base.h:148-163: floor is taken on the scalar part ("PotentiallyUnsafeFloor")
  void CalculateIndexAndInterpolationAmount(T t, int& i0, T& u) const {
    T s = (t - T(meta.t0)) / T(meta.dt);
    i0 = static_cast<int>(std::floor(val(s)));
    u = s - T(i0);
  }
  void CheckRange(T t, int i0) const {  // uniform_se3_spline_trajectory.h:121-127 (same in r3 :44-49, so3 :69-74)
    const size_t N = meta.n;
    if ((N < 4) || (i0 < 0) || (size_t(i0) > (N - 4))) {
      std::stringstream ss; ss << "t=" << val(t) << " i0=" << i0 << " is out of range for spline with ncp=" << N;
      throw std::range_error(ss.str());
    }
  }
  SE3<T> ControlPointSE3(int i) const { const T* p = params[i]; return {{p[0], p[1], p[2], p[3]}, {p[4], p[5], p[6]}}; }
  Quat<T> ControlPointQuat(int i) const { const T* p = params[i]; return {p[0], p[1], p[2], p[3]}; }
  Vec3<T> ControlPointVec(int i) const { const T* p = params[i]; return {p[0], p[1], p[2]}; }

  // trajectories/uniform_se3_spline_trajectory.h:101-194
  void EvaluateSplineSE3(T t, int flags, SE3<T>& P, Mat4<T>& P_prim, Mat4<T>& P_bis) const {
    auto result = std::make_unique<TrajectoryEvaluation<T>>(flags);  // :102 (throw-away, only to read flags)
    int num_derivatives;
    if (result->Acceleration()) num_derivatives = 2;
    else if (result->Velocity() || result->AngularVelocity()) num_derivatives = 1;
    else num_derivatives = 0;
    int i0; T u;
    CalculateIndexAndInterpolationAmount(t, i0, u);
    CheckRange(t, i0);
    T U[4], dU[4], d2U[4], B[4], dB[4], d2B[4];
    for (int i = 0; i < 4; ++i) { dB[i] = T(0.0); d2B[i] = T(0.0); }   // Jet() default == 0 (compat_zero_dB path)
    T u2 = kpow(u, 2.0), u3 = kpow(u, 3.0);
    T dt_inv = T(1.0) / T(meta.dt);
    U[0] = T(1.0); U[1] = u; U[2] = u2; U[3] = u3;
    basis_mul(U, kMcumul, B);
    // :138-141 assigns dB only for Velocity|AngularVelocity although :166-169 reads it whenever
    // num_derivatives >= 1 (accelerometer flags are Orientation|Acceleration, sensors/imu.h:57).
    // Default here = intended math (always assign); compat_zero_dB reproduces the Jet path (dB == 0).
    const bool assign_dB = opt.compat_zero_dB ? (result->AngularVelocity() || result->Velocity()) : (num_derivatives >= 1);
    if (assign_dB) {
      dU[0] = dt_inv * T(0.0); dU[1] = dt_inv * T(1.0); dU[2] = dt_inv * (T(2.0) * u); dU[3] = dt_inv * (T(3.0) * u2);
      basis_mul(dU, kMcumul, dB);
    }
    if (result->Acceleration()) {
      T s = kpow(dt_inv, 2.0);
      d2U[0] = s * T(0.0); d2U[1] = s * T(0.0); d2U[2] = s * T(2.0); d2U[3] = s * (T(6.0) * u);
      basis_mul(d2U, kMcumul, d2B);
    }
    Mat4<T> A[3], A_prim[3], A_bis[3], Aj_prim;
    P = ControlPointSE3(i0);
    const int K = i0 + 4;
    for (int i = i0 + 1; i < K; ++i) {
      int j = i - i0;
      SE3<T> Pa = ControlPointSE3(i - 1), Pb = ControlPointSE3(i);
      Vec6<T> omega = se3_log(se3_mul(se3_inverse(Pa), Pb));
      Mat4<T> omega_hat = se3_hat(omega);
      Vec6<T> scaled; for (int k = 0; k < 6; ++k) scaled.d[k] = B[j] * omega.d[k];
      SE3<T> Aj = se3_exp(scaled);
      P = se3_mul(P, Aj);
      if (num_derivatives >= 1) {
        A[j - 1] = se3_matrix(Aj);
        Aj_prim = (se3_matrix(Aj) * omega_hat) * dB[j];
        A_prim[j - 1] = Aj_prim;
      }
      if (num_derivatives >= 2) {
        A_bis[j - 1] = (Aj_prim * omega_hat) * dB[j] + (se3_matrix(Aj) * omega_hat) * d2B[j];
      }
    }
    SE3<T> P0 = ControlPointSE3(i0);
    if (num_derivatives >= 1) {
      Mat4<T> M1 = A_prim[0] * A[1] * A[2] + A[0] * A_prim[1] * A[2] + A[0] * A[1] * A_prim[2];
      P_prim = se3_matrix(P0) * M1;
    }
    if (num_derivatives >= 2) {
      Mat4<T> M2 = A_bis[0] * A[1] * A[2] + A[0] * A_bis[1] * A[2] + A[0] * A[1] * A_bis[2] +
                   T(2.0) * A_prim[0] * A_prim[1] * A[2] + T(2.0) * A_prim[0] * A[1] * A_prim[2] +
                   T(2.0) * A[0] * A_prim[1] * A_prim[2];
      P_bis = se3_matrix(P0) * M2;
    }
  }

  // trajectories/uniform_se3_spline_trajectory.h:81-99
  Result<T> EvaluateSE3(T t, int flags) const {
    SE3<T> P; Mat4<T> P_prim = mat4_zero<T>(), P_bis = mat4_zero<T>();
    EvaluateSplineSE3(t, flags, P, P_prim, P_bis);
    auto result = std::make_unique<TrajectoryEvaluation<T>>(flags);
    result->position = P.t;
    result->velocity = {P_prim.m[0][3], P_prim.m[1][3], P_prim.m[2][3]};
    result->acceleration = {P_bis.m[0][3], P_bis.m[1][3], P_bis.m[2][3]};
    result->orientation = P.q;
    Mat3<T> Pp, Rt = transpose(qmat(P.q));
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Pp.m[i][j] = P_prim.m[i][j];
    Mat3<T> omega_hat = Pp * Rt;
    result->angular_velocity.x = T(0.5) * (omega_hat.m[2][1] - omega_hat.m[1][2]);
    result->angular_velocity.y = T(0.5) * (omega_hat.m[0][2] - omega_hat.m[2][0]);
    result->angular_velocity.z = T(0.5) * (omega_hat.m[1][0] - omega_hat.m[0][1]);
    return result;
  }

  // math/quaternion_math.h:16-59
  static Quat<T> logq(const Quat<T>& q) {
    T qn = ksqrt(qsqnorm(q));
    if (val(kabs(qn - T(1.0))) > 1e-5) {
      std::stringstream s; s << "logq: Only implemented for unit quaternions. Magnitude was " << val(qn);
      throw std::runtime_error(s.str());
    }
    T k; T v_squared = q.x * q.x + q.y * q.y + q.z * q.z;
    if (val(v_squared) > 1e-16) { T vn = ksqrt(v_squared); k = katan2(vn, q.w) / vn; }
    else k = T(1.0);
    return {q.x * k, q.y * k, q.z * k, T(0.0)};
  }
  // math/quaternion_math.h:62-89
  static Quat<T> expq(const Quat<T>& q) {
    T v_squared = q.x * q.x + q.y * q.y + q.z * q.z;
    T ea = kexp(q.w); T ka, kv;
    if (val(v_squared) > 1e-16) { T v_norm = ksqrt(v_squared); ka = ea * kcos(v_norm); kv = ea * ksin(v_norm) / v_norm; }
    else { ka = ea; kv = ea; }
    return {kv * q.x, kv * q.y, kv * q.z, ka};
  }

  // trajectories/uniform_so3_spline_trajectory.h:46-125
  Result<T> EvaluateSO3(T t, int flags) const {
    auto result = std::make_unique<TrajectoryEvaluation<T>>(flags);
    const T z(0.0);
    if (result->Position()) result->position = {z, z, z};
    if (result->Velocity()) result->velocity = {z, z, z};
    if (result->Acceleration()) result->acceleration = {z, z, z};
    if (!result->FlagsRotation()) return result;
    int i0; T u;
    CalculateIndexAndInterpolationAmount(t, i0, u);
    CheckRange(t, i0);
    T U[4], dU[4], B[4], dB[4];
    T u2 = kpow(u, 2.0), u3 = kpow(u, 3.0);
    T dt_inv = T(1.0) / T(meta.dt);
    U[0] = T(1.0); U[1] = u; U[2] = u2; U[3] = u3;
    basis_mul(U, kMcumul, B);
    if (result->AngularVelocity()) {
      dU[0] = dt_inv * T(0.0); dU[1] = dt_inv * T(1.0); dU[2] = dt_inv * (T(2.0) * u); dU[3] = dt_inv * (T(3.0) * u2);
      basis_mul(dU, kMcumul, dB);
    }
    Quat<T>& q = result->orientation;
    Quat<T> dq_parts[3] = {{z, z, z, T(1.0)}, {z, z, z, T(1.0)}, {z, z, z, T(1.0)}};
    q = ControlPointQuat(i0);
    const int K = i0 + 4;
    for (int i = i0 + 1; i < K; ++i) {
      Quat<T> qa = ControlPointQuat(i - 1), qb = ControlPointQuat(i);
      Quat<T> omega = logq(qmul(qconj(qa), qb));
      const T b = B[i - i0];
      Quat<T> eomegab = expq(Quat<T>{omega.x * b, omega.y * b, omega.z * b, omega.w * b});
      q = qmul(q, eomegab);
      if (result->AngularVelocity()) {
        for (int j = i0 + 1; j < K; ++j) {
          const int m = j - i0 - 1;
          if (i == j) { const T d = dB[i - i0]; dq_parts[m] = qmul(dq_parts[m], Quat<T>{omega.x * d, omega.y * d, omega.z * d, omega.w * d}); }
          dq_parts[m] = qmul(dq_parts[m], eomegab);
        }
      }
    }
    if (result->AngularVelocity()) {
      Quat<T> sum{dq_parts[0].x + dq_parts[1].x + dq_parts[2].x, dq_parts[0].y + dq_parts[1].y + dq_parts[2].y,
                  dq_parts[0].z + dq_parts[1].z + dq_parts[2].z, dq_parts[0].w + dq_parts[1].w + dq_parts[2].w};
      Quat<T> dq = qmul(ControlPointQuat(i0), sum);
      // math/quaternion_math.h:92-95
      Quat<T> w = qmul(dq, qconj(q));
      result->angular_velocity = {T(2.0) * w.x, T(2.0) * w.y, T(2.0) * w.z};
    }
    return result;
  }

  // trajectories/uniform_r3_spline_trajectory.h:34-101
  Result<T> EvaluateR3(T t, int flags) const {
    auto result = std::make_unique<TrajectoryEvaluation<T>>(flags);
    int i0; T u;
    CalculateIndexAndInterpolationAmount(t, i0, u);
    CheckRange(t, i0);
    T Up[4], Uv[4], Ua[4], Bp[4], Bv[4], Ba[4]; T u2(0.0), u3(0.0);
    const T z(0.0);
    Vec3<T>& p = result->position; Vec3<T>& v = result->velocity; Vec3<T>& a = result->acceleration;
    T dt_inv = T(1.0) / T(meta.dt);
    if (result->Position() || result->Velocity()) u2 = kpow(u, 2.0);
    if (flags & EvalPosition) u3 = kpow(u, 3.0);
    if (result->Position()) { Up[0] = T(1.0); Up[1] = u; Up[2] = u2; Up[3] = u3; basis_mul(Up, kM, Bp); p = {z, z, z}; }
    if (result->Velocity()) { Uv[0] = dt_inv * T(0.0); Uv[1] = dt_inv * T(1.0); Uv[2] = dt_inv * (T(2.0) * u); Uv[3] = dt_inv * (T(3.0) * u2); basis_mul(Uv, kM, Bv); v = {z, z, z}; }
    if (result->Acceleration()) { T s = kpow(dt_inv, 2.0); Ua[0] = s * T(0.0); Ua[1] = s * T(0.0); Ua[2] = s * T(2.0); Ua[3] = s * (T(6.0) * u); basis_mul(Ua, kM, Ba); a = {z, z, z}; }
    for (int i = i0; i < i0 + 4; ++i) {
      Vec3<T> cp = ControlPointVec(i);
      if (flags & EvalPosition) p = p + Bp[i - i0] * cp;
      if (flags & EvalVelocity) v = v + Bv[i - i0] * cp;
      if (flags & EvalAcceleration) a = a + Ba[i - i0] * cp;
    }
    if (result->Orientation()) result->orientation = {z, z, z, T(1.0)};
    if (result->AngularVelocity()) result->angular_velocity = {z, z, z};
    return result;
  }

  Result<T> Evaluate(T t, int flags) const {
    switch (kind) { case kSE3: return EvaluateSE3(t, flags); case kSO3: return EvaluateSO3(t, flags); default: return EvaluateR3(t, flags); }
  }
};

// ---- spline view over >=1 segments (trajectories/spline_base.h:168-260) -----------------------------
template <class T>
struct SplineView {
  std::vector<std::shared_ptr<SegmentView<T>>> segments;
  // spline_base.h:178-186: one heap-allocated SegmentView per segment, built on EVERY functor call
  SplineView(const SplineMeta& meta, T const* const* params, SplineKind kind, EvalOptions opt) {
    size_t offset = 0;
    for (auto& sm : meta.segments) {
      auto sv = std::make_shared<SegmentView<T>>();
      sv->meta = sm; sv->params = params + offset; sv->kind = kind; sv->opt = opt;
      segments.push_back(sv); offset += sm.n;
    }
  }
  // spline_base.h:188-202
  Result<T> Evaluate(T t, int flags) const {
    for (auto& seg : segments) if ((val(t) >= seg->meta.MinTime()) && (val(t) < seg->meta.MaxTime())) return seg->Evaluate(t, flags);
    std::stringstream ss; ss << "No segment found for time t=" << val(t);
    throw std::range_error(ss.str());
  }
};

enum TrajKind { kTrajSE3 = 0, kTrajSplit = 1, kTrajR3 = 2, kTrajSO3 = 3 };
struct TrajMeta { TrajKind kind; SplineMeta a, b; EvalOptions opt;   // a: SE3 / R3 ; b: SO3 (split: [R3 | SO3], split_trajectory.h:34-39)
  size_t NumParameters() const { return a.NumParameters() + b.NumParameters(); } };

// Generic trajectory view == entity::Map<TrajectoryModel, T>(params, meta) (cpplib/include/entity/entity.h:80-83)
template <class T>
struct TrajectoryView {
  TrajKind kind; std::unique_ptr<SplineView<T>> va, vb;
  TrajectoryView(const TrajMeta& m, T const* const* params) : kind(m.kind) {
    switch (kind) {
      case kTrajSE3: va = std::make_unique<SplineView<T>>(m.a, params, kSE3, m.opt); break;
      case kTrajR3: va = std::make_unique<SplineView<T>>(m.a, params, kR3, m.opt); break;
      case kTrajSO3: vb = std::make_unique<SplineView<T>>(m.b, params, kSO3, m.opt); break;
      case kTrajSplit:
        va = std::make_unique<SplineView<T>>(m.a, params, kR3, m.opt);
        vb = std::make_unique<SplineView<T>>(m.b, params + m.a.NumParameters(), kSO3, m.opt); break;
    }
  }
  Result<T> Evaluate(T t, int flags) const {
    if (kind == kTrajSE3 || kind == kTrajR3) return va->Evaluate(t, flags);
    if (kind == kTrajSO3) return vb->Evaluate(t, flags);
    // trajectories/split_trajectory.h:41-58
    auto result = std::make_unique<TrajectoryEvaluation<T>>(flags);
    if (result->FlagsLinear()) { auto r = va->Evaluate(t, result->FlagsLinear()); result->position = r->position; result->velocity = r->velocity; result->acceleration = r->acceleration; }
    if (result->FlagsRotation()) { auto r = vb->Evaluate(t, result->FlagsRotation()); result->orientation = r->orientation; result->angular_velocity = r->angular_velocity; }
    return result;
  }
};

// ---- sensors (sensors/sensors.h:27-86; parameter order q_ct(4) p_ct(3) time_offset(1), :139-161) --------
template <class T>
struct SensorView {
  T const* const* params;
  bool has_bias = false;   // ConstantBiasImu: + abias(3), gbias(3) (sensors/constant_bias_imu.h:28-29,106-118)
  Quat<T> relative_orientation() const { const T* p = params[0]; return {p[0], p[1], p[2], p[3]}; }
  Vec3<T> relative_position() const { const T* p = params[1]; return {p[0], p[1], p[2]}; }
  T time_offset() const { return params[2][0]; }
  Vec3<T> accelerometer_bias() const { const T* p = params[3]; return {p[0], p[1], p[2]}; }
  Vec3<T> gyroscope_bias() const { const T* p = params[4]; return {p[0], p[1], p[2]}; }
  size_t NumParameters() const { return has_bias ? 5 : 3; }
};

static const double kStandardGravity = 9.80665;  // constants.h:13,24

// sensors/imu.h:47-52 (+ constant_bias_imu.h:58-61)
template <class T> Vec3<T> imu_gyroscope(const SensorView<T>& imu, const TrajectoryView<T>& traj, T t) {
  auto result = traj.Evaluate(t + imu.time_offset(), EvalOrientation | EvalAngularVelocity);
  Vec3<T> g = qrot(qconj(result->orientation), result->angular_velocity);
  return imu.has_bias ? g + imu.gyroscope_bias() : g;
}
// sensors/imu.h:55-59 (+ constant_bias_imu.h:52-55)
template <class T> Vec3<T> imu_accelerometer(const SensorView<T>& imu, const TrajectoryView<T>& traj, T t) {
  auto result = traj.Evaluate(t + imu.time_offset(), EvalOrientation | EvalAcceleration);
  Vec3<T> grav{T(0.0), T(0.0), T(-kStandardGravity)};
  Vec3<T> a = qrot(qconj(result->orientation), result->acceleration + grav);
  return imu.has_bias ? a + imu.accelerometer_bias() : a;
}

// sensors/camera.h:24-28 + sensors/pinhole_camera.h:20-26 (+ sensors/atan_camera.h:20-52: wc, gamma)
enum CameraModel { kPinhole = 0, kAtan = 1 };
struct CameraMeta { double readout; size_t rows, cols; double K[3][3]; int model = kPinhole; double wc[2] = {0, 0}; double gamma = 0; };

template <class T> Mat3<T> camera_matrix(const CameraMeta& cm) {
  Mat3<T> K; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) K.m[i][j] = T(cm.K[i][j]);
  return K; }

// sensors/pinhole_camera.h:47-60 (EvaluateProjection; dy only when derive) via sensors/camera.h:59-63
template <class T> void pinhole_project(const CameraMeta& cm, const Vec3<T>& X, const Vec3<T>& dX, bool derive, T y[2], T dy[2]) {
  Mat3<T> K = camera_matrix<T>(cm);
  Vec3<T> p = K * X;
  y[0] = p.x / p.z; y[1] = p.y / p.z;
  if (derive) {
    const T z_eps = T(1e-32);
    Vec3<T> dp = K * dX;
    T denominator = (p.z * p.z) + z_eps;
    dy[0] = ((dp.x * p.z) - (p.x * dp.z)) / denominator;
    dy[1] = ((dp.y * p.z) - (p.y * dp.z)) / denominator;
  }
}
// sensors/pinhole_camera.h:63-67 (3x3 inverse recomputed per call, on T)
template <class T> Vec3<T> pinhole_unproject(const CameraMeta& cm, const T y[2]) {
  return inverse3(camera_matrix<T>(cm)) * Vec3<T>{y[0], y[1], T(1.0)};
}
// sensors/atan_camera.h:54-91
template <class T> void atan_project(const CameraMeta& cm, const Vec3<T>& X, const Vec3<T>& dX, bool derive, T y[2], T dy[2]) {
  const T eps = T(1e-32);
  const T gamma = T(cm.gamma), wc0 = T(cm.wc[0]), wc1 = T(cm.wc[1]);
  T A0 = X.x / (X.z + eps), A1 = X.y / (X.z + eps);
  T L0 = A0 - wc0, L1 = A1 - wc1;
  T r = ksqrt((L0 * L0 + L1 * L1) + eps);
  T f = katan(r * gamma) / gamma;
  T g0 = L0 / r, g1 = L1 / r;
  Mat3<T> K = camera_matrix<T>(cm);
  Vec3<T> Y{wc0 + f * g0, wc1 + f * g1, T(1.0)};
  Vec3<T> yk = K * Y;            // "Normalization not needed since Y(2) == 1" (:71-73)
  y[0] = yk.x; y[1] = yk.y;
  if (derive) {
    T dx = (dX.x * X.z - X.x * dX.z) / (X.z * X.z + eps);
    T dyy = (dX.y * X.z - X.y * dX.z) / (X.z * X.z + eps);
    T common = (g0 * dx + g1 * dyy);
    T df = common / (T(1.0) + kpow(gamma, 2.0) * r * r);
    T dgu = (dx * r - L0 * common) / (r * r);
    T du = f * dgu + df * g0;
    T dgv = (dyy * r - L1 * common) / (r * r);
    T dv = f * dgv + df * g1;
    Vec3<T> d = K * Vec3<T>{du, dv, T(0.0)};
    dy[0] = d.x; dy[1] = d.y;
  }
}
// sensors/atan_camera.h:92-103
template <class T> Vec3<T> atan_unproject(const CameraMeta& cm, const T y[2]) {
  const T eps = T(1e-32);
  const T gamma = T(cm.gamma), wc0 = T(cm.wc[0]), wc1 = T(cm.wc[1]);
  Vec3<T> phn = inverse3(camera_matrix<T>(cm)) * Vec3<T>{y[0], y[1], T(1.0)};
  T L0 = phn.x - wc0, L1 = phn.y - wc1;
  T r = ksqrt((L0 * L0 + L1 * L1) + eps);
  T f = ktan(r * gamma) / gamma;
  return {wc0 + f * L0 / r, wc1 + f * L1 / r, T(1.0)};
}
// CameraView::Project / EvaluateProjection / Unproject dispatch (sensors/camera.h:59-63)
template <class T> void camera_project(const CameraMeta& cm, const Vec3<T>& X, const Vec3<T>& dX, bool derive, T y[2], T dy[2]) {
  if (cm.model == kAtan) atan_project(cm, X, dX, derive, y, dy); else pinhole_project(cm, X, dX, derive, y, dy);
}
template <class T> Vec3<T> camera_unproject(const CameraMeta& cm, const T y[2]) {
  return cm.model == kAtan ? atan_unproject(cm, y) : pinhole_unproject(cm, y);
}

// measurements/gyroscope_measurement.h:36-38
template <class T> void gyro_error(double weight, const double w[3], double t, const SensorView<T>& imu, const TrajectoryView<T>& traj, T r[3]) {
  Vec3<T> m = imu_gyroscope(imu, traj, T(t));
  r[0] = T(weight) * (T(w[0]) - m.x); r[1] = T(weight) * (T(w[1]) - m.y); r[2] = T(weight) * (T(w[2]) - m.z);
}
// measurements/position_measurement.h:24-31: Error = p - trajectory.Position(t) (no sensor, no weight in the reference; the
// weight argument is the oracle's own generalisation and is 1 for the reference's measurement)
template <class T> void position_error(double weight, const double p[3], double t, const TrajectoryView<T>& traj, T r[3]) {
  auto result = traj.Evaluate(T(t), EvalPosition);
  r[0] = T(weight) * (T(p[0]) - result->position.x); r[1] = T(weight) * (T(p[1]) - result->position.y); r[2] = T(weight) * (T(p[2]) - result->position.z);
}
// measurements/orientation_measurement.h:27-31: Error = q.angularDistance(trajectory.Orientation(t)), ONE residual, with Eigen 3.3's
// QuaternionBase::angularDistance:  d = (*this) * other.conjugate();  return 2 * atan2(d.vec().norm(), abs(d.w()))
// (Eigen/src/Geometry/Quaternion.h; the CI image's libeigen3-dev is 3.3.4, .circleci/Dockerfile:1,15).  q_meas is (x, y, z, w) here.
template <class T> void orientation_error(const double q_meas[4], double t, const TrajectoryView<T>& traj, T r[1]) {
  auto result = traj.Evaluate(T(t), EvalOrientation);
  const Quat<T> q{T(q_meas[0]), T(q_meas[1]), T(q_meas[2]), T(q_meas[3])};
  const Quat<T> d = qmul(q, qconj(result->orientation));
  const T n = ksqrt(d.x * d.x + d.y * d.y + d.z * d.z);
  r[0] = T(2.0) * katan2(n, kabs(d.w));
}
// measurements/accelerometer_measurement.h:37-39
template <class T> void accel_error(double weight, const double a[3], double t, const SensorView<T>& imu, const TrajectoryView<T>& traj, T r[3]) {
  Vec3<T> m = imu_accelerometer(imu, traj, T(t));
  r[0] = T(weight) * (T(a[0]) - m.x); r[1] = T(weight) * (T(a[1]) - m.y); r[2] = T(weight) * (T(a[2]) - m.z);
}

// measurements/static_rscamera_measurement.h:21-55.  traj_ref / traj_obs are the same view in the
// reference; the oracle can hand in two views over duplicated parameters to separate the two
// contributions for the packed-Jacobian comparison.
template <class T> void reproject_static(const CameraMeta& cm, const double ref_uv[2], double ref_t0, const double obs_uv[2], double obs_t0,
                                         T inverse_depth, const TrajectoryView<T>& traj_ref, const TrajectoryView<T>& traj_obs,
                                         const SensorView<T>& camera, T y_out[2]) {
  T time_offset = camera.time_offset();
  T row_delta = T(cm.readout) / T(double(cm.rows));
  T t_ref = T(ref_t0) + time_offset + T(ref_uv[1]) * row_delta;
  T t_obs = T(obs_t0) + time_offset + T(obs_uv[1]) * row_delta;
  int flags = EvalPosition | EvalOrientation;
  auto eval_ref = traj_ref.Evaluate(t_ref, flags);
  auto eval_obs = traj_obs.Evaluate(t_obs, flags);
  const Vec3<T> p_ct = camera.relative_position();
  const Quat<T> q_ct = camera.relative_orientation();
  T y[2] = {T(ref_uv[0]), T(ref_uv[1])};
  Vec3<T> yh = camera_unproject(cm, y);
  Vec3<T> X_ref = qrot(qconj(q_ct), yh - inverse_depth * p_ct);
  Vec3<T> X = qrot(eval_ref->orientation, X_ref) + eval_ref->position * inverse_depth;
  Vec3<T> X_obs = qrot(qconj(eval_obs->orientation), X - inverse_depth * eval_obs->position);
  Vec3<T> X_camera = qrot(q_ct, X_obs) + p_ct * inverse_depth;
  const Vec3<T> zero{T(0.0), T(0.0), T(0.0)}; T unused[2];
  camera_project(cm, X_camera, zero, false, y_out, unused);   // CameraView::Project, sensors/camera.h:59-63
}

// measurements/lifting_rscamera_measurement.h:21-56: reproject_static with the observation evaluated at the LIFTED time
// t_obs = t0_obs + time_offset + vt * readout, vt in [0, 1] a parameter of the measurement (frame-normalised row time).
template <class T> void reproject_lifting(const CameraMeta& cm, const double ref_uv[2], double ref_t0, double obs_t0, T vt, T inverse_depth,
                                          const TrajectoryView<T>& trajectory, const SensorView<T>& camera, T y_out[2]) {
  T time_offset = camera.time_offset();
  T row_delta = T(cm.readout) / T(double(cm.rows));
  T t_ref = T(ref_t0) + time_offset + T(ref_uv[1]) * row_delta;
  T t_obs = T(obs_t0) + time_offset + vt * T(cm.readout);
  int flags = EvalPosition | EvalOrientation;
  auto eval_ref = trajectory.Evaluate(t_ref, flags);
  auto eval_obs = trajectory.Evaluate(t_obs, flags);
  const Vec3<T> p_ct = camera.relative_position();
  const Quat<T> q_ct = camera.relative_orientation();
  T y[2] = {T(ref_uv[0]), T(ref_uv[1])};
  Vec3<T> yh = camera_unproject(cm, y);
  Vec3<T> X_ref = qrot(qconj(q_ct), yh - inverse_depth * p_ct);
  Vec3<T> X = qrot(eval_ref->orientation, X_ref) + eval_ref->position * inverse_depth;
  Vec3<T> X_obs = qrot(qconj(eval_obs->orientation), X - inverse_depth * eval_obs->position);
  Vec3<T> X_camera = qrot(q_ct, X_obs) + p_ct * inverse_depth;
  const Vec3<T> zero{T(0.0), T(0.0), T(0.0)}; T unused[2];
  camera_project(cm, X_camera, zero, false, y_out, unused);
}

// math/quaternion_math.h:96-115
template <class T> Quat<T> embed_vector(const Vec3<T>& v) { return {v.x, v.y, v.z, T(0.0)}; }
template <class T> Quat<T> dq_from_angular_velocity(const Vec3<T>& w, const Quat<T>& q) {
  Quat<T> p = qmul(embed_vector(w), q);
  return {T(0.5) * p.x, T(0.5) * p.y, T(0.5) * p.z, T(0.5) * p.w}; }
template <class T> Vec3<T> vector_sandwich(const Quat<T>& qa, const Vec3<T>& x, const Quat<T>& qb) {
  Quat<T> r = qmul(qmul(qa, embed_vector(x)), qb);
  return {r.x, r.y, r.z}; }

// measurements/newton_rscamera_measurement.h:23-120.  As in reproject_static, traj_ref / traj_obs are the same view in
// the reference.  Comparisons on T look at the scalar part (ceres::Jet), so the iteration count follows the values.
template <class T> void reproject_newton(const CameraMeta& cm, const double ref_uv[2], double ref_t0, const double obs_uv[2], double obs_t0,
                                         T inverse_depth, const TrajectoryView<T>& traj_ref, const TrajectoryView<T>& traj_obs,
                                         const SensorView<T>& camera, T y_out[2], int max_iterations = 5) {
  T time_offset = camera.time_offset();
  T row_delta = T(cm.readout) / T(double(cm.rows));
  T t0_obs = T(obs_t0) + time_offset;
  T t_ref = T(ref_t0) + time_offset + T(ref_uv[1]) * row_delta;
  T t_obs = t0_obs + T(obs_uv[1]) * row_delta;
  int flags = EvalPosition | EvalVelocity | EvalOrientation | EvalAngularVelocity;
  const Vec3<T> p_ct = camera.relative_position();
  const Quat<T> q_ct = camera.relative_orientation();
  auto eval_ref = traj_ref.Evaluate(t_ref, flags);
  T y[2] = {T(ref_uv[0]), T(ref_uv[1])};
  Vec3<T> yh = camera_unproject(cm, y);
  Vec3<T> X_ref = qrot(qconj(q_ct), yh - inverse_depth * p_ct);
  Vec3<T> X = qrot(eval_ref->orientation, X_ref) + eval_ref->position * inverse_depth;
  const T max_time_delta = T(0.5) * T(cm.readout) / T(double(cm.rows));
  const T max_time_delta_squared = kpow(max_time_delta, 2.0);
  T min_bound = t0_obs;
  T max_bound = t0_obs + T(cm.readout);
  for (int iter = 0; iter < max_iterations; ++iter) {
    auto eval_obs = traj_obs.Evaluate(t_obs, flags);
    Vec3<T> p = eval_obs->position;
    Vec3<T> dp = eval_obs->velocity;
    Quat<T> q = eval_obs->orientation;
    Quat<T> dq = dq_from_angular_velocity(eval_obs->angular_velocity, q);
    Quat<T> dq_inv = qconj(dq);
    Quat<T> q_inv = qconj(q);
    Vec3<T> s = X - inverse_depth * p;
    Vec3<T> ds = (-inverse_depth) * dp;
    Vec3<T> X_obs = qrot(qconj(q), s);
    Vec3<T> X_obs_cam = qrot(q_ct, X_obs) + inverse_depth * p_ct;
    Vec3<T> dX_obs = vector_sandwich(dq_inv, s, q) + vector_sandwich(q_inv, ds, q) + vector_sandwich(q_inv, s, dq);
    Vec3<T> dX_obs_cam = qrot(q_ct, dX_obs) + inverse_depth * p_ct;   // sic (:92)
    T dy[2];
    camera_project(cm, X_obs_cam, dX_obs_cam, true, y_out, dy);
    T v = y_out[1];
    T dv = dy[1];
    T f = v - (T(double(cm.rows)) * (t_obs - t0_obs) / T(cm.readout));
    T df = dv - (T(double(cm.rows)) / T(cm.readout));
    T dt = f / df;
    t_obs -= dt;
    if ((dt * dt) < max_time_delta_squared) break;
    if (t_obs < min_bound) t_obs = min_bound;
    else if (t_obs > max_bound) t_obs = max_bound;
  }
}

// ---- structure rule (trajectories/spline_base.h:361-404) -------------------------------------------------
// Returns the knot indices added (in order) and fills meta.segments, for the ordered spans `times`.
inline void spline_add_to_problem(double master_dt, double master_t0, const std::vector<std::pair<double, double>>& times,
                                  SplineMeta& meta, std::vector<int>& knot_ids) {
  int current_segment_start = 0, current_segment_end = -1;
  for (auto tt : times) {
    int i1, i2;
    i1 = static_cast<int>(std::floor((tt.first - master_t0) / master_dt));   // spline_base.h:148-152
    i2 = static_cast<int>(std::floor((tt.second - master_t0) / master_dt));
    if (i1 > current_segment_end) {
      double segment_t0 = master_t0 + master_dt * i1;
      SplineSegmentMeta sm; sm.dt = master_dt; sm.t0 = segment_t0; sm.n = 0;
      meta.segments.push_back(sm);
      current_segment_start = i1;
    } else {
      i1 = current_segment_end + 1;
    }
    auto& cur = meta.segments.back();
    for (int i = i1; i < (i2 + 4); ++i) { knot_ids.push_back(i); cur.n += 1; }
    current_segment_end = current_segment_start + int(cur.n) - 1;
  }
}

// trajectory_estimator.h:97-122
inline void check_time_spans(const std::vector<std::pair<double, double>>& times, double min_time, double max_time) {
  int i = 0; double t1_prev = 0;
  for (auto& ts : times) {
    double t1 = ts.first, t2 = ts.second;
    if ((t1 < min_time) || (t2 >= max_time)) throw std::range_error("Time span out of range for trajectory");
    if (t1 > t2) throw std::range_error("At least one time span begins before it ends");
    else if ((i > 0) && (t1 < t1_prev)) throw std::range_error("Time spans are not ordered");
    t1_prev = t1; i += 1;
  }
}

// ---- ceres::DynamicAutoDiffCostFunction<Functor, 4>::Evaluate (un-vendored) -------
// functor(T const* const* params, T* residuals); jacobians[k] == nullptr for constant blocks; row-major blocks.
template <class Functor>
bool autodiff_evaluate(const Functor& f, const std::vector<int>& block_sizes, double const* const* parameters,
                       int num_residuals, double* residuals, double** jacobians) {
  if (jacobians == nullptr) return f.template operator()<double>(parameters, residuals);
  constexpr int Stride = 4;
  using Jet = Dual<Stride>;
  const int num_blocks = int(block_sizes.size());
  int num_parameters = 0; for (int s : block_sizes) num_parameters += s;
  std::vector<Jet> input_jets(num_parameters);
  std::vector<Jet> output_jets(num_residuals);
  std::vector<Jet*> jet_parameters(num_blocks, nullptr);
  int num_active_parameters = 0;
  std::vector<int> start_derivative_section;
  int cursor = 0;
  for (int i = 0; i < num_blocks; ++i) {
    jet_parameters[i] = &input_jets[cursor];
    if (jacobians[i] != nullptr) { start_derivative_section.push_back(cursor); num_active_parameters += block_sizes[i]; }
    else start_derivative_section.push_back(-1);
    for (int j = 0; j < block_sizes[i]; ++j, ++cursor) input_jets[cursor].a = parameters[i][j];
  }
  // one "global" derivative index per active scalar parameter
  std::vector<int> active_index(num_parameters, -1);
  { int k = 0; cursor = 0; for (int i = 0; i < num_blocks; ++i) for (int j = 0; j < block_sizes[i]; ++j, ++cursor) if (jacobians[i]) active_index[cursor] = k++; }
  const int num_strides = (num_active_parameters + Stride - 1) / Stride;
  bool first = true;
  for (int pass = 0; pass < std::max(num_strides, 1); ++pass) {
    const int lo = pass * Stride, hi = lo + Stride;
    for (int p = 0; p < num_parameters; ++p) {
      for (int s = 0; s < Stride; ++s) input_jets[p].v[s] = 0.0;
      const int k = active_index[p];
      if (k >= lo && k < hi) input_jets[p].v[k - lo] = 1.0;
    }
    if (!f.template operator()<Jet>(jet_parameters.data(), output_jets.data())) return false;
    if (first) { for (int r = 0; r < num_residuals; ++r) residuals[r] = output_jets[r].a; first = false; }
    cursor = 0;
    for (int i = 0; i < num_blocks; ++i) for (int j = 0; j < block_sizes[i]; ++j, ++cursor) {
      const int k = active_index[cursor];
      if (k >= lo && k < hi) for (int r = 0; r < num_residuals; ++r) jacobians[i][r * block_sizes[i] + j] = output_jets[r].v[k - lo];
    }
  }
  return true;
}

}  // namespace kto

// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
//
// Minimal fixed-size vector / matrix / quaternion / SO3 / SE3 helpers restating the
// semantics of the two un-vendored dependencies that carry the reference's hot-path
// arithmetic:
//   * Eigen 3.3.x   (Quaternion product / conjugate / q*v / toRotationMatrix, 3x3 inverse)
//   * Sophus @ 00f3fd91c153ef04 (.circleci/Dockerfile:19-21, docs/users/install.rst:31)
//     SO3/SE3 exp, log, inverse, operator*, hat, matrix.
// Neither source tree is present under /root/reference, so what is written here is the
// published algorithm of those libraries as recalled; every place where the pinned
// version might differ is marked ASSUMPTION.  Call sites in the reference:
// cpplib/include/kontiki/trajectories/uniform_se3_spline_trajectory.h:87-93,154-168,178-183.
#pragma once
#include "dual.hpp"

namespace kto {

template <class T> struct Vec3 { T x, y, z; };
template <class T> struct Vec6 { T d[6]; };           // Sophus SE3 tangent order: [upsilon(3); omega(3)]
template <class T> struct Mat3 { T m[3][3]; };
template <class T> struct Mat4 { T m[4][4]; };
template <class T> struct Quat { T x, y, z, w; };     // Eigen coefficient order (x, y, z, w)

template <class T> inline Vec3<T> operator+(const Vec3<T>& a, const Vec3<T>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class T> inline Vec3<T> operator-(const Vec3<T>& a, const Vec3<T>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class T> inline Vec3<T> operator*(const T& s, const Vec3<T>& a) { return {s * a.x, s * a.y, s * a.z}; }
template <class T> inline Vec3<T> operator*(const Vec3<T>& a, const T& s) { return {a.x * s, a.y * s, a.z * s}; }
template <class T> inline T dot(const Vec3<T>& a, const Vec3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> inline Vec3<T> cross(const Vec3<T>& a, const Vec3<T>& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

template <class T> inline Mat3<T> mat3_identity() { Mat3<T> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = T(i == j ? 1.0 : 0.0); return r; }
template <class T> inline Mat3<T> operator*(const Mat3<T>& a, const Mat3<T>& b) {
  Mat3<T> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { T s = a.m[i][0] * b.m[0][j]; s += a.m[i][1] * b.m[1][j]; s += a.m[i][2] * b.m[2][j]; r.m[i][j] = s; } return r; }
template <class T> inline Vec3<T> operator*(const Mat3<T>& a, const Vec3<T>& v) {
  return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
          a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
          a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z}; }
template <class T> inline Mat3<T> operator+(const Mat3<T>& a, const Mat3<T>& b) { Mat3<T> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] + b.m[i][j]; return r; }
template <class T> inline Mat3<T> operator*(const T& s, const Mat3<T>& a) { Mat3<T> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = s * a.m[i][j]; return r; }
template <class T> inline Mat3<T> transpose(const Mat3<T>& a) { Mat3<T> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i]; return r; }
template <class T> inline Mat3<T> hat3(const Vec3<T>& w) {
  Mat3<T> r; const T z(0.0);
  r.m[0][0] = z;    r.m[0][1] = -w.z; r.m[0][2] = w.y;
  r.m[1][0] = w.z;  r.m[1][1] = z;    r.m[1][2] = -w.x;
  r.m[2][0] = -w.y; r.m[2][1] = w.x;  r.m[2][2] = z;
  return r; }
// Eigen computes fixed 3x3 inverses by cofactors / determinant (Eigen/src/LU/InverseImpl.h).
template <class T> inline Mat3<T> inverse3(const Mat3<T>& a) {
  Mat3<T> c;
  c.m[0][0] = a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1];
  c.m[0][1] = a.m[0][2] * a.m[2][1] - a.m[0][1] * a.m[2][2];
  c.m[0][2] = a.m[0][1] * a.m[1][2] - a.m[0][2] * a.m[1][1];
  c.m[1][0] = a.m[1][2] * a.m[2][0] - a.m[1][0] * a.m[2][2];
  c.m[1][1] = a.m[0][0] * a.m[2][2] - a.m[0][2] * a.m[2][0];
  c.m[1][2] = a.m[0][2] * a.m[1][0] - a.m[0][0] * a.m[1][2];
  c.m[2][0] = a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0];
  c.m[2][1] = a.m[0][1] * a.m[2][0] - a.m[0][0] * a.m[2][1];
  c.m[2][2] = a.m[0][0] * a.m[1][1] - a.m[0][1] * a.m[1][0];
  const T det = a.m[0][0] * c.m[0][0] + a.m[0][1] * c.m[1][0] + a.m[0][2] * c.m[2][0];
  const T inv = T(1.0) / det;
  return inv * c; }

template <class T> inline Mat4<T> mat4_zero() { Mat4<T> r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = T(0.0); return r; }
template <class T> inline Mat4<T> operator*(const Mat4<T>& a, const Mat4<T>& b) {
  Mat4<T> r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
    T s = a.m[i][0] * b.m[0][j]; s += a.m[i][1] * b.m[1][j]; s += a.m[i][2] * b.m[2][j]; s += a.m[i][3] * b.m[3][j]; r.m[i][j] = s; } return r; }
template <class T> inline Mat4<T> operator+(const Mat4<T>& a, const Mat4<T>& b) { Mat4<T> r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = a.m[i][j] + b.m[i][j]; return r; }
template <class T> inline Mat4<T> operator*(const Mat4<T>& a, const T& s) { Mat4<T> r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = a.m[i][j] * s; return r; }
template <class T> inline Mat4<T> operator*(const T& s, const Mat4<T>& a) { return a * s; }

// ---- Eigen::Quaternion semantics -------------------------------------------------------------
template <class T> inline Quat<T> qmul(const Quat<T>& a, const Quat<T>& b) {  // Hamilton product, no renormalisation
  return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x,
          a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z}; }
template <class T> inline Quat<T> qconj(const Quat<T>& q) { return {-q.x, -q.y, -q.z, q.w}; }
template <class T> inline T qsqnorm(const Quat<T>& q) { return q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w; }
// Eigen QuaternionBase::_transformVector: v + w*(2 u x v) + u x (2 u x v); a polynomial in q,
// i.e. NOT scale invariant when q is not exactly unit.
template <class T> inline Vec3<T> qrot(const Quat<T>& q, const Vec3<T>& v) {
  const Vec3<T> u{q.x, q.y, q.z};
  Vec3<T> uv = cross(u, v); uv = uv + uv;
  return v + q.w * uv + cross(u, uv); }
// Eigen QuaternionBase::toRotationMatrix (the same polynomial as qrot).
template <class T> inline Mat3<T> qmat(const Quat<T>& q) {
  const T tx = T(2.0) * q.x, ty = T(2.0) * q.y, tz = T(2.0) * q.z;
  const T twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const T txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const T tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  Mat3<T> r;
  r.m[0][0] = T(1.0) - (tyy + tzz); r.m[0][1] = txy - twz;            r.m[0][2] = txz + twy;
  r.m[1][0] = txy + twz;            r.m[1][1] = T(1.0) - (txx + tzz); r.m[1][2] = tyz - twx;
  r.m[2][0] = txz - twy;            r.m[2][1] = tyz + twx;            r.m[2][2] = T(1.0) - (txx + tyy);
  return r; }

// ---- Sophus::SO3 / SE3 semantics --------------------------------------------------------------
static const double kSophusEps = 1e-10;  // Sophus::Constants<double>::epsilon()

// ASSUMPTION: every SO3 that Sophus *constructs* from a quaternion (results of inverse(),
// operator*, exp()) is re-normalised (SO3(QuaternionBase) ctor -> normalize(); older revisions use
// the first-order fix q *= 2/(1+|q|^2) inside operator*=, which has the same value to 1e-32 and the
// same first derivative at |q|=1).  A knot that is only *mapped* (Eigen::Map<SE3>) is used as is.
template <class T> inline Quat<T> so3_normalized(const Quat<T>& q) {
  const T n = ksqrt(qsqnorm(q)); const T inv = T(1.0) / n;
  return {q.x * inv, q.y * inv, q.z * inv, q.w * inv}; }
template <class T> inline Quat<T> so3_mul(const Quat<T>& a, const Quat<T>& b) { return so3_normalized(qmul(a, b)); }
template <class T> inline Quat<T> so3_inverse(const Quat<T>& a) { return so3_normalized(qconj(a)); }

// Sophus SO3::exp (expAndTheta).
template <class T> inline Quat<T> so3_exp(const Vec3<T>& omega, T* theta_out = nullptr) {
  const T theta_sq = dot(omega, omega);
  const T theta = ksqrt(theta_sq);
  const T half_theta = T(0.5) * theta;
  T imag_factor, real_factor;
  if (val(theta) < kSophusEps) {
    const T theta_po4 = theta_sq * theta_sq;
    imag_factor = T(0.5) - T(1.0 / 48.0) * theta_sq + T(1.0 / 3840.0) * theta_po4;
    real_factor = T(1.0) - T(1.0 / 8.0) * theta_sq + T(1.0 / 384.0) * theta_po4;
  } else {
    const T sin_half_theta = ksin(half_theta);
    imag_factor = sin_half_theta / theta;
    real_factor = kcos(half_theta);
  }
  if (theta_out) *theta_out = theta;
  return so3_normalized(Quat<T>{imag_factor * omega.x, imag_factor * omega.y, imag_factor * omega.z, real_factor}); }

// Sophus SO3::logAndTheta.  ASSUMPTION: uses atan(n/w) (not atan2) as in the 2018 revisions.
template <class T> inline Vec3<T> so3_log(const Quat<T>& q, T* theta_out = nullptr) {
  const T squared_n = q.x * q.x + q.y * q.y + q.z * q.z;
  const T n = ksqrt(squared_n);
  const T w = q.w;
  T two_atan_nbyw_by_n;
  if (val(n) < kSophusEps) {
    const T squared_w = w * w;
    two_atan_nbyw_by_n = T(2.0) / w - T(2.0) * squared_n / (w * squared_w);
  } else if (kabs(val(w)) < kSophusEps) {
    const double kPi = 3.14159265358979323846;
    two_atan_nbyw_by_n = (val(w) > 0.0) ? T(kPi) / n : T(-kPi) / n;
  } else {
    two_atan_nbyw_by_n = T(2.0) * katan(n / w) / n;
  }
  if (theta_out) *theta_out = two_atan_nbyw_by_n * n;
  return {two_atan_nbyw_by_n * q.x, two_atan_nbyw_by_n * q.y, two_atan_nbyw_by_n * q.z}; }

template <class T> struct SE3 { Quat<T> q; Vec3<T> t; };

// Sophus SE3::inverse: (R^-1, R^-1 * (-t)).
template <class T> inline SE3<T> se3_inverse(const SE3<T>& a) {
  const Quat<T> qi = so3_inverse(a.q);
  const Vec3<T> mt{-a.t.x, -a.t.y, -a.t.z};
  return {qi, qrot(qi, mt)}; }
// Sophus SE3::operator*=: translation += so3 * other.translation; so3 *= other.so3.
template <class T> inline SE3<T> se3_mul(const SE3<T>& a, const SE3<T>& b) {
  return {so3_mul(a.q, b.q), a.t + qrot(a.q, b.t)}; }
// Sophus SE3::exp([upsilon; omega]).
template <class T> inline SE3<T> se3_exp(const Vec6<T>& xi) {
  const Vec3<T> ups{xi.d[0], xi.d[1], xi.d[2]}, omega{xi.d[3], xi.d[4], xi.d[5]};
  T theta;
  const Quat<T> q = so3_exp(omega, &theta);
  const Mat3<T> Omega = hat3(omega);
  const Mat3<T> Omega_sq = Omega * Omega;
  Mat3<T> V;
  if (val(theta) < kSophusEps) {
    V = qmat(q);   // "Note: That is an accurate expansion!" (Sophus se3.hpp)
  } else {
    const T theta_sq = theta * theta;
    V = mat3_identity<T>() + ((T(1.0) - kcos(theta)) / theta_sq) * Omega +
        ((theta - ksin(theta)) / (theta_sq * theta)) * Omega_sq;
  }
  return {q, V * ups}; }
// Sophus SE3::log.
template <class T> inline Vec6<T> se3_log(const SE3<T>& a) {
  T theta;
  const Vec3<T> omega = so3_log(a.q, &theta);
  const Mat3<T> Omega = hat3(omega);
  Mat3<T> V_inv;
  if (kabs(val(theta)) < kSophusEps) {
    V_inv = mat3_identity<T>() + T(-0.5) * Omega + T(1.0 / 12.0) * (Omega * Omega);
  } else {
    const T half_theta = T(0.5) * theta;
    V_inv = mat3_identity<T>() + T(-0.5) * Omega +
            ((T(1.0) - theta * kcos(half_theta) / (T(2.0) * ksin(half_theta))) / (theta * theta)) * (Omega * Omega);
  }
  const Vec3<T> ups = V_inv * a.t;
  Vec6<T> r; r.d[0] = ups.x; r.d[1] = ups.y; r.d[2] = ups.z; r.d[3] = omega.x; r.d[4] = omega.y; r.d[5] = omega.z;
  return r; }
// Sophus SE3::hat: [[hat(omega), upsilon], [0, 0]].
template <class T> inline Mat4<T> se3_hat(const Vec6<T>& xi) {
  Mat4<T> r = mat4_zero<T>();
  const Mat3<T> O = hat3(Vec3<T>{xi.d[3], xi.d[4], xi.d[5]});
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) r.m[i][j] = O.m[i][j]; r.m[i][3] = xi.d[i]; }
  return r; }
// Sophus SE3::matrix: [[R, t], [0, 1]] with R = unit_quaternion().toRotationMatrix().
template <class T> inline Mat4<T> se3_matrix(const SE3<T>& a) {
  Mat4<T> r = mat4_zero<T>();
  const Mat3<T> R = qmat(a.q);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = R.m[i][j];
  r.m[0][3] = a.t.x; r.m[1][3] = a.t.y; r.m[2][3] = a.t.z; r.m[3][3] = T(1.0);
  return r; }

}  // namespace kto

// kontiki_b200 -- small fixed-size fp64 linear algebra and SO(3)/SE(3) closed forms used by the
// fused residual+Jacobian kernels.  Everything is __host__ __device__ so that the very same code
// can be compiled for the host by tests/ (host_check.cpp) and compared with the CPU oracle without
// a GPU; the shipped library only ever runs it on the device.
//
// Conventions (same as the reference):
//   quaternion storage (x, y, z, w); SE3 knot = [qx qy qz qw tx ty tz];
//   SE3 tangent xi = [upsilon(3); omega(3)] (Sophus order); x_world = R x_body + p.
#pragma once
#include <math.h>

#ifndef KB_HD
#if defined(__CUDACC__)
#define KB_HD __host__ __device__ __forceinline__
#else
#define KB_HD inline
#endif
#endif

namespace kb {

struct V3 { double x, y, z; };
template <int N> struct Mr { double a[3 * N]; };   // N x 3, row-major
using M3 = Mr<3>;

// A sequence point for the compiler: independent chains of fp64 work on either side are not interleaved (interleaving two reverse sweeps for
// instruction-level parallelism doubled the live state and cost kilobytes of spill per thread in the 255-register kernels).
#if defined(__CUDA_ARCH__) && !defined(KTK_NO_SEQ)      // -DKTK_NO_SEQ: A/B builds (tools/build_variant.sh)
#define KB_SEQ() asm volatile("" ::: "memory")
#else
#define KB_SEQ() ((void)0)
#endif
KB_HD V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
KB_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
KB_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
KB_HD V3 operator*(double s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
KB_HD V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }
KB_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
KB_HD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

KB_HD M3 m3_identity() { M3 r; r.a[0] = 1; r.a[1] = 0; r.a[2] = 0; r.a[3] = 0; r.a[4] = 1; r.a[5] = 0; r.a[6] = 0; r.a[7] = 0; r.a[8] = 1; return r; }
KB_HD M3 m3_zero() { M3 r; for (int i = 0; i < 9; ++i) r.a[i] = 0.0; return r; }
KB_HD M3 operator*(const M3& A, const M3& B) {
  M3 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.a[3 * i + j] = A.a[3 * i] * B.a[j] + A.a[3 * i + 1] * B.a[3 + j] + A.a[3 * i + 2] * B.a[6 + j];
  return r; }
// A * B^T
KB_HD M3 mul_nt(const M3& A, const M3& B) {
  M3 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.a[3 * i + j] = A.a[3 * i] * B.a[3 * j] + A.a[3 * i + 1] * B.a[3 * j + 1] + A.a[3 * i + 2] * B.a[3 * j + 2];
  return r; }
// A^T * B
KB_HD M3 mul_tn(const M3& A, const M3& B) {
  M3 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.a[3 * i + j] = A.a[i] * B.a[j] + A.a[3 + i] * B.a[3 + j] + A.a[6 + i] * B.a[6 + j];
  return r; }
KB_HD V3 operator*(const M3& A, V3 v) {
  return v3(A.a[0] * v.x + A.a[1] * v.y + A.a[2] * v.z, A.a[3] * v.x + A.a[4] * v.y + A.a[5] * v.z, A.a[6] * v.x + A.a[7] * v.y + A.a[8] * v.z); }
// A^T * v
KB_HD V3 mul_t(const M3& A, V3 v) {
  return v3(A.a[0] * v.x + A.a[3] * v.y + A.a[6] * v.z, A.a[1] * v.x + A.a[4] * v.y + A.a[7] * v.z, A.a[2] * v.x + A.a[5] * v.y + A.a[8] * v.z); }
KB_HD M3 operator+(const M3& A, const M3& B) { M3 r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.a[i] = A.a[i] + B.a[i]; return r; }
KB_HD M3 operator-(const M3& A, const M3& B) { M3 r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.a[i] = A.a[i] - B.a[i]; return r; }
KB_HD M3 operator*(double s, const M3& A) { M3 r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.a[i] = s * A.a[i]; return r; }
KB_HD M3 transpose(const M3& A) { M3 r; r.a[0] = A.a[0]; r.a[1] = A.a[3]; r.a[2] = A.a[6]; r.a[3] = A.a[1]; r.a[4] = A.a[4]; r.a[5] = A.a[7]; r.a[6] = A.a[2]; r.a[7] = A.a[5]; r.a[8] = A.a[8]; return r; }
KB_HD M3 hat(V3 w) { M3 r; r.a[0] = 0; r.a[1] = -w.z; r.a[2] = w.y; r.a[3] = w.z; r.a[4] = 0; r.a[5] = -w.x; r.a[6] = -w.y; r.a[7] = w.x; r.a[8] = 0; return r; }
KB_HD M3 outer(V3 a, V3 b) { M3 r; r.a[0] = a.x * b.x; r.a[1] = a.x * b.y; r.a[2] = a.x * b.z; r.a[3] = a.y * b.x; r.a[4] = a.y * b.y; r.a[5] = a.y * b.z; r.a[6] = a.z * b.x; r.a[7] = a.z * b.y; r.a[8] = a.z * b.z; return r; }
KB_HD double trace(const M3& A) { return A.a[0] + A.a[4] + A.a[8]; }
// A * hat(v): column-wise cross products.  (A hat(v))_{i,:} = A_{i,:} x v ... as a row: a hat(v) = (a x v)^T
KB_HD M3 mul_hat(const M3& A, V3 v) {
  M3 r;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double ax = A.a[3 * i], ay = A.a[3 * i + 1], az = A.a[3 * i + 2];
    r.a[3 * i] = ay * v.z - az * v.y; r.a[3 * i + 1] = az * v.x - ax * v.z; r.a[3 * i + 2] = ax * v.y - ay * v.x; }
  return r; }
// hat(v) * A
KB_HD M3 hat_mul(V3 v, const M3& A) {
  M3 r;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double ax = A.a[j], ay = A.a[3 + j], az = A.a[6 + j];
    r.a[j] = v.y * az - v.z * ay; r.a[3 + j] = v.z * ax - v.x * az; r.a[6 + j] = v.x * ay - v.y * ax; }
  return r; }

// ---- N x 3 row blocks (adjoint rows of an N-row residual) -----------------------------------------------------------
template <int N> KB_HD Mr<N> rmul(const Mr<N>& A, const M3& B) {        // A * B
  Mr<N> r;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.a[3 * i + j] = A.a[3 * i] * B.a[j] + A.a[3 * i + 1] * B.a[3 + j] + A.a[3 * i + 2] * B.a[6 + j];
  return r; }
template <int N> KB_HD Mr<N> rmul_nt(const Mr<N>& A, const M3& B) {     // A * B^T
  Mr<N> r;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.a[3 * i + j] = A.a[3 * i] * B.a[3 * j] + A.a[3 * i + 1] * B.a[3 * j + 1] + A.a[3 * i + 2] * B.a[3 * j + 2];
  return r; }
template <int N> KB_HD Mr<N> rmul_hat(const Mr<N>& A, V3 v) {            // A * hat(v): every row crossed with v
  Mr<N> r;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double ax = A.a[3 * i], ay = A.a[3 * i + 1], az = A.a[3 * i + 2];
    r.a[3 * i] = ay * v.z - az * v.y; r.a[3 * i + 1] = az * v.x - ax * v.z; r.a[3 * i + 2] = ax * v.y - ay * v.x; }
  return r; }
template <int N> KB_HD Mr<N> radd(const Mr<N>& A, const Mr<N>& B) { Mr<N> r;
#pragma unroll
  for (int i = 0; i < 3 * N; ++i) r.a[i] = A.a[i] + B.a[i]; return r; }
template <int N> KB_HD Mr<N> rsub(const Mr<N>& A, const Mr<N>& B) { Mr<N> r;
#pragma unroll
  for (int i = 0; i < 3 * N; ++i) r.a[i] = A.a[i] - B.a[i]; return r; }
template <int N> KB_HD Mr<N> rscale(double s, const Mr<N>& A) { Mr<N> r;
#pragma unroll
  for (int i = 0; i < 3 * N; ++i) r.a[i] = s * A.a[i]; return r; }
template <int N> KB_HD V3 rrow(const Mr<N>& A, int i) { return v3(A.a[3 * i], A.a[3 * i + 1], A.a[3 * i + 2]); }
template <int N> KB_HD Mr<N> rzero() { Mr<N> r;
#pragma unroll
  for (int i = 0; i < 3 * N; ++i) r.a[i] = 0.0; return r; }

// Rotation matrix of a unit quaternion (x,y,z,w) -- same polynomial as Eigen's toRotationMatrix.
KB_HD M3 quat_to_rot(double x, double y, double z, double w) {
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  M3 r;
  r.a[0] = 1.0 - (tyy + tzz); r.a[1] = txy - twz; r.a[2] = txz + twy;
  r.a[3] = txy + twz; r.a[4] = 1.0 - (txx + tzz); r.a[5] = tyz - twx;
  r.a[6] = txz - twy; r.a[7] = tyz + twx; r.a[8] = 1.0 - (txx + tyy);
  return r; }

// ---- coefficient functions of the squared angle x = phi^2 --------------------------------------------
// sa = sin(phi)/phi, cb = (1-cos phi)/phi^2, cc = (phi - sin phi)/phi^3,
// c2 = (phi^2 + 2 cos phi - 2)/(2 phi^4), c3 = (2 phi - 3 sin phi + phi cos phi)/(2 phi^5)
// Each is an entire function of x; the closed forms cancel catastrophically for small phi, so a Taylor series is used
// below a per-group threshold chosen so that BOTH branches are good to <1e-13 relative (series truncation below it,
// cancellation above it): x < 1e-2 for sa/cb/cc (6 terms), x < 0.5 for c2/c3 (7 terms).
// Taylor coefficients of the five functions below (x^0 .. x^5 of sa, cb, cc; x^0 .. x^6 of c2, c3).  On the device they live in the
// constant bank so that the Horner DFMAs read them as c[bank][offset] operands: as literals every coefficient costs two UMOV / IMAD.MOV
// issue slots (64-bit immediates do not exist), ~300 per camera row in a kernel that is bound by issue slots and dependent latency.
#define KB_TAYLOR_COEFS                                                                                                        \
  1.0, -1.0 / 6.0, 1.0 / 120.0, -1.0 / 5040.0, 1.0 / 362880.0, -1.0 / 39916800.0,                                               \
  0.5, -1.0 / 24.0, 1.0 / 720.0, -1.0 / 40320.0, 1.0 / 3628800.0, -1.0 / 479001600.0,                                           \
  1.0 / 6.0, -1.0 / 120.0, 1.0 / 5040.0, -1.0 / 362880.0, 1.0 / 39916800.0, -1.0 / 6227020800.0,                                \
  1.0 / 24.0, -1.0 / 720.0, 1.0 / 40320.0, -1.0 / 3628800.0, 1.0 / 479001600.0, -1.0 / 87178291200.0, 1.0 / 20922789888000.0,   \
  1.0 / 120.0, -1.0 / 2520.0, 1.0 / 120960.0, -1.0 / 9979200.0, 1.0 / 1245404160.0, -1.0 / 217945728000.0, 1.0 / 50812489728000.0
#if defined(__CUDACC__)
__constant__ double kTaylorDev[32] = {KB_TAYLOR_COEFS};
#endif
static const double kTaylorHost[32] = {KB_TAYLOR_COEFS};
#if defined(__CUDA_ARCH__)
#define KB_TC(i) kTaylorDev[i]
#else
#define KB_TC(i) kTaylorHost[i]
#endif
struct AngleCoefs { double sa, cb, cc, c2, c3; };
#define KB_SMALL_X 1.0e-2
#define KB_SMALL_XQ 0.5
// Large-angle branch (closed forms), out of line on the device: spline increments B_j omega_j are small, so the Taylor branch is the
// hot one and is kept straight-line; inlining sincos and three divisions into every exp_part costs branch-merge moves and code size.
// The cold function returns ONE coefficient in registers (which: 0 sa, 1 cb, 2 cc, 3 c2, 4 c3): a struct passed by reference lives on
// the stack, and ptxas stored it there BEFORE the never-taken branch -- 60 STL per 32-row tile, 120 MB of local-memory write traffic per
// H1 evaluation on the SM -> L2 path (ncu l1tex__t_sectors_pipe_lsu_mem_local_op_st, profiles/README.md r2a).
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
double angle_coef_large(double x, int which) {
  const double phi = sqrt(x);
  double s, co;
#if defined(__CUDA_ARCH__)
  sincos(phi, &s, &co);
#else
  s = sin(phi); co = cos(phi);
#endif
  const double ix = 1.0 / x;
  if (which == 0) return s / phi;
  if (which == 1) return (1.0 - co) * ix;
  if (which == 2) return (phi - s) * ix / phi;
  if (which == 3) return (x + 2.0 * co - 2.0) * 0.5 * ix * ix;
  return (2.0 * phi - 3.0 * s + phi * co) * 0.5 * ix * ix / phi;
}
KB_HD AngleCoefs angle_coefs(double x, bool need_q) {
  AngleCoefs c;
  c.sa = KB_TC(0) + x * (KB_TC(1) + x * (KB_TC(2) + x * (KB_TC(3) + x * (KB_TC(4) + x * KB_TC(5)))));
  c.cb = KB_TC(6) + x * (KB_TC(7) + x * (KB_TC(8) + x * (KB_TC(9) + x * (KB_TC(10) + x * KB_TC(11)))));
  c.cc = KB_TC(12) + x * (KB_TC(13) + x * (KB_TC(14) + x * (KB_TC(15) + x * (KB_TC(16) + x * KB_TC(17)))));
  c.c2 = 0.0; c.c3 = 0.0;
  if (need_q) {
    c.c2 = KB_TC(18) + x * (KB_TC(19) + x * (KB_TC(20) + x * (KB_TC(21) + x * (KB_TC(22) + x * (KB_TC(23) + x * KB_TC(24))))));
    c.c3 = KB_TC(25) + x * (KB_TC(26) + x * (KB_TC(27) + x * (KB_TC(28) + x * (KB_TC(29) + x * (KB_TC(30) + x * KB_TC(31))))));
  }
  if (!(x < KB_SMALL_X)) {
    c.sa = angle_coef_large(x, 0); c.cb = angle_coef_large(x, 1); c.cc = angle_coef_large(x, 2);
    if (need_q && !(x < KB_SMALL_XQ)) { c.c2 = angle_coef_large(x, 3); c.c3 = angle_coef_large(x, 4); }
  }
  return c; }

// Q block of the SE(3) left Jacobian J_l([rho; phi]) = [[Jl(phi), Ql],[0, Jl(phi)]] (Barfoot 2017, eq. 7.86),
// reduced with a^ b^ = b a^T - (a.b) I and a^ b^ a^ = -(a.b) a^ :
//   Ql = 1/2 rho^ + c1 (rho phi^T + phi rho^T - 2 s I - s phi^) + c2 (n phi^T - phi n^T + s phi^) - 2 c3 s phi^ phi^
// with s = phi.rho, n = phi x rho.  sign = +1 gives Ql, sign = -1 gives Qr = Ql(-rho, -phi).
KB_HD M3 se3_q_block(V3 rho, V3 phi, const AngleCoefs& c, double x, double sign) {
  const double s = dot(phi, rho);
  const V3 n = cross(phi, rho);
  M3 sym = outer(rho, phi); { M3 t = outer(phi, rho); sym = sym + t; }
  const M3 asym = outer(n, phi) - outer(phi, n);
  const M3 ph = hat(phi);
  // phi^ phi^ = phi phi^T - x I
  M3 pp = outer(phi, phi); pp.a[0] -= x; pp.a[4] -= x; pp.a[8] -= x;
  M3 r = (sign * 0.5) * hat(rho);
  r = r + c.cc * sym;
  r.a[0] -= 2.0 * c.cc * s; r.a[4] -= 2.0 * c.cc * s; r.a[8] -= 2.0 * c.cc * s;
  r = r + (sign * s * (c.c2 - c.cc)) * ph;
  r = r + (sign * c.c2) * asym;
  r = r + (-2.0 * c.c3 * s) * pp;
  return r; }

}  // namespace kb

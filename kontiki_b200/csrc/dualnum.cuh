// kontiki_b200 -- one-direction forward-mode dual number used by the knot-pair prepass (K0) and by the Newton
// rolling-shutter rows (newton_math.cuh), whose derivative IS the forward-mode derivative of an iteration.
//
// The knot-pair log map  omega = log(P_{i-1}^-1 * P_i)  is a function of two *ambient* SE3 knots
// (7 doubles each, quaternion not constrained to unit norm while differentiating).  The reference
// differentiates it with ceres::Jet through Sophus (uniform_se3_spline_trajectory.h:159-162), so the
// derivative along the quaternion-norm direction is whatever the arithmetic of inverse()/operator*/log()
// gives.  K0 runs that arithmetic once per knot pair and direction (n_knots-1 pairs x 14 directions --
// a few 10^4 threads per evaluation point), so a plain dual number is the cheapest exact way to get the
// 6x14 pair Jacobian; all per-measurement work downstream is analytic.
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

struct D1 {
  double a, d;
  KB_HD D1() : a(0.0), d(0.0) {}
  KB_HD D1(double x) : a(x), d(0.0) {}   // NOLINT(implicit)
  KB_HD D1(double x, double dx) : a(x), d(dx) {}
};

KB_HD D1 operator+(D1 x, D1 y) { return D1(x.a + y.a, x.d + y.d); }
KB_HD D1 operator-(D1 x, D1 y) { return D1(x.a - y.a, x.d - y.d); }
KB_HD D1 operator-(D1 x) { return D1(-x.a, -x.d); }
KB_HD D1 operator*(D1 x, D1 y) { return D1(x.a * y.a, x.a * y.d + x.d * y.a); }
KB_HD D1 operator/(D1 x, D1 y) { const double inv = 1.0 / y.a, q = x.a * inv; return D1(q, (x.d - q * y.d) * inv); }

KB_HD double value(double x) { return x; }
KB_HD double value(D1 x) { return x.a; }

KB_HD double t_sqrt(double x) { return sqrt(x); }
KB_HD double t_sin(double x) { return sin(x); }
KB_HD double t_cos(double x) { return cos(x); }
KB_HD double t_atan(double x) { return atan(x); }
KB_HD D1 t_sqrt(D1 x) { const double s = sqrt(x.a); return D1(s, x.d / (2.0 * s)); }
KB_HD D1 t_sin(D1 x) { return D1(sin(x.a), cos(x.a) * x.d); }
KB_HD D1 t_cos(D1 x) { return D1(cos(x.a), -sin(x.a) * x.d); }
KB_HD D1 t_atan(D1 x) { return D1(atan(x.a), x.d / (1.0 + x.a * x.a)); }
KB_HD double t_tan(double x) { return tan(x); }
KB_HD D1 t_tan(D1 x) { const double t = tan(x.a); return D1(t, (1.0 + t * t) * x.d); }
KB_HD double deriv(double) { return 0.0; }
KB_HD double deriv(D1 x) { return x.d; }

}  // namespace kb

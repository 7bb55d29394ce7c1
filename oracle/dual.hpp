// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
// Nothing under oracle/ is shipped or called by the product path (kontiki_b200/).
//
// Forward-mode dual numbers ("Jet") with the semantics the reference gets from
// ceres::Jet<double, N> (Ceres 1.13/1.14, un-vendored dependency of hovren/kontiki;
// see SURVEY.md Appendix B).  The reference instantiates every residual functor
// with T = double and T = ceres::Jet<double, 4>
// (cpplib/include/kontiki/measurements/gyroscope_measurement.h:79 creates a
// ceres::DynamicAutoDiffCostFunction whose default stride is 4).
#pragma once
#include <cmath>

namespace kto {

template <int N>
struct Dual {
  double a;     // scalar part
  double v[N];  // infinitesimal part
  Dual() : a(0.0) { for (int i = 0; i < N; ++i) v[i] = 0.0; }   // Jet() is zero-initialised in Ceres 1.x
  Dual(double x) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0.0; }  // NOLINT implicit like ceres::Jet
  Dual(int x) : a(double(x)) { for (int i = 0; i < N; ++i) v[i] = 0.0; }  // NOLINT
};

#define KTO_DUAL_BIN(op)                                                        \
  template <int N> inline Dual<N> operator op(const Dual<N>& x, double y) { return x op Dual<N>(y); } \
  template <int N> inline Dual<N> operator op(double x, const Dual<N>& y) { return Dual<N>(x) op y; }

template <int N> inline Dual<N> operator+(const Dual<N>& x, const Dual<N>& y) {
  Dual<N> r; r.a = x.a + y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] + y.v[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& x, const Dual<N>& y) {
  Dual<N> r; r.a = x.a - y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] - y.v[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& x, const Dual<N>& y) {
  Dual<N> r; r.a = x.a * y.a; for (int i = 0; i < N; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
template <int N> inline Dual<N> operator/(const Dual<N>& x, const Dual<N>& y) {
  // ceres/jet.h: (a + u) / (b + v) = a/b + (u - (a/b) v)/b
  Dual<N> r; const double inv = 1.0 / y.a; r.a = x.a * inv; const double q = r.a;
  for (int i = 0; i < N; ++i) { r.v[i] = (x.v[i] - q * y.v[i]) * inv; }
  return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& x) {
  Dual<N> r; r.a = -x.a; for (int i = 0; i < N; ++i) r.v[i] = -x.v[i]; return r; }
KTO_DUAL_BIN(+) KTO_DUAL_BIN(-) KTO_DUAL_BIN(*) KTO_DUAL_BIN(/)
#undef KTO_DUAL_BIN
template <int N> inline Dual<N>& operator+=(Dual<N>& x, const Dual<N>& y) { x = x + y; return x; }
template <int N> inline Dual<N>& operator-=(Dual<N>& x, const Dual<N>& y) { x = x - y; return x; }
template <int N> inline Dual<N>& operator*=(Dual<N>& x, const Dual<N>& y) { x = x * y; return x; }

// Comparisons look at the scalar part only (ceres/jet.h).
#define KTO_DUAL_CMP(op)                                                                      \
  template <int N> inline bool operator op(const Dual<N>& x, const Dual<N>& y) { return x.a op y.a; } \
  template <int N> inline bool operator op(const Dual<N>& x, double y) { return x.a op y; }   \
  template <int N> inline bool operator op(double x, const Dual<N>& y) { return x op y.a; }
KTO_DUAL_CMP(<) KTO_DUAL_CMP(<=) KTO_DUAL_CMP(>) KTO_DUAL_CMP(>=) KTO_DUAL_CMP(==) KTO_DUAL_CMP(!=)
#undef KTO_DUAL_CMP

inline double val(double x) { return x; }
template <int N> inline double val(const Dual<N>& x) { return x.a; }

// Elementary functions: double overloads + Dual overloads, named as ceres:: does.
inline double ksqrt(double x) { return std::sqrt(x); }
inline double ksin(double x) { return std::sin(x); }
inline double kcos(double x) { return std::cos(x); }
inline double katan(double x) { return std::atan(x); }
inline double ktan(double x) { return std::tan(x); }
inline double katan2(double y, double x) { return std::atan2(y, x); }
inline double kexp(double x) { return std::exp(x); }
inline double kabs(double x) { return std::fabs(x); }
inline double kpow(double x, double p) { return std::pow(x, p); }

template <int N> inline Dual<N> scale_(const Dual<N>& x, double fa, double d) {
  Dual<N> r; r.a = fa; for (int i = 0; i < N; ++i) r.v[i] = d * x.v[i]; return r; }
template <int N> inline Dual<N> ksqrt(const Dual<N>& x) { const double s = std::sqrt(x.a); return scale_(x, s, 1.0 / (2.0 * s)); }
template <int N> inline Dual<N> ksin(const Dual<N>& x) { return scale_(x, std::sin(x.a), std::cos(x.a)); }
template <int N> inline Dual<N> kcos(const Dual<N>& x) { return scale_(x, std::cos(x.a), -std::sin(x.a)); }
template <int N> inline Dual<N> ktan(const Dual<N>& x) { const double t = std::tan(x.a); return scale_(x, t, 1.0 + t * t); }   // ceres/jet.h tan
template <int N> inline Dual<N> katan(const Dual<N>& x) { return scale_(x, std::atan(x.a), 1.0 / (1.0 + x.a * x.a)); }
template <int N> inline Dual<N> kexp(const Dual<N>& x) { const double e = std::exp(x.a); return scale_(x, e, e); }
template <int N> inline Dual<N> kabs(const Dual<N>& x) { return x.a < 0.0 ? -x : x; }
template <int N> inline Dual<N> kpow(const Dual<N>& x, double p) {
  // ceres/jet.h pow(Jet, double): (a+da)^p ~= a^p + p*a^(p-1) da
  return scale_(x, std::pow(x.a, p), p * std::pow(x.a, p - 1.0)); }
template <int N> inline Dual<N> katan2(const Dual<N>& y, const Dual<N>& x) {
  // ceres/jet.h atan2(g, f): (f dg - g df) / (f^2 + g^2)
  const double inv = 1.0 / (x.a * x.a + y.a * y.a);
  Dual<N> r; r.a = std::atan2(y.a, x.a);
  for (int i = 0; i < N; ++i) r.v[i] = inv * (x.a * y.v[i] - y.a * x.v[i]);
  return r; }

}  // namespace kto

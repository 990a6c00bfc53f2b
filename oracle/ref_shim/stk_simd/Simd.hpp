/* oracle/ref_shim: TEST INFRASTRUCTURE.  Stand-in for <stk_simd/Simd.hpp>: a
 * one-lane "SIMD" double (a distinct type, as in STK's scalar fallback) and
 * the stk::math functions the reference's master elements call.  Written for
 * this repo; not STK code. */
#ifndef NW_REF_SHIM_STK_SIMD_HPP
#define NW_REF_SHIM_STK_SIMD_HPP
#include <algorithm>
#include <cmath>

namespace stk {
namespace simd {

struct Bool
{
  bool b;
  Bool() : b(false) {}
  Bool(bool x) : b(x) {}
};
inline Bool operator&&(Bool a, Bool b) { return Bool(a.b && b.b); }
inline Bool operator||(Bool a, Bool b) { return Bool(a.b || b.b); }
inline Bool operator!(Bool a) { return Bool(!a.b); }

struct Double
{
  double v;
  Double() : v(0.0) {}
  Double(double x) : v(x) {}
  Double& operator+=(Double o) { v += o.v; return *this; }
  Double& operator-=(Double o) { v -= o.v; return *this; }
  Double& operator*=(Double o) { v *= o.v; return *this; }
  Double& operator/=(Double o) { v /= o.v; return *this; }
  Double operator-() const { return Double(-v); }
  Double operator+() const { return *this; }
  double& operator[](int) { return v; }
  const double& operator[](int) const { return v; }
};
inline Double operator+(Double a, Double b) { return Double(a.v + b.v); }
inline Double operator-(Double a, Double b) { return Double(a.v - b.v); }
inline Double operator*(Double a, Double b) { return Double(a.v * b.v); }
inline Double operator/(Double a, Double b) { return Double(a.v / b.v); }
inline Bool operator<(Double a, Double b) { return Bool(a.v < b.v); }
inline Bool operator<=(Double a, Double b) { return Bool(a.v <= b.v); }
inline Bool operator>(Double a, Double b) { return Bool(a.v > b.v); }
inline Bool operator>=(Double a, Double b) { return Bool(a.v >= b.v); }
inline Bool operator==(Double a, Double b) { return Bool(a.v == b.v); }
inline Bool operator!=(Double a, Double b) { return Bool(a.v != b.v); }

static constexpr int ndoubles = 1;
inline double& get_data(Double& d, int) { return d.v; }
inline const double& get_data(const Double& d, int) { return d.v; }
inline void set_data(Double& d, int, double x) { d.v = x; }
inline bool are_any(Bool b, int = 1) { return b.b; }
inline bool are_all(Bool b, int = 1) { return b.b; }

} // namespace simd

namespace math {
using simd::Bool;
using simd::Double;
inline Double sqrt(Double a) { return Double(std::sqrt(a.v)); }
inline Double cbrt(Double a) { return Double(std::cbrt(a.v)); }
inline Double abs(Double a) { return Double(std::fabs(a.v)); }
inline Double cos(Double a) { return Double(std::cos(a.v)); }
inline Double sin(Double a) { return Double(std::sin(a.v)); }
inline Double acos(Double a) { return Double(std::acos(a.v)); }
inline Double tanh(Double a) { return Double(std::tanh(a.v)); }
inline Double exp(Double a) { return Double(std::exp(a.v)); }
inline Double log(Double a) { return Double(std::log(a.v)); }
inline Double erf(Double a) { return Double(std::erf(a.v)); }
inline Double pow(Double a, Double b) { return Double(std::pow(a.v, b.v)); }
inline Double pow(Double a, double b) { return Double(std::pow(a.v, b)); }
inline Double pow(Double a, int b) { return Double(std::pow(a.v, b)); }
inline Double max(Double a, Double b) { return Double(a.v > b.v ? a.v : b.v); }
inline Double min(Double a, Double b) { return Double(a.v < b.v ? a.v : b.v); }
inline Double if_then_else(Bool c, Double a, Double b) { return c.b ? a : b; }
inline Double if_then_else_zero(Bool c, Double a) { return c.b ? a : Double(0.0); }
inline Double copysign(Double a, Double b) { return Double(std::copysign(a.v, b.v)); }

inline double sqrt(double a) { return std::sqrt(a); }
inline double cbrt(double a) { return std::cbrt(a); }
inline double abs(double a) { return std::fabs(a); }
inline double cos(double a) { return std::cos(a); }
inline double sin(double a) { return std::sin(a); }
inline double acos(double a) { return std::acos(a); }
inline double tanh(double a) { return std::tanh(a); }
inline double exp(double a) { return std::exp(a); }
inline double log(double a) { return std::log(a); }
inline double erf(double a) { return std::erf(a); }
inline double pow(double a, double b) { return std::pow(a, b); }
inline double max(double a, double b) { return a > b ? a : b; }
inline double min(double a, double b) { return a < b ? a : b; }
inline double if_then_else(bool c, double a, double b) { return c ? a : b; }
inline double if_then_else_zero(bool c, double a) { return c ? a : 0.0; }
inline double copysign(double a, double b) { return std::copysign(a, b); }
} // namespace math
} // namespace stk
#endif

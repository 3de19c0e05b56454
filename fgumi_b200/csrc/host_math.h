// f64 log-space helpers evaluated on the HOST with the platform libm (product code).
// Spec: fgumi-consensus phred.rs:66-346 (cited per function).  Shared by the table builder and the
// host-side ConsensusBaseBuilder used for RX (UMI) consensus.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <limits>

namespace fgb {
namespace hostmath {

constexpr double kLn10 = 2.302585092994046;
constexpr double kLn2 = 0.6931471805599453;
constexpr double kLnFourThirds = 0.2876820724517809;
static const double kNegInf = -std::numeric_limits<double>::infinity();

inline double ln_err_of_phred(unsigned q) { return -static_cast<double>(q) * kLn10 / 10.0; }  // phred.rs:66

inline double softplus(double x) {  // log(1+e^x), phred.rs:148-158
  if (x <= -37.0) return std::exp(x);
  if (x <= 18.0) return std::log1p(std::exp(x));
  if (x <= 33.3) return x + std::exp(-x);
  return x;
}

inline double ln_1m_exp(double x) {  // log(1-e^x), phred.rs:168-182
  if (x >= 0.0) return kNegInf;
  return x >= -kLn2 ? std::log(-std::expm1(x)) : std::log1p(-std::exp(x));
}

inline double ln_add(double a, double b) {  // phred.rs:274-285
  if (std::isinf(a) && a < 0.0) return b;
  if (std::isinf(b) && b < 0.0) return a;
  double lo = b < a ? b : a, hi = b < a ? a : b;
  return lo + softplus(hi - lo);
}

inline double ln_sub(double a, double b) {  // phred.rs:188-198
  if (std::isinf(b) && b < 0.0) return a;
  if (std::fabs(a - b) < DBL_EPSILON) return kNegInf;
  return a + ln_1m_exp(b - a);
}

inline double two_trials(double p, double r) {  // phred.rs:231-251
  double hi = p < r ? r : p, lo = p < r ? p : r;
  if (hi - lo >= 6.0) return hi;
  return ln_sub(ln_add(hi, lo), kLnFourThirds + hi + lo);
}


// phred.rs:307-330
inline double ln_add_array4(const double* v) {
  double mn = std::numeric_limits<double>::infinity();
  int mi = 0;
  for (int i = 0; i < 4; ++i) if (v[i] < mn) { mi = i; mn = v[i]; }
  if (std::isinf(mn)) return mn;
  double sum = mn;
  for (int i = 0; i < 4; ++i) if (i != mi) sum = ln_add(sum, v[i]);
  return sum;
}

}  // namespace hostmath
}  // namespace fgb

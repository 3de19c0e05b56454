// Host-side construction of the likelihood tables the vote kernel consumes (product code).
//
// Spec: fgumi-consensus base_builder.rs:252-278 (ConsensusBaseBuilder::new) and
// vanilla_caller.rs:463-494 (compute_single_input_consensus_quals), with the log-space helpers of
// phred.rs:66-346.  Evaluated once per handle with the host libm in f64, so the table bits are the
// ones the reference's callers hold; the device only ever adds/subtracts these values.
#include "host_tables.h"

#include <cfloat>
#include <cmath>
#include <limits>

namespace fgb {
namespace {

constexpr double kLn10 = 2.302585092994046;
constexpr double kLn2 = 0.6931471805599453;
constexpr double kLnFourThirds = 0.2876820724517809;
const double kNegInf = -std::numeric_limits<double>::infinity();

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

}  // namespace

unsigned host_ln_prob_to_phred(double ln_prob) {  // phred.rs:119-135
  const double max_as_ln = -93.0 * kLn10 / 10.0;
  if (ln_prob < max_as_ln) return 93;
  double p = std::floor(-10.0 * ln_prob / kLn10 + 0.001);
  if (std::isnan(p)) return 0;
  if (p < 2.0) p = 2.0;
  if (p > 93.0) p = 93.0;
  return static_cast<unsigned>(p);
}

void build_host_tables(unsigned pre, unsigned post, HostTables* t) {
  const double ln_post = ln_err_of_phred(post);
  const double ln3 = std::log(3.0);
  for (unsigned q = 0; q < 94; ++q) {
    double adj = two_trials(ln_post, ln_err_of_phred(q));
    t->correct[q] = ln_1m_exp(adj);
    t->err_alt[q] = adj - ln3;
  }
  t->ln_pre = ln_err_of_phred(pre);
  t->fast_qual = host_ln_prob_to_phred(t->ln_pre);

  const double ln_label = ln_err_of_phred(pre < post ? pre : post);
  for (unsigned q = 0; q < 94; ++q) {
    unsigned v = host_ln_prob_to_phred(two_trials(ln_err_of_phred(q), ln_label));
    t->single_q[q] = static_cast<uint8_t>(v > 93 ? 93 : v);
  }
  t->single_q[94] = t->single_q[95] = 0;

  // Fast-path proof table.  For a pileup of n identical A/C/G/T observations with qualities
  // q_i >= qT, winner_ll - loser_ll = sum_i (correct[q_i] - err_alt[q_i]) >= n * dmono[qT], where
  // dmono[q] = min_{q' >= q} (correct[q'] - err_alt[q']).  qt[n] is the smallest qT with
  // n * dmono[qT] > 23 + margin, so the reference's `> 23.0` test (base_builder.rs:364-375) is
  // guaranteed to pass; the margin (1e-6) dwarfs the <=1e-12 rounding of two Kahan sums.
  double dmono[94];
  double run = std::numeric_limits<double>::infinity();
  for (int q = 93; q >= 0; --q) {
    double d = t->correct[q] - t->err_alt[q];
    if (d < run) run = d;
    dmono[q] = run;
  }
  for (unsigned n = 0; n < 256; ++n) {
    unsigned qt = 255;
    for (unsigned q = 0; q < 94; ++q) {
      if (static_cast<double>(n) * dmono[q] > 23.0 + 1e-6) { qt = q; break; }
    }
    t->qt[n] = static_cast<uint8_t>(qt);
  }
}

}  // namespace fgb

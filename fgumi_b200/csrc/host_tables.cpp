// Host-side construction of the likelihood tables the vote kernel consumes (product code).
//
// Spec: fgumi-consensus base_builder.rs:252-278 (ConsensusBaseBuilder::new) and
// vanilla_caller.rs:463-494 (compute_single_input_consensus_quals), with the log-space helpers of
// phred.rs:66-346.  Evaluated once per handle with the host libm in f64, so the table bits are the
// ones the reference's callers hold; the device only ever adds/subtracts these values.
#include "host_tables.h"
#include "host_math.h"

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <limits>

namespace fgb {
using namespace hostmath;

unsigned host_ln_prob_to_phred(double ln_prob) {  // phred.rs:119-135
  const double max_as_ln = -93.0 * kLn10 / 10.0;
  if (ln_prob < max_as_ln) return 93;
  double p = std::floor(-10.0 * ln_prob / kLn10 + 0.001);
  if (std::isnan(p)) return 0;
  if (p < 2.0) p = 2.0;
  if (p > 93.0) p = 93.0;
  return static_cast<unsigned>(p);
}

namespace {
// base_builder.rs:295-458 for two observations of base index 0 with qualities q1, q2 (in that
// order): the 4-lane Kahan accumulation, the unanimous fast path and the full call().
unsigned pair_quality(const HostTables& t, unsigned q1, unsigned q2) {
  double ll[4] = {0.0, 0.0, 0.0, 0.0}, kc[4] = {0.0, 0.0, 0.0, 0.0};
  const unsigned qs[2] = {q1, q2};
  for (unsigned q : qs) {
    for (int i = 0; i < 4; ++i) {                       // :312-324
      const double v = i == 0 ? t.correct[q] : t.err_alt[q];
      const double y = v - kc[i];
      const double s = ll[i] + y;
      kc[i] = (s - ll[i]) - y;
      ll[i] = s;
    }
  }
  if (ll[0] - ll[1] > 23.0) return t.fast_qual;         // :338-379 (one observed base)
  const double ln_sum = ln_add_array4(ll);              // :401-457
  double mx = -std::numeric_limits<double>::infinity();
  int mi = -1;
  bool tie = false;
  for (int i = 0; i < 4; ++i) {
    const double v = ll[i];
    if (v > mx) { mx = v; mi = i; tie = false; }
    else if (v == mx) tie = true;
    else if (v < mx && std::fabs(v - mx) <= DBL_EPSILON) tie = true;
  }
  if (tie || mi != 0) return 255;                       // cannot happen for finite tables; be literal
  const double post = mx - ln_sum;
  const double err = ln_1m_exp(post);
  return host_ln_prob_to_phred(two_trials(t.ln_pre, err));
}
}  // namespace

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

  for (unsigned q1 = 0; q1 < 94; ++q1)
    for (unsigned q2 = 0; q2 < 94; ++q2)
      t->pair_q[q1 * 94 + q2] =
          (q1 == 0 || q2 == 0) ? 255 : static_cast<uint8_t>(pair_quality(*t, q1, q2));

  // ---- Proof tables (DESIGN.md "exactness") ------------------------------------------------------
  // Let D[q] = correct[q] - err_alt[q] and, for one position, S_b = sum of D[q_i] over the
  // observations of base b (S_b = 0 for an unobserved base).  In exact arithmetic
  // ll[w] - ll[b] = S_w - S_b, so g = S_w - max_{b != w} S_b is the winner's likelihood gap.
  //
  // (1) Reference fast path (base_builder.rs:338-379): one observed base and gap > 23.0
  //     => (w, phred(ln_pre)).
  // (2) "Dominant winner": if g >= G2 the full path (base_builder.rs:401-457) ALSO returns
  //     (w, phred(ln_pre)):  ln_sum = ll[w] + delta with 0 <= delta <= 3e^-g plus <= 64 ulp(|ll|)
  //     of rounding, so |posterior| <= 3e^-g + 64*2^-53*n*Vmax, ln_not(posterior) <= ln of that
  //     (or -inf), and ln_error_prob_two_trials(ln_pre, err) returns ln_pre unchanged as soon as
  //     ln_pre - err >= 6 (phred.rs:238).  G2 and nmax2 make both terms <= 0.5*e^(ln_pre-6.01).
  //     No tie is possible (gaps >> f64::EPSILON) and every table value involved is finite
  //     (quality 0, whose correct[] is -inf, is excluded through the dfix sentinel).
  // Everything here is conservative: a position that fails a proof is evaluated by the literal
  // f64 algorithm, so the proofs can only ever save work, never change a result.
  const double u = 1.1102230246251565e-16;   // 2^-53
  double vmax = 0.0;
  bool finite_ok = true;
  for (unsigned q = 1; q < 94; ++q) {
    if (!std::isfinite(t->correct[q]) || !std::isfinite(t->err_alt[q])) finite_ok = false;
    vmax = std::fmax(vmax, std::fmax(std::fabs(t->correct[q]), std::fabs(t->err_alt[q])));
  }
  const double lim = t->ln_pre - 6.01;
  t->g2 = -lim + std::log(6.0) + 1e-6;
  double nmax = 0.5 * std::exp(lim) / (64.0 * u * (vmax > 1.0 ? vmax : 1.0));
  if (!finite_ok || !(nmax >= 1.0)) nmax = 0.0;
  if (nmax > 1024.0) nmax = 1024.0;          // keeps the int32 fixed-point sums exact
  t->nmax2 = static_cast<uint32_t>(nmax);
  t->g2fix = static_cast<int32_t>(std::ceil(t->g2 * 65536.0));
  for (unsigned q = 0; q < 96; ++q) t->dfix[q] = INT32_MIN;
  for (unsigned q = 1; q < 94; ++q) {
    double d = t->correct[q] - t->err_alt[q];
    if (std::isfinite(d) && std::fabs(d) < 30000.0) t->dfix[q] = static_cast<int32_t>(std::llrint(d * 65536.0));
  }

  // SWAR fast-pass threshold: n identical A/C/G/T observations with qualities q_i >= qT have
  // gap = sum D[q_i] >= n * dmono[qT], dmono[q] = min_{q' >= q} D[q'].  qt[n] is the smallest qT
  // whose bound clears min(23, G2) (+1e-6, which dwarfs the <= 1e-9 rounding of the Kahan sums).
  double dmono[94];
  double run = std::numeric_limits<double>::infinity();
  for (int q = 93; q >= 0; --q) {
    double d = t->correct[q] - t->err_alt[q];
    if (d < run) run = d;
    dmono[q] = run;
  }
  for (unsigned n = 0; n < 256; ++n) {
    double thr = 23.0 + 1e-6;
    // qt[255] also serves every n > 255, so it keeps the reference's own 23.0 threshold
    if (n < 255 && n <= t->nmax2 && t->g2 < thr) thr = t->g2;
    unsigned qt = 255;
    for (unsigned q = 1; q < 94; ++q) {
      if (static_cast<double>(n) * dmono[q] > thr) { qt = q; break; }
    }
    t->qt[n] = static_cast<uint8_t>(qt);
  }

  // Near-unanimous threshold (vote_kernel.cuh vote_tile_deep): with at most kNearK dissenting observations the
  // winner's gap over any other base is at least (n - K) * dmono[q] - K * dmax (dmax = the largest D of the table).
  {
    double dmax = 0.0;
    for (unsigned q = 1; q < 94; ++q) { const double d = t->correct[q] - t->err_alt[q]; if (std::isfinite(d) && d > dmax) dmax = d; }
    for (unsigned n = 0; n < 256; ++n) {
      unsigned qt = 255;
      if (finite_ok && n > 2 * kNearK && n <= t->nmax2 && n < 255) {
        for (unsigned q = 1; q < 94; ++q)
          if (static_cast<double>(n - kNearK) * dmono[q] - static_cast<double>(kNearK) * dmax > t->g2) { qt = q; break; }
      }
      t->qt3[n] = static_cast<uint8_t>(qt);
    }
  }

  // Quality of a unanimous pileup as a step function of its likelihood gap (HostTables::ugap_*).  The tail is the
  // reference's (base_builder.rs:401-457 then phred.rs:119-135) on ll = {0, -g, -g, -g}: only differences of the
  // four sums enter it, and a unanimous pileup has one gap.  quality(g) is non-decreasing in g, so each step's start
  // is found by bisection over the integer gaps.  What the table cannot know is absorbed by kUgapGuard on the device
  // side: the reference evaluates the same tail on Kahan sums that carry a few ulp(|ll|) of absolute noise on
  // s = 3 e^-g -- relative noise 8 * 2^-53 * 32 n / s, at most 4e-4 nat (25 units) for n <= 4 at g = 23 -- the host
  // evaluation here has the same kind of noise, and the fixed-point gap is within 2 n + 1 units of the real one
  // (added separately by the kernel).  160 units = 2.4e-3 nat = 0.011 phred on either side of every step.
  {
    auto quality_at = [&](int32_t units) -> unsigned {
      const double g = static_cast<double>(units) / 65536.0;
      const double ll[4] = {0.0, -g, -g, -g};
      const double ln_sum = ln_add_array4(ll);
      const double err = ln_1m_exp(0.0 - ln_sum);
      return host_ln_prob_to_phred(two_trials(t->ln_pre, err));
    };
    for (int k = 0; k < 128; ++k) { t->ugap_bp[k] = INT32_MAX; t->ugap_q[k] = 0; }
    t->ugap_n = 0;
    const int32_t g_lo = 64, g_hi = 23 * 65536 + 4096;
    unsigned q_cur = quality_at(g_lo);
    int32_t start = g_lo;
    uint32_t n = 0;
    bool ok = finite_ok;
    while (ok && n < 127) {
      t->ugap_bp[n] = start; t->ugap_q[n] = static_cast<uint8_t>(q_cur); ++n;
      if (quality_at(g_hi) <= q_cur) break;              // no further step below the fast-path gap
      int32_t lo = start, hi = g_hi;                     // quality(lo) == q_cur < quality(hi)
      while (hi - lo > 1) {
        const int32_t mid = lo + (hi - lo) / 2;
        if (quality_at(mid) > q_cur) hi = mid; else lo = mid;
      }
      const unsigned q_next = quality_at(hi);
      if (q_next <= q_cur) { ok = false; break; }       // not monotone: leave the table empty
      start = hi; q_cur = q_next;
    }
    if (!ok || n >= 127) { n = 0; for (int k = 0; k < 128; ++k) t->ugap_bp[k] = INT32_MAX; }
    t->ugap_n = n;
  }

  // Sum-of-qualities thresholds (vote_kernel_w.cuh): f[k][s] = the smallest sum of D over k qualities in
  // 1..63 that add up to s, by dynamic programming; sumt[n] = the smallest t such that EVERY such multiset
  // with sum >= t clears the same threshold the per-read minimum test uses.  Exact over the table values;
  // the 1e-6 in the threshold covers the rounding of the reference's Kahan sums as above.
  for (unsigned n = 0; n < 8; ++n) t->sumt[n] = 0xFFFF;
  {
    const double inf = std::numeric_limits<double>::infinity();
    constexpr int kQ = 63, kN = 4;
    static thread_local double f[kN + 1][kQ * kN + 1];
    bool usable = finite_ok;
    for (int q = 1; q <= kQ; ++q) if (t->dfix[q] == INT32_MIN) usable = false;
    for (int k = 0; k <= kN; ++k) for (int s = 0; s <= kQ * kN; ++s) f[k][s] = inf;
    f[0][0] = 0.0;
    for (int k = 1; k <= kN; ++k)
      for (int s = k; s <= kQ * k; ++s) {
        double best = inf;
        for (int q = 1; q <= kQ && q <= s; ++q) {
          const double prev = f[k - 1][s - q];
          if (prev == inf) continue;
          const double v = prev + (t->correct[q] - t->err_alt[q]);
          if (v < best) best = v;
        }
        f[k][s] = best;
      }
    for (int n = 3; usable && n <= kN; ++n) {
      double thr = 23.0 + 1e-6;
      if (static_cast<uint32_t>(n) <= t->nmax2 && t->g2 < thr) thr = t->g2;
      int tmin = kQ * n + 1;                       // scan down while every sum above still clears thr
      for (int s = kQ * n; s >= n; --s) {
        if (f[n][s] > thr) tmin = s; else break;
      }
      if (tmin <= 255 && tmin <= kQ * n) t->sumt[n] = static_cast<uint16_t>(tmin);
    }
  }
}

}  // namespace fgb

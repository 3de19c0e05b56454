// Host-built likelihood tables (see host_tables.cpp).
#pragma once
#include <stdint.h>

namespace fgb {

struct HostTables {
  double correct[94];
  double err_alt[94];
  double ln_pre;
  uint8_t single_q[96];
  uint8_t qt[256];
  unsigned fast_qual;
  // "dominant winner" proof (host_tables.cpp): fixed-point likelihood gaps
  int32_t dfix[96];      // round((correct[q] - err_alt[q]) * 65536); INT32_MIN = unusable quality
  int32_t g2fix;         // ceil(G2 * 65536)
  uint32_t nmax2;        // proof valid for pileups of at most this many observations (0 = off)
  double g2;             // G2 in nats
  // Unanimous two-read pileups: pair_q[q1 * 94 + q2] = the quality base_builder's add/add/call
  // sequence yields for two observations of one base with qualities q1 then q2 (before the
  // min-consensus-quality threshold); 255 = not tabulated (quality 0), evaluate literally.
  uint8_t pair_q[94 * 94];
  // Sum-of-qualities proof of the fast path for shallow pileups: n unanimous observations whose qualities
  // all lie in 1..63 and sum to at least sumt[n] have sum D[q_i] > min(23, G2); 0xFFFF = not available.
  // Used for n = 3, 4 (the byte-wise sums of the kernel stay below 256).
  uint16_t sumt[8];
  // Near-unanimous proof of the deep kernel: a pileup of n A/C/G/T observations of which at most kNearK differ
  // from the rest, every quality >= qt3[n], has a likelihood gap >= (n - K) * dmono[qt3] - K * Dmax >= G2, so the
  // dominant-winner proof applies without looking at the dissenters' qualities; 255 = never.
  uint8_t qt3[256];
  // Unanimous pileups below the fast-path gap: the called quality is a monotone step function of the likelihood gap
  // g = sum D[q_i] alone (the three other bases share one sum).  ugap_bp[k] = the smallest fixed-point gap (units of
  // 2^-16 nat) at which the reference's call() tail -- evaluated on the host in f64 for ll = {0, -g, -g, -g} --
  // reaches ugap_q[k]; the quality of a gap in [ugap_bp[k], ugap_bp[k+1]) is ugap_q[k].  Unused entries hold
  // INT32_MAX.  The kernel accepts a table answer only when the gap's whole uncertainty interval, widened by
  // kUgapGuard units, lies inside one step (vote_kernel.cuh cert_unanimous).
  int32_t ugap_bp[128];
  uint8_t ugap_q[128];
  uint32_t ugap_n;
};
constexpr int32_t kUgapGuard = 160;      // units of 2^-16 nat: see host_tables.cpp
constexpr unsigned kNearK = 3;

void build_host_tables(unsigned pre, unsigned post, HostTables* t);
unsigned host_ln_prob_to_phred(double ln_prob);

}  // namespace fgb

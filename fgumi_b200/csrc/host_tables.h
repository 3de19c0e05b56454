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
};

void build_host_tables(unsigned pre, unsigned post, HostTables* t);
unsigned host_ln_prob_to_phred(double ln_prob);

}  // namespace fgb

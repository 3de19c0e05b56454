// Test harness (not product code): runs the product's header-only host helpers over a stream of
// length-prefixed records under AddressSanitizer / UBSan.  Built and driven by
// tests/test_host_prep.py::test_host_helpers_under_asan.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../fgumi_b200/csrc/host/bam.h"
#include "../../fgumi_b200/csrc/host/overlap.h"
#include "../../fgumi_b200/csrc/host/prep.h"
#include "../../fgumi_b200/csrc/host/record_filter.h"

using namespace fgb;

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  fgb_duplex_filter_params fp{};
  fp.cc.min_reads = 2; fp.cc.min_base_quality = 10; fp.cc.max_read_error_rate = 0.1; fp.cc.max_base_error_rate = 0.2;
  fp.cc.min_mean_base_quality = 20.0; fp.cc.max_no_call_fraction = 0.3;
  fp.ab_min_reads = 1; fp.ba_min_reads = 1; fp.ab_max_read_error_rate = fp.ba_max_read_error_rate = 0.1;
  fp.ab_max_base_error_rate = fp.ba_max_base_error_rate = 0.2; fp.require_ss_agreement = 1;
  uint32_t n = 0;
  unsigned long long checksum = 0, records = 0;
  std::vector<std::vector<uint8_t>> window;               // overlap pre-pass over runs of 8 records
  auto flush_window = [&]() {
    size_t total = 0;
    for (auto& r : window) total += r.size();
    std::vector<uint8_t> blob(total);
    std::vector<uint64_t> off(window.size() + 1, 0);
    for (size_t i = 0; i < window.size(); ++i) {
      std::copy(window[i].begin(), window[i].end(), blob.begin() + off[i]);
      off[i + 1] = off[i] + window[i].size();
    }
    overlap::Caller oc(overlap::kAgreeConsensus, overlap::kDisagreeConsensus);
    oc.apply_group(blob.data(), off.data(), static_cast<uint32_t>(window.size()));
    checksum += oc.stats.overlapping_bases + oc.stats.bases_corrected;
    window.clear();
  };
  while (std::fread(&n, 4, 1, f) == 1) {
    // exact-size heap block: any read past the record is an ASAN error
    std::vector<uint8_t> rec(n);
    if (n && std::fread(rec.data(), 1, n, f) != n) break;
    ++records;
    if (n < 32) continue;
    const bam::View v(rec.data(), n);
    if (!v.cigar_in_bounds() || v.aux_off() > n) continue;      // the guards of the C-ABI entry points
    std::vector<uint32_t> ops;
    bam::cigar_ops(v, &ops);
    checksum += bam::is_fr_pair(v, ops);
    const size_t clip = bam::num_bases_extending_past_mate(v, ops);
    checksum += clip;
    prep::PrepOptions po;
    prep::SourceRead sr;
    if (prep::make_source_read(po, v, 0, clip, &ops, &sr)) checksum += sr.bases.size();
    bam::SimpleCigar sc;
    bam::simplify_cigar(ops, &sc);
    size_t rc = 0;
    checksum += bam::clip_cigar_ops(ops, clip, v.flags() & bam::kReverse, &rc).size() + rc;
    size_t rp = 0;
    if (bam::read_pos_at_ref_pos(ops, static_cast<size_t>(v.pos() + 1), static_cast<size_t>(v.pos() + 5), true, &rp)) checksum += rp;
    window.push_back(rec);                                 // (before the filter masks it)
    if (window.size() == 8) flush_window();
    uint32_t masked = 0;
    checksum += rfilter::filter_record(rec.data(), n, fp, &masked) + masked;
  }
  if (!window.empty()) flush_window();
  std::fclose(f);
  std::printf("records %llu checksum %llu\n", records, checksum);
  return 0;
}

// Test harness (not product code): compiles the product's host caller (caller_host.cpp) with a
// sanitizer and drives planning-only callers (FGB_DEVICE_NONE) over a stream of MI groups, on one
// thread and on several, checking that both queue the same batch.  The engine symbols the caller
// references (fgb_submit...) resolve against the built libfgumi_b200.so and are never called here.
// Built and run by tests/test_caller_planning.py::test_host_caller_under_sanitizers.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/fgumi_b200.h"

static bool same_pending(fgb_caller* a, fgb_caller* b) {
  fgb_batch x, y;
  const fgb_duplex_job *dx, *dy; const fgb_codec_job *cx, *cy;
  uint64_t ndx, ndy, ncx, ncy;
  if (fgb_caller_pending(a, &x, &dx, &ndx, &cx, &ncx) != FGB_OK) return false;
  if (fgb_caller_pending(b, &y, &dy, &ndy, &cy, &ncy) != FGB_OK) return false;
  if (x.n_units != y.n_units || x.n_reads != y.n_reads || x.n_bytes != y.n_bytes || x.n_out != y.n_out ||
      ndx != ndy || ncx != ncy) return false;
  if (x.n_bytes && (std::memcmp(x.bases, y.bases, x.n_bytes) || std::memcmp(x.quals, y.quals, x.n_bytes))) return false;
  if (x.n_reads && std::memcmp(x.reads, y.reads, x.n_reads * sizeof(uint64_t))) return false;
  for (uint64_t u = 0; u < x.n_units; ++u)
    if (x.units[u].out_off != y.units[u].out_off || x.units[u].read_begin != y.units[u].read_begin ||
        x.units[u].cons_len != y.units[u].cons_len) return false;
  for (uint64_t j = 0; j < ndx; ++j)
    if (dx[j].unit_a != dy[j].unit_a || dx[j].unit_b != dy[j].unit_b || dx[j].out_off != dy[j].out_off) return false;
  for (uint64_t j = 0; j < ncx; ++j)
    if (std::memcmp(&cx[j], &cy[j], sizeof(fgb_codec_job))) return false;
  uint64_t sa[FGB_NSTATS], sb[FGB_NSTATS];
  fgb_caller_stats(a, sa); fgb_caller_stats(b, sb);
  return std::memcmp(sa, sb, sizeof(sa)) == 0;
}

int main(int argc, char** argv) {
  // file: u32 mode, u32 n_groups, then per group u32 n_records and per record u32 len + bytes
  if (argc < 3) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  const int threads = std::atoi(argv[2]);
  if (!f) return 2;
  uint32_t mode = 0, n_groups = 0;
  if (std::fread(&mode, 4, 1, f) != 1 || std::fread(&n_groups, 4, 1, f) != 1) return 2;
  std::vector<uint8_t> blob; std::vector<uint64_t> off{0}, grp{0};
  for (uint32_t g = 0; g < n_groups; ++g) {
    uint32_t nr = 0;
    if (std::fread(&nr, 4, 1, f) != 1) return 2;
    for (uint32_t r = 0; r < nr; ++r) {
      uint32_t n = 0;
      if (std::fread(&n, 4, 1, f) != 1) return 2;
      const size_t o = blob.size(); blob.resize(o + n);
      if (n && std::fread(blob.data() + o, 1, n, f) != n) return 2;
      off.push_back(blob.size());
    }
    grp.push_back(off.size() - 1);
  }
  std::fclose(f);
  fgb_caller_options o; std::memset(&o, 0, sizeof(o));
  o.mode = static_cast<uint8_t>(mode); o.error_rate_pre_umi = 45; o.error_rate_post_umi = 40; o.min_input_base_quality = 10;
  o.min_consensus_base_quality = 2; o.produce_per_base_tags = 1; o.min_reads = 1; o.min_xy_reads = 1; o.min_yx_reads = 0;
  o.tag[0] = 'M'; o.tag[1] = 'I'; o.read_name_prefix = "fgumi"; o.read_group_id = "A"; o.min_duplex_length = 1;
  o.consensus_call_overlapping_bases = mode == 2 ? 0 : 1;
  o.codec.single_strand_qual = -1; o.codec.outer_bases_qual = -1; o.codec.outer_bases_length = 5;
  o.codec.max_duplex_disagreements = 0xFFFFFFFFu; o.codec.max_duplex_disagreement_rate = 1.0;
  fgb_caller *one = nullptr, *many = nullptr;
  o.n_threads = 1;
  if (fgb_caller_create(FGB_DEVICE_NONE, &o, &one) != FGB_OK) return 3;
  o.n_threads = static_cast<uint32_t>(threads);
  if (fgb_caller_create(FGB_DEVICE_NONE, &o, &many) != FGB_OK) return 3;
  for (int rep = 0; rep < 2; ++rep) {               // two rounds: the pooled buffers are reused
    for (uint32_t g = 0; g < n_groups; ++g)
      if (fgb_caller_add_group(one, blob.data(), off.data() + grp[g], static_cast<uint32_t>(grp[g + 1] - grp[g])) != FGB_OK) return 4;
    if (fgb_caller_add_groups(many, blob.data(), off.data(), grp.data(), n_groups) != FGB_OK) return 5;
    if (!same_pending(one, many)) { std::printf("MISMATCH\n"); return 6; }
  }
  // the caller's thread pool: many fan-outs in a row on one caller (each add_groups call is one or two)
  for (int rep = 0; rep < 60; ++rep)
    if (fgb_caller_add_groups(many, blob.data(), off.data(), grp.data(), n_groups) != FGB_OK) return 8;
  // corrupted input: every group once more with a few bytes of one record overwritten (header fields,
  // CIGAR, sequence, tags) -- any status is fine, an invalid access is not (the sanitizer aborts)
  {
    uint64_t lcg = 12345;
    auto rnd = [&]() { lcg = lcg * 6364136223846793005ull + 1442695040888963407ull; return static_cast<uint32_t>(lcg >> 33); };
    unsigned long long ok = 0, refused = 0;
    for (uint32_t g = 0; g < n_groups; ++g) {
      const uint64_t r0 = grp[g], r1 = grp[g + 1];
      if (r1 == r0) continue;
      std::vector<uint8_t> copy(blob.begin() + off[r0], blob.begin() + off[r1]);       // exact-size block
      std::vector<uint64_t> o2(r1 - r0 + 1);
      for (uint64_t r = r0; r <= r1; ++r) o2[r - r0] = off[r] - off[r0];
      const uint64_t victim = r0 + rnd() % (r1 - r0);
      const uint64_t vb = off[victim] - off[r0], vl = off[victim + 1] - off[victim];
      const int mode = rnd() % 4;
      if (mode == 0) for (int k = 0; k < 3; ++k) copy[vb + rnd() % vl] = static_cast<uint8_t>(rnd());
      else if (mode == 1 && vl >= 20) { uint32_t v = rnd() % 70000; std::memcpy(&copy[vb + 16], &v, 4); }      // l_seq
      else if (mode == 2 && vl >= 14) { uint16_t v = static_cast<uint16_t>(rnd()); std::memcpy(&copy[vb + 12], &v, 2); }   // n_cigar_op
      else if (vl >= 9) copy[vb + 8] = static_cast<uint8_t>(rnd());                                                  // l_read_name
      fgb_status st = fgb_caller_add_group(one, copy.data(), o2.data(), static_cast<uint32_t>(r1 - r0));
      if (st == FGB_OK) ++ok; else ++refused;
    }
    std::printf("fuzz: %llu accepted, %llu refused\n", ok, refused);
  }
  const uint8_t* d; uint64_t n, c;
  if (fgb_caller_flush(one, &d, &n, &c) != FGB_ERR_NO_DEVICE) return 7;     // planning only: loud refusal
  fgb_batch b; fgb_caller_pending(many, &b, nullptr, nullptr, nullptr, nullptr);
  std::printf("ok units %llu reads %llu\n", (unsigned long long)b.n_units, (unsigned long long)b.n_reads);
  fgb_caller_destroy(one); fgb_caller_destroy(many);
  return 0;
}

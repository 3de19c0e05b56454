// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See fgumi_oracle.hpp for the rules.
// Plain-C entry points over the oracle so tests/ and bench.py's cpu_baseline leg can drive it
// through ctypes.  Batch arrays use the same SoA layout as include/fgumi_b200.h so that the
// CUDA path and the oracle consume the identical buffers.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "fgumi_oracle.hpp"

using namespace fgoracle;

extern "C" {

// ---- phred.rs scalars ------------------------------------------------------------------------
double orc_phred_to_ln_error_prob(uint8_t q) { return phred_to_ln_error_prob(q); }
double orc_phred_to_ln_correct_prob(uint8_t q) { return phred_to_ln_correct_prob(q); }
uint8_t orc_ln_prob_to_phred(double x) { return ln_prob_to_phred(x); }
double orc_log1pexp(double x) { return log1pexp(x); }
double orc_ln_one_minus_exp(double x) { return ln_one_minus_exp(x); }
double orc_ln_a_minus_b(double a, double b) { return ln_a_minus_b(a, b); }
double orc_ln_error_prob_two_trials(double a, double b) { return ln_error_prob_two_trials(a, b); }
double orc_ln_sum_exp(double a, double b) { return ln_sum_exp(a, b); }
double orc_ln_sum_exp_array(const double* v, size_t n) { return ln_sum_exp_array(v, n); }

// ---- base builder ----------------------------------------------------------------------------
// One pileup through ConsensusBaseBuilder: add(base[i], qual[i]) in order, then call().
// obs_out[4] receives the per-base observation counts.
void orc_builder_call(uint8_t pre, uint8_t post, const uint8_t* bases, const uint8_t* quals,
                      size_t n, uint8_t* base_out, uint8_t* qual_out, uint16_t* obs_out,
                      double* ll_out) {
  ConsensusBaseBuilder b(pre, post);
  for (size_t i = 0; i < n; ++i) b.add(bases[i], quals[i]);
  b.call(base_out, qual_out);
  if (obs_out) std::memcpy(obs_out, b.observations, sizeof(b.observations));
  if (ll_out) std::memcpy(ll_out, b.likelihoods, sizeof(b.likelihoods));
}

void orc_tables(uint8_t pre, uint8_t post, double* correct, double* err_alt, double* ln_pre,
                uint8_t* single_q) {
  ConsensusBaseBuilder b(pre, post);
  if (correct) std::memcpy(correct, b.adjusted_correct_table, sizeof(double) * 94);
  if (err_alt) std::memcpy(err_alt, b.adjusted_error_per_alt, sizeof(double) * 94);
  if (ln_pre) *ln_pre = b.ln_error_pre_umi;
  if (single_q) {
    auto v = compute_single_input_consensus_quals(pre, post);
    std::memcpy(single_q, v.data(), 94);
  }
}

// ---- simplex vote over a packed batch --------------------------------------------------------
// reads[r] = (off << 16) | len ; units[u] = {u64 out_off; u32 read_begin; u32 cons_len};
// units[n_units] is the sentinel.  The oracle recomputes the consensus length itself
// (vanilla_caller.rs:1269-1277) and writes it to cons_len_out[u] so tests can check the host's.
struct OrcUnit {
  uint64_t out_off;
  uint32_t read_begin;
  uint32_t cons_len;
};

int orc_simplex_batch(uint64_t n_units, const OrcUnit* units, const uint64_t* reads,
                      const uint8_t* bases, const uint8_t* quals, uint8_t pre, uint8_t post,
                      uint32_t min_reads, uint8_t min_cons_q, uint8_t* out_base,
                      uint8_t* out_qual, uint16_t* out_depth, uint16_t* out_errors,
                      uint32_t* cons_len_out, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  std::atomic<uint64_t> next{0};
  std::atomic<int> rc{0};
  const uint64_t CHUNK = 256;
  auto worker = [&]() {
    // one caller per worker, like the reference's per-thread callers (simplex.rs:574)
    VanillaOptions opt;
    opt.error_rate_pre_umi = pre;
    opt.error_rate_post_umi = post;
    opt.min_reads = min_reads;
    opt.min_consensus_base_quality = min_cons_q;
    ConsensusBaseBuilder builder(pre, post);
    std::vector<uint8_t> single_q = compute_single_input_consensus_quals(pre, post);
    std::vector<SourceRow> rows;
    ConsensusColumns cols;
    for (;;) {
      uint64_t lo = next.fetch_add(CHUNK);
      if (lo >= n_units) break;
      uint64_t hi = std::min(n_units, lo + CHUNK);
      for (uint64_t u = lo; u < hi; ++u) {
        uint32_t rb = units[u].read_begin, re = units[u + 1].read_begin;
        rows.clear();
        for (uint32_t r = rb; r < re; ++r) {
          uint64_t off = reads[r] >> 16;
          size_t len = reads[r] & 0xFFFF;
          rows.push_back(SourceRow{bases + off, quals + off, len});
        }
        if (rows.size() < min_reads || rows.empty()) {  // consensus_call :633-635 returns None
          if (cons_len_out) cons_len_out[u] = 0;
          continue;
        }
        if (!create_consensus_from_source_reads(rows.data(), rows.size(), opt, builder, single_q,
                                                &cols)) {
          rc.store(1);
          continue;
        }
        size_t L = cols.bases.size();
        if (cons_len_out) cons_len_out[u] = static_cast<uint32_t>(L);
        uint64_t o = units[u].out_off;
        std::memcpy(out_base + o, cols.bases.data(), L);
        std::memcpy(out_qual + o, cols.quals.data(), L);
        std::memcpy(out_depth + o, cols.depths.data(), L * 2);
        std::memcpy(out_errors + o, cols.errors.data(), L * 2);
      }
    }
  };
  if (n_threads == 1) {
    worker();
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(worker);
    for (auto& t : th) t.join();
  }
  return rc.load();
}

// ---- duplex combine --------------------------------------------------------------------------
// source rows: n_source × (ptr,len) given as parallel arrays; pass n_source = -1 for the
// approximate error branch.
void orc_duplex_combine(const uint8_t* ab, const uint8_t* aq, const uint16_t* ad,
                        const uint16_t* ae, const uint8_t* bb, const uint8_t* bq,
                        const uint16_t* bd, const uint16_t* be, size_t len,
                        const uint8_t* const* src_bases, const size_t* src_len, long n_source,
                        uint8_t* ob, uint8_t* oq, uint16_t* oe) {
  std::vector<SourceRow> rows;
  if (n_source >= 0)
    for (long i = 0; i < n_source; ++i) rows.push_back(SourceRow{src_bases[i], nullptr, src_len[i]});
  static const SourceRow kEmpty{nullptr, nullptr, 0};
  const SourceRow* src = n_source >= 0 ? (rows.empty() ? &kEmpty : rows.data()) : nullptr;
  duplex_combine(ab, aq, ad, ae, bb, bq, bd, be, len, src, rows.size(), ob, oq, oe);
}

int orc_duplex_job(const uint8_t* ab, const uint8_t* aq, const uint16_t* ad, const uint16_t* ae,
                   size_t la, const uint8_t* bb, const uint8_t* bq, const uint16_t* bd,
                   const uint16_t* be, size_t lb, const uint8_t* const* src_bases,
                   const size_t* src_len, long n_source, uint8_t* ob, uint8_t* oq, uint16_t* oe,
                   size_t* out_len) {
  std::vector<SourceRow> rows;
  for (long i = 0; i < n_source; ++i) rows.push_back(SourceRow{src_bases[i], nullptr, src_len[i]});
  // duplex_caller.rs:2001-2011: `if source_reads.is_empty() { None } else { Some(..) }`
  const SourceRow* src = rows.empty() ? nullptr : rows.data();
  return duplex_consensus_arms(ab, aq, ad, ae, la, bb, bq, bd, be, lb, src, rows.size(), ob, oq, oe,
                               out_len);
}

// ---- codec combine ---------------------------------------------------------------------------
void orc_codec_combine(const uint8_t* ab, const uint8_t* aq, const uint16_t* ad,
                       const uint16_t* ae, const uint8_t* bb, const uint8_t* bq,
                       const uint16_t* bd, const uint16_t* be, size_t len, uint8_t* ob,
                       uint8_t* oq, uint16_t* od, uint16_t* oe, uint64_t* duplex_bases,
                       uint64_t* disagreements) {
  CodecCombineResult r = codec_combine_padded(ab, aq, ad, ae, bb, bq, bd, be, len, ob, oq, od, oe);
  if (duplex_bases) *duplex_bases = r.duplex_bases_count;
  if (disagreements) *disagreements = r.duplex_disagreements;
}

void orc_codec_mask(const uint8_t* cons_bases, uint8_t* cons_quals, size_t len,
                    const uint8_t* r1_bases, const uint8_t* r2_bases, int ss_qual, int outer_qual,
                    size_t outer_len) {
  codec_mask_quals(cons_bases, cons_quals, len, r1_bases, r2_bases, ss_qual, outer_qual, outer_len);
}

// Full CODEC tail for one molecule.  out_* must hold `consensus_length` elements.
int orc_codec_job(const uint8_t* ab, const uint8_t* aq, const uint16_t* ad, const uint16_t* ae,
                  size_t la, const uint8_t* bb, const uint8_t* bq, const uint16_t* bd,
                  const uint16_t* be, size_t lb, int r1_neg, int r2_neg, size_t cons_len,
                  int ss_qual, int outer_qual, size_t outer_len, size_t max_dis, double max_rate,
                  uint8_t* ob, uint8_t* oq, uint16_t* od, uint16_t* oe, uint64_t* duplex_bases,
                  uint64_t* disagreements) {
  SsColumns a, b;
  a.bases.assign(ab, ab + la); a.quals.assign(aq, aq + la);
  a.depths.assign(ad, ad + la); a.errors.assign(ae, ae + la);
  b.bases.assign(bb, bb + lb); b.quals.assign(bq, bq + lb);
  b.depths.assign(bd, bd + lb); b.errors.assign(be, be + lb);
  CodecJobResult r = codec_job(a, b, r1_neg != 0, r2_neg != 0, cons_len, ss_qual, outer_qual,
                               outer_len, max_dis, max_rate);
  size_t n = r.consensus.bases.size();
  std::memcpy(ob, r.consensus.bases.data(), n);
  std::memcpy(oq, r.consensus.quals.data(), n);
  std::memcpy(od, r.consensus.depths.data(), n * 2);
  std::memcpy(oe, r.consensus.errors.data(), n * 2);
  if (duplex_bases) *duplex_bases = r.duplex_bases_count;
  if (disagreements) *disagreements = r.duplex_disagreements;
  return r.status;
}

}  // extern "C"

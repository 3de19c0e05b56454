// Host side of the simplex caller (product code): everything VanillaUmiConsensusCaller does
// around the per-position vote, re-shaped for batches.
//
//   add_group()  = process_group up to the vote (vanilla_caller.rs:1042-1227): filter, sub-group,
//                  mate-overlap clip, create_source_read, CIGAR filter, min-reads / orphan rules,
//                  statistics — and packs the surviving SourceRead rows straight into the SoA
//                  columns of include/fgumi_b200.h;
//   flush()      = one fgb_submit/fgb_wait for everything packed, then
//                  build_consensus_record_into (vanilla_caller.rs:1365-1473) per unit, in input
//                  order, into one ConsensusOutput byte stream (caller.rs:173-178).
// Whether a sub-group yields a consensus is known before the vote (the vote cannot fail), so the
// orphan rule (vanilla_caller.rs:1089-1108) is applied at add time and only emitted units reach
// the GPU.  There is no CPU vote here: without a device, fgb_caller_create fails.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/fgumi_b200.h"
#include "../host_math.h"
#include "../host_tables.h"
#include "bam.h"

namespace fgb {
namespace {

using bam::View;

// Host ConsensusBaseBuilder for the RX (UMI) consensus only (simple_umi.rs:36-134): f64, literal.
struct UmiBuilder {
  HostTables t;
  UmiBuilder() { build_host_tables(90, 90, &t); }   // simple_umi.rs:13-19 defaults (90, 90, Q20)
  // One column of characters, all DNA (A/C/G/T/N, any case); returns the called base.
  uint8_t call(const std::vector<uint8_t>& col) const {
    using namespace hostmath;
    double ll[4] = {0, 0, 0, 0}, kc[4] = {0, 0, 0, 0};
    uint32_t obs[4] = {0, 0, 0, 0};
    const double c = t.correct[20], e = t.err_alt[20];
    for (uint8_t ch : col) {
      int idx;
      switch (ch) {
        case 'A': case 'a': idx = 0; break;
        case 'C': case 'c': idx = 1; break;
        case 'G': case 'g': idx = 2; break;
        case 'T': case 't': idx = 3; break;
        default: idx = -1;
      }
      if (idx < 0) continue;   // 'N' is ignored by add(), base_builder.rs:300
      for (int i = 0; i < 4; ++i) {   // base_builder.rs:312-324
        double v = i == idx ? c : e;
        double y = v - kc[i];
        double s = ll[i] + y;
        kc[i] = (s - ll[i]) - y;
        ll[i] = s;
      }
      obs[idx]++;
    }
    const uint32_t depth = obs[0] + obs[1] + obs[2] + obs[3];
    if (depth == 0) return 'N';
    int kinds = (obs[0] != 0) + (obs[1] != 0) + (obs[2] != 0) + (obs[3] != 0);
    static const char kB[5] = "ACGT";
    if (kinds == 1) {
      int w = obs[0] ? 0 : obs[1] ? 1 : obs[2] ? 2 : 3;
      if (ll[w] - ll[(w + 1) % 4] > 23.0) return kB[w];
    }
    double mx = kNegInf;
    int mi = -1;
    bool tie = false;
    for (int i = 0; i < 4; ++i) {   // base_builder.rs:413-431
      if (ll[i] > mx) { mx = ll[i]; mi = i; tie = false; }
      else if (ll[i] == mx) tie = true;
      else if (ll[i] < mx && std::fabs(ll[i] - mx) <= DBL_EPSILON) tie = true;
    }
    if (tie || mi < 0) return 'N';
    return kB[mi];
  }
};

enum ReadType : uint8_t { kFragment = 0, kR1 = 1, kR2 = 2 };

struct SourceRead {
  uint32_t original_idx;
  std::vector<uint8_t> bases, quals;
  bam::SimpleCigar cigar;
};

struct UnitMeta {
  uint8_t read_type;
  std::string umi;
  std::vector<std::string> rx;      // RX values of the surviving reads, in order
  bool has_cell = false;
  std::string cell;
};

}  // namespace
}  // namespace fgb

using namespace fgb;

struct fgb_caller {
  fgb_caller_options opt{};
  std::string prefix, rg;
  fgb_handle* h = nullptr;
  UmiBuilder umi_builder;
  uint64_t stats[FGB_NSTATS] = {0};
  // packed batch
  std::vector<uint8_t> bases, quals;
  std::vector<uint64_t> reads;
  std::vector<fgb_unit> units;
  std::vector<UnitMeta> metas;
  uint64_t n_out = 0;
  // output of the last flush
  std::vector<uint8_t> out;
  uint64_t out_count = 0;
  std::string last_error;
  // scratch
  std::vector<uint32_t> ops;
  std::vector<uint8_t> seq;
};

namespace {

size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }

// find_quality_trim_point, vanilla_caller.rs:780-804
size_t quality_trim_point(const std::vector<uint8_t>& q, uint8_t trim_qual) {
  size_t length = q.size();
  if (trim_qual < 1 || length == 0) return 0;
  int32_t score = 0, max_score = 0;
  size_t trim_point = length;
  for (size_t i = length; i-- > 0;) {
    score += static_cast<int32_t>(trim_qual) - static_cast<int32_t>(q[i]);
    if (score < 0) break;
    if (score > max_score) { max_score = score; trim_point = i; }
  }
  return trim_point;
}

// create_source_read, vanilla_caller.rs:863-955
bool make_source_read(fgb_caller* c, const View& v, uint32_t idx, size_t mate_clip, SourceRead* sr) {
  const bool neg = v.flags() & bam::kReverse;
  const uint8_t min_bq = c->opt.min_input_base_quality;
  bam::decode_sequence(v, &sr->bases);
  const uint32_t read_len = v.l_seq();
  if (read_len == 0 || v.qual_off() + read_len > v.n) return false;
  sr->quals.assign(v.b + v.qual_off(), v.b + v.qual_off() + read_len);
  bool all_ff = true;
  for (uint8_t q : sr->quals) if (q != 0xFF) { all_ff = false; break; }
  if (all_ff) return false;
  if (neg) {
    std::reverse(sr->bases.begin(), sr->bases.end());
    for (auto& b : sr->bases) b = bam::complement(b);
    std::reverse(sr->quals.begin(), sr->quals.end());
  }
  const size_t trim_to = c->opt.trim ? quality_trim_point(sr->quals, min_bq) : read_len;
  for (size_t i = 0; i < trim_to; ++i)
    if (sr->quals[i] < min_bq) { sr->bases[i] = 'N'; sr->quals[i] = 2; }
  const size_t clip_position = read_len > mate_clip ? read_len - mate_clip : 0;
  size_t final_len = std::min(clip_position, trim_to);
  while (final_len > 0 && sr->bases[final_len - 1] == 'N') --final_len;
  if (final_len == 0) return false;
  sr->bases.resize(final_len);
  sr->quals.resize(final_len);
  bam::cigar_ops(v, &c->ops);
  bam::simplify_cigar(c->ops, &sr->cigar);
  if (neg) std::reverse(sr->cigar.begin(), sr->cigar.end());
  bam::truncate_cigar(&sr->cigar, final_len);
  sr->original_idx = idx;
  return true;
}

// filter_source_reads_by_alignment + select_most_common_alignment_group, vanilla_caller.rs:47-119,961-1013
size_t filter_by_alignment(std::vector<SourceRead>* srs) {
  const size_t n = srs->size();
  if (n < 2) return 0;
  std::vector<uint32_t> order(n);
  for (uint32_t i = 0; i < n; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
    return (*srs)[a].bases.size() > (*srs)[b].bases.size();
  });
  struct Group { const bam::SimpleCigar* cigar; std::vector<uint32_t> members; };
  std::vector<Group> groups;
  for (uint32_t idx : order) {
    const bam::SimpleCigar& cg = (*srs)[idx].cigar;
    bool found = false;
    for (auto& g : groups)
      if (bam::is_cigar_prefix(cg, *g.cigar)) { g.members.push_back(idx); found = true; }   // no break (fgbio)
    if (!found) groups.push_back(Group{&cg, {idx}});
  }
  // Iterator::max_by keeps the LAST maximum: larger group wins, then the smaller CIGAR
  const Group* best = nullptr;
  for (const auto& g : groups) {
    if (!best) { best = &g; continue; }
    int cmp = g.members.size() < best->members.size() ? -1 : (g.members.size() > best->members.size() ? 1 : 0);
    if (cmp == 0) cmp = bam::cmp_cigar(*best->cigar, *g.cigar);
    if (cmp >= 0) best = &g;
  }
  std::vector<char> keep(n, 0);
  for (uint32_t i : best->members) keep[i] = 1;
  size_t kept = 0;
  for (char k : keep) kept += k;
  std::vector<SourceRead> out;
  out.reserve(kept);
  for (size_t i = 0; i < n; ++i) if (keep[i]) out.push_back(std::move((*srs)[i]));
  srs->swap(out);
  return n - kept;
}

void reject(fgb_caller* c, int reason, uint64_t n) {
  c->stats[FGB_STAT_FILTERED_READS] += n;
  c->stats[reason] += n;
}

struct Prepared {
  bool ok = false;
  size_t surviving = 0;
  std::vector<SourceRead> srs;
  std::vector<uint32_t> rec_idx;   // indices (into the group's records) of the surviving reads
};

// process_subgroup up to the vote, vanilla_caller.rs:1124-1227
void prepare_subgroup(fgb_caller* c, const std::vector<View>& recs, const std::vector<uint32_t>& members,
                      Prepared* p) {
  p->ok = false; p->surviving = 0; p->srs.clear(); p->rec_idx.clear();
  const size_t min_reads = c->opt.min_reads;
  if (members.empty()) return;
  if (members.size() < min_reads) { reject(c, FGB_STAT_REJ_INSUFFICIENT_READS, members.size()); return; }
  size_t zero = 0;
  for (uint32_t k = 0; k < members.size(); ++k) {
    const View& v = recs[members[k]];
    bam::cigar_ops(v, &c->ops);
    size_t clip = bam::num_bases_extending_past_mate(v, c->ops);
    SourceRead sr;
    if (make_source_read(c, v, k, clip, &sr)) p->srs.push_back(std::move(sr));
    else ++zero;
  }
  if (zero) reject(c, FGB_STAT_REJ_ZERO_LENGTH, zero);
  if (p->srs.size() < min_reads) {
    if (!p->srs.empty()) reject(c, FGB_STAT_REJ_INSUFFICIENT_READS, p->srs.size());
    return;
  }
  size_t minority = filter_by_alignment(&p->srs);
  if (minority) reject(c, FGB_STAT_REJ_MINORITY_ALIGNMENT, minority);
  if (p->srs.size() < min_reads) {
    if (!p->srs.empty()) reject(c, FGB_STAT_REJ_INSUFFICIENT_READS, p->srs.size());
    return;
  }
  p->ok = true;
  p->surviving = p->srs.size();
  for (auto& sr : p->srs) p->rec_idx.push_back(members[sr.original_idx]);
}

// Appends one unit (its SourceRead rows) to the packed batch.
void pack_unit(fgb_caller* c, const std::vector<View>& recs, const Prepared& p, uint8_t read_type,
               const std::string& umi) {
  fgb_unit u;
  u.out_off = c->n_out;
  u.read_begin = static_cast<uint32_t>(c->reads.size());
  std::vector<size_t> lens;
  for (const auto& sr : p.srs) {
    size_t off = c->bases.size();
    size_t len = sr.bases.size();
    c->reads.push_back(FGB_READ_DESC(off, len));
    c->bases.insert(c->bases.end(), sr.bases.begin(), sr.bases.end());
    c->quals.insert(c->quals.end(), sr.quals.begin(), sr.quals.end());
    size_t pad = round_up(len, FGB_READ_ALIGN) - len;
    c->bases.insert(c->bases.end(), pad, 0);
    c->quals.insert(c->quals.end(), pad, 0);
    lens.push_back(len);
  }
  std::sort(lens.begin(), lens.end(), std::greater<size_t>());
  u.cons_len = static_cast<uint32_t>(lens[c->opt.min_reads - 1]);   // vanilla_caller.rs:1269-1277
  c->units.push_back(u);
  c->n_out += round_up(u.cons_len, FGB_OUT_ALIGN);
  UnitMeta m;
  m.read_type = read_type;
  m.umi = umi;
  for (uint32_t ri : p.rec_idx) {
    const uint8_t* val; size_t n;
    if (bam::find_string_tag(recs[ri], "RX", &val, &n)) m.rx.emplace_back(reinterpret_cast<const char*>(val), n);
  }
  if (c->opt.cell_tag[0] && !p.rec_idx.empty()) {
    const uint8_t* val; size_t n;
    if (bam::find_string_tag(recs[p.rec_idx[0]], c->opt.cell_tag, &val, &n)) {
      m.has_cell = true;
      m.cell.assign(reinterpret_cast<const char*>(val), n);
    }
  }
  c->metas.push_back(std::move(m));
}

// consensus_umis, simple_umi.rs:65-122,236-245.  Returns false for the reference's panics
// (length mismatch, DNA / non-DNA mix).
bool consensus_umis(const fgb_caller* c, const std::vector<std::string>& umis, std::string* out) {
  out->clear();
  if (umis.empty()) return true;
  if (umis.size() == 1) { *out = umis[0]; return true; }
  const std::string& first = umis[0];
  for (const auto& s : umis) if (s.size() != first.size()) return false;
  auto is_dna = [](uint8_t ch) {
    switch (ch) { case 'A': case 'C': case 'G': case 'T': case 'N':
                  case 'a': case 'c': case 'g': case 't': case 'n': return true; default: return false; }
  };
  std::vector<uint8_t> col(umis.size());
  for (size_t i = 0; i < first.size(); ++i) {
    size_t non_dna = 0;
    for (size_t k = 0; k < umis.size(); ++k) {
      col[k] = static_cast<uint8_t>(umis[k][i]);
      if (!is_dna(col[k])) {
        ++non_dna;
        if (col[k] != static_cast<uint8_t>(first[i])) return false;
      }
    }
    if (non_dna == 0) out->push_back(static_cast<char>(c->umi_builder.call(col)));
    else if (non_dna == umis.size()) out->push_back(first[i]);
    else return false;
  }
  return true;
}

}  // namespace

extern "C" {

fgb_status fgb_caller_create(int device, const fgb_caller_options* opt, fgb_caller** out) {
  if (!opt || !out || opt->min_reads == 0 || !opt->read_name_prefix || !opt->read_group_id)
    return FGB_ERR_INVALID_ARG;
  *out = nullptr;
  if (opt->mode != FGB_MODE_SIMPLEX) return FGB_ERR_INVALID_ARG;
  std::unique_ptr<fgb_caller> c(new fgb_caller());
  c->opt = *opt;
  c->prefix = opt->read_name_prefix;
  c->rg = opt->read_group_id;
  c->opt.read_name_prefix = nullptr;
  c->opt.read_group_id = nullptr;
  fgb_params p;
  p.error_rate_pre_umi = opt->error_rate_pre_umi;
  p.error_rate_post_umi = opt->error_rate_post_umi;
  p.min_consensus_base_quality = opt->min_consensus_base_quality;
  p.reserved0 = 0;
  p.min_reads = opt->min_reads;
  fgb_status st = fgb_create(device, &p, &c->h);
  if (st != FGB_OK) return st;
  *out = c.release();
  return FGB_OK;
}

void fgb_caller_destroy(fgb_caller* c) {
  if (!c) return;
  fgb_destroy(c->h);
  delete c;
}

size_t fgb_caller_last_error(const fgb_caller* c, char* buf, size_t buf_len) {
  if (!c) return 0;
  if (buf && buf_len) {
    size_t n = std::min(buf_len - 1, c->last_error.size());
    std::memcpy(buf, c->last_error.data(), n);
    buf[n] = 0;
  }
  return c->last_error.size();
}

// consensus_reads for one MI group (vanilla_caller.rs:1477-1499 + process_group :1042-1114).
fgb_status fgb_caller_add_group(fgb_caller* c, const uint8_t* records, const uint64_t* rec_off,
                                uint32_t n_records) {
  if (!c || (n_records && (!records || !rec_off))) return FGB_ERR_INVALID_ARG;
  if (n_records == 0) return FGB_OK;
  std::vector<View> recs;
  recs.reserve(n_records);
  for (uint32_t i = 0; i < n_records; ++i) {
    size_t len = static_cast<size_t>(rec_off[i + 1] - rec_off[i]);
    if (len < 32) { c->last_error = "BAM record shorter than its fixed header"; return FGB_ERR_INVALID_ARG; }
    recs.emplace_back(records + rec_off[i], len);
    if (recs.back().aux_off() > len) { c->last_error = "truncated BAM record"; return FGB_ERR_INVALID_ARG; }
  }
  const uint8_t* tv; size_t tn;
  if (!bam::find_string_tag(recs[0], c->opt.tag, &tv, &tn)) {   // vanilla_caller.rs:1493-1496
    c->last_error = std::string("Missing UMI tag '") + c->opt.tag[0] + c->opt.tag[1] + "'";
    return FGB_ERR_MISSING_TAG;
  }
  const std::string umi(reinterpret_cast<const char*>(tv), tn);
  c->stats[FGB_STAT_TOTAL_READS] += n_records;
  std::vector<uint32_t> kept;
  for (uint32_t i = 0; i < n_records; ++i) {
    uint16_t f = recs[i].flags();
    if (!(f & bam::kSecondary) && !(f & bam::kSupplementary)) kept.push_back(i);
  }
  if (kept.size() != n_records) reject(c, FGB_STAT_REJ_SECONDARY_SUPPLEMENTARY, n_records - kept.size());
  if (kept.empty()) return FGB_OK;
  if (kept.size() < c->opt.min_reads) { reject(c, FGB_STAT_REJ_INSUFFICIENT_READS, kept.size()); return FGB_OK; }
  std::vector<uint32_t> frag, r1, r2;
  for (uint32_t i : kept) {   // subgroup_reads, vanilla_caller.rs:1018-1039
    uint16_t f = recs[i].flags();
    if (!(f & bam::kPaired)) frag.push_back(i);
    else if (f & bam::kFirst) r1.push_back(i);
    else if (f & bam::kLast) r2.push_back(i);
  }
  Prepared pf, p1, p2;
  prepare_subgroup(c, recs, frag, &pf);
  if (pf.ok) { pack_unit(c, recs, pf, kFragment, umi); c->stats[FGB_STAT_CONSENSUS_READS] += 1; }
  prepare_subgroup(c, recs, r1, &p1);
  prepare_subgroup(c, recs, r2, &p2);
  if (p1.ok && p2.ok) {
    pack_unit(c, recs, p1, kR1, umi);
    pack_unit(c, recs, p2, kR2, umi);
    c->stats[FGB_STAT_CONSENSUS_READS] += 2;
  } else if (p1.ok) {
    reject(c, FGB_STAT_REJ_ORPHAN_CONSENSUS, p1.surviving);
  } else if (p2.ok) {
    reject(c, FGB_STAT_REJ_ORPHAN_CONSENSUS, p2.surviving);
  }
  return FGB_OK;
}

fgb_status fgb_caller_flush(fgb_caller* c, const uint8_t** out_data, uint64_t* out_len,
                            uint64_t* out_count) {
  if (!c || !out_data || !out_len || !out_count) return FGB_ERR_INVALID_ARG;
  c->out.clear();
  c->out_count = 0;
  const uint64_t U = c->units.size();
  if (U) {
    fgb_unit sentinel;
    sentinel.out_off = c->n_out;
    sentinel.read_begin = static_cast<uint32_t>(c->reads.size());
    sentinel.cons_len = 0;
    c->units.push_back(sentinel);
    const uint64_t n_bytes = c->bases.size();
    c->bases.resize(round_up(n_bytes + 1, 16), 0);
    c->quals.resize(c->bases.size(), 0);
    const uint64_t R = c->reads.size();
    c->reads.resize(R + 2, 0);
    uint64_t n_tiles = 0;
    fgb_status st = fgb_plan_tiles(c->units.data(), U, c->reads.data(), R, nullptr, 0, &n_tiles);
    if (st != FGB_OK) { c->last_error = "fgb_plan_tiles failed"; return st; }
    std::vector<fgb_tile> tiles(n_tiles ? n_tiles : 1);
    st = fgb_plan_tiles(c->units.data(), U, c->reads.data(), R, tiles.data(), n_tiles, &n_tiles);
    if (st != FGB_OK) return st;
    std::vector<uint8_t> ob(c->n_out + 8), oq(c->n_out + 8);
    std::vector<uint16_t> od(c->n_out + 8), oe(c->n_out + 8);
    fgb_batch b;
    b.n_units = U; b.n_reads = R; b.n_bytes = n_bytes; b.n_out = c->n_out; b.n_tiles = n_tiles;
    b.bases = c->bases.data(); b.quals = c->quals.data(); b.reads = c->reads.data();
    b.units = c->units.data(); b.tiles = tiles.data();
    fgb_columns cols{ob.data(), oq.data(), od.data(), oe.data()};
    st = fgb_submit(c->h, &b, &cols);
    if (st == FGB_OK) st = fgb_wait(c->h);
    if (st != FGB_OK) {
      char buf[256];
      fgb_last_error(c->h, buf, sizeof(buf));
      c->last_error = buf;
      return st;
    }
    // ---- build_consensus_record_into, vanilla_caller.rs:1365-1473 ----
    bam::Writer w(&c->out);
    std::string rx;
    for (uint64_t i = 0; i < U; ++i) {
      const fgb_unit& u = c->units[i];
      const UnitMeta& m = c->metas[i];
      const uint32_t L = u.cons_len;
      const uint8_t* bases = ob.data() + u.out_off;
      const uint8_t* quals = oq.data() + u.out_off;
      const uint16_t* depths = od.data() + u.out_off;
      const uint16_t* errors = oe.data() + u.out_off;
      uint16_t flag = bam::kUnmapped;
      if (m.read_type == kR1) flag |= bam::kPaired | bam::kFirst | bam::kMateUnmapped;
      else if (m.read_type == kR2) flag |= bam::kPaired | bam::kLast | bam::kMateUnmapped;
      std::string name = c->prefix + ":" + m.umi;
      if (name.size() >= 255) { c->last_error = "read name too long"; return FGB_ERR_INVALID_ARG; }
      w.begin(name, flag, bases, quals, L);
      w.str("RG", c->rg.data(), c->rg.size());
      uint32_t max_d = 0, min_d = L ? 0xFFFFFFFFu : 0;
      uint64_t tot_e = 0, tot_d = 0;
      for (uint32_t k = 0; k < L; ++k) {
        max_d = std::max<uint32_t>(max_d, depths[k]);
        min_d = std::min<uint32_t>(min_d, depths[k]);
        tot_e += errors[k];
        tot_d += depths[k];
      }
      float rate = tot_d > 0 ? static_cast<float>(tot_e) / static_cast<float>(tot_d) : 0.0f;
      w.integer("cD", static_cast<int32_t>(max_d));
      w.integer("cM", static_cast<int32_t>(min_d));
      w.real("cE", rate);
      if (c->opt.produce_per_base_tags) {
        w.i16_array("cd", depths, L);
        w.i16_array("ce", errors, L);
      }
      w.str("MI", m.umi.data(), m.umi.size());
      if (m.has_cell) w.str(c->opt.cell_tag, m.cell.data(), m.cell.size());
      if (!m.rx.empty()) {
        if (!consensus_umis(c, m.rx, &rx)) {
          c->last_error = "RX values of a family have different lengths or mix DNA and non-DNA characters";
          return FGB_ERR_INVALID_ARG;   // the reference panics here (simple_umi.rs:78-116)
        }
        w.str("RX", rx.data(), rx.size());
      }
      w.end();
      ++c->out_count;
    }
  }
  c->bases.clear(); c->quals.clear(); c->reads.clear(); c->units.clear(); c->metas.clear();
  c->n_out = 0;
  *out_data = c->out.data();
  *out_len = c->out.size();
  *out_count = c->out_count;
  return FGB_OK;
}

fgb_status fgb_caller_stats(const fgb_caller* c, uint64_t stats[FGB_NSTATS]) {
  if (!c || !stats) return FGB_ERR_INVALID_ARG;
  std::memcpy(stats, c->stats, sizeof(c->stats));
  return FGB_OK;
}

}  // extern "C"

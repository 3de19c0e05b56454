// `fgumi filter` applied to one assembled consensus record on the host: the duplex arm
// (crates/fgumi-consensus/src/filter.rs:477-557 filter_duplex_read, :702-806 mask_duplex_bases), the
// single-strand arm for records without aD / bD (filter.rs:453-471, 650-696) and the shared read-level
// gates (src/lib/commands/filter.rs:909-968).  The simplex caller runs its filter on the device columns
// (filter_kernel.cuh); duplex reads carry three depth/error tiers and optional strand-agreement
// strings, which exist only in the assembled record -- so this arm reads the record like the
// reference does.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cstring>

#include "../../../include/fgumi_b200.h"
#include "bam.h"

namespace fgb {
namespace rfilter {

struct TagRef {
  uint8_t type = 0;        // 0 = absent; 'B' arrays carry `sub` and `count`
  uint8_t sub = 0;
  uint32_t count = 0;
  const uint8_t* data = nullptr;
};

inline size_t fixed_size(uint8_t t) {
  switch (t) {
    case 'A': case 'c': case 'C': return 1;
    case 's': case 'S': return 2;
    case 'i': case 'I': case 'f': return 4;
    default: return 0;
  }
}

// bam_fields::find_tag_position semantics: first occurrence, stops at malformed data.
inline TagRef find_tag(const uint8_t* a, size_t an, const char tag[2]) {
  TagRef r;
  size_t p = 0;
  while (p + 3 <= an) {
    const uint8_t vt = a[p + 2];
    size_t size = fixed_size(vt);
    uint8_t sub = 0;
    uint32_t count = 0;
    if (!size) {
      if (vt == 'Z' || vt == 'H') {
        const void* z = std::memchr(a + p + 3, 0, an - (p + 3));
        if (!z) return r;
        size = static_cast<const uint8_t*>(z) - (a + p + 3) + 1;
      } else if (vt == 'B') {
        if (an - (p + 3) < 5) return r;
        sub = a[p + 3];
        const size_t es = fixed_size(sub);
        if (!es) return r;
        count = bam::rd32(a + p + 4);
        size = 5 + static_cast<size_t>(count) * es;
      } else {
        return r;
      }
    }
    if (p + 3 + size > an) return r;
    if (a[p] == static_cast<uint8_t>(tag[0]) && a[p + 1] == static_cast<uint8_t>(tag[1])) {
      r.type = vt; r.sub = sub; r.count = vt == 'B' ? count : static_cast<uint32_t>(size);
      r.data = a + p + 3 + (vt == 'B' ? 5 : 0);
      return r;
    }
    p += 3 + size;
  }
  return r;
}

inline bool tag_int(const TagRef& t, int64_t* v) {            // raw-bam tags.rs:118-150: any integer type
  switch (t.type) {
    case 'c': *v = static_cast<int8_t>(t.data[0]); return true;
    case 'C': *v = t.data[0]; return true;
    case 's': *v = static_cast<int16_t>(bam::rd16(t.data)); return true;
    case 'S': *v = bam::rd16(t.data); return true;
    case 'i': *v = bam::rdi32(t.data); return true;
    case 'I': *v = bam::rd32(t.data); return true;
    default: return false;
  }
}
inline bool tag_float(const TagRef& t, float* v) {            // 'f' only
  if (t.type != 'f') return false;
  std::memcpy(v, t.data, 4);
  return true;
}
inline uint16_t array_u16(const TagRef& t, size_t i) {        // raw-bam tags.rs:479-497
  if (t.type != 'B' || i >= t.count) return 0;
  switch (t.sub) {
    case 'C': return t.data[i];
    case 'S': return bam::rd16(t.data + 2 * i);
    case 's': { int16_t v = static_cast<int16_t>(bam::rd16(t.data + 2 * i)); return v > 0 ? static_cast<uint16_t>(v) : 0; }
    case 'c': { int8_t v = static_cast<int8_t>(t.data[i]); return v > 0 ? static_cast<uint16_t>(v) : 0; }
    default: return 0;
  }
}

struct Tier {               // FilterThresholds, filter.rs:31-40
  uint64_t min_reads;
  double max_read_error_rate, max_base_error_rate;
};

inline int filter_read(const uint8_t* aux, size_t an, const Tier& th) {          // filter.rs:453-471
  int64_t depth;
  if (tag_int(find_tag(aux, an, "cD"), &depth)) {
    const int64_t mr = th.min_reads > static_cast<uint64_t>(INT64_MAX) ? INT64_MAX : static_cast<int64_t>(th.min_reads);
    if (depth < mr) return FGB_FILTER_INSUFFICIENT_READS;
  }
  float er;
  if (tag_float(find_tag(aux, an, "cE"), &er) && static_cast<double>(er) > th.max_read_error_rate)
    return FGB_FILTER_EXCESSIVE_ERROR_RATE;
  return FGB_FILTER_PASS;
}

inline int filter_duplex_read(const uint8_t* aux, size_t an, const Tier& cc, const Tier& ab, const Tier& ba) {
  int r = filter_read(aux, an, cc);                                                // filter.rs:477-557
  if (r != FGB_FILTER_PASS) return r;
  int64_t a_d = 0, b_d = 0;
  bool has_a = tag_int(find_tag(aux, an, "aD"), &a_d) || tag_int(find_tag(aux, an, "aM"), &a_d);
  bool has_b = tag_int(find_tag(aux, an, "bD"), &b_d) || tag_int(find_tag(aux, an, "bM"), &b_d);
  float a_e = 0, b_e = 0;
  const bool has_ae = tag_float(find_tag(aux, an, "aE"), &a_e), has_be = tag_float(find_tag(aux, an, "bE"), &b_e);
  int64_t worst_d, best_d;
  if (has_a && has_b) { if (a_d < b_d) { worst_d = a_d; best_d = b_d; } else { worst_d = b_d; best_d = a_d; } }
  else if (has_a) { worst_d = 0; best_d = a_d; }
  else if (has_b) { worst_d = 0; best_d = b_d; }
  else return FGB_FILTER_PASS;
  float best_e, worst_e;
  if (has_ae && has_be) { if (a_e < b_e) { best_e = a_e; worst_e = b_e; } else { best_e = b_e; worst_e = a_e; } }
  else if (has_ae) best_e = worst_e = a_e;
  else if (has_be) best_e = worst_e = b_e;
  else best_e = worst_e = 0.0f;
  // `(depth as usize) < min_reads`: a negative depth wraps to a huge usize and passes
  if (static_cast<uint64_t>(best_d) < ab.min_reads) return FGB_FILTER_INSUFFICIENT_READS;
  if (static_cast<double>(best_e) > ab.max_read_error_rate) return FGB_FILTER_EXCESSIVE_ERROR_RATE;
  if (static_cast<uint64_t>(worst_d) < ba.min_reads) return FGB_FILTER_INSUFFICIENT_READS;
  if (static_cast<double>(worst_e) > ba.max_read_error_rate) return FGB_FILTER_EXCESSIVE_ERROR_RATE;
  return FGB_FILTER_PASS;
}

inline uint8_t get_base(const uint8_t* rec, size_t so, size_t i) {
  static const char kCodes[] = "=ACMGRSVTWYHKDBN";
  const uint8_t b = rec[so + i / 2];
  return static_cast<uint8_t>(kCodes[(i & 1) ? (b & 0xF) : (b >> 4)]);
}
inline void mask_base(uint8_t* rec, size_t so, size_t i) {     // nibble 15 = N
  uint8_t& b = rec[so + i / 2];
  b = (i & 1) ? static_cast<uint8_t>(b | 0x0F) : static_cast<uint8_t>(b | 0xF0);
}

// Z string or B:C / B:c array (filter.rs:621-639); absent / other types -> nullptr
inline bool string_or_u8(const TagRef& t, const uint8_t** p, size_t* n) {
  if (t.type == 'Z') { *p = t.data; *n = t.count ? t.count - 1 : 0; return true; }
  if (t.type == 'B' && (t.sub == 'C' || t.sub == 'c')) { *p = t.data; *n = t.count; return true; }
  return false;
}

inline uint32_t mask_bases(uint8_t* rec, size_t n, const Tier& th, int min_bq) {   // filter.rs:650-696
  const bam::View v(rec, n);
  const size_t so = v.seq_off(), qo = v.qual_off(), L = v.l_seq(), ao = v.aux_off();
  const uint8_t* aux = rec + ao;
  const size_t an = n - ao;
  const TagRef cd = find_tag(aux, an, "cd"), ce = find_tag(aux, an, "ce");
  uint32_t masked = 0;
  for (size_t i = 0; i < L; ++i) {
    const uint16_t depth = array_u16(cd, i), errors = array_u16(ce, i);
    const uint8_t q = rec[qo + i];
    const bool should = (min_bq >= 0 && q < min_bq) || depth < th.min_reads ||
                        (depth > 0 && static_cast<double>(errors) / static_cast<double>(depth) > th.max_base_error_rate);
    if (should) {
      if (get_base(rec, so, i) != 'N') ++masked;
      mask_base(rec, so, i);
      rec[qo + i] = 2;
    }
  }
  return masked;
}

inline uint32_t mask_duplex_bases(uint8_t* rec, size_t n, const Tier& cc, const Tier& ab, const Tier& ba,
                                  int min_bq, bool require_ss_agreement) {          // filter.rs:702-806
  const bam::View v(rec, n);
  const size_t so = v.seq_off(), qo = v.qual_off(), L = v.l_seq(), ao = v.aux_off();
  const uint8_t* aux = rec + ao;
  const size_t an = n - ao;
  const TagRef ad = find_tag(aux, an, "ad"), ae = find_tag(aux, an, "ae");
  const TagRef bd = find_tag(aux, an, "bd"), be = find_tag(aux, an, "be");
  const uint8_t *acp = nullptr, *bcp = nullptr;
  size_t acn = 0, bcn = 0;
  bool has_ac = false, has_bc = false;
  if (require_ss_agreement) {
    has_ac = string_or_u8(find_tag(aux, an, "ac"), &acp, &acn);
    has_bc = string_or_u8(find_tag(aux, an, "bc"), &bcp, &bcn);
  }
  uint32_t masked = 0;
  for (size_t i = 0; i < L; ++i) {
    if (get_base(rec, so, i) == 'N') continue;
    const uint16_t a_d = array_u16(ad, i), b_d = array_u16(bd, i), a_e = array_u16(ae, i), b_e = array_u16(be, i);
    const uint16_t best_d = std::max(a_d, b_d), worst_d = std::min(a_d, b_d);
    const double a_r = a_d > 0 ? static_cast<double>(a_e) / static_cast<double>(a_d) : 0.0;
    const double b_r = b_d > 0 ? static_cast<double>(b_e) / static_cast<double>(b_d) : 0.0;
    const double best_r = std::min(a_r, b_r), worst_r = std::max(a_r, b_r);
    const uint32_t tot_d = static_cast<uint32_t>(a_d) + b_d;
    const double tot_r = tot_d > 0 ? static_cast<double>(static_cast<uint32_t>(a_e) + b_e) / static_cast<double>(tot_d) : 0.0;
    const uint8_t q = rec[qo + i];
    const bool should = (min_bq >= 0 && q < min_bq) || tot_d < cc.min_reads || tot_r > cc.max_base_error_rate ||
                        best_d < ab.min_reads || best_r > ab.max_base_error_rate ||
                        worst_d < ba.min_reads || worst_r > ba.max_base_error_rate;
    bool ss_dis = false;
    if (require_ss_agreement && a_d > 0 && b_d > 0) {
      const uint8_t x = (has_ac && i < acn) ? acp[i] : 'N', y = (has_bc && i < bcn) ? bcp[i] : 'N';
      ss_dis = x != y;
    }
    if (should || ss_dis) {
      ++masked;
      mask_base(rec, so, i);
      rec[qo + i] = 2;
    }
  }
  return masked;
}

// compute_read_stats + check_no_call_and_quality (filter.rs:565-590, commands/filter.rs:909-929)
inline int no_call_and_quality(const uint8_t* rec, size_t n, double min_mean_q, double max_no_call) {
  const bam::View v(rec, n);
  const size_t so = v.seq_off(), qo = v.qual_off(), L = v.l_seq();
  uint64_t n_count = 0, qsum = 0;
  for (size_t i = 0; i < L; ++i) {
    if (get_base(rec, so, i) == 'N') ++n_count; else qsum += rec[qo + i];
  }
  const uint64_t non_n = L - n_count;
  const double mean_q = non_n ? static_cast<double>(qsum) / static_cast<double>(non_n) : 0.0;
  if (min_mean_q >= 0.0 && mean_q < min_mean_q) return FGB_FILTER_LOW_MEAN_QUALITY;
  if (max_no_call >= 1.0) {
    if (static_cast<double>(n_count) > max_no_call) return FGB_FILTER_TOO_MANY_NO_CALLS;
  } else {
    const double frac = L ? static_cast<double>(n_count) / static_cast<double>(L) : 0.0;
    if (frac > max_no_call) return FGB_FILTER_TOO_MANY_NO_CALLS;
  }
  return FGB_FILTER_PASS;
}

// One record through the filter (commands/filter.rs:770-793 masking, :949-968 read-level gates).
inline int filter_record(uint8_t* rec, size_t n, const fgb_duplex_filter_params& p, uint32_t* masked) {
  const Tier cc{p.cc.min_reads, p.cc.max_read_error_rate, p.cc.max_base_error_rate};
  const Tier ab{p.ab_min_reads, p.ab_max_read_error_rate, p.ab_max_base_error_rate};
  const Tier ba{p.ba_min_reads, p.ba_max_read_error_rate, p.ba_max_base_error_rate};
  const bam::View v(rec, n);
  const size_t ao = v.aux_off();
  const bool duplex = find_tag(rec + ao, n - ao, "aD").type || find_tag(rec + ao, n - ao, "bD").type;
  int st;
  if (duplex) {
    *masked = mask_duplex_bases(rec, n, cc, ab, ba, p.cc.min_base_quality, p.require_ss_agreement != 0);
    st = filter_duplex_read(rec + ao, n - ao, cc, ab, ba);
  } else {
    *masked = mask_bases(rec, n, cc, p.cc.min_base_quality);
    st = filter_read(rec + ao, n - ao, cc);
  }
  if (st != FGB_FILTER_PASS) return st;
  return no_call_and_quality(rec, n, p.cc.min_mean_base_quality, p.cc.max_no_call_fraction);
}

}  // namespace rfilter
}  // namespace fgb

// K2 (duplex) and K3 (CODEC) strand-combine kernels for sm_100a — pure integer/byte work.
//
// Replaces (reference = /root/reference/crates/fgumi-consensus/src/):
//   duplex_caller.rs:838-1015   DuplexConsensusCaller::duplex_consensus (methylation disabled)
//   codec_caller.rs:507-520     reverse_complement_ss
//   codec_caller.rs:980-1023    pad_consensus (lowercase 'n', Q0, depth 0, errors 0)
//   codec_caller.rs:1029-1178   build_duplex_consensus_from_padded (+ disagreement gate)
//   codec_caller.rs:1183-1212   mask_consensus_quals_query_based
//   fgumi-dna/src/dna.rs:30-40  complement_base
//
// One warp per job, lanes stride the position axis (coalesced byte/u16 loads and stores); the
// job-level reductions (any-depth, disagreement counts) are warp shuffles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fgumi_b200.h"

namespace fgb {

constexpr int kCombineThreads = 256;
constexpr int kCombineJobsPerCta = kCombineThreads / 32;
constexpr int kCodecJobsPerCta = kCombineThreads / 32;

struct DuplexArgs {
  const uint8_t* bases;        // source base column (for the exact error recount)
  const uint64_t* reads;
  const fgb_unit* units;
  const uint8_t* ss_base;
  const uint8_t* ss_qual;
  const uint16_t* ss_depth;
  const uint16_t* ss_errors;
  const fgb_duplex_job* jobs;
  uint64_t n_jobs;
  uint8_t* out_base;
  uint8_t* out_qual;
  uint16_t* out_errors;
  uint8_t* out_status;
  unsigned long long* counters;
};

// duplex_caller.rs:783-791
__device__ __forceinline__ uint32_t cap_quality(int32_t s) {
  return s < 2 ? 2u : (s > 93 ? 93u : static_cast<uint32_t>(s));
}

// One job's descriptors, fetched one job ahead of the combine (the loads of job j + stride are in flight while
// job j is combined: two dependent rounds of global-memory latency leave the critical path).
struct DuplexJobRegs {
  fgb_duplex_job job;
  fgb_unit ua, ub;
  uint32_t ra1, rb1;      // read_begin of the units after a and b
};
__device__ __forceinline__ DuplexJobRegs load_duplex_job(const DuplexArgs& a, uint64_t j) {
  DuplexJobRegs r;
  r.job = a.jobs[j];
  r.ua = a.units[r.job.unit_a];
  r.ub = a.units[r.job.unit_b];
  r.ra1 = a.units[r.job.unit_a + 1].read_begin;
  r.rb1 = a.units[r.job.unit_b + 1].read_begin;
  return r;
}

__global__ void __launch_bounds__(kCombineThreads) duplex_combine_kernel(const DuplexArgs a) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warp0 = static_cast<uint64_t>(blockIdx.x) * kCombineJobsPerCta + (threadIdx.x >> 5);
  const uint64_t wstride = static_cast<uint64_t>(gridDim.x) * kCombineJobsPerCta;
  uint32_t done = 0;
  DuplexJobRegs nx;
  if (warp0 < a.n_jobs) nx = load_duplex_job(a, warp0);
  for (uint64_t j = warp0; j < a.n_jobs; j += wstride) {
    const DuplexJobRegs cur = nx;
    if (j + wstride < a.n_jobs) nx = load_duplex_job(a, j + wstride);
    const fgb_duplex_job job = cur.job;
    const fgb_unit ua = cur.ua, ub = cur.ub;
    const uint32_t la = ua.cons_len, lb = ub.cons_len;
    const uint32_t len = la < lb ? la : lb;                       // duplex_caller.rs:846-849
    const uint32_t ra0 = ua.read_begin, ra1 = cur.ra1;
    const uint32_t rb0 = ub.read_begin, rb1 = cur.rb1;
    // Word path: 8 positions per lane with byte-parallel arithmetic.  Needs 8-aligned rows (the layout rule
    // for inputs; job.out_off is the caller's) and per-position error counts that fit a byte.
    const bool words = ((ua.out_off | ub.out_off | job.out_off) & 7u) == 0 && (ra1 - ra0) + (rb1 - rb0) <= 255u;
    uint8_t status;
    if (words && len <= 256u) {
      // ---- one block per job (reads up to 256 bases): everything the job needs is requested at once ----
      const uint32_t p0 = lane * 8u;
      const bool active = p0 < len;
      uint2 ab2 = make_uint2(0, 0), bb2 = ab2, aq2 = ab2, bq2 = ab2;
      uint4 ad4 = make_uint4(0, 0, 0, 0), bd4 = ad4;
      if (active) {
        ab2 = *reinterpret_cast<const uint2*>(a.ss_base + ua.out_off + p0);
        bb2 = *reinterpret_cast<const uint2*>(a.ss_base + ub.out_off + p0);
        aq2 = *reinterpret_cast<const uint2*>(a.ss_qual + ua.out_off + p0);
        bq2 = *reinterpret_cast<const uint2*>(a.ss_qual + ub.out_off + p0);
        ad4 = *reinterpret_cast<const uint4*>(a.ss_depth + ua.out_off + p0);
        bd4 = *reinterpret_cast<const uint4*>(a.ss_depth + ub.out_off + p0);
      }
      const uint32_t na = ra1 - ra0, nr = na + (rb1 - rb0);
      // descriptors of the pooled source rows (AB rows then BA rows), one per lane, 32 at a time
      const uint64_t first_desc = lane < nr ? a.reads[lane < na ? ra0 + lane : rb0 + (lane - na)] : 0ull;
      // :852-853 strands with no coverage inside the truncated region are dropped (rows are padded with
      // zero depth, but a longer strand has real depths behind `len`: mask the last word)
      const uint32_t live = active ? (len - p0 < 8u ? len - p0 : 8u) : 0u;
      auto any16 = [&](const uint4& d) {
        const uint32_t w[4] = {d.x, d.y, d.z, d.w};
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t keep = live >= 2u * k + 2u ? 0xFFFFFFFFu : (live == 2u * k + 1u ? 0x0000FFFFu : 0u);
          acc |= w[k] & keep;
        }
        return acc != 0u;
      };
      const bool a_any = __any_sync(0xFFFFFFFFu, any16(ad4));
      const bool b_any = __any_sync(0xFFFFFFFFu, any16(bd4));
      if (a_any && b_any) {
        status = FGB_DUPLEX_BOTH;
        uint32_t ob[2], oq[2], rawb[2], cnt[2] = {0u, 0u};
        const uint32_t abw[2] = {ab2.x, ab2.y}, bbw[2] = {bb2.x, bb2.y};
        const uint32_t aqw[2] = {aq2.x, aq2.y}, bqw[2] = {bq2.x, bq2.y};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t eq = __vcmpeq4(abw[h], bbw[h]);                    // :912-927, bytewise
          const uint32_t sum = __vminu4(__vaddus4(aqw[h], bqw[h]), 0x5D5D5D5Du);
          const uint32_t dif = __vminu4(__vabsdiffu4(aqw[h], bqw[h]), 0x5D5D5D5Du);
          const uint32_t rq = __vmaxu4((eq & sum) | (~eq & dif), 0x02020202u);   // cap_quality; equal-quality dissent -> 2
          const uint32_t b_wins = ~eq & __vcmpgtu4(bqw[h], aqw[h]);
          rawb[h] = (bbw[h] & b_wins) | (abw[h] & ~b_wins);
          const uint32_t mask = __vcmpeq4(abw[h], 0x4E4E4E4Eu) | __vcmpeq4(bbw[h], 0x4E4E4E4Eu) |
                                __vcmpeq4(rq, 0x02020202u);                 // :930-935
          ob[h] = (0x4E4E4E4Eu & mask) | (rawb[h] & ~mask);
          oq[h] = (0x02020202u & mask) | (rq & ~mask);
        }
        // :943-951 exact error recount against the pooled source reads
        auto recount_word = [&](uint64_t d) {
          const uint32_t rl = static_cast<uint32_t>(d & 0xFFFFu);
          if (active && rl > p0) {
            const uint2 sb = *reinterpret_cast<const uint2*>(a.bases + (d >> 16) + p0);
            const uint32_t cov = rl - p0;                                   // covered positions of this word
            const uint32_t c0 = cov >= 4u ? 0xFFFFFFFFu : ((1u << (8u * cov)) - 1u);
            const uint32_t c1 = cov >= 8u ? 0xFFFFFFFFu : (cov > 4u ? ((1u << (8u * (cov - 4u))) - 1u) : 0u);
            const uint32_t sw[2] = {sb.x, sb.y}, cw[2] = {c0, c1};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint32_t ne = ~__vcmpeq4(sw[h], rawb[h]) & ~__vcmpeq4(sw[h], 0x4E4E4E4Eu) & cw[h];
              cnt[h] += ne & 0x01010101u;
            }
          }
        };
        for (uint32_t c0 = 0; c0 < nr; c0 += 32u) {
          const uint32_t k = c0 + lane;
          const uint64_t mine = c0 == 0 ? first_desc : (k < nr ? a.reads[k < na ? ra0 + k : rb0 + (k - na)] : 0ull);
          const uint32_t m = nr - c0 < 32u ? nr - c0 : 32u;
#pragma unroll 8
          for (uint32_t r = 0; r < m; ++r) recount_word(__shfl_sync(0xFFFFFFFFu, mine, r));
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) cnt[h] &= ~__vcmpeq4(rawb[h], 0x4E4E4E4Eu);   // raw base N: no recount
        if (active) {
          *reinterpret_cast<uint2*>(a.out_base + job.out_off + p0) = make_uint2(ob[0], ob[1]);
          *reinterpret_cast<uint2*>(a.out_qual + job.out_off + p0) = make_uint2(oq[0], oq[1]);
          *reinterpret_cast<uint4*>(a.out_errors + job.out_off + p0) =
              make_uint4(__byte_perm(cnt[0], 0u, 0x4140u), __byte_perm(cnt[0], 0u, 0x4342u),
                         __byte_perm(cnt[1], 0u, 0x4140u), __byte_perm(cnt[1], 0u, 0x4342u));
        }
      } else if (a_any || b_any) {
        // :855-882 single-strand passthrough keeps the FULL length of the surviving strand
        status = a_any ? FGB_DUPLEX_A_ONLY : FGB_DUPLEX_B_ONLY;
        const fgb_unit us = a_any ? ua : ub;
        for (uint32_t i = lane; i < us.cons_len; i += 32) {
          a.out_base[job.out_off + i] = a.ss_base[us.out_off + i];
          a.out_qual[job.out_off + i] = a.ss_qual[us.out_off + i];
          a.out_errors[job.out_off + i] = a.ss_errors[us.out_off + i];
        }
      } else {
        status = FGB_DUPLEX_NONE;
      }
      if (lane == 0) {
        if (a.out_status) a.out_status[j] = status;
        ++done;
      }
      continue;
    }
    // ---- general path (long reads, unaligned rows, more than 255 pooled source reads) ----
    // :852-853 strands with no coverage inside the truncated region are dropped
    bool a_any = false, b_any = false;
    for (uint32_t i = lane; i < len; i += 32) {
      a_any |= a.ss_depth[ua.out_off + i] > 0;
      b_any |= a.ss_depth[ub.out_off + i] > 0;
    }
    a_any = __any_sync(0xFFFFFFFFu, a_any);
    b_any = __any_sync(0xFFFFFFFFu, b_any);
    if (a_any && b_any) {
      status = FGB_DUPLEX_BOTH;
      if (words) {
        for (uint32_t base0 = 0; base0 < len; base0 += 256u) {
          const uint32_t p0 = base0 + lane * 8u;
          const bool active = p0 < len;
          uint2 ab2 = make_uint2(0, 0), bb2 = ab2, aq2 = ab2, bq2 = ab2;
          if (active) {
            ab2 = *reinterpret_cast<const uint2*>(a.ss_base + ua.out_off + p0);
            bb2 = *reinterpret_cast<const uint2*>(a.ss_base + ub.out_off + p0);
            aq2 = *reinterpret_cast<const uint2*>(a.ss_qual + ua.out_off + p0);
            bq2 = *reinterpret_cast<const uint2*>(a.ss_qual + ub.out_off + p0);
          }
          uint32_t ob[2], oq[2], rawb[2], cnt[2] = {0u, 0u};
          const uint32_t abw[2] = {ab2.x, ab2.y}, bbw[2] = {bb2.x, bb2.y};
          const uint32_t aqw[2] = {aq2.x, aq2.y}, bqw[2] = {bq2.x, bq2.y};
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t eq = __vcmpeq4(abw[h], bbw[h]);                    // :912-927, bytewise
            const uint32_t sum = __vminu4(__vaddus4(aqw[h], bqw[h]), 0x5D5D5D5Du);
            const uint32_t dif = __vminu4(__vabsdiffu4(aqw[h], bqw[h]), 0x5D5D5D5Du);
            const uint32_t rq = __vmaxu4((eq & sum) | (~eq & dif), 0x02020202u);   // cap_quality; equal-quality dissent -> 2
            const uint32_t b_wins = ~eq & __vcmpgtu4(bqw[h], aqw[h]);
            rawb[h] = (bbw[h] & b_wins) | (abw[h] & ~b_wins);
            const uint32_t mask = __vcmpeq4(abw[h], 0x4E4E4E4Eu) | __vcmpeq4(bbw[h], 0x4E4E4E4Eu) |
                                  __vcmpeq4(rq, 0x02020202u);                 // :930-935
            ob[h] = (0x4E4E4E4Eu & mask) | (rawb[h] & ~mask);
            oq[h] = (0x02020202u & mask) | (rq & ~mask);
          }
          // :943-951 exact error recount against the pooled source reads (AB rows then BA rows)
          // descriptors are fetched 32 at a time, one per lane, and broadcast by shuffle, so the row
          // loads of a chunk are independent of each other (all lanes take part in the shuffles)
          auto recount_word = [&](uint64_t d) {
            const uint32_t rl = static_cast<uint32_t>(d & 0xFFFFu);
            if (active && rl > p0) {
              const uint2 sb = *reinterpret_cast<const uint2*>(a.bases + (d >> 16) + p0);
              const uint32_t cov = rl - p0;                                   // covered positions of this word
              const uint32_t c0 = cov >= 4u ? 0xFFFFFFFFu : ((1u << (8u * cov)) - 1u);
              const uint32_t c1 = cov >= 8u ? 0xFFFFFFFFu : (cov > 4u ? ((1u << (8u * (cov - 4u))) - 1u) : 0u);
              const uint32_t sw[2] = {sb.x, sb.y}, cw[2] = {c0, c1};
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint32_t ne = ~__vcmpeq4(sw[h], rawb[h]) & ~__vcmpeq4(sw[h], 0x4E4E4E4Eu) & cw[h];
                cnt[h] += ne & 0x01010101u;
              }
            }
          };
          const uint32_t na = ra1 - ra0, nr = na + (rb1 - rb0);
          for (uint32_t c0 = 0; c0 < nr; c0 += 32u) {
            const uint32_t k = c0 + lane;
            const uint64_t mine = k < nr ? a.reads[k < na ? ra0 + k : rb0 + (k - na)] : 0ull;
            const uint32_t m = nr - c0 < 32u ? nr - c0 : 32u;
#pragma unroll 4
            for (uint32_t r = 0; r < m; ++r) recount_word(__shfl_sync(0xFFFFFFFFu, mine, r));
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) cnt[h] &= ~__vcmpeq4(rawb[h], 0x4E4E4E4Eu);   // raw base N: no recount
          if (active) {
            *reinterpret_cast<uint2*>(a.out_base + job.out_off + p0) = make_uint2(ob[0], ob[1]);
            *reinterpret_cast<uint2*>(a.out_qual + job.out_off + p0) = make_uint2(oq[0], oq[1]);
            *reinterpret_cast<uint4*>(a.out_errors + job.out_off + p0) =
                make_uint4(__byte_perm(cnt[0], 0u, 0x4140u), __byte_perm(cnt[0], 0u, 0x4342u),
                           __byte_perm(cnt[1], 0u, 0x4140u), __byte_perm(cnt[1], 0u, 0x4342u));
          }
        }
      } else {
      for (uint32_t i = lane; i < len; i += 32) {
        uint32_t a_base = a.ss_base[ua.out_off + i], b_base = a.ss_base[ub.out_off + i];
        int32_t a_qual = a.ss_qual[ua.out_off + i], b_qual = a.ss_qual[ub.out_off + i];
        uint32_t raw_base, raw_qual;                               // :912-927
        if (a_base == b_base) { raw_base = a_base; raw_qual = cap_quality(a_qual + b_qual); }
        else if (a_qual > b_qual) { raw_base = a_base; raw_qual = cap_quality(a_qual - b_qual); }
        else if (b_qual > a_qual) { raw_base = b_base; raw_qual = cap_quality(b_qual - a_qual); }
        else { raw_base = a_base; raw_qual = 2u; }
        bool mask = a_base == 'N' || b_base == 'N' || raw_qual == 2u;   // :930-935
        // :943-951 exact error recount against the pooled source reads (AB rows then BA rows)
        int32_t nerr = 0;
        if (raw_base != 'N') {
          for (uint32_t r = ra0; r < ra1; ++r) {
            uint64_t d = a.reads[r];
            if ((d & 0xFFFFu) > i) {
              uint32_t sb = a.bases[(d >> 16) + i];
              nerr += (sb != 'N' && sb != raw_base);
            }
          }
          for (uint32_t r = rb0; r < rb1; ++r) {
            uint64_t d = a.reads[r];
            if ((d & 0xFFFFu) > i) {
              uint32_t sb = a.bases[(d >> 16) + i];
              nerr += (sb != 'N' && sb != raw_base);
            }
          }
        }
        nerr = nerr > 32767 ? 32767 : nerr;
        a.out_base[job.out_off + i] = mask ? 'N' : static_cast<uint8_t>(raw_base);
        a.out_qual[job.out_off + i] = mask ? 2 : static_cast<uint8_t>(raw_qual);
        a.out_errors[job.out_off + i] = static_cast<uint16_t>(nerr);
      }
      }
    } else if (a_any || b_any) {
      // :855-882 single-strand passthrough keeps the FULL length of the surviving strand
      status = a_any ? FGB_DUPLEX_A_ONLY : FGB_DUPLEX_B_ONLY;
      const fgb_unit us = a_any ? ua : ub;
      for (uint32_t i = lane; i < us.cons_len; i += 32) {
        a.out_base[job.out_off + i] = a.ss_base[us.out_off + i];
        a.out_qual[job.out_off + i] = a.ss_qual[us.out_off + i];
        a.out_errors[job.out_off + i] = a.ss_errors[us.out_off + i];
      }
    } else {
      status = FGB_DUPLEX_NONE;
    }
    if (lane == 0) {
      if (a.out_status) a.out_status[j] = status;
      ++done;
    }
  }
  if (lane == 0 && done) atomicAdd(a.counters + FGB_CTR_COMBINED, static_cast<unsigned long long>(done));
}

// ---- CODEC ------------------------------------------------------------------------------------
struct CodecArgs {
  const fgb_unit* units;
  const uint8_t* ss_base;
  const uint8_t* ss_qual;
  const uint16_t* ss_depth;
  const uint16_t* ss_errors;
  const fgb_codec_job* jobs;
  uint64_t n_jobs;
  fgb_codec_params cp;
  uint8_t* out_base;
  uint8_t* out_qual;
  uint16_t* out_depth;
  uint16_t* out_errors;
  uint8_t* status;
  uint32_t* disagreements;
  uint32_t* duplex_bases;
  unsigned long long* counters;
};

// fgumi-dna dna.rs:30-40
__device__ __forceinline__ uint32_t complement_base(uint32_t b) {
  switch (b) {
    case 'A': case 'a': return 'T';
    case 'T': case 't': return 'A';
    case 'C': case 'c': return 'G';
    case 'G': case 'g': return 'C';
    default: return b;   // 'N' stays 'N', 'n' stays 'n', anything else unchanged
  }
}

struct SsCol {
  uint32_t base, qual, depth, err;
};

// Column `i` of a single-strand consensus after orientation (reverse_complement_ss) and padding.
__device__ __forceinline__ SsCol padded_column(const CodecArgs& a, const fgb_unit& un, uint32_t i,
                                               uint32_t pad_left, bool rc) {
  SsCol c;
  uint32_t p = i - pad_left;                 // wraps to a huge value when i < pad_left
  if (i >= pad_left && p < un.cons_len) {
    uint32_t s = rc ? un.cons_len - 1u - p : p;
    uint32_t b = a.ss_base[un.out_off + s];
    c.base = rc ? complement_base(b) : b;
    c.qual = a.ss_qual[un.out_off + s];
    c.depth = a.ss_depth[un.out_off + s];
    c.err = a.ss_errors[un.out_off + s];
  } else {                                   // pad_consensus, codec_caller.rs:991-995
    c.base = 'n'; c.qual = 0; c.depth = 0; c.err = 0;
  }
  return c;
}

__global__ void __launch_bounds__(kCombineThreads) codec_combine_kernel(const CodecArgs a) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warp0 = static_cast<uint64_t>(blockIdx.x) * kCodecJobsPerCta + (threadIdx.x >> 5);
  const uint64_t wstride = static_cast<uint64_t>(gridDim.x) * kCodecJobsPerCta;
  unsigned long long tot_bases = 0, tot_dis = 0;
  uint32_t done = 0;
  for (uint64_t j = warp0; j < a.n_jobs; j += wstride) {
    const fgb_codec_job job = a.jobs[j];
    const fgb_unit ua = a.units[job.unit_a], ub = a.units[job.unit_b];
    const uint32_t len = job.len;
    const uint32_t outer_len = a.cp.outer_bases_length;
    const uint32_t outer_hi = len > outer_len ? len - outer_len : 0u;   // saturating_sub, :1205
    uint32_t n_dup = 0, n_dis = 0;
    for (uint32_t i = lane; i < len; i += 32) {
      SsCol A = padded_column(a, ua, i, job.pad_a_left, job.rc_a != 0);
      SsCol B = padded_column(a, ub, i, job.pad_b_left, job.rc_b != 0);
      const bool a_has = A.base != 'N' && A.base != 'n';           // :1064-1065
      const bool b_has = B.base != 'N' && B.base != 'n';
      uint32_t dbase, dqual, depth, err;
      if (a_has && b_has) {                                        // :1068-1113
        ++n_dup;
        uint32_t raw_base, raw_qual;
        if (A.base == B.base) {
          raw_base = A.base;
          uint32_t s = A.qual + B.qual;
          raw_qual = s > 93u ? 93u : s;
        } else if (A.qual > B.qual) {
          ++n_dis; raw_base = A.base;
          uint32_t d = A.qual - B.qual; raw_qual = d < 2u ? 2u : d;
        } else if (B.qual > A.qual) {
          ++n_dis; raw_base = B.base;
          uint32_t d = B.qual - A.qual; raw_qual = d < 2u ? 2u : d;
        } else {
          ++n_dis; raw_base = A.base; raw_qual = 2u;
        }
        if (raw_qual == 2u) { dbase = 'N'; dqual = 2u; } else { dbase = raw_base; dqual = raw_qual; }
        if (A.base == B.base) err = A.err + B.err;
        else if (A.base == raw_base) err = A.err + (B.depth > B.err ? B.depth - B.err : 0u);
        else err = B.err + (A.depth > A.err ? A.depth - A.err : 0u);
        depth = A.depth + B.depth;
      } else if (a_has) {                                          // :1115-1122
        if (A.qual == 2u) { dbase = 'N'; dqual = 2u; } else { dbase = A.base; dqual = A.qual; }
        depth = A.depth; err = A.err;
      } else if (b_has) {                                          // :1124-1131
        if (B.qual == 2u) { dbase = 'N'; dqual = 2u; } else { dbase = B.base; dqual = B.qual; }
        depth = B.depth; err = B.err;
      } else {                                                     // :1133-1139
        dbase = 'N'; dqual = 2u; depth = 0; err = A.err + B.err;
      }
      if (A.base == 'N' || B.base == 'N') { dbase = 'N'; dqual = 2u; }   // :1145-1149
      // mask_consensus_quals_query_based, :1191-1209
      if ((A.base == 'N' || B.base == 'N') && dbase != 'N') {
        if (a.cp.single_strand_qual >= 0) dqual = static_cast<uint32_t>(a.cp.single_strand_qual);
      }
      if (a.cp.outer_bases_qual >= 0) {
        if (i < outer_len || i >= outer_hi) {
          uint32_t oq = static_cast<uint32_t>(a.cp.outer_bases_qual);
          dqual = dqual < oq ? dqual : oq;
        }
      }
      // final re-orientation, :783-784
      uint32_t oi = job.rc_out ? len - 1u - i : i;
      uint32_t ob = job.rc_out ? complement_base(dbase) : dbase;
      a.out_base[job.out_off + oi] = static_cast<uint8_t>(ob);
      a.out_qual[job.out_off + oi] = static_cast<uint8_t>(dqual);
      a.out_depth[job.out_off + oi] = static_cast<uint16_t>(depth);
      a.out_errors[job.out_off + oi] = static_cast<uint16_t>(err);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      n_dup += __shfl_xor_sync(0xFFFFFFFFu, n_dup, off);
      n_dis += __shfl_xor_sync(0xFFFFFFFFu, n_dis, off);
    }
    if (lane == 0) {
      uint8_t st = FGB_CODEC_OK;
      if (n_dup > 0) {                                             // :1155-1166
        double rate = static_cast<double>(n_dis) / static_cast<double>(n_dup);
        tot_bases += n_dup;
        tot_dis += n_dis;
        if (n_dis > a.cp.max_duplex_disagreements) st = FGB_CODEC_HIGH_DISAGREEMENT_COUNT;
        else if (rate > a.cp.max_duplex_disagreement_rate) st = FGB_CODEC_HIGH_DISAGREEMENT_RATE;
      }
      a.status[j] = st;
      if (a.disagreements) a.disagreements[j] = n_dis;
      if (a.duplex_bases) a.duplex_bases[j] = n_dup;
      ++done;
    }
  }
  if (lane == 0) {
    if (tot_bases) atomicAdd(a.counters + FGB_CTR_DUPLEX_BASES, tot_bases);
    if (tot_dis) atomicAdd(a.counters + FGB_CTR_DUPLEX_DISAGREE, tot_dis);
    if (done) atomicAdd(a.counters + FGB_CTR_COMBINED, static_cast<unsigned long long>(done));
  }
}

}  // namespace fgb

// K4 — consensus filter as an epilogue of the vote, single-strand reads.
// What `fgumi filter` does to a simplex consensus record, on the columns while they are still in
// HBM: per-base masking (crates/fgumi-consensus/src/filter.rs:655-696 mask_bases), the per-read depth
// and error-rate gates (:453-471 filter_read, fed by the cD / cE values the caller would write,
// caller.rs:322-329) and the no-call / mean-quality gates (commands/filter.rs:909-929).
// One warp per unit; integer work plus one f32 and two f64 divisions per unit.
#pragma once
#include <stdint.h>

#include "../../include/fgumi_b200.h"

namespace fgb {

struct FilterArgs {
  const fgb_unit* units;        // biased like the vote's
  uint64_t unit_begin, unit_end;
  uint8_t* base;                // consensus columns, masked in place
  uint8_t* qual;
  const uint16_t* depth;
  const uint16_t* errors;
  const uint16_t* emax;         // [65536]: largest error count e with (double)e / (double)d <= max_base_error_rate
  uint8_t* status;              // per unit (index u - unit_begin)
  uint32_t* masked;             // per unit newly masked bases, may be null
  unsigned long long* counters; // device counter block (FGB_CTR_FILTER_*)
  uint32_t min_reads;
  uint32_t min_base_quality;    // 0 = none (no quality is below 0)
  uint32_t per_base_tags;       // 0: the record carries no cd/ce arrays -> depth 0 everywhere (filter.rs:677-678)
  double max_read_error_rate;
  double min_mean_base_quality; // < 0 = none
  double max_no_call_fraction;
};

__global__ void __launch_bounds__(256) filter_simplex_kernel(const FilterArgs a) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
  unsigned long long pass_cnt = 0, masked_cnt = 0, rec_cnt = 0;
  for (uint64_t u = a.unit_begin + ((static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5);
       u < a.unit_end; u += warps) {
    const fgb_unit un = a.units[u];
    const uint32_t L = un.cons_len;
    if (L == 0) {                               // no consensus read was produced for this unit
      if (lane == 0) { a.status[u - a.unit_begin] = FGB_FILTER_NO_RECORD; if (a.masked) a.masked[u - a.unit_begin] = 0; }
      continue;
    }
    uint32_t maxd = 0, td = 0, te = 0, ncount = 0, qsum = 0, newly = 0;
    for (uint32_t p = lane; p < L; p += 32u) {
      const uint64_t o = un.out_off + p;
      uint32_t b = a.base[o], q = a.qual[o];
      const uint32_t d = a.depth[o], e = a.errors[o];
      maxd = d > maxd ? d : maxd; td += d; te += e;
      // the filter reads cd / ce back from the record, where the simplex caller stored them clamped to
      // i16::MAX (vanilla_caller.rs:1410-1412)
      const uint32_t dt = a.per_base_tags ? (d < 32767u ? d : 32767u) : 0u, et = a.per_base_tags ? (e < 32767u ? e : 32767u) : 0u;
      const bool mask = q < a.min_base_quality || dt < a.min_reads || (dt > 0u && et > a.emax[dt]);
      if (mask) {
        newly += (b != 'N');
        b = 'N'; q = 2u;
        a.base[o] = 'N'; a.qual[o] = 2;
      }
      if (b == 'N') ++ncount; else qsum += q;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const uint32_t m = __shfl_xor_sync(0xFFFFFFFFu, maxd, off);
      maxd = m > maxd ? m : maxd;
      td += __shfl_xor_sync(0xFFFFFFFFu, td, off);
      te += __shfl_xor_sync(0xFFFFFFFFu, te, off);
      ncount += __shfl_xor_sync(0xFFFFFFFFu, ncount, off);
      qsum += __shfl_xor_sync(0xFFFFFFFFu, qsum, off);
      newly += __shfl_xor_sync(0xFFFFFFFFu, newly, off);
    }
    if (lane == 0) {
      uint32_t st = FGB_FILTER_PASS;
      const float ce = td == 0u ? 0.0f : __fdiv_rn(__uint2float_rn(te), __uint2float_rn(td));   // caller.rs:322-329
      if (maxd < a.min_reads) st = FGB_FILTER_INSUFFICIENT_READS;                               // cD
      else if (static_cast<double>(ce) > a.max_read_error_rate) st = FGB_FILTER_EXCESSIVE_ERROR_RATE;
      else {
        const uint32_t non_n = L - ncount;
        const double mean = non_n == 0u ? 0.0 : __ddiv_rn(static_cast<double>(qsum), static_cast<double>(non_n));
        if (a.min_mean_base_quality >= 0.0 && mean < a.min_mean_base_quality) st = FGB_FILTER_LOW_MEAN_QUALITY;
        else if (a.max_no_call_fraction >= 1.0) {
          if (static_cast<double>(ncount) > a.max_no_call_fraction) st = FGB_FILTER_TOO_MANY_NO_CALLS;
        } else if (__ddiv_rn(static_cast<double>(ncount), static_cast<double>(L)) > a.max_no_call_fraction) {
          st = FGB_FILTER_TOO_MANY_NO_CALLS;
        }
      }
      a.status[u - a.unit_begin] = static_cast<uint8_t>(st);
      if (a.masked) a.masked[u - a.unit_begin] = newly;
      ++rec_cnt; pass_cnt += (st == FGB_FILTER_PASS); masked_cnt += newly;
    }
  }
  if (lane == 0 && rec_cnt) {
    atomicAdd(a.counters + FGB_CTR_FILTER_RECORDS, rec_cnt);
    atomicAdd(a.counters + FGB_CTR_FILTER_PASSED, pass_cnt);
    atomicAdd(a.counters + FGB_CTR_FILTER_BASES_MASKED, masked_cnt);
  }
}

}  // namespace fgb

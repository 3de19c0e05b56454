// K5 — consensus-record assembly on the device (for simplex units).
// build_consensus_record_into (vanilla_caller.rs:1365-1473) over UnmappedSamBuilder (raw-bam builder.rs:90-230)
// and the tag encoders (tags.rs:512-667): the device writes the finished `[u32 block_size][BAM record]` bytes of
// every unit at its final offset in the ConsensusOutput stream, from the columns the vote just left in HBM.
// The host supplies only what it alone knows -- per unit: read type, UMI, cell barcode, the RX consensus
// (simple_umi.rs:236-245, computed at planning time) -- and the record's offset; sizes are known up front because
// cD / cM fit one byte when a unit has at most 255 reads (units beyond that take the host assembly).
// One warp per unit; every field is written with byte stores at consecutive addresses across the lanes.
#pragma once
#include <stdint.h>

#include "../../include/fgumi_b200.h"

namespace fgb {

struct AssembleArgs {
  const fgb_unit* units;            // biased like the vote's (absolute unit indices)
  const fgb_record_job* jobs;       // biased the same way; jobs[u].out_off = byte offset of the record in the stream
  uint64_t unit_begin, unit_end;
  const uint8_t* base;              // consensus columns (biased by the chunk's output origin)
  const uint8_t* qual;
  const uint16_t* depth;
  const uint16_t* errors;
  const uint8_t* strings;           // string blob: [prefix][read group id] then the units' UMI / cell / RX bytes
  uint32_t prefix_len, rg_len;      // prefix at strings[0], read group id at strings[prefix_len]
  uint8_t cell_tag[2];
  uint8_t per_base_tags;
  uint8_t pad;
  uint8_t* out;                     // output stream, biased so that out[jobs[u].out_off] is the record's first byte
};

__device__ __forceinline__ void put_bytes(uint8_t* dst, const uint8_t* src, uint32_t n, uint32_t lane) {
  for (uint32_t i = lane; i < n; i += 32u) dst[i] = src[i];
}

__global__ void __launch_bounds__(256) assemble_simplex_kernel(const AssembleArgs a) {
  // base letter -> 4-bit code ("=ACMGRSVTWYHKDBN", either case; anything else 15: sequence.rs:183-209)
  __shared__ uint8_t code[256];
  {
    uint32_t c = 15u;
    switch (threadIdx.x) {
      case '=': c = 0; break;
      case 'A': case 'a': c = 1; break;  case 'C': case 'c': c = 2; break;  case 'M': case 'm': c = 3; break;
      case 'G': case 'g': c = 4; break;  case 'R': case 'r': c = 5; break;  case 'S': case 's': c = 6; break;
      case 'V': case 'v': c = 7; break;  case 'T': case 't': c = 8; break;  case 'W': case 'w': c = 9; break;
      case 'Y': case 'y': c = 10; break; case 'H': case 'h': c = 11; break; case 'K': case 'k': c = 12; break;
      case 'D': case 'd': c = 13; break; case 'B': case 'b': c = 14; break;
      default: break;
    }
    code[threadIdx.x] = static_cast<uint8_t>(c);
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
  for (uint64_t u = a.unit_begin + ((static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5);
       u < a.unit_end; u += warps) {
    const fgb_record_job jb = a.jobs[u];
    if (jb.flags & FGB_RECJOB_SKIP) continue;                    // no record for this unit
    const fgb_unit un = a.units[u];
    const uint32_t L = un.cons_len;
    const uint8_t* bs = a.base + un.out_off;
    const uint8_t* qs = a.qual + un.out_off;
    const uint16_t* ds = a.depth + un.out_off;
    const uint16_t* es = a.errors + un.out_off;
    uint8_t* const rec = a.out + jb.out_off;
    const uint8_t* str = a.strings + jb.str_off;                 // [umi][cell][rx]
    const uint32_t name_len = a.prefix_len + 1u + jb.umi_len;
    // ---- depth / error statistics (cD, cM, cE: caller.rs:322-329) ----
    uint32_t mx = 0, mn = L ? 0xFFFFFFFFu : 0u;
    unsigned long long td = 0, te = 0;
    for (uint32_t p = lane; p < L; p += 32u) {
      const uint32_t d = ds[p];
      mx = d > mx ? d : mx; mn = d < mn ? d : mn;
      td += d; te += es[p];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const uint32_t omx = __shfl_xor_sync(0xFFFFFFFFu, mx, off), omn = __shfl_xor_sync(0xFFFFFFFFu, mn, off);
      mx = omx > mx ? omx : mx; mn = omn < mn ? omn : mn;
      td += __shfl_xor_sync(0xFFFFFFFFu, td, off);
      te += __shfl_xor_sync(0xFFFFFFFFu, te, off);
    }
    // ---- fixed fields, name (builder.rs:113-139) ----
    uint32_t o = 0;
    if (lane == 0) {
      const uint32_t bsz = jb.size - 4u;
      uint16_t flag = 0x4;                                        // unmapped
      if (jb.read_type == 1) flag |= 0x1 | 0x40 | 0x8;           // R1: paired, first, mate unmapped (vanilla_caller.rs:1379-1388)
      else if (jb.read_type == 2) flag |= 0x1 | 0x80 | 0x8;
      const uint8_t hdr[36] = {
          static_cast<uint8_t>(bsz), static_cast<uint8_t>(bsz >> 8), static_cast<uint8_t>(bsz >> 16), static_cast<uint8_t>(bsz >> 24),
          0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF,         // refID -1, pos -1
          static_cast<uint8_t>(name_len + 1u), 0,                  // l_read_name, mapq
          static_cast<uint8_t>(4680 & 0xFF), static_cast<uint8_t>(4680 >> 8), 0, 0,   // bin, n_cigar
          static_cast<uint8_t>(flag), static_cast<uint8_t>(flag >> 8),
          static_cast<uint8_t>(L), static_cast<uint8_t>(L >> 8), static_cast<uint8_t>(L >> 16), static_cast<uint8_t>(L >> 24),
          0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0, 0, 0, 0};   // next refID -1, next pos -1, tlen 0
      for (int i = 0; i < 36; ++i) rec[i] = hdr[i];
    }
    o = 36;
    put_bytes(rec + o, a.strings, a.prefix_len, lane); o += a.prefix_len;
    if (lane == 0) rec[o] = ':';
    o += 1;
    put_bytes(rec + o, str, jb.umi_len, lane); o += jb.umi_len;
    if (lane == 0) rec[o] = 0;
    o += 1;
    // ---- packed sequence, qualities ----
    const uint32_t nb = (L + 1u) >> 1;
    for (uint32_t k = lane; k < nb; k += 32u) {
      const uint32_t hi = code[bs[2 * k]], lo = 2 * k + 1 < L ? code[bs[2 * k + 1]] : 0u;
      rec[o + k] = static_cast<uint8_t>((hi << 4) | lo);
    }
    o += nb;
    put_bytes(rec + o, qs, L, lane); o += L;
    // ---- tags (vanilla_caller.rs:1393-1444): RG, cD, cM, cE, [cd, ce], MI, [cell], [RX] ----
    if (lane == 0) { rec[o] = 'R'; rec[o + 1] = 'G'; rec[o + 2] = 'Z'; rec[o + 3 + a.rg_len] = 0; }
    put_bytes(rec + o + 3, a.strings + a.prefix_len, a.rg_len, lane);
    o += 3 + a.rg_len + 1;
    if (lane == 0) {
      // smallest integer type, tags.rs:533-553; the values are <= 255 here (the unit has at most 255 reads)
      rec[o] = 'c'; rec[o + 1] = 'D'; rec[o + 2] = mx <= 127u ? 'c' : 'C'; rec[o + 3] = static_cast<uint8_t>(mx);
      rec[o + 4] = 'c'; rec[o + 5] = 'M'; rec[o + 6] = mn <= 127u ? 'c' : 'C'; rec[o + 7] = static_cast<uint8_t>(mn);
      const float ce = td > 0 ? __fdiv_rn(__ull2float_rn(te), __ull2float_rn(td)) : 0.0f;
      const uint32_t cb = __float_as_uint(ce);
      rec[o + 8] = 'c'; rec[o + 9] = 'E'; rec[o + 10] = 'f';
      rec[o + 11] = static_cast<uint8_t>(cb); rec[o + 12] = static_cast<uint8_t>(cb >> 8);
      rec[o + 13] = static_cast<uint8_t>(cb >> 16); rec[o + 14] = static_cast<uint8_t>(cb >> 24);
    }
    o += 15;
    if (a.per_base_tags) {
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const uint16_t* v = which ? es : ds;
        if (lane == 0) {
          rec[o] = 'c'; rec[o + 1] = which ? 'e' : 'd'; rec[o + 2] = 'B'; rec[o + 3] = 's';
          rec[o + 4] = static_cast<uint8_t>(L); rec[o + 5] = static_cast<uint8_t>(L >> 8);
          rec[o + 6] = static_cast<uint8_t>(L >> 16); rec[o + 7] = static_cast<uint8_t>(L >> 24);
        }
        for (uint32_t p = lane; p < L; p += 32u) {                // values clamp to i16::MAX (:1410-1412)
          const uint32_t x = v[p] > 32767u ? 32767u : v[p];
          rec[o + 8 + 2 * p] = static_cast<uint8_t>(x);
          rec[o + 9 + 2 * p] = static_cast<uint8_t>(x >> 8);
        }
        o += 8 + 2 * L;
      }
    }
    if (lane == 0) { rec[o] = 'M'; rec[o + 1] = 'I'; rec[o + 2] = 'Z'; rec[o + 3 + jb.umi_len] = 0; }
    put_bytes(rec + o + 3, str, jb.umi_len, lane);
    o += 3 + jb.umi_len + 1;
    if (jb.flags & FGB_RECJOB_HAS_CELL) {
      if (lane == 0) { rec[o] = a.cell_tag[0]; rec[o + 1] = a.cell_tag[1]; rec[o + 2] = 'Z'; rec[o + 3 + jb.cell_len] = 0; }
      put_bytes(rec + o + 3, str + jb.umi_len, jb.cell_len, lane);
      o += 3 + jb.cell_len + 1;
    }
    if (jb.flags & FGB_RECJOB_HAS_RX) {
      if (lane == 0) { rec[o] = 'R'; rec[o + 1] = 'X'; rec[o + 2] = 'Z'; rec[o + 3 + jb.rx_len] = 0; }
      put_bytes(rec + o + 3, str + jb.umi_len + jb.cell_len, jb.rx_len, lane);
      o += 3 + jb.rx_len + 1;
    }
  }
}

}  // namespace fgb

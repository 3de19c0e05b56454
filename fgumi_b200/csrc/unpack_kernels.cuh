// K0 — input unpack kernels: turn a compact host transfer format into the two byte columns
// (bases[], quals[]) the vote kernel stages.  Pure byte work at HBM speed; they exist so that the
// PCIe link, which bounds end-to-end throughput, carries fewer bytes per observation.
//
//   PACK8  one byte per observation of an already prepared SourceRead row
//          (vanilla_caller.rs:129-146): bits 7..6 = A,C,G,T, bits 5..0 = quality 0..61;
//          0x3E = (N, Q2), the masked base of create_source_read (:908-916).
//   BAM4   the record's own payload: 4-bit packed sequence (raw-bam sequence.rs:9-35) + raw quality
//          bytes; the kernel does the per-base part of create_source_read (:893-916): orientation
//          (reverse-complement + reversed qualities), and the min-input-quality mask.
#pragma once
#include <stdint.h>

#include "../../include/fgumi_b200.h"

namespace fgb {

// 4 packed observations -> 4 base bytes + 4 quality bytes
__device__ __forceinline__ void unpack8_word(uint32_t w, uint32_t* base, uint32_t* qual) {
  const uint32_t qf = w & 0x3F3F3F3Fu;
  const uint32_t sf = (w >> 6) & 0x03030303u;
  const uint32_t hi = (sf >> 1) & 0x01010101u;                 // G or T
  const uint32_t t = sf & hi;                                  // T
  uint32_t b = 0x41414141u + 2u * sf + 2u * hi + 0x0Bu * t;    // A 41, C 43, G 47, T 54 (no carries)
  const uint32_t x = qf ^ 0x3E3E3E3Eu;                         // zero byte <=> (N, Q2)
  const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
  const uint32_t m = (z >> 7) * 0xFFu;
  *base = (b & ~m) | (0x4E4E4E4Eu & m);
  *qual = (qf & ~m) | (0x02020202u & m);
}

// n16 = number of 16-byte chunks; all three pointers 16-byte aligned
__global__ void __launch_bounds__(256) unpack8_kernel(const uint4* __restrict__ packed,
                                                      uint4* __restrict__ bases,
                                                      uint4* __restrict__ quals, uint64_t n16) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) {
    const uint4 p = packed[i];
    uint4 b, q;
    unpack8_word(p.x, &b.x, &q.x);
    unpack8_word(p.y, &b.y, &q.y);
    unpack8_word(p.z, &b.z, &q.z);
    unpack8_word(p.w, &b.w, &q.w);
    bases[i] = b;
    quals[i] = q;
  }
}

}  // namespace fgb

namespace fgb {

struct Bam4Args {
  const uint8_t* seq4;          // biased: nibble i of the batch lives in seq4[i >> 1] (high nibble first)
  const uint8_t* quals_raw;     // biased: quals_raw[i]
  const fgb_raw_read* raw_reads;   // biased: raw_reads[r] for absolute read index r
  const uint64_t* reads;        // layout descriptors (off << 16 | final_len), biased the same way
  uint8_t* bases;               // byte columns being built (biased by the chunk origin)
  uint8_t* quals;
  uint64_t read_begin, read_end;   // absolute read range of this launch
  uint64_t raw_lo, raw_hi;      // raw index range [raw_lo, raw_hi) that is resident (bounds check)
  uint32_t* bad;                // set to 1 when a raw span breaks the layout rules (read is skipped)
  uint32_t min_q;               // min_input_base_quality (0 = no masking)
};

// One warp per read, lanes over the row's 64-bit words.  Output position p of the row comes from
// raw base p (forward strand) or raw_len-1-p complemented (reverse strand) of the kept raw span
// (vanilla_caller.rs:893-898); q < min_q turns the observation into (N, Q2) (:908-916).
__global__ void __launch_bounds__(256) unpack_bam4_kernel(const Bam4Args a) {
  // "=ACMGRSVTWYHKDBN" and its complement (A<->T, C<->G, everything else unchanged; fgumi-dna dna.rs:30-60)
  __shared__ uint8_t lut[32];
  if (threadIdx.x < 32) {
    const char* f = "=ACMGRSVTWYHKDBN";
    const char* c = "=TGMCRSVAWYHKDBN";
    lut[threadIdx.x] = static_cast<uint8_t>(threadIdx.x < 16 ? f[threadIdx.x] : c[threadIdx.x - 16]);
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
  for (uint64_t r = a.read_begin + ((static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5);
       r < a.read_end; r += warps) {
    const fgb_raw_read rr = a.raw_reads[r];
    const uint64_t d = a.reads[r];
    const uint32_t len = static_cast<uint32_t>(d & 0xFFFFu);
    const uint64_t off = d >> 16;
    // layout rules of fgb_raw_columns, checked where the data is touched (no host pass over the reads)
    if ((rr.src_off & 1u) || rr.src_off < a.raw_lo || rr.src_off + rr.raw_len > a.raw_hi || len > rr.raw_len) {
      if (lane == 0) atomicOr(a.bad, 1u);
      continue;
    }
    const bool rev = rr.flags & 1u;
    const uint8_t* tab = lut + (rev ? 16 : 0);
    const uint32_t words = (len + 7u) >> 3;
    for (uint32_t w = lane; w < words; w += 32u) {
      uint64_t wb = 0, wq = 0;
#pragma unroll
      for (uint32_t j = 0; j < 8u; ++j) {
        const uint32_t p = w * 8u + j;
        if (p < len) {
          const uint64_t i = rr.src_off + (rev ? rr.raw_len - 1u - p : p);
          const uint32_t byte = __ldg(a.seq4 + (i >> 1));
          const uint32_t nib = (i & 1u) ? (byte & 15u) : (byte >> 4);
          uint32_t b = tab[nib];
          uint32_t q = __ldg(a.quals_raw + i);
          if (q < a.min_q) { b = 'N'; q = 2u; }
          wb |= static_cast<uint64_t>(b) << (8u * j);
          wq |= static_cast<uint64_t>(q) << (8u * j);
        }
      }
      *reinterpret_cast<uint64_t*>(a.bases + off + w * 8u) = wb;
      *reinterpret_cast<uint64_t*>(a.quals + off + w * 8u) = wq;
    }
  }
}

// Output narrowing in front of the device->host copy: 8 u16 -> 8 u8 per thread and column.
__global__ void __launch_bounds__(256) narrow_u16_kernel(const uint4* __restrict__ depth16,
                                                         const uint4* __restrict__ errors16,
                                                         uint2* __restrict__ depth8,
                                                         uint2* __restrict__ errors8, uint64_t n8) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 d = depth16[i], e = errors16[i];
    depth8[i] = make_uint2(__byte_perm(d.x, d.y, 0x6420u), __byte_perm(d.z, d.w, 0x6420u));
    errors8[i] = make_uint2(__byte_perm(e.x, e.y, 0x6420u), __byte_perm(e.z, e.w, 0x6420u));
  }
}

}  // namespace fgb

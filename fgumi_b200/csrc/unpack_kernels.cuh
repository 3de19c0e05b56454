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

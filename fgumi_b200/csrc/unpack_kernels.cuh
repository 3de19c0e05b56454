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
#include "swar.cuh"

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

// ---- RECORDS: source-read rows built straight from the BAM records' own bytes -----------------------
// The host ships the records as they are (one DMA from pinned memory, no per-base host work) plus one
// fgb_raw_read per surviving read: src_off = byte offset of the record's packed-sequence field inside the
// record blob, raw_len = l_seq (the quality bytes start (l_seq + 1) / 2 bytes after src_off, raw-bam
// fields.rs:6-23), flags bit 0 = reverse strand.  The kernel does the per-base part of create_source_read
// (vanilla_caller.rs:893-916) and of to_source_read_for_codec_raw (codec_caller.rs:414-469): row position p
// comes from raw base p (forward) or l_seq - 1 - p complemented (reverse), `q < min_q` turns the
// observation into (N, Q2); the row length (after mate clip / quality trim / trailing-N strip) is the
// host's decision and comes with the layout descriptor.
namespace fgb {

struct RecordsArgs {
  const uint8_t* records;          // biased: blob byte i lives at records[i]
  const fgb_raw_read* raw_reads;   // biased: raw_reads[r] for absolute read index r
  const uint64_t* reads;           // layout descriptors (off << 16 | row length), biased the same way
  uint8_t* bases;                  // byte columns being built (biased by the chunk origin)
  uint8_t* quals;
  uint64_t read_begin, read_end;   // absolute read range of this launch
  uint64_t rec_lo, rec_hi;         // blob byte range [rec_lo, rec_hi) that is resident (bounds check)
  uint32_t* bad;                   // set to 1 when a record span breaks the layout rules (read is skipped)
  uint32_t min_q;                  // min_input_base_quality (0 = no masking)
};

// ---- the row builder shared by RECORDS and BAM4: flat (read, word) items over chunks of reads ----------------
// One warp per read left 13 of 32 lanes idle on a 150-base row (19 words) and walked the reads one at a time.  Here a
// CTA takes kRowChunk consecutive reads, 64 threads fetch and check their descriptors, a prefix sum of the rows' word
// counts goes to shared memory and the 256 threads take 8-position words from ONE flat index (6-step search).
#ifndef FGB_ROW_CHUNK
#define FGB_ROW_CHUNK 256
#endif
constexpr uint32_t kRowChunk = FGB_ROW_CHUNK;     // 32 .. 256 (power of two): reads whose descriptors are fetched in one round

struct RowSrc {                    // one read of the chunk, as the word builder needs it
  const uint8_t* seq;              // packed sequence, high nibble first (raw-bam sequence.rs:9-35)
  const uint8_t* qual;             // raw qualities
  uint64_t off;                    // row offset in the byte columns
  uint32_t L;                      // raw bases in the record
  uint32_t len_rev;                // row length | reverse strand << 31
};

// 4 / 8 bytes at an arbitrary address through aligned 32-bit loads; with Guard, words outside [lo, hi) (4-aligned
// byte addresses: the resident part of a transfer column) read as zero instead of being touched.
template <bool Guard>
__device__ __forceinline__ uint32_t ldg_word(const uint32_t* w, const uintptr_t lo, const uintptr_t hi) {
  if (Guard && reinterpret_cast<uintptr_t>(w) - lo >= hi - lo) return 0u;
  return __ldg(w);
}
template <bool Guard>
__device__ __forceinline__ uint32_t ldg_u32_at(const uint8_t* p, const uintptr_t lo, const uintptr_t hi) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
  return __funnelshift_r(ldg_word<Guard>(w, lo, hi), ldg_word<Guard>(w + 1, lo, hi), static_cast<uint32_t>(a & 3u) * 8u);
}
template <bool Guard>
__device__ __forceinline__ uint2 ldg_u64_at(const uint8_t* p, const uintptr_t lo, const uintptr_t hi) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
  const uint32_t sh = static_cast<uint32_t>(a & 3u) * 8u;
  const uint32_t w0 = ldg_word<Guard>(w, lo, hi), w1 = ldg_word<Guard>(w + 1, lo, hi), w2 = ldg_word<Guard>(w + 2, lo, hi);
  return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}

// Word w (row positions 8w .. 8w+7) of one source-read row: the per-base part of create_source_read
// (vanilla_caller.rs:893-916).  lut_f / lut_r: one packed byte -> two ASCII bytes in row order (reverse: complemented).
template <bool Guard>
__device__ __forceinline__ void build_row_word(const RowSrc& r, const uint32_t w, const uint16_t* lut_f,
                                               const uint16_t* lut_r, const uint32_t min_q,
                                               const uintptr_t s_lo, const uintptr_t s_hi, const uintptr_t q_lo_b,
                                               const uintptr_t q_hi_b, uint2* ob, uint2* oq) {
  const uint32_t len = r.len_rev & 0x7FFFFFFFu;
  const bool rev = (r.len_rev >> 31) != 0u;
  uint32_t x, q_lo, q_hi;
  if (!rev) {
    x = ldg_u32_at<Guard>(r.seq + 4u * w, s_lo, s_hi);       // byte k = bases 8w+2k (high nibble), 8w+2k+1
    const uint2 q = ldg_u64_at<Guard>(r.qual + 8u * w, q_lo_b, q_hi_b);
    q_lo = q.x; q_hi = q.y;
  } else {
    // row positions 8w+j <- raw bases e-j, e = L-1-8w: the eight nibbles ending at raw index e, as a 32-bit
    // value whose nibble j (from the least significant end) is raw base e-j
    const int32_t e = static_cast<int32_t>(r.L) - 1 - static_cast<int32_t>(8u * w);
    const int32_t first = e - 7;                              // may be negative on the row's last word
    const int32_t s0 = first >> 1;                            // floor: byte holding raw base `first`
    const uint2 v = ldg_u64_at<Guard>(r.seq + s0, s_lo, s_hi);   // bytes s0 .. s0+7, little endian
    const uint64_t be = (static_cast<uint64_t>(__byte_perm(v.x, 0u, 0x0123u)) << 8) | (v.y & 0xFFu);   // bytes s0 .. s0+4, big endian
    const uint32_t t0 = static_cast<uint32_t>(first - 2 * s0);   // 0 or 1
    x = static_cast<uint32_t>(be >> (4u * (2u - t0)));
    const uint2 q = ldg_u64_at<Guard>(r.qual + first, q_lo_b, q_hi_b);   // raw qualities first .. e
    q_lo = __byte_perm(q.y, 0u, 0x0123u);                     // reversed: position j <- raw e - j
    q_hi = __byte_perm(q.x, 0u, 0x0123u);
  }
  const uint16_t* lut = rev ? lut_r : lut_f;
  uint32_t b_lo = static_cast<uint32_t>(lut[x & 0xFFu]) | (static_cast<uint32_t>(lut[(x >> 8) & 0xFFu]) << 16);
  uint32_t b_hi = static_cast<uint32_t>(lut[(x >> 16) & 0xFFu]) | (static_cast<uint32_t>(lut[x >> 24]) << 16);
  if (min_q) {                                                // vanilla_caller.rs:908-916
    uint32_t ok_lo, ok_hi;
    if (min_q <= 127u) {
      const uint32_t ts = min_q * 0x01010101u;                // high bit survives iff (q & 0x7F) >= t; q >= 128 passes outright
      ok_lo = (((q_lo | 0x80808080u) - ts) | q_lo) & 0x80808080u;
      ok_hi = (((q_hi | 0x80808080u) - ts) | q_hi) & 0x80808080u;
    } else {
      ok_lo = ok_hi = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ok_lo |= (((q_lo >> (8 * j)) & 0xFFu) >= min_q) ? (0x80u << (8 * j)) : 0u;
        ok_hi |= (((q_hi >> (8 * j)) & 0xFFu) >= min_q) ? (0x80u << (8 * j)) : 0u;
      }
    }
    const uint32_t m_lo = spread_msb(ok_lo ^ 0x80808080u);    // 0xFF where the base is masked
    const uint32_t m_hi = spread_msb(ok_hi ^ 0x80808080u);
    b_lo = (b_lo & ~m_lo) | (0x4E4E4E4Eu & m_lo); q_lo = (q_lo & ~m_lo) | (0x02020202u & m_lo);
    b_hi = (b_hi & ~m_hi) | (0x4E4E4E4Eu & m_hi); q_hi = (q_hi & ~m_hi) | (0x02020202u & m_hi);
  }
  const uint32_t live = len - 8u * w;                         // positions of this word inside the row (>= 1)
  if (live < 8u) {                                            // zero the row padding
    const uint32_t k_lo = live >= 4u ? 0xFFFFFFFFu : ((1u << (8u * live)) - 1u);
    const uint32_t k_hi = live <= 4u ? 0u : ((1u << (8u * (live - 4u))) - 1u);
    b_lo &= k_lo; q_lo &= k_lo; b_hi &= k_hi; q_hi &= k_hi;
  }
  *ob = make_uint2(b_lo, b_hi);
  *oq = make_uint2(q_lo, q_hi);
}

__device__ __forceinline__ void fill_pair_luts(uint16_t* lut_f, uint16_t* lut_r) {
  // pair tables: one packed byte (two bases) -> two ASCII bytes; forward: (high, low) nibble in row order,
  // reverse: (low, high) nibble complemented ("=ACMGRSVTWYHKDBN", A<->T, C<->G; fgumi-dna dna.rs:30-60)
  const char* f = "=ACMGRSVTWYHKDBN";
  const char* c = "=TGMCRSVAWYHKDBN";
  const uint32_t b = threadIdx.x;
  lut_f[b] = static_cast<uint16_t>(static_cast<uint8_t>(f[b >> 4]) | (static_cast<uint32_t>(static_cast<uint8_t>(f[b & 15])) << 8));
  lut_r[b] = static_cast<uint16_t>(static_cast<uint8_t>(c[b & 15]) | (static_cast<uint32_t>(static_cast<uint8_t>(c[b >> 4])) << 8));
}

// The chunk loop.  Describe(r, &src) fills the RowSrc of absolute read r and returns false when its spans break the
// layout rules (the read is skipped and *bad set).
template <bool Guard, class Describe>
__device__ __forceinline__ void unpack_rows(const uint64_t read_begin, const uint64_t read_end, uint8_t* bases,
                                            uint8_t* quals, const uint32_t min_q, uint32_t* bad, const uintptr_t s_lo,
                                            const uintptr_t s_hi, const uintptr_t q_lo, const uintptr_t q_hi,
                                            Describe&& describe) {
  __shared__ uint16_t lut_f[256], lut_r[256];
  __shared__ RowSrc s_src[kRowChunk];
  __shared__ uint32_t s_pref[kRowChunk + 1];
  __shared__ uint32_t s_wsum[kRowChunk / 32];
  fill_pair_luts(lut_f, lut_r);
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  const uint64_t n = read_end - read_begin;
  const uint64_t chunks = (n + kRowChunk - 1) / kRowChunk;
  for (uint64_t c = blockIdx.x; c < chunks; c += gridDim.x) {
    __syncthreads();                                          // the tables are filled / the last chunk's items are done
    const uint64_t r0 = read_begin + c * kRowChunk;
    const uint32_t nr = static_cast<uint32_t>(read_end - r0 < kRowChunk ? read_end - r0 : kRowChunk);
    if (tid < kRowChunk) {
      uint32_t items = 0;
      if (tid < nr) {
        RowSrc src;
        if (describe(r0 + tid, &src)) items = ((src.len_rev & 0x7FFFFFFFu) + 7u) >> 3;
        else atomicOr(bad, 1u);
        s_src[tid] = src;
      }
      uint32_t incl = items;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if (lane >= static_cast<uint32_t>(off)) incl += v;
      }
      if (lane == 31u) s_wsum[tid >> 5] = incl;
      s_pref[tid + 1] = incl;                                 // within the warp; the warps in front are added below
    }
    __syncthreads();
    if (kRowChunk > 32u && tid >= 32u && tid < kRowChunk) {
      uint32_t add = 0;
      for (uint32_t w = 0; w < (tid >> 5); ++w) add += s_wsum[w];
      s_pref[tid + 1] += add;
    }
    if (tid == 0) s_pref[0] = 0u;
    __syncthreads();
    const uint32_t total = s_pref[kRowChunk];
    for (uint32_t it = tid; it < total; it += blockDim.x) {
      uint32_t rl = 0;                                        // last read with s_pref[rl] <= it
#pragma unroll
      for (uint32_t step = kRowChunk / 2; step > 0; step >>= 1)
        if (s_pref[rl + step] <= it) rl += step;
      const uint32_t w = it - s_pref[rl];
      uint2 ob, oq;
      build_row_word<Guard>(s_src[rl], w, lut_f, lut_r, min_q, s_lo, s_hi, q_lo, q_hi, &ob, &oq);
      const uint64_t o = s_src[rl].off + w * 8u;
      *reinterpret_cast<uint2*>(bases + o) = ob;
      *reinterpret_cast<uint2*>(quals + o) = oq;
    }
  }
}

__global__ void __launch_bounds__(256) unpack_records_kernel(const RecordsArgs a) {
  unpack_rows<false>(a.read_begin, a.read_end, a.bases, a.quals, a.min_q, a.bad, 0, 0, 0, 0,
                     [&](uint64_t r, RowSrc* s) {
    const fgb_raw_read rr = a.raw_reads[r];
    const uint64_t d = a.reads[r];
    const uint32_t len = static_cast<uint32_t>(d & 0xFFFFu);
    const uint32_t L = rr.raw_len;
    const uint64_t qoff = rr.src_off + ((static_cast<uint64_t>(L) + 1u) >> 1);
    s->seq = a.records + rr.src_off; s->qual = a.records + qoff; s->off = d >> 16; s->L = L;
    s->len_rev = len | ((rr.flags & 1u) << 31);
    // the sequence and quality fields must lie inside the resident blob (with 16 bytes of slack either side
    // for the aligned window loads), and the row cannot be longer than the record's sequence
    return !(rr.src_off < a.rec_lo + 16u || qoff + L + 16u > a.rec_hi || len > L);
  });
}

// BAM4 through the same builder: the packed sequence and the raw qualities come in two columns (nibble i of the batch
// in seq4[i >> 1], high nibble first; a read's span starts on an even nibble), so a read looks exactly like a record's
// sequence / quality fields.  The window loads are guarded: nothing outside the resident part of the columns is
// touched (4-byte granularity; fgb_raw_columns asks for columns padded to a multiple of 4 bytes).
__global__ void __launch_bounds__(256) unpack_bam4_words_kernel(const Bam4Args a) {
  const uintptr_t s_lo = reinterpret_cast<uintptr_t>(a.seq4 + (a.raw_lo >> 1)) & ~static_cast<uintptr_t>(3);
  const uintptr_t s_hi = (reinterpret_cast<uintptr_t>(a.seq4 + ((a.raw_hi + 1u) >> 1)) + 3u) & ~static_cast<uintptr_t>(3);
  const uintptr_t q_lo = reinterpret_cast<uintptr_t>(a.quals_raw + a.raw_lo) & ~static_cast<uintptr_t>(3);
  const uintptr_t q_hi = (reinterpret_cast<uintptr_t>(a.quals_raw + a.raw_hi) + 3u) & ~static_cast<uintptr_t>(3);
  unpack_rows<true>(a.read_begin, a.read_end, a.bases, a.quals, a.min_q, a.bad, s_lo, s_hi, q_lo, q_hi,
                    [&](uint64_t r, RowSrc* s) {
    const fgb_raw_read rr = a.raw_reads[r];
    const uint64_t d = a.reads[r];
    const uint32_t len = static_cast<uint32_t>(d & 0xFFFFu);
    s->seq = a.seq4 + (rr.src_off >> 1); s->qual = a.quals_raw + rr.src_off; s->off = d >> 16; s->L = rr.raw_len;
    s->len_rev = len | ((rr.flags & 1u) << 31);
    // layout rules of fgb_raw_columns, checked where the data is touched (no host pass over the reads)
    return !((rr.src_off & 1u) || rr.src_off < a.raw_lo || rr.src_off + rr.raw_len > a.raw_hi || len > rr.raw_len);
  });
}

}  // namespace fgb

// ---- K0o: overlapping-bases pre-pass on the uploaded records --------------------------
// OverlappingBasesConsensusCaller::call (overlapping.rs:236-337) per base, in place on the record blob in HBM,
// ahead of the row builder.  The host plans the runs from the headers and CIGARs (csrc/host/overlap.h
// plan_group); one thread walks the runs of one pair in order (they may meet inside a packed sequence byte).
namespace fgb {

struct OverlapArgs {
  uint8_t* records;                // biased: blob byte i lives at records[i]
  const fgb_overlap_run* runs;
  uint64_t n_runs;
  uint64_t rec_lo, rec_hi;         // resident blob range
  unsigned long long* stats;       // device u64[4]: overlapping, agreeing, disagreeing, corrected
  uint32_t* bad;
  uint32_t agree, disagree;        // FGB_OVERLAP_AGREE_* / FGB_OVERLAP_DISAGREE_*
};

__global__ void __launch_bounds__(256) overlap_kernel(const OverlapArgs a) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  unsigned long long n_ov = 0, n_ag = 0, n_dis = 0, n_cor = 0;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n_runs; i += stride) {
    if (a.runs[i].flags & FGB_RUN_CONTINUES) continue;           // walked by the thread of the pair's first run
    for (uint64_t k = i; k < a.n_runs && (k == i || (a.runs[k].flags & FGB_RUN_CONTINUES)); ++k) {
      const fgb_overlap_run r = a.runs[k];
      const uint64_t q1o = r.seq1_off + ((static_cast<uint64_t>(r.l_seq1) + 1u) >> 1);
      const uint64_t q2o = r.seq2_off + ((static_cast<uint64_t>(r.l_seq2) + 1u) >> 1);
      if (r.seq1_off < a.rec_lo || r.seq2_off < a.rec_lo || q1o + r.l_seq1 > a.rec_hi || q2o + r.l_seq2 > a.rec_hi ||
          static_cast<uint64_t>(r.o1) + r.len > r.l_seq1 || static_cast<uint64_t>(r.o2) + r.len > r.l_seq2) {
        atomicOr(a.bad, 1u);
        continue;
      }
      uint8_t* s1 = a.records + r.seq1_off; uint8_t* q1 = a.records + q1o;
      uint8_t* s2 = a.records + r.seq2_off; uint8_t* q2 = a.records + q2o;
      for (uint32_t t = 0; t < r.len; ++t) {
        const uint32_t i1 = r.o1 + t, i2 = r.o2 + t;
        const uint32_t b1 = s1[i1 >> 1], b2 = s2[i2 >> 1];
        const uint32_t c1 = (i1 & 1u) ? (b1 & 15u) : (b1 >> 4), c2 = (i2 & 1u) ? (b2 & 15u) : (b2 >> 4);
        if (c1 == 15u || c2 == 15u) continue;                     // a no-call in either mate: the position is skipped
        ++n_ov;
        const uint32_t x = q1[i1], y = q2[i2];
        uint32_t oc1 = c1, oc2 = c2, oq1 = x, oq2 = y;
        if (c1 == c2) {
          ++n_ag;
          if (a.agree == FGB_OVERLAP_AGREE_PASS_THROUGH) continue;
          const uint32_t nq = a.agree == FGB_OVERLAP_AGREE_CONSENSUS ? (x + y > 93u ? 93u : x + y) : (x > y ? x : y);
          oq1 = oq2 = nq;
          if (nq != x || nq != y) ++n_cor;
        } else {
          ++n_dis;
          if (a.disagree == FGB_OVERLAP_DISAGREE_CONSENSUS) {      // higher quality wins with the difference; tie -> N
            uint32_t code = 15u, q = 2u;
            if (x > y) { code = c1; q = x - y < 2u ? 2u : x - y; }
            else if (y > x) { code = c2; q = y - x < 2u ? 2u : y - x; }
            oc1 = oc2 = code; oq1 = oq2 = q;
            n_cor += 2;
          } else if (a.disagree == FGB_OVERLAP_DISAGREE_MASK_BOTH || x == y) {
            oc1 = oc2 = 15u; oq1 = oq2 = 2u;
            n_cor += 2;
          } else if (x < y) {
            oc1 = 15u; oq1 = 2u; ++n_cor;
          } else {
            oc2 = 15u; oq2 = 2u; ++n_cor;
          }
        }
        if (oc1 != c1) s1[i1 >> 1] = static_cast<uint8_t>((i1 & 1u) ? ((b1 & 0xF0u) | oc1) : ((oc1 << 4) | (b1 & 0x0Fu)));
        if (oc2 != c2) {
          const uint32_t bb = s2[i2 >> 1];                        // re-read: s1 and s2 never alias, but keep it simple
          s2[i2 >> 1] = static_cast<uint8_t>((i2 & 1u) ? ((bb & 0xF0u) | oc2) : ((oc2 << 4) | (bb & 0x0Fu)));
        }
        if (oq1 != x) q1[i1] = static_cast<uint8_t>(oq1);
        if (oq2 != y) q2[i2] = static_cast<uint8_t>(oq2);
      }
    }
  }
  // warp-reduce the counters, one atomic per warp and counter
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    n_ov += __shfl_down_sync(0xFFFFFFFFu, n_ov, off);
    n_ag += __shfl_down_sync(0xFFFFFFFFu, n_ag, off);
    n_dis += __shfl_down_sync(0xFFFFFFFFu, n_dis, off);
    n_cor += __shfl_down_sync(0xFFFFFFFFu, n_cor, off);
  }
  if ((threadIdx.x & 31u) == 0) {
    if (n_ov) atomicAdd(a.stats + 0, n_ov);
    if (n_ag) atomicAdd(a.stats + 1, n_ag);
    if (n_dis) atomicAdd(a.stats + 2, n_dis);
    if (n_cor) atomicAdd(a.stats + 3, n_cor);
  }
}

}  // namespace fgb

// DEFLATE (RFC 1951) decoder for one BGZF member, written once for the host and the device (FGB_HD): the device kernel
// (inflate_kernel.cuh) runs one member per thread with the decoder's tables in shared memory, the CPU tests run the very
// same functions against zlib.  BGZF members are independent and hold at most 64 KiB, which is what makes the input
// side of a file-level run a device workload: the link carries the compressed members, a third
// of the bytes, and the records appear in HBM where the row builder reads them.
//
// The role of the reference's fgumi-bgzf reader (crates/fgumi-bgzf/src/reader.rs over libdeflate).  Canonical-code
// decoding by length counts (no per-symbol tables to build: 2 x 16 counters + the symbols in code order), every
// access to input and output bounds-checked: a corrupt member ends with an error code, never with a stray access.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define FGB_HD __host__ __device__ __forceinline__
#else
#define FGB_HD inline
#endif

namespace fgb {
namespace inflate {

enum : uint32_t {
  kOk = 0,
  kErrInput = 1,        // ran out of input bits
  kErrOutput = 2,       // more output than the member declares
  kErrBlockType = 3,
  kErrStoredLen = 4,
  kErrCodeLengths = 5,  // invalid or over-subscribed code-length / literal / distance code
  kErrSymbol = 6,       // invalid symbol or a distance beyond the start of the output
  kErrShort = 7,        // the stream ended before the declared output size
  kErrCrc = 8
};

template <int N>
struct Huffman {        // canonical code: number of codes of each length, symbols ordered by (length, value)
  uint16_t count[16];
  uint16_t symbol[N];
};
struct Tables {         // 1 044 bytes: one per thread in shared memory on the device
  Huffman<288> lit;     // also holds the code-length code while a dynamic block's lengths are read
  Huffman<30> dist;
  uint8_t lengths[19 + 286 + 30 + 9];
};

struct Bits {
  const uint8_t* in;
  uint32_t n_in, pos;   // bytes
  uint64_t acc;
  uint32_t cnt;         // valid bits in acc
  uint32_t starved;     // set when bits were asked for past the end of the input
};

FGB_HD void bits_init(Bits& b, const uint8_t* in, uint32_t n) { b.in = in; b.n_in = n; b.pos = 0; b.acc = 0; b.cnt = 0; b.starved = 0; }
FGB_HD void bits_fill(Bits& b) {
  while (b.cnt <= 56 && b.pos < b.n_in) { b.acc |= static_cast<uint64_t>(b.in[b.pos++]) << b.cnt; b.cnt += 8; }
}
FGB_HD uint32_t bits_get(Bits& b, uint32_t n) {            // n <= 16
  if (b.cnt < n) { bits_fill(b); if (b.cnt < n) { b.starved = 1; b.cnt = 0; b.acc = 0; return 0; } }
  const uint32_t v = static_cast<uint32_t>(b.acc) & ((1u << n) - 1u);
  b.acc >>= n; b.cnt -= n;
  return v;
}

// Builds the canonical decoding structure for n code lengths; returns 0 when the code is complete, >0 when it is
// incomplete (allowed only for a single-code distance alphabet), <0 when it is over-subscribed.
template <int N>
FGB_HD int huffman_build(Huffman<N>& h, const uint8_t* length, int n) {
  for (int l = 0; l < 16; ++l) h.count[l] = 0;
  for (int s = 0; s < n; ++s) h.count[length[s]]++;
  if (h.count[0] == n) return 0;                           // no codes: complete, but decoding will fail
  int left = 1;
  for (int l = 1; l < 16; ++l) {
    left <<= 1;
    left -= h.count[l];
    if (left < 0) return left;
  }
  uint16_t offs[16];
  offs[1] = 0;
  for (int l = 1; l < 15; ++l) offs[l + 1] = static_cast<uint16_t>(offs[l] + h.count[l]);
  for (int s = 0; s < n; ++s)
    if (length[s] != 0) h.symbol[offs[length[s]]++] = static_cast<uint16_t>(s);
  return left;
}

// One symbol: walk the code lengths (bit-reversed codes arrive LSB first, one bit at a time).
template <int N>
FGB_HD int huffman_decode(Bits& b, const Huffman<N>& h) {
  if (b.cnt < 15) bits_fill(b);
  int code = 0, first = 0, index = 0;
  uint64_t acc = b.acc;
  const uint32_t avail = b.cnt;
  for (uint32_t len = 1; len <= 15; ++len) {
    if (len > avail) { b.starved = 1; return -1; }
    code |= static_cast<int>(acc & 1u);
    acc >>= 1;
    const int count = h.count[len];
    if (code - count < first) {
      b.acc = acc; b.cnt = avail - len;
      return h.symbol[index + (code - first)];
    }
    index += count;
    first += count;
    first <<= 1;
    code <<= 1;
  }
  return -1;
}

struct Consts {
  uint16_t len_base[29], dist_base[30];
  uint8_t len_extra[29], dist_extra[30];
  uint8_t order[19];
};
FGB_HD void consts_init(Consts& c) {
  const uint16_t lb[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  const uint8_t le[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  const uint16_t db[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  const uint8_t de[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  const uint8_t od[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  for (int i = 0; i < 29; ++i) { c.len_base[i] = lb[i]; c.len_extra[i] = le[i]; }
  for (int i = 0; i < 30; ++i) { c.dist_base[i] = db[i]; c.dist_extra[i] = de[i]; }
  for (int i = 0; i < 19; ++i) c.order[i] = od[i];
}

// Literal / length / distance symbols of one block into out[*op ...), never past out_len.
FGB_HD uint32_t inflate_codes(Bits& b, const Tables& t, const Consts& k, uint8_t* out, uint32_t out_len, uint32_t* op) {
  uint32_t o = *op;
  for (;;) {
    const int sym = huffman_decode(b, t.lit);
    if (sym < 0) return b.starved ? kErrInput : kErrSymbol;
    if (sym < 256) {
      if (o >= out_len) return kErrOutput;
      out[o++] = static_cast<uint8_t>(sym);
    } else if (sym == 256) {
      *op = o;
      return kOk;
    } else {
      const int ls = sym - 257;
      if (ls >= 29) return kErrSymbol;
      uint32_t len = k.len_base[ls] + bits_get(b, k.len_extra[ls]);
      const int ds = huffman_decode(b, t.dist);
      if (ds < 0) return b.starved ? kErrInput : kErrSymbol;
      if (ds >= 30) return kErrSymbol;
      const uint32_t dist = k.dist_base[ds] + bits_get(b, k.dist_extra[ds]);
      if (b.starved) return kErrInput;
      if (dist > o) return kErrSymbol;
      if (len > out_len - o) return kErrOutput;
      const uint8_t* from = out + (o - dist);
      for (uint32_t i = 0; i < len; ++i) out[o + i] = from[i];      // may overlap forward: byte order matters
      o += len;
    }
  }
}

// One whole DEFLATE stream (all its blocks) of one member.  `t` is scratch (shared memory on the device).
FGB_HD uint32_t inflate_member(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t out_len, Tables& t, const Consts& k) {
  Bits b;
  bits_init(b, in, in_len);
  uint32_t o = 0;
  for (;;) {
    const uint32_t last = bits_get(b, 1), type = bits_get(b, 2);
    if (b.starved) return kErrInput;
    if (type == 0) {                                          // stored
      b.acc >>= (b.cnt & 7u); b.cnt -= (b.cnt & 7u);          // to the byte boundary
      const uint32_t len = bits_get(b, 16), nlen = bits_get(b, 16);
      if (b.starved) return kErrInput;
      if ((len ^ 0xFFFFu) != nlen) return kErrStoredLen;
      if (len > out_len - o) return kErrOutput;
      // bytes still in the accumulator first, then straight from the input
      uint32_t i = 0;
      while (i < len && b.cnt >= 8) { out[o + i++] = static_cast<uint8_t>(b.acc); b.acc >>= 8; b.cnt -= 8; }
      if (len - i > b.n_in - b.pos) return kErrInput;
      for (; i < len; ++i) out[o + i] = b.in[b.pos++];
      o += len;
    } else if (type == 1) {                                   // fixed codes
      for (int s = 0; s < 144; ++s) t.lengths[s] = 8;
      for (int s = 144; s < 256; ++s) t.lengths[s] = 9;
      for (int s = 256; s < 280; ++s) t.lengths[s] = 7;
      for (int s = 280; s < 288; ++s) t.lengths[s] = 8;
      huffman_build(t.lit, t.lengths, 288);
      for (int s = 0; s < 30; ++s) t.lengths[s] = 5;
      huffman_build(t.dist, t.lengths, 30);
      const uint32_t st = inflate_codes(b, t, k, out, out_len, &o);
      if (st != kOk) return st;
    } else if (type == 2) {                                   // dynamic codes
      const uint32_t nlen = bits_get(b, 5) + 257, ndist = bits_get(b, 5) + 1, ncode = bits_get(b, 4) + 4;
      if (b.starved) return kErrInput;
      if (nlen > 286 || ndist > 30) return kErrCodeLengths;
      for (uint32_t i = 0; i < 19; ++i) t.lengths[i] = 0;
      for (uint32_t i = 0; i < ncode; ++i) t.lengths[k.order[i]] = static_cast<uint8_t>(bits_get(b, 3));
      if (b.starved) return kErrInput;
      if (huffman_build(t.lit, t.lengths, 19) != 0) return kErrCodeLengths;   // the code-length code must be complete
      uint32_t idx = 0;
      while (idx < nlen + ndist) {
        const int sym = huffman_decode(b, t.lit);
        if (sym < 0) return b.starved ? kErrInput : kErrCodeLengths;
        if (sym < 16) { t.lengths[19 + idx++] = static_cast<uint8_t>(sym); continue; }
        uint32_t rep, val = 0;
        if (sym == 16) {
          if (idx == 0) return kErrCodeLengths;
          val = t.lengths[19 + idx - 1];
          rep = 3 + bits_get(b, 2);
        } else if (sym == 17) rep = 3 + bits_get(b, 3);
        else rep = 11 + bits_get(b, 7);
        if (b.starved) return kErrInput;
        if (idx + rep > nlen + ndist) return kErrCodeLengths;
        while (rep--) t.lengths[19 + idx++] = static_cast<uint8_t>(val);
      }
      if (t.lengths[19 + 256] == 0) return kErrCodeLengths;   // no end-of-block code
      // (the lengths live behind the 19 code-length lengths: both builds read them before anything overwrites)
      int r = huffman_build(t.lit, t.lengths + 19, static_cast<int>(nlen));
      if (r < 0 || (r > 0 && nlen - t.lit.count[0] != 1)) return kErrCodeLengths;
      r = huffman_build(t.dist, t.lengths + 19 + nlen, static_cast<int>(ndist));
      if (r < 0 || (r > 0 && ndist - t.dist.count[0] != 1)) return kErrCodeLengths;
      const uint32_t st = inflate_codes(b, t, k, out, out_len, &o);
      if (st != kOk) return st;
    } else {
      return kErrBlockType;
    }
    if (last) break;
  }
  return o == out_len ? kOk : kErrShort;
}

// CRC-32 (IEEE 802.3), one byte at a time over a 256-entry table (crc_table_init fills it).
FGB_HD void crc_table_init(uint32_t* table, uint32_t i) {   // entry i
  uint32_t c = i;
  for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
  table[i] = c;
}
FGB_HD uint32_t crc32_bytes(const uint32_t* table, const uint8_t* p, uint32_t n) {
  uint32_t c = 0xFFFFFFFFu;
  for (uint32_t i = 0; i < n; ++i) c = (c >> 8) ^ table[(c ^ p[i]) & 0xFFu];
  return ~c;
}

}  // namespace inflate
}  // namespace fgb

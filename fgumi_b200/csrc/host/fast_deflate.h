// A small, fast DEFLATE (RFC 1951) encoder for BGZF members: one block of at most 65 280 input bytes -> one
// dynamic-Huffman block (or a stored block when that is smaller).  It is what fgb_bgzf_compress uses at level 1:
// the image's zlib manages ~60 MB/s per thread there, and the consensus BAM stream is half 6-bit-random quality
// bytes on which an elaborate match search finds nothing.  Greedy LZ77 with one hash probe per position (4-byte
// hash, 8 K-entry table, block-local 16-bit positions), matches extended 8 bytes at a time, then exact Huffman code
// lengths for the block's own symbol statistics (length-limited by frequency halving), canonical codes, run-length
// coded header.  No dependency on zlib; the CRC-32 (IEEE 802.3) of the gzip trailer is slicing-by-8.
//
// The role of the reference's fgumi-bgzf writer over libdeflate (crates/fgumi-bgzf); the algorithm is RFC 1951's, the
// code is this repo's.  Host code, header-only, re-entrant (all state lives in a caller-provided Scratch).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>

namespace fgb {
namespace fastdeflate {

constexpr uint32_t kMaxIn = 0xFF00;
constexpr uint32_t kHashBits = 13;

struct Scratch {
  uint32_t tokens[kMaxIn + 8];          // literal: byte; match: 1 << 31 | (len - 3) << 16 | (dist - 1)
  uint16_t head[1u << kHashBits];       // position + 1 of the last occurrence of a hash, 0 = none
  uint32_t lfreq[288], dfreq[32];
  uint8_t llen[288], dlen[32];
  uint16_t lcode[288], dcode[32];
};

// ---- CRC-32 -------------------------------------------------------------------------------------
struct CrcTables {
  uint32_t t[8][256];
  CrcTables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFFu];
  }
};
inline uint32_t crc32(const uint8_t* p, size_t n) {
  static const CrcTables T;
  uint32_t c = 0xFFFFFFFFu;
  while (n >= 8) {
    uint32_t a, b;
    std::memcpy(&a, p, 4); std::memcpy(&b, p + 4, 4);
    a ^= c;
    c = T.t[7][a & 0xFFu] ^ T.t[6][(a >> 8) & 0xFFu] ^ T.t[5][(a >> 16) & 0xFFu] ^ T.t[4][a >> 24] ^
        T.t[3][b & 0xFFu] ^ T.t[2][(b >> 8) & 0xFFu] ^ T.t[1][(b >> 16) & 0xFFu] ^ T.t[0][b >> 24];
    p += 8; n -= 8;
  }
  while (n--) c = (c >> 8) ^ T.t[0][(c ^ *p++) & 0xFFu];
  return ~c;
}

// ---- symbol tables of RFC 1951 section 3.2.5 ------------------------------------------------------
struct SymTables {
  uint8_t len_sym[256];      // len - 3 -> length code - 257
  uint8_t len_extra[29];
  uint16_t len_base[29];
  uint8_t dist_sym[512];     // zlib's trick: d - 1 < 256 ? [d - 1] : [256 + ((d - 1) >> 7)]
  uint8_t dist_extra[30];
  uint16_t dist_base[30];
  SymTables() {
    static const uint8_t le[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    uint32_t l = 3;
    for (int c = 0; c < 29; ++c) {
      len_extra[c] = le[c];
      len_base[c] = static_cast<uint16_t>(c == 28 ? 258 : l);
      if (c < 28) for (uint32_t k = 0; k < (1u << le[c]); ++k) len_sym[l++ - 3] = static_cast<uint8_t>(c);
    }
    len_sym[255] = 28;                                       // length 258 has its own code
    uint32_t d = 1;
    for (int c = 0; c < 30; ++c) {
      const uint32_t e = c < 2 ? 0 : static_cast<uint32_t>(c / 2 - 1);
      dist_extra[c] = static_cast<uint8_t>(e);
      dist_base[c] = static_cast<uint16_t>(d);
      for (uint32_t k = 0; k < (1u << e); ++k, ++d) {
        const uint32_t i = d - 1;
        if (i < 256) dist_sym[i] = static_cast<uint8_t>(c);
        else dist_sym[256 + (i >> 7)] = static_cast<uint8_t>(c);   // codes >= 16 span whole multiples of 128
      }
    }
  }
};
inline const SymTables& sym() { static const SymTables S; return S; }
inline uint32_t dist_symbol(uint32_t dist_m1) { return dist_m1 < 256 ? sym().dist_sym[dist_m1] : sym().dist_sym[256 + (dist_m1 >> 7)]; }

// ---- Huffman code lengths, limited to `limit` bits --------------------------------------------------
// Plain Huffman on the used symbols (two-queue merge over the sorted frequencies); if the tree is deeper than the
// limit the frequencies are halved (floor 1) and the tree rebuilt -- flatter every time, optimal when it was not
// needed.  A lone used symbol gets one bit (RFC 1951: a single distance code is sent with one bit).
inline void huffman_lengths(const uint32_t* freq, int n, int limit, uint8_t* len) {
  struct Node { uint64_t w; int16_t l, r; };
  int order[288], used = 0;
  for (int i = 0; i < n; ++i) { len[i] = 0; if (freq[i]) order[used++] = i; }
  if (used == 0) return;
  if (used == 1) { len[order[0]] = 1; return; }
  uint32_t f[288];
  for (int i = 0; i < used; ++i) f[i] = freq[order[i]];
  for (;;) {
    int idx[288];
    for (int i = 0; i < used; ++i) idx[i] = i;
    std::sort(idx, idx + used, [&](int a, int b) { return f[a] != f[b] ? f[a] < f[b] : a < b; });
    Node nodes[2 * 288];
    for (int i = 0; i < used; ++i) nodes[i] = Node{f[idx[i]], -1, -1};
    int leaf = 0, in0 = used, in1 = used;                     // queue of leaves / queue of internal nodes [in0, in1)
    auto take = [&]() {
      if (leaf < used && (in0 >= in1 || nodes[leaf].w <= nodes[in0].w)) return leaf++;
      return in0++;
    };
    for (int k = 0; k < used - 1; ++k) {
      const int a = take(), b = take();
      nodes[in1] = Node{nodes[a].w + nodes[b].w, static_cast<int16_t>(a), static_cast<int16_t>(b)};
      ++in1;
    }
    uint8_t depth[2 * 288];
    const int root = in1 - 1;                                 // used >= 2: root >= used
    if (root < used || root >= 2 * 288) return;
    depth[root] = 0;
    int maxd = 0;
    for (int k = root; k >= used; --k) {                      // children have smaller indices than their parent
      depth[nodes[k].l] = depth[nodes[k].r] = static_cast<uint8_t>(depth[k] + 1);
    }
    for (int i = 0; i < used; ++i) maxd = std::max<int>(maxd, depth[i]);
    if (maxd <= limit) {
      for (int i = 0; i < used; ++i) len[order[idx[i]]] = depth[i];
      return;
    }
    for (int i = 0; i < used; ++i) f[i] = (f[i] + 1) >> 1;
  }
}

// Canonical codes (RFC 1951 3.2.2), bit-reversed for the LSB-first bit stream.
inline void canonical_codes(const uint8_t* len, int n, uint16_t* code) {
  uint32_t count[16] = {0}, next[16] = {0};
  for (int i = 0; i < n; ++i) count[len[i]]++;
  count[0] = 0;
  uint32_t c = 0;
  for (int b = 1; b <= 15; ++b) { c = (c + count[b - 1]) << 1; next[b] = c; }
  for (int i = 0; i < n; ++i) {
    const uint32_t l = len[i];
    if (!l) { code[i] = 0; continue; }
    uint32_t v = next[l]++, r = 0;
    for (uint32_t k = 0; k < l; ++k) { r = (r << 1) | (v & 1u); v >>= 1; }
    code[i] = static_cast<uint16_t>(r);
  }
}

struct BitWriter {                                            // branch-free: fewer than 8 bits are ever pending
  uint8_t* p;
  uint64_t acc = 0;
  uint32_t n = 0;
  explicit BitWriter(uint8_t* out) : p(out) {}
  inline void put(uint32_t v, uint32_t bits) {                // bits <= 32; the caller keeps 8 bytes of slack
    acc |= static_cast<uint64_t>(v) << n;
    n += bits;
    std::memcpy(p, &acc, 8);
    p += n >> 3;
    acc >>= n & ~7u;
    n &= 7u;
  }
  inline uint8_t* finish() {
    if (n) { *p++ = static_cast<uint8_t>(acc); n = 0; }
    return p;
  }
};

inline uint32_t load32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
inline uint64_t load64(const uint8_t* p) { uint64_t v; std::memcpy(&v, p, 8); return v; }

// Compresses in[0, n) (n <= kMaxIn) into out (capacity `cap` >= n + 64) as ONE final DEFLATE block.
// Returns the number of bytes written (never more than n + 5: the stored form is the fallback).
inline size_t deflate_block(Scratch& S, const uint8_t* in, uint32_t n, uint8_t* out, size_t cap) {
  auto stored = [&]() -> size_t {
    out[0] = 0x01;                                            // BFINAL = 1, BTYPE = 00, padding
    out[1] = static_cast<uint8_t>(n); out[2] = static_cast<uint8_t>(n >> 8);
    out[3] = static_cast<uint8_t>(~n); out[4] = static_cast<uint8_t>((~n) >> 8);
    if (n) std::memcpy(out + 5, in, n);
    return 5u + n;
  };
  if (n < 16 || cap < static_cast<size_t>(n) + 64) return stored();
  const SymTables& T = sym();
  // ---- LZ77: greedy, one probe ----
  std::memset(S.head, 0, sizeof(S.head));
  std::memset(S.lfreq, 0, sizeof(S.lfreq));
  std::memset(S.dfreq, 0, sizeof(S.dfreq));
  uint32_t nt = 0, i = 0, miss = 0;
  const uint32_t last = n - 8;                                // positions below `last` may load 8 bytes
  while (i < last) {
    const uint32_t w = load32(in + i);
    const uint32_t h = (w * 2654435761u) >> (32 - kHashBits);
    const uint32_t cand = S.head[h];
    S.head[h] = static_cast<uint16_t>(i + 1);
    if (cand && i - (cand - 1) <= 32768u && load32(in + cand - 1) == w) {
      const uint32_t c = cand - 1;
      uint32_t len = 4;
      const uint32_t maxlen = std::min<uint32_t>(258, n - i);
      while (len + 8 <= maxlen) {
        const uint64_t x = load64(in + i + len) ^ load64(in + c + len);
        if (x) { len += static_cast<uint32_t>(__builtin_ctzll(x)) >> 3; goto matched; }
        len += 8;
      }
      while (len < maxlen && in[i + len] == in[c + len]) ++len;
    matched:
      if (len > maxlen) len = maxlen;
      const uint32_t dm1 = i - c - 1;
      S.tokens[nt++] = 0x80000000u | ((len - 3) << 16) | dm1;
      S.lfreq[257 + T.len_sym[len - 3]]++;
      S.dfreq[dist_symbol(dm1)]++;
      // index the second position of the match too (cheap, and it is what repeats of tag arrays hit)
      if (i + 1 < last) {
        const uint32_t h2 = (load32(in + i + 1) * 2654435761u) >> (32 - kHashBits);
        S.head[h2] = static_cast<uint16_t>(i + 2);
      }
      i += len;
      miss = 0;
    } else {
      // incompressible stretches (quality bytes, packed bases): after 32 misses in a row the search looks at every
      // second position, after 64 at every third, ... (the skipped bytes go out as literals)
      uint32_t step = 1u + (miss >> 5);
      if (step > 8u) step = 8u;
      if (i + step > last) step = last - i;
      ++miss;
      for (uint32_t k = 0; k < step; ++k) { S.tokens[nt++] = in[i + k]; S.lfreq[in[i + k]]++; }
      i += step;
    }
  }
  for (; i < n; ++i) { S.tokens[nt++] = in[i]; S.lfreq[in[i]]++; }
  S.lfreq[256] = 1;                                           // end of block
  // ---- codes ----
  huffman_lengths(S.lfreq, 286, 15, S.llen);
  huffman_lengths(S.dfreq, 30, 15, S.dlen);
  int n_lit = 286, n_dist = 30;
  while (n_lit > 257 && S.llen[n_lit - 1] == 0) --n_lit;
  while (n_dist > 1 && S.dlen[n_dist - 1] == 0) --n_dist;
  if (n_dist == 1 && S.dlen[0] == 0) S.dlen[0] = 1;           // no match at all: one unused 1-bit distance code
  canonical_codes(S.llen, n_lit, S.lcode);
  canonical_codes(S.dlen, n_dist, S.dcode);
  // ---- header: the two length tables, zeros run-length coded (symbols 17 / 18), other lengths as they are ----
  uint8_t seq[288 + 32 + 8], extra[288 + 32 + 8];
  int ns = 0;
  {
    uint8_t all[288 + 32];
    std::memcpy(all, S.llen, n_lit);
    std::memcpy(all + n_lit, S.dlen, n_dist);
    const int tot = n_lit + n_dist;
    for (int k = 0; k < tot;) {
      if (all[k] == 0) {
        int r = 1;
        while (k + r < tot && all[k + r] == 0 && r < 138) ++r;
        if (r >= 11) { seq[ns] = 18; extra[ns++] = static_cast<uint8_t>(r - 11); k += r; }
        else if (r >= 3) { seq[ns] = 17; extra[ns++] = static_cast<uint8_t>(r - 3); k += r; }
        else { seq[ns] = 0; extra[ns++] = 0; ++k; }
      } else { seq[ns] = all[k]; extra[ns++] = 0; ++k; }
    }
  }
  uint32_t cfreq[19] = {0};
  for (int k = 0; k < ns; ++k) cfreq[seq[k]]++;
  uint8_t clen[19];
  uint16_t ccode[19];
  huffman_lengths(cfreq, 19, 7, clen);
  {                                                           // the code-length code must be complete (zlib rejects
    int used = 0, only = 0;                                   // an incomplete one even with a single symbol)
    for (int k = 0; k < 19; ++k) if (clen[k]) { ++used; only = k; }
    if (used == 1) clen[only == 0 ? 1 : 0] = 1;
  }
  canonical_codes(clen, 19, ccode);
  static const uint8_t kOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  int n_clc = 19;
  while (n_clc > 4 && clen[kOrder[n_clc - 1]] == 0) --n_clc;
  // ---- size check before writing: header + body in bits ----
  uint64_t bits = 3 + 5 + 5 + 4 + 3ull * n_clc;
  for (int k = 0; k < ns; ++k) bits += clen[seq[k]] + (seq[k] == 17 ? 3 : seq[k] == 18 ? 7 : 0);
  for (int s = 0; s < n_lit; ++s) bits += static_cast<uint64_t>(S.lfreq[s]) * (S.llen[s] + (s >= 257 ? T.len_extra[s - 257] : 0));
  for (int s = 0; s < n_dist; ++s) bits += static_cast<uint64_t>(S.dfreq[s]) * (S.dlen[s] + T.dist_extra[s]);
  const size_t bytes = static_cast<size_t>((bits + 7) >> 3);
  if (bytes >= static_cast<size_t>(n) + 5 || bytes + 16 > cap) return stored();
  // ---- write ----
  BitWriter bw(out);
  bw.put(1, 1); bw.put(2, 2);                                 // BFINAL, BTYPE = 10 (dynamic Huffman)
  bw.put(static_cast<uint32_t>(n_lit - 257), 5); bw.put(static_cast<uint32_t>(n_dist - 1), 5);
  bw.put(static_cast<uint32_t>(n_clc - 4), 4);
  for (int k = 0; k < n_clc; ++k) bw.put(clen[kOrder[k]], 3);
  for (int k = 0; k < ns; ++k) {
    bw.put(ccode[seq[k]], clen[seq[k]]);
    if (seq[k] == 17) bw.put(extra[k], 3);
    else if (seq[k] == 18) bw.put(extra[k], 7);
  }
  uint32_t lit[256];                                          // a literal's code and length in one load
  for (int b = 0; b < 256; ++b) lit[b] = S.lcode[b] | (static_cast<uint32_t>(S.llen[b]) << 16);
  for (uint32_t k = 0; k < nt; ++k) {
    const uint32_t t = S.tokens[k];
    if (!(t & 0x80000000u)) { const uint32_t e = lit[t]; bw.put(e & 0xFFFFu, e >> 16); continue; }
    const uint32_t lm3 = (t >> 16) & 0xFFu, dm1 = t & 0x7FFFu;
    const uint32_t ls = T.len_sym[lm3], ds = dist_symbol(dm1);
    // code + extra bits of the length (<= 15 + 5), then of the distance (<= 15 + 13)
    bw.put(S.lcode[257 + ls] | ((lm3 + 3u - T.len_base[ls]) << S.llen[257 + ls]), S.llen[257 + ls] + T.len_extra[ls]);
    bw.put(S.dcode[ds] | ((dm1 + 1u - T.dist_base[ds]) << S.dlen[ds]), S.dlen[ds] + T.dist_extra[ds]);
  }
  bw.put(S.lcode[256], S.llen[256]);
  return static_cast<size_t>(bw.finish() - out);
}

}  // namespace fastdeflate
}  // namespace fgb

// BGZF framing of the callers' output stream (the role of the reference's fgumi-bgzf
// writer, crates/fgumi-bgzf): the `ConsensusOutput` bytes are already BAM records with their
// block_size words, so a BAM file is  header | records  cut into <= 64 KiB gzip members with the "BC"
// extra field, followed by the 28-byte EOF member (SAM spec §4.1).  Host code, blocks compressed on
// several threads.  zlib is bound at call time (dlopen "libz.so.1"), so the engine library itself
// carries no dependency on it; without zlib the calls fail with FGB_ERR_INVALID_ARG and a message.
#include <dlfcn.h>
#include <zlib.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../../include/fgumi_b200.h"
#include "fast_deflate.h"
#include "../inflate_core.h"

namespace {

constexpr size_t kBlockInput = 0xFF00;      // uncompressed bytes per member (htslib's choice)
constexpr size_t kHeader = 18, kFooter = 8;
const uint8_t kEof[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0x00, 0x42, 0x43,
                          0x02, 0x00, 0x1b, 0x00, 0x03, 0x00, 0, 0, 0, 0, 0, 0, 0, 0};

struct Zlib {
  void* handle = nullptr;
  int (*deflateInit2_)(z_streamp, int, int, int, int, int, const char*, int) = nullptr;
  int (*deflate)(z_streamp, int) = nullptr;
  int (*deflateEnd)(z_streamp) = nullptr;
  int (*deflateReset)(z_streamp) = nullptr;
  uLong (*crc32)(uLong, const Bytef*, uInt) = nullptr;
  int (*inflateInit2_)(z_streamp, int, const char*, int) = nullptr;
  int (*inflate)(z_streamp, int) = nullptr;
  int (*inflateEnd)(z_streamp) = nullptr;
  int (*inflateReset)(z_streamp) = nullptr;
  bool ok = false, ok_inflate = false;
};

const Zlib& zlib() {
  static Zlib z;
  static std::once_flag once;
  std::call_once(once, []() {
    z.handle = dlopen("libz.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!z.handle) z.handle = dlopen("libz.so", RTLD_NOW | RTLD_LOCAL);
    if (!z.handle) return;
    z.deflateInit2_ = reinterpret_cast<decltype(z.deflateInit2_)>(dlsym(z.handle, "deflateInit2_"));
    z.deflate = reinterpret_cast<decltype(z.deflate)>(dlsym(z.handle, "deflate"));
    z.deflateEnd = reinterpret_cast<decltype(z.deflateEnd)>(dlsym(z.handle, "deflateEnd"));
    z.deflateReset = reinterpret_cast<decltype(z.deflateReset)>(dlsym(z.handle, "deflateReset"));
    z.crc32 = reinterpret_cast<decltype(z.crc32)>(dlsym(z.handle, "crc32"));
    z.ok = z.deflateInit2_ && z.deflate && z.deflateEnd && z.deflateReset && z.crc32;
    z.inflateInit2_ = reinterpret_cast<decltype(z.inflateInit2_)>(dlsym(z.handle, "inflateInit2_"));
    z.inflate = reinterpret_cast<decltype(z.inflate)>(dlsym(z.handle, "inflate"));
    z.inflateEnd = reinterpret_cast<decltype(z.inflateEnd)>(dlsym(z.handle, "inflateEnd"));
    z.inflateReset = reinterpret_cast<decltype(z.inflateReset)>(dlsym(z.handle, "inflateReset"));
    z.ok_inflate = z.crc32 && z.inflateInit2_ && z.inflate && z.inflateEnd && z.inflateReset;
  });
  return z;
}

void put16(uint8_t* p, uint32_t v) { p[0] = static_cast<uint8_t>(v); p[1] = static_cast<uint8_t>(v >> 8); }
void put32(uint8_t* p, uint32_t v) { put16(p, v & 0xFFFFu); put16(p + 2, v >> 16); }

// One member; returns its size, 0 on failure.  `out` holds at least 64 KiB.
size_t compress_block(const Zlib& z, z_stream* zs, const uint8_t* in, size_t n, uint8_t* out) {
  if (z.deflateReset(zs) != Z_OK) return 0;
  zs->next_in = const_cast<Bytef*>(in);
  zs->avail_in = static_cast<uInt>(n);
  zs->next_out = out + kHeader;
  zs->avail_out = static_cast<uInt>(65536 - kHeader - kFooter);
  if (z.deflate(zs, Z_FINISH) != Z_STREAM_END) return 0;
  const size_t clen = (65536 - kHeader - kFooter) - zs->avail_out;
  const size_t total = kHeader + clen + kFooter;
  static const uint8_t kHead[16] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0x00, 0x42, 0x43, 0x02, 0x00};
  std::memcpy(out, kHead, 16);
  put16(out + 16, static_cast<uint32_t>(total - 1));                       // BSIZE
  put32(out + kHeader + clen, static_cast<uint32_t>(z.crc32(z.crc32(0, nullptr, 0), in, static_cast<uInt>(n))));
  put32(out + kHeader + clen + 4, static_cast<uint32_t>(n));               // ISIZE
  return total;
}

// One member through the built-in encoder (level 1): no zlib involved.
size_t compress_block_fast(fgb::fastdeflate::Scratch& S, const uint8_t* in, size_t n, uint8_t* out) {
  const size_t clen = fgb::fastdeflate::deflate_block(S, in, static_cast<uint32_t>(n), out + kHeader, 65536 - kHeader - kFooter);
  if (!clen || clen > 65536 - kHeader - kFooter) return 0;
  const size_t total = kHeader + clen + kFooter;
  static const uint8_t kHead[16] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0x00, 0x42, 0x43, 0x02, 0x00};
  std::memcpy(out, kHead, 16);
  put16(out + 16, static_cast<uint32_t>(total - 1));                       // BSIZE
  put32(out + kHeader + clen, fgb::fastdeflate::crc32(in, n));
  put32(out + kHeader + clen + 4, static_cast<uint32_t>(n));               // ISIZE
  return total;
}

}  // namespace

extern "C" {

size_t fgb_bgzf_bound(size_t len) {
  const size_t blocks = (len + kBlockInput - 1) / kBlockInput;
  return blocks * 65536 + sizeof(kEof);
}

fgb_status fgb_bgzf_compress(const uint8_t* data, size_t len, int level, uint32_t n_threads, int append_eof,
                             uint8_t* out, size_t cap, size_t* out_len) {
  if ((len && !data) || !out || !out_len || level < 0 || level > 9) return FGB_ERR_INVALID_ARG;
  // level 1 = the built-in encoder (fast_deflate.h); every other level is zlib's
  const bool fast = level == 1 && !std::getenv("FGB_BGZF_ZLIB");
  const Zlib& z = zlib();
  if (!fast && !z.ok) return FGB_ERR_INVALID_ARG;                          // zlib not available on this host
  const size_t blocks = (len + kBlockInput - 1) / kBlockInput;
  if (cap < blocks * 65536 + (append_eof ? sizeof(kEof) : 0)) return FGB_ERR_INVALID_ARG;   // fgb_bgzf_bound(len)
  uint32_t T = std::max<uint32_t>(1, std::min<uint32_t>(n_threads ? n_threads : 1, static_cast<uint32_t>(std::max<size_t>(blocks, 1))));
  std::vector<uint32_t> sizes(blocks, 0);
  std::vector<int> failed(T, 0);
  // every block is compressed into its own 64 KiB slot of `out`, then the slots are closed up
  auto work = [&](uint32_t t) {
    if (fast) {
      std::unique_ptr<fgb::fastdeflate::Scratch> S(new (std::nothrow) fgb::fastdeflate::Scratch);
      if (!S) { failed[t] = 1; return; }
      for (size_t b = t; b < blocks; b += T) {
        const size_t o = b * kBlockInput, n = std::min(kBlockInput, len - o);
        const size_t s = compress_block_fast(*S, data + o, n, out + b * 65536);
        if (!s) { failed[t] = 1; break; }
        sizes[b] = static_cast<uint32_t>(s);
      }
      return;
    }
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (z.deflateInit2_(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY, ZLIB_VERSION, static_cast<int>(sizeof(z_stream))) != Z_OK) { failed[t] = 1; return; }
    for (size_t b = t; b < blocks; b += T) {
      const size_t o = b * kBlockInput, n = std::min(kBlockInput, len - o);
      const size_t s = compress_block(z, &zs, data + o, n, out + b * 65536);
      if (!s) { failed[t] = 1; break; }
      sizes[b] = static_cast<uint32_t>(s);
    }
    z.deflateEnd(&zs);
  };
  std::vector<std::thread> th;
  for (uint32_t t = 1; t < T; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  for (int f : failed) if (f) return FGB_ERR_INVALID_ARG;
  size_t w = 0;
  for (size_t b = 0; b < blocks; ++b) {
    if (w != b * 65536) std::memmove(out + w, out + b * 65536, sizes[b]);
    w += sizes[b];
  }
  if (append_eof) { std::memcpy(out + w, kEof, sizeof(kEof)); w += sizeof(kEof); }
  *out_len = w;
  return FGB_OK;
}

// "BAM\1" | l_text | text | n_ref = 0: consensus reads are unmapped (ref_id -1), so no reference
// dictionary is needed for the records to be valid (commands write @HD / @RG / @PG here).
fgb_status fgb_bam_header(const char* sam_text, size_t l_text, uint8_t* out, size_t cap, size_t* out_len) {
  if ((l_text && !sam_text) || !out || !out_len || l_text > 0x7FFFFFFFu) return FGB_ERR_INVALID_ARG;
  const size_t total = 4 + 4 + l_text + 4;
  if (cap < total) return FGB_ERR_INVALID_ARG;
  std::memcpy(out, "BAM\1", 4);
  put32(out + 4, static_cast<uint32_t>(l_text));
  if (l_text) std::memcpy(out + 8, sam_text, l_text);
  put32(out + 8 + l_text, 0);
  *out_len = total;
  return FGB_OK;
}

// ---- reading (the input side of a file-level run: fgumi-bgzf reader.rs, raw-bam record framing) ----------------
namespace {
uint32_t get16(const uint8_t* p) { return p[0] | (static_cast<uint32_t>(p[1]) << 8); }
uint32_t get32(const uint8_t* p) { return get16(p) | (get16(p + 2) << 16); }

struct Member { size_t off, size, isize, out_off; };

// Walks the gzip members of a BGZF stream (SAM spec 4.1: BSIZE in the "BC" extra subfield).
bool scan_members(const uint8_t* d, size_t len, std::vector<Member>* out, size_t* total) {
  size_t p = 0, o = 0;
  while (p < len) {
    if (len - p < kHeader + kFooter || d[p] != 0x1f || d[p + 1] != 0x8b || d[p + 2] != 8 || !(d[p + 3] & 4)) return false;
    const size_t xlen = get16(d + p + 10);
    size_t q = p + 12, bsize = 0;
    if (q + xlen > len) return false;
    while (q + 4 <= p + 12 + xlen) {
      const size_t slen = get16(d + q + 2);
      if (d[q] == 'B' && d[q + 1] == 'C' && slen == 2 && q + 6 <= len) bsize = get16(d + q + 4) + 1u;
      q += 4 + slen;
    }
    if (bsize < 12 + xlen + kFooter || p + bsize > len) return false;
    const size_t isize = get32(d + p + bsize - 4);
    out->push_back(Member{p + 12 + xlen, bsize - 12 - xlen - kFooter, isize, o});
    o += isize;
    p += bsize;
  }
  *total = o;
  return true;
}
}  // namespace

// Sum of the members' ISIZE fields: the size fgb_bgzf_decompress needs.
fgb_status fgb_bgzf_uncompressed_size(const uint8_t* data, size_t len, size_t* size) {
  if ((len && !data) || !size) return FGB_ERR_INVALID_ARG;
  std::vector<Member> m;
  if (!scan_members(data, len, &m, size)) return FGB_ERR_LAYOUT;
  return FGB_OK;
}

// The member table fgb_bgzf_inflate_device takes (capi.cu): framing only, nothing is inflated here.
fgb_status fgb_bgzf_scan_members(const uint8_t* data, size_t len, fgb_bgzf_member* members, uint64_t cap,
                                 uint64_t* n_members, uint64_t* out_len) {
  if ((len && !data) || !n_members || !out_len) return FGB_ERR_INVALID_ARG;
  std::vector<Member> m;
  size_t total = 0;
  if (!scan_members(data, len, &m, &total)) return FGB_ERR_LAYOUT;
  for (const Member& mb : m) if (mb.isize > 65536 || mb.size > 0xFFFFFFFFull) return FGB_ERR_LAYOUT;
  *n_members = m.size();
  *out_len = total;
  if (!members || cap < m.size()) return FGB_OK;              // sizing call
  for (size_t i = 0; i < m.size(); ++i) {
    fgb_bgzf_member& o = members[i];
    o.in_off = m[i].off; o.out_off = m[i].out_off;
    o.in_len = static_cast<uint32_t>(m[i].size); o.out_len = static_cast<uint32_t>(m[i].isize);
    o.crc = get32(data + m[i].off + m[i].size);
    o.reserved = 0;
  }
  return FGB_OK;
}

// The device decoder's code on the host (csrc/inflate_core.h), one member.
uint32_t fgb_host_inflate_member(const uint8_t* payload, uint32_t in_len, uint8_t* out, uint32_t out_len) {
  if ((in_len && !payload) || (out_len && !out)) return fgb::inflate::kErrInput;
  static const fgb::inflate::Consts k = [] { fgb::inflate::Consts c; fgb::inflate::consts_init(c); return c; }();
  fgb::inflate::Tables t;
  return fgb::inflate::inflate_member(payload, in_len, out, out_len, t, k);
}

// Inflates every member at its place in `out`, members dealt to n_threads threads; CRC32 and ISIZE are checked.
fgb_status fgb_bgzf_decompress(const uint8_t* data, size_t len, uint32_t n_threads, uint8_t* out, size_t cap,
                               size_t* out_len) {
  if ((len && !data) || !out_len || (cap && !out)) return FGB_ERR_INVALID_ARG;
  const Zlib& z = zlib();
  const bool own = !z.ok_inflate || std::getenv("FGB_BGZF_OWN_INFLATE");      // no zlib: the repo's decoder (slower)
  std::vector<Member> m;
  size_t total = 0;
  if (!scan_members(data, len, &m, &total)) return FGB_ERR_LAYOUT;
  if (total > cap) return FGB_ERR_INVALID_ARG;
  const uint32_t T = std::max<uint32_t>(1, std::min<uint32_t>(n_threads ? n_threads : 1, static_cast<uint32_t>(std::max<size_t>(m.size(), 1))));
  std::vector<int> failed(T, 0);
  auto work = [&](uint32_t t) {
    // contiguous runs of members per thread (sequential output per thread)
    const size_t a = m.size() * t / T, e = m.size() * (t + 1) / T;
    if (own) {
      for (size_t i = a; i < e; ++i) {
        const Member& mb = m[i];
        if (mb.isize > 65536 || mb.size > 0xFFFFFFFFull ||
            fgb_host_inflate_member(data + mb.off, static_cast<uint32_t>(mb.size), out + mb.out_off, static_cast<uint32_t>(mb.isize)) != 0 ||
            fgb::fastdeflate::crc32(out + mb.out_off, mb.isize) != get32(data + mb.off + mb.size)) { failed[t] = 1; break; }
      }
      return;
    }
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (z.inflateInit2_(&zs, -15, ZLIB_VERSION, static_cast<int>(sizeof(z_stream))) != Z_OK) { failed[t] = 1; return; }
    for (size_t i = a; i < e; ++i) {
      const Member& mb = m[i];
      if (z.inflateReset(&zs) != Z_OK) { failed[t] = 1; break; }
      zs.next_in = const_cast<Bytef*>(data + mb.off);
      zs.avail_in = static_cast<uInt>(mb.size);
      zs.next_out = out + mb.out_off;
      zs.avail_out = static_cast<uInt>(mb.isize);
      const int rc = z.inflate(&zs, Z_FINISH);
      if (!(rc == Z_STREAM_END || (rc == Z_OK && mb.isize == 0) || (rc == Z_BUF_ERROR && mb.isize == 0)) || zs.avail_out != 0) { failed[t] = 1; break; }
      const uint32_t crc = static_cast<uint32_t>(z.crc32(z.crc32(0, nullptr, 0), out + mb.out_off, static_cast<uInt>(mb.isize)));
      if (crc != get32(data + mb.off + mb.size)) { failed[t] = 1; break; }
    }
    z.inflateEnd(&zs);
  };
  std::vector<std::thread> th;
  for (uint32_t t = 1; t < T; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  for (int f : failed) if (f) return FGB_ERR_LAYOUT;
  *out_len = total;
  return FGB_OK;
}

// The header of an (uncompressed) BAM stream: "BAM\1", l_text, text, n_ref, references (l_name, name, l_ref).
// *text_off / *text_len locate the SAM text, *records_off the first record's block_size word.
fgb_status fgb_bam_read_header(const uint8_t* bam, size_t len, size_t* text_off, size_t* text_len, uint32_t* n_ref,
                               size_t* records_off) {
  if (!bam || !records_off || len < 12 || std::memcmp(bam, "BAM\1", 4) != 0) return FGB_ERR_LAYOUT;
  const size_t l_text = get32(bam + 4);
  if (8 + l_text + 4 > len) return FGB_ERR_LAYOUT;
  size_t p = 8 + l_text;
  const uint32_t nr = get32(bam + p);
  p += 4;
  for (uint32_t i = 0; i < nr; ++i) {
    if (p + 4 > len) return FGB_ERR_LAYOUT;
    const size_t l_name = get32(bam + p);
    if (p + 4 + l_name + 4 > len) return FGB_ERR_LAYOUT;
    p += 4 + l_name + 4;
  }
  if (text_off) *text_off = 8;
  if (text_len) *text_len = l_text;
  if (n_ref) *n_ref = nr;
  *records_off = p;
  return FGB_OK;
}

// The record section of a BAM stream ([u32 block_size][record]...) as the callers take it: the record bodies
// back to back in `bodies` (may be `stream` itself: the copy only moves bytes down) and rec_off[0..n] with record
// i at [rec_off[i], rec_off[i+1]).  Stops at the first incomplete record; *consumed = bytes of `stream` used.
fgb_status fgb_bam_split_records(const uint8_t* stream, size_t len, uint8_t* bodies, uint64_t* rec_off, uint64_t cap_records,
                                 uint64_t* n_records, size_t* consumed) {
  if ((len && !stream) || !bodies || !rec_off || !n_records) return FGB_ERR_INVALID_ARG;
  size_t p = 0, w = 0;
  uint64_t n = 0;
  rec_off[0] = 0;
  while (p + 4 <= len && n < cap_records) {
    const size_t bs = get32(stream + p);
    if (bs < 32) return FGB_ERR_LAYOUT;
    if (p + 4 + bs > len) break;
    std::memmove(bodies + w, stream + p + 4, bs);
    w += bs;
    p += 4 + bs;
    rec_off[++n] = w;
  }
  *n_records = n;
  if (consumed) *consumed = p;
  return FGB_OK;
}

}  // extern "C"

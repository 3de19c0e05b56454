// TEST INFRASTRUCTURE.  The repo's DEFLATE encoder (fgumi_b200/csrc/host/fast_deflate.h) and decoder
// (fgumi_b200/csrc/inflate_core.h, the code the device kernel runs) under ASan / UBSan against zlib:
//   * every block the encoder writes must inflate back with zlib, and with the own decoder;
//   * every stream zlib writes (levels 0-9, default / fixed / Huffman-only strategies) must inflate with the own decoder;
//   * damaged and truncated streams must end with a status, never with an access outside [out, out + out_len)
//     (redzones of the allocation are what ASan watches).
// usage: deflate_fuzz [iterations]
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../fgumi_b200/csrc/host/fast_deflate.h"
#include "../../fgumi_b200/csrc/inflate_core.h"

static std::vector<uint8_t> shape(std::mt19937& rng, int kind, size_t n) {
  std::vector<uint8_t> d(n);
  for (size_t i = 0; i < n; ++i) {
    switch (kind) {
      case 0: d[i] = static_cast<uint8_t>(rng()); break;                                  // incompressible
      case 1: d[i] = 0; break;                                                             // maximal matches
      case 2: d[i] = static_cast<uint8_t>("ACGT"[rng() & 3]); break;
      case 3: d[i] = static_cast<uint8_t>(i / 7); break;
      case 4: d[i] = (rng() % 100 < 90) ? 'x' : static_cast<uint8_t>(rng()); break;
      case 5: d[i] = (i % 337 < 200) ? static_cast<uint8_t>(40 + rng() % 30) : static_cast<uint8_t>(i % 337); break;   // record-like
      default: d[i] = static_cast<uint8_t>(rng() & 15); break;
    }
  }
  return d;
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? std::atoi(argv[1]) : 1500;
  std::mt19937 rng(7);
  fgb::inflate::Consts k;
  fgb::inflate::consts_init(k);
  fgb::inflate::Tables* t = new fgb::inflate::Tables;
  fgb::fastdeflate::Scratch* S = new fgb::fastdeflate::Scratch;
  const size_t sizes[] = {0, 1, 2, 15, 16, 17, 100, 1000, 40000, 65279, 65280};
  int checked = 0;
  for (int trial = 0; trial < iters; ++trial) {
    const int kind = trial % 7;
    const size_t n = sizes[rng() % (sizeof(sizes) / sizeof(sizes[0]))];
    const std::vector<uint8_t> d = shape(rng, kind, n);
    // ---- the own encoder -> zlib and the own decoder ----
    {
      std::vector<uint8_t> c(n + 64);                          // exactly the capacity the encoder is promised
      const size_t cl = fgb::fastdeflate::deflate_block(*S, d.data(), static_cast<uint32_t>(n), c.data(), c.size());
      if (cl == 0 || cl > n + 5) { std::printf("encoder size %zu for %zu\n", cl, n); return 1; }
      std::vector<uint8_t> o(n ? n : 1);
      z_stream zs;
      std::memset(&zs, 0, sizeof(zs));
      inflateInit2(&zs, -15);
      zs.next_in = c.data(); zs.avail_in = static_cast<uInt>(cl); zs.next_out = o.data(); zs.avail_out = static_cast<uInt>(n);
      const int rc = inflate(&zs, Z_FINISH);
      inflateEnd(&zs);
      if (!(rc == Z_STREAM_END || (n == 0 && (rc == Z_OK || rc == Z_BUF_ERROR))) || zs.avail_out != 0 || (n && std::memcmp(o.data(), d.data(), n))) {
        std::printf("zlib rejects the encoder's block: trial %d kind %d n %zu rc %d\n", trial, kind, n, rc); return 1;
      }
      std::vector<uint8_t> o2(n);
      const uint32_t st = fgb::inflate::inflate_member(c.data(), static_cast<uint32_t>(cl), o2.data(), static_cast<uint32_t>(n), *t, k);
      if (st != 0 || (n && std::memcmp(o2.data(), d.data(), n))) { std::printf("own decoder rejects the encoder's block: trial %d st %u\n", trial, st); return 1; }
      if (fgb::fastdeflate::crc32(d.data(), n) != crc32(crc32(0, nullptr, 0), d.data(), static_cast<uInt>(n))) { std::printf("crc\n"); return 1; }
    }
    // ---- zlib -> the own decoder, good / damaged / truncated ----
    {
      const int level = static_cast<int>(rng() % 10);
      const int strategy = trial % 11 == 0 ? Z_FIXED : (trial % 13 == 0 ? Z_HUFFMAN_ONLY : Z_DEFAULT_STRATEGY);
      z_stream zs;
      std::memset(&zs, 0, sizeof(zs));
      deflateInit2(&zs, level, Z_DEFLATED, -15, 8, strategy);
      std::vector<uint8_t> c(n + n / 8 + 128);
      zs.next_in = const_cast<Bytef*>(d.data()); zs.avail_in = static_cast<uInt>(n); zs.next_out = c.data(); zs.avail_out = static_cast<uInt>(c.size());
      if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { std::printf("zlib deflate failed\n"); return 1; }
      const size_t cl = c.size() - zs.avail_out;
      deflateEnd(&zs);
      std::vector<uint8_t> tight(c.begin(), c.begin() + cl);   // no slack behind the stream: ASan sees an over-read
      std::vector<uint8_t> o(n);
      const uint32_t st = fgb::inflate::inflate_member(tight.data(), static_cast<uint32_t>(cl), o.data(), static_cast<uint32_t>(n), *t, k);
      if (st != 0 || (n && std::memcmp(o.data(), d.data(), n))) { std::printf("own decoder fails on zlib level %d: trial %d kind %d n %zu st %u\n", level, trial, kind, n, st); return 1; }
      if (cl > 4) {
        std::vector<uint8_t> bad(tight);
        for (int f = 0; f < 3; ++f) bad[rng() % cl] ^= static_cast<uint8_t>(1u << (rng() % 8));
        std::vector<uint8_t> o3(n);
        (void)fgb::inflate::inflate_member(bad.data(), static_cast<uint32_t>(cl), o3.data(), static_cast<uint32_t>(n), *t, k);
        std::vector<uint8_t> half(tight.begin(), tight.begin() + cl / 2);
        const uint32_t s3 = fgb::inflate::inflate_member(half.data(), static_cast<uint32_t>(half.size()), o3.data(), static_cast<uint32_t>(n), *t, k);
        if (s3 == 0 && n > 100 && level != 0) { std::printf("a truncated stream passed: trial %d\n", trial); return 1; }
        // an output size that is too small / too large
        if (n > 1) {
          std::vector<uint8_t> o4(n - 1);
          if (fgb::inflate::inflate_member(tight.data(), static_cast<uint32_t>(cl), o4.data(), static_cast<uint32_t>(n - 1), *t, k) == 0) { std::printf("short output accepted\n"); return 1; }
          std::vector<uint8_t> o5(n + 1);
          if (fgb::inflate::inflate_member(tight.data(), static_cast<uint32_t>(cl), o5.data(), static_cast<uint32_t>(n + 1), *t, k) != fgb::inflate::kErrShort) { std::printf("long output not flagged\n"); return 1; }
        }
      }
    }
    ++checked;
  }
  std::printf("deflate_fuzz ok: %d cases\n", checked);
  delete t;
  delete S;
  return 0;
}

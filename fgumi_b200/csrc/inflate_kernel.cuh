// K0z -- BGZF members inflated on the device: one member per warp, the decoder's tables (1 044 bytes,
// inflate_core.h) in shared memory, four warps per CTA.  A member is an independent DEFLATE stream of at most 64 KiB of
// output (SAM spec 4.1), a batch of records is thousands of them: the host only frames the members
// (fgb_bgzf_scan_members: magic, BSIZE, CRC and ISIZE words), the link carries the compressed bytes, and the records
// appear in HBM where the row builder (unpack_kernels.cuh) reads them.  The role of the reference's fgumi-bgzf reader
// (crates/fgumi-bgzf/src/reader.rs); DESIGN.md section 8 has the numbers that make this the next step of a file-level
// run.  Every access is bounds-checked by the decoder: a corrupt member ends with a status byte, nothing else.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fgumi_b200.h"
#include "inflate_core.h"

namespace fgb {

// One decoder per WARP (lane 0 runs it; the other lanes wait): thirty-two decoders in one warp take data-dependent
// branches at every symbol and serialise each other (measured: 0.7 GB/s for 6 129 members, a warp per SM), one
// decoder per warp has the warp's issue slots to itself and ~40 of them fit an SM.
constexpr int kInflateWarps = 4;                              // decoders (warps) per CTA
constexpr int kInflateThreads = 32 * kInflateWarps;

struct InflateArgs {
  const uint8_t* in;                 // the compressed stream (device)
  const fgb_bgzf_member* members;    // device
  uint64_t n_members;
  uint8_t* out;                      // the inflated stream (device)
  uint8_t* status;                   // one byte per member: 0 = ok, else inflate::kErr*
  uint32_t check_crc;
  unsigned long long* n_bad;         // optional device counter of failed members
};

__global__ void __launch_bounds__(kInflateThreads) bgzf_inflate_kernel(const InflateArgs a) {
  __shared__ inflate::Tables tabs[kInflateWarps];
  __shared__ inflate::Consts k;
  __shared__ uint32_t crc_table[256];
  if (threadIdx.x == 0) inflate::consts_init(k);
  for (uint32_t i = threadIdx.x; i < 256u; i += kInflateThreads) inflate::crc_table_init(crc_table, i);
  __syncthreads();
  if ((threadIdx.x & 31u) != 0u) return;
  const uint32_t warp = threadIdx.x >> 5;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kInflateWarps;
  for (uint64_t m = static_cast<uint64_t>(blockIdx.x) * kInflateWarps + warp; m < a.n_members; m += stride) {
    const fgb_bgzf_member mem = a.members[m];
    uint8_t* const out = a.out + mem.out_off;
    uint32_t st = inflate::inflate_member(a.in + mem.in_off, mem.in_len, out, mem.out_len, tabs[warp], k);
    if (st == inflate::kOk && a.check_crc && inflate::crc32_bytes(crc_table, out, mem.out_len) != mem.crc) st = inflate::kErrCrc;
    a.status[m] = static_cast<uint8_t>(st);
    if (st != inflate::kOk && a.n_bad) atomicAdd(a.n_bad, 1ull);
  }
}

}  // namespace fgb

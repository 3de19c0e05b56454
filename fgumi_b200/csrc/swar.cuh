// Byte-parallel helpers shared by the kernels.
#pragma once
#include <stdint.h>

namespace fgb {

// 0xFF in every byte whose high bit is set, 0x00 elsewhere: PRMT with sign-replicating selectors (the
// __byte_perm intrinsic masks the replicate bit away, hence the PTX).
__device__ __forceinline__ uint32_t spread_msb(uint32_t x) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %1, 0xBA98;" : "=r"(r) : "r"(x));
  return r;
}

}  // namespace fgb

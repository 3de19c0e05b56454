// Compile-time geometry shared by the tile planner (host) and the vote kernel (device).
#pragma once
#include <stdint.h>

namespace fgb {

// One CTA = 256 threads, 2 CTAs per SM.  A tile is a run of consecutive units whose byte range in
// each column fits one shared-memory stage.
constexpr int kConsumerWarps = 8;
constexpr int kVoteThreads = kConsumerWarps * 32;        // threads that vote
constexpr int kThreads = kVoteThreads + 32;              // + one TMA producer warp
constexpr int kStages = 2;
constexpr uint32_t kTileCapBytes = 20480;  // per column per stage
constexpr uint32_t kTileMaxReads = 512;    // read descriptors per stage (8 B each)
constexpr uint32_t kTileMaxUnits = 127;    // unit descriptors per stage (16 B each, +1 sentinel)
#ifndef FGB_WARP_QUEUE_CAP
#define FGB_WARP_QUEUE_CAP 96
#endif
constexpr uint32_t kWarpQueueCap = FGB_WARP_QUEUE_CAP;     // undecided positions buffered per warp per tile
constexpr uint32_t kQtEntries = 256;       // fast-path quality threshold table, indexed by min(n,255)

constexpr uint32_t kTileFlagDirect = 1u;   // unit too large for a stage: kernel reads it from HBM
constexpr uint32_t kTileFlagRegular = 2u;  // all reads one length L, rows back to back at stride
                                           // round_up(L,8), every unit calls L positions (L > 8)
constexpr uint32_t kTileFlagSkew8 = 4u;    // regular tiles: first row starts 8 bytes into the stage
constexpr uint32_t kTileFlagShallow = 8u;  // no unit of the tile has more than 64 reads
// bits 8..31: items (8-position words) per unit when uniform over the tile (2..4096), else 0

}  // namespace fgb

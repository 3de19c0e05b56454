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
constexpr uint32_t kWarpQueueCap = 96;     // undecided positions buffered per warp per tile
constexpr uint32_t kQtEntries = 256;       // fast-path quality threshold table, indexed by min(n,255)

constexpr uint32_t kTileFlagDirect = 1u;   // unit too large for a stage: kernel reads it from HBM

}  // namespace fgb

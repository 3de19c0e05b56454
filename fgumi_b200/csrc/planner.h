// Tile planner (host, header-only): greedy segmentation of a run of units into shared-memory stages.
// Shared by fgb_plan_tiles (capi.cu) and the record-level callers, which plan the ranges their worker
// threads packed in parallel (csrc/host/caller_host.cpp).  Validates the layout rules of
// include/fgumi_b200.h while it walks the descriptors.
#pragma once
#include <algorithm>
#include <cstring>

#include "../../include/fgumi_b200.h"
#include "fgb_config.h"

namespace fgb {

// Plans units [u_begin, u_end) (units[u_end] must exist: the next unit or the sentinel).  A tile never
// spans the range's ends.  *prev_read_end carries the end of the last row seen (rows must ascend) in and
// out.  emit(const fgb_tile&) receives the tiles in order.
//
// `glue` (optional, one byte per unit index): glue[u] != 0 asks for unit u to share a tile with unit u - 1.  A run of
// glued units (a duplex molecule's single-strand units: fgb_plan_tiles_jobs) is placed as a whole when it fits an
// empty stage -- the open tile is closed early rather than cut through the run -- and its tile takes the run's
// common class, or the general class when the run mixes classes (every vote kernel is correct for any tile).  A run
// that does not fit a stage is planned unit by unit as if it were not glued.
template <class Emit>
inline fgb_status plan_tiles_range(const fgb_unit* units, uint64_t u_begin, uint64_t u_end,
                                   const fgb_read_desc* reads, uint64_t n_reads, uint64_t* prev_read_end_io,
                                   Emit&& emit_tile, const uint8_t* glue = nullptr) {
  fgb_tile cur{};
  bool open = false;
  uint64_t cur_end = 0;   // exclusive end (unaligned) of the open tile's byte range
  uint64_t prev_read_end = *prev_read_end_io;

  uint32_t cur_items = 0;     // 8-position items per unit if uniform so far, 0xFFFFFFFF = mixed
  bool regular = false;       // open tile: equal-length rows packed at stride round_up(len, 8)
  uint32_t reg_len = 0;
  uint64_t reg_next = 0;      // where the next row must start for the tile to stay regular
  uint32_t max_reads_in_unit = 0;
  uint32_t cur_class = 0;     // class of the open tile (fgb_config.h unit_class)
  auto emit = [&]() {
    cur.byte_len = static_cast<uint32_t>(((cur_end + 15u) & ~15ull) - cur.byte_begin);
    if (cur.flags & kTileFlagDirect) { cur.byte_len = 0; }
    const bool uniform = cur_items != 0xFFFFFFFFu && cur_items >= 2 && cur_items <= 4096;   // umulhi exactness
    if (uniform) cur.flags |= cur_items << 8;    // hint: unit index = item / cur_items
    if (!(cur.flags & kTileFlagDirect)) {
      if (regular && uniform && cur.n_reads > 0) {
        cur.flags |= kTileFlagRegular;
        if (FGB_READ_OFF(reads[cur.read_begin]) != cur.byte_begin) cur.flags |= kTileFlagSkew8;
      }
      if (max_reads_in_unit <= 64) cur.flags |= kTileFlagShallow;
    }
    cur.flags |= cur_class << kTileClassShift;
    emit_tile(cur);
    open = false;
  };

  // the glued run unit u belongs to (glue only)
  uint64_t run_begin = u_begin, run_end = u_begin;
  bool run_atomic = false;
  uint32_t run_class = 0, run_reads = 0;
  uint64_t run_byte_end = 0;

  for (uint64_t u = u_begin; u < u_end; ++u) {
    const fgb_unit& un = units[u];
    const fgb_unit& nx = units[u + 1];
    if (glue && u == run_end) {                       // a new run starts here: measure it
      run_begin = u;
      uint64_t v = u + 1;
      while (v < u_end && glue[v]) ++v;
      run_end = v;
      run_atomic = false;
      if (v - u > 1) {
        uint64_t gb = 0, ge = 0;
        bool any = false, mixed = false;
        uint32_t cls = 0;
        uint64_t total = 0;
        for (uint64_t w = u; w < v; ++w) {
          const uint32_t r0 = units[w].read_begin, r1 = units[w + 1].read_begin;
          if (r1 < r0 || r1 > n_reads) return FGB_ERR_LAYOUT;
          const uint32_t c = unit_class(r1 - r0);
          if (w == u) cls = c; else if (c != cls) mixed = true;
          total += r1 - r0;
          if (r1 > r0) {
            if (!any) { gb = FGB_READ_OFF(reads[r0]); any = true; }
            ge = FGB_READ_OFF(reads[r1 - 1]) + FGB_READ_LEN(reads[r1 - 1]);
          }
        }
        const uint64_t span = any && ge >= gb ? ((ge + 15u) & ~15ull) - (gb & ~15ull) : 0;
        run_atomic = span <= kTileCapBytes && v - u <= kTileMaxUnits && total + 1 <= kTileMaxReads;
        run_class = mixed ? 0u : cls;
        run_reads = static_cast<uint32_t>(total);
        run_byte_end = ge;
      }
    }
    const bool in_run = glue && run_atomic;           // this unit is placed with its run
    if (nx.read_begin < un.read_begin || nx.read_begin > n_reads) return FGB_ERR_LAYOUT;
    uint32_t nr = nx.read_begin - un.read_begin;
    if (un.out_off % FGB_OUT_ALIGN) return FGB_ERR_LAYOUT;
    if (nx.out_off != un.out_off + ((static_cast<uint64_t>(un.cons_len) + (FGB_OUT_ALIGN - 1u)) & ~static_cast<uint64_t>(FGB_OUT_ALIGN - 1u)))
      return FGB_ERR_LAYOUT;   // output rows are dense, each padded to FGB_OUT_ALIGN
    if (un.cons_len > FGB_MAX_READ_LEN) return FGB_ERR_UNIT_TOO_LARGE;
    if (nr == 0 && un.cons_len != 0) return FGB_ERR_LAYOUT;
    if (nr > 0xFFFFu) return FGB_ERR_UNIT_TOO_LARGE;   // u16 observation counters, base_builder.rs:236
    uint64_t ub = 0, ue = 0;   // byte range of this unit
    uint32_t maxlen = 0;
    for (uint32_t r = un.read_begin; r < nx.read_begin; ++r) {
      uint64_t off = FGB_READ_OFF(reads[r]);
      uint32_t len = FGB_READ_LEN(reads[r]);
      if (off % FGB_READ_ALIGN || len == 0) return FGB_ERR_LAYOUT;   // no empty rows
      if (off < prev_read_end) return FGB_ERR_LAYOUT;   // rows ascend and do not overlap
      prev_read_end = off + len;
      if (r == un.read_begin) ub = off;
      ue = off + len;
      maxlen = std::max(maxlen, len);
    }
    if (un.cons_len > maxlen) return FGB_ERR_LAYOUT;
    if (nr == 0) { ub = ue = open ? cur_end : prev_read_end; }

    // Can the unit join the open tile?  (The first unit of an atomic run asks for the whole run; its other units
    // follow it whatever their class -- the stage limits are still checked: a run measured too kindly is cut.)
    if (open) {
      const bool follower = in_run && u > run_begin;
      const bool leader = in_run && !follower;
      const uint64_t end = leader ? std::max(ue, run_byte_end) : ue;
      const uint32_t want_reads = leader ? run_reads : nr;
      const uint64_t want_units = leader ? run_end - run_begin : 1;
      const uint32_t cls = follower ? cur_class : (leader ? run_class : unit_class(nr));
      uint64_t span = ((end + 15u) & ~15ull) - cur.byte_begin;
      uint32_t skew = cur.read_begin & 1u;
      bool fits = !(cur.flags & kTileFlagDirect) && span <= kTileCapBytes && cls == cur_class &&
                  cur.n_units + want_units <= kTileMaxUnits &&
                  cur.n_reads + want_reads + skew <= kTileMaxReads;
      if (!fits) emit();
    }
    if (!open) {
      std::memset(&cur, 0, sizeof(cur));
      cur.byte_begin = ub & ~15ull;
      cur.unit_begin = static_cast<uint32_t>(u);
      cur.read_begin = un.read_begin;
      cur_end = ub;
      open = true;
      uint64_t span = ((ue + 15u) & ~15ull) - cur.byte_begin;
      if (span > kTileCapBytes || nr + (cur.read_begin & 1u) > kTileMaxReads)
        cur.flags |= kTileFlagDirect;   // oversize unit: kernel votes it straight from HBM
      regular = nr > 0;
      reg_len = nr ? FGB_READ_LEN(reads[un.read_begin]) : 0;
      reg_next = ub;
      max_reads_in_unit = 0;
      cur_class = in_run ? run_class : unit_class(nr);
    }
    if (regular) {   // still regular with this unit?
      const uint64_t stride = (static_cast<uint64_t>(reg_len) + 7u) & ~7ull;
      if (nr == 0 || un.cons_len != reg_len) regular = false;
      for (uint32_t r = un.read_begin; regular && r < nx.read_begin; ++r) {
        if (FGB_READ_LEN(reads[r]) != reg_len || FGB_READ_OFF(reads[r]) != reg_next) regular = false;
        reg_next += stride;
      }
    }
    max_reads_in_unit = std::max(max_reads_in_unit, nr);
    {
      uint32_t items = (un.cons_len + 7u) >> 3;
      if (cur.n_units == 0) cur_items = items;
      else if (cur_items != items) cur_items = 0xFFFFFFFFu;
    }
    cur.n_units += 1;
    cur.n_reads += nr;
    cur_end = std::max(cur_end, ue);
    if (cur.flags & kTileFlagDirect) emit();
  }
  if (open) emit();
  *prev_read_end_io = prev_read_end;
  return FGB_OK;
}

}  // namespace fgb

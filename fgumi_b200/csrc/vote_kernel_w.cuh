// K1w — the vote kernel, second work decomposition: ONE WARP PER UNIT, lanes across the unit's
// 8-position items, rows (reads) walked in sequence; units deeper than kSliceMin reads are voted by
// all eight consumer warps at once, each taking a slice of the rows, and the warp that arrives last
// folds the eight partial results.
//
// Same algorithm, tables, proofs, stages and TMA producer as vote_kernel (vote_kernel.cuh, which
// cites the reference lines); what changes is who does what:
//   * vote_kernel deals items (8 positions of one unit) to threads: every thread looks its unit up,
//     loads descriptors, walks the whole depth axis on its own.  Cheap per byte at depth 6..20, but a
//     deep unit is a tile of 19..38 items (most of the CTA idles while a few threads walk 50..100
//     reads), and the per-item bookkeeping is paid 19 times per unit;
//   * here a warp owns a unit: lane l holds item l (a 150 bp unit uses 19 lanes), the unit's
//     descriptors are read once per warp (broadcast), each row is one conflict-free 8-byte access per
//     lane, and the undecided positions of the units a warp finished are resolved by that warp.
//     Two-read units take the pair table in line (one lookup per position), three- and four-read
//     units use a sum-of-qualities proof of the reference's fast path instead of the per-read
//     minimum (host_tables.cpp: sum thresholds by dynamic programming over the likelihood-gap table).
#pragma once
#include "vote_kernel.cuh"

namespace fgb {

constexpr uint32_t kSliceMin = 24;   // units with more reads than this are voted by all eight warps

struct __align__(128) VoteSmemW {
  VoteSmem v;                                        // same layout first: the shared helpers take VoteSmem&
  uint32_t part[2][kConsumerWarps][5][32];           // [buffer][warp][word][lane] partial results of a sliced unit
  uint32_t arrive[2];                                // arrivals on a buffer, monotonic (8 per use)
  uint32_t released[2];                              // uses of a buffer whose partials have been folded
  uint16_t sumt[8];                                  // sum-of-qualities thresholds by depth (3, 4 used)
};

// 0x80 in every byte of `sum` (bytes <= 255) that is >= t (t <= 255)
__device__ __forceinline__ uint32_t bytes_ge(uint32_t sum, uint32_t t) {
  const uint32_t add = (0x100u - t) * 0x00010001u;
  const uint32_t ev = ((sum & 0x00FF00FFu) + add) & 0x01000100u;          // bytes 0, 2 -> bits 8, 24
  const uint32_t od = (((sum >> 8) & 0x00FF00FFu) + add) & 0x01000100u;   // bytes 1, 3
  return (ev >> 1) | (od << 7);
}

template <class M, bool Regular>
__device__ __forceinline__ void vote_tile_w(const VoteArgs& a, VoteSmemW& SW, const Stage& st,
                                            const TileView<M>& tv, uint32_t warp, uint32_t lane,
                                            uint32_t& rot, uint32_t& gen, LocalStats& ls) {
  VoteSmem& S = SW.v;
  const uint32_t n_units = st.tile.n_units;
  const uint32_t min_reads = a.min_reads, min_cons_q = a.min_cons_q, fast_qual = a.fast_qual;
  const bool fast_masked = fast_qual < min_cons_q;
  const uint32_t fq4 = (fast_masked ? 2u : fast_qual) * 0x01010101u;
  const uint32_t base32 = static_cast<uint32_t>(tv.byte_base);
  uint32_t* const wqueue = S.queue[warp];
  uint32_t* const wcount = &S.q_count[warp];
  // regular tiles: every row has length reg_len and sits at stride uni_m * 8 from the tile's first row
  const uint32_t uni_m = st.tile.flags >> 8;
  const uint32_t reg_len = st.units[0].cons_len;
  const uint32_t reg_row0 = (st.tile.flags & kTileFlagSkew8) ? 8u : 0u;

  for (uint32_t u = 0; u < n_units; ++u) {
    const fgb_unit un = st.units[u];
    const uint32_t rb = un.read_begin;
    const uint32_t n = st.units[u + 1].read_begin - rb;
    const uint32_t cons_len = un.cons_len;
    const bool sliced = n > kSliceMin;
    uint32_t r0 = 0, r1 = n;
    if (!sliced) {
      const bool mine = ((rot++) & (kConsumerWarps - 1u)) == warp;
      if (!mine) continue;
    } else {
      r0 = (n * warp) >> 3;
      r1 = (n * (warp + 1u)) >> 3;
    }
    const uint32_t items = (cons_len + 7u) >> 3;
    const uint32_t row_first = rb - tv.read_base;          // index of the unit's first row in the tile

    for (uint32_t ib = 0; ib < items; ib += 32u) {
      const uint32_t item = ib + lane;
      const bool active = item < items;
      const uint32_t p0 = active ? item << 3 : 0u;          // idle lanes shadow item 0; nothing of theirs is stored
      const uint32_t real = active ? (cons_len - p0 < 8u ? cons_len - p0 : 8u) : 0u;
      const uint32_t rm_lo = low_bytes_mask(real), rm_hi = low_bytes_mask(real > 4u ? real - 4u : 0u);
      const uint64_t o = un.out_off + p0;

      // address of this lane's word in row r of the unit, and the row's length
      auto row_at = [&](uint32_t r, uint32_t* len) -> typename M::off_t {
        if (Regular) {
          *len = reg_len;
          return static_cast<typename M::off_t>(reg_row0 + ((row_first + r) * uni_m << 3) + p0);
        }
        const uint64_t d = M::ld64(tv.reads + static_cast<typename M::off_t>(row_first + r) * 8u);
        const uint32_t l = static_cast<uint32_t>(d) & 0xFFFFu;
        *len = l;
        return M::row_offset(d, tv.byte_base, base32) + (l > p0 ? p0 : 0u);   // an uncovered row points at its own start
      };

      uint32_t wb_lo = 0, wb_hi = 0, wq_lo = 0, wq_hi = 0;
      uint4 dep = make_uint4(0, 0, 0, 0);
      const uint4 err = make_uint4(0, 0, 0, 0);
      uint32_t todo_lo = 0, todo_hi = 0;

      if (n == 1u) {
        // single-read consensus, vanilla_caller.rs:1285-1316
        uint32_t len;
        const typename M::off_t row = row_at(0, &len);
        uint64_t rbw = 0, rqw = 0;
        if (p0 < len) { rbw = M::ld64(tv.bases + row); rqw = M::ld64(tv.quals + row); }
        uint64_t obw = 0, oqw = 0, odw_lo = 0, odw_hi = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t pos = p0 + j;
          if (static_cast<uint32_t>(j) < real) {
            const uint32_t b = static_cast<uint32_t>(rbw >> (8 * j)) & 0xFFu;
            const uint32_t q = static_cast<uint32_t>(rqw >> (8 * j)) & 0xFFu;
            uint32_t ob = 'N', oq = 2, od = 0;
            if (pos < len) {
              const uint32_t adj = q < FGB_NTABLE ? S.single_q[q] : 0u;   // `.get(idx).unwrap_or(0)`
              if (adj >= min_cons_q) { ob = b; oq = adj; }
              od = (b != 'N');
            }
            obw |= static_cast<uint64_t>(ob) << (8 * j);
            oqw |= static_cast<uint64_t>(oq) << (8 * j);
            if (j < 4) odw_lo |= static_cast<uint64_t>(od) << (16 * j);
            else odw_hi |= static_cast<uint64_t>(od) << (16 * (j - 4));
            ls.nocall += (ob == 'N');
          }
        }
        wb_lo = static_cast<uint32_t>(obw); wb_hi = static_cast<uint32_t>(obw >> 32);
        wq_lo = static_cast<uint32_t>(oqw); wq_hi = static_cast<uint32_t>(oqw >> 32);
        dep = make_uint4(static_cast<uint32_t>(odw_lo), static_cast<uint32_t>(odw_lo >> 32),
                         static_cast<uint32_t>(odw_hi), static_cast<uint32_t>(odw_hi >> 32));
      } else if (n == 2u && min_reads <= 2u) {
        // two-read units: where both reads cover the position and agree on an A/C/G/T base, the result
        // is the host-evaluated outcome of the reference's add / add / call sequence (host_tables.cpp
        // pair_quality); everything else is queued for the literal path.
        uint32_t l0, l1;
        const typename M::off_t ra = row_at(0, &l0), rc = row_at(1, &l1);
        const uint64_t b0w = M::ld64(tv.bases + ra), b1w = M::ld64(tv.bases + rc);
        const uint64_t q0w = M::ld64(tv.quals + ra), q1w = M::ld64(tv.quals + rc);
        const uint32_t ml = l0 < l1 ? l0 : l1;
        const uint32_t covered = ml > p0 ? ml - p0 : 0u;
        const uint32_t b0_lo = static_cast<uint32_t>(b0w), b0_hi = static_cast<uint32_t>(b0w >> 32);
        const uint32_t eq_lo = zero_bytes(b0_lo ^ static_cast<uint32_t>(b1w)) & acgt_bytes(b0_lo) &
                               low_bytes_mask(covered) & rm_lo;
        const uint32_t eq_hi = zero_bytes(b0_hi ^ static_cast<uint32_t>(b1w >> 32)) & acgt_bytes(b0_hi) &
                               low_bytes_mask(covered > 4u ? covered - 4u : 0u) & rm_hi;
        uint64_t obw = 0, oqw = 0;
        uint32_t ok_lo = 0, ok_hi = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t bit = ((j < 4 ? eq_lo >> (8 * j) : eq_hi >> (8 * (j - 4))) >> 7) & 1u;
          if (bit) {
            uint32_t qa = static_cast<uint32_t>(q0w >> (8 * j)) & 0xFFu, qb = static_cast<uint32_t>(q1w >> (8 * j)) & 0xFFu;
            qa = qa > FGB_MAX_PHRED ? FGB_MAX_PHRED : qa;
            qb = qb > FGB_MAX_PHRED ? FGB_MAX_PHRED : qb;
            const uint32_t cq = __ldg(S.pair_q + qa * 94u + qb);
            if (cq != 255u) {
              const bool masked = cq < min_cons_q;                             // vanilla_caller.rs:1347-1348
              const uint32_t ob = masked ? 'N' : (static_cast<uint32_t>(b0w >> (8 * j)) & 0xFFu);
              obw |= static_cast<uint64_t>(ob) << (8 * j);
              oqw |= static_cast<uint64_t>(masked ? 2u : cq) << (8 * j);
              ls.nocall += masked;
              if (j < 4) ok_lo |= 0x80u << (8 * j); else ok_hi |= 0x80u << (8 * (j - 4));
            }
          }
        }
        wb_lo = static_cast<uint32_t>(obw); wb_hi = static_cast<uint32_t>(obw >> 32);
        wq_lo = static_cast<uint32_t>(oqw); wq_hi = static_cast<uint32_t>(oqw >> 32);
        const uint32_t fb_lo = spread_msb(ok_lo), fb_hi = spread_msb(ok_hi);   // 0x80 -> 0xFF
        dep.x = 0x00020002u & __byte_perm(fb_lo, 0u, 0x1100u);
        dep.y = 0x00020002u & __byte_perm(fb_lo, 0u, 0x3322u);
        dep.z = 0x00020002u & __byte_perm(fb_hi, 0u, 0x1100u);
        dep.w = 0x00020002u & __byte_perm(fb_hi, 0u, 0x3322u);
        todo_lo = rm_lo & ~ok_lo; todo_hi = rm_hi & ~ok_hi;
      } else {
        // ---- general pileup: SWAR proof of the unanimous fast path over this warp's rows ----
        const uint32_t qt = S.qt[n < kQtEntries ? n : kQtEntries - 1];
        const uint32_t sumt = (n == 3u || n == 4u) ? SW.sumt[n] : 0xFFFFu;
        const bool by_sum = sumt <= 255u;
        const bool fast_ok = (by_sum || qt <= FGB_MAX_PHRED) && (n >= min_reads) && (n <= 0xFFFFu);
        uint32_t b0_lo = 0, b0_hi = 0;
        uint32_t diff_lo = 0, diff_hi = 0, okq_lo = 0x80808080u, okq_hi = 0x80808080u, minlen = 0xFFFFFFFFu;
        if (fast_ok) {
          {
            uint32_t l;
            const typename M::off_t row = row_at(0, &l);
            const uint64_t w = M::ld64(tv.bases + row);
            b0_lo = static_cast<uint32_t>(w); b0_hi = static_cast<uint32_t>(w >> 32);
          }
          if (by_sum) {
            // three / four reads: every quality in 1..63 and the SUM of the qualities at or above the
            // threshold proves sum D[q_i] > min(23, G2) (host_tables.cpp, exact dynamic programme)
            uint32_t s_lo = 0, s_hi = 0, or_lo = 0, or_hi = 0, nz_lo = 0x40404040u, nz_hi = 0x40404040u;
            for (uint32_t r = 0; r < n; ++r) {
              uint32_t l;
              const typename M::off_t row = row_at(r, &l);
              minlen = l < minlen ? l : minlen;
              const uint64_t wb = M::ld64(tv.bases + row);
              const uint64_t wq = M::ld64(tv.quals + row);
              diff_lo |= static_cast<uint32_t>(wb) ^ b0_lo;
              diff_hi |= static_cast<uint32_t>(wb >> 32) ^ b0_hi;
              const uint32_t ql = static_cast<uint32_t>(wq), qh = static_cast<uint32_t>(wq >> 32);
              s_lo += ql; s_hi += qh;                      // bytes stay <= 4 * 63 while or_* proves q < 64
              or_lo |= ql; or_hi |= qh;
              nz_lo &= ql + 0x3F3F3F3Fu; nz_hi &= qh + 0x3F3F3F3Fu;   // bit 6 survives iff every q >= 1 (q < 64: no carries)
            }
            // a quality >= 64 anywhere in a word voids the whole word (its carry may have touched a neighbour)
            okq_lo = (or_lo & 0xC0C0C0C0u) ? 0u : ((nz_lo << 1) & bytes_ge(s_lo, sumt));
            okq_hi = (or_hi & 0xC0C0C0C0u) ? 0u : ((nz_hi << 1) & bytes_ge(s_hi, sumt));
          } else {
            const uint32_t tsplat = qt * 0x01010101u;
#pragma unroll 4
            for (uint32_t r = r0; r < r1; ++r) {
              uint32_t l;
              const typename M::off_t row = row_at(r, &l);
              minlen = l < minlen ? l : minlen;
              const uint64_t wb = M::ld64(tv.bases + row);
              const uint64_t wq = M::ld64(tv.quals + row);
              diff_lo |= static_cast<uint32_t>(wb) ^ b0_lo;
              diff_hi |= static_cast<uint32_t>(wb >> 32) ^ b0_hi;
              // byte high bit survives iff q >= qT (no borrows: every minuend byte is >= 0x80 > qT)
              okq_lo &= (static_cast<uint32_t>(wq) | 0x80808080u) - tsplat;
              okq_hi &= (static_cast<uint32_t>(wq >> 32) | 0x80808080u) - tsplat;
            }
          }
        }
        if (sliced) {
          // ---- fold the eight slices: the warp that arrives last carries on, the others move on ----
          const uint32_t g = gen++;
          const uint32_t buf = g & 1u, use = g >> 1;
          if (lane == 0) {
            while (*reinterpret_cast<volatile uint32_t*>(&SW.released[buf]) < use) __nanosleep(32);
          }
          __syncwarp();
          SW.part[buf][warp][0][lane] = diff_lo; SW.part[buf][warp][1][lane] = diff_hi;
          SW.part[buf][warp][2][lane] = okq_lo;  SW.part[buf][warp][3][lane] = okq_hi;
          SW.part[buf][warp][4][lane] = minlen;
          __threadfence_block();
          __syncwarp();
          uint32_t last = 0;
          if (lane == 0) last = atomicAdd(&SW.arrive[buf], 1u) == use * kConsumerWarps + (kConsumerWarps - 1u);
          last = __shfl_sync(0xFFFFFFFFu, last, 0);
          if (!last) continue;
          __threadfence_block();
#pragma unroll
          for (uint32_t w = 0; w < static_cast<uint32_t>(kConsumerWarps); ++w) {
            diff_lo |= SW.part[buf][w][0][lane]; diff_hi |= SW.part[buf][w][1][lane];
            okq_lo &= SW.part[buf][w][2][lane];  okq_hi &= SW.part[buf][w][3][lane];
            const uint32_t ml = SW.part[buf][w][4][lane];
            minlen = ml < minlen ? ml : minlen;
          }
          __syncwarp();
          if (lane == 0) {
            __threadfence_block();
            *reinterpret_cast<volatile uint32_t*>(&SW.released[buf]) = use + 1u;
          }
        }
        uint32_t fm_lo = 0, fm_hi = 0;
        if (fast_ok) {
          // per-byte verdict: unanimous & quality-proven & A/C/G/T & covered by every read
          const uint32_t covered = minlen > p0 ? minlen - p0 : 0u;
          fm_lo = zero_bytes(diff_lo) & okq_lo & acgt_bytes(b0_lo) & low_bytes_mask(covered) & rm_lo;
          fm_hi = zero_bytes(diff_hi) & okq_hi & acgt_bytes(b0_hi) &
                  low_bytes_mask(covered > 4u ? covered - 4u : 0u) & rm_hi;
        }
        // proven positions: constant quality, depth n, no errors
        const uint32_t fb_lo = spread_msb(fm_lo), fb_hi = spread_msb(fm_hi);
        wb_lo = (fast_masked ? 0x4E4E4E4Eu : b0_lo) & fb_lo;
        wb_hi = (fast_masked ? 0x4E4E4E4Eu : b0_hi) & fb_hi;
        wq_lo = fq4 & fb_lo;
        wq_hi = fq4 & fb_hi;
        const uint32_t nn = n | (n << 16);
        dep.x = nn & __byte_perm(fb_lo, 0u, 0x1100u);
        dep.y = nn & __byte_perm(fb_lo, 0u, 0x3322u);
        dep.z = nn & __byte_perm(fb_hi, 0u, 0x1100u);
        dep.w = nn & __byte_perm(fb_hi, 0u, 0x3322u);
        ls.nocall += fast_masked ? (__popc(fm_lo) + __popc(fm_hi)) : 0;
        todo_lo = rm_lo & ~fm_lo; todo_hi = rm_hi & ~fm_hi;
      }
      ls.positions += real;
      if (active) {
        *reinterpret_cast<uint2*>(a.out_base + o) = make_uint2(wb_lo, wb_hi);
        *reinterpret_cast<uint2*>(a.out_qual + o) = make_uint2(wq_lo, wq_hi);
        *reinterpret_cast<uint4*>(a.out_depth + o) = dep;
        *reinterpret_cast<uint4*>(a.out_errors + o) = err;
      }
      if (todo_lo | todo_hi) {
        // undecided positions go to this warp's queue; what does not fit is resolved in place
        const uint32_t cnt = static_cast<uint32_t>(__popc(todo_lo) + __popc(todo_hi));
        uint32_t slot = atomicAdd(wcount, cnt);
        const uint32_t ent = (u << 16) | p0;
        uint32_t keep_lo = 0, keep_hi = 0;
        for (uint32_t t = todo_lo; t; t &= t - 1u, ++slot) {
          if (slot < kWarpQueueCap) wqueue[slot] = ent + ((__ffs(t) - 1) >> 3);
          else keep_lo |= t & (0u - t);
        }
        for (uint32_t t = todo_hi; t; t &= t - 1u, ++slot) {
          if (slot < kWarpQueueCap) wqueue[slot] = ent + 4u + ((__ffs(t) - 1) >> 3);
          else keep_hi |= t & (0u - t);
        }
        for (uint32_t j = 0; j < 8u; ++j) {
          if ((j < 4u ? keep_lo >> (8u * j) : keep_hi >> (8u * (j - 4u))) & 0x80u) {
            const Called c = resolve_position<M>(tv, S, rb, n, p0 + j, a, ls);
            write_called(a, o + j, c);
          }
        }
      }
      __syncwarp();
      // drain before the queue can overflow on the next block (an item adds at most 8 per lane)
      uint32_t qn = *wcount;
      if (qn >= kWarpQueueCap / 2u) {
        qn = qn < kWarpQueueCap ? qn : kWarpQueueCap;
        slow_pass<M>(a, S, st, tv, wqueue, qn, lane, ls);
        __syncwarp();
        if (lane == 0) *wcount = 0;
        __syncwarp();
      }
    }
  }
  __syncwarp();
  uint32_t qn = *wcount;
  qn = qn < kWarpQueueCap ? qn : kWarpQueueCap;
  if (qn) slow_pass<M>(a, S, st, tv, wqueue, qn, lane, ls);
  __syncwarp();
  if (lane == 0) *wcount = 0;
}

__global__ void __launch_bounds__(kThreads) vote_kernel_w(const VoteArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  VoteSmemW& SW = *reinterpret_cast<VoteSmemW*>(smem_raw);
  VoteSmem& S = SW.v;
  const uint32_t tid = threadIdx.x;
  const uint32_t warp = tid >> 5;

  for (uint32_t i = tid; i < FGB_NTABLE; i += kThreads) {
    S.correct[i] = a.tables->correct[i];
    S.err_alt[i] = a.tables->err_alt[i];
  }
  for (uint32_t i = tid; i < 96; i += kThreads) S.single_q[i] = a.tables->single_q[i];
  for (uint32_t i = tid; i < kQtEntries; i += kThreads) S.qt[i] = a.tables->qt[i];
  for (uint32_t i = tid; i < 96; i += kThreads) S.dfix[i] = a.tables->dfix[i];
  if (tid < kConsumerWarps) { S.q_count[tid] = 0; SW.sumt[tid] = a.tables->sumt[tid]; }
  if (tid == 0) {
    S.ln_pre = a.tables->ln_pre;
    S.g2fix = a.tables->g2fix;
    S.nmax2 = a.tables->nmax2;
    S.pair_q = a.tables->pair_q;
    SW.arrive[0] = SW.arrive[1] = 0;
    SW.released[0] = SW.released[1] = 0;
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&S.full[s], 1);
      mbar_init(&S.empty[s], kConsumerWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const uint32_t grid = gridDim.x;
  const uint32_t n_tiles = static_cast<uint32_t>(a.n_tiles);

  if (warp == kConsumerWarps) {
    // ================= PRODUCER WARP (as in vote_kernel) =================
    if ((tid & 31u) == 0) {
      uint32_t k = 0;
      uint64_t n_units_done = 0, n_reads_done = 0;
      for (uint32_t t = blockIdx.x; t < n_tiles; t += grid, ++k) {
        const int s = k % kStages;
        const uint32_t use = k / kStages;
        if (use > 0) mbar_wait(&S.empty[s], (use - 1u) & 1u, 20000u);
        Stage& st = S.st[s];
        const uint4* gt = reinterpret_cast<const uint4*>(a.tiles + t);
        const uint4 t0 = __ldg(gt), t1 = __ldg(gt + 1);
        *reinterpret_cast<uint4*>(&st.tile) = t0;
        *(reinterpret_cast<uint4*>(&st.tile) + 1) = t1;
        const uint64_t byte_begin = (static_cast<uint64_t>(t0.y) << 32) | t0.x;
        const uint32_t byte_len = t0.z, unit_begin = t0.w, n_units = t1.x, read_begin = t1.y,
                       n_reads = t1.z, flags = t1.w;
        const uint32_t units_bytes = (n_units + 1u) * 16u;
        const bool direct = (flags & kTileFlagDirect) != 0;
        const uint32_t rskew = read_begin & 1u;
        const uint32_t rbytes = ((n_reads + rskew + 1u) & ~1u) * 8u;
        const uint32_t tx = units_bytes + (direct ? 0u : 2u * byte_len + rbytes);
        n_units_done += n_units; n_reads_done += n_reads;
        mbar_arrive_expect_tx(&S.full[s], tx);
        tma_load_1d(st.units, a.units + unit_begin, units_bytes, &S.full[s]);
        if (!direct) {
          if (byte_len) {
            tma_load_1d(st.bases, a.bases + byte_begin, byte_len, &S.full[s]);
            tma_load_1d(st.quals, a.quals + byte_begin, byte_len, &S.full[s]);
          }
          if (rbytes) tma_load_1d(st.reads, a.reads + (read_begin - rskew), rbytes, &S.full[s]);
        }
      }
      if (n_units_done) atomicAdd(a.counters + FGB_CTR_UNITS, static_cast<unsigned long long>(n_units_done));
      if (n_reads_done) atomicAdd(a.counters + FGB_CTR_INPUT_READS, static_cast<unsigned long long>(n_reads_done));
    }
    return;
  }

  // ================= CONSUMER WARPS =================
  LocalStats ls = {0, 0, 0};
  const uint32_t lane = tid & 31u;
  uint32_t k = 0, rot = 0, gen = 0;
  for (uint32_t t = blockIdx.x; t < n_tiles; t += grid, ++k) {
    const int s = k % kStages;
    mbar_wait(&S.full[s], (k / kStages) & 1u, 2000u);
    const Stage& st = S.st[s];
    if (st.tile.flags & kTileFlagDirect) {
      TileView<GlMem> tv;
      tv.bases = a.bases; tv.quals = a.quals;
      tv.reads = reinterpret_cast<const uint8_t*>(a.reads + st.tile.read_begin);
      tv.byte_base = 0; tv.read_base = st.tile.read_begin;
      vote_tile_w<GlMem, false>(a, SW, st, tv, warp, lane, rot, gen, ls);
    } else {
      TileView<ShMem> tv;
      tv.bases = st.bases; tv.quals = st.quals;
      tv.reads = reinterpret_cast<const uint8_t*>(st.reads) + (st.tile.read_begin & 1u) * 8u;
      tv.byte_base = st.tile.byte_begin; tv.read_base = st.tile.read_begin;
      if (st.tile.flags & kTileFlagRegular) vote_tile_w<ShMem, true>(a, SW, st, tv, warp, lane, rot, gen, ls);
      else vote_tile_w<ShMem, false>(a, SW, st, tv, warp, lane, rot, gen, ls);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.empty[s]);
  }

  uint32_t v0 = ls.positions, v1 = ls.exact, v2 = ls.nocall;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    v0 += __shfl_down_sync(0xFFFFFFFFu, v0, off);
    v1 += __shfl_down_sync(0xFFFFFFFFu, v1, off);
    v2 += __shfl_down_sync(0xFFFFFFFFu, v2, off);
  }
  if (lane == 0) {
    if (v0) atomicAdd(a.counters + FGB_CTR_POSITIONS, static_cast<unsigned long long>(v0));
    if (v1) atomicAdd(a.counters + FGB_CTR_EXACT_POSITIONS, static_cast<unsigned long long>(v1));
    if (v2) atomicAdd(a.counters + FGB_CTR_NOCALL_POSITIONS, static_cast<unsigned long long>(v2));
  }
}

}  // namespace fgb

// stem_kernel_b200/csrc/stem_tile.cu -- the fast stem kernel with the column sweep done on TILES of 16 rows.
//
// Same recurrence, same pre-scaled quantities and same phases A / B1 / C as stem_fast.cu (read its header first;
// reference: StemKernel<ST,D>::operator(), stem_kernel_lite/stem_kernel.cpp:14-95).  What changes is who does the
// column sweep B2.  In stem_fast.cu a warp sweeps the two rows of its own block with lanes <-> (row, node): a DAG
// level offers ~6 nodes at or above a row's band window, so 11 of 32 lanes work and two thirds of the kernel's
// shared-memory wavefronts serve a third of a wavefront's width.  Here the sweep has lanes <-> ROWS:
//   * a TILE is a [column j][16 rows] block of doubles in shared memory (row slot r of column j at
//     128 j + 8 (r ^ (j & 15)): a column's 16 rows are one 128-byte line in any order, and a row's entries of 16
//     consecutive columns hit 16 different 8-byte banks);
//   * a row worker -- any warp -- does phases A and B1 of ONE row in its private HQ buffer, is granted a row slot of
//     the tile that is filling, clears that slot's column entries and drops the row's MATCH values in (phase Z);
//   * when a tile is full (or has to be closed: see "progress") ONE warp sweeps it: half-warp <-> node, lane <-> row
//     slot, so the node record and its child list are uniform loads and every gather is a 128-byte line: a wavefront
//     carries 16 useful doubles.  Rows of a tile come from any pairs of the group and any sub-level: they only share y;
//   * the swept rows are drained one by one -- any warp -- into the pair's slab (phase C) and published.
// Work is handed out by a small scheduler in shared memory (one lock, touched a few times per row): drain tasks first,
// then sweeps, then new rows.  A warp that has to wait -- for the flags of a row's inner pairs, for a free row slot --
// runs those tasks meanwhile.
// Progress: a row waits only for rows with earlier tickets; such a row is being worked on, or sits in the filling tile
// -- which every waiting warp seals first, so it gets swept even if not full -- or in a tile that is closed, being
// swept or drained, which the waiting warps themselves serve.  Tickets exhausted: the filling tile is sealed.
#include "kernels.cuh"
#include "stem_common.cuh"

namespace stemk {

namespace {

constexpr uint32_t kGroup = kFastGroup;
constexpr uint32_t kTileRows = 16;
#ifndef STEMK_TILES
#define STEMK_TILES 2
#endif
constexpr uint32_t kTiles = STEMK_TILES;
constexpr uint32_t kNone = 0xffffffffu;
enum : uint32_t { T_FREE = 0, T_FILL, T_SEALED, T_CLOSED, T_SWEEP, T_DRAIN };

struct TileMeta {   // protected by the CTA's scheduler lock
  uint32_t state, granted, ready, drain_next, drained, lo;   // lo = lowest band-window bound of its rows
};
struct __align__(16) TileRow {   // who sits in a row slot: written by the depositor before `ready` counts it
  uint32_t sl, i, lo, pad;
  double xup;
  double pad2;
};

struct TileLayout {
  uint32_t tab, yB2, yB1a, yB1b, yC, yLev, yPerm, done, rows, row_bytes, tiles, tile_bytes, total;
};

__host__ __device__ inline TileLayout tile_layout(uint32_t nwarps, uint32_t nx_cap, uint32_t ny_cap, uint32_t e4_cap,
                                                  uint32_t lev_cap) {
  TileLayout L;
  uint32_t off = 0;
  auto take = [&](uint32_t bytes) { uint32_t at = off; off += (bytes + 15u) & ~15u; return at; };
  L.tab = take(8 * 256);
  L.yB2 = take(16 * ny_cap);   // {child list offset << 8 | bcode, deg4 | len << 16, coef = up*a*s2}   (sweep)
  L.yB1a = take(16 * ny_cap);  // {s2*bfreq*up, el*bfreq*up}                                              (MATCH)
  L.yB1b = take(16 * ny_cap);  // {paths*dn, the two integer words of the sweep record}                   (MATCH)
  L.yC = take(2 * e4_cap);
  L.yLev = take(8 * (lev_cap + 1));
  L.yPerm = take(4 * ny_cap);
  L.done = take(nx_cap * kGroup);
  L.row_bytes = max((8u * (ny_cap + 2u) + 15u) & ~15u, 16u * 128u);   // also parks a row's band values (128 x 16 bytes)
  L.rows = take(L.row_bytes * nwarps);
  off = (off + 127u) & ~127u;
  L.tile_bytes = 128u * (ny_cap + 1u);
  L.tiles = take(L.tile_bytes * kTiles);
  L.total = off;
  return L;
}

// byte offset of (column j, row slot r) inside a tile
__device__ __forceinline__ uint32_t tile_at(uint32_t j, uint32_t r) { return 128u * j + 8u * (r ^ (j & 15u)); }
// same, from a child entry of the 16-bit lists (o = 8 * column)
__device__ __forceinline__ uint32_t tile_at_off(uint32_t o, uint32_t r) { return 16u * o + 8u * (r ^ ((o >> 3) & 15u)); }

template <int NCH>
__global__ void __launch_bounds__(32 * kFastRegWarps, 1) stem_tile_kernel(const StemFastLaunch P) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ unsigned long long s_item;
  __shared__ uint32_t s_next_blk, s_g, s_maxblk;
  __shared__ PairSlot s_slot[kGroup];
  __shared__ uint32_t s_lock, s_cur_fill, s_rows_done, s_rows_total, s_hint;   // s_hint: tasks the scheduler holds (closed tiles + undrained rows)
  __shared__ TileMeta s_meta[kTiles];
  __shared__ TileRow s_trow[kTiles][kTileRows];
  const uint32_t nwarps = blockDim.x >> 5;
  const TileLayout L = tile_layout(nwarps, P.nx_cap, P.ny_cap, P.e4_cap, P.lev_cap);
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(sm);

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t band = P.len_band;
  for (uint32_t t = tid; t < 256; t += blockDim.x) sts_f64(sb + (L.tab + 8 * t), P.pair_tab[t]);
  double* __restrict__ slab = P.scratch + (size_t)blockIdx.x * P.scratch_stride;
  const unsigned long long slot_stride = P.scratch_stride / kGroup;
  const SetView& X = P.X;
  const SetView& Y = P.Y;
  const unsigned long long n_items = P.count[P.bucket];
  const uint32_t* __restrict__ order = P.order + P.start[P.bucket];
  const uint32_t rb = pin(sb + L.rows + L.row_bytes * warp);   // this warp's HQ row
  const uint32_t tiles = pin(sb + L.tiles);
  const uint32_t tile_bytes = pin(L.tile_bytes);
  const uint32_t yPerm = pin(sb + L.yPerm), yB2 = pin(sb + L.yB2), yC = pin(sb + L.yC), yLev = pin(sb + L.yLev);
  unsigned long long item = 0, item_end = 0;
#ifdef TILE_DEBUG
  __shared__ uint32_t s_abort;
  __shared__ uint32_t s_where[32];
  if (tid == 0) s_abort = 0;
  if (lane == 0) s_where[warp] = 0;
  const long long dbg_t0 = clock64();
  // a wait loop that has been spinning for more than ~1.5 s of SM clock records where and what it sees, once per CTA
  auto watchdog = [&](uint32_t where, uint32_t a, uint32_t b) -> bool {
    if (*(volatile uint32_t*)&s_abort) return true;
    if (clock64() - dbg_t0 > 3000000000ll) {
      if (lane == 0 && atomicCAS(&s_abort, 0u, 1u) == 0u && P.prof && atomicAdd(P.prof, 1ull) == 0ull) {
        P.prof[1] = where; P.prof[2] = warp; P.prof[3] = a; P.prof[4] = b;
        P.prof[5] = s_rows_done; P.prof[6] = s_rows_total; P.prof[7] = s_hint; P.prof[8] = s_cur_fill; P.prof[9] = s_lock;
        for (uint32_t t = 0; t < kTiles; ++t) { P.prof[10 + 3 * t] = s_meta[t].state; P.prof[11 + 3 * t] = s_meta[t].granted | (s_meta[t].ready << 8) | (s_meta[t].drain_next << 16) | (s_meta[t].drained << 24); P.prof[12 + 3 * t] = s_meta[t].lo; }
        P.prof[16] = s_next_blk; P.prof[17] = blockIdx.x;
        for (uint32_t q = 0; q < 12; ++q) P.prof[18 + q] = (unsigned long long)s_where[2 * q] | ((unsigned long long)s_where[2 * q + 1] << 32);
      }
      return true;
    }
    return false;
  };
#define WATCHDOG(where, a, b) if (watchdog(where, a, b)) break
#define ABORTED() (*(volatile uint32_t*)&s_abort != 0u)
#define WHERE(code) do { if (lane == 0) { *(volatile uint32_t*)&s_where[warp] = (code); if (blockIdx.x == 0 && P.prof) { \
    volatile unsigned long long* gp_ = P.prof; gp_[32 + warp] = (code); \
    gp_[64 + warp] = (unsigned long long)s_rows_done | ((unsigned long long)s_hint << 16) | ((unsigned long long)(s_cur_fill & 0xff) << 32) | ((unsigned long long)s_lock << 40) | ((unsigned long long)s_next_blk << 44); \
    gp_[96 + warp] = (unsigned long long)s_meta[0].state | ((unsigned long long)s_meta[0].granted << 4) | ((unsigned long long)s_meta[0].ready << 12) | ((unsigned long long)s_meta[0].drain_next << 20) | ((unsigned long long)s_meta[0].drained << 28) | \
                     ((unsigned long long)s_meta[1].state << 36) | ((unsigned long long)s_meta[1].granted << 40) | ((unsigned long long)s_meta[1].ready << 48) | ((unsigned long long)s_meta[1].drain_next << 56); } } } while (0)
#else
#define WATCHDOG(where, a, b) do {} while (0)
#define ABORTED() false
#define WHERE(code) do {} while (0)
#endif

  // ---- the scheduler (lane 0 of the calling warp, under the lock; results broadcast by the caller)
  auto lock = [&]() { while (atomicCAS(&s_lock, 0u, 1u) != 0u) { __nanosleep(20); WATCHDOG(1u, 0u, 0u); } __threadfence_block(); };
  auto unlock = [&]() { __threadfence_block(); atomicExch(&s_lock, 0u); };
  // a drain task (1 | tile << 4 | slot << 8 | rows << 16), a sweep (2 | tile << 4 | rows << 16 ; lo in *lo_out) or 0
  auto take_work = [&](uint32_t* lo_out) -> uint32_t {
    uint32_t w = 0u;
    if (lane == 0 && *(volatile uint32_t*)&s_hint != 0u) {
      lock();
      for (uint32_t t = 0; t < kTiles && !w; ++t)
        if (s_meta[t].state == T_DRAIN && s_meta[t].drain_next < s_meta[t].granted) {
          w = 1u | (t << 4) | (s_meta[t].drain_next++ << 8) | (s_meta[t].granted << 16);
          --s_hint;
        }
      for (uint32_t t = 0; t < kTiles && !w; ++t)
        if (s_meta[t].state == T_CLOSED) {
          s_meta[t].state = T_SWEEP;
          w = 2u | (t << 4) | (s_meta[t].granted << 16);
          *lo_out = s_meta[t].lo;
          --s_hint;
        }
      unlock();
    }
    w = __shfl_sync(0xffffffffu, w, 0);
    *lo_out = __shfl_sync(0xffffffffu, *lo_out, 0);
    return w;
  };
  // a row slot of the filling tile: 1 | tile << 4 | slot << 8, or 0 when every tile is busy
  auto grant = [&](uint32_t lo) -> uint32_t {
    uint32_t w = 0u;
    if (lane == 0) {
      lock();
      if (s_cur_fill == kNone)
        for (uint32_t t = 0; t < kTiles; ++t)
          if (s_meta[t].state == T_FREE) { s_meta[t].state = T_FILL; s_meta[t].granted = 0; s_meta[t].ready = 0; s_meta[t].lo = kNone; s_cur_fill = t; break; }
      if (s_cur_fill != kNone) {
        const uint32_t t = s_cur_fill;
        const uint32_t c = s_meta[t].granted++;
        s_meta[t].lo = min(s_meta[t].lo, lo);
        if (s_meta[t].granted == kTileRows) { s_meta[t].state = T_SEALED; s_cur_fill = kNone; }
        w = 1u | (t << 4) | (c << 8);
      }
      unlock();
    }
    return __shfl_sync(0xffffffffu, w, 0);
  };
  auto deposited = [&](uint32_t t) {
    if (lane == 0) {
      lock();
      if (++s_meta[t].ready == s_meta[t].granted && s_meta[t].state == T_SEALED) { s_meta[t].state = T_CLOSED; ++s_hint; }
      unlock();
    }
  };
  auto seal_current = [&]() {
    if (lane == 0 && *(volatile uint32_t*)&s_cur_fill != kNone) {
      lock();
      if (s_cur_fill != kNone && s_meta[s_cur_fill].granted > 0u) {
        const uint32_t t = s_cur_fill;
        if (s_meta[t].ready == s_meta[t].granted) { s_meta[t].state = T_CLOSED; ++s_hint; } else s_meta[t].state = T_SEALED;
        s_cur_fill = kNone;
      }
      unlock();
    }
  };
  auto sweep_done = [&](uint32_t t) {
    if (lane == 0) { lock(); s_meta[t].state = T_DRAIN; s_meta[t].drain_next = 0; s_meta[t].drained = 0; s_hint += s_meta[t].granted; unlock(); }
  };
  auto drain_done = [&](uint32_t t) {
    if (lane == 0) { lock(); if (++s_meta[t].drained == s_meta[t].granted) s_meta[t].state = T_FREE; unlock(); }
  };

  for (;;) {
    __syncthreads();  // previous group fully retired (also orders the tab fill on the first trip)
    if (tid == 0) {
      if (item >= item_end) { item = atomicAdd(P.counter, (unsigned long long)kGroup); item_end = item + kGroup; }
      if (item_end > n_items) item_end = n_items;
      uint32_t g = 0, rows = 0;
      if (item < item_end) {
        const uint32_t y0 = P.yi[order[item]];
        uint32_t maxblk = 0;
        while (item + g < item_end && g < kGroup && P.yi[order[item + g]] == y0) {
          const uint32_t k = order[item + g];
          const RecDev rx = X.rec[P.xi[k]];
          PairSlot ps;
          ps.k = k; ps.N = rx.N; ps.node0 = rx.node0; ps.blk0 = rx.blk0; ps.nblk = rx.nblk; ps.plr = rx.plr;
          s_slot[g] = ps;
          maxblk = max(maxblk, rx.nblk);
          rows += rx.N;
          ++g;
        }
        s_maxblk = maxblk;
        s_item = item;
        item += g;
      }
      s_g = g;
      s_next_blk = 0;
      s_lock = 0; s_cur_fill = kNone; s_rows_done = 0; s_rows_total = rows; s_hint = 0;
      for (uint32_t t = 0; t < kTiles; ++t) { s_meta[t].state = T_FREE; s_meta[t].granted = 0; s_meta[t].ready = 0; s_meta[t].drain_next = 0; s_meta[t].drained = 0; s_meta[t].lo = kNone; }
    }
    __syncthreads();
    const uint32_t g = s_g;
    if (g == 0) break;
    const RecDev ry = Y.rec[P.yi[order[s_item]]];
    const uint32_t Ny = ry.N;
    const uint32_t NYS = (Ny + 1u) & ~1u;
    const uint32_t n_tickets = g * s_maxblk;

    // ---- stage the y record, clear the row flags
    for (uint32_t j = tid; j < Ny; j += blockDim.x) {
      const uint32_t gy = ry.node0 + j;
      const double ys2 = Y.s2[gy], yup = Y.up[gy], ydn = Y.dn[gy];
      const NodeI ni = Y.nodei[gy];
      const double coef = yup * (Y.a[gy] * ys2);
      sts_v4u32(yB2 + 16 * j, make_uint4(ni.e4_bcode, (uint32_t)ni.deg4 | ((uint32_t)ni.len << 16),
                                         (uint32_t)__double2loint(coef), (uint32_t)__double2hiint(coef)));
      const double bfu = Y.bfreq[gy] * yup;
      sts_v2f64(sb + (L.yB1a + 16 * j), make_double2(ys2 * bfu, Y.el[gy] * bfu));
      sts_v2f64(sb + (L.yB1b + 16 * j), make_double2(Y.paths[gy] * ydn, __hiloint2double((int)((uint32_t)ni.deg4 | ((uint32_t)ni.len << 16)), (int)ni.e4_bcode)));
      sts_u32(yPerm + 4 * j, Y.lperm[gy]);
    }
    {
      const uint2* __restrict__ src = reinterpret_cast<const uint2*>(Y.c16 + ry.c16_0);
      for (uint32_t e = tid; e < ry.e4 / 4u; e += blockDim.x) { const uint2 v = src[e]; sts_u32(yC + 8 * e, v.x); sts_u32(yC + 8 * e + 4, v.y); }
    }
    for (uint32_t l = tid; l < ry.nlev; l += blockDim.x) { sts_u32(yLev + 8 * l, Y.lev_off[ry.lev0 + l]); sts_u32(yLev + 8 * l + 4, Y.lev_off[ry.lev0 + l + 1]); }
    for (uint32_t i = tid; i < (g * P.nx_cap + 3u) / 4u; i += blockDim.x) sts_u32(sb + (L.done + 4 * i), 0u);
    if (lane < 2u) sts_f64(rb + 8u * (Ny + lane), 0.0);   // the dummy column (and the pad after it) of this warp's HQ row
    __syncthreads();

    // ---- phase C of one swept row: slab(i,:) += up_x * H(i,:), publish
    auto drain_row = [&](uint32_t t, uint32_t c) {
      WHERE(0x100u | (t << 4) | c);
      const TileRow tr = s_trow[t][c];
      const uint32_t tb = tiles + tile_bytes * t;
      double* __restrict__ g0 = slab + tr.sl * slot_stride + (size_t)tr.i * NYS;
      const double xup = tr.xup;
      for (uint32_t jb = 0; jb < Ny; jb += 128u) {
        double h[4], o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t j = jb + lane + 32u * k;
          h[k] = 0.0;
          if (j < Ny) h[k] = lds_f64(tb + tile_at(j, c));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { o[k] = 0.0; if (h[k] != 0.0) o[k] = __ldcg(g0 + jb + lane + 32u * k); }
#pragma unroll
        for (int k = 0; k < 4; ++k) if (h[k] != 0.0) g0[jb + lane + 32u * k] = fma(xup, h[k], o[k]);
      }
      __threadfence_block();
      __syncwarp();
      if (lane == 0) {
        asm volatile("st.volatile.shared.u8 [%0], %1;" ::"r"(sb + L.done + tr.sl * P.nx_cap + tr.i), "r"(1u) : "memory");
        atomicAdd(&s_rows_done, 1u);
      }
      drain_done(t);
    };

    // ---- phase B2 of a closed tile: half-warp <-> node, lane <-> row slot
    auto sweep_tile = [&](uint32_t t, uint32_t rows, uint32_t lo_tile) {
      WHERE(0x200u | (t << 8) | rows);
      const uint32_t tb = tiles + tile_bytes * t;
      const uint32_t h = lane >> 4, r = lane & 15u;
      const bool mine = r < rows;
      const uint32_t nlev = ry.nlev;
      for (uint32_t base = ry.sub1 & ~31u; base < nlev; base += 32u) {
        const uint32_t sl0 = base + lane;
        bool act = sl0 >= ry.sub1 && sl0 < nlev;
        if (act) act = (lds_u32(yB2 + 16u * lds_u32(yLev + 8u * sl0) + 4u) >> 16) >= lo_tile;   // nodes of a sub-level are sorted by length
        uint32_t mask = __ballot_sync(0xffffffffu, act);
        while (mask != 0u) {
          const uint32_t sl = base + (uint32_t)__ffs((int)mask) - 1u;
          mask &= mask - 1u;
          const uint2 jj = lds_v2u32(yLev + 8u * sl);
          // two nodes per half-warp and trip (independent inside a sub-level): their loads overlap
          for (uint32_t j0 = jj.x; j0 < jj.y; j0 += 4u) {
            const uint32_t ja = j0 + h, jb2 = j0 + 2u + h;
            uint4 ra = make_uint4(0u, 0u, 0u, 0u), rbq = ra;
            if (ja < jj.y) ra = lds_v4u32(yB2 + 16u * ja);
            if (jb2 < jj.y) rbq = lds_v4u32(yB2 + 16u * jb2);
            const bool oka = mine && ja < jj.y && (ra.y >> 16) >= lo_tile;
            const bool okb = mine && jb2 < jj.y && (rbq.y >> 16) >= lo_tile;
            uint32_t ea = yC + 2u * (ra.x >> 8), eb = yC + 2u * (rbq.x >> 8);
            const uint32_t ea_end = ea + 8u * (ra.y & 0xffu), eb_end = eb + 8u * (rbq.y & 0xffu);
            uint2 ca = make_uint2(0u, 0u), cb = ca;
            if (oka && ea < ea_end) ca = lds_v2u32(ea);
            if (okb && eb < eb_end) cb = lds_v2u32(eb);
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0, hva = 0.0, hvb = 0.0;
            if (oka) {
              hva = lds_f64(tb + tile_at(ja, r));
              if (ea < ea_end) {
                a0 = lds_f64(tb + tile_at_off(ca.x & 0xffffu, r)); a1 = lds_f64(tb + tile_at_off(ca.x >> 16, r));
                a2 = lds_f64(tb + tile_at_off(ca.y & 0xffffu, r)); a3 = lds_f64(tb + tile_at_off(ca.y >> 16, r));
              }
            }
            if (okb) {
              hvb = lds_f64(tb + tile_at(jb2, r));
              if (eb < eb_end) {
                b0 = lds_f64(tb + tile_at_off(cb.x & 0xffffu, r)); b1 = lds_f64(tb + tile_at_off(cb.x >> 16, r));
                b2 = lds_f64(tb + tile_at_off(cb.y & 0xffffu, r)); b3 = lds_f64(tb + tile_at_off(cb.y >> 16, r));
              }
            }
            if (oka) {
              double S0 = a0 + a1, S1 = a2 + a3;
#pragma unroll 1
              for (ea += 8u; ea < ea_end; ea += 8u) {
                const uint2 c4 = lds_v2u32(ea);
                S0 += lds_f64(tb + tile_at_off(c4.x & 0xffffu, r)) + lds_f64(tb + tile_at_off(c4.x >> 16, r));
                S1 += lds_f64(tb + tile_at_off(c4.y & 0xffffu, r)) + lds_f64(tb + tile_at_off(c4.y >> 16, r));
              }
              sts_f64(tb + tile_at(ja, r), fma(__hiloint2double((int)ra.w, (int)ra.z), S0 + S1, hva));
            }
            if (okb) {
              double S0 = b0 + b1, S1 = b2 + b3;
#pragma unroll 1
              for (eb += 8u; eb < eb_end; eb += 8u) {
                const uint2 c4 = lds_v2u32(eb);
                S0 += lds_f64(tb + tile_at_off(c4.x & 0xffffu, r)) + lds_f64(tb + tile_at_off(c4.x >> 16, r));
                S1 += lds_f64(tb + tile_at_off(c4.y & 0xffffu, r)) + lds_f64(tb + tile_at_off(c4.y >> 16, r));
              }
              sts_f64(tb + tile_at(jb2, r), fma(__hiloint2double((int)rbq.w, (int)rbq.z), S0 + S1, hvb));
            }
          }
          __syncwarp();
        }
      }
      __threadfence_block();
      sweep_done(t);
    };

    // one scheduler task, if there is one
    auto help = [&]() -> bool {
      uint32_t lo_t = 0u;
      const uint32_t w = take_work(&lo_t);
      if ((w & 15u) == 1u) { drain_row((w >> 4) & 15u, (w >> 8) & 255u); return true; }
      if ((w & 15u) == 2u) { sweep_tile((w >> 4) & 15u, w >> 16, lo_t); return true; }
      return false;
    };

    bool more_tickets = true;
    for (;;) {
      WHERE(0x10u);
      if (help()) continue;
      WHERE(0x11u);
      uint32_t t = n_tickets;
      if (more_tickets) {
        if (lane == 0) t = atomicAdd(&s_next_blk, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
      }
      if (t >= n_tickets) {
        more_tickets = false;
        WHERE(0x20u);
        seal_current();
        // one lane decides for the warp: lanes 1..31 get here before lane 0 leaves seal_current, and a count that
        // changes in between must not split the warp between the barrier below and the next help()
        uint32_t fin = 0u;
        if (lane == 0) fin = *(volatile uint32_t*)&s_rows_done == s_rows_total;
        if (__shfl_sync(0xffffffffu, fin, 0)) break;
        if (!help()) __nanosleep(200);
        WATCHDOG(2u, t, 0u);
        continue;
      }
      const uint32_t sl = t % g, b = t / g;
      const PairSlot ps = s_slot[sl];
      if (b >= ps.nblk) continue;
      const uint32_t blk = X.blk[ps.blk0 + b];
      const uint32_t i0 = blk & 0xffffu, cnt = blk >> 16;
      double* __restrict__ G0 = slab + sl * slot_stride;
      const uint32_t done = sb + L.done + sl * P.nx_cap;
      double* __restrict__ rowacc = P.rowacc + ((size_t)blockIdx.x * kGroup + sl) * P.nx_cap;

      for (uint32_t r = 0; r < cnt; ++r) {
        const uint32_t i = i0 + r;
        const XNode* __restrict__ xn = X.xnode + ps.node0 + i;
        const double2 x01 = __ldg(reinterpret_cast<const double2*>(xn));       // {s2, a}
        const double2 x23 = __ldg(reinterpret_cast<const double2*>(xn) + 1);   // {up, ql}
        const double2 x45 = __ldg(reinterpret_cast<const double2*>(xn) + 2);   // {bfreq, paths}
        const uint4 xi4 = __ldg(reinterpret_cast<const uint4*>(xn) + 3);       // {e0, e1, len, bcode}
        const uint32_t e0 = xi4.x, e1 = xi4.y, xl = xi4.z, xbc = xi4.w;
        const double xs2 = x01.x, pc = x23.x * x01.y;   // pc = up_x * a_x
        double* __restrict__ g0row = G0 + (size_t)i * NYS;

        // ---- wait until the rows of ALL inner pairs are published, serving the scheduler meanwhile
        WHERE(0x30000000u | (sl << 24) | i);
        for (uint32_t eb = e0; eb < e1; eb += 32u) {
          const uint32_t ne = min(32u, e1 - eb);
          uint32_t c = 0u;
          if (lane < ne) c = X.cidx[eb + lane];
          while (!__all_sync(0xffffffffu, lane >= ne || ld_flag_f(done + c) != 0u)) {
            seal_current();
            if (!help()) __nanosleep(128);
            WATCHDOG(3u, sl, i);
          }
        }
        __threadfence_block();  // acquire: the G0 rows behind the flags just seen
        WHERE(0x40000000u | (sl << 24) | i);

        // ---- phase A: HQ(i,:) = s2_x * sum over inner pairs of G0ss(c,:), the row in registers
        {
          constexpr int CPR = NCH <= 4 ? 3 : 2;
          const uint32_t j = 2u * lane;
          double2 q[NCH];
#pragma unroll
          for (int c = 0; c < NCH; ++c) q[c] = make_double2(0.0, 0.0);
          for (uint32_t eb = e0; eb < e1; eb += 32u) {
            const uint32_t ne = min(32u, e1 - eb);
            uint32_t c = 0u;
            if (lane < ne) c = X.cidx[eb + lane];
            const uint32_t off_l = c * NYS;
            prefetch_rows(G0, off_l, lane < ne, NYS, Ny, lane, ne);
#pragma unroll 1
            for (uint32_t tt = 0; tt < ne; tt += CPR) {
              double2 tv[CPR][NCH];
#pragma unroll
              for (int k = 0; k < CPR; ++k) {
                const bool ck = tt + k < ne;
                const double2* __restrict__ sk = reinterpret_cast<const double2*>(G0 + __shfl_sync(0xffffffffu, off_l, (tt + k) & 31u) + j);
#pragma unroll
                for (int cc = 0; cc < NCH; ++cc) {
                  tv[k][cc] = make_double2(0.0, 0.0);
                  if (ck && j + 64u * cc < Ny) tv[k][cc] = __ldcg(sk + 32 * cc);
                }
              }
#pragma unroll
              for (int cc = 0; cc < NCH; ++cc) {
                double sx = tv[0][cc].x, sy = tv[0][cc].y;
#pragma unroll
                for (int k = 1; k < CPR; ++k) { sx += tv[k][cc].x; sy += tv[k][cc].y; }
                q[cc].x += sx; q[cc].y += sy;
              }
            }
          }
#pragma unroll
          for (int cc = 0; cc < NCH; ++cc) {
            const uint32_t jc = j + 64u * cc;
            if (jc < Ny) {
              const double sx = xs2 * q[cc].x, sy = xs2 * q[cc].y;
              *reinterpret_cast<double2*>(g0row + jc) = make_double2(pc * sx, pc * sy);
              sts_v2f64(rb + 8u * jc, make_double2(sx, jc + 1u < Ny ? sy : 0.0));
            }
          }
        }
        __syncwarp();

        // ---- phase B1: MATCH on the y nodes inside the band of row i (a range of the length-sorted node list)
        uint32_t lo = 0u, hi = Ny;
        if (band != 0u) {
          lo = xl > band ? count_len_below(yPerm, Ny, xl - band, lane) : 0u;
          hi = count_len_below(yPerm, Ny, xl + band + 1u, lane);
        }
        uint32_t jr[4];
        double mr[4];
        {
          const double xql = x23.y, xbf = x45.x;
          const uint32_t tabx = sb + L.tab + 128u * xbc;
          const uint32_t yB1a = sb + L.yB1a, yB1b = sb + L.yB1b;
          double racc = 0.0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t tq = lo + lane + 32u * k;
            jr[k] = kNone; mr[k] = 0.0;
            if (tq < hi) {
              const uint32_t jn = lds_u32(yPerm + 4u * tq) & 0xffffu;
              const double2 ya = lds_v2f64(yB1a + 16u * jn);   // {s2_y*bfreq_y*up_y, el_y*bfreq_y*up_y}
              const double2 yb = lds_v2f64(yB1b + 16u * jn);   // {paths_y*dn_y, record words}
              const uint32_t rx = (uint32_t)__double2loint(yb.y), ry2 = (uint32_t)__double2hiint(yb.y);
              const double tvv = lds_f64(tabx + 8u * (rx & 0xffu));
              uint32_t e = yC + 2u * (rx >> 8);
              const uint32_t eend = e + 8u * (ry2 & 0xffffu);
              double R0 = 0.0, R1 = 0.0;
#pragma unroll 1
              for (; e < eend; e += 8u) {
                const uint2 c4 = lds_v2u32(e);
                R0 += lds_f64(rb + (c4.x & 0xffffu)) + lds_f64(rb + (c4.x >> 16));
                R1 += lds_f64(rb + (c4.y & 0xffffu)) + lds_f64(rb + (c4.y >> 16));
              }
              const double mp = (tvv * xbf) * fma(ya.y, xql, ya.x * (R0 + R1));   // up_y * M(i,j)
              racc = fma(yb.x, mp, racc);
              jr[k] = jn; mr[k] = mp;
            }
          }
          racc = warp_sum_all(racc);
          if (lane == 0) rowacc[i] = x45.y * racc;   // per-row slot in global scratch (L2)
        }
        __syncwarp();

        // ---- a row slot of the filling tile (every tile busy: park the band values in the HQ row and serve)
        const uint32_t len_lo = (band != 0u && xl > band) ? xl - band : 0u;   // below the window G1 is identically 0
        WHERE(0x50000000u | (sl << 24) | i);
        uint32_t gw = grant(len_lo);
        if (gw == 0u) {
#pragma unroll
          for (int k = 0; k < 4; ++k) { sts_f64(rb + 16u * (lane + 32u * k), mr[k]); sts_u32(rb + 16u * (lane + 32u * k) + 8u, jr[k]); }
          do { WHERE(0x60000000u | (sl << 24) | i); if (!help()) __nanosleep(128); gw = grant(len_lo); if (ABORTED()) break; WATCHDOG(4u, sl, i); } while (gw == 0u);
          if (gw == 0u) break;
#pragma unroll
          for (int k = 0; k < 4; ++k) { mr[k] = lds_f64(rb + 16u * (lane + 32u * k)); jr[k] = lds_u32(rb + 16u * (lane + 32u * k) + 8u); }
          if (lane < 2u) sts_f64(rb + 8u * (Ny + lane), 0.0);   // the HQ row's dummy column, for the next row
        }
        const uint32_t tt_ = (gw >> 4) & 15u, cc_ = (gw >> 8) & 255u;
        const uint32_t tb = tiles + tile_bytes * tt_;
        // ---- phase Z: the slot's entries of every column (the dummy column too) start at zero, then the band's up_y*M
        for (uint32_t j = lane; j <= Ny; j += 32u) sts_f64(tb + tile_at(j, cc_), 0.0);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k) if (jr[k] != kNone) sts_f64(tb + tile_at(jr[k], cc_), mr[k]);
        if (lane == 0) {
          TileRow tr;
          tr.sl = sl; tr.i = i; tr.lo = len_lo; tr.pad = 0; tr.xup = x23.x; tr.pad2 = 0.0;
          s_trow[tt_][cc_] = tr;
        }
        __threadfence_block();
        __syncwarp();
        deposited(tt_);
      }
      if (ABORTED()) break;
    }
    if (ABORTED()) break;

    // ---- fixed-order sum of the per-row slots, one warp per pair of the group
    WHERE(0x70u);
    __syncthreads();
    if (warp < g) {
      const PairSlot ps = s_slot[warp];
      double t = 0.0;
      const double* __restrict__ ra = P.rowacc + ((size_t)blockIdx.x * kGroup + warp) * P.nx_cap;
      for (uint32_t i = lane; i < ps.N; i += 32u) t += __ldcg(ra + i);
      t = warp_sum_all(t);
      if (lane == 0) P.out[ps.k] = t + ps.plr * (double)ry.lr;
    }
  }
}

using TileKernelFn = void (*)(StemFastLaunch);
TileKernelFn tile_kernel_for(uint32_t ny_cap) {
  const uint32_t nch = (ny_cap + 63u) / 64u;
  if (nch <= 4) return stem_tile_kernel<4>;
  if (nch == 5) return stem_tile_kernel<5>;
  return stem_tile_kernel<6>;
}

}  // namespace

// the tile kernel serves staged records of up to 384 nodes whose bands hold at most 128 nodes
bool stem_tile_serves(uint32_t ny_cap, uint32_t band_cap) { return ny_cap <= 384u && band_cap <= 128u; }

size_t stem_tile_smem_bytes(uint32_t nwarps, uint32_t nx_cap, uint32_t ny_cap, uint32_t e4_cap, uint32_t lev_cap) {
  return tile_layout(nwarps, nx_cap, ny_cap, e4_cap, lev_cap).total + 128u;   // + alignment slack of the dynamic window
}

cudaError_t launch_stem_tile(const StemFastLaunch& p, int grid, int nwarps, size_t smem, cudaStream_t stream) {
  const TileKernelFn fn = tile_kernel_for(p.ny_cap);
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  fn<<<grid, nwarps * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace stemk

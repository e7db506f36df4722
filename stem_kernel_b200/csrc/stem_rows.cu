// stem_kernel_b200/csrc/stem_rows.cu -- the stem (base-pair DAG) kernel, row-block path ("lanes are rows").
//
// Same restated recurrence as stem_fast.cu (StemKernel<ST,D>::operator(), stem_kernel_lite/stem_kernel.cpp:14-95;
// node / edge scores score_table.cpp:56-101,162-201) on pre-scaled rows,
//     H (i,j) = up_y(j) * G1(i,j)          HQ(i,j) = up_y(j) * Q(i,j)          G0s(i,j) = up_x(i) * G0(i,j)
//     Q (i,j) = s2_x(i) * sum_cx G0s(cx,j)
//     R (i,j) = s2_y(j) * sum_cy HQ(i,cy) + el_y(j)*ql_x(i)         S(i,j) = s2_y(j) * sum_cy H(i,cy)
//     M = in_band ? v_s*R : 0      G1 = M + a_y(j)*S      G0 = G1 + a_x(i)*Q = dn_y(j) * (H + a_x(i)*HQ)
//     k(x,y) = sum_i paths_x(i) * sum_j paths_y(j) * M(i,j)  (+ plr_x * lr_y)
// but with the opposite thread mapping.  stem_fast.cu gives a warp ONE row and spreads the ~14 nodes of a y level
// over its lanes (44 % of the lanes busy, every instruction of the level sweep paid per row).  Here a BLOCK of
// R = 8 (or 4) rows of one DAG level of one x record is swept in lockstep by a TEAM of warps: a lane owns
// (row of the block, node slot), a warp takes 32/R y nodes at a time for all rows of the block, so
//   * the y-side node record and child list are shared by the R rows of a slot (broadcast loads),
//   * the gathers H[cy][row] of a slot are one contiguous 8R-byte piece of shared memory,
//   * the instructions of a node are shared by R rows, and
//   * nodes below every row's length window cost one test and one store.
// A CTA (one per SM) stages the y record once for a GROUP of pairs sharing it and runs as many teams as fit next
// to it in shared memory (each team owns an H and an HQ tile of R rows x Ny columns).  Teams pull blocks from a
// ticket queue; tickets interleave the pairs of the group level by level (levels aligned at the top), a pair's
// own blocks keep their level order, and a block waits on per-row flags for the rows of its inner pairs, so waiting
// only ever points to earlier tickets: no deadlock.  Per block:
//   A  warp <-> row, lanes <-> columns: HQ row = up_y * s2_x * sum of the finished pre-scaled G0 rows of the row's
//      inner pairs (coalesced L2 reads of the per-pair slab), stored transposed into the tile
//   B  y level by y level, a named barrier of the team between levels; the node record and first children of the
//      next level are fetched BEFORE the barrier so that only gather -> add -> store sits between two barriers
//   C  warp <-> row, lanes <-> columns: G0s row = up_x * dn_y * (H + a_x*HQ) to the slab (coalesced), row flag
// Row sums of the path-weighted MATCH terms are reduced in a fixed order: results are bit-reproducible and do not
// depend on which team ran which block.
#include "kernels.cuh"

namespace stemk {

namespace {

constexpr uint32_t kG = kFastGroup;

struct RowsLayout {
  uint32_t tab, yI, yD1, yB, yDn, yC, yLev, cnt, start, pref, blk, done, teams, team_bytes, part, H, HQ, total;
};

// Tile pitch R+1 doubles: lanes <-> rows at a fixed column is contiguous, lanes <-> columns at a fixed row strides
// by (R+1)*8 bytes, which walks all 32 banks for R = 8 and R = 4 (72 = 18 words, 40 = 10 words: both coprime
// enough with 32 that 16 consecutive columns hit 16 distinct bank pairs).
__host__ __device__ inline RowsLayout rows_layout(uint32_t R, uint32_t nteams, uint32_t team_warps, uint32_t nx_cap,
                                                  uint32_t ny_cap, uint32_t e4_cap, uint32_t ylev_cap, uint32_t xlev_cap) {
  RowsLayout L;
  uint32_t off = 0;
  auto take = [&](uint32_t bytes) { uint32_t at = off; off += (bytes + 15u) & ~15u; return at; };
  L.tab = take(8 * 256);
  L.yI = take(8 * ny_cap);      // {child list offset << 8 | bcode, deg | len << 16}
  L.yD1 = take(16 * ny_cap);    // {up*a*s2, up}
  L.yB = take(32 * ny_cap);     // {s2, el, paths, bfreq}   (MATCH cells only)
  L.yDn = take(8 * ny_cap);     // dn
  L.yC = take(2 * e4_cap);      // child lists: 8 * child number (16 bit), padded with 8*N (the zero row)
  L.yLev = take(4 * (ylev_cap + 1));
  L.cnt = take(2 * kG * xlev_cap);     // per (round, pair): rows of the pair's level
  L.start = take(2 * kG * xlev_cap);   // ... first row
  L.pref = take(2 * kG * xlev_cap);    // ... first block
  L.blk = take(4 * kG * (nx_cap / R + xlev_cap + 1));   // block list: pair | first row << 4 | rows << 20
  L.done = take(kG * nx_cap);          // one byte per row, per pair of the group
  L.part = 0;                          // inside a team's area
  const uint32_t tile = 8u * (R + 1u) * (ny_cap + 1u);   // + the zero row
  L.H = (8u * 32u * team_warps + 15u) & ~15u;
  L.HQ = L.H + ((tile + 15u) & ~15u);
  L.team_bytes = L.HQ + ((tile + 15u) & ~15u);
  L.teams = take(L.team_bytes * nteams);
  L.total = off;
  return L;
}

__device__ __forceinline__ uint32_t ld_flag(uint32_t addr) {
  uint32_t v; asm volatile("ld.volatile.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v;
}
// raw 32-bit shared addresses (see stem_fast.cu)
__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ double2 lds_v2f64(uint32_t a) { double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds_v2u32(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_v2f64(uint32_t a, double2 v) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_v2u32(uint32_t a, uint2 v) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory"); }
__device__ __forceinline__ void team_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct RowSlot {         // one pair of the group in flight
  uint32_t k;            // pair number (index into xi / yi / out)
  uint32_t N, node0, lev0, nlev, shift;   // shift = maxlev - nlev: the pair joins at round `shift`
  double plr;
};

struct NodePre {         // what a lane needs about its next y node, fetched ahead of the level barrier
  uint2 ni;              // {child list offset << 8 | bcode, deg | len << 16}
  double2 d1;            // {up*a*s2, up}
  uint2 c4;              // its first four children
};

template <int RSH>
__global__ void __launch_bounds__(kRowsMaxThreads, 1) stem_rows_kernel(const StemRowsLaunch P) {
  constexpr uint32_t R = 1u << RSH;         // rows of a block
  constexpr uint32_t NS = 32u >> RSH;       // y nodes a warp sweeps side by side
  constexpr uint32_t RP = R + 1u;           // tile pitch, doubles
  constexpr uint32_t ROWB = 8u * RP;
  extern __shared__ __align__(16) unsigned char sm[];
  __shared__ unsigned long long s_item;
  __shared__ uint32_t s_g, s_maxlev, s_nblocks, s_ticket;
  __shared__ uint32_t s_tk[16];
  __shared__ RowSlot s_slot[kG];
  const uint32_t TW = P.team_warps, TT = 32u * TW;
  const uint32_t nteams = blockDim.x / TT;
  const RowsLayout L = rows_layout(R, nteams, TW, P.nx_cap, P.ny_cap, P.e4_cap, P.ylev_cap, P.xlev_cap);
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(sm);

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t team = warp / TW, tw = warp % TW, bar_id = 1u + team;
  const uint32_t band = P.len_band;
  for (uint32_t t = tid; t < 256; t += blockDim.x) sts_f64(sb + (L.tab + 8 * t), P.pair_tab[t]);
  double* __restrict__ slab = P.scratch + (size_t)blockIdx.x * P.scratch_stride;
  const unsigned long long slot_stride = P.scratch_stride / kG;
  double* __restrict__ rowacc0 = P.rowacc + (size_t)blockIdx.x * P.rowacc_stride;   // kG x nx_cap row slots
  const SetView& X = P.X;
  const SetView& Y = P.Y;
  const unsigned long long n_items = P.count[P.bucket];
  const uint32_t* __restrict__ order = P.order + P.start[P.bucket];
  unsigned long long item = 0, item_end = 0;
  const uint32_t rr = lane & (R - 1u), q = lane >> RSH;   // phase B: this lane's row of the block, node slot
  const uint32_t tbase = sb + L.teams + L.team_bytes * team;   // this team's area: partial sums | H tile | HQ tile
  const uint32_t tH = tbase + L.H, tHQ = tbase + L.HQ;

  for (;;) {
    __syncthreads();
    if (tid == 0) {
      if (item >= item_end) { item = atomicAdd(P.counter, (unsigned long long)kG); item_end = item + kG; }
      if (item_end > n_items) item_end = n_items;
      uint32_t g = 0, maxlev = 0;
      if (item < item_end) {
        const uint32_t y0 = P.yi[order[item]];
        while (item + g < item_end && g < kG && P.yi[order[item + g]] == y0) {
          const uint32_t k = order[item + g];
          const RecDev rx = X.rec[P.xi[k]];
          RowSlot ps;
          ps.k = k; ps.N = rx.N; ps.node0 = rx.node0; ps.lev0 = rx.lev0; ps.nlev = rx.nlev; ps.shift = 0; ps.plr = rx.plr;
          s_slot[g] = ps;
          maxlev = max(maxlev, rx.nlev);
          ++g;
        }
        for (uint32_t s = 0; s < g; ++s) s_slot[s].shift = maxlev - s_slot[s].nlev;
        s_item = item;
        item += g;
      }
      s_g = g;
      s_maxlev = maxlev;
      s_ticket = 0;
    }
    __syncthreads();
    const uint32_t g = s_g;
    if (g == 0) break;
    const uint32_t maxlev = s_maxlev;
    const RecDev ry = Y.rec[P.yi[order[s_item]]];
    const uint32_t Ny = ry.N;
    const uint32_t NYS = (Ny + 3u) & ~3u;  // row stride of the G0 slabs (32-byte sectors)

    // ---- stage the y record, clear the row flags, count the rows of every (round, pair)
    for (uint32_t j = tid; j < Ny; j += blockDim.x) {
      const uint32_t gy = ry.node0 + j;
      const NodeB nb = Y.yband[gy];
      const double yup = Y.up[gy];
      sts_v2f64(sb + (L.yD1 + 16 * j), make_double2(yup * (Y.a[gy] * nb.s2), yup));
      sts_v2f64(sb + (L.yB + 32 * j), make_double2(nb.s2, nb.el));
      sts_v2f64(sb + (L.yB + 32 * j + 16), make_double2(nb.paths, nb.bfreq));
      sts_f64(sb + (L.yDn + 8 * j), Y.dn[gy]);
      const NodeI ni = Y.nodei[gy];
      const uint32_t deg = Y.coff[ry.coff0 + j + 1] - Y.coff[ry.coff0 + j];
      sts_v2u32(sb + (L.yI + 8 * j), make_uint2(ni.e4_bcode, deg | ((uint32_t)ni.len << 16)));
    }
    {
      const uint2* __restrict__ src = reinterpret_cast<const uint2*>(Y.c16 + ry.c16_0);
      for (uint32_t e = tid; e < ry.e4 / 4u; e += blockDim.x) sts_v2u32(sb + (L.yC + 8 * e), src[e]);
    }
    for (uint32_t l = tid; l <= ry.nlev; l += blockDim.x) sts_u32(sb + (L.yLev + 4 * l), Y.lev_off[ry.lev0 + l]);
    for (uint32_t i = tid; i < (g * P.nx_cap + 3u) / 4u; i += blockDim.x) sts_u32(sb + (L.done + 4 * i), 0u);
    for (uint32_t t = tid; t < maxlev * g; t += blockDim.x) {
      const uint32_t k = t / g, s = t - k * g;
      const RowSlot ps = s_slot[s];
      uint32_t r0 = 0, n = 0;
      if (k >= ps.shift) {
        const uint32_t l = k - ps.shift;
        r0 = X.lev_off[ps.lev0 + l];
        n = X.lev_off[ps.lev0 + l + 1] - r0;
      }
      sts_u16(sb + (L.cnt + 2u * t), n);
      sts_u16(sb + (L.start + 2u * t), r0);
    }
    // the zero row (index Ny) of this team's tiles: what the padding entries of the child lists point to
    if (tw == 0 && lane < 2u * R) sts_f64((lane < R ? tH : tHQ) + ROWB * Ny + 8u * (lane & (R - 1u)), 0.0);
    __syncthreads();
    if (tid == 0) {
      uint32_t acc = 0;
      for (uint32_t t = 0; t < maxlev * g; ++t) {
        sts_u16(sb + (L.pref + 2u * t), acc);
        acc += (lds_u16(sb + (L.cnt + 2u * t)) + R - 1u) >> RSH;
      }
      s_nblocks = acc;
    }
    __syncthreads();
    for (uint32_t t = tid; t < maxlev * g; t += blockDim.x) {
      const uint32_t k = t / g, s = t - k * g;
      const uint32_t n = lds_u16(sb + (L.cnt + 2u * t)), r0 = lds_u16(sb + (L.start + 2u * t));
      uint32_t b = lds_u16(sb + (L.pref + 2u * t));
      for (uint32_t r = 0; r < n; r += R, ++b) sts_u32(sb + (L.blk + 4u * b), s | ((r0 + r) << 4) | (min(R, n - r) << 20));
    }
    __syncthreads();
    const uint32_t nblocks = s_nblocks;

    for (;;) {
      if (tw == 0 && lane == 0) s_tk[team] = atomicAdd(&s_ticket, 1u);
      team_sync(bar_id, TT);
      const uint32_t t = *reinterpret_cast<volatile uint32_t*>(&s_tk[team]);
      if (t >= nblocks) break;
      const uint32_t bw = lds_u32(sb + (L.blk + 4u * t));
      const uint32_t sl = bw & 15u, row0 = (bw >> 4) & 0xffffu, cnt = bw >> 20;
      const RowSlot& ps = s_slot[sl];
      double* __restrict__ G0 = slab + sl * slot_stride;
      const uint32_t done = sb + L.done + sl * P.nx_cap;

      // ---- phase A: HQ(:, r) = up_y * s2_x(i) * sum over inner pairs c of G0s(c, :)
      for (uint32_t r = tw; r < cnt; r += TW) {
        const XNode* __restrict__ xn = X.xnode + ps.node0 + row0 + r;
        const uint4 xi4 = __ldg(reinterpret_cast<const uint4*>(xn) + 3);
        const uint32_t e0 = xi4.x, e1 = xi4.y;
        const double xs2 = __ldg(&xn->s2);
        const uint32_t hq = tHQ + 8u * r;
        for (uint32_t eb = e0; eb < e1 || eb == e0; eb += 32u) {
          const uint32_t ne = min(32u, e1 - eb);
          const bool first = eb == e0, last = eb + 32u >= e1;
          uint32_t off_l = 0u;
          if (lane < ne) {
            const uint32_t c = __ldg(X.cidx + eb + lane);
            off_l = c * NYS;
            while (ld_flag(done + c) == 0u) __nanosleep(32);   // wait until that row is published
          }
          __syncwarp();
          __threadfence_block();   // acquire: the G0 rows behind the flags just seen
          for (uint32_t jb = 0; jb < Ny; jb += 256u) {
            const uint32_t j = jb + lane;
            double qv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) qv[u] = (!first && j + 32u * u < Ny) ? lds_f64(hq + ROWB * (j + 32u * u)) : 0.0;
#pragma unroll 2
            for (uint32_t tt = 0; tt < ne; ++tt) {
              const double* __restrict__ src = G0 + __shfl_sync(0xffffffffu, off_l, tt) + j;
#pragma unroll
              for (int u = 0; u < 8; ++u) if (j + 32u * u < Ny) {
#ifndef ROWS_ABL_NO_A
                qv[u] += __ldcg(src + 32 * u);
#else
                qv[u] += (double)(size_t)src * 1e-300;
#endif
              }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const uint32_t jj = j + 32u * u;
              if (jj < Ny) sts_f64(hq + ROWB * jj, last ? lds_v2f64(sb + (L.yD1 + 16u * jj)).y * (xs2 * qv[u]) : qv[u]);
            }
          }
          if (e1 == e0) break;
        }
      }
      team_sync(bar_id, TT);

      // ---- phase B: sweep the y DAG level by level; lanes <-> (node slot, row of the block), warps <-> more slots
      const bool valid = rr < cnt;
      uint32_t xl = 0, xbc = 0;
      double xql = 0.0, xbf = 0.0;
      if (valid) {
        const XNode* __restrict__ xr = X.xnode + ps.node0 + row0 + rr;
        const double2 x12 = __ldg(reinterpret_cast<const double2*>(xr) + 1), x22 = __ldg(reinterpret_cast<const double2*>(xr) + 2);
        const uint4 xr4 = __ldg(reinterpret_cast<const uint4*>(xr) + 3);
        xql = x12.y; xbf = x22.x; xl = xr4.z; xbc = xr4.w;
      }
      // MATCH needs |len_x - len_y| <= band; below the window G1 is identically 0 (length-monotone DAG)
      const uint32_t len_lo = !valid ? 0xffffu : ((band != 0u && xl > band) ? xl - band : 0u);
      const uint32_t len_hi = band != 0u ? xl + band : 0xffffffffu;
      const uint32_t hbase = tH + 8u * rr, hqbase = tHQ + 8u * rr;
      const uint32_t tabrow = sb + L.tab + 8u * 16u * xbc;
      const uint32_t slot_id = tw * NS + q, step = TW * NS;
      double racc = 0.0;
      auto fetch = [&](uint32_t j, uint32_t jend) {
        NodePre n;
        n.ni = make_uint2(0u, 0u); n.d1 = make_double2(0.0, 0.0); n.c4 = make_uint2(0u, 0u);
        if (j < jend) {
          n.ni = lds_v2u32(sb + (L.yI + 8u * j));
          n.d1 = lds_v2f64(sb + (L.yD1 + 16u * j));
          n.c4 = lds_v2u32(sb + L.yC + 2u * (n.ni.x >> 8));
        }
        return n;
      };
      uint32_t jbeg = lds_u32(sb + L.yLev);
      uint32_t jend = lds_u32(sb + (L.yLev + 4u));
      NodePre nx = fetch(jbeg + slot_id, jend);
#ifdef ROWS_ABL_NO_B
      for (uint32_t ly = 0; ly < 1; ++ly) {
#else
      for (uint32_t ly = 0; ly < ry.nlev; ++ly) {
#endif
        for (uint32_t j = jbeg + slot_id; j < jend; j += step) {
          const NodePre n = nx;
          if (j + step < jend) nx = fetch(j + step, jend);
          const uint32_t yl = n.ni.y >> 16, deg = n.ni.y & 0xffffu;
          if (yl < len_lo) {  // G1 == 0 here and below
            if (valid) sts_f64(hbase + ROWB * j, 0.0);
            continue;
          }
          const bool in_band = yl <= len_hi;
          double S0 = 0.0, S1 = 0.0, m = 0.0;
          uint32_t e = sb + L.yC + 2u * (n.ni.x >> 8) + 8u;   // past the four prefetched children
          const uint32_t eend = e - 8u + 2u * deg;
          if (in_band) {
            double R0 = 0.0, R1 = 0.0;
            if (deg > 0u) {
              const uint32_t o0 = (n.c4.x & 0xffffu) * RP, o1 = (n.c4.x >> 16) * RP;
              S0 += lds_f64(hbase + o0); R0 += lds_f64(hqbase + o0);
              S1 += lds_f64(hbase + o1); R1 += lds_f64(hqbase + o1);
            }
            if (deg > 2u) {
              const uint32_t o0 = (n.c4.y & 0xffffu) * RP, o1 = (n.c4.y >> 16) * RP;
              S0 += lds_f64(hbase + o0); R0 += lds_f64(hqbase + o0);
              S1 += lds_f64(hbase + o1); R1 += lds_f64(hqbase + o1);
            }
            for (; e < eend; e += 4u) {
              const uint32_t c2 = lds_u32(e);
              const uint32_t o0 = (c2 & 0xffffu) * RP, o1 = (c2 >> 16) * RP;
              S0 += lds_f64(hbase + o0); R0 += lds_f64(hqbase + o0);
              S1 += lds_f64(hbase + o1); R1 += lds_f64(hqbase + o1);
            }
            const double2 b0v = lds_v2f64(sb + (L.yB + 32u * j));        // {s2_y, el_y}
            const double2 b1v = lds_v2f64(sb + (L.yB + 32u * j + 16u));  // {paths_y, bfreq_y}
            const double vs = lds_f64(tabrow + 8u * (n.ni.x & 0xffu)) * xbf * b1v.y;
            m = vs * fma(b0v.y, xql, b0v.x * (R0 + R1));
            racc = fma(b1v.x, m, racc);
          } else {
            if (deg > 0u) {
              S0 += lds_f64(hbase + (n.c4.x & 0xffffu) * RP);
              S1 += lds_f64(hbase + (n.c4.x >> 16) * RP);
            }
            if (deg > 2u) {
              S0 += lds_f64(hbase + (n.c4.y & 0xffffu) * RP);
              S1 += lds_f64(hbase + (n.c4.y >> 16) * RP);
            }
            for (; e < eend; e += 4u) {
              const uint32_t c2 = lds_u32(e);
              S0 += lds_f64(hbase + (c2 & 0xffffu) * RP);
              S1 += lds_f64(hbase + (c2 >> 16) * RP);
            }
          }
          sts_f64(hbase + ROWB * j, fma(n.d1.x, S0 + S1, n.d1.y * m));   // up_y * (M + a_y*s2_y*sum)
        }
        jbeg = jend;
        if (ly + 1u < ry.nlev) {   // next level's node record and first children, fetched ahead of the barrier
          jend = lds_u32(sb + (L.yLev + 4u * ly + 8u));
          nx = fetch(jbeg + slot_id, jend);
        } else {
          sts_f64(tbase + 8u * (tw * 32u + lane), racc);
        }
        team_sync(bar_id, TT);
      }

      // ---- phase C: finished rows G0s(i,:) = up_x(i) * dn_y * (H + a_x*HQ) to the slab; row sums; publish
      for (uint32_t r = tw; r < cnt; r += TW) {
        const uint32_t i = row0 + r;
        const XNode* __restrict__ xc = X.xnode + ps.node0 + i;
        const double xa2 = __ldg(&xc->a), xup = __ldg(&xc->up), xpaths = __ldg(&xc->paths);
        double* __restrict__ g0row = G0 + (size_t)i * NYS;
        const uint32_t h2 = tH + 8u * r, hq2 = tHQ + 8u * r;
#ifdef ROWS_ABL_NO_C
        for (uint32_t j = lane; j < 32u; j += 32u)
#else
        for (uint32_t j = lane; j < Ny; j += 32u)
#endif
          __stcg(g0row + j, xup * (lds_f64(sb + (L.yDn + 8u * j)) * fma(xa2, lds_f64(hq2 + ROWB * j), lds_f64(h2 + ROWB * j))));
        // partial sums of this row: thread (warp w, slot q', row r) = part[w*32 + q'*R + r], added in a fixed order
        double t2 = 0.0;
        for (uint32_t u = 0; u < TW * NS; ++u) t2 += lds_f64(tbase + 8u * ((u / NS) * 32u + (u % NS) * R + r));
        if (lane == 0) __stcg(rowacc0 + (size_t)sl * P.nx_cap + i, xpaths * t2);
        __threadfence_block();
        __syncwarp();
        if (lane == 0) asm volatile("st.volatile.shared.u8 [%0], %1;" ::"r"(done + i), "r"(1u) : "memory");
      }
    }

    // ---- fixed-order sum of the per-row slots, one warp per pair of the group
    __syncthreads();
    for (uint32_t s = warp; s < g; s += blockDim.x >> 5) {
      const RowSlot ps = s_slot[s];
      double t = 0.0;
      const double* __restrict__ ra = rowacc0 + (size_t)s * P.nx_cap;
      for (uint32_t i = lane; i < ps.N; i += 32u) t += __ldcg(ra + i);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) P.out[ps.k] = t + ps.plr * (double)ry.lr;
    }
  }
}

template <int RSH>
cudaError_t launch_rows_t(const StemRowsLaunch& p, int grid, int nthreads, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(stem_rows_kernel<RSH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  stem_rows_kernel<RSH><<<grid, nthreads, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace

size_t stem_rows_smem_bytes(uint32_t rows, uint32_t nteams, uint32_t team_warps, uint32_t nx_cap, uint32_t ny_cap,
                            uint32_t e4_cap, uint32_t ylev_cap, uint32_t xlev_cap) {
  return rows_layout(rows, nteams, team_warps, nx_cap, ny_cap, e4_cap, ylev_cap, xlev_cap).total;
}

cudaError_t launch_stem_rows(const StemRowsLaunch& p, uint32_t rows, int grid, int nteams, size_t smem, cudaStream_t stream) {
  const int nthreads = nteams * (int)p.team_warps * 32;
  switch (rows) {
    case 16: return launch_rows_t<4>(p, grid, nthreads, smem, stream);
    case 8: return launch_rows_t<3>(p, grid, nthreads, smem, stream);
    case 4: return launch_rows_t<2>(p, grid, nthreads, smem, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace stemk

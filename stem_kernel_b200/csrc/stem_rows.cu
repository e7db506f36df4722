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
  uint32_t tab, yN, yB, yDn, yC, yLev, cnt, start, pref, blk, done, teams, team_bytes, part, H, HQ, total;
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
  L.yN = take(32 * ny_cap);     // {rest of the child list << 8 | bcode, deg | len << 16, first four children, up*a*s2, up}
  L.yB = take(32 * ny_cap);     // {s2, el, paths, bfreq}   (MATCH cells only)
  L.yDn = take(8 * ny_cap);     // dn
  L.yC = take(2 * e4_cap);      // child lists as tile offsets (pitch * child number, 16 bit), padded with the zero row
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
__device__ __forceinline__ void team_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct RowSlot {         // one pair of the group in flight
  uint32_t k;            // pair number (index into xi / yi / out)
  uint32_t N, node0, lev0, nlev, shift;   // shift = maxlev - nlev: the pair joins at round `shift`
  double plr;
};

struct NodePre {         // what a lane needs about its next y node (one 32-byte record), fetched a level ahead
  uint4 ni;              // {child list offset << 8 | bcode, deg | len << 16, children 0-1, children 2-3 (tile offsets)}
  double2 d1;            // {up*a*s2, up}
};
__device__ __forceinline__ uint4 lds_v4u32(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void sts_v4u32(uint32_t a, uint4 v) { asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }

#ifdef ROWS_PROF
#define PROF_T(var) const long long var = clock64()
#define PROF_ADD(k, a, b) do { if (tw == 0 && lane == 0) prof_acc[k] += (unsigned long long)((b) - (a)); } while (0)
#else
#define PROF_T(var)
#define PROF_ADD(k, a, b)
#endif

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
  __shared__ __align__(8) unsigned long long s_mbar[16];   // one transaction barrier per team (phase A's bulk copies)
  const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar[team]);
  uint32_t mphase = 0;
  if (tw == 0 && lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

#ifdef ROWS_PROF
  unsigned long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned long long prof_acc2[2] = {0, 0};
  unsigned long long prof_acc3[3] = {0, 0, 0};
  unsigned long long prof_acc4[2] = {0, 0};
  const long long prof_k0 = clock64();
#endif
  for (;;) {
    PROF_T(pg0);
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
      sts_v2f64(sb + (L.yB + 32 * j), make_double2(nb.s2, nb.el));
      sts_v2f64(sb + (L.yB + 32 * j + 16), make_double2(nb.paths, nb.bfreq));
      sts_f64(sb + (L.yDn + 8 * j), Y.dn[gy]);
      const NodeI ni = Y.nodei[gy];
      const uint32_t deg = Y.coff[ry.coff0 + j + 1] - Y.coff[ry.coff0 + j];
      // first four children as tile offsets (c16 holds 8 * child; the padding entries 8 * N are the zero row)
      uint2 c4 = make_uint2(0u, 0u);
      if (deg > 0u) c4 = *reinterpret_cast<const uint2*>(Y.c16 + ry.c16_0 + (ni.e4_bcode >> 8));
      const uint32_t zero2 = (8u * Ny * RP) | ((8u * Ny * RP) << 16);
      const uint32_t c01 = deg > 0u ? ((c4.x & 0xffffu) * RP) | (((c4.x >> 16) * RP) << 16) : zero2;
      const uint32_t c23 = deg > 2u ? ((c4.y & 0xffffu) * RP) | (((c4.y >> 16) * RP) << 16) : zero2;
      sts_v4u32(sb + (L.yN + 32 * j), make_uint4(ni.e4_bcode, deg | ((uint32_t)ni.len << 16), c01, c23));
      sts_v2f64(sb + (L.yN + 32 * j + 16), make_double2(yup * (Y.a[gy] * nb.s2), yup));
    }
    for (uint32_t e = tid; e < ry.e4; e += blockDim.x) sts_u16(sb + (L.yC + 2 * e), (uint32_t)Y.c16[ry.c16_0 + e] * RP);
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
    PROF_T(pg1);
    PROF_ADD(0, pg0, pg1);   // group setup

    for (;;) {
      PROF_T(pt0);
      if (tw == 0 && lane == 0) s_tk[team] = atomicAdd(&s_ticket, 1u);
      team_sync(bar_id, TT);
      const uint32_t t = *reinterpret_cast<volatile uint32_t*>(&s_tk[team]);
      if (t >= nblocks) break;
      const uint32_t bw = lds_u32(sb + (L.blk + 4u * t));
      const uint32_t sl = bw & 15u, row0 = (bw >> 4) & 0xffffu, cnt = bw >> 20;
      const RowSlot& ps = s_slot[sl];
      double* __restrict__ G0 = slab + sl * slot_stride;
      const uint32_t done = sb + L.done + sl * P.nx_cap;

      PROF_T(pa0);
      PROF_ADD(1, pt0, pa0);   // ticket
      // ---- phase A: HQ(:, r) = up_y * s2_x(i) * sum over inner pairs c of G0s(c, :)
      // The finished rows of the inner pairs are fetched by TMA bulk copies (one 8*NYS-byte copy per (row, inner pair)
      // edge, issued by the lanes of the team's first warp, completion on the team's mbarrier) into the team's H
      // tile, which is free until phase B; a warp's own loads cannot do this: a warp sustains only ~2-4 B/clk of
      // global loads however many it keeps in flight (scripts/micro/rowsum.cu).
      uint32_t e0r = 0, degr = 0;
      if (lane < cnt) {
        const uint4 xi4 = __ldg(reinterpret_cast<const uint4*>(X.xnode + ps.node0 + row0 + lane) + 3);
        e0r = xi4.x; degr = xi4.y - xi4.x;
      }
      uint32_t pre = degr;   // inclusive prefix of the rows' inner-pair counts over lanes 0..R-1
#pragma unroll
      for (uint32_t o = 1; o < R; o <<= 1) { const uint32_t t2 = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += t2; }
      const uint32_t etot = __shfl_sync(0xffffffffu, pre, R - 1u);
      const uint32_t row_bytes = 8u * NYS;
      const uint32_t kst = min(32u, (ROWB * (P.ny_cap + 1u)) / row_bytes);   // rows the H tile can stage
      for (uint32_t r = tw; r < cnt; r += TW)   // rows whose inner pairs are all hairpin loops: Q == 0
        if (__shfl_sync(0xffffffffu, degr, r) == 0u)
          for (uint32_t j = lane; j < Ny; j += 32u) sts_f64(tHQ + 8u * r + ROWB * j, 0.0);
      for (uint32_t b0 = 0; b0 < etot; b0 += kst) {
        const uint32_t nb = min(kst, etot - b0);
        team_sync(bar_id, TT);   // the staging rows are free (previous batch summed, previous block's phase C done)
        PROF_T(px0);
        if (tw == 0) {
          const uint32_t x = b0 + lane;
          uint32_t row = 0;
#pragma unroll
          for (uint32_t r = 0; r < R; ++r) row += (x >= __shfl_sync(0xffffffffu, pre, r)) ? 1u : 0u;
          row = min(row, R - 1u);
          const uint32_t e0x = __shfl_sync(0xffffffffu, e0r, row), exx = __shfl_sync(0xffffffffu, pre - degr, row);
          uint32_t c = 0;
          if (lane < nb) {
            c = __ldg(X.cidx + e0x + (x - exx));
            while (ld_flag(done + c) == 0u) __nanosleep(32);   // wait until that row is published
          }
          __syncwarp();
          __threadfence_block();   // acquire: the G0 rows behind the flags just seen
          asm volatile("fence.proxy.async;" ::: "memory");   // ... and hand them (and the staging rows) to the async proxy
          if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(nb * row_bytes) : "memory");
          __syncwarp();
          if (lane < nb)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(tH + lane * row_bytes), "l"(G0 + (size_t)c * NYS), "r"(row_bytes), "r"(mbar) : "memory");
        }
        {
          uint32_t ok = 0;
          while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(mbar), "r"(mphase) : "memory");
          mphase ^= 1u;
        }
#ifdef ROWS_PROF
        const long long px1 = clock64();
        if (tw == 0 && lane == 0) { prof_acc3[0] += (unsigned long long)(px1 - px0); prof_acc3[2] += 1; }
#endif
        for (uint32_t r = tw; r < cnt; r += TW) {   // a warp owns the rows r = tw, tw + TW, ...
          const uint32_t pin = __shfl_sync(0xffffffffu, pre, r), dg = __shfl_sync(0xffffffffu, degr, r);
          const uint32_t ex = pin - dg, lo = max(ex, b0), hi = min(pin, b0 + nb);
          if (lo >= hi) continue;
          const double xs2 = __ldg(&(X.xnode + ps.node0 + row0 + r)->s2);
          const uint32_t hq = tHQ + 8u * r;
          const bool cont = lo > ex, fin = hi == pin;
          for (uint32_t jb = 0; jb < Ny; jb += 128u) {   // four columns per lane: four independent sums
            const uint32_t j = jb + lane;
            const bool p0 = j < Ny, p1 = j + 32u < Ny, p2 = j + 64u < Ny, p3 = j + 96u < Ny;
            double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
            if (cont) {
              if (p0) q0 = lds_f64(hq + ROWB * j);
              if (p1) q1 = lds_f64(hq + ROWB * (j + 32u));
              if (p2) q2 = lds_f64(hq + ROWB * (j + 64u));
              if (p3) q3 = lds_f64(hq + ROWB * (j + 96u));
            }
#ifndef ROWS_ABL_NO_A
            uint32_t src = tH + (lo - b0) * row_bytes + 8u * j;
            for (uint32_t x = lo; x < hi; ++x, src += row_bytes) {
              const double v0 = p0 ? lds_f64(src) : 0.0, v1 = p1 ? lds_f64(src + 256u) : 0.0;
              const double v2 = p2 ? lds_f64(src + 512u) : 0.0, v3 = p3 ? lds_f64(src + 768u) : 0.0;
              q0 += v0; q1 += v1; q2 += v2; q3 += v3;
            }
#endif
            if (fin) {
              if (p0) q0 = lds_f64(sb + (L.yN + 32u * j + 24u)) * (xs2 * q0);
              if (p1) q1 = lds_f64(sb + (L.yN + 32u * (j + 32u) + 24u)) * (xs2 * q1);
              if (p2) q2 = lds_f64(sb + (L.yN + 32u * (j + 64u) + 24u)) * (xs2 * q2);
              if (p3) q3 = lds_f64(sb + (L.yN + 32u * (j + 96u) + 24u)) * (xs2 * q3);
            }
            if (p0) sts_f64(hq + ROWB * j, q0);
            if (p1) sts_f64(hq + ROWB * (j + 32u), q1);
            if (p2) sts_f64(hq + ROWB * (j + 64u), q2);
            if (p3) sts_f64(hq + ROWB * (j + 96u), q3);
          }
        }
#ifdef ROWS_PROF
        { const long long px2 = clock64(); if (tw == 0 && lane == 0) prof_acc3[1] += (unsigned long long)(px2 - px1); }
#endif
      }
      if (etot != 0u) {   // the staged rows ran over the H tile's zero row: put it back once everybody has summed
        team_sync(bar_id, TT);
        if (tw == 0 && lane < R) sts_f64(tH + ROWB * Ny + 8u * lane, 0.0);
      }
      PROF_T(pa1);
      team_sync(bar_id, TT);
      PROF_T(pb0);
      PROF_ADD(2, pa0, pa1);   // phase A (this warp)
      PROF_ADD(3, pa1, pb0);   // wait for the team's other warps

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
      const uint32_t hbase = tH + 8u * rr;
      const uint32_t tabrow = sb + L.tab + 8u * 16u * xbc;
      const uint32_t slot_id = tw * NS + q, step = TW * NS;
      double racc = 0.0;
      const uint32_t hq_off = tHQ - tH;
      const uint32_t yN0 = sb + L.yN, yB0 = sb + L.yB, yC0 = sb + L.yC;
      // One step = one node per slot for this lane's row.  The step is kept as short as possible -- a warp alone on
      // its scheduler runs dependent code at ~5 cycles per instruction, so the level time is the instruction count:
      //   votes (everything below every row's window / some lane in band / some list longer than four) keep the
      //   control flow warp-uniform; rows below their window compute an exact 0 from zeros; slots past the level's
      //   end compute on a zero record and do not store.
      uint32_t jbeg = lds_u32(sb + L.yLev);
      uint32_t jend = lds_u32(sb + (L.yLev + 4u));
      uint4 ni = make_uint4(0u, 0u, 0u, 0u);
      double2 d1 = make_double2(0.0, 0.0);
      if (jbeg + slot_id < jend) { ni = lds_v4u32(yN0 + 32u * (jbeg + slot_id)); d1 = lds_v2f64(yN0 + 32u * (jbeg + slot_id) + 16u); }
#ifdef ROWS_ABL_NO_B
      for (uint32_t ly = 0; ly < 1; ++ly) {
#else
      for (uint32_t ly = 0; ly < ry.nlev; ++ly) {
#endif
        for (uint32_t jw = jbeg + tw * NS; jw < jend;) {   // warp-uniform trip count (wide levels take several steps)
          const uint32_t j = jw + q;
          const bool act = j < jend;
          const uint32_t yl = ni.y >> 16, deg = ni.y & 0xffffu;
          const bool live = act && yl >= len_lo;
#ifdef ROWS_PROF
          const long long ps0 = clock64() + (long long)(ni.y & 0u);
          long long ps1 = ps0, ps2 = ps0;
#endif
          if (__any_sync(0xffffffffu, live)) {
            const bool in_band = live && yl <= len_hi;
            const bool any_band = __any_sync(0xffffffffu, in_band);
            const uint32_t a0 = hbase + (ni.z & 0xffffu), a1 = hbase + (ni.z >> 16);
            const uint32_t a2 = hbase + (ni.w & 0xffffu), a3 = hbase + (ni.w >> 16);
#ifdef ROWS_PROF
            ps1 = clock64() + (long long)(a0 & 0u);
#endif
            double S0 = lds_f64(a0), S1 = lds_f64(a1), S2 = lds_f64(a2), S3 = lds_f64(a3);
            double R0 = 0.0, R1 = 0.0, R2 = 0.0, R3 = 0.0, s2y = 0.0, ely = 0.0, pathsy = 0.0, vs = 0.0;
            if (any_band) {
              R0 = lds_f64(a0 + hq_off); R1 = lds_f64(a1 + hq_off); R2 = lds_f64(a2 + hq_off); R3 = lds_f64(a3 + hq_off);
              if (act) {
                const double2 b0v = lds_v2f64(yB0 + 32u * j);        // {s2_y, el_y}
                const double2 b1v = lds_v2f64(yB0 + 32u * j + 16u);  // {paths_y, bfreq_y}
                s2y = b0v.x; ely = b0v.y; pathsy = b1v.x;
                vs = lds_f64(tabrow + 8u * (ni.x & 0xffu)) * xbf * b1v.y;
              }
            }
            if (__any_sync(0xffffffffu, deg > 4u)) {   // lists are padded to multiples of four with the zero row
              const uint32_t cl = yC0 + 2u * (ni.x >> 8);
              uint32_t gq = 4u;
#pragma unroll 1
              do {
                if (gq < deg) {
                  const uint2 c = lds_v2u32(cl + 2u * gq);
                  const uint32_t o0 = hbase + (c.x & 0xffffu), o1 = hbase + (c.x >> 16);
                  const uint32_t o2 = hbase + (c.y & 0xffffu), o3 = hbase + (c.y >> 16);
                  const double h0 = lds_f64(o0), h1 = lds_f64(o1), h2 = lds_f64(o2), h3 = lds_f64(o3);
                  if (any_band) {
                    const double r0 = lds_f64(o0 + hq_off), r1 = lds_f64(o1 + hq_off), r2 = lds_f64(o2 + hq_off), r3 = lds_f64(o3 + hq_off);
                    R0 += r0; R1 += r1; R2 += r2; R3 += r3;
                  }
                  S0 += h0; S1 += h1; S2 += h2; S3 += h3;
                }
                gq += 4u;
              } while (__any_sync(0xffffffffu, gq < deg));
            }
            double h = d1.x * ((S0 + S1) + (S2 + S3));   // up_y * a_y*s2_y * sum
#ifdef ROWS_PROF
            ps2 = clock64() + (long long)(h == 1.2345e-300);
#endif
            if (any_band) {
              const double m = in_band ? vs * fma(ely, xql, s2y * ((R0 + R1) + (R2 + R3))) : 0.0;
              h = fma(d1.y, m, h);                       // + up_y * M
              if (in_band) racc = fma(pathsy, m, racc);
            }
            if (act) sts_f64(hbase + ROWB * j, h);
          } else if (act) {
            sts_f64(hbase + ROWB * j, 0.0);   // G1 == 0 here and below for every row of the warp
          }
#ifdef ROWS_PROF
          { const long long ps3 = clock64();
            if (tw == 0 && lane == 0) { prof_acc2[0] += (unsigned long long)(ps1 - ps0); prof_acc2[1] += (unsigned long long)(ps2 - ps1);
                                        prof_acc3[0] += 0; prof_acc4[0] += (unsigned long long)(ps3 - ps2); prof_acc4[1] += 1; } }
#endif
          jw += step;
          if (jw < jend) {   // wide level: the next step's record, unpipelined
            ni = make_uint4(0u, 0u, 0u, 0u);
            if (jw + q < jend) { ni = lds_v4u32(yN0 + 32u * (jw + q)); d1 = lds_v2f64(yN0 + 32u * (jw + q) + 16u); }
          }
        }
        // the next level's record is requested ahead of the barrier
        jbeg = jend;
        jend = lds_u32(sb + (L.yLev + 4u * min(ly + 2u, ry.nlev)));
        ni = make_uint4(0u, 0u, 0u, 0u);
        if (jbeg + slot_id < jend) { ni = lds_v4u32(yN0 + 32u * (jbeg + slot_id)); d1 = lds_v2f64(yN0 + 32u * (jbeg + slot_id) + 16u); }
        if (ly + 1u == ry.nlev) sts_f64(tbase + 8u * (tw * 32u + lane), racc);
        team_sync(bar_id, TT);
      }

      PROF_T(pc0);
      PROF_ADD(4, pb0, pc0);   // phase B
#ifdef ROWS_PROF
      if (tw == 0 && lane == 0) { prof_acc[6] += 1; prof_acc[7] += ry.nlev; }
#endif
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
      PROF_T(pc1);
      PROF_ADD(5, pc0, pc1);   // phase C
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
#ifdef ROWS_PROF
  if (P.prof != nullptr && tw == 0 && lane == 0) {
    for (int k = 0; k < 8; ++k) atomicAdd(P.prof + k, prof_acc[k]);
    atomicAdd(P.prof + 8, (unsigned long long)(clock64() - prof_k0));
    atomicAdd(P.prof + 9, 1ull);
    atomicAdd(P.prof + 10, prof_acc2[0]);
    atomicAdd(P.prof + 11, prof_acc2[1]);
    atomicAdd(P.prof + 12, prof_acc3[0]);
    atomicAdd(P.prof + 13, prof_acc3[1]);
    atomicAdd(P.prof + 14, prof_acc3[2]);
    atomicAdd(P.prof + 15, prof_acc4[0]);
    atomicAdd(P.prof + 7, 0ull);
    atomicAdd(P.prof + 6, 0ull);
    P.prof[16 + 0] = 0;
    atomicAdd(P.prof + 17, prof_acc4[1]);
  }
#endif
}

#undef PROF_T
#undef PROF_ADD

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

// stem_kernel_b200/csrc/stem_lanes.cu -- the stem (base-pair DAG) kernel, "lanes are rows" mapping.
//
// Same pre-scaled recurrence as stem_fast.cu (StemKernel<ST,D>::operator(), stem_kernel_lite/stem_kernel.cpp:14-95;
// node / edge scores score_table.cpp:56-101,162-201) for REC_FAST records:
//     H (i,j) = up_y(j) * G1(i,j)          HQ(i,j) = up_y(j) * Q(i,j)          G0s(i,j) = up_x(i) * G0(i,j)
//     Q (i,j) = s2_x(i) * sum_cx G0s(cx,j)
//     R (i,j) = s2_y(j) * sum_cy HQ(i,cy) + el_y(j)*ql_x(i)         S(i,j) = s2_y(j) * sum_cy H(i,cy)
//     M = in_band ? v_s*R : 0      G1 = M + a_y(j)*S      G0 = G1 + a_x(i)*Q = dn_y(j) * (H + a_x(i)*HQ)
//     k(x,y) = sum_i paths_x(i) * sum_j paths_y(j) * M(i,j)  (+ plr_x * lr_y)
//
// Mapping.  The y record is the same for every row of every pair of a group, so the sweep over the y DAG is made
// WARP-UNIFORM: a BLOCK of R rows (R = 32, 28, ... 8, the tallest whose tile fits, chosen per size bucket) is held in shared
// memory as a tile [column j][H(0..R-1) | HQ(0..R-1) | pad]; in the sweep a lane owns a ROW of the block and a warp
// owns 32/R NODES of the current y level: the child list of the node is read once per warp (a broadcast), the
// gathers H[child][lane] are contiguous, there is no per-lane control flow, and one warp instruction serves 32
// cells (stem_fast.cu pays every instruction of the sweep per row, with lanes 44 % filled).  The rows of a block
// are independent: they are the rows of ONE DAG level of the x records of a GROUP of up to kLanesGroup pairs that
// share their y record (pair lists are y-major), level by level, so no flags or tickets are needed -- the CTA is
// synchronous:
//   A  warp <-> (row, 128 columns), lanes <-> columns: sums of the finished pre-scaled G0 rows of the inner pairs
//      (coalesced L2 reads of the per-pair slab), written transposed into the tile (odd column stride: no conflicts)
//   B  y level by y level on the first kLB warps with a named barrier; levels in which every row of the block is
//      below its band window (G1 == 0, length-monotone DAGs) are zero-filled without a barrier
//   C  warp <-> (row, 128 columns): the finished rows are scaled and written to the slab
// A row's path-weighted MATCH sum is reduced over the warps in a fixed order, so results are reproducible run to run.
#include "kernels.cuh"

namespace stemk {

namespace {

constexpr uint32_t kLG = kLanesGroup;
constexpr uint32_t kLW = kLanesWarps;
#ifndef STEMK_LANES_BWARPS
#define STEMK_LANES_BWARPS 10
#endif
constexpr uint32_t kLB = STEMK_LANES_BWARPS < STEMK_LANES_WARPS ? STEMK_LANES_BWARPS : STEMK_LANES_WARPS;   // warps that run the y sweep (a level has 10-40 nodes; a barrier over fewer warps is cheaper)
constexpr uint32_t kLCh = 24;   // inner pairs of a row whose slab offsets are staged in shared memory (more: read from the record)

struct LanesLayout {
  uint32_t tab, yD0, yD1, yD2, yD3, yI, yC, yLev, yLmax, tile, total;
};

__host__ __device__ inline LanesLayout lanes_layout(uint32_t R, uint32_t ny_cap, uint32_t e4_cap, uint32_t lev_cap) {
  LanesLayout L;
  uint32_t off = 0;
  auto take = [&](uint32_t bytes) { uint32_t at = off; off += (bytes + 15u) & ~15u; return at; };
  L.tab = take(8 * 256);
  L.yD0 = take(16 * ny_cap);  // {s2, el}     (MATCH cells only)
  L.yD1 = take(16 * ny_cap);  // {up*a*s2, up}
  L.yD2 = take(16 * ny_cap);  // {paths, bfreq}
  L.yD3 = take(8 * ny_cap);   // dn
  L.yI = take(8 * ny_cap);    // NodeI
  L.yC = take(2 * e4_cap);    // child lists, in tile columns (units of (2R+1) doubles)
  L.yLev = take(4 * (lev_cap + 1));
  L.yLmax = take(4 * (lev_cap + 1));
  uint32_t tile = (2u * R + 1u) * 8u * (ny_cap + 1u);   // + the all-zero dummy column
  const uint32_t red = 8u * kLW * 32u;                  // the per-warp partial sums alias the tile between blocks
  if (tile < red) tile = red;
  L.tile = take(tile);
  L.total = off;
  return L;
}

__device__ __forceinline__ double warp_sum_l(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ double2 lds_v2f64(uint32_t a) { double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds_v2u32(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_v2f64(uint32_t a, double2 v) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ void sts_v2u32(uint32_t a, uint2 v) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory"); }

struct LSlot {           // one pair of the group in flight
  uint32_t k;            // pair number (index into xi / yi / out)
  uint32_t N, node0, lev0, nlev;
  double plr;
};

template <int R>
__global__ void __launch_bounds__(32 * kLanesWarps, 1) stem_lanes_kernel(const StemLanesLaunch P) {
  constexpr uint32_t NS = 32u / R;              // nodes of a y level a warp sweeps side by side (lanes >= NS*R idle: R need not divide 32)
  constexpr uint32_t LB = NS == 1u ? kLW : kLB; // warps of the sweep: one node per warp needs them all, two or more per warp half of them
  constexpr uint32_t JS = (2u * R + 1u) * 8u;   // bytes per tile column: H[R], HQ[R], one pad double
  constexpr uint32_t HQO = 8u * R;              // HQ part of a column
  extern __shared__ __align__(16) unsigned char sm[];
  __shared__ unsigned long long s_item;
  __shared__ uint32_t s_g, s_maxlev;
  __shared__ LSlot s_slot[kLG];
  __shared__ uint32_t s_lo[kLG], s_cnt[kLG];    // rows of the current x level, per pair of the group
  // the rows of the block in flight: pair of the group, row, inner pairs (first kLCh staged as slab row offsets)
  __shared__ uint32_t s_rsl[32], s_ri[32], s_re0[32], s_rne[32];
  __shared__ double s_rs2[32], s_ra[32], s_rup[32];
  __shared__ __align__(16) uint32_t s_rch[32 * kLCh];
  const LanesLayout L = lanes_layout(R, P.ny_cap, P.e4_cap, P.lev_cap);
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(sm);

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t band = P.len_band;
  for (uint32_t t = tid; t < 256; t += blockDim.x) sts_f64(sb + (L.tab + 8 * t), P.pair_tab[t]);
  double* __restrict__ slab = P.scratch + (size_t)blockIdx.x * P.scratch_stride;
  const unsigned long long slot_stride = P.scratch_stride / kLG;
  double* __restrict__ rowacc0 = P.rowacc + (size_t)blockIdx.x * kLG * P.nx_cap;
  const SetView& X = P.X;
  const SetView& Y = P.Y;
  const unsigned long long n_items = P.count[P.bucket];
  const uint32_t* __restrict__ order = P.order + P.start[P.bucket];
  unsigned long long item = 0, item_end = 0;  // the CTA's current run of queue positions
  const uint32_t tileb = sb + L.tile;
#ifdef LANES_PROF
  unsigned long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pt = clock64();   // thread 0: stage, A, B, C, red, blocks, levels swept, group tail
#define PROF(k) do { const unsigned long long now_ = clock64(); pf[k] += now_ - pt; pt = now_; } while (0)
#else
#define PROF(k) do { } while (0)
#endif

  // sweep mapping: lane = slot * R + row
  const uint32_t r_b = lane % R, s_b = lane / R;
  const bool lane_ok = lane < NS * R;
  // Base addresses of the sweep as opaque registers: left to itself the compiler rebuilds the whole shared-memory
  // carve-up (constant-bank loads, S2R SR_CgaCtaId, a dozen integer instructions) inside the per-node loop.
  auto pin = [](uint32_t v) { uint32_t o; asm volatile("mov.u32 %0, %1;" : "=r"(o) : "r"(v)); return o; };
  const uint32_t hb = pin(tileb + 8u * r_b);   // H(j, r_b) at hb + JS*j, HQ at + HQO
  const uint32_t aI = pin(sb + L.yI), aD0 = pin(sb + L.yD0), aD1 = pin(sb + L.yD1), aD2 = pin(sb + L.yD2);
  const uint32_t aC = pin(sb + L.yC), aTab = pin(sb + L.tab), aLev = pin(sb + L.yLev), aLmax = pin(sb + L.yLmax);

  for (;;) {
    __syncthreads();  // previous group fully retired (also orders the tab fill on the first trip)
    if (tid == 0) {
      if (item >= item_end) { item = atomicAdd(P.counter, (unsigned long long)kLG); item_end = item + kLG; }
      if (item_end > n_items) item_end = n_items;
      uint32_t g = 0;
      if (item < item_end) {
        const uint32_t y0 = P.yi[order[item]];
        uint32_t maxlev = 0;
        while (item + g < item_end && g < kLG && P.yi[order[item + g]] == y0) {
          const uint32_t k = order[item + g];
          const RecDev rx = X.rec[P.xi[k]];
          LSlot ps;
          ps.k = k; ps.N = rx.N; ps.node0 = rx.node0; ps.lev0 = rx.lev0; ps.nlev = rx.nlev; ps.plr = rx.plr;
          s_slot[g] = ps;
          maxlev = max(maxlev, rx.nlev);
          ++g;
        }
        s_maxlev = maxlev;
        s_item = item;
        item += g;
      }
      s_g = g;
    }
    __syncthreads();
    const uint32_t g = s_g;
    if (g == 0) break;
    const RecDev ry = Y.rec[P.yi[order[s_item]]];
    const uint32_t Ny = ry.N;
    const uint32_t NYS = (Ny + 1u) & ~1u;  // row stride of the G0 slabs
    const uint32_t maxlev = s_maxlev;

    // ---- stage the y record
    for (uint32_t j = tid; j < Ny; j += blockDim.x) {
      const uint32_t gy = ry.node0 + j;
      const double ys2 = Y.s2[gy], yup = Y.up[gy];
      sts_v2f64(sb + (L.yD0 + 16 * j), make_double2(ys2, Y.el[gy]));
      sts_v2f64(sb + (L.yD1 + 16 * j), make_double2(yup * (Y.a[gy] * ys2), yup));   // H = up*a*s2 * sum(H children) (+ up*M)
      sts_v2f64(sb + (L.yD2 + 16 * j), make_double2(Y.paths[gy], Y.bfreq[gy]));
      sts_f64(sb + (L.yD3 + 8 * j), Y.dn[gy]);
      const NodeI ni = Y.nodei[gy];
      sts_v2u32(sb + (L.yI + 8 * j), make_uint2(ni.e4_bcode, (uint32_t)ni.deg4 | ((uint32_t)ni.len << 16)));
    }
    {
      // child lists: byte offset 8*c in a row  ->  tile column c in units of 8 bytes: c * (2R+1)
      const uint16_t* __restrict__ src = Y.c16 + ry.c16_0;
      for (uint32_t e = tid; e < ry.e4; e += blockDim.x) sts_u16(sb + (L.yC + 2 * e), ((uint32_t)src[e] >> 3) * (2u * R + 1u));
    }
    for (uint32_t l = tid; l <= ry.nlev; l += blockDim.x) sts_u32(sb + (L.yLev + 4 * l), Y.lev_off[ry.lev0 + l]);
    __syncthreads();
    for (uint32_t l = tid; l < ry.nlev; l += blockDim.x) {   // longest node of each level (levels skipped by the band window)
      uint32_t mx = 0;
      const uint32_t j1 = lds_u32(sb + (L.yLev + 4 * l + 4));
      for (uint32_t j = lds_u32(sb + (L.yLev + 4 * l)); j < j1; ++j) mx = max(mx, lds_v2u32(sb + (L.yI + 8u * j)).y >> 16);
      sts_u32(sb + (L.yLmax + 4 * l), mx);
    }

    PROF(0);
    for (uint32_t lvl = 0; lvl < maxlev; ++lvl) {
      __syncthreads();   // previous level's rows are in the slab; s_lo / s_cnt free
      if (tid < g) {
        const LSlot ps = s_slot[tid];
        uint32_t lo = 0, hi = 0;
        if (lvl < ps.nlev) { lo = X.lev_off[ps.lev0 + lvl]; hi = X.lev_off[ps.lev0 + lvl + 1]; }
        s_lo[tid] = lo; s_cnt[tid] = hi - lo;
      }
      __syncthreads();
      uint32_t total = 0;
      for (uint32_t p = 0; p < g; ++p) total += s_cnt[p];
      // q-th row of the level step -> (pair of the group, row)
      auto locate = [&](uint32_t q, uint32_t* sl, uint32_t* i) {
        uint32_t p = 0;
        while (p + 1u < g && q >= s_cnt[p]) { q -= s_cnt[p]; ++p; }
        *sl = p; *i = s_lo[p] + q;
      };

      for (uint32_t b0 = 0; b0 < total; b0 += R) {
        const uint32_t cnt = min((uint32_t)R, total - b0);
        // the dummy column (the partial sums of the previous block aliased the tile)
        if (tid < 2u * R) sts_f64(tileb + JS * Ny + 8u * tid, 0.0);
        // the block's rows: one thread per row fetches the row record, then one thread per (row, inner pair) its slab offset
        if (tid < cnt) {
          uint32_t sl, i;
          locate(b0 + tid, &sl, &i);
          const XNode* __restrict__ xn = X.xnode + s_slot[sl].node0 + i;
          const double2 x0 = __ldg(reinterpret_cast<const double2*>(xn));        // {s2, a}
          const double2 x1 = __ldg(reinterpret_cast<const double2*>(xn) + 1);    // {up, ql}
          const uint4 x3 = __ldg(reinterpret_cast<const uint4*>(xn) + 3);        // {e0, e1, len, bcode}
          s_rsl[tid] = sl; s_ri[tid] = i; s_re0[tid] = x3.x; s_rne[tid] = x3.y - x3.x;
          s_rs2[tid] = x0.x; s_ra[tid] = x0.y; s_rup[tid] = x1.x;
        }
        __syncthreads();
        for (uint32_t t = tid; t < cnt * kLCh; t += blockDim.x) {
          const uint32_t r = t / kLCh, c = t % kLCh;
          if (c < s_rne[r]) s_rch[t] = __ldg(X.cidx + s_re0[r] + c) * NYS;
        }

        // ---- this lane's row of the block (sweep mapping)
        const bool live = lane_ok && r_b < cnt;
        uint32_t slB = 0, iB = 0;
        if (live) locate(b0 + r_b, &slB, &iB);
        const XNode* __restrict__ xr = X.xnode + s_slot[slB].node0 + iB;
        const double2 x12 = __ldg(reinterpret_cast<const double2*>(xr) + 1), x22 = __ldg(reinterpret_cast<const double2*>(xr) + 2);
        const uint4 xr4 = __ldg(reinterpret_cast<const uint4*>(xr) + 3);
        const double xql = x12.y, xbf = x22.x, xpaths = x22.y;
        const uint32_t xl = xr4.z, xbc = xr4.w;
        // MATCH needs |len_x - len_y| <= band; below the window G1 is identically 0 (length-monotone DAG)
        const uint32_t len_lo = !live ? 0xffffffffu : ((band != 0u && xl > band) ? xl - band : 0u);
        const uint32_t len_hi = band != 0u ? xl + band : 0xffffffffu;
        const uint32_t blk_lo = __reduce_min_sync(0xffffffffu, len_lo);

        // ---- phase A: HQ(r,:) = up_y * s2_x(i) * sum over inner pairs c of G0s(c,:)
        __syncthreads();   // s_rch complete
        const uint32_t nchunk = (Ny + 127u) >> 7;
        for (uint32_t it = warp; it < cnt * nchunk; it += kLW) {
          const uint32_t r = it % cnt, ch = it / cnt;
          const uint32_t ne_all = s_rne[r], nst = min(ne_all, kLCh);
          const double xs2 = s_rs2[r];
          const uint32_t j = (ch << 7) + lane;
          const double* __restrict__ G0 = slab + s_rsl[r] * slot_stride + j;
          const bool v0 = j < Ny, v1 = j + 32u < Ny, v2 = j + 64u < Ny, v3 = j + 96u < Ny;
          double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
          for (uint32_t c = 0; c < nst; c += 4u) {
            // four inner pairs at a time: sixteen loads in flight per lane before the first add
            const uint4 o4 = *reinterpret_cast<const uint4*>(&s_rch[r * kLCh + c]);
            const bool k1 = c + 1u < nst, k2 = c + 2u < nst, k3 = c + 3u < nst;
            const double* __restrict__ p0 = G0 + o4.x;
            const double* __restrict__ p1 = G0 + o4.y;
            const double* __restrict__ p2 = G0 + o4.z;
            const double* __restrict__ p3 = G0 + o4.w;
            const double a00 = v0 ? __ldcg(p0) : 0.0, a01 = v1 ? __ldcg(p0 + 32) : 0.0, a02 = v2 ? __ldcg(p0 + 64) : 0.0, a03 = v3 ? __ldcg(p0 + 96) : 0.0;
            const double a10 = (k1 && v0) ? __ldcg(p1) : 0.0, a11 = (k1 && v1) ? __ldcg(p1 + 32) : 0.0, a12 = (k1 && v2) ? __ldcg(p1 + 64) : 0.0, a13 = (k1 && v3) ? __ldcg(p1 + 96) : 0.0;
            const double a20 = (k2 && v0) ? __ldcg(p2) : 0.0, a21 = (k2 && v1) ? __ldcg(p2 + 32) : 0.0, a22 = (k2 && v2) ? __ldcg(p2 + 64) : 0.0, a23 = (k2 && v3) ? __ldcg(p2 + 96) : 0.0;
            const double a30 = (k3 && v0) ? __ldcg(p3) : 0.0, a31 = (k3 && v1) ? __ldcg(p3 + 32) : 0.0, a32 = (k3 && v2) ? __ldcg(p3 + 64) : 0.0, a33 = (k3 && v3) ? __ldcg(p3 + 96) : 0.0;
            q0 += a00; q1 += a01; q2 += a02; q3 += a03;
            q0 += a10; q1 += a11; q2 += a12; q3 += a13;
            q0 += a20; q1 += a21; q2 += a22; q3 += a23;
            q0 += a30; q1 += a31; q2 += a32; q3 += a33;
          }
          for (uint32_t c = kLCh; c < ne_all; ++c) {   // rows with more inner pairs than are staged
            const double* __restrict__ src = G0 + __ldg(X.cidx + s_re0[r] + c) * NYS;
            if (v0) q0 += __ldcg(src);
            if (v1) q1 += __ldcg(src + 32);
            if (v2) q2 += __ldcg(src + 64);
            if (v3) q3 += __ldcg(src + 96);
          }
          const uint32_t dst = tileb + HQO + 8u * r + JS * j;
          if (v0) sts_f64(dst, lds_v2f64(sb + (L.yD1 + 16u * j)).y * (xs2 * q0));
          if (v1) sts_f64(dst + 32u * JS, lds_v2f64(sb + (L.yD1 + 16u * (j + 32u))).y * (xs2 * q1));
          if (v2) sts_f64(dst + 64u * JS, lds_v2f64(sb + (L.yD1 + 16u * (j + 64u))).y * (xs2 * q2));
          if (v3) sts_f64(dst + 96u * JS, lds_v2f64(sb + (L.yD1 + 16u * (j + 96u))).y * (xs2 * q3));
        }

        // ---- phase B: sweep the y DAG level by level; lane <-> row, warp <-> NS nodes of the level
#ifdef LANES_PROF
        __syncthreads();
        PROF(1);
        pf[5] += 1;
#endif
        double racc = 0.0;
        __syncthreads();   // HQ complete
        if (warp < LB) {   // the sweep runs on the first LB warps, with its own named barrier
        uint32_t jbeg = lds_u32(aLev);
        for (uint32_t ly = 0; ly < ry.nlev; ++ly) {
          const uint32_t jend = lds_u32(aLev + 4u * ly + 4u);
          if (lds_u32(aLmax + 4u * ly) < blk_lo) {
            // every row of the block is below its window on every node of this level: G1 == 0
            for (uint32_t t = jbeg * R + tid; t < jend * R; t += 32u * LB) sts_f64(tileb + JS * (t / R) + 8u * (t % R), 0.0);
            jbeg = jend;
            continue;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(32 * LB) : "memory");   // the rows of the levels below are complete
#ifdef LANES_PROF
          pf[6] += 1;
#endif
          for (uint32_t jj = jbeg + warp * NS; jj < jend; jj += LB * NS) {
            const uint32_t j = jj + s_b;
            const bool valid = lane_ok && j < jend;
            const uint32_t jc = valid ? j : jend - 1u;
            const uint2 niw = lds_v2u32(aI + 8u * jc);
            const uint32_t yl = niw.y >> 16;
            const bool skip = !valid || yl < len_lo;
            const bool inb = !skip && yl <= len_hi;
            if (__all_sync(0xffffffffu, skip)) {
              if (valid) sts_f64(hb + JS * jc, 0.0);
              continue;
            }
            const bool any_inb = __any_sync(0xffffffffu, inb);
            uint32_t e = aC + 2u * (niw.x >> 8);
            const uint32_t eend = e + 8u * (niw.y & 0xffffu);
            uint2 c4 = lds_v2u32(e);                                 // first four children (the word after a list is readable)
            const double2 d1 = lds_v2f64(aD1 + 16u * jc);   // {up*a*s2, up}
            double S0 = 0.0, S1 = 0.0, m = 0.0;
            if (any_inb) {
              // the node constants of a MATCH cell are requested before the gathers, not after them
              const double2 d0 = lds_v2f64(aD0 + 16u * jc);  // {s2_y, el_y}
              const double2 d2 = lds_v2f64(aD2 + 16u * jc);  // {paths_y, bfreq_y}
              const double tv = lds_f64(aTab + 8u * (xbc * 16u + (niw.x & 0xffu)));
              double R0 = 0.0, R1 = 0.0;
#pragma unroll 1
              while (e < eend) {
                e += 8u;
                const uint2 nx = lds_v2u32(e);                       // next four, in flight during this round's gathers
                const uint32_t a0 = hb + ((c4.x & 0xffffu) << 3), a1 = hb + ((c4.x >> 16) << 3);
                const uint32_t a2 = hb + ((c4.y & 0xffffu) << 3), a3 = hb + ((c4.y >> 16) << 3);
                const double h0 = lds_f64(a0), g0 = lds_f64(a0 + HQO), h1 = lds_f64(a1), g1 = lds_f64(a1 + HQO);
                const double h2 = lds_f64(a2), g2 = lds_f64(a2 + HQO), h3 = lds_f64(a3), g3 = lds_f64(a3 + HQO);
                S0 += h0; R0 += g0; S1 += h1; R1 += g1;
                S0 += h2; R0 += g2; S1 += h3; R1 += g3;
                c4 = nx;
              }
              if (inb) {
                const double vs = tv * xbf * d2.y;
                m = vs * fma(d0.y, xql, d0.x * (R0 + R1));
                racc = fma(d2.x, m, racc);
              }
            } else {
#pragma unroll 1
              while (e < eend) {
                e += 8u;
                const uint2 nx = lds_v2u32(e);
                const double h0 = lds_f64(hb + ((c4.x & 0xffffu) << 3)), h1 = lds_f64(hb + ((c4.x >> 16) << 3));
                const double h2 = lds_f64(hb + ((c4.y & 0xffffu) << 3)), h3 = lds_f64(hb + ((c4.y >> 16) << 3));
                S0 += h0; S1 += h1; S0 += h2; S1 += h3;
                c4 = nx;
              }
            }
            const double h = skip ? 0.0 : fma(d1.x, S0 + S1, d1.y * m);   // up_y * (M + a_y*s2_y*sum)
            if (valid) sts_f64(hb + JS * jc, h);
          }
          jbeg = jend;
        }
        }
        __syncthreads();   // the block's H and HQ are complete
        PROF(2);

        // ---- phase C: finished rows G0s(i,:) = up_x(i) * dn_y * (H + a_x*HQ) -> slab
        for (uint32_t it = warp; it < cnt * nchunk; it += kLW) {
          const uint32_t r = it % cnt, ch = it / cnt;
          const double xa2 = s_ra[r], xup = s_rup[r];
          double* __restrict__ g0row = slab + s_rsl[r] * slot_stride + (size_t)s_ri[r] * NYS;
          const uint32_t jhi = min(Ny, (ch << 7) + 128u);
          for (uint32_t j = (ch << 7) + lane; j < jhi; j += 32u) {
            const uint32_t a = tileb + 8u * r + JS * j;
            g0row[j] = xup * (lds_f64(sb + (L.yD3 + 8u * j)) * fma(xa2, lds_f64(a + HQO), lds_f64(a)));
          }
        }
        __syncthreads();   // tile free
        PROF(3);

        // ---- the rows' path-weighted MATCH sums: partial sums per (warp, node slot), added in a fixed order
        sts_f64(tileb + 8u * (warp * 32u + lane), racc);
        __syncthreads();
        if (tid < cnt) {
          double t = 0.0;
          for (uint32_t w = 0; w < kLW; ++w)
            for (uint32_t s = 0; s < NS; ++s) t += lds_f64(tileb + 8u * (w * 32u + s * R + tid));
          // tid < cnt <= R: this thread's own sweep row is row `tid` (r_b == tid for the first R threads)
          rowacc0[(size_t)slB * P.nx_cap + iB] = xpaths * t;
        }
        __syncthreads();   // partial sums consumed before the next block's dummy column / phase A overwrite them
        PROF(4);
      }
    }

    // ---- fixed-order sum of the per-row slots, one warp per pair of the group
    __syncthreads();
    if (warp < g) {
      const LSlot ps = s_slot[warp];
      double t = 0.0;
      const double* __restrict__ ra = rowacc0 + (size_t)warp * P.nx_cap;
      for (uint32_t i = lane; i < ps.N; i += 32u) t += __ldcg(ra + i);
      t = warp_sum_l(t);
      if (lane == 0) P.out[ps.k] = t + ps.plr * (double)ry.lr;
    }
    PROF(7);
  }
#ifdef LANES_PROF
  if (tid == 0 && P.prof) for (int k = 0; k < 8; ++k) atomicAdd(P.prof + k, pf[k]);
#endif
}

template <int R>
cudaError_t launch_r(const StemLanesLaunch& p, int grid, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(stem_lanes_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  stem_lanes_kernel<R><<<grid, kLanesWarps * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace

size_t stem_lanes_smem_bytes(uint32_t rows, uint32_t ny_cap, uint32_t e4_cap, uint32_t lev_cap) {
  return lanes_layout(rows, ny_cap, e4_cap, lev_cap).total;
}

cudaError_t launch_stem_lanes(const StemLanesLaunch& p, uint32_t rows, int grid, size_t smem, cudaStream_t stream) {
  if (rows == 32) return launch_r<32>(p, grid, smem, stream);
  if (rows == 28) return launch_r<28>(p, grid, smem, stream);
  if (rows == 24) return launch_r<24>(p, grid, smem, stream);
  if (rows == 20) return launch_r<20>(p, grid, smem, stream);
  if (rows == 16) return launch_r<16>(p, grid, smem, stream);
  if (rows == 12) return launch_r<12>(p, grid, smem, stream);
  if (rows == 8) return launch_r<8>(p, grid, smem, stream);
  return cudaErrorInvalidValue;
}

}  // namespace stemk

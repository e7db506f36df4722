// stem_kernel_b200/csrc/stem_fast.cu -- the stem (base-pair DAG) kernel, fast path.
//
// Same restated recurrence as stem_kernel.cu (StemKernel<ST,D>::operator(), stem_kernel_lite/stem_kernel.cpp:14-95;
// node / edge scores score_table.cpp:56-101,162-201), for records whose skip-edge factors are separable
// (REC_FAST, compile_set.cpp):  e(p,c) = g^(len_p-len_c-2) = s2(p) * up(c).  Rows are therefore kept PRE-SCALED,
//     H (i,j) = up_y(j) * G1(i,j)          HQ(i,j) = up_y(j) * Q(i,j)          G0s(i,j) = up_x(i) * G0(i,j)
// and every gather of the recurrence becomes a plain sum over 2-byte child indices:
//     Q (i,j) = s2_x(i) * sum_cx G0s(cx,j)
//     R (i,j) = s2_y(j) * sum_cy HQ(i,cy) + el_y(j)*ql_x(i)         S(i,j) = s2_y(j) * sum_cy H(i,cy)
//     M = in_band ? v_s*R : 0      G1 = M + a_y(j)*S      G0 = G1 + a_x(i)*Q = dn_y(j) * (H + a_x(i)*HQ)
//     k(x,y) = sum_i paths_x(i) * sum_j paths_y(j) * M(i,j)  (+ plr_x * lr_y)
// No FMA operand has to be fetched per edge any more, child lists are padded to multiples of four (index N is an
// all-zero dummy column), so the inner loop is one 8-byte index load + four 8-byte gathers + four adds.
//
// Mapping.  One persistent CTA per SM; launches are bucketed by the size of the staged record so that shared memory
// is sized for the bucket, not for the largest record of the set.  A CTA takes a GROUP of up to kGroup consecutive
// queue positions that share their y record (callers order pair lists y-major), stages that record in shared memory
// once and runs the group's pairs concurrently: a DAG level of one x record has ~14 rows, too few to keep the CTA's
// warps busy on its own.  Warps pull row BLOCKS of the x records -- up to kFastRows rows of one DAG level, precomputed per record --
// from a shared-memory queue, wait on per-row flags until the rows of the block's inner pairs are published, and
// run the block alone:
//   A  per row, lanes <-> columns: sum of the finished pre-scaled G0 rows (coalesced L2 reads of the per-CTA slab)
//   B  y level by y level, lanes <-> (row of the block) x (node of the level): both rows share every index load;
//      only __syncwarp between levels
//   C  per row, lanes <-> columns: the finished row is scaled, written to the slab, fenced, its flag raised
// The path-weighted MATCH sum of a row goes to a per-row shared-memory slot and the slots are added in a fixed
// order at the end of the pair, so a pair's value does not depend on which warp happened to run which row.
#include "kernels.cuh"

namespace stemk {

namespace {

#ifndef STEMK_GROUP
#define STEMK_GROUP 6
#endif
constexpr uint32_t kGroup = STEMK_GROUP;  // pairs sharing one staged y record that a CTA runs concurrently

struct FastLayout {
  uint32_t tab, yD0, yD1, yD2, yD3, yI, yC, yLev, done, rows, row_bytes, total;
};

// nwarps warps, each with kFastRows x (HQ row, H row)
__host__ __device__ inline FastLayout fast_layout(uint32_t nwarps, uint32_t nx_cap, uint32_t ny_cap, uint32_t e4_cap,
                                                  uint32_t lev_cap) {
  FastLayout L;
  uint32_t off = 0;
  auto take = [&](uint32_t bytes) { uint32_t at = off; off += (bytes + 15u) & ~15u; return at; };
  L.tab = take(8 * 256);
  L.yD0 = take(16 * ny_cap);  // {s2, el}     (MATCH cells only)
  L.yD1 = take(16 * ny_cap);  // {up*a*s2, up}
  L.yD2 = take(16 * ny_cap);  // {paths, bfreq}
  L.yD3 = take(8 * ny_cap);   // dn
  L.yI = take(8 * ny_cap);    // NodeI
  L.yC = take(2 * e4_cap);    // child lists
  L.yLev = take(4 * (lev_cap + 1));
  L.done = take(nx_cap * kGroup);        // one byte per row, per pair of the group
  L.row_bytes = (8u * (ny_cap + 1u) + 15u) & ~15u;  // + the dummy column
  L.rows = take(2u * kFastRows * L.row_bytes * nwarps);
  L.total = off;
  return L;
}

__device__ __forceinline__ double warp_sum_all(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint32_t ld_flag_f(uint32_t addr) {
  uint32_t v; asm volatile("ld.volatile.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v;
}

// Shared-memory accessors on raw 32-bit shared addresses.  Going through generic pointers makes the compiler rebuild
// the shared window base (S2R SR_CgaCtaId + LEA) next to the accesses of the inner loops; these keep it to one
// cvta per kernel.  All of them are volatile: they keep their program order around __syncwarp / __syncthreads.
__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ double2 lds_v2f64(uint32_t a) { double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds_v2u32(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ NodeI lds_nodei(uint32_t a) {
  const uint2 v = lds_v2u32(a);
  NodeI n; n.e4_bcode = v.x; n.deg4 = (uint16_t)(v.y & 0xffffu); n.len = (uint16_t)(v.y >> 16);
  return n;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_v2f64(uint32_t a, double2 v) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_v2u32(uint32_t a, uint2 v) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory"); }
__device__ __forceinline__ void sts_nodei(uint32_t a, NodeI n) { sts_v2u32(a, make_uint2(n.e4_bcode, (uint32_t)n.deg4 | ((uint32_t)n.len << 16))); }

struct PairSlot {        // one pair of the group in flight
  uint32_t k;            // pair number (index into xi / yi / out)
  uint32_t N, node0, coff0, blk0, nblk;
  double plr;
};

__global__ void __launch_bounds__(32 * kFastMaxWarps, 1) stem_fast_kernel(const StemFastLaunch P) {
  extern __shared__ __align__(16) unsigned char sm[];
  __shared__ unsigned long long s_item;
  __shared__ uint32_t s_next_blk, s_g, s_maxblk;
  __shared__ PairSlot s_slot[kGroup];
  const uint32_t nwarps = blockDim.x >> 5;
  const FastLayout L = fast_layout(nwarps, P.nx_cap, P.ny_cap, P.e4_cap, P.lev_cap);
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(sm);  // raw 32-bit shared addresses: LDS/STS [reg+imm]

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t band = P.len_band;
  for (uint32_t t = tid; t < 256; t += blockDim.x) sts_f64(sb + (L.tab + 8 * t), P.pair_tab[t]);
  double* __restrict__ slab = P.scratch + (size_t)blockIdx.x * P.scratch_stride;
  const unsigned long long slot_stride = P.scratch_stride / kGroup;
  const SetView& X = P.X;
  const SetView& Y = P.Y;
  const unsigned long long n_items = P.count[P.bucket];
  const uint32_t* __restrict__ order = P.order + P.start[P.bucket];
  // this warp's rows: [r][HQ | H]
  const uint32_t wrows = L.rows + 2u * kFastRows * L.row_bytes * warp;
  unsigned long long item = 0, item_end = 0;  // the CTA's current run of queue positions

  for (;;) {
    __syncthreads();  // previous group fully retired (also orders the tab fill on the first trip)
    if (tid == 0) {
      // take kGroup queue positions at a time; a group = the leading positions that share their y record
      if (item >= item_end) { item = atomicAdd(P.counter, (unsigned long long)kGroup); item_end = item + kGroup; }
      if (item_end > n_items) item_end = n_items;
      uint32_t g = 0;
      if (item < item_end) {
        const uint32_t y0 = P.yi[order[item]];
        uint32_t maxblk = 0;
        while (item + g < item_end && g < kGroup && P.yi[order[item + g]] == y0) {
          const uint32_t k = order[item + g];
          const RecDev rx = X.rec[P.xi[k]];
          PairSlot ps;
          ps.k = k; ps.N = rx.N; ps.node0 = rx.node0; ps.coff0 = rx.coff0; ps.blk0 = rx.blk0; ps.nblk = rx.nblk; ps.plr = rx.plr;
          s_slot[g] = ps;
          maxblk = max(maxblk, rx.nblk);
          ++g;
        }
        s_maxblk = maxblk;
        s_item = item;
        item += g;
      }
      s_g = g;
      s_next_blk = 0;
    }
    __syncthreads();
    const uint32_t g = s_g;
    if (g == 0) break;
    const RecDev ry = Y.rec[P.yi[order[s_item]]];
    const uint32_t Ny = ry.N;
    const uint32_t NYS = (Ny + 1u) & ~1u;  // row stride of the G0 slabs
    const uint32_t n_tickets = g * s_maxblk;

    // ---- stage the y record, clear the row flags
    for (uint32_t j = tid; j < Ny; j += blockDim.x) {
      const uint32_t gy = ry.node0 + j;
      const double ys2 = Y.s2[gy], yup = Y.up[gy];
      sts_v2f64(sb + (L.yD0 + 16 * j), make_double2(ys2, Y.el[gy]));
      sts_v2f64(sb + (L.yD1 + 16 * j), make_double2(yup * (Y.a[gy] * ys2), yup));   // H = up*a*s2 * sum(H children) (+ up*M)
      sts_v2f64(sb + (L.yD2 + 16 * j), make_double2(Y.paths[gy], Y.bfreq[gy]));
      sts_f64(sb + (L.yD3 + 8 * j), Y.dn[gy]);
      sts_nodei(sb + (L.yI + 8 * j), Y.nodei[gy]);
    }
    {
      const uint2* __restrict__ src = reinterpret_cast<const uint2*>(Y.c16 + ry.c16_0);  // c16_0 is a multiple of 4
      for (uint32_t e = tid; e < ry.e4 / 4u; e += blockDim.x) sts_v2u32(sb + (L.yC + 8 * e), src[e]);
    }
    for (uint32_t l = tid; l <= ry.nlev; l += blockDim.x) sts_u32(sb + (L.yLev + 4 * l), Y.lev_off[ry.lev0 + l]);
    for (uint32_t i = tid; i < (g * P.nx_cap + 3u) / 4u; i += blockDim.x) sts_u32(sb + (L.done + 4 * i), 0u);
    // the dummy column of every row of this warp
    if (lane < 2u * kFastRows) sts_f64(sb + (wrows + L.row_bytes * lane + 8u * Ny), 0.0);
    __syncthreads();

    for (;;) {
      // tickets interleave the pairs of the group block by block; a pair's own blocks keep their order, which is
      // all the dependencies need (a block only waits for earlier blocks of the same pair)
      uint32_t t = 0;
      if (lane == 0) t = atomicAdd(&s_next_blk, 1u);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t >= n_tickets) break;
      const uint32_t sl = t % g, b = t / g;
      const PairSlot ps = s_slot[sl];
      if (b >= ps.nblk) continue;
      const uint32_t blk = kFastRows == 1u ? (b | (1u << 16)) : X.blk[ps.blk0 + b];   // one row per block: block b = row b
      const uint32_t i0 = blk & 0xffffu, cnt = blk >> 16;  // cnt in 1..kFastRows
      double* __restrict__ G0 = slab + sl * slot_stride;
      const uint32_t done = L.done + sl * P.nx_cap;
      double* __restrict__ rowacc = P.rowacc + ((size_t)blockIdx.x * kGroup + sl) * P.nx_cap;

      // ---- phase A: HQ(r,:) = up_y * s2_x(i) * sum over inner pairs c of G0s(c,:)
      for (uint32_t r = 0; r < cnt; ++r) {
        const uint32_t i = i0 + r;
        const uint32_t hq = wrows + 2u * L.row_bytes * r;
        const XNode* __restrict__ xn = X.xnode + ps.node0 + i;   // one 64-byte line, the same address for every lane
        const uint4 xi4 = __ldg(reinterpret_cast<const uint4*>(xn) + 3);
        const uint32_t e0 = xi4.x, e1 = xi4.y;
        const double xs2 = __ldg(&xn->s2);
        for (uint32_t eb = e0; eb < e1 || eb == e0; eb += 32u) {
          const uint32_t ne = min(32u, e1 - eb);
          const bool last = eb + 32u >= e1;
          uint32_t off_l = 0u;
          if (lane < ne) {
            const uint32_t c = X.cidx[eb + lane];
            off_l = c * NYS;
            #ifndef ABL_NO_WAIT
            while (ld_flag_f(sb + done + c) == 0u) __nanosleep(32);
#endif  // wait until that row is published
          }
          __syncwarp();
          __threadfence_block();  // acquire: the G0 rows behind the flags just seen
#ifndef STEMK_NO_PREFETCH
          // The slabs of all CTAs together are several times the L2, so about half of these rows come from
          // DRAM: ask for every 128-byte line of every child row at once instead of discovering the misses
          // sixteen loads at a time.
          for (uint32_t tt = 0; tt < ne; ++tt) {
            const double* __restrict__ src = G0 + __shfl_sync(0xffffffffu, off_l, tt);
            for (uint32_t ln = lane * 16u; ln < Ny; ln += 512u) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + ln));
          }
#endif
          for (uint32_t jb = 0; jb < Ny; jb += 128u) {  // uniform trip count: the shuffles below need every lane
            const uint32_t j = jb + lane;
            const bool v0 = j < Ny, v1 = j + 32u < Ny, v2 = j + 64u < Ny, v3 = j + 96u < Ny;
            const bool more = eb != e0;
            double q0 = (v0 && more) ? lds_f64(sb + (hq + 8u * j)) : 0.0;
            double q1 = (v1 && more) ? lds_f64(sb + (hq + 8u * (j + 32u))) : 0.0;
            double q2 = (v2 && more) ? lds_f64(sb + (hq + 8u * (j + 64u))) : 0.0;
            double q3 = (v3 && more) ? lds_f64(sb + (hq + 8u * (j + 96u))) : 0.0;
#pragma unroll 4
            for (uint32_t tt = 0; tt < ne; ++tt) {
              const double* __restrict__ src = G0 + __shfl_sync(0xffffffffu, off_l, tt) + j;
#ifndef ABL_NO_A
              if (v0) q0 += __ldcg(src);
              if (v1) q1 += __ldcg(src + 32);
              if (v2) q2 += __ldcg(src + 64);
              if (v3) q3 += __ldcg(src + 96);
#else
              q0 += (double)(size_t)src * 1e-300;
#endif
            }
            if (last) {  // scale: HQ = up_y(j) * (s2_x * sum)
              if (v0) q0 = lds_v2f64(sb + (L.yD1 + 16u * j)).y * (xs2 * q0);
              if (v1) q1 = lds_v2f64(sb + (L.yD1 + 16u * (j + 32u))).y * (xs2 * q1);
              if (v2) q2 = lds_v2f64(sb + (L.yD1 + 16u * (j + 64u))).y * (xs2 * q2);
              if (v3) q3 = lds_v2f64(sb + (L.yD1 + 16u * (j + 96u))).y * (xs2 * q3);
            }
            if (v0) sts_f64(sb + (hq + 8u * j), q0);
            if (v1) sts_f64(sb + (hq + 8u * (j + 32u)), q1);
            if (v2) sts_f64(sb + (hq + 8u * (j + 64u)), q2);
            if (v3) sts_f64(sb + (hq + 8u * (j + 96u)), q3);
          }
          if (e1 == e0) break;
        }
      }
      __syncwarp();

      // ---- phase B: sweep the y DAG level by level; lanes <-> (row r, node slot s)
      // rows of the block rounded up to 1, 2, 4: lane = row * nslot + slot (all compile-time for one row per warp)
      const uint32_t rsh = kFastRows == 1u ? 0u : (cnt > 2u ? 2u : (cnt > 1u ? 1u : 0u));
      const uint32_t nslot = 32u >> rsh;
      const uint32_t r_raw = lane >> (5u - rsh), slot = lane & (nslot - 1u);
      const bool live = kFastRows == 1u ? true : r_raw < cnt;
      const uint32_t r = live ? r_raw : 0u;
      const uint32_t gx = ps.node0 + i0 + r;
      const XNode* __restrict__ xr = X.xnode + gx;
      const double2 x12 = __ldg(reinterpret_cast<const double2*>(xr) + 1), x22 = __ldg(reinterpret_cast<const double2*>(xr) + 2);
      const uint4 xr4 = __ldg(reinterpret_cast<const uint4*>(xr) + 3);
      const double xql = x12.y, xbf = x22.x, xpaths = x22.y;
      const uint32_t xl = xr4.z, xbc = xr4.w;
      const uint32_t hqrow = wrows + 2u * L.row_bytes * r, hrow = hqrow + L.row_bytes;
      // MATCH needs |len_x - len_y| <= band; below the window G1 is identically 0 (length-monotone DAG)
      const uint32_t len_lo = (band != 0u && xl > band) ? xl - band : 0u;
      const uint32_t len_hi = band != 0u ? xl + band : 0xffffffffu;
      double racc = 0.0;
      uint32_t jbeg = lds_u32(sb + (L.yLev));
#ifdef ABL_NO_B
      for (uint32_t ly = 0; ly < 1; ++ly) {
#else
      for (uint32_t ly = 0; ly < ry.nlev; ++ly) {
#endif
        const uint32_t jend = lds_u32(sb + (L.yLev + 4u * ly + 4u));
        for (uint32_t j = jbeg + slot; live && j < jend; j += nslot) {
          const NodeI ni = lds_nodei(sb + (L.yI + 8u * j));
          const uint32_t yl = ni.len;
          if (yl < len_lo) {  // G1 == 0 here and below
            sts_f64(sb + (hrow + 8u * j), 0.0);
            continue;
          }
          const bool in_band = yl <= len_hi;
          const double2 d1 = lds_v2f64(sb + (L.yD1 + 16u * j));  // {up*a*s2, up}
          uint32_t e = L.yC + 2u * (ni.e4_bcode >> 8);
          const uint32_t eend = e + 8u * ni.deg4;
          double S0 = 0.0, S1 = 0.0, m = 0.0;
          if (in_band) {
            double R0 = 0.0, R1 = 0.0;
#pragma unroll 1
            for (; e < eend; e += 8u) {
              const uint2 c4 = lds_v2u32(sb + (e));
              const uint32_t o0 = c4.x & 0xffffu, o1 = c4.x >> 16, o2 = c4.y & 0xffffu, o3 = c4.y >> 16;
              S0 += lds_f64(sb + (hrow + o0)); R0 += lds_f64(sb + (hqrow + o0));
              S1 += lds_f64(sb + (hrow + o1)); R1 += lds_f64(sb + (hqrow + o1));
              S0 += lds_f64(sb + (hrow + o2)); R0 += lds_f64(sb + (hqrow + o2));
              S1 += lds_f64(sb + (hrow + o3)); R1 += lds_f64(sb + (hqrow + o3));
            }
            const double2 d0 = lds_v2f64(sb + (L.yD0 + 16u * j));  // {s2_y, el_y}
            const double2 d2 = lds_v2f64(sb + (L.yD2 + 16u * j));  // {paths_y, bfreq_y}
            const double vs = lds_f64(sb + (L.tab + 8u * (xbc * 16u + (ni.e4_bcode & 0xffu)))) * xbf * d2.y;
            m = vs * fma(d0.y, xql, d0.x * (R0 + R1));
            racc = fma(d2.x, m, racc);
          } else {
#pragma unroll 1
            for (; e < eend; e += 8u) {
              const uint2 c4 = lds_v2u32(sb + (e));
              S0 += lds_f64(sb + hrow + (c4.x & 0xffffu));
              S1 += lds_f64(sb + hrow + (c4.x >> 16));
              S0 += lds_f64(sb + hrow + (c4.y & 0xffffu));
              S1 += lds_f64(sb + hrow + (c4.y >> 16));
            }
          }
          sts_f64(sb + (hrow + 8u * j), fma(d1.x, S0 + S1, d1.y * m));   // up_y * (M + a_y*s2_y*sum)
        }
        jbeg = jend;
        __syncwarp();
      }
      // per-row path-weighted MATCH sum (lanes of one row are contiguous: reduce inside the half / full warp)

      for (uint32_t o = nslot >> 1; o > 0u; o >>= 1) racc += __shfl_xor_sync(0xffffffffu, racc, o);
      if (slot == 0u && live) rowacc[i0 + r] = xpaths * racc;   // per-row slot in global scratch (L2)

      // ---- phase C: finished rows G0s(i,:) = up_x(i) * dn_y * (H + a_x*HQ), then publish them
      for (uint32_t rr = 0; rr < cnt; ++rr) {
        const uint32_t i = i0 + rr;
        const uint32_t hq2 = wrows + 2u * L.row_bytes * rr, h2 = hq2 + L.row_bytes;
        const XNode* __restrict__ xc = X.xnode + ps.node0 + i;
        const double xa2 = __ldg(&xc->a), xup = __ldg(&xc->up);
        double* __restrict__ g0row = G0 + (size_t)i * NYS;
#ifdef ABL_NO_C
        for (uint32_t j = lane; j < 32u; j += 32u)
#else
        for (uint32_t j = lane; j < Ny; j += 32u)
#endif
          g0row[j] = xup * (lds_f64(sb + (L.yD3 + 8u * j)) * fma(xa2, lds_f64(sb + (hq2 + 8u * j)), lds_f64(sb + (h2 + 8u * j))));
      }
      __threadfence_block();
      __syncwarp();
      if (lane < cnt) asm volatile("st.volatile.shared.u8 [%0], %1;" ::"r"(sb + done + i0 + lane), "r"(1u) : "memory");
    }

    // ---- fixed-order sum of the per-row slots, one warp per pair of the group
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

// ---- pair classification: which kernel / size bucket runs a pair -----------------------------------------
// bucket 0 = general kernel; 1 + b = fast kernel with shared memory sized for bucket b; trivial pairs (an empty DAG
// on either side) are finished here.
__device__ __forceinline__ int pair_bucket(const StemClassify& C, const RecDev& rx, const RecDev& ry) {
  if (rx.N == 0 || ry.N == 0) return -1;
  if (!C.allow_fast || !(rx.flags & REC_FAST) || !(ry.flags & REC_FAST)) return 0;
  for (int b = 0; b < C.n_caps; ++b) if (ry.N <= C.caps[b]) return 1 + b;
  return 0;
}

__global__ void classify_pairs_kernel(const StemClassify C) {
  const unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (k >= C.n_pairs) return;
  const RecDev& rx = C.X.rec[C.xi[k]];
  const RecDev& ry = C.Y.rec[C.yi[k]];
  const int bucket = pair_bucket(C, rx, ry);
  if (bucket < 0) C.out[k] = rx.plr * (double)ry.lr;
  else atomicAdd(&C.count[bucket], 1ull);
}

__global__ void bucket_starts_kernel(const unsigned long long* count, unsigned long long* start,
                                     unsigned long long* counters, int n) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long s = 0;
    for (int b = 0; b < n; ++b) { start[b] = s; s += count[b]; counters[b] = 0; }
  }
}

// One block per bucket walks the pair list in order and appends its own pairs, so every bucket keeps the caller's
// pair order (big pairs first), which is what keeps the tail of each work queue short.
__global__ void bucket_fill_stable_kernel(const StemClassify C) {
  const int bucket = blockIdx.x;
  __shared__ unsigned long long s_base;
  __shared__ uint32_t wsum[32];
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  for (unsigned long long k0 = 0; k0 < C.n_pairs; k0 += blockDim.x) {
    const unsigned long long k = k0 + threadIdx.x;
    int mine = 0;
    if (k < C.n_pairs) mine = pair_bucket(C, C.X.rec[C.xi[k]], C.Y.rec[C.yi[k]]) == bucket;
    const unsigned m = __ballot_sync(0xffffffffu, mine);
    if (lane == 0) wsum[w] = __popc(m);
    __syncthreads();
    uint32_t before = 0, total = 0;
    for (uint32_t q = 0; q < (blockDim.x >> 5); ++q) { if (q < w) before += wsum[q]; total += wsum[q]; }
    if (mine) C.order[C.start[bucket] + s_base + before + __popc(m & ((1u << lane) - 1u))] = (uint32_t)k;
    __syncthreads();
    if (threadIdx.x == 0) s_base += total;
    __syncthreads();
  }
}

}  // namespace

size_t stem_fast_smem_bytes(uint32_t nwarps, uint32_t nx_cap, uint32_t ny_cap, uint32_t e4_cap, uint32_t lev_cap) {
  return fast_layout(nwarps, nx_cap, ny_cap, e4_cap, lev_cap).total;
}

cudaError_t launch_stem_fast(const StemFastLaunch& p, int grid, int nwarps, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(stem_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  stem_fast_kernel<<<grid, nwarps * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

int stem_fast_ctas_per_sm(int nwarps, size_t smem) {
  int n = 0;
  if (cudaFuncSetAttribute(stem_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, stem_fast_kernel, nwarps * 32, smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

cudaError_t launch_classify(const StemClassify& c, int n_buckets, unsigned long long* counters, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(c.count, 0, sizeof(unsigned long long) * n_buckets, stream);
  if (e != cudaSuccess) return e;
  const unsigned blocks = (unsigned)((c.n_pairs + 255) / 256);
  classify_pairs_kernel<<<blocks, 256, 0, stream>>>(c);
  bucket_starts_kernel<<<1, 32, 0, stream>>>(c.count, c.start, counters, n_buckets);
  bucket_fill_stable_kernel<<<n_buckets, 1024, 0, stream>>>(c);
  return cudaGetLastError();
}

}  // namespace stemk

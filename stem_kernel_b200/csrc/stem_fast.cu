// stem_kernel_b200/csrc/stem_fast.cu -- the stem (base-pair DAG) kernel, fast path.
//
// Same restated recurrence as stem_kernel.cu (StemKernel<ST,D>::operator(), stem_kernel_lite/stem_kernel.cpp:14-95;
// node / edge scores score_table.cpp:56-101,162-201), for records whose skip-edge factors are separable
// (REC_FAST, compile_set.cpp):  e(p,c) = g^(len_p-len_c-2) = s2(p) * up(c).  Everything is kept PRE-SCALED,
//     H (i,j) = up_y(j) * G1(i,j)        HQ(i,j) = up_y(j) * Q(i,j)        G0ss(i,j) = up_x(i) * up_y(j) * G0(i,j)
// so that every gather of the recurrence is a plain sum over child offsets and no phase multiplies per column:
//     HQ(i,j) = s2_x(i) * sum_cx G0ss(cx,j)
//     R (i,j) = s2_y(j) * sum_cy HQ(i,cy)/up_y(j) ...   in the scaled form:  up_y*M = v_s*bfreq_y*up_y*(s2_y*sum_cy HQ(i,cy) + el_y*ql_x)
//     H (i,j) = up_y*M(i,j) + up_y*a_y*s2_y * sum_cy H(i,cy)             (M only inside the length band)
//     G0ss(i,j) = up_x(i) * (H(i,j) + a_x(i)*HQ(i,j))
//     k(x,y) = sum_i paths_x(i) * sum_j paths_y(j) * M(i,j)  (+ plr_x * lr_y)
// Child lists are padded to multiples of four (index N is an all-zero dummy column).
//
// A warp computes a BLOCK of kFastRows = 2 rows of one x sub-level.  MATCH (stem_kernel.cpp:46-59) only exists inside
// the length band, ~13 % of the cells, and R only depends on the finished rows of i's inner pairs -- not on the row's
// own sweep -- so it is taken out of the sweep and done with full lanes:
//   A   per row, lanes <-> column pairs: q = sum of the finished G0ss rows of i's inner pairs.  The rows of a pair live
//       in a per-pair slab in global memory (L2); one TMA bulk prefetch per inner pair asks the L2 for the whole row,
//       then coalesced 16-byte reads, the row accumulated in REGISTERS (template on the row width: 7-12 independent
//       loads in flight per lane);  HQ(i,:) = s2_x*q goes to the row buffer, up_x*a_x*HQ straight to the slab
//   B1  per row, lanes <-> the y nodes INSIDE THE BAND of i (a contiguous range of the record's length-sorted node
//       list, found by two warp-wide searches): R gathered from the row buffer, up_y*M(i,j) kept in registers, the
//       path-weighted MATCH sum of the row accumulated here
//   Z   the row buffer is cleared and the band's up_y*M values are dropped into it: it is now the H row
//   B2  y sub-level by sub-level (at most 16 nodes, sorted by length), lanes <-> (row of the block) x (node):
//       H(i,j) += coef_j * sum_cy H(i,cy) for the nodes at or above the band (G1 is identically 0 below it:
//       length-monotone DAG).  Only the sub-levels that hold such a node are visited (one ballot per 32 sub-levels)
//   C   the buffers hold up_x*H (B1 scales the band values by up_x, the sweep is linear): a row is added to its slab row
//       by the TMA unit (one cp.reduce.async.bulk .add.f64 per row, no load/store wavefronts), then the flags are raised
//
// What bounds it (ncu, profiles/r02_*): the LSU data pipe -- shared-memory wavefronts of the gathers (two thirds of
// them in B2) plus the global wavefronts of phases A and C -- at 66-70 % of its peak; time follows the wavefront count
// of a variant, not its instruction count, its warps (16 to 24 give the same throughput) or the latency of a phase.
// Hence: child lists are 16-bit byte offsets (one 8-byte index load per four children) whose order is chosen at upload
// so that the nodes a warp sweeps together hit different banks (compile_set.cpp), rows are polled with one uniform
// vote, the prefetch goes through the TMA unit (a prefetch.global.L2 per line cost an eighth of all wavefronts), no
// phase loads a per-column scale factor, and the band's values never take a round trip through shared memory.
//
// Mapping.  One persistent CTA per SM; launches are bucketed by the size of the staged record so that shared memory
// is sized for the bucket, not for the largest record of the set.  A CTA takes a GROUP of up to kGroup consecutive
// queue positions that share their y record (callers order pair lists y-major), stages that record in shared memory
// once and runs the group's pairs concurrently: a DAG level of one x record has ~14 rows, too few to keep the CTA's
// warps busy on its own.  Warps pull row blocks of the x records -- precomputed per record -- from a shared-memory
// ticket counter, wait on per-row flags until the rows of the block's inner pairs are published, and run the block
// alone.  The path-weighted MATCH sum of a row goes to a per-row slot and the slots are added in a fixed order at the
// end of the pair, so a pair's value does not depend on which warp happened to run which row: results are
// bit-reproducible run to run and independent of the pair's position in a list (which is what makes the multi-device
// matrices of stemk_gram_multi bit-identical to the single-device ones).
#include "kernels.cuh"
#include "stem_common.cuh"

namespace stemk {

namespace {

constexpr uint32_t kGroup = kFastGroup;  // pairs sharing one staged y record that a CTA runs concurrently
constexpr uint32_t R = kFastRows;        // rows of a block (1, 2 or 4)
#ifndef STEMK_A_TWO_PASS
#define STEMK_A_TWO_PASS 0
#endif
#ifndef STEMK_POLL_NS0
#define STEMK_POLL_NS0 128
#endif
#ifndef STEMK_POLL_NSMAX
#define STEMK_POLL_NSMAX 2048
#endif
constexpr uint32_t kBandRegs = 128;      // MATCH values of a row kept in registers (four per lane) instead of the per-warp band buffer
constexpr uint32_t kPollNs0 = STEMK_POLL_NS0, kPollNsMax = STEMK_POLL_NSMAX;   // sleep between two polls of the row flags: first, longest
static_assert(R == 1 || R == 2 || R == 4, "kFastRows must be 1, 2 or 4");

struct FastLayout {
  uint32_t tab, yB2, yB1a, yB1b, yC, yLev, yPerm, done, warps, mbuf_bytes, row_bytes, warp_bytes, total;
};

// per warp: the band buffer, then R row buffers of ny_cap + 1 columns (the last one is the all-zero dummy column)
__host__ __device__ inline FastLayout fast_layout(uint32_t nwarps, uint32_t nx_cap, uint32_t ny_cap, uint32_t e4_cap,
                                                  uint32_t lev_cap, uint32_t band_cap) {
  FastLayout L;
  uint32_t off = 0;
  auto take = [&](uint32_t bytes) { uint32_t at = off; off += (bytes + 15u) & ~15u; return at; };
  L.tab = take(8 * 256);
  L.yB2 = take(16 * ny_cap);   // {child list offset << 8 | bcode, deg4 | len << 16, coef = up*a*s2}   (sweep)
  L.yB1a = take(16 * ny_cap);  // {s2*bfreq*up, el*bfreq*up}                                              (MATCH)
  L.yB1b = take(16 * ny_cap);  // {paths*dn, the two integer words of the sweep record}                   (MATCH)
  L.yC = take(2 * e4_cap);     // child lists as 16-bit byte offsets into a row buffer, padded to multiples of four
  L.yLev = take(8 * (lev_cap + 1));   // per sub-level {first node, end}: one 8-byte broadcast load per step of the sweep
  L.yPerm = take(4 * ny_cap);  // nodes sorted by length: len << 16 | node
  L.done = take(nx_cap * kGroup);  // one byte per row, per pair of the group
  L.mbuf_bytes = band_cap > kBandRegs ? ((8u * band_cap + 15u) & ~15u) : 0u;   // bands of up to kBandRegs nodes stay in registers
  L.row_bytes = (8u * (ny_cap + 2u) + 15u) & ~15u;
  L.warp_bytes = L.mbuf_bytes + R * L.row_bytes;
  L.warps = take(L.warp_bytes * nwarps);
  L.total = off;
  return L;
}

#ifdef FAST_PROF
#define PROF_T(v) const long long v = clock64()
#define PROF_ADD(slot, a, b) do { if (lane == 0) prof_acc[slot] += (b) - (a); } while (0)
#else
#define PROF_T(v) do {} while (0)
#define PROF_ADD(slot, a, b) do {} while (0)
#endif

// NCH > 0: the staged record has at most 64*NCH nodes and phase A keeps the whole row in registers (NCH 16-byte
// accumulators per lane); NCH = 0: any size up to kFastMaxN, phase A walks the row in blocks of 256 columns.
template <int NCH>
__global__ void __launch_bounds__(32 * (NCH > 0 ? kFastRegWarps : kFastMaxWarps), 1) stem_fast_kernel(const StemFastLaunch P) {
  extern __shared__ __align__(16) unsigned char sm[];
  __shared__ unsigned long long s_item;
  __shared__ uint32_t s_next_blk, s_g, s_maxblk;
  __shared__ PairSlot s_slot[kGroup];
  const uint32_t nwarps = blockDim.x >> 5;
  const FastLayout L = fast_layout(nwarps, P.nx_cap, P.ny_cap, P.e4_cap, P.lev_cap, P.band_cap);
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
  // this warp's buffers: [band buffer | row 0 | row 1 ...]; pinned so that they live in registers through the loops
  const uint32_t mbuf = pin(sb + L.warps + L.warp_bytes * warp);
  const uint32_t wrows = pin(mbuf + L.mbuf_bytes);
  const uint32_t row_bytes = pin(L.row_bytes);
  const uint32_t yPerm = pin(sb + L.yPerm), yB2 = pin(sb + L.yB2), yC = pin(sb + L.yC), yLev = pin(sb + L.yLev);
  unsigned long long item = 0, item_end = 0;  // the CTA's current run of queue positions
#ifdef FAST_PROF
  long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long prof_gwait = 0;
  const long long prof_t00 = clock64();
#endif

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
          ps.k = k; ps.N = rx.N; ps.node0 = rx.node0; ps.blk0 = rx.blk0; ps.nblk = rx.nblk; ps.plr = rx.plr;
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
    const uint32_t NYS = (Ny + 1u) & ~1u;  // row stride of the G0 slabs (even: rows are 16-byte aligned)
    const uint32_t n_tickets = g * s_maxblk;

    // ---- stage the y record, clear the row flags
    for (uint32_t j = tid; j < Ny; j += blockDim.x) {
      const uint32_t gy = ry.node0 + j;
      const double ys2 = Y.s2[gy], yup = Y.up[gy], ydn = Y.dn[gy];
      const NodeI ni = Y.nodei[gy];
      const double coef = yup * (Y.a[gy] * ys2);   // H = up*a*s2 * sum(H children) (+ up*M)
      sts_v4u32(yB2 + 16 * j, make_uint4(ni.e4_bcode, (uint32_t)ni.deg4 | ((uint32_t)ni.len << 16),
                                         (uint32_t)__double2loint(coef), (uint32_t)__double2hiint(coef)));
      const double bfu = Y.bfreq[gy] * yup;
      sts_v2f64(sb + (L.yB1a + 16 * j), make_double2(ys2 * bfu, Y.el[gy] * bfu));
      sts_v2f64(sb + (L.yB1b + 16 * j), make_double2(Y.paths[gy] * ydn, __hiloint2double((int)((uint32_t)ni.deg4 | ((uint32_t)ni.len << 16)), (int)ni.e4_bcode)));
      sts_u32(yPerm + 4 * j, Y.lperm[gy]);
    }
    {
      const uint2* __restrict__ src = reinterpret_cast<const uint2*>(Y.c16 + ry.c16_0);  // c16_0 is a multiple of 4
      for (uint32_t e = tid; e < ry.e4 / 4u; e += blockDim.x) { const uint2 v = src[e]; sts_u32(yC + 8 * e, v.x); sts_u32(yC + 8 * e + 4, v.y); }
    }
    for (uint32_t l = tid; l < ry.nlev; l += blockDim.x) { sts_u32(yLev + 8 * l, Y.lev_off[ry.lev0 + l]); sts_u32(yLev + 8 * l + 4, Y.lev_off[ry.lev0 + l + 1]); }
    for (uint32_t i = tid; i < (g * P.nx_cap + 3u) / 4u; i += blockDim.x) sts_u32(sb + (L.done + 4 * i), 0u);
    // the dummy column (and the pad after it) of every row buffer of this warp
    if (lane < 2u * R) sts_f64(wrows + row_bytes * (lane >> 1) + 8u * (Ny + (lane & 1u)), 0.0);
    __syncthreads();

    for (;;) {
      // tickets interleave the pairs of the group block by block; a pair's own blocks keep their order, which is
      // all the dependencies need (a block only waits for earlier blocks of the same pair)
      PROF_T(t_0);
      uint32_t t = 0;
      if (lane == 0) t = atomicAdd(&s_next_blk, 1u);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t >= n_tickets) break;
      const uint32_t sl = t % g, b = t / g;
      const PairSlot ps = s_slot[sl];
      if (b >= ps.nblk) continue;
      const uint32_t blk = R == 1u ? (b | (1u << 16)) : X.blk[ps.blk0 + b];   // one row per block: block b = row b
      const uint32_t i0 = blk & 0xffffu, cnt = blk >> 16;  // cnt in 1..R
      double* __restrict__ G0 = slab + sl * slot_stride;
      const uint32_t done = sb + L.done + sl * P.nx_cap;
      double* __restrict__ rowacc = P.rowacc + ((size_t)blockIdx.x * kGroup + sl) * P.nx_cap;

      for (uint32_t r = 0; r < cnt; ++r) {
        const uint32_t rb = wrows + row_bytes * r;   // this row's buffer
        PROF_T(t_a0);
        const uint32_t i = i0 + r;
        const XNode* __restrict__ xn = X.xnode + ps.node0 + i;   // one 64-byte line, the same address for every lane
        const double2 x01 = __ldg(reinterpret_cast<const double2*>(xn));       // {s2, a}
        const double2 x23 = __ldg(reinterpret_cast<const double2*>(xn) + 1);   // {up, ql}
        const double2 x45 = __ldg(reinterpret_cast<const double2*>(xn) + 2);   // {bfreq, paths}
        const uint4 xi4 = __ldg(reinterpret_cast<const uint4*>(xn) + 3);       // {e0, e1, len, bcode}
        const uint32_t e0 = xi4.x, e1 = xi4.y, xl = xi4.z, xbc = xi4.w;
        const double xs2 = x01.x, pc = x23.x * x01.y;   // pc = up_x * a_x
        double* __restrict__ g0row = G0 + (size_t)i * NYS;

        // ---- phase A: q = sum over inner pairs c of G0s(c,:);  HQ = up_y*s2_x*q -> buffer, up_x*a_x*s2_x*q -> slab
        if constexpr (NCH > 0) {
          // The row lives in registers: lane <-> columns 2*lane, 2*lane+1 of each 64-column chunk; CPR inner pairs
          // x W chunks = 7 to 12 independent 16-byte loads are in flight per lane in every round.
          constexpr int W = STEMK_A_TWO_PASS && NCH > 6 ? (NCH + 1) / 2 : NCH;   // two passes measured slower (-6 %): off
          constexpr int CPR = W <= 4 ? 3 : (W <= 6 ? 2 : 1);
          const uint32_t j = 2u * lane;
#pragma unroll
          for (int c0 = 0; c0 < NCH; c0 += W) {
            double2 q[W];
#pragma unroll
            for (int c = 0; c < W; ++c) q[c] = make_double2(0.0, 0.0);
            for (uint32_t eb = e0; eb < e1; eb += 32u) {
              const uint32_t ne = min(32u, e1 - eb);
              uint32_t c = 0u;
              if (lane < ne) c = X.cidx[eb + lane];
              const uint32_t off_l = c * NYS;
              if (c0 == 0) {
                // wait until the rows of all inner pairs are published: one poll per lane and round, the whole warp
                // sleeps in between, longer every time
                PROF_T(t_w0);
                for (uint32_t ns = kPollNs0; !__all_sync(0xffffffffu, lane >= ne || ld_flag_f(done + c) != 0u); ns = min(2u * ns, kPollNsMax))
                  __nanosleep(ns);
                __threadfence_block();  // acquire: the G0 rows behind the flags just seen
                PROF_T(t_w1);
                PROF_ADD(6, t_w0, t_w1);
                prefetch_rows(G0, off_l, lane < ne, NYS, Ny, lane, ne);
              }
#pragma unroll 1
              for (uint32_t tt = 0; tt < ne; tt += CPR) {
                double2 t[CPR][W];
#pragma unroll
                for (int k = 0; k < CPR; ++k) {
                  const bool ck = tt + k < ne;
                  const double2* __restrict__ sk = reinterpret_cast<const double2*>(G0 + __shfl_sync(0xffffffffu, off_l, (tt + k) & 31u) + j);
#pragma unroll
                  for (int cc = 0; cc < W; ++cc) {
                    t[k][cc] = make_double2(0.0, 0.0);
                    if (c0 + cc < NCH && ck && j + 64u * (c0 + cc) < Ny) t[k][cc] = __ldcg(sk + 32 * (c0 + cc));
                  }
                }
#pragma unroll
                for (int cc = 0; cc < W; ++cc) {
                  double sx = t[0][cc].x, sy = t[0][cc].y;
#pragma unroll
                  for (int k = 1; k < CPR; ++k) { sx += t[k][cc].x; sy += t[k][cc].y; }
                  q[cc].x += sx; q[cc].y += sy;
                }
              }
            }
            // the second column of a pair may be the slab row's pad (j + 1 == Ny, Ny odd): its value is never used, and
            // the buffer's dummy column must stay zero
#pragma unroll
            for (int cc = 0; cc < W; ++cc) {
              const uint32_t jc = j + 64u * (c0 + cc);
              if (c0 + cc < NCH && jc < Ny) {
                const double sx = xs2 * q[cc].x, sy = xs2 * q[cc].y;   // HQ: the slab rows carry up_y already
                *reinterpret_cast<double2*>(g0row + jc) = make_double2(pc * sx, pc * sy);
                sts_v2f64(rb + 8u * jc, make_double2(sx, jc + 1u < Ny ? sy : 0.0));
              }
            }
          }
        } else {
        for (uint32_t eb = e0; eb < e1 || eb == e0; eb += 32u) {
            const uint32_t ne = min(32u, e1 - eb);
            const bool last = eb + 32u >= e1;
            const bool more = eb != e0;
            PROF_T(t_w0);
            uint32_t c = 0u;
            if (lane < ne) c = X.cidx[eb + lane];
            const uint32_t off_l = c * NYS;
            // wait until the rows of all inner pairs are published: one poll per lane and round, the whole warp sleeps
            // in between, longer every time (polling is shared-memory traffic the sweeps of the other warps pay for)
            for (uint32_t ns = kPollNs0; !__all_sync(0xffffffffu, lane >= ne || ld_flag_f(done + c) != 0u); ns = min(2u * ns, kPollNsMax))
              __nanosleep(ns);
            __threadfence_block();  // acquire: the G0 rows behind the flags just seen
            PROF_T(t_w1);
            PROF_ADD(6, t_w0, t_w1);
            prefetch_rows(G0, off_l, lane < ne, NYS, Ny, lane, ne);
            for (uint32_t jb = 0; jb < Ny; jb += 256u) {  // uniform trip count: the shuffles below need every lane
              const uint32_t j = jb + 2u * lane;           // this lane: columns j, j+1 of four 64-column chunks
              const bool v0 = j < Ny, v1 = j + 64u < Ny, v2 = j + 128u < Ny, v3 = j + 192u < Ny;
              const uint32_t rj = rb + 8u * j;
              double2 q0 = make_double2(0.0, 0.0), q1 = q0, q2 = q0, q3 = q0;
              if (more) {
                if (v0) q0 = lds_v2f64(rj);
                if (v1) q1 = lds_v2f64(rj + 512u);
                if (v2) q2 = lds_v2f64(rj + 1024u);
                if (v3) q3 = lds_v2f64(rj + 1536u);
              }
              // two inner pairs x four column chunks per round: eight independent 16-byte loads in flight per lane
#pragma unroll 1
              for (uint32_t tt = 0; tt < ne; tt += 2u) {
                const double2* __restrict__ s0 = reinterpret_cast<const double2*>(G0 + __shfl_sync(0xffffffffu, off_l, tt) + j);
                const double2* __restrict__ s1 = reinterpret_cast<const double2*>(G0 + __shfl_sync(0xffffffffu, off_l, (tt + 1u) & 31u) + j);
                const bool c1 = tt + 1u < ne;
                const double2 z = make_double2(0.0, 0.0);
                double2 t00 = z, t01 = z, t02 = z, t03 = z, t10 = z, t11 = z, t12 = z, t13 = z;
                if (v0) t00 = __ldcg(s0);
                if (v1) t01 = __ldcg(s0 + 32);
                if (v2) t02 = __ldcg(s0 + 64);
                if (v3) t03 = __ldcg(s0 + 96);
                if (v0 && c1) t10 = __ldcg(s1);
                if (v1 && c1) t11 = __ldcg(s1 + 32);
                if (v2 && c1) t12 = __ldcg(s1 + 64);
                if (v3 && c1) t13 = __ldcg(s1 + 96);
                q0.x += t00.x + t10.x; q0.y += t00.y + t10.y;
                q1.x += t01.x + t11.x; q1.y += t01.y + t11.y;
                q2.x += t02.x + t12.x; q2.y += t02.y + t12.y;
                q3.x += t03.x + t13.x; q3.y += t03.y + t13.y;
              }
              if (last) {
                // the second column of a pair may be the slab row's pad (j + 1 == Ny, Ny odd): its value is never used
                // (and the dummy column of the buffer must stay zero)
                if (v0) { const double sx = xs2 * q0.x, sy = xs2 * q0.y;
                        *reinterpret_cast<double2*>(g0row + j) = make_double2(pc * sx, pc * sy); q0 = make_double2(sx, j + 1u < Ny ? sy : 0.0); }
                if (v1) { const double sx = xs2 * q1.x, sy = xs2 * q1.y;
                        *reinterpret_cast<double2*>(g0row + j + 64u) = make_double2(pc * sx, pc * sy); q1 = make_double2(sx, j + 65u < Ny ? sy : 0.0); }
                if (v2) { const double sx = xs2 * q2.x, sy = xs2 * q2.y;
                        *reinterpret_cast<double2*>(g0row + j + 128u) = make_double2(pc * sx, pc * sy); q2 = make_double2(sx, j + 129u < Ny ? sy : 0.0); }
                if (v3) { const double sx = xs2 * q3.x, sy = xs2 * q3.y;
                        *reinterpret_cast<double2*>(g0row + j + 192u) = make_double2(pc * sx, pc * sy); q3 = make_double2(sx, j + 193u < Ny ? sy : 0.0); }
              } else {   // an intermediate round: the dummy column must stay zero
                if (j + 1u >= Ny) q0.y = 0.0;
                if (j + 65u >= Ny) q1.y = 0.0;
                if (j + 129u >= Ny) q2.y = 0.0;
                if (j + 193u >= Ny) q3.y = 0.0;
              }
              if (v0) sts_v2f64(rj, q0);
              if (v1) sts_v2f64(rj + 512u, q1);
              if (v2) sts_v2f64(rj + 1024u, q2);
              if (v3) sts_v2f64(rj + 1536u, q3);
            }
            if (e1 == e0) break;
          }
        }
        __syncwarp();
        PROF_T(t_a1);
        PROF_ADD(1, t_a0, t_a1);

        // ---- phase B1: MATCH on the y nodes inside the band of row i (a range of the length-sorted node list)
        uint32_t lo = 0u, hi = Ny;
        if (band != 0u) {
          lo = xl > band ? count_len_below(yPerm, Ny, xl - band, lane) : 0u;
          hi = count_len_below(yPerm, Ny, xl + band + 1u, lane);
        }
        {
          const double xql = x23.y, xbf = x45.x, xup = x23.x;   // the H row is kept pre-scaled by up_x: the sweep is linear
          const uint32_t tabx = sb + L.tab + 128u * xbc;
          const uint32_t yB1a = sb + L.yB1a, yB1b = sb + L.yB1b;
          double racc = 0.0;
          // up_y * M(i,j) of one band node: R gathered from the row buffer (HQ)
          auto match = [&](uint32_t j) -> double {
            const double2 ya = lds_v2f64(yB1a + 16u * j);   // {s2_y*bfreq_y*up_y, el_y*bfreq_y*up_y}
            const double2 yb = lds_v2f64(yB1b + 16u * j);   // {paths_y*dn_y, record words}
            const uint32_t rx = (uint32_t)__double2loint(yb.y), ry2 = (uint32_t)__double2hiint(yb.y);
            const double tv = lds_f64(tabx + 8u * (rx & 0xffu));
            uint32_t e = yC + 2u * (rx >> 8);
            const uint32_t eend = e + 8u * (ry2 & 0xffffu);
            double R0 = 0.0, R1 = 0.0;
#pragma unroll 1
            for (; e < eend; e += 8u) {
              const uint2 c4 = lds_v2u32(e);
              R0 += lds_f64(rb + (c4.x & 0xffffu)) + lds_f64(rb + (c4.x >> 16));
              R1 += lds_f64(rb + (c4.y & 0xffffu)) + lds_f64(rb + (c4.y >> 16));
            }
            const double mp = (tv * xbf) * fma(ya.y, xql, ya.x * (R0 + R1));
            racc = fma(yb.x, mp, racc);
            return mp;
          };
          if (P.band_cap <= kBandRegs) {
            // the band's values wait in registers while the buffer is cleared: no trip through shared memory
            uint32_t jr[4];
            double mr[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t tq = lo + lane + 32u * k;
              jr[k] = 0xffffffffu; mr[k] = 0.0;
              if (tq < hi) { jr[k] = lds_u32(yPerm + 4u * tq) & 0xffffu; mr[k] = match(jr[k]); }
            }
            __syncwarp();
            // ---- Z: the row buffer becomes the H row: zero, then the band's up_y*M
            for (uint32_t j = 2u * lane; j <= Ny; j += 64u) sts_v2f64(rb + 8u * j, make_double2(0.0, 0.0));
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 4; ++k) if (jr[k] != 0xffffffffu) sts_f64(rb + 8u * jr[k], xup * mr[k]);
          } else {
            for (uint32_t tq = lo + lane; tq < hi; tq += 32u) sts_f64(mbuf + 8u * (tq - lo), match(lds_u32(yPerm + 4u * tq) & 0xffffu));
            __syncwarp();
            for (uint32_t j = 2u * lane; j <= Ny; j += 64u) sts_v2f64(rb + 8u * j, make_double2(0.0, 0.0));
            __syncwarp();
            for (uint32_t tq = lo + lane; tq < hi; tq += 32u)
              sts_f64(rb + 8u * (lds_u32(yPerm + 4u * tq) & 0xffffu), xup * lds_f64(mbuf + 8u * (tq - lo)));
          }
          racc = warp_sum_all(racc);
          if (lane == 0) rowacc[i] = x45.y * racc;   // per-row slot in global scratch (L2)
        }
        __syncwarp();
        PROF_T(t_b1);
        PROF_ADD(2, t_a1, t_b1);
      }

      // ---- phase B2: sweep the y DAG level by level; lanes <-> (row r of the block, node slot s): the rows of the
      // block share every instruction of the sweep
      PROF_T(t_b2a);
      {
        constexpr uint32_t nslot = 32u / R;
        const uint32_t r = lane / nslot, slot = lane % nslot;
        const bool live = r < cnt;
        const uint32_t xl = __ldg(&X.xnode[ps.node0 + i0 + (live ? r : 0u)].len);
        // below the window (len_y + band < len_x) G1 is identically 0: length-monotone DAG
        const uint32_t len_lo = (band != 0u && xl > band) ? xl - band : 0u;
        const uint32_t hrow = wrows + row_bytes * r;
        // Only the sub-levels that hold a node at or above the window of one of the block's rows are visited: their
        // nodes are sorted by length, so the first node decides.  Lanes <-> sub-levels find them 32 at a time (one
        // ballot); the sub-levels before ry.sub1 hold the nodes without inner pairs and have nothing to add.
        uint32_t lo_blk = live ? len_lo : 0xffffffffu;
#pragma unroll
        for (uint32_t o = 16u; o >= nslot; o >>= 1) lo_blk = min(lo_blk, __shfl_xor_sync(0xffffffffu, lo_blk, o));
        const uint32_t nlev = ry.nlev;
        for (uint32_t base = ry.sub1 & ~31u; base < nlev; base += 32u) {
          const uint32_t sl0 = base + lane;
          bool act = sl0 >= ry.sub1 && sl0 < nlev;
          if (act) act = (lds_u32(yB2 + 16u * lds_u32(yLev + 8u * sl0) + 4u) >> 16) >= lo_blk;
          uint32_t mask = __ballot_sync(0xffffffffu, act);
          while (mask != 0u) {
            const uint32_t sl = base + (uint32_t)__ffs((int)mask) - 1u;
            mask &= mask - 1u;
            const uint2 jj = lds_v2u32(yLev + 8u * sl);
            const uint32_t j = jj.x + slot;
            if (live && j < jj.y) {
              const uint4 rec = lds_v4u32(yB2 + 16u * j);
              if ((rec.y >> 16) >= len_lo) {   // below the window G1 is identically 0
                uint32_t e = yC + 2u * (rec.x >> 8);
                const uint32_t eend = e + 8u * (rec.y & 0xffffu);
                const uint32_t hj = hrow + 8u * j;
                const double hv = lds_f64(hj);
                double S0 = 0.0, S1 = 0.0;
#pragma unroll 1
                for (; e < eend; e += 8u) {
                  const uint2 c4 = lds_v2u32(e);
                  S0 += lds_f64(hrow + (c4.x & 0xffffu)) + lds_f64(hrow + (c4.x >> 16));
                  S1 += lds_f64(hrow + (c4.y & 0xffffu)) + lds_f64(hrow + (c4.y >> 16));
                }
                sts_f64(hj, fma(__hiloint2double((int)rec.w, (int)rec.z), S0 + S1, hv));   // up_y*M + up_y*a_y*s2_y*sum
              }
            }
            __syncwarp();
          }
        }
      }
      PROF_T(t_b2b);
      PROF_ADD(3, t_b2a, t_b2b);

      // ---- phase C: G0ss(i,:) += up_x * H(i,:), then publish the rows.  The buffers hold up_x * H already (B1 scales the
      // band values, the sweep is linear), so a row is ADDED to its slab row by the TMA unit -- one bulk reduce-add from
      // shared to global memory per row, no load/store wavefronts (reading the buffer, reading and writing the slab row
      // was an eighth of the kernel's wavefronts).  One add per element onto the value phase A stored: deterministic.
      asm volatile("fence.proxy.async;" ::: "memory");   // this lane's writes (buffer: sweep; slab: phase A) before the async proxy
      __syncwarp();
      if (lane == 0) {
        for (uint32_t rr = 0; rr < cnt; ++rr)
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                       ::"l"(__cvta_generic_to_global(G0 + (size_t)(i0 + rr) * NYS)), "r"(wrows + row_bytes * rr), "r"(8u * NYS) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the adds are performed: the rows may be published, the buffers reused
      }
      __threadfence_block();
      __syncwarp();
      if (lane < cnt) asm volatile("st.volatile.shared.u8 [%0], %1;" ::"r"(done + i0 + lane), "r"(1u) : "memory");
      PROF_T(t_c);
      PROF_ADD(4, t_b2b, t_c);
      PROF_ADD(0, t_0, t_c);
#ifdef FAST_PROF
      if (lane == 0) prof_acc[5] += cnt;
#endif
    }

    // ---- fixed-order sum of the per-row slots, one warp per pair of the group
    PROF_T(t_g0);
    __syncthreads();
    PROF_T(t_g1);
#ifdef FAST_PROF
    if (lane == 0) prof_gwait += t_g1 - t_g0;
#endif
    if (warp < g) {
      const PairSlot ps = s_slot[warp];
      double t = 0.0;
      const double* __restrict__ ra = P.rowacc + ((size_t)blockIdx.x * kGroup + warp) * P.nx_cap;
      for (uint32_t i = lane; i < ps.N; i += 32u) t += __ldcg(ra + i);
      t = warp_sum_all(t);
      if (lane == 0) P.out[ps.k] = t + ps.plr * (double)ry.lr;
    }
  }
#ifdef FAST_PROF
  if (lane == 0 && P.prof) {
    for (int q = 0; q < 8; ++q) atomicAdd(P.prof + q, (unsigned long long)prof_acc[q]);
    atomicAdd(P.prof + 8, (unsigned long long)prof_gwait);
    atomicAdd(P.prof + 9, (unsigned long long)(clock64() - prof_t00));
  }
#endif
}

// ---- pair classification: which kernel / size bucket runs a pair -----------------------------------------
// bucket 0 = general kernel; 1 + b = fast kernel with shared memory sized for bucket b; big_bucket = the unstaged
// general kernel (records of any size); trivial pairs (an empty DAG on either side) are finished here.
__device__ __forceinline__ int pair_bucket(const StemClassify& C, const RecDev& rx, const RecDev& ry) {
  if (rx.N == 0 || ry.N == 0) return -1;
  if (C.allow_fast && (rx.flags & REC_FAST) && (ry.flags & REC_FAST))
    for (int b = 0; b < C.n_caps; ++b) if (ry.N <= C.caps[b]) return 1 + b;
  // general kernel; records its shared-memory carve-up cannot hold go to the unstaged kernel
  if (C.big_bucket >= 0) {
    const uint32_t ey = C.Y.coff[ry.coff0 + ry.N] - C.Y.coff[ry.coff0];
    if (rx.N > C.gen_nx || ry.N > C.gen_ny || ey > C.gen_ey) return C.big_bucket;
  }
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

size_t stem_fast_smem_bytes(uint32_t nwarps, uint32_t nx_cap, uint32_t ny_cap, uint32_t e4_cap, uint32_t lev_cap,
                            uint32_t band_cap) {
  return fast_layout(nwarps, nx_cap, ny_cap, e4_cap, lev_cap, band_cap).total;
}

// kernel variant for a bucket: the register-resident phase A where the staged record fits 64*NCH columns
using FastKernelFn = void (*)(StemFastLaunch);
static FastKernelFn fast_kernel_for(uint32_t ny_cap) {
  const uint32_t nch = (ny_cap + 63u) / 64u;
  if (nch <= 4) return stem_fast_kernel<4>;
  if (nch == 5) return stem_fast_kernel<5>;
  if (nch == 6) return stem_fast_kernel<6>;
  if (nch == 7) return stem_fast_kernel<7>;
  if (nch == 8) return stem_fast_kernel<8>;
  return stem_fast_kernel<0>;
}

int stem_fast_max_warps(uint32_t ny_cap) { return (ny_cap + 63u) / 64u <= 8u ? kFastRegWarps : kFastMaxWarps; }

cudaError_t launch_stem_fast(const StemFastLaunch& p, int grid, int nwarps, size_t smem, cudaStream_t stream) {
  const FastKernelFn fn = fast_kernel_for(p.ny_cap);
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  fn<<<grid, nwarps * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

int stem_fast_ctas_per_sm(uint32_t ny_cap, int nwarps, size_t smem) {
  const FastKernelFn fn = fast_kernel_for(ny_cap);
  int n = 0;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, nwarps * 32, smem) != cudaSuccess) {
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

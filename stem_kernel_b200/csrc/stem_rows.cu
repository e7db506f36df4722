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
// R = 32 / 16 / 8 independent rows -- rows of one DAG level of the x records of the CTA's pair group -- is swept
// in lockstep: a lane owns a row, a warp takes one y node (32/R nodes) at a time for all rows of the block, so
//   * the y-side node record and child list are warp-uniform (broadcast loads),
//   * the gathers H[cy][lane] are conflict-free (rows are the fast index of the shared-memory tiles),
//   * the instructions of a node are shared by R rows, and
//   * nodes below every row's length window cost one vote and one store.
// The whole CTA works on one block at a time (shared memory holds the H and HQ tiles of R rows x Ny columns):
//   A  warp <-> row, lanes <-> columns: HQ row = up_y * s2_x * sum of the finished pre-scaled G0 rows of the row's
//      inner pairs (coalesced L2 reads of the per-pair slab), stored transposed into the tile
//   B  y level by y level, __syncthreads between levels; warps <-> nodes of the level, lanes <-> rows
//   C  warp <-> row, lanes <-> columns: G0s row = up_x * dn_y * (H + a_x*HQ) written to the slab (coalesced)
// Blocks are formed round by round: round k holds level k - (maxlev - nlev_s) of every pair s of the group (levels
// are aligned at the top, so the thin top levels of all pairs share blocks); rows of a round are independent of
// each other, rounds run in order, so every dependency is satisfied by construction: no flags, no spinning.
// Row sums of the path-weighted MATCH terms are reduced in a fixed order: results are bit-reproducible.
#include "kernels.cuh"

namespace stemk {

namespace {

constexpr uint32_t kG = kFastGroup;

struct RowsLayout {
  uint32_t tab, yI, yD1, yC, yLev, pref, start, H, HQ, row_pitch_bytes, total;
};

// R rows per block; row pitch R+1 doubles makes both access directions conflict-free:
// lanes <-> rows at fixed j is contiguous, lanes <-> columns at fixed row strides by (R+1)*8 bytes = 2 banks mod 32.
__host__ __device__ inline RowsLayout rows_layout(uint32_t R, uint32_t ny_cap, uint32_t e4_cap, uint32_t ylev_cap,
                                                  uint32_t xlev_cap) {
  RowsLayout L;
  uint32_t off = 0;
  auto take = [&](uint32_t bytes) { uint32_t at = off; off += (bytes + 15u) & ~15u; return at; };
  L.tab = take(8 * 256);
  L.yI = take(8 * ny_cap);      // {child list offset << 8 | bcode, deg | len << 16}
  L.yD1 = take(16 * ny_cap);    // {up*a*s2, up}
  L.yC = take(2 * e4_cap);      // child lists: 8 * child number (16 bit)
  L.yLev = take(4 * (ylev_cap + 1));
  L.pref = take(2 * (kG + 1) * xlev_cap);   // per round: prefix of row counts over the pairs of the group
  L.start = take(2 * kG * xlev_cap);        // per round and pair: first row of the pair's level
  L.row_pitch_bytes = 8u * (R + 1u);
  L.H = take(L.row_pitch_bytes * ny_cap);
  L.HQ = take(L.row_pitch_bytes * ny_cap);
  L.total = off;
  return L;
}

__device__ __forceinline__ double warp_sum_all(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
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

struct RowSlot {         // one pair of the group in flight
  uint32_t k;            // pair number (index into xi / yi / out)
  uint32_t N, node0, lev0, nlev, shift;   // shift = maxlev - nlev: the pair joins at round `shift`
  double plr;
};

template <int RSH>
__global__ void __launch_bounds__(1024, 1) stem_rows_kernel(const StemRowsLaunch P) {
  constexpr uint32_t R = 1u << RSH;         // rows of a block
  constexpr uint32_t NS = 32u >> RSH;       // y nodes a warp sweeps side by side
  constexpr uint32_t RP = R + 1u;           // row pitch of the tiles, doubles
  constexpr uint32_t ROWB = 8u * RP;
  extern __shared__ __align__(16) unsigned char sm[];
  __shared__ unsigned long long s_item;
  __shared__ uint32_t s_g, s_maxlev;
  __shared__ RowSlot s_slot[kG];
  const uint32_t nwarps = blockDim.x >> 5;
  const RowsLayout L = rows_layout(R, P.ny_cap, P.e4_cap, P.ylev_cap, P.xlev_cap);
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(sm);

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t band = P.len_band;
  for (uint32_t t = tid; t < 256; t += blockDim.x) sts_f64(sb + (L.tab + 8 * t), P.pair_tab[t]);
  double* __restrict__ slab = P.scratch + (size_t)blockIdx.x * P.scratch_stride;
  const unsigned long long slot_stride = P.scratch_stride / kG;
  double* __restrict__ rowacc0 = P.rowacc + (size_t)blockIdx.x * P.rowacc_stride;   // kG x nx_cap row slots
  double* __restrict__ part = rowacc0 + (size_t)kG * P.nx_cap;                         // nwarps x 32 partial sums
  const SetView& X = P.X;
  const SetView& Y = P.Y;
  const unsigned long long n_items = P.count[P.bucket];
  const uint32_t* __restrict__ order = P.order + P.start[P.bucket];
  unsigned long long item = 0, item_end = 0;
  const uint32_t rr = lane & (R - 1u), q = lane >> RSH;   // phase B: this lane's row of the block, node slot

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
    }
    __syncthreads();
    const uint32_t g = s_g;
    if (g == 0) break;
    const uint32_t maxlev = s_maxlev;
    const RecDev ry = Y.rec[P.yi[order[s_item]]];
    const uint32_t Ny = ry.N;
    const uint32_t NYS = (Ny + 3u) & ~3u;  // row stride of the G0 slabs (32-byte sectors)

    // ---- stage the y record and the round tables
    for (uint32_t j = tid; j < Ny; j += blockDim.x) {
      const uint32_t gy = ry.node0 + j;
      const double ys2 = Y.s2[gy], yup = Y.up[gy];
      sts_v2f64(sb + (L.yD1 + 16 * j), make_double2(yup * (Y.a[gy] * ys2), yup));
      const NodeI ni = Y.nodei[gy];
      const uint32_t deg = Y.coff[ry.coff0 + j + 1] - Y.coff[ry.coff0 + j];
      sts_v2u32(sb + (L.yI + 8 * j), make_uint2(ni.e4_bcode, deg | ((uint32_t)ni.len << 16)));
    }
    {
      const uint2* __restrict__ src = reinterpret_cast<const uint2*>(Y.c16 + ry.c16_0);
      for (uint32_t e = tid; e < ry.e4 / 4u; e += blockDim.x) sts_v2u32(sb + (L.yC + 8 * e), src[e]);
    }
    for (uint32_t l = tid; l <= ry.nlev; l += blockDim.x) sts_u32(sb + (L.yLev + 4 * l), Y.lev_off[ry.lev0 + l]);
    for (uint32_t k = tid; k < maxlev; k += blockDim.x) {
      uint32_t acc = 0;
      for (uint32_t s = 0; s < g; ++s) {
        const RowSlot ps = s_slot[s];
        sts_u16(sb + (L.pref + 2u * (k * (kG + 1u) + s)), acc);
        uint32_t r0 = 0;
        if (k >= ps.shift) {
          const uint32_t l = k - ps.shift;
          r0 = X.lev_off[ps.lev0 + l];
          acc += X.lev_off[ps.lev0 + l + 1] - r0;
        }
        sts_u16(sb + (L.start + 2u * (k * kG + s)), r0);
      }
      for (uint32_t s = g; s <= kG; ++s) sts_u16(sb + (L.pref + 2u * (k * (kG + 1u) + s)), acc);
    }
    __syncthreads();

    // (slot, row) of entry t of round k
    auto locate = [&](uint32_t k, uint32_t t, uint32_t& s, uint32_t& i) {
      const uint32_t pr = sb + L.pref + 2u * (k * (kG + 1u));
      s = 0;
#pragma unroll
      for (uint32_t u = 1; u < kG; ++u) s += (t >= lds_u16(pr + 2u * u)) ? 1u : 0u;
      i = lds_u16(sb + (L.start + 2u * (k * kG + s))) + (t - lds_u16(pr + 2u * s));
    };

    for (uint32_t k = 0; k < maxlev; ++k) {
      const uint32_t nrows = lds_u16(sb + (L.pref + 2u * (k * (kG + 1u) + kG)));
      for (uint32_t b0 = 0; b0 < nrows; b0 += R) {
        // ---- phase A: HQ(:, r) = up_y * s2_x(i) * sum over inner pairs c of G0s(c, :)
        for (uint32_t r = warp; r < R && b0 + r < nrows; r += nwarps) {
          uint32_t s, i;
          locate(k, b0 + r, s, i);
          const RowSlot& ps = s_slot[s];
          const XNode* __restrict__ xn = X.xnode + ps.node0 + i;
          const uint4 xi4 = __ldg(reinterpret_cast<const uint4*>(xn) + 3);
          const uint32_t e0 = xi4.x, e1 = xi4.y;
          const double xs2 = __ldg(&xn->s2);
          const double* __restrict__ G0 = slab + s * slot_stride;
          const uint32_t hq = sb + L.HQ + 8u * r;
          for (uint32_t jb = 0; jb < Ny; jb += 256u) {
            const uint32_t j = jb + lane;
            double qv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) qv[u] = 0.0;
            for (uint32_t eb = e0; eb < e1; eb += 32u) {
              const uint32_t ne = min(32u, e1 - eb);
              uint32_t off_l = 0u;
              if (lane < ne) off_l = __ldg(X.cidx + eb + lane) * NYS;
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
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const uint32_t jj = j + 32u * u;
              if (jj < Ny) sts_f64(hq + ROWB * jj, lds_v2f64(sb + (L.yD1 + 16u * jj)).y * (xs2 * qv[u]));
            }
          }
        }
        __syncthreads();

        // ---- phase B: sweep the y DAG level by level; lanes <-> rows of the block, warps <-> nodes of the level
        const bool valid = b0 + rr < nrows;
        uint32_t xl = 0, xbc = 0;
        double xql = 0.0, xbf = 0.0;
        if (valid) {
          uint32_t s, i;
          locate(k, b0 + rr, s, i);
          const XNode* __restrict__ xr = X.xnode + s_slot[s].node0 + i;
          const double2 x12 = __ldg(reinterpret_cast<const double2*>(xr) + 1), x22 = __ldg(reinterpret_cast<const double2*>(xr) + 2);
          const uint4 xr4 = __ldg(reinterpret_cast<const uint4*>(xr) + 3);
          xql = x12.y; xbf = x22.x; xl = xr4.z; xbc = xr4.w;
        }
        // MATCH needs |len_x - len_y| <= band; below the window G1 is identically 0 (length-monotone DAG)
        const uint32_t len_lo = !valid ? 0xffffffffu : ((band != 0u && xl > band) ? xl - band : 0u);
        const uint32_t len_hi = band != 0u ? xl + band : 0xffffffffu;
        const uint32_t hbase = sb + L.H + 8u * rr, hqbase = sb + L.HQ + 8u * rr;
        const uint32_t tabrow = sb + L.tab + 8u * 16u * xbc;
        double racc = 0.0;
        uint32_t jbeg = lds_u32(sb + L.yLev);
        for (uint32_t ly = 0; ly < ry.nlev; ++ly) {
          const uint32_t jend = lds_u32(sb + (L.yLev + 4u * ly + 4u));
#ifdef ROWS_ABL_NO_B
          if (ly + 1u < ry.nlev) { jbeg = jend; continue; }
#endif
          for (uint32_t j = jbeg + warp * NS + q; j < jend; j += nwarps * NS) {
            const uint2 ni = lds_v2u32(sb + (L.yI + 8u * j));
            const uint32_t yl = ni.y >> 16, deg = ni.y & 0xffffu;
            if (yl < len_lo) {  // G1 == 0 here and below
              if (valid) sts_f64(hbase + ROWB * j, 0.0);
              continue;
            }
            const bool in_band = yl <= len_hi;
            const double2 d1 = lds_v2f64(sb + (L.yD1 + 16u * j));  // {up*a*s2, up}
            uint32_t e = sb + L.yC + 2u * (ni.x >> 8);
            const uint32_t eend = e + 2u * deg;
            double S0 = 0.0, S1 = 0.0, m = 0.0;
            if (in_band) {
              double R0 = 0.0, R1 = 0.0;
              for (; e + 4u <= eend; e += 4u) {
                const uint32_t c2 = lds_u32(e);
                const uint32_t o0 = (c2 & 0xffffu) * RP, o1 = (c2 >> 16) * RP;
                S0 += lds_f64(hbase + o0); R0 += lds_f64(hqbase + o0);
                S1 += lds_f64(hbase + o1); R1 += lds_f64(hqbase + o1);
              }
              if (e < eend) {
                const uint32_t o0 = lds_u16(e) * RP;
                S0 += lds_f64(hbase + o0); R0 += lds_f64(hqbase + o0);
              }
              const uint32_t gy = ry.node0 + j;
              const double2 b0v = __ldg(reinterpret_cast<const double2*>(Y.yband + gy));       // {s2_y, el_y}
              const double2 b1v = __ldg(reinterpret_cast<const double2*>(Y.yband + gy) + 1);   // {paths_y, bfreq_y}
              const double vs = lds_f64(tabrow + 8u * (ni.x & 0xffu)) * xbf * b1v.y;
              m = vs * fma(b0v.y, xql, b0v.x * (R0 + R1));
              racc = fma(b1v.x, m, racc);
            } else {
              for (; e + 4u <= eend; e += 4u) {
                const uint32_t c2 = lds_u32(e);
                S0 += lds_f64(hbase + (c2 & 0xffffu) * RP);
                S1 += lds_f64(hbase + (c2 >> 16) * RP);
              }
              if (e < eend) S0 += lds_f64(hbase + lds_u16(e) * RP);
            }
            sts_f64(hbase + ROWB * j, fma(d1.x, S0 + S1, d1.y * m));   // up_y * (M + a_y*s2_y*sum)
          }
          jbeg = jend;
          if (ly + 1u == ry.nlev) __stcg(part + tid, racc);
          __syncthreads();
        }

        // ---- phase C: finished rows G0s(i,:) = up_x(i) * dn_y * (H + a_x*HQ) to the slab; row sums
        for (uint32_t r = warp; r < R && b0 + r < nrows; r += nwarps) {
          uint32_t s, i;
          locate(k, b0 + r, s, i);
          const RowSlot& ps = s_slot[s];
          const XNode* __restrict__ xc = X.xnode + ps.node0 + i;
          const double xa2 = __ldg(&xc->a), xup = __ldg(&xc->up), xpaths = __ldg(&xc->paths);
          double* __restrict__ g0row = slab + s * slot_stride + (size_t)i * NYS;
          const uint32_t h2 = sb + L.H + 8u * r, hq2 = sb + L.HQ + 8u * r;
          const double* __restrict__ ydn = Y.dn + ry.node0;
#ifdef ROWS_ABL_NO_C
          for (uint32_t j = lane; j < 32u; j += 32u)
#else
          for (uint32_t j = lane; j < Ny; j += 32u)
#endif
            __stcg(g0row + j, xup * (__ldg(ydn + j) * fma(xa2, lds_f64(hq2 + ROWB * j), lds_f64(h2 + ROWB * j))));
          // partial sums of this row: thread (w, slot q', row r) = part[w*32 + q'*R + r], added in a fixed order
          double t = 0.0;
          for (uint32_t u = lane; u < nwarps * NS; u += 32u) t += __ldcg(part + (u / NS) * 32u + (u % NS) * R + r);
          t = warp_sum_all(t);
          if (lane == 0) __stcg(rowacc0 + (size_t)s * P.nx_cap + i, xpaths * t);
        }
        __syncthreads();
      }
    }

    // ---- fixed-order sum of the per-row slots, one warp per pair of the group
    for (uint32_t s = warp; s < g; s += nwarps) {
      const RowSlot ps = s_slot[s];
      double t = 0.0;
      const double* __restrict__ ra = rowacc0 + (size_t)s * P.nx_cap;
      for (uint32_t i = lane; i < ps.N; i += 32u) t += __ldcg(ra + i);
      t = warp_sum_all(t);
      if (lane == 0) P.out[ps.k] = t + ps.plr * (double)ry.lr;
    }
  }
}

template <int RSH>
cudaError_t launch_rows_t(const StemRowsLaunch& p, int grid, int nwarps, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(stem_rows_kernel<RSH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  stem_rows_kernel<RSH><<<grid, nwarps * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace

size_t stem_rows_smem_bytes(uint32_t rows, uint32_t ny_cap, uint32_t e4_cap, uint32_t ylev_cap, uint32_t xlev_cap) {
  return rows_layout(rows, ny_cap, e4_cap, ylev_cap, xlev_cap).total;
}

cudaError_t launch_stem_rows(const StemRowsLaunch& p, uint32_t rows, int grid, int nwarps, size_t smem, cudaStream_t stream) {
  switch (rows) {
    case 32: return launch_rows_t<5>(p, grid, nwarps, smem, stream);
    case 16: return launch_rows_t<4>(p, grid, nwarps, smem, stream);
    case 8: return launch_rows_t<3>(p, grid, nwarps, smem, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace stemk

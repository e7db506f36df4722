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
// Mapping.  One CTA per pair, several CTAs per SM (launches are bucketed by the size of the staged record so that
// shared memory is sized for the bucket, not for the largest record of the set).  The y record is staged in shared
// memory.  Warps pull row BLOCKS of the x record -- up to kFastRows rows of one DAG level, precomputed per record --
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

struct FastLayout {
  uint32_t tab, yD0, yD1, yD2, yD3, yI, yC, yLev, done, rowacc, rows, row_bytes, total;
};

// nwarps warps, each with kFastRows x (HQ row, H row)
__host__ __device__ inline FastLayout fast_layout(uint32_t nwarps, uint32_t nx_cap, uint32_t ny_cap, uint32_t e4_cap,
                                                  uint32_t lev_cap) {
  FastLayout L;
  uint32_t off = 0;
  auto take = [&](uint32_t bytes) { uint32_t at = off; off += (bytes + 15u) & ~15u; return at; };
  L.tab = take(8 * 256);
  L.yD0 = take(16 * ny_cap);  // {a, el}
  L.yD1 = take(16 * ny_cap);  // {s2, up}
  L.yD2 = take(16 * ny_cap);  // {paths, bfreq}
  L.yD3 = take(8 * ny_cap);   // dn
  L.yI = take(8 * ny_cap);    // NodeI
  L.yC = take(2 * e4_cap);    // child lists
  L.yLev = take(4 * (lev_cap + 1));
  L.done = take(4 * nx_cap);
  L.rowacc = take(8 * nx_cap);
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

__device__ __forceinline__ uint32_t ld_flag_f(const unsigned char* sm, uint32_t byteoff) {
  return *reinterpret_cast<const volatile uint32_t*>(sm + byteoff);
}

__global__ void __launch_bounds__(256, 2) stem_fast_kernel(const StemFastLaunch P) {
  extern __shared__ __align__(16) unsigned char sm[];
  __shared__ unsigned long long s_pair;
  __shared__ uint32_t s_next_blk;
  const uint32_t nwarps = blockDim.x >> 5;
  const FastLayout L = fast_layout(nwarps, P.nx_cap, P.ny_cap, P.e4_cap, P.lev_cap);
#define SM(T, byteoff) (*reinterpret_cast<T*>(sm + (byteoff)))

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t band = P.len_band;
  for (uint32_t t = tid; t < 256; t += blockDim.x) SM(double, L.tab + 8 * t) = P.pair_tab[t];
  double* __restrict__ G0 = P.scratch + (size_t)blockIdx.x * P.scratch_stride;
  const SetView& X = P.X;
  const SetView& Y = P.Y;
  const unsigned long long n_items = P.count[P.bucket];
  const uint32_t* __restrict__ order = P.order + P.start[P.bucket];
  // this warp's rows: [r][HQ | H]
  const uint32_t wrows = L.rows + 2u * kFastRows * L.row_bytes * warp;

  for (;;) {
    __syncthreads();  // previous pair fully retired (also orders the tab fill on the first trip)
    if (tid == 0) { s_pair = atomicAdd(P.counter, 1ull); s_next_blk = 0; }
    __syncthreads();
    if (s_pair >= n_items) break;
    const uint32_t k = order[s_pair];
    const RecDev rx = X.rec[P.xi[k]];
    const RecDev ry = Y.rec[P.yi[k]];
    const uint32_t Nx = rx.N, Ny = ry.N;
    const uint32_t NYS = (Ny + 1u) & ~1u;  // row stride of the G0 slab

    // ---- stage the y record, clear the row flags
    for (uint32_t j = tid; j < Ny; j += blockDim.x) {
      const uint32_t g = ry.node0 + j;
      SM(double2, L.yD0 + 16 * j) = make_double2(Y.a[g], Y.el[g]);
      SM(double2, L.yD1 + 16 * j) = make_double2(Y.s2[g], Y.up[g]);
      SM(double2, L.yD2 + 16 * j) = make_double2(Y.paths[g], Y.bfreq[g]);
      SM(double, L.yD3 + 8 * j) = Y.dn[g];
      SM(NodeI, L.yI + 8 * j) = Y.nodei[g];
    }
    {
      const uint2* __restrict__ src = reinterpret_cast<const uint2*>(Y.c16 + ry.c16_0);  // c16_0 is a multiple of 4
      for (uint32_t e = tid; e < ry.e4 / 4u; e += blockDim.x) SM(uint2, L.yC + 8 * e) = src[e];
    }
    for (uint32_t l = tid; l <= ry.nlev; l += blockDim.x) SM(uint32_t, L.yLev + 4 * l) = Y.lev_off[ry.lev0 + l];
    for (uint32_t i = tid; i < Nx; i += blockDim.x) SM(uint32_t, L.done + 4 * i) = 0u;
    // the dummy column of every row of this warp
    if (lane < 2u * kFastRows) SM(double, wrows + L.row_bytes * lane + 8u * Ny) = 0.0;
    __syncthreads();

    const uint32_t* __restrict__ xcoff = X.coff + rx.coff0;
    const uint32_t* __restrict__ xblk = X.blk + rx.blk0;

    for (;;) {
      uint32_t b = 0;
      if (lane == 0) b = atomicAdd(&s_next_blk, 1u);
      b = __shfl_sync(0xffffffffu, b, 0);
      if (b >= rx.nblk) break;
      const uint32_t blk = xblk[b];
      const uint32_t i0 = blk & 0xffffu, cnt = blk >> 16;  // cnt in 1..kFastRows

      // ---- phase A: HQ(r,:) = up_y * s2_x(i) * sum over inner pairs c of G0s(c,:)
      for (uint32_t r = 0; r < cnt; ++r) {
        const uint32_t i = i0 + r;
        const uint32_t hq = wrows + 2u * L.row_bytes * r;
        const uint32_t e0 = xcoff[i], e1 = xcoff[i + 1];
        const double xs2 = X.s2[rx.node0 + i];
        for (uint32_t eb = e0; eb < e1 || eb == e0; eb += 32u) {
          const uint32_t ne = min(32u, e1 - eb);
          uint32_t off_l = 0u;
          if (lane < ne) {
            const uint32_t c = X.cidx[eb + lane];
            off_l = c * NYS;
            while (ld_flag_f(sm, L.done + 4u * c) == 0u) __nanosleep(32);  // wait until that row is published
          }
          __syncwarp();
          __threadfence_block();  // acquire: the G0 rows behind the flags just seen
          for (uint32_t jb = 0; jb < Ny; jb += 64u) {  // uniform trip count: the shuffles below need every lane
            const uint32_t j = jb + lane;
            const bool one = j < Ny, two = j + 32u < Ny;
            double q0 = (one && eb != e0) ? SM(double, hq + 8u * j) : 0.0;
            double q1 = (two && eb != e0) ? SM(double, hq + 8u * (j + 32u)) : 0.0;
#pragma unroll 4
            for (uint32_t t = 0; t < ne; ++t) {
              const uint32_t off = __shfl_sync(0xffffffffu, off_l, t);
              if (one) q0 += __ldcg(G0 + off + j);
              if (two) q1 += __ldcg(G0 + off + j + 32u);
            }
            if (one) SM(double, hq + 8u * j) = q0;
            if (two) SM(double, hq + 8u * (j + 32u)) = q1;
          }
          if (e1 == e0) break;
        }
        __syncwarp();
        // scale: HQ = up_y(j) * (s2_x * sum)
        for (uint32_t j = lane; j < Ny; j += 32u)
          SM(double, hq + 8u * j) = SM(double2, L.yD1 + 16u * j).y * (xs2 * SM(double, hq + 8u * j));
      }
      __syncwarp();

      // ---- phase B: sweep the y DAG level by level; lanes <-> (row r, node slot s)
      const uint32_t r = (cnt == 2u) ? (lane >> 4) : 0u;
      const uint32_t slot = (cnt == 2u) ? (lane & 15u) : lane;
      const uint32_t nslot = (cnt == 2u) ? 16u : 32u;
      const uint32_t gx = rx.node0 + i0 + r;
      const double xql = X.ql[gx], xbf = X.bfreq[gx];
      const uint32_t xl = X.len[gx], xbc = X.bcode[gx];
      const uint32_t hqrow = wrows + 2u * L.row_bytes * r, hrow = hqrow + L.row_bytes;
      double racc = 0.0;
      uint32_t jbeg = SM(uint32_t, L.yLev);
      for (uint32_t ly = 0; ly < ry.nlev; ++ly) {
        const uint32_t jend = SM(uint32_t, L.yLev + 4u * ly + 4u);
        for (uint32_t j = jbeg + slot; j < jend; j += nslot) {
          const NodeI ni = SM(NodeI, L.yI + 8u * j);
          const uint32_t yl = ni.len;
          if (band != 0u && yl + band < xl) {  // G1 == 0 here and below (length-monotone DAG)
            SM(double, hrow + 8u * j) = 0.0;
            continue;
          }
          const uint32_t dl = xl > yl ? xl - yl : yl - xl;
          const bool in_band = (band == 0u) || (dl <= band);
          const double2 d0 = SM(double2, L.yD0 + 16u * j);  // {a_y, el_y}
          const double2 d1 = SM(double2, L.yD1 + 16u * j);  // {s2_y, up_y}
          uint32_t e = L.yC + 2u * (ni.e4_bcode >> 8);
          const uint32_t eend = e + 8u * ni.deg4;
          double S0 = 0.0, S1 = 0.0, m = 0.0;
          if (in_band) {
            double R0 = 0.0, R1 = 0.0;
#pragma unroll 1
            for (; e < eend; e += 8u) {
              const uint2 c4 = SM(uint2, e);
              const uint32_t o0 = (c4.x & 0xffffu) * 8u, o1 = (c4.x >> 16) * 8u, o2 = (c4.y & 0xffffu) * 8u, o3 = (c4.y >> 16) * 8u;
              S0 += SM(double, hrow + o0); R0 += SM(double, hqrow + o0);
              S1 += SM(double, hrow + o1); R1 += SM(double, hqrow + o1);
              S0 += SM(double, hrow + o2); R0 += SM(double, hqrow + o2);
              S1 += SM(double, hrow + o3); R1 += SM(double, hqrow + o3);
            }
            const double2 d2 = SM(double2, L.yD2 + 16u * j);  // {paths_y, bfreq_y}
            const double vs = SM(double, L.tab + 8u * (xbc * 16u + (ni.e4_bcode & 0xffu))) * xbf * d2.y;
            m = vs * fma(d0.y, xql, d1.x * (R0 + R1));
            racc = fma(d2.x, m, racc);
          } else {
#pragma unroll 1
            for (; e < eend; e += 8u) {
              const uint2 c4 = SM(uint2, e);
              S0 += SM(double, hrow + (c4.x & 0xffffu) * 8u);
              S1 += SM(double, hrow + (c4.x >> 16) * 8u);
              S0 += SM(double, hrow + (c4.y & 0xffffu) * 8u);
              S1 += SM(double, hrow + (c4.y >> 16) * 8u);
            }
          }
          const double g1 = fma(d0.x, d1.x * (S0 + S1), m);
          SM(double, hrow + 8u * j) = d1.y * g1;
        }
        jbeg = jend;
        __syncwarp();
      }
      // per-row path-weighted MATCH sum (lanes of one row are contiguous: reduce inside the half / full warp)
      if (cnt == 2u) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) racc += __shfl_xor_sync(0xffffffffu, racc, o);
        if ((lane & 15u) == 0u) SM(double, L.rowacc + 8u * (i0 + r)) = X.paths[gx] * racc;
      } else {
        racc = warp_sum_all(racc);
        if (lane == 0u) SM(double, L.rowacc + 8u * i0) = X.paths[gx] * racc;
      }

      // ---- phase C: finished rows G0s(i,:) = up_x(i) * dn_y * (H + a_x*HQ), then publish them
      for (uint32_t rr = 0; rr < cnt; ++rr) {
        const uint32_t i = i0 + rr;
        const uint32_t hq2 = wrows + 2u * L.row_bytes * rr, h2 = hq2 + L.row_bytes;
        const double xa2 = X.a[rx.node0 + i], xup = X.up[rx.node0 + i];
        double* __restrict__ g0row = G0 + (size_t)i * NYS;
        for (uint32_t j = lane; j < Ny; j += 32u)
          g0row[j] = xup * (SM(double, L.yD3 + 8u * j) * fma(xa2, SM(double, hq2 + 8u * j), SM(double, h2 + 8u * j)));
      }
      __threadfence_block();
      __syncwarp();
      if (lane < cnt) *reinterpret_cast<volatile uint32_t*>(sm + L.done + 4u * (i0 + lane)) = 1u;
    }

    // ---- fixed-order sum of the per-row slots
    __syncthreads();
    if (warp == 0) {
      double t = 0.0;
      for (uint32_t i = lane; i < Nx; i += 32u) t += SM(double, L.rowacc + 8u * i);
      t = warp_sum_all(t);
      if (lane == 0) P.out[k] = t + rx.plr * (double)ry.lr;
    }
  }
#undef SM
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

// stem_kernel_b200/csrc/stem_kernel.cu -- the stem (base-pair DAG) kernel, one CTA per pair.
//
// Replaces StemKernel<ST,D>::operator() (stem_kernel_lite/stem_kernel.cpp:14-95) with its node and
// edge scores (score_table.cpp:14-53,56-101,162-201) for a whole list of pairs.
//
// Restated recurrence (DESIGN.md has the derivation; tests/restated.py is the executable model):
// on non-leaf nodes i of x and j of y, with c ranging over non-leaf children,
//     Q (i,j) = sum_cx  e_x(i,cx) * G0(cx,j)                                   (column-local)
//     R (i,j) = sum_cy  e_y(j,cy) * Q (i,cy)  +  el_y(j) * ql_x(i)             (row-local)
//     S (i,j) = sum_cy  e_y(j,cy) * G1(i,cy)                                   (row-local)
//     M (i,j) = in_band(i,j) ? v_s(i,j) * R(i,j) : 0                           MATCH, stem_kernel.cpp:46-59
//     G1(i,j) = M(i,j) + a_y(j) * S(i,j)                                       IY,    stem_kernel.cpp:62-67
//     G0(i,j) = G1(i,j) + a_x(i) * Q(i,j)                                      IX,    stem_kernel.cpp:70-77
//     k(x,y)  = sum_ij paths_x(i) * paths_y(j) * M(i,j)  (+ plr_x * lr_y)      result, stem_kernel.cpp:88-93
// The K tables are gone (they only count root->node paths), MATCH costs deg_y instead of
// deg_x*deg_y per cell, and only G0 needs history across rows; Q and G1 live for one row block.
//
// Mapping: rows of one x-level are independent, so they are processed RB at a time:
//   phase A  threads <-> columns j : Q rows of the block from earlier G0 rows (coalesced reads of the
//            per-CTA G0 slab in global memory / L2), Q kept in shared memory
//   phase B  y-level by y-level, threads <-> (row, node of the level) with rows fastest: R, S from
//            shared memory (conflict-free: Q has an odd row stride, G1 is stored [node][row])
//   phase C  threads <-> columns : G0 rows of the block written back coalesced
#include <cstdio>

#include "kernels.cuh"

namespace stemk {

namespace {

constexpr int kStemThreads = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// node_score(xx,yy,i,j): score_table.cpp:14-53 (simple) / :162-201 (subst); the table decides which
__device__ __forceinline__ double node_match(const SetView& X, const SetView& Y, const RecDev& rx, const RecDev& ry,
                                             uint32_t i, uint32_t j, const double* __restrict__ tab, double ax,
                                             double ay) {
  const uint32_t gi = rx.node0 + i, gj = ry.node0 + j;
  const uint8_t bx = X.bcode[gi], by = Y.bcode[gj];
  double v;
  if (bx != 0xFF && by != 0xFF) {
    v = tab[bx * 16 + by] * X.bfreq[gi] * Y.bfreq[gj];
  } else {
    v = 0.0;
    const uint32_t x0 = X.boff[rx.boff0 + i], x1 = X.boff[rx.boff0 + i + 1];
    const uint32_t y0 = Y.boff[ry.boff0 + j], y1 = Y.boff[ry.boff0 + j + 1];
    for (uint32_t p = x0; p < x1; ++p) {
      const double cx = X.bfq[p];
      const uint32_t ab = X.bab[p];
      for (uint32_t q = y0; q < y1; ++q) v += tab[ab * 16 + Y.bab[q]] * cx * Y.bfq[q];
    }
  }
  v += ay * X.gapt[gi];
  v += ax * Y.gapt[gj];
  return v;
}

__global__ void __launch_bounds__(kStemThreads) stem_pairs_kernel(const StemLaunch P) {
  extern __shared__ double smem[];
  const uint32_t RB = P.rb;
  const uint32_t RBP = RB | 1u;                 // odd stride of the [node][row] G1 tile
  const uint32_t NYP_CAP = P.ny_cap | 1u;       // odd stride of the [row][node] Q tile
  double* tab = smem;                            // 256
  double* red = tab + 256;                       // 8 warp partials
  double* Qs = red + 8;                          // RB * NYP_CAP
  double* G1s = Qs + (size_t)RB * NYP_CAP;       // ny_cap * RBP
  __shared__ unsigned long long s_pair;

  const int tid = threadIdx.x;
  for (int t = tid; t < 256; t += kStemThreads) tab[t] = P.pair_tab[t];
  double* __restrict__ G0 = P.scratch + (size_t)blockIdx.x * P.scratch_stride;
  const SetView& X = P.X;
  const SetView& Y = P.Y;

  for (;;) {
    __syncthreads();  // previous pair fully retired (also orders the tab fill on the first trip)
    if (tid == 0) s_pair = atomicAdd(P.counter, 1ull);
    __syncthreads();
    const unsigned long long k = s_pair;
    if (k >= P.n_pairs) break;
    const RecDev rx = X.rec[P.xi[k]];
    const RecDev ry = Y.rec[P.yi[k]];
    const uint32_t Nx = rx.N, Ny = ry.N;
    const double extra = rx.plr * (double)ry.lr;
    if (Nx == 0 || Ny == 0) {
      if (tid == 0) P.out[k] = extra;
      continue;
    }
    const uint32_t NYP = Ny | 1u;               // Q row stride for this pair (odd, <= NYP_CAP)
    const uint32_t NYG = (Ny + 1u) & ~1u;       // G0 row stride in the global slab (16-byte rows)
    const double* __restrict__ xa = X.a + rx.node0;
    const double* __restrict__ ya = Y.a + ry.node0;
    const uint32_t* __restrict__ xcoff = X.coff + rx.coff0;
    const uint32_t* __restrict__ ycoff = Y.coff + ry.coff0;
    const uint32_t* __restrict__ xlev = X.lev_off + rx.lev0;
    const uint32_t* __restrict__ ylev = Y.lev_off + ry.lev0;
    const uint32_t* __restrict__ xlen = X.len + rx.node0;
    const uint32_t* __restrict__ ylen = Y.len + ry.node0;
    const double* __restrict__ xql = X.ql + rx.node0;
    const double* __restrict__ yel = Y.el + ry.node0;
    const double* __restrict__ xpaths = X.paths + rx.node0;
    const double* __restrict__ ypaths = Y.paths + ry.node0;
    double acc = 0.0;

    for (uint32_t lx = 0; lx < rx.nlev; ++lx) {
      const uint32_t lev_end = xlev[lx + 1];
      for (uint32_t row0 = xlev[lx]; row0 < lev_end; row0 += RB) {
        const uint32_t nrows = min(RB, lev_end - row0);

        // ---- phase A: Q(block rows, all columns)
        for (uint32_t j = tid; j < Ny; j += kStemThreads) {
          for (uint32_t r = 0; r < nrows; ++r) {
            const uint32_t i = row0 + r;
            double q = 0.0;
            const uint32_t e1 = xcoff[i + 1];
            for (uint32_t e = xcoff[i]; e < e1; ++e) q = fma(X.ce[e], G0[(size_t)X.cidx[e] * NYG + j], q);
            Qs[r * NYP + j] = q;
          }
        }
        __syncthreads();

        // ---- phase B: sweep the y DAG level by level; cells = (row, node), rows fastest
        for (uint32_t ly = 0; ly < ry.nlev; ++ly) {
          const uint32_t j0 = ylev[ly];
          const uint32_t ncell = nrows * (ylev[ly + 1] - j0);
          for (uint32_t cell = tid; cell < ncell; cell += kStemThreads) {
            const uint32_t jj = cell / nrows;
            const uint32_t r = cell - jj * nrows;
            const uint32_t j = j0 + jj, i = row0 + r;
            const uint32_t lxi = xlen[i], lyj = ylen[j];
            const uint32_t dl = lxi > lyj ? lxi - lyj : lyj - lxi;
            const bool in_band = (P.len_band == 0u) || (dl <= P.len_band);
            double S = 0.0, R = 0.0;
            const uint32_t e1 = ycoff[j + 1];
            const double* __restrict__ qrow = Qs + r * NYP;
            for (uint32_t e = ycoff[j]; e < e1; ++e) {
              const uint32_t c = Y.cidx[e];
              const double ce = Y.ce[e];
              S = fma(ce, G1s[c * RBP + r], S);
              if (in_band) R = fma(ce, qrow[c], R);
            }
            const double ay = ya[j];
            double m = 0.0;
            if (in_band) {
              R = fma(yel[j], xql[i], R);
              m = node_match(X, Y, rx, ry, i, j, tab, xa[i], ay) * R;
              acc = fma(xpaths[i] * ypaths[j], m, acc);
            }
            G1s[j * RBP + r] = fma(ay, S, m);
          }
          __syncthreads();
        }

        // ---- phase C: G0 rows of the block, coalesced
        for (uint32_t j = tid; j < Ny; j += kStemThreads) {
          for (uint32_t r = 0; r < nrows; ++r)
            G0[(size_t)(row0 + r) * NYG + j] = fma(xa[row0 + r], Qs[r * NYP + j], G1s[j * RBP + r]);
        }
        __syncthreads();
      }
    }

    // ---- block reduction of the path-weighted MATCH sum
    acc = warp_sum(acc);
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int w = 0; w < kStemThreads / 32; ++w) s += red[w];
      P.out[k] = s + extra;
    }
  }
}

}  // namespace

size_t stem_smem_bytes(uint32_t rb, uint32_t ny_cap) {
  const size_t nyp = ny_cap | 1u, rbp = rb | 1u;
  return sizeof(double) * (256 + 8 + (size_t)rb * nyp + (size_t)ny_cap * rbp);
}

int stem_max_ctas_per_sm(size_t smem) {
  int n = 0;
  cudaFuncSetAttribute(stem_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, stem_pairs_kernel, kStemThreads, smem) != cudaSuccess) return 0;
  return n;
}

cudaError_t launch_stem(const StemLaunch& p, int grid, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(stem_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  stem_pairs_kernel<<<grid, kStemThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace stemk

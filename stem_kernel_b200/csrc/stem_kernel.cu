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
// With a length band, G1(i,j) is identically 0 when len_j + band < len_i (MATCH is out of band for
// j and for every descendant of j, which are all shorter), so those cells are skipped.
//
// Mapping.  The y record (DAG in CSR form, per-node constants) is staged in shared memory once per
// pair.  Rows of one x-level are independent and are processed in blocks of up to RB rows:
//   phase A  threads <-> columns j : Q rows of the block from earlier G0 rows -- coalesced reads of the
//            per-CTA G0 slab in global memory (L2-resident), the block's x edges staged in shared
//            memory so that the loads of one thread are independent and can be in flight together
//   phase B  y-level by y-level, threads <-> cells (row, node of the level) with rows fastest, so the
//            lanes of a warp share a node (same child list, no degree divergence) and read
//            consecutive shared-memory words of the node-major Q / G1 tiles ([node][RB|1])
//   phase C  threads <-> columns : finished G0 rows written back coalesced
// One CTA barrier per y-level and three per row block.  Shared memory is addressed through 32-bit
// byte offsets from one base so that every access is an LDS/STS with a register+immediate address.
#include <cstdio>

#include "kernels.cuh"

namespace stemk {

namespace {

constexpr int kStemThreads = 512;
constexpr int kStemWarps = kStemThreads / 32;
constexpr uint32_t kXEdgeCap = 768;  // x edges staged per row block

struct __align__(16) EdgeRec {   // one staged DAG edge
  double ce;                     // g^gaps * edge weight
  uint32_t off;                  // x edges: element offset of the child's G0 row; y edges: byte offset of the child's tile row
  uint32_t pad;
};
struct __align__(16) NodeInt {   // integer part of a staged y node
  uint32_t e0, e1;               // its edges in the staged edge array
  uint32_t len;                  // last - first
  uint32_t bcode;                // single-entry base-pair profile code, 0xFF otherwise
};

// byte offsets of the shared-memory carve-up; everything 16-byte aligned
struct StemLayout {
  uint32_t tab, red, yA, yB, yG, yI, yE, yLev, xA, xB, xG, xI, xE, xOff, Q, G1, total;
};

__host__ __device__ inline StemLayout stem_layout(uint32_t rb, uint32_t ny_cap, uint32_t ey_cap, uint32_t lev_cap) {
  StemLayout L;
  uint32_t off = 0;
  auto take = [&](uint32_t bytes) { uint32_t at = off; off += (bytes + 15u) & ~15u; return at; };
  const uint32_t rbp = rb | 1u;
  L.tab = take(8 * 256);
  L.red = take(8 * kStemWarps);
  L.yA = take(16 * ny_cap);        // double2 {a, el}
  L.yB = take(16 * ny_cap);        // double2 {paths, bfreq}
  L.yG = take(8 * ny_cap);         // gapt
  L.yI = take(16 * ny_cap);        // NodeInt
  L.yE = take(16 * ey_cap);        // EdgeRec
  L.yLev = take(4 * (lev_cap + 1));
  L.xA = take(16 * rb);            // double2 {a, ql}
  L.xB = take(16 * rb);            // double2 {paths, bfreq}
  L.xG = take(8 * rb);             // gapt
  L.xI = take(8 * rb);             // uint2 {len, bcode}
  L.xE = take(16 * kXEdgeCap);     // EdgeRec
  L.xOff = take(4 * (rb + 1));
  L.Q = take(8 * ny_cap * rbp);    // [node][rbp]
  L.G1 = take(8 * ny_cap * rbp);
  L.total = off;
  return L;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// general node_score for records whose base-pair profiles have several entries (alignments, IUPAC):
// score_table.cpp:14-41 / :162-186
__device__ __noinline__ double node_match_general(const uint32_t* __restrict__ xboff, const uint8_t* __restrict__ xbab,
                                                  const double* __restrict__ xbfq, const uint32_t* __restrict__ yboff,
                                                  const uint8_t* __restrict__ ybab, const double* __restrict__ ybfq,
                                                  const double* tab) {
  double v = 0.0;
  const uint32_t x0 = xboff[0], x1 = xboff[1], y0 = yboff[0], y1 = yboff[1];
  for (uint32_t p = x0; p < x1; ++p) {
    const double cx = xbfq[p];
    const uint32_t ab = xbab[p];
    for (uint32_t q = y0; q < y1; ++q) v += tab[ab * 16 + ybab[q]] * cx * ybfq[q];
  }
  return v;
}

__global__ void __launch_bounds__(kStemThreads, 1) stem_pairs_kernel(const StemLaunch P) {
  extern __shared__ __align__(16) unsigned char sm[];
  __shared__ unsigned long long s_pair;
  const StemLayout L = stem_layout(P.rb, P.ny_cap, P.ey_cap, P.lev_cap);
#define SM(T, byteoff) (*reinterpret_cast<T*>(sm + (byteoff)))

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t RB = P.rb, RBP = P.rb | 1u;
  const uint32_t band = P.len_band;
  for (uint32_t t = tid; t < 256; t += kStemThreads) SM(double, L.tab + 8 * t) = P.pair_tab[t];
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
    const uint32_t NYS = (Ny + 1u) & ~1u;   // row stride of the G0 slab (16-byte rows)
    const uint32_t ye0 = Y.coff[ry.coff0];  // first edge of y in the global edge arrays
    const uint32_t Ey = Y.coff[ry.coff0 + Ny] - ye0;

    // ---- stage the y record
    for (uint32_t j = tid; j < Ny; j += kStemThreads) {
      const uint32_t g = ry.node0 + j;
      SM(double2, L.yA + 16 * j) = make_double2(Y.a[g], Y.el[g]);
      SM(double2, L.yB + 16 * j) = make_double2(Y.paths[g], Y.bfreq[g]);
      SM(double, L.yG + 8 * j) = Y.gapt[g];
      NodeInt ni;
      ni.e0 = Y.coff[ry.coff0 + j] - ye0; ni.e1 = Y.coff[ry.coff0 + j + 1] - ye0; ni.len = Y.len[g]; ni.bcode = Y.bcode[g];
      SM(NodeInt, L.yI + 16 * j) = ni;
    }
    for (uint32_t e = tid; e < Ey; e += kStemThreads) {
      EdgeRec er;
      er.ce = Y.ce[ye0 + e]; er.off = Y.cidx[ye0 + e] * RBP * 8u; er.pad = 0;
      SM(EdgeRec, L.yE + 16 * e) = er;
    }
    for (uint32_t l = tid; l <= ry.nlev; l += kStemThreads) SM(uint32_t, L.yLev + 4 * l) = Y.lev_off[ry.lev0 + l];
    const uint32_t* __restrict__ xcoff = X.coff + rx.coff0;
    const uint32_t* __restrict__ xlev = X.lev_off + rx.lev0;
    const bool simple_bpf = (rx.flags & REC_SIMPLE_BPF) && (ry.flags & REC_SIMPLE_BPF);
    const bool skip_short = band != 0u && (ry.flags & REC_LEN_MONOTONE);  // zero-G1 shortcut is valid
    double acc = 0.0;

    for (uint32_t lx = 0; lx < rx.nlev; ++lx) {
      const uint32_t lev_end = xlev[lx + 1];
      uint32_t row0 = xlev[lx];
      while (row0 < lev_end) {
        // rows of this block: at most RB, and at most kXEdgeCap staged edges
        uint32_t nrows = min(RB, lev_end - row0);
        const uint32_t xe0 = xcoff[row0];
        while (nrows > 1 && xcoff[row0 + nrows] - xe0 > kXEdgeCap) --nrows;
        const uint32_t nedge_all = xcoff[row0 + nrows] - xe0;
        const bool staged = nedge_all <= kXEdgeCap;  // false only for a single row wider than the staging area

        __syncthreads();  // previous block is done with Q/G1 and the x staging area
        for (uint32_t r = tid; r < nrows; r += kStemThreads) {
          const uint32_t g = rx.node0 + row0 + r;
          SM(double2, L.xA + 16 * r) = make_double2(X.a[g], X.ql[g]);
          SM(double2, L.xB + 16 * r) = make_double2(X.paths[g], X.bfreq[g]);
          SM(double, L.xG + 8 * r) = X.gapt[g];
          SM(uint2, L.xI + 8 * r) = make_uint2(X.len[g], X.bcode[g]);
        }
        for (uint32_t r = tid; r <= nrows; r += kStemThreads) SM(uint32_t, L.xOff + 4 * r) = xcoff[row0 + r] - xe0;
        if (staged)
          for (uint32_t e = tid; e < nedge_all; e += kStemThreads) {
            EdgeRec er;
            er.ce = X.ce[xe0 + e]; er.off = X.cidx[xe0 + e] * NYS; er.pad = 0;
            SM(EdgeRec, L.xE + 16 * e) = er;
          }
        __syncthreads();

        // ---- phase A: Q(block rows, all columns), one column per thread, loads of a row unrolled
        if (lx == 0) {
          for (uint32_t t = tid; t < Ny * RBP; t += kStemThreads) SM(double, L.Q + 8 * t) = 0.0;  // hairpin rows: no inner pair
        } else if (staged) {
          for (uint32_t j = tid; j < Ny; j += kStemThreads) {
            const double* __restrict__ g0c = G0 + j;
            uint32_t e = 0;
            for (uint32_t r = 0; r < nrows; ++r) {
              const uint32_t e1 = SM(uint32_t, L.xOff + 4 * (r + 1));
              double q0 = 0.0, q1 = 0.0;
#pragma unroll 1
              for (; e + 3 < e1; e += 4) {
                const EdgeRec a = SM(EdgeRec, L.xE + 16 * e), b = SM(EdgeRec, L.xE + 16 * e + 16),
                              c = SM(EdgeRec, L.xE + 16 * e + 32), d = SM(EdgeRec, L.xE + 16 * e + 48);
                const double va = g0c[a.off], vb = g0c[b.off], vc = g0c[c.off], vd = g0c[d.off];
                q0 = fma(a.ce, va, q0); q1 = fma(b.ce, vb, q1); q0 = fma(c.ce, vc, q0); q1 = fma(d.ce, vd, q1);
              }
#pragma unroll 1
              for (; e < e1; ++e) {
                const EdgeRec a = SM(EdgeRec, L.xE + 16 * e);
                q0 = fma(a.ce, g0c[a.off], q0);
              }
              SM(double, L.Q + 8 * (j * RBP + r)) = q0 + q1;
            }
          }
        } else {
          for (uint32_t j = tid; j < Ny; j += kStemThreads) {
            double q = 0.0;
            for (uint32_t e = xe0; e < xcoff[row0 + 1]; ++e) q = fma(X.ce[e], G0[(size_t)X.cidx[e] * NYS + j], q);
            SM(double, L.Q + 8 * (j * RBP)) = q;
          }
        }
        __syncthreads();

        // ---- phase B: sweep the y DAG level by level; cells = (row, node), rows fastest.
        // (r, jj) of this thread's first cell and the step to its next cell (tid + kStemThreads)
        const uint32_t jj_first = tid / nrows, r_first = tid - jj_first * nrows;
        const uint32_t jj_step = kStemThreads / nrows, r_step = kStemThreads - jj_step * nrows;
        for (uint32_t ly = 0; ly < ry.nlev; ++ly) {
          const uint32_t j0 = SM(uint32_t, L.yLev + 4 * ly);
          const uint32_t nj = SM(uint32_t, L.yLev + 4 * ly + 4) - j0;
          uint32_t jj = jj_first, r = r_first;
          while (jj < nj) {
            const uint32_t j = j0 + jj;
            const NodeInt ni = SM(NodeInt, L.yI + 16 * j);
            const uint2 xi = SM(uint2, L.xI + 8 * r);
            const uint32_t cell_off = 8u * (j * RBP + r);
            // G1 == 0 for every node shorter than len_i - band (see header); nothing to do for the cell
            if (skip_short && ni.len + band < xi.x) {
              SM(double, L.G1 + cell_off) = 0.0;
            } else {
              const uint32_t dl = xi.x > ni.len ? xi.x - ni.len : ni.len - xi.x;
              const bool in_band = (band == 0u) || (dl <= band);
              const double2 yA = SM(double2, L.yA + 16 * j);                   // {a_y, el_y}
              double S0 = 0.0, S1 = 0.0;
              double m = 0.0;
              const uint32_t g1r = L.G1 + 8u * r, qr = L.Q + 8u * r;
              uint32_t e = L.yE + 16u * ni.e0;
              const uint32_t eend = L.yE + 16u * ni.e1;
              if (in_band) {
                const double2 xA = SM(double2, L.xA + 16 * r);                 // {a_x, ql_x}
                const double2 xB = SM(double2, L.xB + 16 * r);                 // {paths_x, bfreq_x}
                const double2 yB = SM(double2, L.yB + 16 * j);                 // {paths_y, bfreq_y}
                double vs;
                if (simple_bpf) vs = SM(double, L.tab + 8u * (xi.y * 16u + ni.bcode)) * xB.y * yB.y;
                else vs = node_match_general(X.boff + rx.boff0 + row0 + r, X.bab, X.bfq, Y.boff + ry.boff0 + j, Y.bab, Y.bfq,
                                             reinterpret_cast<const double*>(sm + L.tab));
                vs = fma(yA.x, SM(double, L.xG + 8 * r), vs);
                vs = fma(xA.x, SM(double, L.yG + 8 * j), vs);
                double R0 = 0.0, R1 = 0.0;
#pragma unroll 1
                for (; e + 16u < eend; e += 32u) {
                  const EdgeRec a = SM(EdgeRec, e), b = SM(EdgeRec, e + 16u);
                  S0 = fma(a.ce, SM(double, g1r + a.off), S0); R0 = fma(a.ce, SM(double, qr + a.off), R0);
                  S1 = fma(b.ce, SM(double, g1r + b.off), S1); R1 = fma(b.ce, SM(double, qr + b.off), R1);
                }
                if (e < eend) {
                  const EdgeRec a = SM(EdgeRec, e);
                  S0 = fma(a.ce, SM(double, g1r + a.off), S0); R0 = fma(a.ce, SM(double, qr + a.off), R0);
                }
                const double R = fma(yA.y, xA.y, R0 + R1);
                m = vs * R;
                acc = fma(xB.x * yB.x, m, acc);
              } else {
#pragma unroll 1
                for (; e + 16u < eend; e += 32u) {
                  const EdgeRec a = SM(EdgeRec, e), b = SM(EdgeRec, e + 16u);
                  S0 = fma(a.ce, SM(double, g1r + a.off), S0);
                  S1 = fma(b.ce, SM(double, g1r + b.off), S1);
                }
                if (e < eend) {
                  const EdgeRec a = SM(EdgeRec, e);
                  S0 = fma(a.ce, SM(double, g1r + a.off), S0);
                }
              }
              SM(double, L.G1 + cell_off) = fma(yA.x, S0 + S1, m);
            }
            jj += jj_step; r += r_step;
            if (r >= nrows) { r -= nrows; ++jj; }
          }
          __syncthreads();
        }

        // ---- phase C: finished rows, G0(i,:) = G1 + a_x * Q, coalesced over columns
        for (uint32_t j = tid; j < Ny; j += kStemThreads) {
          for (uint32_t r = 0; r < nrows; ++r) {
            const uint32_t o = 8u * (j * RBP + r);
            G0[(size_t)(row0 + r) * NYS + j] = fma(SM(double2, L.xA + 16 * r).x, SM(double, L.Q + o), SM(double, L.G1 + o));
          }
        }
        row0 += nrows;
      }
    }

    // ---- block reduction of the path-weighted MATCH sum
    acc = warp_sum(acc);
    __syncthreads();
    if (lane == 0) SM(double, L.red + 8 * warp) = acc;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < kStemWarps; ++w) t += SM(double, L.red + 8 * w);
      P.out[k] = t + extra;
    }
  }
#undef SM
}

}  // namespace

size_t stem_smem_bytes(uint32_t rb, uint32_t ny_cap, uint32_t ey_cap, uint32_t lev_cap) {
  return stem_layout(rb, ny_cap, ey_cap, lev_cap).total;
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

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
// Mapping: dataflow over rows.  The y record (DAG in CSR form, per-node constants) is staged in shared
// memory once per pair.  Each of the CTA's warps then repeatedly takes the next row i of x (rows are
// numbered level by level, so every child of a row has a smaller number), waits on shared-memory flags
// until the rows of i's inner pairs are finished, and processes the row on its own:
//   A  lanes <-> columns: Q(i,:) from the finished G0 rows (coalesced L2 reads of the per-CTA G0 slab;
//      the row's edge list sits in lane registers and is broadcast by shuffles)
//   B  y-level by y-level, lanes <-> nodes of the level: R, S gathered from the warp's private Q / G1
//      rows in shared memory; only __syncwarp between levels
//   C  lanes <-> columns: the finished G0 row is written back, fenced, and the row's flag is raised
// There is no CTA-wide barrier inside a pair: warps sit at different rows, levels and phases, which is
// what hides the shared-memory and L2 latency of this pointer-chasing recursion.  Shared memory is
// addressed through 32-bit byte offsets from one base (LDS/STS with register+immediate addresses).
#include <cstdio>

#include "kernels.cuh"

namespace stemk {

namespace {

constexpr int kStemThreads = 512;
constexpr int kStemWarps = kStemThreads / 32;

struct __align__(16) EdgeRec {   // one staged y edge
  double ce;                     // g^gaps * edge weight
  uint32_t off;                  // byte offset of the child inside a Q / G1 row
  uint32_t pad;
};
struct __align__(16) NodeInt {   // integer part of a staged y node
  uint32_t e0, e1;               // its edges in the staged edge array
  uint32_t len;                  // last - first
  uint32_t bcode;                // single-entry base-pair profile code, 0xFF otherwise
};

// byte offsets of the shared-memory carve-up; everything 16-byte aligned
struct StemLayout {
  uint32_t tab, red, yA, yB, yG, yI, yE, yLev, done, rows, row_bytes, nslots, total;
};

// nslots warps get a private (Q row, G1 row) pair; the rest of the CTA's warps stay idle for that launch
__host__ __device__ inline StemLayout stem_layout(uint32_t nslots, uint32_t nx_cap, uint32_t ny_cap, uint32_t ey_cap,
                                                  uint32_t lev_cap) {
  StemLayout L;
  uint32_t off = 0;
  auto take = [&](uint32_t bytes) { uint32_t at = off; off += (bytes + 15u) & ~15u; return at; };
  L.tab = take(8 * 256);
  L.red = take(8 * kStemWarps);
  L.yA = take(16 * ny_cap);        // double2 {a, el}
  L.yB = take(16 * ny_cap);        // double2 {paths, bfreq}
  L.yG = take(8 * ny_cap);         // gapt
  L.yI = take(16 * ny_cap);        // NodeInt
  L.yE = take(16 * ey_cap);        // EdgeRec
  L.yLev = take(4 * (lev_cap + 1));
  L.done = take(4 * nx_cap);       // row-finished flags
  L.row_bytes = (8u * ny_cap + 15u) & ~15u;
  L.nslots = nslots;
  L.rows = take(2u * L.row_bytes * nslots);  // per slot: Q row then G1 row
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

__device__ __forceinline__ uint32_t ld_flag(const unsigned char* sm, uint32_t byteoff) {
  return *reinterpret_cast<const volatile uint32_t*>(sm + byteoff);
}

__global__ void __launch_bounds__(kStemThreads, 1) stem_pairs_kernel(const StemLaunch P) {
  extern __shared__ __align__(16) unsigned char sm[];
  __shared__ unsigned long long s_pair;
  __shared__ uint32_t s_next_row;
  const StemLayout L = stem_layout(P.nslots, P.nx_cap, P.ny_cap, P.ey_cap, P.lev_cap);
#define SM(T, byteoff) (*reinterpret_cast<T*>(sm + (byteoff)))

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t band = P.len_band;
  const unsigned long long n_items = P.order ? *P.n_items_dev : P.n_pairs;
  for (uint32_t t = tid; t < 256; t += kStemThreads) SM(double, L.tab + 8 * t) = P.pair_tab[t];
  double* __restrict__ G0 = P.scratch + (size_t)blockIdx.x * P.scratch_stride;
  const SetView& X = P.X;
  const SetView& Y = P.Y;
  const uint32_t qrow = L.rows + 2u * L.row_bytes * warp;  // this warp's Q row; its G1 row follows
  const uint32_t g1row = qrow + L.row_bytes;

  for (;;) {
    __syncthreads();  // previous pair fully retired (also orders the tab fill on the first trip)
    if (tid == 0) { s_pair = atomicAdd(P.counter, 1ull); s_next_row = 0; }
    __syncthreads();
    if (s_pair >= n_items) break;
    const unsigned long long k = P.order ? (unsigned long long)P.order[s_pair] : s_pair;
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

    // ---- stage the y record, clear the row flags
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
      er.ce = Y.ce[ye0 + e]; er.off = Y.cidx[ye0 + e] * 8u; er.pad = 0;
      SM(EdgeRec, L.yE + 16 * e) = er;
    }
    for (uint32_t l = tid; l <= ry.nlev; l += kStemThreads) SM(uint32_t, L.yLev + 4 * l) = Y.lev_off[ry.lev0 + l];
    for (uint32_t i = tid; i < Nx; i += kStemThreads) SM(uint32_t, L.done + 4 * i) = 0u;
    __syncthreads();

    const uint32_t* __restrict__ xcoff = X.coff + rx.coff0;
    const bool simple_bpf = (rx.flags & REC_SIMPLE_BPF) && (ry.flags & REC_SIMPLE_BPF);
    const bool skip_short = band != 0u && (ry.flags & REC_LEN_MONOTONE);  // zero-G1 shortcut is valid
    double acc = 0.0;

    if (warp < L.nslots) {
      for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(&s_next_row, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= Nx) break;
        const uint32_t gx = rx.node0 + i;
        const uint32_t e0 = xcoff[i], e1 = xcoff[i + 1];
        const double xa = X.a[gx], xql = X.ql[gx], xpath = X.paths[gx], xbf = X.bfreq[gx], xgap = X.gapt[gx];
        const uint32_t xl = X.len[gx], xbc = X.bcode[gx];

        // ---- phase A: Q(i,:) = sum over inner pairs c of e * G0(c,:)
        for (uint32_t eb = e0; eb < e1 || eb == e0; eb += 32u) {
          // up to 32 edges of the row live in lane registers
          const uint32_t ne = min(32u, e1 - eb);
          double ce_l = 0.0;
          uint32_t off_l = 0u;
          if (lane < ne) {
            const uint32_t c = X.cidx[eb + lane];
            ce_l = X.ce[eb + lane];
            off_l = c * NYS;
            while (ld_flag(sm, L.done + 4u * c) == 0u) __nanosleep(40);  // wait until that row is finished
          }
          __syncwarp();
          __threadfence_block();  // acquire: the G0 rows behind the flags just seen
          for (uint32_t jb = 0; jb < Ny; jb += 64u) {  // uniform trip count: the shuffles below need every lane
            const uint32_t j = jb + lane;
            const bool one = j < Ny, two = j + 32u < Ny;
            double q0 = (one && eb != e0) ? SM(double, qrow + 8u * j) : 0.0;
            double q1 = (two && eb != e0) ? SM(double, qrow + 8u * (j + 32u)) : 0.0;
#pragma unroll 4
            for (uint32_t t = 0; t < ne; ++t) {
              const double ce = __shfl_sync(0xffffffffu, ce_l, t);
              const uint32_t off = __shfl_sync(0xffffffffu, off_l, t);
              if (one) q0 = fma(ce, __ldcg(G0 + off + j), q0);
              if (two) q1 = fma(ce, __ldcg(G0 + off + j + 32u), q1);
            }
            if (one) SM(double, qrow + 8u * j) = q0;
            if (two) SM(double, qrow + 8u * (j + 32u)) = q1;
          }
          if (e1 == e0) break;
        }
        __syncwarp();

        // ---- phase B: sweep the y DAG level by level, lanes <-> nodes of the level
        double racc = 0.0;
        uint32_t jbeg = SM(uint32_t, L.yLev);
        for (uint32_t ly = 0; ly < ry.nlev; ++ly) {
          const uint32_t jend = SM(uint32_t, L.yLev + 4u * ly + 4u);
          for (uint32_t j = jbeg + lane; j < jend; j += 32u) {
            const NodeInt ni = SM(NodeInt, L.yI + 16u * j);
            if (skip_short && ni.len + band < xl) {  // G1 == 0 here and below (see header)
              SM(double, g1row + 8u * j) = 0.0;
              continue;
            }
            const uint32_t dl = xl > ni.len ? xl - ni.len : ni.len - xl;
            const bool in_band = (band == 0u) || (dl <= band);
            const double2 yA = SM(double2, L.yA + 16u * j);  // {a_y, el_y}
            double S0 = 0.0, S1 = 0.0, m = 0.0;
            uint32_t e = L.yE + 16u * ni.e0;
            const uint32_t eend = L.yE + 16u * ni.e1;
            if (in_band) {
              const double2 yB = SM(double2, L.yB + 16u * j);  // {paths_y, bfreq_y}
              double vs;
              if (simple_bpf) vs = SM(double, L.tab + 8u * (xbc * 16u + ni.bcode)) * xbf * yB.y;
              else vs = node_match_general(X.boff + rx.boff0 + i, X.bab, X.bfq, Y.boff + ry.boff0 + j, Y.bab, Y.bfq,
                                           reinterpret_cast<const double*>(sm + L.tab));
              vs = fma(yA.x, xgap, vs);
              vs = fma(xa, SM(double, L.yG + 8u * j), vs);
              double R0 = 0.0, R1 = 0.0;
#pragma unroll 1
              for (; e + 16u < eend; e += 32u) {
                const EdgeRec a = SM(EdgeRec, e), b = SM(EdgeRec, e + 16u);
                S0 = fma(a.ce, SM(double, g1row + a.off), S0); R0 = fma(a.ce, SM(double, qrow + a.off), R0);
                S1 = fma(b.ce, SM(double, g1row + b.off), S1); R1 = fma(b.ce, SM(double, qrow + b.off), R1);
              }
              if (e < eend) {
                const EdgeRec a = SM(EdgeRec, e);
                S0 = fma(a.ce, SM(double, g1row + a.off), S0); R0 = fma(a.ce, SM(double, qrow + a.off), R0);
              }
              m = vs * fma(yA.y, xql, R0 + R1);
              racc = fma(yB.x, m, racc);
            } else {
#pragma unroll 1
              for (; e + 16u < eend; e += 32u) {
                const EdgeRec a = SM(EdgeRec, e), b = SM(EdgeRec, e + 16u);
                S0 = fma(a.ce, SM(double, g1row + a.off), S0);
                S1 = fma(b.ce, SM(double, g1row + b.off), S1);
              }
              if (e < eend) {
                const EdgeRec a = SM(EdgeRec, e);
                S0 = fma(a.ce, SM(double, g1row + a.off), S0);
              }
            }
            SM(double, g1row + 8u * j) = fma(yA.x, S0 + S1, m);
          }
          jbeg = jend;
          __syncwarp();
        }
        acc = fma(xpath, racc, acc);

        // ---- phase C: finished row G0(i,:) = G1 + a_x * Q, then publish it
        double* __restrict__ g0row = G0 + (size_t)i * NYS;
        for (uint32_t j = lane; j < Ny; j += 32u) g0row[j] = fma(xa, SM(double, qrow + 8u * j), SM(double, g1row + 8u * j));
        __threadfence_block();
        __syncwarp();
        if (lane == 0) *reinterpret_cast<volatile uint32_t*>(sm + L.done + 4u * i) = 1u;
      }
    }

    // ---- block reduction of the path-weighted MATCH sum
    acc = warp_sum(acc);
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

// ---------------------------------------------------------------------------------------------------------------
// The same recurrence with NOTHING staged: records of any size (the reference's operator() has no size limit).
// The y record is read where the upload put it (global memory through L1/L2), every warp's Q and G1 rows and the row
// flags live in the CTA's global scratch next to the G0 slab.  Same dataflow -- a warp takes the next row, waits for
// the rows of its inner pairs, runs phases A, B, C alone -- and the same arithmetic per cell as stem_pairs_kernel
// (the two y-edge sums are kept in one accumulator instead of two).  Only the pairs the classifier finds too large
// for the staged kernel's shared-memory carve-up run here (bucket StemClassify::big_bucket).
struct UnstagedLayout { unsigned long long pitch, rows, flags, total; };
__host__ __device__ inline UnstagedLayout unstaged_layout(uint32_t nx_cap, uint32_t ny_cap) {
  UnstagedLayout U;
  U.pitch = ((unsigned long long)ny_cap + 1ull) & ~1ull;
  U.rows = (unsigned long long)nx_cap * U.pitch;
  U.flags = U.rows + 2ull * kStemWarps * U.pitch;
  U.total = U.flags + ((unsigned long long)nx_cap + 1ull) / 2ull;
  U.total = (U.total + 1ull) & ~1ull;
  return U;
}

__global__ void __launch_bounds__(kStemThreads, 1) stem_pairs_unstaged_kernel(const StemBigLaunch P) {
  __shared__ double s_tab[256];
  __shared__ double s_red[kStemWarps];
  __shared__ unsigned long long s_pair;
  __shared__ uint32_t s_next_row;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t band = P.len_band;
  const unsigned long long n_items = P.count[P.bucket];
  const uint32_t* __restrict__ order = P.order + P.start[P.bucket];
  for (uint32_t t = tid; t < 256; t += kStemThreads) s_tab[t] = P.pair_tab[t];
  const UnstagedLayout U = unstaged_layout(P.nx_cap, P.ny_cap);
  double* __restrict__ base = P.scratch + (size_t)blockIdx.x * P.scratch_stride;
  double* __restrict__ G0 = base;
  double* qrow = base + U.rows + 2ull * warp * U.pitch;   // this warp's Q row; its G1 row follows
  double* g1row = qrow + U.pitch;
  volatile uint32_t* done = reinterpret_cast<volatile uint32_t*>(base + U.flags);
  const SetView& X = P.X;
  const SetView& Y = P.Y;

  for (;;) {
    __syncthreads();
    if (tid == 0) { s_pair = atomicAdd(P.counter, 1ull); s_next_row = 0; }
    __syncthreads();
    if (s_pair >= n_items) break;
    const unsigned long long k = order[s_pair];
    const RecDev rx = X.rec[P.xi[k]];
    const RecDev ry = Y.rec[P.yi[k]];
    const uint32_t Nx = rx.N, Ny = ry.N;
    const double extra = rx.plr * (double)ry.lr;
    if (Nx == 0 || Ny == 0) {
      if (tid == 0) P.out[k] = extra;
      continue;
    }
    const unsigned long long NYS = ((unsigned long long)Ny + 1ull) & ~1ull;
    for (uint32_t i = tid; i < Nx; i += kStemThreads) done[i] = 0u;
    __threadfence_block();
    __syncthreads();

    const uint32_t* __restrict__ xcoff = X.coff + rx.coff0;
    const uint32_t* __restrict__ ycoff = Y.coff + ry.coff0;
    const uint32_t* __restrict__ ylev = Y.lev_off + ry.lev0;
    const bool simple_bpf = (rx.flags & REC_SIMPLE_BPF) && (ry.flags & REC_SIMPLE_BPF);
    const bool skip_short = band != 0u && (ry.flags & REC_LEN_MONOTONE);
    double acc = 0.0;

    for (;;) {
      uint32_t i = 0;
      if (lane == 0) i = atomicAdd(&s_next_row, 1u);
      i = __shfl_sync(0xffffffffu, i, 0);
      if (i >= Nx) break;
      const uint32_t gx = rx.node0 + i;
      const uint32_t e0 = xcoff[i], e1 = xcoff[i + 1];
      const double xa = X.a[gx], xql = X.ql[gx], xpath = X.paths[gx], xbf = X.bfreq[gx], xgap = X.gapt[gx];
      const uint32_t xl = X.len[gx], xbc = X.bcode[gx];

      // ---- phase A: Q(i,:) = sum over inner pairs c of e * G0(c,:)
      for (uint32_t eb = e0; eb < e1 || eb == e0; eb += 32u) {
        const uint32_t ne = min(32u, e1 - eb);
        double ce_l = 0.0;
        unsigned long long off_l = 0ull;
        if (lane < ne) {
          const uint32_t c = X.cidx[eb + lane];
          ce_l = X.ce[eb + lane];
          off_l = (unsigned long long)c * NYS;
          while (done[c] == 0u) __nanosleep(100);   // wait until that row is finished
        }
        __syncwarp();
        __threadfence_block();   // acquire: the G0 rows behind the flags just seen
        for (uint32_t jb = 0; jb < Ny; jb += 32u) {   // uniform trip count: the shuffles below need every lane
          const uint32_t j = jb + lane;
          const bool one = j < Ny;
          double q = (one && eb != e0) ? __ldcg(qrow + j) : 0.0;
          for (uint32_t t = 0; t < ne; ++t) {
            const double ce = __shfl_sync(0xffffffffu, ce_l, t);
            const unsigned long long off = __shfl_sync(0xffffffffu, off_l, t);
            if (one) q = fma(ce, __ldcg(G0 + off + j), q);
          }
          if (one) qrow[j] = q;
        }
        if (e1 == e0) break;
      }
      __syncwarp();

      // ---- phase B: sweep the y DAG level by level, lanes <-> nodes of the level
      double racc = 0.0;
      uint32_t jbeg = ylev[0];
      for (uint32_t ly = 0; ly < ry.nlev; ++ly) {
        const uint32_t jend = ylev[ly + 1];
        for (uint32_t j = jbeg + lane; j < jend; j += 32u) {
          const uint32_t gy = ry.node0 + j;
          const uint32_t yl = Y.len[gy];
          if (skip_short && yl + band < xl) {   // G1 == 0 here and below (see the header)
            g1row[j] = 0.0;
            continue;
          }
          const uint32_t dl = xl > yl ? xl - yl : yl - xl;
          const bool in_band = (band == 0u) || (dl <= band);
          const double ya = Y.a[gy];
          uint32_t e = ycoff[j];
          const uint32_t eend = ycoff[j + 1];
          double S = 0.0, m = 0.0;
          if (in_band) {
            double vs;
            if (simple_bpf) vs = s_tab[xbc * 16u + Y.bcode[gy]] * xbf * Y.bfreq[gy];
            else vs = node_match_general(X.boff + rx.boff0 + i, X.bab, X.bfq, Y.boff + ry.boff0 + j, Y.bab, Y.bfq, s_tab);
            vs = fma(ya, xgap, vs);
            vs = fma(xa, Y.gapt[gy], vs);
            double R = 0.0;
            for (; e < eend; ++e) {
              const uint32_t c = Y.cidx[e];
              const double ce = Y.ce[e];
              S = fma(ce, __ldcg(g1row + c), S);
              R = fma(ce, __ldcg(qrow + c), R);
            }
            m = vs * fma(Y.el[gy], xql, R);
            racc = fma(Y.paths[gy], m, racc);
          } else {
            for (; e < eend; ++e) S = fma(Y.ce[e], __ldcg(g1row + Y.cidx[e]), S);
          }
          g1row[j] = fma(ya, S, m);
        }
        jbeg = jend;
        __syncwarp();
      }
      acc = fma(xpath, racc, acc);

      // ---- phase C: finished row G0(i,:) = G1 + a_x * Q, then publish it
      double* __restrict__ g0row = G0 + (unsigned long long)i * NYS;
      for (uint32_t j = lane; j < Ny; j += 32u) g0row[j] = fma(xa, __ldcg(qrow + j), __ldcg(g1row + j));
      __threadfence_block();
      __syncwarp();
      if (lane == 0) done[i] = 1u;
    }

    acc = warp_sum(acc);
    if (lane == 0) s_red[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < kStemWarps; ++w) t += s_red[w];
      P.out[k] = t + extra;
    }
  }
}

}  // namespace

unsigned long long stem_unstaged_scratch_doubles(uint32_t nx_cap, uint32_t ny_cap) { return unstaged_layout(nx_cap, ny_cap).total; }

cudaError_t launch_stem_unstaged(const StemBigLaunch& p, int grid, cudaStream_t stream) {
  stem_pairs_unstaged_kernel<<<grid, kStemThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

size_t stem_smem_bytes(uint32_t nslots, uint32_t nx_cap, uint32_t ny_cap, uint32_t ey_cap, uint32_t lev_cap) {
  return stem_layout(nslots, nx_cap, ny_cap, ey_cap, lev_cap).total;
}
int stem_warps_per_cta() { return kStemWarps; }

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

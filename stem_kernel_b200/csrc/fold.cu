// stem_kernel_b200/csrc/fold.cu -- base-pair probabilities on the device.
//
// Replaces the per-sequence ViennaRNA call of the reference's front end (fold / init_pf_fold / pf_fold and the copy
// bp(i,j) = pr[iindx[i]-j] under a process-wide mutex, common/bpmatrix.cpp:141-177) by a batched McCaskill partition
// function for the loop model of include/stemk.h (stemk_fold_model).  Parity with ViennaRNA is unpinned (the package
// is absent from this image and from /root/reference); the checker is oracle/stemk_fold_oracle.c.
//
// Recursions (weights carry the per-nucleotide scaling s^-1 of the loop that owns the nucleotide, so the tables hold
// Z / s^length and the recursions themselves are scale-free):
//   inside, by increasing span d = j - i
//     Qb (i,j) = HP(i,j) + sum_{k,l} Qb(k,l) IL(i,j,k,l) + MLclose(i,j) sum_{k} Qm(i+1,k-1) Qm1(k,j-1)
//     Qm1(i,j) = Qm1(i,j-1) u + Qb(i,j) MLstem(i,j)          one stem starting at i, unpaired bases after it
//     Qq (i,j) = Qq (i,j-1) / s + Qb(i,j) EXTstem(i,j)       the same in the exterior loop
//     Qm (i,j) = sum_{k=i..j} (u^(k-i) + Qm(i,k-1)) Qm1(k,j)
//     Q  (i,j) = s^-(d+1) + sum_{k=i..j} Q(i,k-1) Qq(k,j)
//   outside, by decreasing span
//     Ob (i,j) = Q(1,i-1) Q(j+1,n) EXTstem(i,j) + sum_{p,q} Ob(p,q) IL(p,q,i,j)
//                + MLstem(i,j) sum_{p<i} [ Qm(p+1,i-1) A(p,j) + u^(i-p-1) B(p,j) ]
//     W  (p,q) = Ob(p,q) MLclose(p,q)
//     A  (p,j) = sum_{q>j} W(p,q) (u^(q-j-1) + Qm(j+1,q-1)),   B(p,j) = sum_{q>j} W(p,q) Qm(j+1,q-1)
//   P(i,j) = Qb(i,j) Ob(i,j) / Q(1,n)
// (the multiloop term of Ob: the other stems of the loop closed by (p,q) sit left of i, right of j or on both sides,
// at least one of them; A and B are per-(p,j) sums that make the whole outside pass O(n^3) like the inside one).
//
// Mapping: one CTA per sequence (persistent, sequences drawn longest first from a counter), a warp per cell of the
// running diagonal, lanes over the summation index; every table that is summed along its second index by one
// recursion and along its first by another is kept in both orientations (Qm / QmT, Qm1T, QqT, AT, BT) so that all the
// O(n) sums read consecutive doubles.  The tables of the sequence in flight live in the CTA's global scratch
// (10 x (n+2)^2 doubles: 7 MB at 300 nt, L2-resident); the Boltzmann factors of the model are computed on the host with
// libm's exp and staged in shared memory.  One __syncthreads per diagonal.  The pairs at or above the cut-off are
// compacted per sequence in ascending (i,j) with a block-wide scan into one output range reserved by a single atomic.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>

#include "kernels.cuh"


#ifdef FOLD_PROF
#define FP_T(v) const long long v = clock64()
#define FP_ADD(slot, a, b) do { if (lane == 0) fprof[slot] += (b) - (a); } while (0)
#else
#define FP_T(v) do {} while (0)
#define FP_ADD(slot, a, b) do {} while (0)
#endif

namespace stemk {

namespace {

constexpr int kTurn = 3, kMaxLoop = 30;
constexpr int kTables = 10;

struct FoldTab {          // Boltzmann factors of the model; *S = with the scaling of the loop's own nucleotides
  double stackS[8][8], stackB[8][8];
  double mmH[8][5][5], mmI[8][5][5];
  double d5[8][5], d3[8][5];
  double bulgeS[31], interiorS[31], ninioB[31];
  double mlintern[8], tau[8];
  double mlclosingS;      // exp(-ml_closing/kT) / s^2
  double u1, s1;          // exp(-ml_base/kT)/s,  1/s
  double kT, log_s;
};

struct FoldLaunch {
  const FoldTab* tab;
  const double* hpS;          // [cap + 1] hairpin initiation by size, with s^-(u+2)
  const double* sp;           // [cap + 3] s^-k
  const double* up;           // [cap + 3] (u/s)^k
  const uint8_t* codes;       // concatenated base codes
  const uint64_t* seq_off;    // [n_seqs + 1]
  const uint32_t* order;      // sequences, longest first
  uint32_t n_seqs, cap;       // cap = longest sequence + 2
  uint32_t pb_shared;         // byte offset of the pairable-column bits in dynamic shared memory, 0: they live in the scratch
  int no_gu;
  double cutoff;
  double* scratch;            // per CTA: kTables tables of cap^2 doubles, then cap rows of pairable-column bits
  unsigned long long scratch_stride;   // doubles per CTA
  unsigned long long* counter;         // [0] next sequence, [1] output cursor
  uint64_t* out_start;        // [n_seqs] first entry of the sequence's pairs in out_*
  uint32_t* out_count;        // [n_seqs]
  uint32_t* out_i;
  uint32_t* out_j;
  double* out_p;
  double* unpaired;           // concatenated like codes
  double* ensemble;           // [n_seqs]
  int* status;                // [n_seqs] 1: partition function out of range
  unsigned long long* prof;   // FOLD_PROF builds: per-phase cycle counters summed over warps
  double* dense;              // optional
  const uint64_t* dense_off;  // [n_seqs]
};

__device__ __forceinline__ int d_pair_type(int a, int b, int no_gu) {
  if (a == 2 && b == 3) return 1;
  if (a == 3 && b == 2) return 2;
  if (a == 3 && b == 4) return no_gu ? 0 : 3;
  if (a == 4 && b == 3) return no_gu ? 0 : 4;
  if (a == 1 && b == 4) return 5;
  if (a == 4 && b == 1) return 6;
  return 0;
}
__device__ __forceinline__ int d_rtype(int t) { return t == 0 ? 0 : ((t - 1) ^ 1) + 1; }

__device__ __forceinline__ double warp_sum_f(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// a pair of type `type` closes the loop, a pair of type tkl is the inner pair; mm_close / mm_inner = the interior-loop mismatch
// factors of the two pairs, mmI[type][S[i+1]][S[j-1]] and mmI[rtype(tkl)][S[l+1]][S[k-1]]
__device__ __forceinline__ double il_weight(const FoldTab& T, int u1, int u2, int type, int tkl, double mm_close, double mm_inner) {
  const int t2 = d_rtype(tkl);
  if (u1 != 0 && u2 != 0) {
    const int d = u1 > u2 ? u1 - u2 : u2 - u1;
    return T.interiorS[u1 + u2] * T.ninioB[d] * (mm_close * mm_inner);
  }
  const int u = u1 + u2;
  if (u == 0) return T.stackS[type][t2];
  return T.bulgeS[u] * (u == 1 ? T.stackB[type][t2] : T.tau[type] * T.tau[t2]);
}

// Sum over the interior loops (stacks and bulges included) of one cell.
// OUTSIDE = false: (i,j) closes, tab = Qb of the inner pair (k,l) = (i+1+u1, j-1-u2);
// OUTSIDE = true:  (i,j) is the inner pair, tab = Ob of the closing pair (p,q) = (i-1-u1, j+1+u2).
// Only ~1/4 of the 496 (u1, u2) combinations are pairs inside the sequence, and the owning warp's instruction stream is
// what bounds the kernel, so the combinations that exist are found WITHOUT touching the others: the
// sequence's pairable columns are a bit vector per row (pb, built once per sequence; bit l + 32 of row a = (a,l) can
// pair), lane <-> row u1 cuts its 31-column window out of it with one funnel shift, a warp scan of the popcounts gives
// every row its place in the warp's staging area, each lane writes the codes of its set bits there, and the table
// entries are then fetched and weighted with every lane busy -- consecutive lanes mostly read consecutive entries of a
// row.
template <bool OUTSIDE>
__device__ __forceinline__ double interior_sum(const FoldTab& T, const uint8_t* S, const uint8_t* s_pt, const uint32_t* pb, int pbw,
                                               uint16_t* stage, const double* __restrict__ tab, int W, int n, int i, int j, int type,
                                               int lane) {
  double acc = 0.0;
  const double mm_ij = T.mmI[OUTSIDE ? d_rtype(type) : type][S[OUTSIDE ? j + 1 : i + 1]][S[OUTSIDE ? i - 1 : j - 1]];   // this cell's own mismatch factor
  const int u1_end = min(kMaxLoop, OUTSIDE ? i - 2 : j - i - 3 - kTurn);   // last u1 with a row inside the sequence / a pair that can close
  // this lane's row: its pairable columns inside the window, bit q <-> u2 = q (outside) or 30 - q (inside)
  uint32_t x = 0;
  if (lane <= u1_end) {
    const int a = OUTSIDE ? i - 1 - lane : i + 1 + lane;
    const int lo = (OUTSIDE ? j + 1 : j - 1 - kMaxLoop) + 32;
    const uint32_t* row = pb + (size_t)a * pbw + (lo >> 5);
    x = __funnelshift_r(row[0], row[1], lo & 31) & 0x7fffffffu;
    x &= OUTSIDE ? (0x7fffffffu >> lane) : ~((1u << lane) - 1u);   // u2 <= 30 - u1
  }
  const int cnt = __popc(x);
  int at = cnt;   // inclusive scan of the rows' counts
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, at, o); if (lane >= o) at += v; }
  const int total = __shfl_sync(0xffffffffu, at, 31);
  at -= cnt;
  while (x) {
    const int q = __ffs((int)x) - 1;
    x &= x - 1u;
    stage[at++] = (uint16_t)((lane << 5) | (OUTSIDE ? q : kMaxLoop - q));
  }
  __syncwarp();
  for (int t = lane; t < total; t += 32) {
    const int code = stage[t];
    const int u1 = code >> 5, u2 = code & 31;
    const int a = OUTSIDE ? i - 1 - u1 : i + 1 + u1, bb = OUTSIDE ? j + 1 + u2 : j - 1 - u2;
    const int tp = s_pt[S[a] * 5 + S[bb]];
    const double v = tab[(size_t)a * W + bb];
    // il_weight(closing type, inner type, closing pair's mismatch factor, inner pair's)
    acc += v * (OUTSIDE ? il_weight(T, u1, u2, tp, type, T.mmI[tp][S[a + 1]][S[bb - 1]], mm_ij)
                        : il_weight(T, u1, u2, type, tp, mm_ij, T.mmI[d_rtype(tp)][S[bb + 1]][S[a - 1]]));
  }
  __syncwarp();   // the staging area is reused by the warp's next cell
  return acc;
}

// THREADS x CTAS: 1024 x 1 for sequences of 128 nt and more -- the fewer sequences an SM has in flight, the more of
// their tables the L2 holds (256 x 4: 108 ms, 512 x 2: 103 ms, 1024 x 1: 96 ms per 2 000 sequences of 150-300 nt) --
// and 256 x 4 for short ones, whose diagonals cannot feed 32 warps.
template <int kFoldThreads, int kFoldCtas>
__global__ void __launch_bounds__(kFoldThreads, kFoldCtas) fold_kernel(const FoldLaunch P) {
  constexpr int kFoldWarps = kFoldThreads / 32;
  extern __shared__ __align__(16) unsigned char dyn[];   // base codes of the sequence in flight: cap bytes
  __shared__ FoldTab T;
  __shared__ unsigned long long s_seq, s_base;
  __shared__ uint32_t s_wsum[kFoldWarps];
  __shared__ uint32_t s_run;
  __shared__ uint8_t s_pt[32];                        // pair type by (code of i) * 5 + (code of j)
  __shared__ uint16_t s_stage[kFoldWarps][512];       // per warp: codes (u1 << 5 | u2) of the interior-loop combinations that exist
  __shared__ int s_cell[2];   // next cell of the running diagonal (two counters, alternating: the other one is reset under the barrier)
  uint8_t* S = dyn;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    const double* src = reinterpret_cast<const double*>(P.tab);
    double* dst = reinterpret_cast<double*>(&T);
    for (int t = tid; t < (int)(sizeof(FoldTab) / sizeof(double)); t += kFoldThreads) dst[t] = src[t];
    if (tid < 32) s_pt[tid] = tid < 25 ? (uint8_t)d_pair_type(tid / 5, tid % 5, P.no_gu) : (uint8_t)0;
  }
#ifdef FOLD_PROF
  long long fprof[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const long long fp_t00 = clock64();
#endif
  const size_t cap2 = (size_t)P.cap * P.cap;
  double* base = P.scratch + (size_t)blockIdx.x * P.scratch_stride;
  const double* __restrict__ sp = P.sp;
  const double* __restrict__ up = P.up;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_seq = atomicAdd(P.counter, 1ull);
    __syncthreads();
    if (s_seq >= P.n_seqs) break;
    FP_T(t_s0);
    const uint32_t sid = P.order[s_seq];
    const uint64_t c0 = P.seq_off[sid];
    const int n = (int)(P.seq_off[sid + 1] - c0);
    const int W = n + 2;
    const size_t W2 = (size_t)W * W;
    // tables of this sequence: pitch W, index [a*W + b] with 0 <= a, b <= n + 1
    double* Qb = base;
    double* Qm = base + cap2;
    double* QmT = base + 2 * cap2;
    double* Qm1T = base + 3 * cap2;
    double* Q = base + 4 * cap2;
    double* QqT = base + 5 * cap2;
    double* Ob = base + 6 * cap2;
    double* Wc = base + 7 * cap2;
    double* AT = base + 8 * cap2;
    double* BT = base + 9 * cap2;
#define IX(a, b) ((size_t)(a) * W + (size_t)(b))
#define PT(a, b) ((b) - (a) > kTurn ? (int)s_pt[S[a] * 5 + S[b]] : 0)   /* pair type of (a,b), 1 <= a, b <= n */

    for (int t = tid; t < W; t += kFoldThreads) S[t] = (t >= 1 && t <= n) ? P.codes[c0 + t - 1] : (uint8_t)0;
    // pairable columns of every row as bits (bit l + 32 of row a: (a,l) can pair, l - a > 3): in shared memory when the
    // launch reserved room for the longest sequence (P.pb_shared), in the CTA's scratch otherwise
    const int pbw = (n + 33) / 32 + 2;   // a window starts at most at bit n + 33 and spans two words
    uint32_t* pb = P.pb_shared ? reinterpret_cast<uint32_t*>(dyn + P.pb_shared) : reinterpret_cast<uint32_t*>(base + (size_t)kTables * cap2);
    for (int tb = 0; tb < kTables; ++tb) {
      double* q = base + (size_t)tb * cap2;
      for (size_t t = tid; t < W2; t += kFoldThreads) q[t] = 0.0;
    }
    __syncthreads();
    for (int t = tid; t < W * pbw; t += kFoldThreads) {
      const int a = t / pbw, w = t % pbw;
      uint32_t bits = 0;
      if (a >= 1 && a <= n)
        for (int q = 0; q < 32; ++q) {
          const int l = 32 * w + q - 32;
          if (l <= n && l - a > kTurn && s_pt[S[a] * 5 + S[l]]) bits |= 1u << q;
        }
      pb[t] = bits;
    }
    // spans without a pair: only the exterior table is non-zero; Q(i,i-1) = 1 is the empty interval
    for (int t = tid; t < (n + 1) * (kTurn + 2); t += kFoldThreads) {
      const int i = 1 + t / (kTurn + 2), j = i - 1 + t % (kTurn + 2);
      if (j <= n) Q[IX(i, j)] = sp[j - i + 1];
    }
    if (tid == 0) { s_cell[0] = 0; s_cell[1] = 0; }
    __syncthreads();
    FP_T(t_s1);
    FP_ADD(0, t_s0, t_s1);
    // ---------------------------------------------------------------- inside
    for (int d = kTurn + 1; d < n; ++d) {
      FP_T(t_d0);
      if (tid == 0) s_cell[(d + 1) & 1] = 0;
      // cells are drawn from a counter: a pair cell costs ~4x a cell that is no pair, a static deal leaves warps idle
      for (;;) {
        int i = 0;
        if (lane == 0) i = 1 + atomicAdd(&s_cell[d & 1], 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i + d > n) break;
        const int j = i + d;
        FP_T(t_c0);
        const int type = PT(i, j);
        const double qm1_prev = Qm1T[IX(j - 1, i)], qq_prev = QqT[IX(j - 1, i)];   // asked for before the sums below need them
        double qb = 0.0;
        if (type) {
          double acc = interior_sum<false>(T, S, s_pt, pb, pbw, s_stage[warp], Qb, W, n, i, j, type, lane);
          FP_T(t_c1);
          FP_ADD(1, t_c0, t_c1);
          // multiloop: sum_k Qm(i+1,k-1) Qm1(k,j-1)
          double ml = 0.0;
          const double* qm_row = Qm + IX(i + 1, 0);
          const double* qm1_row = Qm1T + IX(j - 1, 0);
#pragma unroll 4
          for (int k = i + 2 + lane; k <= j - 1; k += 32) ml += qm_row[k - 1] * qm1_row[k];
          const int tt = d_rtype(type);
          acc += ml * (T.mlclosingS * T.mlintern[tt] * T.d3[tt][S[i + 1]] * T.d5[tt][S[j - 1]]);
          qb = warp_sum_f(acc);
          const int u = d - 1;
          qb += P.hpS[u] * (u == 3 ? T.tau[type] : T.mmH[type][S[i + 1]][S[j - 1]]);
          FP_T(t_c2);
          FP_ADD(2, t_c1, t_c2);
        }
        FP_T(t_c3);
        double dang = 1.0;
        if (type) dang = (i > 1 ? T.d5[type][S[i - 1]] : 1.0) * (j < n ? T.d3[type][S[j + 1]] : 1.0);
        const double qm1 = qm1_prev * T.u1 + (type ? qb * T.mlintern[type] * dang : 0.0);
        const double qq = qq_prev * T.s1 + (type ? qb * T.tau[type] * dang : 0.0);
        // Qm, Q: the k = i terms are this cell's own Qm1 / Qq
        double sm = 0.0, sq = 0.0;
        const double* qmi = Qm + IX(i, 0);
        const double* qi = Q + IX(i, 0);
        const double* qm1j = Qm1T + IX(j, 0);
        const double* qqj = QqT + IX(j, 0);
#pragma unroll 4
        for (int k = i + 1 + lane; k <= j; k += 32) {
          sm += (up[k - i] + qmi[k - 1]) * qm1j[k];
          sq += qi[k - 1] * qqj[k];
        }
        sm = warp_sum_f(sm) + qm1;
        sq = warp_sum_f(sq) + qq + sp[d + 1];
        if (lane == 0) {
          Qb[IX(i, j)] = qb;
          Qm1T[IX(j, i)] = qm1;
          QqT[IX(j, i)] = qq;
          Qm[IX(i, j)] = sm;
          QmT[IX(j, i)] = sm;
          Q[IX(i, j)] = sq;
        }
        FP_T(t_c4);
        FP_ADD(3, t_c3, t_c4);
#ifdef FOLD_PROF
        if (lane == 0) { fprof[8] += 1; if (type) fprof[9] += 1; }
#endif
      }
      FP_T(t_d1);
      __syncthreads();
      FP_T(t_d2);
      FP_ADD(4, t_d0, t_d1);
      FP_ADD(5, t_d1, t_d2);
    }
    FP_T(t_o0);
    const double Z = n > 0 ? Q[IX(1, n)] : 1.0;
    const bool bad = !(Z > 0.0) || isinf(Z) || isnan(Z);

    // ---------------------------------------------------------------- outside
    if (tid == 0) { s_cell[0] = 0; s_cell[1] = 0; }
    __syncthreads();
    for (int d = n - 1; d >= 1 && !bad; --d) {
      if (tid == 0) s_cell[(d + 1) & 1] = 0;
      for (;;) {
        int i = 0;
        if (lane == 0) i = 1 + atomicAdd(&s_cell[d & 1], 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i + d > n) break;
        const int j = i + d;
        const int type = PT(i, j);
        if (type) {
          double acc = interior_sum<true>(T, S, s_pt, pb, pbw, s_stage[warp], Ob, W, n, i, j, type, lane);
          double ml = 0.0;
          const double* qmt = QmT + IX(i - 1, 0);   // QmT[i-1][p+1] = Qm(p+1, i-1)
          const double* at = AT + IX(j, 0);
          const double* bt = BT + IX(j, 0);
#pragma unroll 4
          for (int p = 1 + lane; p < i; p += 32) ml += qmt[p + 1] * at[p] + up[i - p - 1] * bt[p];
          const double dang = (i > 1 ? T.d5[type][S[i - 1]] : 1.0) * (j < n ? T.d3[type][S[j + 1]] : 1.0);
          acc += ml * (T.mlintern[type] * dang);
          double ob = warp_sum_f(acc);
          ob += Q[IX(1, i - 1)] * Q[IX(j + 1, n)] * (T.tau[type] * dang);
          if (lane == 0) {
            const int tt = d_rtype(type);
            Ob[IX(i, j)] = ob;
            Wc[IX(i, j)] = ob * (T.mlclosingS * T.mlintern[tt] * T.d3[tt][S[i + 1]] * T.d5[tt][S[j - 1]]);
          }
        }
        if (j < n) {
          double a = 0.0, b = 0.0;
          const double* wrow = Wc + IX(i, 0);
          const double* qmr = Qm + IX(j + 1, 0);
#pragma unroll 4
          for (int q = j + 1 + lane; q <= n; q += 32) {
            const double w = wrow[q], qm = qmr[q - 1];   // both asked for at once: W is mostly 0, but a dependent load costs a round trip
            a += w * (up[q - j - 1] + qm);
            b += w * qm;
          }
          a = warp_sum_f(a);
          b = warp_sum_f(b);
          if (lane == 0) { AT[IX(j, i)] = a; BT[IX(j, i)] = b; }
        }
      }
      __syncthreads();
    }

    FP_T(t_o1);
    FP_ADD(6, t_o0, t_o1);
    // ---------------------------------------------------------------- probabilities, per-position sums, pair lists
    const double invZ = bad ? 0.0 : 1.0 / Z;
    for (size_t t = tid; t < W2; t += kFoldThreads) {   // Ob becomes P
      const int i = (int)(t / W), j = (int)(t % W);
      Ob[t] = (i >= 1 && j <= n && PT(i, j)) ? Qb[t] * Ob[t] * invZ : 0.0;
    }
    __syncthreads();
    for (int t = 1 + tid; t <= n; t += kFoldThreads) {
      double s = 0.0;
      for (int q = t + 1; q <= n; ++q) s += Ob[IX(t, q)];
      for (int p = 1; p < t; ++p) s += Ob[IX(p, t)];
      P.unpaired[c0 + t - 1] = fmax(0.0, 1.0 - s);
    }
    if (P.dense) {
      double* dd = P.dense + P.dense_off[sid];
      const int W1 = n + 1;
      for (size_t t = tid; t < (size_t)W1 * W1; t += kFoldThreads) {
        const int i = (int)(t / W1), j = (int)(t % W1);
        dd[t] = Ob[IX(i, j)];
      }
    }
    const double cut = P.cutoff;
    auto listed = [&](double p) { return p > 0.0 && p >= cut; };
    uint32_t mine = 0;
    for (size_t t = tid; t < W2; t += kFoldThreads) mine += listed(Ob[t]) ? 1u : 0u;
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (lane == 0) s_wsum[warp] = mine;
    __syncthreads();
    if (tid == 0) {
      uint32_t total = 0;
      for (int w = 0; w < kFoldWarps; ++w) total += s_wsum[w];
      s_base = atomicAdd(P.counter + 1, (unsigned long long)total);
      P.out_start[sid] = s_base;
      P.out_count[sid] = total;
      P.ensemble[sid] = bad ? 0.0 : -T.kT * (log(Z) + n * T.log_s);
      P.status[sid] = bad ? 1 : 0;
      s_run = 0;
    }
    __syncthreads();
    // row-major walk in chunks of the block: ascending (i, j)
    for (size_t t0 = 0; t0 < W2; t0 += kFoldThreads) {
      const size_t t = t0 + tid;
      const double p = t < W2 ? Ob[t] : 0.0;
      const bool f = listed(p);
      const unsigned m = __ballot_sync(0xffffffffu, f);
      if (lane == 0) s_wsum[warp] = __popc(m);
      __syncthreads();
      uint32_t before = 0, total = 0;
      for (int w = 0; w < kFoldWarps; ++w) { if (w < warp) before += s_wsum[w]; total += s_wsum[w]; }
      if (f) {
        const unsigned long long at = s_base + s_run + before + __popc(m & ((1u << lane) - 1u));
        P.out_i[at] = (uint32_t)(t / W);
        P.out_j[at] = (uint32_t)(t % W);
        P.out_p[at] = p;
      }
      __syncthreads();
      if (tid == 0) s_run += total;
    }
    FP_T(t_o2);
    FP_ADD(7, t_o1, t_o2);
#undef IX
#undef PT
  }
#ifdef FOLD_PROF
  if (lane == 0 && P.prof) {
    for (int q = 0; q < 10; ++q) atomicAdd(P.prof + q, (unsigned long long)fprof[q]);
    atomicAdd(P.prof + 10, (unsigned long long)(clock64() - fp_t00));
  }
#endif
}

uint8_t base_code(char c) {
  switch (c) {
    case 'a': case 'A': return 1;
    case 'c': case 'C': return 2;
    case 'g': case 'G': return 3;
    case 'u': case 'U': case 't': case 'T': return 4;
    default: return 0;
  }
}

}  // namespace

cudaError_t run_fold(const stemk_fold_model& m, uint32_t n_seqs, const uint64_t* seq_off, const char* text, double cutoff,
                     bool want_dense, int sm_count, cudaStream_t stream, FoldResult* res, std::string* err, void** scratch_p,
                     size_t* scratch_bytes) {
  res->pair_off.assign((size_t)n_seqs + 1, 0);
  res->bi.clear(); res->bj.clear(); res->bp.clear(); res->unpaired.clear(); res->ensemble.assign(n_seqs, 0.0); res->dense.clear();
  if (n_seqs == 0) return cudaSuccess;
  const uint64_t n_chars = seq_off[n_seqs];
  uint32_t max_len = 0;
  for (uint32_t k = 0; k < n_seqs; ++k) {
    if (seq_off[k + 1] < seq_off[k]) { if (err) *err = "seq_off must be non-decreasing"; return cudaErrorInvalidValue; }
    if (seq_off[k + 1] - seq_off[k] > 30000) { if (err) *err = "sequence longer than 30000 characters"; return cudaErrorInvalidValue; }
    max_len = std::max<uint32_t>(max_len, (uint32_t)(seq_off[k + 1] - seq_off[k]));
  }
  const uint32_t cap = max_len + 2;
  // ---- Boltzmann factors (libm exp, like the checker)
  const double kT = (m.temperature + 273.15) * 1.98717e-3;
  const double s = m.pf_scale > 0 ? m.pf_scale : std::exp(-(-185.0 + (m.temperature - 37.0) * 7.27) / (1000.0 * kT));
  auto B = [&](double e) { return std::exp(-e / kT); };
  std::vector<double> sp(cap + 3), up(cap + 3), hpS(cap + 1, 0.0);
  const double u1 = B(m.ml_base) / s;
  sp[0] = up[0] = 1.0;
  for (size_t k = 1; k < sp.size(); ++k) { sp[k] = sp[k - 1] / s; up[k] = up[k - 1] * u1; }
  for (uint32_t u = 3; u <= cap; ++u)
    hpS[u] = B(u <= 30 ? m.hairpin[u] : m.hairpin[30] + m.lxc * std::log(u / 30.0)) * (u + 2 < sp.size() ? sp[u + 2] : 0.0);
  FoldTab T;
  std::memset(&T, 0, sizeof(T));
  for (int a = 0; a < 8; ++a) {
    for (int b = 0; b < 8; ++b) { T.stackB[a][b] = B(m.stack[a][b]); T.stackS[a][b] = T.stackB[a][b] * sp[2]; }
    for (int x = 0; x < 5; ++x) {
      for (int y = 0; y < 5; ++y) { T.mmH[a][x][y] = B(m.mismatch_h[a][x][y]); T.mmI[a][x][y] = B(m.mismatch_i[a][x][y]); }
      T.d5[a][x] = B(m.dangle5[a][x]); T.d3[a][x] = B(m.dangle3[a][x]);
    }
    T.mlintern[a] = B(m.ml_intern[a]);
    T.tau[a] = a > 2 ? B(m.terminal_au) : 1.0;
  }
  for (int u = 0; u <= 30; ++u) {
    T.bulgeS[u] = B(m.bulge[u]) * sp[u + 2];
    T.interiorS[u] = B(m.interior[u]) * sp[u + 2];
    T.ninioB[u] = B(std::min(m.max_ninio, u * m.ninio));
  }
  T.mlclosingS = B(m.ml_closing) * sp[2];
  T.u1 = u1; T.s1 = 1.0 / s; T.kT = kT; T.log_s = std::log(s);

  std::vector<uint8_t> codes(n_chars);
  for (uint64_t c = 0; c < n_chars; ++c) codes[c] = base_code(text[c]);
  std::vector<uint32_t> order(n_seqs);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return seq_off[a + 1] - seq_off[a] > seq_off[b + 1] - seq_off[b]; });
  // capacity of the pair lists: sum_j P(i,j) <= 1 bounds the pairs at or above the cut-off by L / (2 cutoff)
  uint64_t out_cap = 0;
  std::vector<uint64_t> dense_off(n_seqs, 0);
  uint64_t dense_total = 0;
  for (uint32_t k = 0; k < n_seqs; ++k) {
    const uint64_t L = seq_off[k + 1] - seq_off[k];
    const uint64_t all = L * L / 2 + 1;
    out_cap += cutoff > 0 ? std::min<uint64_t>(all, (uint64_t)(L / (2.0 * cutoff)) + 2) : all;
    dense_off[k] = dense_total;
    dense_total += (L + 1) * (L + 1);
  }

  std::vector<void*> to_free;
  auto cleanup = [&]() { for (void* q : to_free) cudaFree(q); to_free.clear(); };
  struct Guard { decltype(cleanup)& f; ~Guard() { f(); } } guard{cleanup};   // also on a host-side std::bad_alloc below
  cudaError_t e = cudaSuccess;
  auto dalloc = [&](void** d, size_t bytes) -> bool {
    *d = nullptr;
    e = cudaMalloc(d, std::max<size_t>(bytes, 16));
    if (e != cudaSuccess) return false;
    to_free.push_back(*d);
    return true;
  };
  auto upload = [&](void** d, const void* h, size_t bytes) -> bool {
    if (!dalloc(d, bytes)) return false;
    if (bytes) e = cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, stream);
    return e == cudaSuccess;
  };
  FoldLaunch L;
  std::memset(&L, 0, sizeof(L));
  void *d_tab, *d_hp, *d_sp, *d_up, *d_codes, *d_off, *d_order, *d_cnt, *d_start, *d_count, *d_i, *d_j, *d_p, *d_unp, *d_ens, *d_status,
      *d_scratch, *d_dense = nullptr, *d_doff = nullptr;
  const uint32_t pbw_cap = (max_len + 33) / 32 + 2;
  const size_t pb_bytes = (size_t)cap * pbw_cap * 4;
  const unsigned long long stride = ((unsigned long long)kTables * cap * cap + (pb_bytes + 7) / 8 + 1) & ~1ull;
  const bool big_cta = max_len >= 128;
  const int fold_threads = big_cta ? 1024 : 256, fold_ctas = big_cta ? 1 : 4;
  (void)fold_threads;
  int grid = (int)std::min<uint64_t>(n_seqs, (uint64_t)sm_count * fold_ctas);
  grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)grid, ((uint64_t)8 << 30) / (stride * sizeof(double))));
  bool ok = upload(&d_tab, &T, sizeof(T)) && upload(&d_hp, hpS.data(), hpS.size() * 8) && upload(&d_sp, sp.data(), sp.size() * 8) &&
            upload(&d_up, up.data(), up.size() * 8) && upload(&d_codes, codes.data(), codes.size()) &&
            upload(&d_off, seq_off, ((size_t)n_seqs + 1) * 8) && upload(&d_order, order.data(), (size_t)n_seqs * 4) &&
            dalloc(&d_cnt, 16) && dalloc(&d_start, (size_t)n_seqs * 8) && dalloc(&d_count, (size_t)n_seqs * 4) &&
            dalloc(&d_i, out_cap * 4) && dalloc(&d_j, out_cap * 4) && dalloc(&d_p, out_cap * 8) && dalloc(&d_unp, n_chars * 8) &&
            dalloc(&d_ens, (size_t)n_seqs * 8) && dalloc(&d_status, (size_t)n_seqs * 4);
  if (ok && *scratch_bytes < stride * sizeof(double) * grid) {   // the DP tables of the sequences in flight: kept by the context
    if (*scratch_p) cudaFree(*scratch_p);
    *scratch_p = nullptr; *scratch_bytes = 0;
    e = cudaMalloc(scratch_p, stride * sizeof(double) * grid);
    ok = e == cudaSuccess;
    if (ok) *scratch_bytes = stride * sizeof(double) * grid;
  }
  d_scratch = *scratch_p;
  if (ok && want_dense) ok = dalloc(&d_dense, dense_total * 8) && upload(&d_doff, dense_off.data(), (size_t)n_seqs * 8);
  if (ok) { e = cudaMemsetAsync(d_cnt, 0, 16, stream); ok = e == cudaSuccess; }
  if (!ok) { cleanup(); if (err && e == cudaErrorMemoryAllocation) *err = "base-pair probabilities: out of device memory (split the batch or raise the cut-off)"; return e; }
  L.tab = (const FoldTab*)d_tab; L.hpS = (const double*)d_hp; L.sp = (const double*)d_sp; L.up = (const double*)d_up;
  L.codes = (const uint8_t*)d_codes; L.seq_off = (const uint64_t*)d_off; L.order = (const uint32_t*)d_order;
  L.n_seqs = n_seqs; L.cap = cap; L.no_gu = m.no_gu ? 1 : 0; L.cutoff = cutoff;
  L.scratch = (double*)d_scratch; L.scratch_stride = stride; L.counter = (unsigned long long*)d_cnt;
  L.out_start = (uint64_t*)d_start; L.out_count = (uint32_t*)d_count; L.out_i = (uint32_t*)d_i; L.out_j = (uint32_t*)d_j;
  L.out_p = (double*)d_p; L.unpaired = (double*)d_unp; L.ensemble = (double*)d_ens; L.status = (int*)d_status;
  L.dense = (double*)d_dense; L.dense_off = (const uint64_t*)d_doff;
#ifdef FOLD_PROF
  unsigned long long* d_prof = nullptr;
  cudaMalloc((void**)&d_prof, 16 * sizeof(unsigned long long));
  cudaMemsetAsync(d_prof, 0, 16 * sizeof(unsigned long long), stream);
  L.prof = d_prof;
#endif
  size_t smem = ((size_t)cap + 15) & ~(size_t)15;
  L.pb_shared = 0;
  if (pb_bytes <= 40 * 1024) { L.pb_shared = (uint32_t)smem; smem += pb_bytes; }   // longer sequences: the bits stay in the scratch (L1/L2)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEventCreate(&ev0); cudaEventCreate(&ev1);
  cudaEventRecord(ev0, stream);
  if (big_cta) {
    cudaFuncSetAttribute(fold_kernel<1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    fold_kernel<1024, 1><<<grid, 1024, smem, stream>>>(L);
  } else {
    cudaFuncSetAttribute(fold_kernel<256, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    fold_kernel<256, 4><<<grid, 256, smem, stream>>>(L);
  }
  e = cudaGetLastError();
  cudaEventRecord(ev1, stream);
  // ---- results back to the host, pair lists put in sequence order
  std::vector<uint64_t> start(n_seqs);
  std::vector<uint32_t> count(n_seqs);
  std::vector<int> status(n_seqs);
  unsigned long long cnt[2] = {0, 0};
  if (e == cudaSuccess) e = cudaMemcpyAsync(cnt, d_cnt, 16, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(start.data(), d_start, (size_t)n_seqs * 8, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(count.data(), d_count, (size_t)n_seqs * 4, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(status.data(), d_status, (size_t)n_seqs * 4, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(res->ensemble.data(), d_ens, (size_t)n_seqs * 8, cudaMemcpyDeviceToHost, stream);
  res->unpaired.resize(n_chars);
  if (e == cudaSuccess && n_chars) e = cudaMemcpyAsync(res->unpaired.data(), d_unp, n_chars * 8, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess && want_dense) {
    res->dense.resize(dense_total);
    e = cudaMemcpyAsync(res->dense.data(), d_dense, dense_total * 8, cudaMemcpyDeviceToHost, stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  res->kernel_ms = 0.0;
  if (e == cudaSuccess) { float ms = 0; if (cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess) res->kernel_ms = ms; }
  cudaEventDestroy(ev0); cudaEventDestroy(ev1);
#ifdef FOLD_PROF
  {
    unsigned long long h[16];
    cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d_prof);
    const double tot = h[10] ? (double)h[10] : 1.0, cells = h[8] ? (double)h[8] : 1.0, pc = h[9] ? (double)h[9] : 1.0;
    std::fprintf(stderr, "fold prof: grid %d x %d threads, %u seqs | of all warp time: setup %.1f%% inside cells %.1f%% inside barrier %.1f%% outside %.1f%% output %.1f%% | inside, cycles per cell: interior %.0f (per pair cell) ml+hairpin %.0f (per pair cell) Qm/Q sums %.0f (per cell); cells %.0f of which pairs %.0f\n",
                 grid, fold_threads, n_seqs, 100.0 * h[0] / tot, 100.0 * h[4] / tot, 100.0 * h[5] / tot, 100.0 * h[6] / tot, 100.0 * h[7] / tot,
                 h[1] / pc, h[2] / pc, h[3] / cells, cells, pc);
  }
#endif
  if (e != cudaSuccess) { cleanup(); return e; }
  const uint64_t total = cnt[1];
  std::vector<uint32_t> ti(total), tj(total);
  std::vector<double> tp(total);
  if (total) {
    e = cudaMemcpy(ti.data(), d_i, total * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(tj.data(), d_j, total * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(tp.data(), d_p, total * 8, cudaMemcpyDeviceToHost);
  }
  cleanup();
  if (e != cudaSuccess) return e;
  for (uint32_t k = 0; k < n_seqs; ++k)
    if (status[k]) {
      if (err) *err = "base-pair probabilities: the partition function of sequence " + std::to_string(k) + " is out of range under this pf_scale";
      return cudaErrorInvalidValue;
    }
  res->bi.resize(total); res->bj.resize(total); res->bp.resize(total);
  uint64_t at = 0;
  for (uint32_t k = 0; k < n_seqs; ++k) {
    res->pair_off[k] = at;
    std::copy_n(ti.begin() + start[k], count[k], res->bi.begin() + at);
    std::copy_n(tj.begin() + start[k], count[k], res->bj.begin() + at);
    std::copy_n(tp.begin() + start[k], count[k], res->bp.begin() + at);
    at += count[k];
  }
  res->pair_off[n_seqs] = at;
  return cudaSuccess;
}

}  // namespace stemk

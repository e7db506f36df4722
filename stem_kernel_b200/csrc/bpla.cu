// stem_kernel_b200/csrc/bpla.cu -- BPLA / local-alignment kernels.
//
// BPLAKernel<double, MData>::operator() of bpla_kernel/bpla_kernel.cpp:160-175: the sum over all local alignments
// (local_alignment_exp, :64-118, five tables M, X, Y, X2, Y2) or the Smith-Waterman score (local_alignment_max,
// :120-157), with the plain substitution score (LAScore, :16-45) or the base-pairing profile score (BPLAScore,
// :47-62).  Restated for a row-by-row sweep:
//   * M, X, X2 of row i depend on row i-1 only, so a row is computed 32 columns at a time, lanes <-> columns;
//   * Y(i,j) = e^{beta gap} (M(i,j-1) + X(i,j-1)) + e^{beta ext} Y(i,j-1) is a first-order linear recurrence along
//     the row: an inclusive warp scan with the powers of e^{beta ext}, the carry handed from chunk to chunk
//     (max-plus scan with multiples of ext for Smith-Waterman);
//   * X2(i,j) = sum_{i'<i} M(i',j) is a per-column accumulator and Y2 is only ever read at (Lx, Ly), where it is
//     sum_{j<Ly} (M(Lx,j) + X2(Lx,j)): no Y2 table;
//   * the float / double mix of the two score functors is kept operation by operation, so a cell's score is the
//     reference's bit for bit; exp() is CUDA's (<= 1 ulp), the scans re-associate the Y recurrence: results agree
//     with the reference to ~1e-13 relative, the parity gate is 1e-9.
// One warp per pair, persistent grid with an atomic work queue; a warp's rows (previous and current M, X, Y, X2)
// live in shared memory, 64 bytes per column.
#include <algorithm>

#include "kernels.cuh"

namespace stemk {

namespace {

constexpr int kBplaWarps = 4;

struct BplaDev {            // one record set on the device
  const uint32_t* col_off;
  const float* profile;    // 5 per column
  const float* pl;
  const float* pr;
  const float* pu;
  const uint8_t* code;     // per column: 0..3 one-hot base (count 1), 4 all-zero column, 5 general profile
};

struct BplaLaunch {
  BplaDev X, Y;
  const uint32_t* xi;
  const uint32_t* yi;
  unsigned long long n_pairs;
  unsigned long long* counter;
  double* out;
  double gap, ext, alpha, beta;
  double score[16];
  int no_bp, sw;
  uint32_t ly_cap;
};

__device__ __forceinline__ double la_score(const double* __restrict__ tab, const float* x, const float* y) {
  double v = 0.0;
  float n = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (x[k] == 0.f) continue;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      if (y[l] == 0.f) continue;
      n = __fadd_rn(n, __fmul_rn(x[k], y[l]));                      // float, no contraction (bpla_kernel.cpp:36)
      v = __dadd_rn(v, __dmul_rn(__dmul_rn(tab[k * 4 + l], (double)x[k]), (double)y[l]));   // :37
    }
  }
  return n == 0.f ? 0.0 : v / (double)n;
}

template <bool sw, bool no_bp>
__global__ void __launch_bounds__(32 * kBplaWarps) bpla_pairs_kernel(const BplaLaunch P) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double s_tab[16];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (threadIdx.x < 16) s_tab[threadIdx.x] = P.score[threadIdx.x];
  __syncthreads();
  const uint32_t W = P.ly_cap + 1u;                 // row length in shared memory
  double* rows = sm + (size_t)warp * 8u * W;         // [prev M, X, Y, X2 | cur M, X, Y, X2]
  const double bg = sw ? P.gap : exp(P.beta * P.gap), be = sw ? P.ext : exp(P.beta * P.ext);
  // per lane: be^(lane+1) (sum form) or (lane+1)*ext (max form); per scan step o: be^o or o*ext
  double pw_lane = sw ? 0.0 : 1.0, pw_o[5];
  for (uint32_t k = 0; k <= lane; ++k) pw_lane = sw ? pw_lane + be : pw_lane * be;
  {
    double t = be;
    for (int s = 0; s < 5; ++s) { pw_o[s] = t; t = sw ? t + t : t * t; }
  }

  for (;;) {
    unsigned long long k = 0;
    if (lane == 0) k = atomicAdd(P.counter, 1ull);
    k = __shfl_sync(0xffffffffu, k, 0);
    if (k >= P.n_pairs) break;
    const uint32_t xr = P.xi[k], yr = P.yi[k];
    const uint32_t x0 = P.X.col_off[xr], lx = P.X.col_off[xr + 1] - x0;
    const uint32_t y0 = P.Y.col_off[yr], ly = P.Y.col_off[yr + 1] - y0;
    double *pM = rows, *pX = rows + W, *pY = rows + 2 * W, *pX2 = rows + 3 * W;
    double *cM = rows + 4 * W, *cX = rows + 5 * W, *cY = rows + 6 * W, *cX2 = rows + 7 * W;
    for (uint32_t j = lane; j <= ly; j += 32u) { pM[j] = 0.0; pX[j] = 0.0; pY[j] = 0.0; pX2[j] = 0.0; }
    if (lane == 0) { cM[0] = 0.0; cX[0] = 0.0; cY[0] = 0.0; cX2[0] = 0.0; }
    __syncwarp();
    double mmax = 0.0;
    for (uint32_t i = 1; i <= lx; ++i) {
      // the x column of this row (uniform)
      float xc[4], xpl = 0.f, xpr = 0.f, xpu = 0.f;
      const uint32_t xcode = __ldg(P.X.code + x0 + i - 1u);
#pragma unroll
      for (int a = 0; a < 4; ++a) xc[a] = __ldg(P.X.profile + 5u * (x0 + i - 1u) + a);
      if (!no_bp) { xpl = __ldg(P.X.pl + x0 + i - 1u); xpr = __ldg(P.X.pr + x0 + i - 1u); xpu = __ldg(P.X.pu + x0 + i - 1u); }
      double carry = 0.0;   // Y(i, j0 - 1) of the chunk's first column (Y(i,0) = 0)
      for (uint32_t j0 = 1; j0 <= ly; j0 += 32u) {
        const uint32_t j = j0 + lane;
        const bool act = j <= ly;
        double m = 0.0, xv = 0.0, x2 = 0.0;
        if (act) {
          // one-hot columns (count exactly 1): n = 1, v = table entry, v / n = the entry bit for bit; an all-zero
          // column gives n = 0 and the score 0 (bpla_kernel.cpp:31-44); everything else takes the general loop
          const uint32_t ycode = __ldg(P.Y.code + y0 + j - 1u);
          double s;
          if (xcode < 4u && ycode < 4u) s = s_tab[xcode * 4u + ycode];
          else if (xcode == 4u || ycode == 4u) s = 0.0;
          else {
            float yc[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) yc[a] = __ldg(P.Y.profile + 5u * (y0 + j - 1u) + a);
            s = la_score(s_tab, xc, yc);
          }
          if (!no_bp) {
            const float ypl = __ldg(P.Y.pl + y0 + j - 1u), ypr = __ldg(P.Y.pr + y0 + j - 1u), ypu = __ldg(P.Y.pu + y0 + j - 1u);
            const float pair = __fadd_rn(__fmul_rn(xpr, ypr), __fmul_rn(xpl, ypl));      // bpla_kernel.cpp:58
            s = __dadd_rn(__dmul_rn(P.alpha, (double)pair), __dmul_rn((double)__fmul_rn(xpu, ypu), s));   // :58-59
          }
          if (!sw) {
            m = exp(P.beta * s) * (((1.0 + pX[j - 1]) + pY[j - 1]) + pM[j - 1]);       // :103-104
            xv = bg * pM[j] + be * pX[j];                                               // :105-106
            x2 = pM[j] + pX2[j];                                                        // :111
          } else {
            m = fmax(fmax(fmax(0.0, pM[j - 1]), pX[j - 1]), pY[j - 1]) + s;            // :143-146
            mmax = fmax(mmax, m);
            xv = fmax(pM[j] + bg, pX[j] + be);                                          // :148
          }
          cM[j] = m; cX[j] = xv; cX2[j] = x2;
        }
        __syncwarp();
        // Y(i,j) from the left neighbours of the same row: inclusive scan of b_j = f(M(i,j-1), X(i,j-1))
        double v = sw ? -1e300 : 0.0;
        if (act) v = sw ? fmax(cM[j - 1] + bg, cX[j - 1] + bg) : bg * (cM[j - 1] + cX[j - 1]);   // :108-109 / :149
#pragma unroll
        for (int s = 0; s < 5; ++s) {
          const double t = __shfl_up_sync(0xffffffffu, v, 1u << s);
          if (lane >= (1u << s)) v = sw ? fmax(v, t + pw_o[s]) : v + pw_o[s] * t;
        }
        v = sw ? fmax(v, carry + pw_lane) : v + pw_lane * carry;
        if (act) cY[j] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
        __syncwarp();
      }
      double* t;
      t = pM; pM = cM; cM = t;  t = pX; pX = cX; cX = t;  t = pY; pY = cY; cY = t;  t = pX2; pX2 = cX2; cX2 = t;
    }
    // result (:113-114): 1 + X2(Lx,Ly) + Y2(Lx,Ly) + M(Lx,Ly), Y2(Lx,Ly) = sum_{j<Ly} (M(Lx,j) + X2(Lx,j)), in j order per lane
    double res;
    if (!sw) {
      double y2 = 0.0;
      if (lx > 0) for (uint32_t j = 1u + lane; j < ly; j += 32u) y2 += pM[j] + pX2[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) y2 += __shfl_xor_sync(0xffffffffu, y2, o);
      res = (lx > 0 && ly > 0) ? ((1.0 + pX2[ly]) + y2) + pM[ly] : 1.0;
    } else {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mmax = fmax(mmax, __shfl_xor_sync(0xffffffffu, mmax, o));
      res = mmax;
    }
    if (lane == 0) P.out[k] = res;
    __syncwarp();
  }
}


// ---- BPLAKernel::compute_gradients (bpla_kernel.cpp:176-402) -----------------------------------------------------
// Forward tables M, IX, IY (BPLA_Forward :178-243; LX / LY are closed forms: LX(i,j) = [i >= 1 or j == 0],
// LY(i,j) = j for i >= 1, [i == 0] on the border, so the start term of a cell is 2, 1 or j), backward tables
// (BPLA_Backward :245-314, scatter form in the reference) restated as a gather:
//     D (i,j) = e^{beta s(i+1,j+1)} BM(i+1,j+1)
//     BIY(i,j) = D + e^{beta ext} BIY(i,j+1)                       (suffix recurrence along the row: warp scan)
//     BIX(i,j) = D + e^{beta ext} BIX(i+1,j) + e^{beta gap} BIY(i,j+1)
//     BM (i,j) = D + e^{beta gap} (BIX(i+1,j) + BIY(i,j+1)) + 1    (RX == 1 and RY == [i == Lx] on the grid)
// and the four partial derivatives (BPLA_ForwardBackword :335-385) accumulated in the backward sweep, lanes <-> columns.
// One warp per pair; the forward tables and e^{beta s} of the pair live in a per-warp global scratch (L2).
struct BplaGradLaunch {
  BplaLaunch B;
  double* scratch;                 // per warp: 4 tables of (lx_cap+1) x (ly_cap+1) doubles
  unsigned long long scratch_stride;
  double* grad;                    // 4 per pair: d/d{alpha, beta, gap, ext}
};

__global__ void __launch_bounds__(32 * kBplaWarps) bpla_grad_kernel(const BplaGradLaunch G) {
  const BplaLaunch& P = G.B;
  extern __shared__ __align__(16) double sm[];
  __shared__ double s_tab[16];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (threadIdx.x < 16) s_tab[threadIdx.x] = P.score[threadIdx.x];
  __syncthreads();
  const uint32_t W = P.ly_cap + 2u;                 // row length in shared memory (columns 0 .. ly+1)
  double* rows = sm + (size_t)warp * 8u * W;
  double* scr = G.scratch + (size_t)(blockIdx.x * kBplaWarps + warp) * G.scratch_stride;
  const double bg = exp(P.beta * P.gap), be = exp(P.beta * P.ext);
  double pw_lane = 1.0, pw_o[5];
  for (uint32_t k = 0; k <= lane; ++k) pw_lane *= be;
  {
    double t = be;
    for (int s = 0; s < 5; ++s) { pw_o[s] = t; t = t * t; }
  }
  // score of cell (x column loaded by the caller, y column cy): w_pair (a float expression) and w_unpair
  auto cell = [&](uint32_t xcode, const float* xc, float xpl, float xpr, float xpu, uint32_t cy, double* wp, double* wu) {
    const uint32_t ycode = __ldg(P.Y.code + cy);
    double la;
    if (xcode < 4u && ycode < 4u) la = s_tab[xcode * 4u + ycode];
    else if (xcode == 4u || ycode == 4u) la = 0.0;
    else {
      float yc[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) yc[a] = __ldg(P.Y.profile + 5u * cy + a);
      la = la_score(s_tab, xc, yc);
    }
    const float ypl = __ldg(P.Y.pl + cy), ypr = __ldg(P.Y.pr + cy), ypu = __ldg(P.Y.pu + cy);
    *wp = (double)__fadd_rn(__fmul_rn(xpr, ypr), __fmul_rn(xpl, ypl));       // bpla_kernel.cpp:357
    *wu = __dmul_rn((double)__fmul_rn(xpu, ypu), la);                          // :358
  };

  for (;;) {
    unsigned long long k = 0;
    if (lane == 0) k = atomicAdd(P.counter, 1ull);
    k = __shfl_sync(0xffffffffu, k, 0);
    if (k >= P.n_pairs) break;
    const uint32_t xr = P.xi[k], yr = P.yi[k];
    const uint32_t x0 = P.X.col_off[xr], lx = P.X.col_off[xr + 1] - x0;
    const uint32_t y0 = P.Y.col_off[yr], ly = P.Y.col_off[yr + 1] - y0;
    const size_t T = ly + 1u, plane = (size_t)(lx + 1u) * T;
    double *FM = scr, *FIX = scr + plane, *FIY = scr + 2 * plane, *BS = scr + 3 * plane;
    double *pM = rows, *pX = rows + W, *pY = rows + 2 * W, *pR = rows + 3 * W;
    double *cM = rows + 4 * W, *cX = rows + 5 * W, *cY = rows + 6 * W, *cR = rows + 7 * W;
    // ---- forward (:178-243)
    for (uint32_t j = lane; j <= ly; j += 32u) {
      const double m0 = j == 0u ? 1.0 : 0.0;          // F[M][0][0] = 1 (:196)
      pM[j] = m0; pX[j] = 0.0; pY[j] = 0.0; pR[j] = 0.0;
      FM[j] = m0; FIX[j] = 0.0; FIY[j] = 0.0;
    }
    __syncwarp();
    for (uint32_t i = 1; i <= lx; ++i) {
      float xc[4];
      const uint32_t xcode = __ldg(P.X.code + x0 + i - 1u);
#pragma unroll
      for (int a = 0; a < 4; ++a) xc[a] = __ldg(P.X.profile + 5u * (x0 + i - 1u) + a);
      const float xpl = __ldg(P.X.pl + x0 + i - 1u), xpr = __ldg(P.X.pr + x0 + i - 1u), xpu = __ldg(P.X.pu + x0 + i - 1u);
      if (lane == 0) { cM[0] = 0.0; cX[0] = 0.0; cY[0] = 0.0; cR[0] = 0.0; FM[i * T] = 0.0; FIX[i * T] = 0.0; FIY[i * T] = 0.0; }
      __syncwarp();
      double carry = 0.0;
      for (uint32_t j0 = 1; j0 <= ly; j0 += 32u) {
        const uint32_t j = j0 + lane;
        const bool act = j <= ly;
        if (act) {
          double wp, wu;
          cell(xcode, xc, xpl, xpr, xpu, y0 + j - 1u, &wp, &wu);
          const double bs = exp(P.beta * __dadd_rn(__dmul_rn(P.alpha, wp), wu));     // :210-213
          // LX + LY of the diagonal neighbour (i-1, j-1): 2 at the origin, 1 on the borders, 1 + (j-1) inside
          const double st = (i == 1u && j == 1u) ? 2.0 : ((i == 1u || j == 1u) ? 1.0 : (double)j);
          const double m = bs * (((pM[j - 1] + pX[j - 1]) + pY[j - 1]) + st);         // :214-218
          const double xv = bg * pM[j] + be * pX[j];                                    // :220-221
          cM[j] = m; cX[j] = xv; cR[j] = pM[j] + pR[j];                                 // :232-233
          FM[i * T + j] = m; FIX[i * T + j] = xv; BS[i * T + j] = bs;
        }
        __syncwarp();
        double v = 0.0;
        if (act) v = bg * (cM[j - 1] + cX[j - 1]);                                      // :223-225
#pragma unroll
        for (int s = 0; s < 5; ++s) {
          const double t = __shfl_up_sync(0xffffffffu, v, 1u << s);
          if (lane >= (1u << s)) v += pw_o[s] * t;
        }
        v += pw_lane * carry;
        if (act) { cY[j] = v; FIY[i * T + j] = v; }
        carry = __shfl_sync(0xffffffffu, v, 31);
        __syncwarp();
      }
      double* t;
      t = pM; pM = cM; cM = t;  t = pX; pX = cX; cX = t;  t = pY; pY = cY; cY = t;  t = pR; pR = cR; cR = t;
    }
    // value (:240-241): 1 + M + RX + RY at (Lx, Ly); RY(Lx,Ly) = sum_{j<Ly} (M(Lx,j) + RX(Lx,j))
    double res;
    {
      double ry = 0.0;
      if (lx > 0) for (uint32_t j = 1u + lane; j < ly; j += 32u) ry += pM[j] + pR[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ry += __shfl_xor_sync(0xffffffffu, ry, o);
      res = (lx > 0 && ly > 0) ? ((1.0 + pM[ly]) + pR[ly]) + ry : ((lx == 0 && ly == 0) ? 2.0 : 1.0);
    }
    __syncwarp();
    // ---- backward (:245-314) fused with the derivative sums (:335-385)
    double *nBM = rows, *nBX = rows + W, *cBM = rows + 4 * W, *cBX = rows + 5 * W;
    for (uint32_t j = lane; j <= ly + 1u; j += 32u) { nBM[j] = 0.0; nBX[j] = 0.0; cBM[j] = 0.0; cBX[j] = 0.0; }
    __syncwarp();
    double da = 0.0, db = 0.0, dg = 0.0, de = 0.0;
    for (uint32_t i = lx; i >= 1u; --i) {
      float xc[4];
      const uint32_t xcode = __ldg(P.X.code + x0 + i - 1u);
#pragma unroll
      for (int a = 0; a < 4; ++a) xc[a] = __ldg(P.X.profile + 5u * (x0 + i - 1u) + a);
      const float xpl = __ldg(P.X.pl + x0 + i - 1u), xpr = __ldg(P.X.pr + x0 + i - 1u), xpu = __ldg(P.X.pu + x0 + i - 1u);
      double carry = 0.0;                         // BIY(i, ly + 1)
      for (int jr = (int)ly; jr >= 1; jr -= 32) {
        const int js = jr - (int)lane;            // lane 0 is the rightmost column of the chunk
        const bool act = js >= 1;
        const uint32_t j = act ? (uint32_t)js : 1u;
        double D = 0.0;
        if (act && i < lx && j < ly) D = BS[(size_t)(i + 1u) * T + j + 1u] * nBM[j + 1u];
        double v = act ? D : 0.0;
#pragma unroll
        for (int s = 0; s < 5; ++s) {
          const double t = __shfl_up_sync(0xffffffffu, v, 1u << s);
          if (lane >= (1u << s)) v += pw_o[s] * t;
        }
        v += pw_lane * carry;                     // BIY(i,j)
        double right = __shfl_up_sync(0xffffffffu, v, 1);   // BIY(i,j+1)
        if (lane == 0) right = carry;
        carry = __shfl_sync(0xffffffffu, v, 31);
        if (act) {
          const double nbx = nBX[j];
          const double bix = (D + be * nbx) + bg * right;
          const double bm = (D + bg * (nbx + right)) + 1.0;
          cBM[j] = bm; cBX[j] = bix;
          double wp, wu;
          cell(xcode, xc, xpl, xpr, xpu, y0 + j - 1u, &wp, &wu);
          const double bs = BS[(size_t)i * T + j];
          const size_t d1 = (size_t)(i - 1u) * T + (j - 1u), up = (size_t)(i - 1u) * T + j, lf = (size_t)i * T + (j - 1u);
          const double st = (i == 1u && j == 1u) ? 2.0 : ((i == 1u || j == 1u) ? 1.0 : (double)j);
          const double vs = ((((FM[d1] + FIX[d1]) + FIY[d1]) + st) * bs) * bm;          // :361-370
          da += (P.beta * wp) * vs;
          db += (P.alpha * wp + wu) * vs;
          const double vg = (FM[up] * bg) * bix + ((FM[lf] + FIX[lf]) * bg) * v;        // :372-373, :377-380
          const double ve = (FIX[up] * be) * bix + (FIY[lf] * be) * v;                  // :374-375, :381-382
          db += P.gap * vg + P.ext * ve;
          dg += P.beta * vg;
          de += P.beta * ve;
        }
      }
      __syncwarp();
      double* t;
      t = nBM; nBM = cBM; cBM = t;  t = nBX; nBX = cBX; cBX = t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      da += __shfl_xor_sync(0xffffffffu, da, o); db += __shfl_xor_sync(0xffffffffu, db, o);
      dg += __shfl_xor_sync(0xffffffffu, dg, o); de += __shfl_xor_sync(0xffffffffu, de, o);
    }
    if (lane == 0) {
      P.out[k] = res;
      G.grad[4 * k] = da; G.grad[4 * k + 1] = db; G.grad[4 * k + 2] = dg; G.grad[4 * k + 3] = de;
    }
    __syncwarp();
  }
}

}  // namespace

// Host side: copies the two sets to the device, runs the kernel, copies the values back.  Everything on `stream`.
cudaError_t run_bpla(const stemk_bpla_params& p, const stemk_bpla_set& x, const stemk_bpla_set& y, size_t n_pairs,
                     const uint32_t* xi, const uint32_t* yi, double* out, double* grad, int sm_count, size_t smem_optin,
                     cudaStream_t stream, std::string* err) {
  struct Up { void* d = nullptr; };
  std::vector<void*> to_free;
  auto up = [&](const void* h, size_t bytes, const void** d) -> cudaError_t {
    *d = nullptr;
    if (!h || bytes == 0) return cudaSuccess;
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess) return e;
    to_free.push_back(q);
    *d = q;
    return cudaMemcpyAsync(q, h, bytes, cudaMemcpyHostToDevice, stream);
  };
  auto cleanup = [&]() { for (void* q : to_free) cudaFree(q); };
  BplaLaunch L;
  cudaError_t e = cudaSuccess;
  uint32_t ly_cap = 1;
  std::vector<std::vector<uint8_t>> codes;   // host staging of the column codes, alive until the copies are done
  codes.reserve(2);
  auto up_set = [&](const stemk_bpla_set& s, BplaDev* d, bool is_y) -> cudaError_t {
    const size_t ncol = s.n_seqs ? s.col_off[s.n_seqs] : 0;
    if (is_y) for (uint32_t r = 0; r < s.n_seqs; ++r) ly_cap = std::max(ly_cap, s.col_off[r + 1] - s.col_off[r]);
    cudaError_t q;
    if ((q = up(s.col_off, sizeof(uint32_t) * (s.n_seqs + 1), (const void**)&d->col_off)) != cudaSuccess) return q;
    if ((q = up(s.profile, sizeof(float) * 5 * ncol, (const void**)&d->profile)) != cudaSuccess) return q;
    codes.emplace_back(ncol);
    std::vector<uint8_t>& cd = codes.back();
    for (size_t c = 0; c < ncol; ++c) {
      const float* pc = s.profile + 5 * c;
      int ones = 0, zeros = 0, which = 0;
      for (int a = 0; a < 4; ++a) { if (pc[a] == 1.0f) { ++ones; which = a; } if (pc[a] == 0.0f) ++zeros; }
      cd[c] = (ones == 1 && zeros == 3) ? (uint8_t)which : (zeros == 4 ? 4 : 5);
    }
    if ((q = up(cd.data(), ncol, (const void**)&d->code)) != cudaSuccess) return q;
    if ((q = up(s.p_left, sizeof(float) * ncol, (const void**)&d->pl)) != cudaSuccess) return q;
    if ((q = up(s.p_right, sizeof(float) * ncol, (const void**)&d->pr)) != cudaSuccess) return q;
    return up(s.p_unpair, sizeof(float) * ncol, (const void**)&d->pu);
  };
  if ((e = up_set(x, &L.X, false)) != cudaSuccess || (e = up_set(y, &L.Y, true)) != cudaSuccess) { cleanup(); return e; }
  const void *dxi, *dyi;
  void* dout = nullptr;
  unsigned long long* dcnt = nullptr;
  if ((e = up(xi, sizeof(uint32_t) * n_pairs, &dxi)) != cudaSuccess || (e = up(yi, sizeof(uint32_t) * n_pairs, &dyi)) != cudaSuccess ||
      (e = cudaMalloc(&dout, sizeof(double) * n_pairs)) != cudaSuccess) { cleanup(); return e; }
  to_free.push_back(dout);
  if ((e = cudaMalloc((void**)&dcnt, sizeof(unsigned long long))) != cudaSuccess) { cleanup(); return e; }
  to_free.push_back(dcnt);
  cudaMemsetAsync(dcnt, 0, sizeof(unsigned long long), stream);
  L.xi = (const uint32_t*)dxi; L.yi = (const uint32_t*)dyi; L.n_pairs = n_pairs; L.counter = dcnt; L.out = (double*)dout;
  L.gap = p.gap; L.ext = p.ext; L.alpha = p.alpha; L.beta = p.beta; L.no_bp = p.no_bp; L.sw = p.sw; L.ly_cap = ly_cap;
  for (int k = 0; k < 16; ++k) L.score[k] = p.score[k];
  if (grad) {   // compute_gradients: forward tables in a per-warp scratch, values + four derivatives per pair
    uint32_t lx_cap = 1;
    for (uint32_t r = 0; r < x.n_seqs; ++r) lx_cap = std::max(lx_cap, x.col_off[r + 1] - x.col_off[r]);
    const size_t gsmem = sizeof(double) * 8 * (ly_cap + 2) * kBplaWarps;
    if (gsmem > smem_optin) { cleanup(); if (err) *err = "BPLA gradient kernel: sequence too long for shared memory"; return cudaErrorInvalidValue; }
    if ((e = cudaFuncSetAttribute(bpla_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem)) != cudaSuccess) { cleanup(); return e; }
    const unsigned long long stride = 4ull * (lx_cap + 1ull) * (ly_cap + 1ull);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bpla_grad_kernel, 32 * kBplaWarps, gsmem);
    per_sm = std::max(1, std::min(per_sm, 8));
    size_t ctas = std::min<size_t>((n_pairs + kBplaWarps - 1) / kBplaWarps, (size_t)sm_count * per_sm);
    const size_t budget = (size_t)4 << 30;   // scratch of all warps in flight
    ctas = std::max<size_t>(1, std::min<size_t>(ctas, (size_t)(budget / (sizeof(double) * stride * kBplaWarps))));
    void *dscr = nullptr, *dgrad = nullptr;
    if ((e = cudaMalloc(&dscr, sizeof(double) * stride * kBplaWarps * ctas)) != cudaSuccess) { cleanup(); return e; }
    to_free.push_back(dscr);
    if ((e = cudaMalloc(&dgrad, sizeof(double) * 4 * n_pairs)) != cudaSuccess) { cleanup(); return e; }
    to_free.push_back(dgrad);
    BplaGradLaunch GL;
    GL.B = L; GL.scratch = (double*)dscr; GL.scratch_stride = stride; GL.grad = (double*)dgrad;
    bpla_grad_kernel<<<(int)ctas, 32 * kBplaWarps, gsmem, stream>>>(GL);
    if ((e = cudaGetLastError()) != cudaSuccess) { cleanup(); return e; }
    e = cudaMemcpyAsync(out, dout, sizeof(double) * n_pairs, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(grad, dgrad, sizeof(double) * 4 * n_pairs, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cleanup();
    return e;
  }
  const size_t smem = sizeof(double) * 8 * (ly_cap + 1) * kBplaWarps;
  if (smem > smem_optin) { cleanup(); if (err) *err = "BPLA kernel: sequence too long for shared memory"; return cudaErrorInvalidValue; }
  void (*kern)(const BplaLaunch) = p.sw ? (p.no_bp ? bpla_pairs_kernel<true, true> : bpla_pairs_kernel<true, false>)
                                        : (p.no_bp ? bpla_pairs_kernel<false, true> : bpla_pairs_kernel<false, false>);
  if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) { cleanup(); return e; }
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * kBplaWarps, smem);
  per_sm = std::max(1, std::min(per_sm, 8));
  const int grid = (int)std::min<size_t>((n_pairs + kBplaWarps - 1) / kBplaWarps, (size_t)sm_count * per_sm);
  kern<<<grid, 32 * kBplaWarps, smem, stream>>>(L);
  if ((e = cudaGetLastError()) != cudaSuccess) { cleanup(); return e; }
  e = cudaMemcpyAsync(out, dout, sizeof(double) * n_pairs, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cleanup();
  return e;
}

}  // namespace stemk

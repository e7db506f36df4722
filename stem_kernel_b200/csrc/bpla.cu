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

}  // namespace

// Host side: copies the two sets to the device, runs the kernel, copies the values back.  Everything on `stream`.
cudaError_t run_bpla(const stemk_bpla_params& p, const stemk_bpla_set& x, const stemk_bpla_set& y, size_t n_pairs,
                     const uint32_t* xi, const uint32_t* yi, double* out, int sm_count, size_t smem_optin,
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

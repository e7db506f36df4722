// stem_kernel_b200/csrc/nstem.cu -- the naive stem kernel.
//
// StemKernel<double, BPMat>::full_dp of the stem_kernel/ program (stem_kernel/stem_kernel.cpp:282-351, dp_init /
// dp_update :85-111): eight tables over (i,j) x (k,l), all pairs of base pairs of two raw sequences.  Restated:
//   * the four K tables are plain prefix sums of the MATCH increments in all four directions, so the result
//     K0(0,Lx,0,Ly) is 1 + the sum of every increment: no K table;
//   * per plane (i,j):  G3(k,l) = g G3(k+1,l) + [pair(i,j-1), pair(k,l-1), same bases] G0(i+1,j-1,k+1,l-1)
//                       G2(k,l) = g G2(k,l-1) + G3(k,l)
//                       G1(i,j) = g G1(i+1,j) + G2            G0(i,j) = g G0(i,j-1) + G1(i,j)
//     with G0(j,j,k,l) = g^(l-k); G3 is a suffix recurrence along k (thread <-> column l, sequential in k),
//     G2 a prefix recurrence along l (warp <-> row k, inclusive warp scan with powers of g, carry between chunks).
// One CTA per pair, x = the longer sequence (the kernel is symmetric in its arguments up to rounding).  The G0
// planes of the current column live in global memory (one (Ly+1)^2 plane per i, updated in place column by column;
// the plane of i+1 from the previous column, which the MATCH term of i needs, is kept in one of two spare planes),
// G3 and G1 of the running plane in shared memory.  Planes whose (i,j-1) is not a base pair (almost all) skip the
// G3 pass.  Only practical for short sequences, like the reference: the work is Lx^2 Ly^2 / 4 cells.
#include <algorithm>

#include "kernels.cuh"

namespace stemk {

namespace {

constexpr int kNstemThreads = 1024;   // 32 warps: a warp streams only ~2-4 B/clk, the G0 planes need many of them

struct NstemDev {
  const uint32_t* off;
  const char* text;
  const unsigned long long* bp_off;
  const float* bp;
};

struct NstemLaunch {
  NstemDev X, Y;
  const uint32_t* xi;
  const uint32_t* yi;
  unsigned long long n_pairs;
  unsigned long long* counter;
  double* out;
  double* scratch;                  // per CTA: (lx_cap + 3) planes of pitch^2 doubles
  unsigned long long scratch_stride;
  double gap, stack, subst;
  float bp_bound;
  int bp_mode, use_gu;
  uint32_t loop, pitch;             // pitch = ly_cap + 1 (odd)
};

__device__ __forceinline__ float pair_prob(const NstemLaunch& P, const char* s, uint32_t len, const float* tab, uint32_t i, uint32_t j) {
  if (P.bp_mode == 1) return __ldg(tab + (size_t)i * len + j);
  const char a = s[i], b = s[j];
  bool ok = (a == 'a' && b == 'u') || (a == 'u' && b == 'a') || (a == 'g' && b == 'c') || (a == 'c' && b == 'g');
  if (P.use_gu) ok = ok || (a == 'g' && b == 'u') || (a == 'u' && b == 'g');
  return (i + 1u + P.loop <= j && ok) ? 1.0f : 0.0f;
}

__global__ void __launch_bounds__(kNstemThreads) nstem_pairs_kernel(const NstemLaunch P) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double s_red[kNstemThreads / 32];
  __shared__ unsigned long long s_k;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, nwarps = kNstemThreads / 32;
  const uint32_t PT = P.pitch;
  const size_t plane = (size_t)PT * PT;
  double* P3 = sm;                 // G3 of the running plane
  double* G1 = sm + plane;         // G1(i+1, j) -> G1(i, j)
  double* base = P.scratch + (size_t)blockIdx.x * P.scratch_stride;
  double* T0 = base;               // spare planes: G0(i+1, j-1) for the MATCH term
  double* T1 = base + plane;
  double* G0 = base + 2 * plane;   // G0[i]: plane of (i, current or previous column)
  const double g = P.gap;
  double pw_lane = 1.0, pw_o[5];   // g^(lane+1), g^(2^s)
  for (uint32_t k = 0; k <= lane; ++k) pw_lane *= g;
  { double t = g; for (int s = 0; s < 5; ++s) { pw_o[s] = t; t *= t; } }

  for (;;) {
    __syncthreads();
    if (tid == 0) s_k = atomicAdd(P.counter, 1ull);
    __syncthreads();
    const unsigned long long pk = s_k;
    if (pk >= P.n_pairs) break;
    // x = the longer sequence
    uint32_t xr = P.xi[pk], yr = P.yi[pk];
    const NstemDev* DX = &P.X;
    const NstemDev* DY = &P.Y;
    {
      const uint32_t la = P.X.off[xr + 1] - P.X.off[xr], lb = P.Y.off[yr + 1] - P.Y.off[yr];
      if (lb > la) { const NstemDev* t = DX; DX = DY; DY = t; const uint32_t u = xr; xr = yr; yr = u; }
    }
    const char* x = DX->text + DX->off[xr];
    const char* y = DY->text + DY->off[yr];
    const uint32_t lx = DX->off[xr + 1] - DX->off[xr], ly = DY->off[yr + 1] - DY->off[yr];
    const float* tx = P.bp_mode == 1 ? DX->bp + DX->bp_off[xr] : nullptr;
    const float* ty = P.bp_mode == 1 ? DY->bp + DY->bp_off[yr] : nullptr;
    double ksum = 0.0;

    // start plane (j,j): G0 = g^(l-k) on and above the diagonal (powers by repeated multiplication, :297-303)
    auto start_plane = [&](uint32_t j) {
      for (uint32_t k = tid; k <= ly; k += kNstemThreads) {
        double v = 1.0;
        double* row = G0 + (size_t)j * plane + (size_t)k * PT;
        for (uint32_t l = 0; l < k; ++l) row[l] = 0.0;
        for (uint32_t l = k; l <= ly; ++l) { row[l] = v; v *= g; }
      }
    };
    start_plane(0);
    double* Told = T0;
    double* Tnew = T1;
    for (uint32_t j = 1; j <= lx; ++j) {
      start_plane(j);                                     // used as G0(j, j) by column j+1
      for (uint32_t c = tid; c < (ly + 1u) * PT; c += kNstemThreads) G1[c] = 0.0;   // G1(j, j) = 0
      __syncthreads();
      for (uint32_t i = j; i-- > 0;) {
        const float bp_ij = (i + 1u < j) ? pair_prob(P, x, lx, tx, i, j - 1u) : 0.0f;
        const bool active = bp_ij > P.bp_bound;
        // the old plane of i is only needed by the MATCH pass of i-1: keep it when that pass will run
        const bool keep_old = i > 0u && pair_prob(P, x, lx, tx, i - 1u, j - 1u) > P.bp_bound;
        if (active) {
          // ---- G3 (suffix recurrence along k) and the MATCH increments; thread <-> column l
          const char xa = x[i], xb = x[j - 1u];
          for (uint32_t l = 1u + tid; l <= ly; l += kNstemThreads) {
            double g3 = 0.0;
            const char yb = y[l - 1u];
            for (uint32_t k = l; k-- > 0;) {
              double mg = 0.0;
              if (k + 1u < l) {
                const float bp_kl = pair_prob(P, y, ly, ty, k, l - 1u);
                if (bp_kl > P.bp_bound) {
                  const double t = Told[(size_t)(k + 1u) * PT + (l - 1u)];
                  const bool same = xa == y[k] && xb == yb;
                  ksum += same ? ((t * P.stack) * (double)bp_ij) * (double)bp_kl
                               : (((t * P.stack) * P.subst) * (double)bp_ij) * (double)bp_kl;     // stem_kernel.cpp:323-330
                  if (same) mg = t;
                }
              }
              g3 = g3 * g + mg;
              P3[(size_t)k * PT + l] = g3;
            }
          }
          __syncthreads();
        }
        // ---- G2 (prefix recurrence along l), G1, G0; warp <-> row k, lanes <-> l
        double* g0p = G0 + (size_t)i * plane;
        for (uint32_t k = warp; k <= ly; k += nwarps) {
          double carry = 0.0;
          for (uint32_t l0 = k; l0 <= ly; l0 += 32u) {
            const uint32_t l = l0 + lane;
            const bool act = l <= ly;
            double v = (active && act && l > k) ? P3[(size_t)k * PT + l] : 0.0;
            if (active) {
#pragma unroll
              for (int s = 0; s < 5; ++s) {
                const double t = __shfl_up_sync(0xffffffffu, v, 1u << s);
                if (lane >= (1u << s)) v += pw_o[s] * t;
              }
              v += pw_lane * carry;
              carry = __shfl_sync(0xffffffffu, v, 31);
            }
            if (act) {
              const size_t at = (size_t)k * PT + l;
              const double g1 = G1[at] * g + v;
              G1[at] = g1;
              const double old = g0p[at];
              if (keep_old) Tnew[at] = old;
              g0p[at] = old * g + g1;
            }
          }
        }
        __syncthreads();
        double* t = Told; Told = Tnew; Tnew = t;
      }
    }
    // ---- result: 1 + every MATCH increment
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ksum += __shfl_xor_sync(0xffffffffu, ksum, o);
    if (lane == 0) s_red[warp] = ksum;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (uint32_t w = 0; w < nwarps; ++w) t += s_red[w];
      P.out[pk] = 1.0 + t;
    }
  }
}

}  // namespace

cudaError_t run_nstem(const stemk_nstem_params& p, const stemk_nstem_set& x, const stemk_nstem_set& y, size_t n_pairs,
                      const uint32_t* xi, const uint32_t* yi, double* out, int sm_count, size_t smem_optin,
                      cudaStream_t stream, std::string* err) {
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
  // x of a pair = the longer sequence, y = the shorter: caps over the pair list
  uint32_t lx_cap = 1, ly_cap = 1;
  for (size_t k = 0; k < n_pairs; ++k) {
    const uint32_t a = x.off[xi[k] + 1] - x.off[xi[k]], b = y.off[yi[k] + 1] - y.off[yi[k]];
    lx_cap = std::max(lx_cap, std::max(a, b));
    ly_cap = std::max(ly_cap, std::min(a, b));
  }
  const uint32_t pitch = (ly_cap + 1u) | 1u;
  const size_t plane = (size_t)pitch * pitch;
  const size_t smem = 2 * plane * sizeof(double);
  if (smem > smem_optin) { if (err) *err = "naive stem kernel: the shorter sequence of a pair is too long for shared memory"; return cudaErrorInvalidValue; }
  NstemLaunch L;
  cudaError_t e;
  auto up_set = [&](const stemk_nstem_set& s, NstemDev* d) -> cudaError_t {
    const size_t nch = s.n_seqs ? s.off[s.n_seqs] : 0;
    cudaError_t q;
    if ((q = up(s.off, sizeof(uint32_t) * (s.n_seqs + 1), (const void**)&d->off)) != cudaSuccess) return q;
    if ((q = up(s.text, nch, (const void**)&d->text)) != cudaSuccess) return q;
    d->bp_off = nullptr; d->bp = nullptr;
    if (p.bp_mode == 1) {
      size_t nbp = 0;
      for (uint32_t r = 0; r < s.n_seqs; ++r) { const size_t l = s.off[r + 1] - s.off[r]; nbp = std::max(nbp, (size_t)s.bp_off[r] + l * l); }
      if ((q = up(s.bp_off, sizeof(uint64_t) * s.n_seqs, (const void**)&d->bp_off)) != cudaSuccess) return q;
      if ((q = up(s.bp, sizeof(float) * nbp, (const void**)&d->bp)) != cudaSuccess) return q;
    }
    return cudaSuccess;
  };
  if ((e = up_set(x, &L.X)) != cudaSuccess || (e = up_set(y, &L.Y)) != cudaSuccess) { cleanup(); return e; }
  const void *dxi, *dyi;
  void *dout = nullptr, *dscr = nullptr;
  unsigned long long* dcnt = nullptr;
  const int grid = (int)std::min<size_t>(n_pairs, (size_t)sm_count);
  const unsigned long long stride = (unsigned long long)(lx_cap + 3u) * plane;
  if ((e = up(xi, sizeof(uint32_t) * n_pairs, &dxi)) != cudaSuccess || (e = up(yi, sizeof(uint32_t) * n_pairs, &dyi)) != cudaSuccess ||
      (e = cudaMalloc(&dout, sizeof(double) * n_pairs)) != cudaSuccess) { cleanup(); return e; }
  to_free.push_back(dout);
  if ((e = cudaMalloc(&dscr, sizeof(double) * stride * grid)) != cudaSuccess) { cleanup(); return e; }
  to_free.push_back(dscr);
  if ((e = cudaMalloc((void**)&dcnt, sizeof(unsigned long long))) != cudaSuccess) { cleanup(); return e; }
  to_free.push_back(dcnt);
  cudaMemsetAsync(dcnt, 0, sizeof(unsigned long long), stream);
  L.xi = (const uint32_t*)dxi; L.yi = (const uint32_t*)dyi; L.n_pairs = n_pairs; L.counter = dcnt; L.out = (double*)dout;
  L.scratch = (double*)dscr; L.scratch_stride = stride;
  L.gap = p.gap; L.stack = p.stack; L.subst = p.subst; L.bp_bound = p.bp_bound; L.bp_mode = p.bp_mode; L.use_gu = p.use_gu;
  L.loop = p.loop; L.pitch = pitch;
  if ((e = cudaFuncSetAttribute(nstem_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) { cleanup(); return e; }
  nstem_pairs_kernel<<<grid, kNstemThreads, smem, stream>>>(L);
  if ((e = cudaGetLastError()) != cudaSuccess) { cleanup(); return e; }
  e = cudaMemcpyAsync(out, dout, sizeof(double) * n_pairs, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cleanup();
  return e;
}

}  // namespace stemk

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


// ---- StemKernel::partial_dp with the band-only constraints (stem_kernel.cpp:14-83, 113-280) ----------------------
// Row i of x may pair with the columns c_low[i] .. c_high[i] of y (a band of half-width `band` around the diagonal,
// :77-83), so a plane (i,j) only has cells k in [c_low[i], c_high[i]], l in [c_low[j], c_high[j]], k < l; everything
// else keeps the plane's fill value.  At the window borders the reference substitutes "approximation" cells for the
// missing neighbours, which breaks the prefix-sum structure the full kernel uses to drop the K tables: all eight
// tables are kept, on band-compressed planes (window of (2 band + 2)^2 cells per table) in a per-warp global scratch
// holding two columns j of planes.  Restated facts the kernel uses (each follows from the statements of :130-262):
//   * the planes (j,j) are constants: K0 = 1, G0(k,l) = g^(l-k), the other six tables 0;
//   * the diagonal cells of a plane (i,j) are K0(l,l) = 1, G0(l,l) = g^(j-i) for l in the window of j, else 0, and 0
//     in the other six tables, so the two fall-back sums over K3/G3(ll,ll) (:203-209, :215-216) are 0;
//   * a cell outside its plane's window reads as 0.
// One warp per pair: pass 1, lane <-> column l, the K3/G3 suffix recurrences along k with the MATCH term; pass 2,
// lane <-> row k, K2/G2 along l, then K1/G1 from the plane (i+1,j) and K0/G0 from the plane (i,j-1), statement order
// and operand order of the reference kept.
struct NstemBandLaunch {
  NstemLaunch B;
  uint32_t band, wcap;              // wcap = plane pitch: min(2 band + 2, ly_cap + 2), or the widest caller window + 2
  // caller-supplied constraints instead of the band (alignment_constraints with ali_bound > 0, stem_kernel.cpp:25-67):
  // pair k owns entries win_off[k] .. win_off[k] + lx of c_low / c_high (row i of x pairs with columns c_low..c_high of y)
  const uint32_t* win_off;
  const uint32_t* c_low;
  const uint32_t* c_high;
};
constexpr int kNbandWarps = 4;
enum { BK0 = 0, BK1, BK2, BK3, BG0, BG1, BG2, BG3 };

__global__ void __launch_bounds__(32 * kNbandWarps) nstem_banded_kernel(const NstemBandLaunch Q) {
  const NstemLaunch& P = Q.B;
  extern __shared__ __align__(16) double sm[];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t WP = Q.wcap;                                   // pitch of a compressed plane
  const size_t tsz = (size_t)WP * WP;                            // one table of one plane
  double* scr = P.scratch + (size_t)(blockIdx.x * kNbandWarps + warp) * P.scratch_stride;
  const double g = P.gap;
  const uint32_t band = Q.band;

  for (;;) {
    unsigned long long kq = 0;
    if (lane == 0) kq = atomicAdd(P.counter, 1ull);
    kq = __shfl_sync(0xffffffffu, kq, 0);
    if (kq >= P.n_pairs) break;
    const uint32_t xr = P.xi[kq], yr = P.yi[kq];
    const uint32_t lx = P.X.off[xr + 1] - P.X.off[xr], ly = P.Y.off[yr + 1] - P.Y.off[yr];
    const char* x = P.X.text + P.X.off[xr];
    const char* y = P.Y.text + P.Y.off[yr];
    const float* tx = P.bp_mode == 1 ? P.X.bp + P.X.bp_off[xr] : nullptr;
    const float* ty = P.bp_mode == 1 ? P.Y.bp + P.Y.bp_off[yr] : nullptr;
    if (lx == 0 || ly == 0) {                                    // K0(0,lx,0,ly) is a diagonal cell or a (j,j) plane: 1
      if (lane == 0) P.out[kq] = 1.0;
      continue;
    }
    // g^n by repeated multiplication (the chains G0(i,j,l,l) = G0(i+1,j,l,l) * g and :119-121), per warp in shared memory
    double* gp = sm + (size_t)warp * (P.pitch + 2u);
    __syncwarp();
    if (lane == 0) { gp[0] = 1.0; for (uint32_t n = 1; n <= max(lx, ly); ++n) gp[n] = gp[n - 1] * g; }
    __syncwarp();
    const uint32_t* wl = Q.win_off ? Q.c_low + Q.win_off[kq] : nullptr;
    const uint32_t* wh = Q.win_off ? Q.c_high + Q.win_off[kq] : nullptr;
    auto clo = [&](uint32_t i) { if (wl) return __ldg(wl + i); const uint32_t c = (uint32_t)((double)i / (double)lx * (double)ly + 0.5); return c < band ? 0u : c - band; };
    auto chi = [&](uint32_t i) { if (wh) return __ldg(wh + i); const uint32_t c = (uint32_t)((double)i / (double)lx * (double)ly + 0.5); return c + band > ly ? ly : c + band; };
    // plane (i, column parity): 8 tables of WP x WP, cell (k,l) at [k - clo(i)][l - clo(j)]
    auto tab = [&](uint32_t t, uint32_t i, uint32_t j) { return scr + (((size_t)(j & 1u) * (lx + 1u) + i) * 8u + t) * tsz; };
    // value of table t of plane (i,j) at (k,l), for a plane that is complete (fill values, diagonals, (j,j) planes)
    auto get = [&](uint32_t t, uint32_t i, uint32_t j, uint32_t k, uint32_t l) -> double {
      if (i == j) return t == BK0 ? 1.0 : ((t == BG0 && k <= l) ? gp[l - k] : 0.0);
      const uint32_t ci = clo(i), cj = clo(j);
      if (l < cj || l > chi(j)) return 0.0;
      if (k == l) return t == BK0 ? 1.0 : (t == BG0 ? gp[j - i] : 0.0);
      if (k < ci || k > chi(i) || k > l) return 0.0;
      return tab(t, i, j)[(size_t)(k - ci) * WP + (l - cj)];
    };

    for (uint32_t j = 1; j <= lx; ++j) {
      const uint32_t cj = clo(j), hj = chi(j), hj1 = chi(j - 1u);
      for (uint32_t i = j - 1u;; --i) {
        const uint32_t ci = clo(i), hi = chi(i), ci1 = clo(i + 1u);
        const float bp_ij = pair_prob(P, x, lx, tx, i, j - 1u);
        double* T[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) T[t] = tab(t, i, j);
        for (size_t c = lane; c < 8u * tsz; c += 32u) T[0][c] = 0.0;            // the 8 tables of a plane are contiguous
        __syncwarp();
        // ---- pass 1: K3 / G3 along k (descending), lane <-> l
        for (uint32_t l = cj + lane; l <= hj; l += 32u) {
          if (l == 0u) continue;
          uint32_t k = min(l - 1u, hi);
          if (k < ci) continue;
          double k3 = 0.0, g3 = 0.0;                                               // K3 / G3 (k+1, l): 0 above the window and on the diagonal
          for (;; --k) {
            if (k + 1u <= hi) g3 = g3 * g;                                         // :211-213 (K3 is copied)
            else { k3 = 0.0; g3 = 0.0; }                                           // :214-217: K3/G3(l,l) = 0
            if (bp_ij > P.bp_bound) {                                              // :219-231
              const float bp_kl = pair_prob(P, y, ly, ty, k, l - 1u);
              if (bp_kl > P.bp_bound) {
                const double g0m = get(BG0, i + 1u, j - 1u, k + 1u, l - 1u);
                if (x[i] == y[k] && x[j - 1u] == y[l - 1u]) {
                  k3 += g0m * P.stack * (double)bp_ij * (double)bp_kl;
                  g3 += g0m;
                } else {
                  k3 += g0m * P.stack * P.subst * (double)bp_ij * (double)bp_kl;
                }
              }
            }
            T[BK3][(size_t)(k - ci) * WP + (l - cj)] = k3;
            T[BG3][(size_t)(k - ci) * WP + (l - cj)] = g3;
            if (k == ci) break;
          }
        }
        __syncwarp();
        // ---- pass 2: K2 / G2 along l (ascending), then K1 / G1 and K0 / G0, lane <-> k
        for (uint32_t k = ci + lane; k <= hi; k += 32u) {
          double k2 = 0.0, g2 = 0.0;                                               // K2 / G2 (k, l-1)
          for (uint32_t l = max(cj, k + 1u); l <= hj; ++l) {
            // :198-210: the left neighbour inside the window (0 on the diagonal), nothing left of the window
            if (l - 1u >= cj) g2 = g2 * g; else { k2 = 0.0; g2 = 0.0; }
            const size_t at = (size_t)(k - ci) * WP + (l - cj);
            k2 += T[BK3][at];                                                      // dp_update :102-103
            g2 += T[BG3][at];
            double k1, g1;                                                         // :179-186
            if (k >= ci1) { k1 = get(BK1, i + 1u, j, k, l); g1 = get(BG1, i + 1u, j, k, l) * g; }
            else { k1 = get(BK1, i + 1u, j, ci1, l); g1 = get(BG1, i + 1u, j, ci1, l) * g * g; }
            k1 += k2;                                                              // :104-105
            g1 += g2;
            double k0, g0;                                                         // :170-177
            if (l <= hj1) { k0 = get(BK0, i, j - 1u, k, l); g0 = get(BG0, i, j - 1u, k, l) * g; }
            else { k0 = get(BK0, i, j - 1u, k, hj1); g0 = get(BG0, i, j - 1u, k, hj1) * g * g; }
            k0 += k1;                                                              // :106-107
            g0 += g1;
            T[BK2][at] = k2; T[BG2][at] = g2; T[BK1][at] = k1; T[BG1][at] = g1; T[BK0][at] = k0; T[BG0][at] = g0;
          }
        }
        __syncwarp();
        if (i == 0u) break;
      }
    }
    if (lane == 0) P.out[kq] = get(BK0, 0u, lx, 0u, ly);
    __syncwarp();
  }
}

}  // namespace

cudaError_t run_nstem(const stemk_nstem_params& p, const stemk_nstem_set& x, const stemk_nstem_set& y, size_t n_pairs,
                      const uint32_t* xi, const uint32_t* yi, double* out, uint32_t band, int sm_count, size_t smem_optin,
                      cudaStream_t stream, std::string* err, const uint32_t* win_off, const uint32_t* c_low, const uint32_t* c_high) {
  const bool windows = win_off != nullptr;
  if (windows) band = 1;   // the banded kernel, its windows read from the caller's arrays
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
  // x of a pair = the longer sequence, y = the shorter: caps over the pair list (banded: the arguments keep their roles)
  uint32_t lx_cap = 1, ly_cap = 1;
  for (size_t k = 0; k < n_pairs; ++k) {
    const uint32_t a = x.off[xi[k] + 1] - x.off[xi[k]], b = y.off[yi[k] + 1] - y.off[yi[k]];
    lx_cap = std::max(lx_cap, band ? a : std::max(a, b));
    ly_cap = std::max(ly_cap, band ? b : std::min(a, b));
  }
  const uint32_t pitch = (ly_cap + 1u) | 1u;
  const size_t plane = (size_t)pitch * pitch;
  const size_t smem = band ? 0 : 2 * plane * sizeof(double);
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
  uint32_t wcap = std::min(2u * band + 2u, ly_cap + 2u);
  const void *d_woff = nullptr, *d_clo = nullptr, *d_chi = nullptr;
  if (windows) {
    uint32_t widest = 0;
    const size_t n_win = win_off[n_pairs];
    for (size_t q = 0; q < n_win; ++q) widest = std::max(widest, c_high[q] >= c_low[q] ? c_high[q] - c_low[q] : 0u);
    wcap = std::min(widest + 2u, ly_cap + 2u);
    if ((e = up(win_off, sizeof(uint32_t) * (n_pairs + 1), &d_woff)) != cudaSuccess || (e = up(c_low, sizeof(uint32_t) * n_win, &d_clo)) != cudaSuccess ||
        (e = up(c_high, sizeof(uint32_t) * n_win, &d_chi)) != cudaSuccess) { cleanup(); return e; }
  }
  int grid = (int)std::min<size_t>(n_pairs, (size_t)sm_count);
  unsigned long long stride = (unsigned long long)(lx_cap + 3u) * plane;
  if (band) {   // per warp: two columns of (lx_cap + 1) planes, eight tables of wcap^2 cells each
    stride = 2ull * (lx_cap + 1u) * 8ull * wcap * wcap;
    const size_t budget = (size_t)8 << 30;
    size_t ctas = std::min<size_t>((n_pairs + kNbandWarps - 1) / kNbandWarps, (size_t)sm_count * 4);
    ctas = std::max<size_t>(1, std::min<size_t>(ctas, (size_t)(budget / (sizeof(double) * stride * kNbandWarps))));
    grid = (int)ctas;
  }
  if ((e = up(xi, sizeof(uint32_t) * n_pairs, &dxi)) != cudaSuccess || (e = up(yi, sizeof(uint32_t) * n_pairs, &dyi)) != cudaSuccess ||
      (e = cudaMalloc(&dout, sizeof(double) * n_pairs)) != cudaSuccess) { cleanup(); return e; }
  to_free.push_back(dout);
  if ((e = cudaMalloc(&dscr, sizeof(double) * stride * grid * (band ? kNbandWarps : 1))) != cudaSuccess) { cleanup(); return e; }
  to_free.push_back(dscr);
  if ((e = cudaMalloc((void**)&dcnt, sizeof(unsigned long long))) != cudaSuccess) { cleanup(); return e; }
  to_free.push_back(dcnt);
  cudaMemsetAsync(dcnt, 0, sizeof(unsigned long long), stream);
  L.xi = (const uint32_t*)dxi; L.yi = (const uint32_t*)dyi; L.n_pairs = n_pairs; L.counter = dcnt; L.out = (double*)dout;
  L.scratch = (double*)dscr; L.scratch_stride = stride;
  L.gap = p.gap; L.stack = p.stack; L.subst = p.subst; L.bp_bound = p.bp_bound; L.bp_mode = p.bp_mode; L.use_gu = p.use_gu;
  L.loop = p.loop; L.pitch = pitch;
  if (band) {
    NstemBandLaunch BL;
    BL.B = L; BL.band = band; BL.wcap = wcap;
    BL.win_off = (const uint32_t*)d_woff; BL.c_low = (const uint32_t*)d_clo; BL.c_high = (const uint32_t*)d_chi;
    BL.B.pitch = std::max(pitch, (lx_cap + 1u) | 1u);                    // the powers of g, per warp: up to max(lx, ly)
    const size_t bsmem = sizeof(double) * (BL.B.pitch + 2u) * kNbandWarps;
    nstem_banded_kernel<<<grid, 32 * kNbandWarps, bsmem, stream>>>(BL);
  } else {
    if ((e = cudaFuncSetAttribute(nstem_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) { cleanup(); return e; }
    nstem_pairs_kernel<<<grid, kNstemThreads, smem, stream>>>(L);
  }
  if ((e = cudaGetLastError()) != cudaSuccess) { cleanup(); return e; }
  e = cudaMemcpyAsync(out, dout, sizeof(double) * n_pairs, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cleanup();
  return e;
}

}  // namespace stemk

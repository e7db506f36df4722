// stem_kernel_b200/csrc/matrix_ops.cu -- the KernelMatrix bookkeeping around the pair evaluator, on the device:
// work-order pair lists (so that no O(n^2) list is built on the host or copied to the device), scatter of pair
// values into the square / rectangular matrix, normalisation (common/kernel_matrix.cpp:42-56, 560-571, 735-748).
// All of it is HBM-bound element-wise work, a fraction of a percent of a Gram matrix.
#include "kernels.cuh"

namespace stemk {

namespace {

// Square matrix, y-major work order: for every record b = perm[q] (biggest first) all partners a = perm[p] with
// a <= b, in perm order (biggest first).  The reference evaluates kernel(x_i, x_j) with i <= j
// (kernel_matrix.cpp:47-50), so x is the smaller ORIGINAL index.  Row q owns b + 1 pairs starting at off[q].
__global__ void gram_pairs_kernel(const uint32_t* __restrict__ perm, const unsigned long long* __restrict__ off,
                                  uint32_t n, uint32_t* __restrict__ xi, uint32_t* __restrict__ yi) {
  const uint32_t q = blockIdx.x;
  const uint32_t b = perm[q];
  __shared__ uint32_t wsum[32];
  __shared__ unsigned long long s_base;
  if (threadIdx.x == 0) s_base = off[q];
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (uint32_t p0 = 0; p0 < n; p0 += blockDim.x) {
    const uint32_t p = p0 + threadIdx.x;
    const uint32_t a = p < n ? perm[p] : 0xffffffffu;
    const bool mine = p < n && a <= b;
    const unsigned m = __ballot_sync(0xffffffffu, mine);
    if (lane == 0) wsum[w] = __popc(m);
    __syncthreads();
    uint32_t before = 0, total = 0;
    for (uint32_t k = 0; k < nw; ++k) { if (k < w) before += wsum[k]; total += wsum[k]; }
    if (mine) {
      const unsigned long long at = s_base + before + __popc(m & ((1u << lane) - 1u));
      xi[at] = a;
      yi[at] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base += total;
    __syncthreads();
  }
}

// Rectangular matrix: pair k = (train column cperm[k % nc], test row tperm[k / nc]); the train record is the FIRST
// kernel argument (kernel_matrix.cpp:159,168).
__global__ void cross_pairs_kernel(const uint32_t* __restrict__ tperm, const uint32_t* __restrict__ cperm, uint32_t nt,
                                   uint32_t nc, uint32_t* __restrict__ xi, uint32_t* __restrict__ yi) {
  const unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (k >= (unsigned long long)nt * nc) return;
  xi[k] = cperm[k % nc];
  yi[k] = tperm[k / nc];
}

// out[row(k) * ld + col(k)] = vals[k]: row = y index (test record) or its rank, col = x index (train column) or its rank
__global__ void scatter_cross_kernel(const double* __restrict__ vals, const uint32_t* __restrict__ xi,
                                     const uint32_t* __restrict__ yi, unsigned long long n_pairs,
                                     const uint32_t* __restrict__ col_of, double* __restrict__ out, uint32_t ld) {
  const unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (k >= n_pairs) return;
  const uint32_t c = col_of ? col_of[xi[k]] : xi[k];
  out[(size_t)yi[k] * ld + c] = vals[k];
}

// out[i * ld + c] /= sqrt(self[i] * diag[col[c]])   (kernel_matrix.cpp:735-748)
__global__ void normalize_cross_kernel(double* __restrict__ out, uint32_t nt, uint32_t nc, uint32_t ld,
                                       const double* __restrict__ selfv, const double* __restrict__ diag,
                                       const uint32_t* __restrict__ cols) {
  const unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (k >= (unsigned long long)nt * nc) return;
  const uint32_t i = (uint32_t)(k / nc), c = (uint32_t)(k % nc);
  out[(size_t)i * ld + c] /= sqrt(selfv[i] * diag[cols ? cols[c] : c]);
}

__global__ void scatter_vec_kernel(const double* __restrict__ vals, const uint32_t* __restrict__ idx, uint32_t n,
                                   double* __restrict__ out) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[idx[k]] = vals[k];
}

// Multi-device Gram matrix: device r of W owns positions r, r + W, ... of the global pair order
__global__ void deal_pairs_kernel(const uint32_t* __restrict__ xi, const uint32_t* __restrict__ yi, unsigned long long n_pairs,
                                  uint32_t rank, uint32_t world, uint32_t* __restrict__ out_x, uint32_t* __restrict__ out_y) {
  const unsigned long long m = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  const unsigned long long k = m * world + rank;
  if (k >= n_pairs) return;
  out_x[m] = xi[k];
  out_y[m] = yi[k];
}

// ... and the gathered values [world][per] go back into the global order: position k sits at [k % world][k / world]
__global__ void undeal_kernel(const double* __restrict__ gathered, unsigned long long per, uint32_t world,
                              unsigned long long n_pairs, double* __restrict__ vals) {
  const unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (k >= n_pairs) return;
  vals[k] = gathered[(k % world) * per + k / world];
}

__global__ void iota_kernel(uint32_t* __restrict__ a, uint32_t n) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) a[k] = k;
}

}  // namespace

cudaError_t launch_gram_pairs(const uint32_t* perm, const unsigned long long* off, uint32_t n, uint32_t* xi, uint32_t* yi,
                              cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  gram_pairs_kernel<<<n, 256, 0, stream>>>(perm, off, n, xi, yi);
  return cudaGetLastError();
}

cudaError_t launch_cross_pairs(const uint32_t* tperm, const uint32_t* cperm, uint32_t nt, uint32_t nc, uint32_t* xi,
                               uint32_t* yi, cudaStream_t stream) {
  const unsigned long long tot = (unsigned long long)nt * nc;
  if (tot == 0) return cudaSuccess;
  cross_pairs_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(tperm, cperm, nt, nc, xi, yi);
  return cudaGetLastError();
}

cudaError_t launch_scatter_cross(const double* vals, const uint32_t* xi, const uint32_t* yi, unsigned long long n_pairs,
                                 const uint32_t* col_of, double* out, uint32_t ld, cudaStream_t stream) {
  if (n_pairs == 0) return cudaSuccess;
  scatter_cross_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, stream>>>(vals, xi, yi, n_pairs, col_of, out, ld);
  return cudaGetLastError();
}

cudaError_t launch_normalize_cross(double* out, uint32_t nt, uint32_t nc, uint32_t ld, const double* selfv,
                                   const double* diag, const uint32_t* cols, cudaStream_t stream) {
  const unsigned long long tot = (unsigned long long)nt * nc;
  if (tot == 0) return cudaSuccess;
  normalize_cross_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(out, nt, nc, ld, selfv, diag, cols);
  return cudaGetLastError();
}

cudaError_t launch_scatter_vec(const double* vals, const uint32_t* idx, uint32_t n, double* out, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  scatter_vec_kernel<<<(n + 255) / 256, 256, 0, stream>>>(vals, idx, n, out);
  return cudaGetLastError();
}

cudaError_t launch_deal_pairs(const uint32_t* xi, const uint32_t* yi, unsigned long long n_pairs, uint32_t rank, uint32_t world,
                              uint32_t* out_x, uint32_t* out_y, cudaStream_t stream) {
  const unsigned long long mine = n_pairs > rank ? (n_pairs - rank + world - 1) / world : 0;
  if (mine == 0) return cudaSuccess;
  deal_pairs_kernel<<<(unsigned)((mine + 255) / 256), 256, 0, stream>>>(xi, yi, n_pairs, rank, world, out_x, out_y);
  return cudaGetLastError();
}

cudaError_t launch_undeal(const double* gathered, unsigned long long per, uint32_t world, unsigned long long n_pairs,
                          double* vals, cudaStream_t stream) {
  if (n_pairs == 0) return cudaSuccess;
  undeal_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, stream>>>(gathered, per, world, n_pairs, vals);
  return cudaGetLastError();
}

cudaError_t launch_iota(uint32_t* a, uint32_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  iota_kernel<<<(n + 255) / 256, 256, 0, stream>>>(a, n);
  return cudaGetLastError();
}

}  // namespace stemk

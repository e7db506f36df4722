// Phase A access pattern, two more ways: 16-byte loads per lane, and TMA bulk copies (cp.async.bulk + mbarrier) into
// shared memory followed by a shared-memory sum.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int NB>
__global__ void k128(const double* slab, size_t stride_cta, uint32_t rows_per_slab, uint32_t ny, double* out, long long* cyc, int iters) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* base = slab + (size_t)blockIdx.x * stride_cta;
  uint32_t seed = blockIdx.x * 7919u + warp * 104729u + 1u;
  double acc = 0.0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    seed = seed * 1664525u + 1013904223u;
    const uint32_t off_l = (((seed >> 8) + lane * 2654435761u) % rows_per_slab) * ny;
    double2 v[NB][2];
#pragma unroll
    for (int t = 0; t < NB; ++t) {
      const double2* src = reinterpret_cast<const double2*>(base + __shfl_sync(0xffffffffu, off_l, t)) + lane;
#pragma unroll
      for (int u = 0; u < 2; ++u) v[t][u] = __ldcg(src + 32 * u);   // 2 x 512 B = the same 128 columns as 4 x 256 B
    }
#pragma unroll
    for (int t = 0; t < NB; ++t) acc += (v[t][0].x + v[t][0].y) + (v[t][1].x + v[t][1].y);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
// one team of `blockDim.x` threads: thread 0 issues NB bulk copies of one whole row (ny doubles) each, everybody waits
// on the mbarrier, then lanes <-> columns sum the staged rows.
template <int NB>
__global__ void ktma(const double* slab, size_t stride_cta, uint32_t rows_per_slab, uint32_t ny, double* out, long long* cyc, int iters) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long mbar;
  const uint32_t tid = threadIdx.x;
  const double* base = slab + (size_t)blockIdx.x * stride_cta;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm), mb = (uint32_t)__cvta_generic_to_shared(&mbar);
  if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
  __syncthreads();
  uint32_t seed = blockIdx.x * 7919u + 1u, phase = 0;
  double acc = 0.0;
  const uint32_t row_bytes = ny * 8;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(row_bytes * NB) : "memory");
      for (int t = 0; t < NB; ++t) {
        seed = seed * 1664525u + 1013904223u;
        const double* src = base + (size_t)((seed >> 8) % rows_per_slab) * ny;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(sbase + t * row_bytes), "l"(src), "r"(row_bytes), "r"(mb) : "memory");
      }
    }
    uint32_t done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mb), "r"(phase) : "memory");
    phase ^= 1;
    for (uint32_t j = tid; j < ny; j += blockDim.x)
#pragma unroll
      for (int t = 0; t < NB; ++t) acc += *reinterpret_cast<const double*>(sm + t * row_bytes + 8 * j);
    __syncthreads();
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + tid] = acc;
  if (tid == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* dc; cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&dc, 8);
  const uint32_t ny = 384; const int iters = 400;
  for (size_t mb : {1, 13}) {
    const size_t stride = mb * 1024 * 1024 / 8;
    double* slab; cudaMalloc(&slab, stride * 8 * 148); cudaMemset(slab, 0, stride * 8 * 148);
    const uint32_t rows = (uint32_t)(stride / ny);
    long long h;
    for (int warps : {1, 4, 12}) {
      k128<4><<<148, 32 * warps>>>(slab, stride, rows, ny, out, dc, iters); cudaDeviceSynchronize();
      k128<4><<<148, 32 * warps>>>(slab, stride, rows, ny, out, dc, iters); cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
      printf("slab %2zu MB/CTA %2d warps, 16-byte loads, 4 rows x 128 columns (8 loads/lane): %6.0f cycles/batch %5.1f B/clk/SM\n", mb, warps, (double)h / iters, warps * 4 * 1024.0 / ((double)h / iters));
      k128<8><<<148, 32 * warps>>>(slab, stride, rows, ny, out, dc, iters); cudaDeviceSynchronize();
      k128<8><<<148, 32 * warps>>>(slab, stride, rows, ny, out, dc, iters); cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
      printf("slab %2zu MB/CTA %2d warps, 16-byte loads, 8 rows x 128 columns (16 loads/lane): %6.0f cycles/batch %5.1f B/clk/SM\n", mb, warps, (double)h / iters, warps * 8 * 1024.0 / ((double)h / iters));
    }
    cudaFuncSetAttribute(ktma<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * 3072);
    cudaFuncSetAttribute(ktma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * 3072);
    for (int thr : {128, 256}) {
      ktma<9><<<148, thr, 9 * 3072>>>(slab, stride, rows, ny, out, dc, iters); cudaDeviceSynchronize();
      ktma<9><<<148, thr, 9 * 3072>>>(slab, stride, rows, ny, out, dc, iters); cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
      printf("slab %2zu MB/CTA TMA bulk, 9 rows of 3 KB per batch, %d threads: %6.0f cycles/batch %5.1f B/clk/SM (%s)\n", mb, thr, (double)h / iters, 9 * 3072.0 / ((double)h / iters), cudaGetErrorString(cudaGetLastError()));
      ktma<4><<<148, thr, 9 * 3072>>>(slab, stride, rows, ny, out, dc, iters); cudaDeviceSynchronize();
      ktma<4><<<148, thr, 9 * 3072>>>(slab, stride, rows, ny, out, dc, iters); cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
      printf("slab %2zu MB/CTA TMA bulk, 4 rows of 3 KB per batch, %d threads: %6.0f cycles/batch %5.1f B/clk/SM (%s)\n", mb, thr, (double)h / iters, 4 * 3072.0 / ((double)h / iters), cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(slab);
  }
  return 0;
}

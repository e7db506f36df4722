// Micro-benchmark of the stem kernels' phase A access pattern: a warp sums NB rows picked at random from its CTA's
// slab, NB*NC independent 8-byte loads per lane in flight (rows x 32-column pieces), no branches around the loads.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int NB, int NC>
__global__ void k(const double* slab, size_t stride_cta, uint32_t rows_per_slab, uint32_t ny, double* out, long long* cyc, int iters) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* base = slab + (size_t)blockIdx.x * stride_cta;
  uint32_t seed = blockIdx.x * 7919u + warp * 104729u + 1u;
  double acc = 0.0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    seed = seed * 1664525u + 1013904223u;
    const uint32_t off_l = (((seed >> 8) + lane * 2654435761u) % rows_per_slab) * ny;   // lane t holds row t's offset
    double v[NB][NC];
#pragma unroll
    for (int t = 0; t < NB; ++t) {
      const double* src = base + __shfl_sync(0xffffffffu, off_l, t) + lane;
#pragma unroll
      for (int u = 0; u < NC; ++u) v[t][u] = __ldcg(src + 32 * u);
    }
#pragma unroll
    for (int t = 0; t < NB; ++t)
#pragma unroll
      for (int u = 0; u < NC; ++u) acc += v[t][u];
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int NB, int NC>
void run(const double* slab, size_t stride, uint32_t rows, double* out, long long* dc, int warps, size_t mb) {
  const int iters = 400;
  k<NB, NC><<<148, 32 * warps>>>(slab, stride, rows, 384, out, dc, iters); cudaDeviceSynchronize();
  k<NB, NC><<<148, 32 * warps>>>(slab, stride, rows, 384, out, dc, iters);
  long long h; cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / iters;
  printf("slab %4zu MB/CTA %2d warps/CTA batch %d rows x %d pieces (%2d loads/lane): %6.0f cycles/batch, %5.1f B/clk/SM\n", mb, warps, NB, NC,
         NB * NC, per, warps * NB * NC * 256.0 / per);
}
int main() {
  double* out; long long* dc; cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&dc, 8);
  for (size_t mb : {1, 13}) {
    const size_t stride = mb * 1024 * 1024 / 8;
    double* slab; cudaMalloc(&slab, stride * 8 * 148); cudaMemset(slab, 0, stride * 8 * 148);
    const uint32_t rows = (uint32_t)(stride / 384);
    for (int warps : {1, 4, 12, 24}) {
      run<1, 4>(slab, stride, rows, out, dc, warps, mb);
      run<2, 4>(slab, stride, rows, out, dc, warps, mb);
      run<4, 4>(slab, stride, rows, out, dc, warps, mb);
      run<8, 4>(slab, stride, rows, out, dc, warps, mb);
      run<4, 2>(slab, stride, rows, out, dc, warps, mb);
      run<8, 1>(slab, stride, rows, out, dc, warps, mb);
    }
    cudaFree(slab);
  }
  return 0;
}

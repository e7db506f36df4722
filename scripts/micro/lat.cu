// Latency micro-benchmarks for the stem kernels' design assumptions (dependent DFMA / DADD, LDS, named barrier, LDG).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_dfma(double* out, long long* cyc, double a, double b) {
  double x = a;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) { x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a); }
  long long t1 = clock64();
  double y = a;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) { y = y + b; y = y + b; y = y + b; y = y + b; }
  long long t2 = clock64();
  out[threadIdx.x] = x + y;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
}
__global__ void k_lds(uint32_t* out, long long* cyc) {
  __shared__ uint32_t s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 37 + 5) & 1023;
  __syncthreads();
  uint32_t p = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) { p = s[p]; p = s[p]; p = s[p]; p = s[p]; }
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_bar(long long* cyc, int n) {
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1024; ++i) asm volatile("bar.sync %0, %1;" ::"r"(1), "r"(n) : "memory");
  long long t1 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1024; ++i) __syncthreads();
  long long t2 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
}
__global__ void k_ldg(const uint32_t* buf, uint32_t* out, long long* cyc, int cg) {
  uint32_t p = threadIdx.x * 977u;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) p = cg ? __ldcg(buf + p) : __ldg(buf + p);
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_fill(uint32_t* buf, uint32_t n, uint32_t mul) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) buf[i] = (uint32_t)(((unsigned long long)i * mul + 12345u) % n);
}
int main() {
  double* dout; long long* dc; uint32_t* uo;
  cudaMalloc(&dout, 8 * 1024); cudaMalloc(&dc, 64); cudaMalloc(&uo, 4 * 1024);
  long long h[2];
  for (int rep = 0; rep < 2; ++rep) {
    k_dfma<<<1, 32>>>(dout, dc, 1.0000001, 0.9999999); cudaMemcpy(h, dc, 16, cudaMemcpyDeviceToHost);
    printf("dependent DFMA %.1f cyc, DADD %.1f cyc (1 warp)\n", h[0] / 1024.0, h[1] / 1024.0);
    k_dfma<<<1, 512>>>(dout, dc, 1.0000001, 0.9999999); cudaMemcpy(h, dc, 16, cudaMemcpyDeviceToHost);
    printf("dependent DFMA %.1f cyc, DADD %.1f cyc (16 warps)\n", h[0] / 1024.0, h[1] / 1024.0);
    k_lds<<<1, 32>>>(uo, dc); cudaMemcpy(h, dc, 8, cudaMemcpyDeviceToHost);
    printf("dependent LDS %.1f cyc\n", h[0] / 1024.0);
    for (int n : {32, 128, 256, 512}) {
      k_bar<<<1, n>>>(dc, n); cudaMemcpy(h, dc, 16, cudaMemcpyDeviceToHost);
      printf("bar.sync named (%d thr) %.1f cyc, __syncthreads %.1f cyc\n", n, h[0] / 1024.0, h[1] / 1024.0);
    }
    for (size_t mb : {8, 64, 2048}) {
      uint32_t n = (uint32_t)(mb * 1024 * 1024 / 4); uint32_t* buf; cudaMalloc(&buf, (size_t)n * 4);
      k_fill<<<1024, 256>>>(buf, n, 2654435761u); cudaDeviceSynchronize();
      for (int cg = 0; cg < 2; ++cg) {
        k_ldg<<<1, 32>>>(buf, uo, dc, cg); cudaMemcpy(h, dc, 8, cudaMemcpyDeviceToHost);
        printf("dependent LDG%s over %zu MB: %.0f cyc\n", cg ? ".cg" : ".nc", mb, h[0] / 256.0);
      }
      cudaFree(buf);
    }
  }
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("clock %d kHz\n", clk);
  return 0;
}

// stem_kernel_b200/csrc/stem_common.cuh -- device helpers shared by the two fast stem kernels (stem_fast.cu, stem_tile.cu):
// shared-memory accessors on raw 32-bit addresses, the TMA row prefetch, the band search, the pair slot of a group.
#pragma once
#include "kernels.cuh"

namespace stemk {
namespace {

__device__ __forceinline__ double warp_sum_all(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint32_t ld_flag_f(uint32_t addr) {
  uint32_t v; asm volatile("ld.volatile.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v;
}

// Shared-memory accessors on raw 32-bit shared addresses.  Going through generic pointers makes the compiler rebuild
// the shared window base (S2R SR_CgaCtaId + LEA) next to the accesses of the inner loops; these keep it to one
// cvta per kernel.  All of them are volatile: they keep their program order around __syncwarp / __syncthreads.
__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ double2 lds_v2f64(uint32_t a) { double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds_v2u32(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds_v4u32(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_v2f64(uint32_t a, double2 v) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_v4u32(uint32_t a, uint4 v) { asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
// Makes a value opaque to the optimiser: base addresses of the shared-memory carve-up are otherwise rematerialised
// from the kernel parameters inside the innermost loops (a dozen integer instructions per DP cell).
__device__ __forceinline__ uint32_t pin(uint32_t v) { asm volatile("mov.b32 %0, %0;" : "+r"(v)); return v; }

// The slabs of all CTAs together are several times the L2, so about half of the finished rows a parent row sums have
// gone back to DRAM by the time it needs them.  Ask the L2 for all of them up front instead of discovering the misses a
// few loads at a time: one bulk prefetch per inner pair, issued by the lane that holds its row offset.  It goes
// through the TMA unit, not through the load/store pipe -- the prefetch.global.L2 loop this replaces cost one
// tag wavefront per 128-byte line, an eighth of all the load/store wavefronts of the kernel.
__device__ __forceinline__ void prefetch_rows(const double* G0, uint32_t off_l, bool mine, uint32_t NYS, uint32_t Ny,
                                              uint32_t lane, uint32_t ne) {
#if defined(STEMK_NO_PREFETCH)
  (void)G0; (void)off_l; (void)mine; (void)NYS; (void)Ny; (void)lane; (void)ne;
#elif defined(STEMK_LSU_PREFETCH)
  (void)mine; (void)NYS;
  for (uint32_t tt = 0; tt < ne; ++tt) {
    const double* __restrict__ src = G0 + __shfl_sync(0xffffffffu, off_l, tt);
    for (uint32_t ln = lane * 16u; ln < Ny; ln += 512u) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + ln));
  }
#else
  (void)Ny; (void)lane; (void)ne;
  if (mine) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(G0 + off_l), "r"(8u * NYS) : "memory");
#endif
}

struct PairSlot {        // one pair of the group in flight
  uint32_t k;            // pair number (index into xi / yi / out)
  uint32_t N, node0, blk0, nblk;
  double plr;
};

// number of sorted entries (len << 16 | node, ascending) whose length is below `key`; Ny <= 1024, whole warp
__device__ __forceinline__ uint32_t count_len_below(uint32_t perm, uint32_t Ny, uint32_t key, uint32_t lane) {
  const uint32_t step = (Ny + 31u) >> 5;  // <= 32
  const uint32_t t1 = lane * step;
  const uint32_t c1 = __popc(__ballot_sync(0xffffffffu, t1 < Ny && (lds_u32(perm + 4u * t1) >> 16) < key));
  const uint32_t base = c1 ? (c1 - 1u) * step : 0u;
  const uint32_t t2 = base + lane;
  return base + __popc(__ballot_sync(0xffffffffu, t2 < Ny && (lds_u32(perm + 4u * t2) >> 16) < key));
}

}  // namespace
}  // namespace stemk

// stem_kernel_b200/csrc/kernels.cuh -- launch descriptors shared by the .cu files.
#pragma once
#include <cuda_runtime.h>

#include "stemk_internal.h"

namespace stemk {

struct StemLaunch {
  SetView X, Y;
  const uint32_t* xi;  // device pair list: out[k] = k_stem(X[xi[k]], Y[yi[k]])
  const uint32_t* yi;
  unsigned long long n_pairs;
  double* out;
  unsigned long long* counter;  // work queue head (zeroed before launch)
  double* scratch;              // per-CTA slab holding the G0 rows of the pair in flight
  unsigned long long scratch_stride;  // doubles per CTA
  const double* pair_tab;       // 256 doubles
  uint32_t len_band;
  uint32_t nslots;              // warps of a CTA that own a (Q row, G1 row) pair in shared memory
  uint32_t nx_cap;              // largest Nx over the x set (row flags)
  uint32_t ny_cap;              // largest Ny / Ey / level count over the y set (shared-memory carve-up)
  uint32_t ey_cap;
  uint32_t lev_cap;
  const uint32_t* order;                 // optional: pair numbers to run (bucket 0 of the classifier); NULL = 0..n_pairs-1
  const unsigned long long* n_items_dev; // with `order`: how many of them (device memory)
};

constexpr int kMaxFastBuckets = 8;
#ifndef STEMK_GROUP
#define STEMK_GROUP 8
#endif
constexpr uint32_t kFastGroup = STEMK_GROUP;  // == kGroup of stem_fast.cu
#ifndef STEMK_MAXWARPS
#define STEMK_MAXWARPS 24
#endif
constexpr int kFastMaxWarps = STEMK_MAXWARPS;  // warps per CTA of the fast stem kernel (its launch bound)
#ifndef STEMK_REGWARPS
#define STEMK_REGWARPS 20
#endif
constexpr int kFastRegWarps = STEMK_REGWARPS;  // ... of its variants that keep a whole row in registers in phase A (102 registers per thread)

// fast (separable) stem kernel: runs the pairs order[start[bucket] .. + count[bucket])
struct StemFastLaunch {
  SetView X, Y;
  const uint32_t* xi;
  const uint32_t* yi;
  double* out;
  const uint32_t* order;
  const unsigned long long* start;   // [n_buckets] device
  const unsigned long long* count;   // [n_buckets] device
  unsigned long long* counter;       // this bucket's work-queue head (zeroed by the classifier)
  int bucket;
  double* scratch;                   // per-CTA slabs of pre-scaled G0 rows (one per pair of a group)
  unsigned long long scratch_stride; // doubles per CTA (all slabs of the group)
  double* rowacc;                    // per CTA: kFastGroup x nx_cap per-row result slots
  const double* pair_tab;
  uint32_t len_band, nx_cap, ny_cap, e4_cap, lev_cap;
  uint32_t band_cap;                 // entries of the per-warp MATCH buffer (>= the y set's max_band_cnt)
  unsigned long long* prof;          // optional (FAST_PROF builds): per-phase cycle counters summed over warps
};

struct StemClassify {
  SetView X, Y;
  const uint32_t* xi;
  const uint32_t* yi;
  unsigned long long n_pairs;
  double* out;                       // trivial pairs (an empty DAG on either side) are finished by the classifier
  unsigned long long* count;         // [n_buckets]
  unsigned long long* start;         // [n_buckets]
  uint32_t* order;                   // [n_pairs]
  uint32_t caps[kMaxFastBuckets];    // staged-record size limit of fast bucket b (bucket number 1+b)
  int n_caps;
  int allow_fast;
  // pairs of the general kernel whose records do not fit its shared-memory carve-up (x rows > gen_nx, y nodes > gen_ny
  // or y inner edges > gen_ey) go to bucket `big_bucket` (the unstaged kernel); big_bucket < 0: no such bucket
  int big_bucket;
  uint32_t gen_nx, gen_ny, gen_ey;
};

// unstaged general stem kernel: runs the pairs order[start[bucket] .. + count[bucket]) with nothing in shared memory
struct StemBigLaunch {
  SetView X, Y;
  const uint32_t* xi;
  const uint32_t* yi;
  double* out;
  const uint32_t* order;
  const unsigned long long* start;   // [n_buckets] device
  const unsigned long long* count;   // [n_buckets] device
  unsigned long long* counter;       // this bucket's work-queue head (zeroed by the classifier)
  int bucket;
  double* scratch;                   // per CTA: G0 slab nx_cap x ny_pitch | per warp a Q row and a G1 row | nx_cap row flags
  unsigned long long scratch_stride; // doubles per CTA (stem_unstaged_scratch_doubles)
  const double* pair_tab;
  uint32_t len_band, nx_cap, ny_cap;
};

struct StringLaunch {
  SetView X, Y;
  const uint32_t* xi;
  const uint32_t* yi;
  unsigned long long n_pairs;
  double* out;
  unsigned long long* counter;
  double* carry;                // per-group column carries for sequences wider than one tile
  unsigned long long carry_stride;  // doubles per lane group
  uint32_t pow_cap;             // largest power of the gap the kernel needs (max sequence length + 1)
  const double* subst;          // 16 doubles
  double gap;
  int naive;                    // exact-match kernel on raw characters
};

// host-callable launchers (defined in the .cu files); all asynchronous on `stream`
size_t stem_smem_bytes(uint32_t nslots, uint32_t nx_cap, uint32_t ny_cap, uint32_t ey_cap, uint32_t lev_cap);
int stem_warps_per_cta();
cudaError_t launch_stem(const StemLaunch& p, int grid, size_t smem, cudaStream_t stream);
int stem_max_ctas_per_sm(size_t smem);
unsigned long long stem_unstaged_scratch_doubles(uint32_t nx_cap, uint32_t ny_cap);
cudaError_t launch_stem_unstaged(const StemBigLaunch& p, int grid, cudaStream_t stream);
size_t stem_fast_smem_bytes(uint32_t nwarps, uint32_t nx_cap, uint32_t ny_cap, uint32_t e4_cap, uint32_t lev_cap,
                            uint32_t band_cap);
int stem_fast_ctas_per_sm(uint32_t ny_cap, int nwarps, size_t smem);
int stem_fast_max_warps(uint32_t ny_cap);   // launch bound of the variant that serves staged records of up to ny_cap nodes
cudaError_t launch_stem_fast(const StemFastLaunch& p, int grid, int nwarps, size_t smem, cudaStream_t stream);
cudaError_t launch_classify(const StemClassify& c, int n_buckets, unsigned long long* counters, cudaStream_t stream);
void string_shape_for(uint32_t ly_cap, int mode, int* cw, int* tp);
// mode: 0 plain one-hot columns, 1 weighted, 2 naive characters, 3 general (see string_kernel.cu)
cudaError_t launch_string(const StringLaunch& p, int cw, int tp, int mode, int grid, cudaStream_t stream);
int string_warps_per_cta();
cudaError_t launch_combine(int kind, double alpha, double beta, const double* stem, const double* str, double* out,
                           unsigned long long n, cudaStream_t stream);
cudaError_t launch_scatter_square(const double* vals, const uint32_t* xi, const uint32_t* yi, unsigned long long n_pairs,
                                  double* matrix, uint32_t n, cudaStream_t stream);
cudaError_t launch_normalize_square(double* matrix, uint32_t n, cudaStream_t stream);
// matrix_ops.cu: work-order pair lists and the rectangular matrix, on the device
cudaError_t launch_gram_pairs(const uint32_t* perm, const unsigned long long* off, uint32_t n, uint32_t* xi, uint32_t* yi,
                              cudaStream_t stream);
cudaError_t launch_cross_pairs(const uint32_t* tperm, const uint32_t* cperm, uint32_t nt, uint32_t nc, uint32_t* xi,
                               uint32_t* yi, cudaStream_t stream);
cudaError_t launch_scatter_cross(const double* vals, const uint32_t* xi, const uint32_t* yi, unsigned long long n_pairs,
                                 const uint32_t* col_of, double* out, uint32_t ld, cudaStream_t stream);
cudaError_t launch_normalize_cross(double* out, uint32_t nt, uint32_t nc, uint32_t ld, const double* selfv,
                                   const double* diag, const uint32_t* cols, cudaStream_t stream);
cudaError_t launch_scatter_vec(const double* vals, const uint32_t* idx, uint32_t n, double* out, cudaStream_t stream);
cudaError_t launch_iota(uint32_t* a, uint32_t n, cudaStream_t stream);
cudaError_t launch_deal_pairs(const uint32_t* xi, const uint32_t* yi, unsigned long long n_pairs, uint32_t rank, uint32_t world,
                              uint32_t* out_x, uint32_t* out_y, cudaStream_t stream);
cudaError_t launch_undeal(const double* gathered, unsigned long long per, uint32_t world, unsigned long long n_pairs,
                          double* vals, cudaStream_t stream);
// BPLA / local-alignment kernels (bpla.cu): host buffers in, host buffer out, synchronous on `stream`
cudaError_t run_bpla(const stemk_bpla_params& p, const stemk_bpla_set& x, const stemk_bpla_set& y, size_t n_pairs,
                     const uint32_t* xi, const uint32_t* yi, double* out, double* grad, int sm_count, size_t smem_optin,
                     cudaStream_t stream, std::string* err);   // grad != NULL: BPLAKernel::compute_gradients, 4 doubles per pair
// naive stem kernel (nstem.cu): host buffers in, host buffer out, synchronous on `stream`
cudaError_t run_nstem(const stemk_nstem_params& p, const stemk_nstem_set& x, const stemk_nstem_set& y, size_t n_pairs,
                      const uint32_t* xi, const uint32_t* yi, double* out, uint32_t band, int sm_count, size_t smem_optin,
                      cudaStream_t stream, std::string* err,   // band > 0: partial_dp with the band-only constraints
                      const uint32_t* win_off = nullptr, const uint32_t* c_low = nullptr, const uint32_t* c_high = nullptr);   // or the caller's per-row windows
// base-pair probabilities (fold.cu): host buffers in, host vectors out, synchronous on `stream`
struct FoldResult {
  std::vector<uint64_t> pair_off;     // [n_seqs + 1]
  std::vector<uint32_t> bi, bj;       // 1-based, i < j, per sequence in ascending (i, j)
  std::vector<double> bp, unpaired, ensemble, dense;
  double kernel_ms = 0;               // device time of the fold kernel (CUDA events on the launching stream)
};
cudaError_t run_fold(const stemk_fold_model& m, uint32_t n_seqs, const uint64_t* seq_off, const char* text, double cutoff,
                     bool want_dense, int sm_count, cudaStream_t stream, FoldResult* res, std::string* err, void** scratch_p,
                     size_t* scratch_bytes);   // *scratch_p: device buffer the caller keeps between calls (grown here)
cudaError_t launch_fp64_peak(double* sink, int grid, int block, int iters, cudaStream_t stream);

}  // namespace stemk

/* include/stemk.h -- C ABI of the B200 stem-kernel / string-kernel Gram-matrix builder.
 *
 * This is the drop-in boundary for ONE path of keio-bioinformatics/stem_kernel: evaluating a
 * kernel functor over many sequence pairs.  In the reference that path is entered through
 *
 *     value_type Kernel::operator()(const Data&, const Data&) const
 *         stem_kernel_lite/stem_kernel.h:18, string_kernel.h:20, def_kernel.h:21,45,72,100,130,156,184
 *         string_kernel/string_kernel.h:19 (naive kernel on std::string)
 *
 * called one pair at a time from the Calc* functors of common/kernel_matrix.cpp (lines 50, 95,
 * 104, 159, 168, 176) on behalf of the four KernelMatrix entry points
 *
 *     calculate(train, kernel, normalize, n_th)                 kernel_matrix.h:67-69   -> stemk_gram
 *     calculate(test, train, kernel, norm_test, normalize,n_th) kernel_matrix.h:71-74   -> stemk_cross
 *     calculate(row, example, train, sv_index, kernel, ...)     kernel_matrix.h:76-82   -> stemk_cross (1 row)
 *     diagonal(diag, train, sv_index, kernel, n_th)             kernel_matrix.h:94-97   -> stemk_diag
 *
 * A reference maintainer binds these entry points from a KernelMatrix specialisation (see
 * INTEGRATION.md); stem_kernel_b200/host/ holds a C++ mirror of the reference's kernel classes and
 * KernelMatrix built on them.
 *
 * Conventions: plain pointers and sizes only; all host buffers are owned by the caller and may be
 * freed as soon as the call returns; device buffers are owned by the context / set.  Every function
 * returns STEMK_OK (0) or a negative error code; stemk_last_error() gives the message.  There is no
 * CPU fallback: without a CUDA device every compute entry point fails with STEMK_ERR_CUDA.
 * All arithmetic is IEEE fp64 on the device (the reference's ValueType is double,
 * stem_kernel.cpp:103, string_kernel.cpp:139).
 */
#ifndef STEMK_H_
#define STEMK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STEMK_OK 0
#define STEMK_ERR_ARG (-1)    /* bad argument / inconsistent descriptor */
#define STEMK_ERR_CUDA (-2)   /* CUDA runtime error or no device */
#define STEMK_ERR_NOMEM (-3)
#define STEMK_ERR_STATE (-4)  /* set does not carry what the kernel needs (e.g. no DAG for a stem kernel) */

/* Kernel selection = the reference's kernel classes (stem_kernel_lite/def_kernel.h, main.cpp:180-215). */
typedef enum {
  STEMK_SI_STEM = 0,      /* SiStemKernel:     stem DAG kernel, match/mismatch node score (def_kernel.h:12)   */
  STEMK_SU_STEM = 1,      /* SuStemKernel:     stem DAG kernel, exp(beta*RIBOSUM) node score (def_kernel.h:36) */
  STEMK_SI_STEM_STR = 2,  /* SiStemStrKernel = SiStem + StringKernel(gap,match,mismatch) (def_kernel.h:59)   */
  STEMK_SU_STEM_STR = 3,  /* SuStemStrKernel = SuStem + StringKernel(gap,alpha)          (def_kernel.h:87)   */
  STEMK_LSU_STEM = 4,     /* beta*log(SuStem)                                            (def_kernel.h:114)  */
  STEMK_LSU_STR = 5,      /* alpha*log(StringKernel(gap,alpha))                          (def_kernel.h:140)  */
  STEMK_LSU_STEM_STR = 6, /* sum of the two above                                        (def_kernel.h:166)  */
  STEMK_STR_SUBST = 7,    /* lite StringKernel(gap, alpha)             (stem_kernel_lite/string_kernel.cpp:11) */
  STEMK_STR_SIMPLE = 8,   /* lite StringKernel(gap, match, mismatch)   (stem_kernel_lite/string_kernel.cpp:24) */
  STEMK_STR_NAIVE = 9     /* exact-match gap-weighted kernel on raw characters (string_kernel/string_kernel.cpp:11) */
} stemk_kind;

typedef struct {
  int32_t kind;        /* stemk_kind */
  uint32_t len_band;   /* --length-band; 0 = no band (stem_kernel.cpp:46-48) */
  double loop_gap;     /* -g  (stem) */
  double beta;         /* -b  (stem, RIBOSUM weight) */
  double stack;        /* -s  (stem, --no-ribosum match) */
  double covar;        /* -v  (stem, --no-ribosum mismatch) */
  double gap;          /* -G  (string); STR_NAIVE: the float-parsed -g of string_kernel/main.cpp:40, widened */
  double alpha;        /* -a  (string, RIBOSUM weight) */
  double match;        /* --match    (string, --no-ribosum) */
  double mismatch;     /* --mismatch (string, --no-ribosum) */
} stemk_params;

/* A set of sequence records, flattened: the fields of the reference's MData
 * (stem_kernel_lite/data.h:33-37; DAG::Node / DAG::Edge, dag.h:17-157) as concatenated arrays.
 * Node order inside a record must be the reference's (children before parents, data.cpp:193-244).
 * A record with zero nodes is legal (no pair reached the threshold: k_stem = 0). */
typedef struct {
  uint32_t n_seqs;
  /* DAG: per-record node ranges, then CSR over all nodes */
  const uint32_t* node_off;     /* [n_seqs+1] */
  const uint32_t* node_first;   /* [n_nodes] 0-based column of the 5' base */
  const uint32_t* node_last;    /* [n_nodes] first==last marks a leaf */
  const float* node_weight;     /* [n_nodes] */
  const uint32_t* edge_off;     /* [n_nodes+1] */
  const uint32_t* edge_to;      /* [n_edges] child, as an index LOCAL to its record */
  const uint32_t* edge_gaps;    /* [n_edges] */
  const float* edge_weight;     /* [n_edges] */
  const uint32_t* bpf_off;      /* [n_nodes+1] */
  const uint8_t* bpf_a;         /* [n_bpf] */
  const uint8_t* bpf_b;         /* [n_bpf] */
  const float* bpf_freq;        /* [n_bpf] */
  const uint32_t* root_off;     /* [n_seqs+1] */
  const uint32_t* root;         /* [n_roots] local node indices */
  /* sequence profile */
  const uint32_t* col_off;      /* [n_seqs+1] */
  const float* profile;         /* [n_cols*5]  A,C,G,U,GAP per column (common/profile.h:15-16) */
  const float* n_rows;          /* [n_seqs]    ProfileSequence::n_seqs() */
  const uint32_t* weight_off;   /* [n_seqs+1]  empty range = record has no weight vector */
  const float* col_weight;      /* [n_weights] MData::weight (data.cpp:437-453) */
  const uint8_t* text;          /* [n_cols] raw characters of the first row (STEMK_STR_NAIVE); may be NULL */
} stemk_seqset_desc;

typedef struct stemk_ctx stemk_ctx;
typedef struct stemk_set stemk_set;

/* library / device */
const char* stemk_version(void);
int stemk_device_count(void);

/* context: one CUDA device + one kernel object (immutable, like the reference's kernel classes).
 * device = STEMK_DEVICE_NONE makes a host-only context: stemk_upload compiles the records and stemk_set_stats /
 * stemk_pair_cost (schedulers, work model) work, every entry point that computes a kernel value fails with
 * STEMK_ERR_CUDA. */
#define STEMK_DEVICE_NONE (-1)
int stemk_create(stemk_ctx** ctx, const stemk_params* params, int device);
void stemk_destroy(stemk_ctx* ctx);
const char* stemk_last_error(const stemk_ctx* ctx); /* ctx may be NULL: last creation error */

/* Context options (test and diagnostic hooks; there is no environment-variable dispatch):
 *   STEMK_OPT_FORCE_GENERAL  1: every stem pair runs on the general kernel (stem_kernel.cu) instead of the separable
 *                            fast path -- what alignments / IUPAC records take anyway; lets tests compare the two
 *   STEMK_OPT_TIMING         1: stemk_upload / stemk_gram print a host-side time breakdown on stderr
 *   STEMK_OPT_FORCE_UNSTAGED 1: the pairs of the general kernel all run on its unstaged variant (the one that takes
 *                            records of any size); lets tests cover it with small records */
#define STEMK_OPT_FORCE_GENERAL 1
#define STEMK_OPT_TIMING 2
#define STEMK_OPT_FORCE_UNSTAGED 3
int stemk_set_option(stemk_ctx* ctx, int option, int value);

/* Threading: a context is single-threaded and has ONE call in flight -- its scratch buffers, work queues and DP
 * slabs belong to the call.  stemk_pairs_device / stemk_assemble_device are asynchronous on the caller's stream; a
 * call on a DIFFERENT stream than the previous one first waits (on the device) for the previous call's work.  Use one
 * context per host thread (the kernel object is immutable, like the reference's functors, and cheap to create).
 * Sets are bound to the device, loop gap and length band of the context that uploaded them (the derived per-record
 * tables depend on them); using a set with another context fails with STEMK_ERR_ARG. */

/* Record sizes.  The stem kernels take records of ANY size, like the reference's operator(): the fast stem kernel
 * runs records of up to 1024 non-leaf DAG nodes whose gap powers g^(len/2) stay within 1e-125 .. 1e125 (about 360 nt of
 * pair span at g = 0.2) with unit edge weights, single-entry base-pair profiles and no gap columns under a node --
 * every record the reference's front end makes from one sequence; everything else (alignments, IUPAC codes, hand-made
 * DAGs, longer records) runs on the general stem kernel, which stages the second record of a pair in shared memory
 * (up to about 1000 non-leaf nodes / 8000 inner edges), and pairs with a larger record run on its unstaged variant,
 * which keeps everything in global memory (slower per pair; bounded only by device memory: 8 B x nodes_x x nodes_y of
 * scratch per pair in flight).  The string kernels have no length limit. */

/* upload a flattened set (copied; derived per-record tables are built here) */
int stemk_upload(stemk_ctx* ctx, const stemk_seqset_desc* desc, stemk_set** set);
void stemk_set_free(stemk_ctx* ctx, stemk_set* set);
uint32_t stemk_set_size(const stemk_set* set);
/* Per-record sizes for cost models / schedulers: #V and #E of the DAG (leaves and leaf edges included, as the
 * reference counts them) and the number of columns.  Any of the three output arrays [n] may be NULL. */
void stemk_set_stats(const stemk_set* set, uint32_t* n_nodes, uint32_t* n_edges, uint32_t* length);
/* Bytes of device memory the uploaded set occupies (= host->device bytes copied by stemk_upload). */
uint64_t stemk_set_device_bytes(const stemk_set* set);

/* ---- several devices (the reference's MPI path, kernel_matrix.cpp:186-262, 495-526, 560-571, a device per rank) --
 * One context per device, all created with the same parameters.
 *   stemk_set_clone     copies an uploaded set to the device of `ctx`, device to device (NVLink when the devices are
 *                       peers): no second host compile, no second host->device copy.  Record headers and statistics
 *                       are shared with the source set; either set may be freed first.
 *   stemk_upload_multi  stemk_upload on ctxs[0] + stemk_set_clone onto ctxs[1..n_ctx): sets[d] lives on ctxs[d]'s device.
 *   stemk_gram_multi    KernelMatrix::calculate(train, kernel, normalize) over n_ctx devices driven by ONE host thread:
 *                       the y-major pair order is dealt round-robin, every device evaluates its share on its own stream,
 *                       device 0 gathers the shares over peer copies, assembles and normalises.  out: n*n row-major,
 *                       the same values as stemk_gram on one device.  Errors of device d > 0 are reported through
 *                       stemk_last_error(ctxs[0]).
 *   stemk_set_export / stemk_set_import  the same for one PROCESS per device (torchrun, MPI): the set as one byte
 *                       string in a caller-provided DEVICE buffer of stemk_set_export_bytes(set) bytes, to be moved
 *                       with the caller's collective (one NCCL broadcast) and turned back into a set on the receiving
 *                       device.  The byte layout is private to the library version that wrote it. */
int stemk_set_clone(stemk_ctx* ctx, const stemk_set* src, stemk_set** set);
int stemk_upload_multi(stemk_ctx* const* ctxs, int n_ctx, const stemk_seqset_desc* desc, stemk_set** sets);
int stemk_gram_multi(stemk_ctx* const* ctxs, const stemk_set* const* sets, int n_ctx, int normalize, double* out);
uint64_t stemk_set_export_bytes(const stemk_set* set);
int stemk_set_export(stemk_ctx* ctx, const stemk_set* set, void* d_dst, void* stream);
int stemk_set_import(stemk_ctx* ctx, const void* d_src, uint64_t bytes, stemk_set** set);

/* KernelMatrix::calculate(train, kernel, normalize) -- kernel_matrix.cpp:485-575.
 * out: n*n row-major, both triangles.  normalize: K_ij /= sqrt(K_ii K_jj), K_ii = 1. */
int stemk_gram(stemk_ctx* ctx, const stemk_set* train, int normalize, double* out);

/* KernelMatrix::calculate(test, train, ...) and the static one-row calculate -- kernel_matrix.cpp:635-754.
 * out: n_test*n_train row-major.  sv_index (may be NULL, n_sv = 0): only those train columns are
 * computed AND WRITTEN, the other entries of out are left untouched (the static row calculate writes vec[sv] only,
 * kernel_matrix.cpp:164-171; a caller that wants the zero-initialised rows of the rectangular calculate,
 * kernel_matrix.cpp:713, clears out first).  self_out (may be NULL): k(test_i, test_i).
 * normalize != 0: out_ij /= sqrt(self_i * k(train_j,train_j)) as kernel_matrix.cpp:735-748, applied to the written
 * columns (train diagonals are computed internally). */
int stemk_cross(stemk_ctx* ctx, const stemk_set* test, const stemk_set* train, const uint32_t* sv_index,
                uint32_t n_sv, int normalize, double* out, double* self_out);

/* KernelMatrix::diagonal -- kernel_matrix.cpp:578-633.  out: n; with sv_index only those entries are written. */
int stemk_diag(stemk_ctx* ctx, const stemk_set* train, const uint32_t* sv_index, uint32_t n_sv, double* out);

/* Arbitrary pair list: out[k] = kernel(x[xi[k]], y[yi[k]]).  The three calls above are built on it;
 * multi-GPU drivers use it for their tile of the matrix. */
int stemk_pairs(stemk_ctx* ctx, const stemk_set* x, const stemk_set* y, size_t n_pairs, const uint32_t* xi,
                const uint32_t* yi, double* out);

/* Same, asynchronous on `stream` (a cudaStream_t, or NULL) with DEVICE index/result buffers.
 * xi/yi must stay valid until the stream reaches the end of the call's work. */
int stemk_pairs_device(stemk_ctx* ctx, const stemk_set* x, const stemk_set* y, size_t n_pairs,
                       const uint32_t* d_xi, const uint32_t* d_yi, double* d_out, void* stream);

/* Rank-0 half of a distributed KernelMatrix::calculate (kernel_matrix.cpp:504-526 gathers the ranks' values,
 * :560-571 normalises): scatter n_pairs gathered values into the n*n DEVICE matrix, mirroring (i,j) to (j,i),
 * then optionally K_ij /= sqrt(K_ii K_jj), K_ii = 1.  Asynchronous on `stream`. */
int stemk_assemble_device(stemk_ctx* ctx, size_t n_pairs, const uint32_t* d_xi, const uint32_t* d_yi,
                          const double* d_vals, uint32_t n, int normalize, double* d_matrix, void* stream);

/* Work model of SURVEY 8(d), per pair: DP cells and algorithmic flops
 * (stem: 2*U_match + 3*U_bf + 3*U_skip; string: 9, 7 or 4(+3 per match) per cell). Host only. */
int stemk_pair_cost(stemk_ctx* ctx, const stemk_set* x, const stemk_set* y, size_t n_pairs, const uint32_t* xi,
                    const uint32_t* yi, double* cells, double* flops);

/* Launch accounting for benchmarks: kernels launched / device ms (CUDA events on the library's own
 * stream) spent inside them since the last reset. */
void stemk_stats_reset(stemk_ctx* ctx);
void stemk_stats_get(stemk_ctx* ctx, uint64_t* launches, double* stem_ms, double* string_ms);

/* FP64 FMA micro-benchmark on the context's device: returns sustained Tflop/s (the roofline
 * denominator, since MEASURED_PEAKS.json carries no fp64 entry). */
int stemk_fp64_peak(stemk_ctx* ctx, double seconds, double* tflops);

/* ---- BPLA / local-alignment kernels: bpla_kernel/bpla_kernel.cpp:64-175 -------------------
 * BPLAKernel<double, MData>::operator() -- the local-alignment kernel with (or, no_bp, without) base-pairing
 * profiles, in its sum-over-alignments form (local_alignment_exp, :64-118) or, sw, its Smith-Waterman form
 * (local_alignment_max, :120-157).  A record is the reference's `Data` (bpla_kernel/data.h:30-53) flattened: per
 * column the 5 profile counts [A,C,G,U,GAP] of ProfileSequence (common/profile.h) and the three base-pairing
 * profiles p_left / p_right / p_unpair as data.cpp:19-46 leaves them (square roots, float).  score: the 4 x 4
 * substitution table (bpla_kernel/main.cpp:20-26 holds the default).  Parameter defaults: main.cpp:69-73. */
typedef struct stemk_bpla_params {
  int32_t no_bp;      /* --noBP: LAScore instead of BPLAScore (bpla_kernel.cpp:16-62) */
  int32_t sw;         /* --SW */
  double gap, ext;    /* -8, -0.75 */
  double alpha, beta; /* 4.5, 0.11 */
  double score[16];   /* [a*4+b] */
} stemk_bpla_params;

typedef struct stemk_bpla_set {
  uint32_t n_seqs;
  const uint32_t* col_off;   /* [n+1] */
  const float* profile;      /* 5 per column */
  const float* p_left;       /* per column; may be NULL when no_bp */
  const float* p_right;
  const float* p_unpair;
} stemk_bpla_set;

/* out[k] = k_bpla(x[xi[k]], y[yi[k]]) on the context's device (host buffers; the two sets are copied to the
 * device by the call).  The context's own kernel kind is irrelevant here.  No CPU path.  Length limit: a warp keeps
 * one row of the five tables in shared memory (64 B per column of the second sequence, 4 warps per CTA): second
 * sequences of up to about 900 columns; longer ones fail the call with STEMK_ERR_NOMEM. */
int stemk_bpla_pairs(stemk_ctx* ctx, const stemk_bpla_params* params, const stemk_bpla_set* x, const stemk_bpla_set* y,
                     size_t n_pairs, const uint32_t* xi, const uint32_t* yi, double* out);

/* BPLAKernel<double, MData>::compute_gradients (bpla_kernel/bpla_kernel.cpp:176-402; bpla_kernel.h:30-34), the call
 * bpla_optimizer.cpp:77,114,151,187,233,244 makes per pair: value[k] = the kernel value of the optimizer's model
 * (BPLA_Forward :178-243 -- not the same number as operator()) and grad[4k .. 4k+3] = its partial derivatives with
 * respect to alpha, beta, gap, ext (params->no_bp and params->sw must be 0; param vector order of :188-191). */
int stemk_bpla_gradients(stemk_ctx* ctx, const stemk_bpla_params* params, const stemk_bpla_set* x, const stemk_bpla_set* y,
                         size_t n_pairs, const uint32_t* xi, const uint32_t* yi, double* value, double* grad);

/* ---- naive stem kernel: stem_kernel/stem_kernel.cpp:282-351 (full_dp) --------------------
 * StemKernel<double, BPMat>::operator() of the stem_kernel/ program with band 0 and no alignment constraint: the
 * O(Lx^2 Ly^2) dynamic program over all pairs of base pairs (i,j) x (k,l) on the raw (lower-case) sequences.
 * bp_mode 0: canonical pairs from the text, NormalBasePair / WobbleBasePair (:353-392; use_gu selects g-u pairs);
 * bp_mode 1: prob(i,j) read from a dense row-major L x L float table per sequence (the role of the
 * ViennaRNA-backed BPMatrix class, :394-420).  Parameter defaults: stem_kernel/main.cpp:46-64 (gap 0.8, stack 1,
 * loop 3, subst 0.5).  The banded partial_dp (:113-280) is stemk_nstem_pairs_banded below. */
typedef struct stemk_nstem_params {
  int32_t bp_mode;
  int32_t use_gu;
  uint32_t loop;
  float bp_bound;      /* a pair counts when prob > bp_bound */
  double gap, stack, subst;
} stemk_nstem_params;

typedef struct stemk_nstem_set {
  uint32_t n_seqs;
  const uint32_t* off;     /* [n+1] character offsets into text */
  const char* text;
  const uint64_t* bp_off;  /* [n] offset of each sequence's L x L table in bp (bp_mode 1) */
  const float* bp;
} stemk_nstem_set;

/* out[k] = k_stem_naive(x[xi[k]], y[yi[k]]) on the context's device (host buffers).  No CPU path.  Length limit: two
 * planes of the SHORTER sequence of a pair live in shared memory (16 B x (L+2)^2: about 115 characters); as in the
 * reference this O(Lx^2 Ly^2) kernel is for short sequences.  Longer pairs fail the call with STEMK_ERR_NOMEM. */
int stemk_nstem_pairs(stemk_ctx* ctx, const stemk_nstem_params* params, const stemk_nstem_set* x, const stemk_nstem_set* y,
                      size_t n_pairs, const uint32_t* xi, const uint32_t* yi, double* out);

/* The same kernel with band > 0 (stem_kernel.h:51-54: operator() takes partial_dp, stem_kernel.cpp:113-280, with the
 * band-only alignment constraints of :77-83; the pair-HMM constraints of ali_bound > 0 are not provided): row i of x may
 * pair with the columns of y within `band` of the diagonal.  The arguments keep their roles (the band is not symmetric). */
int stemk_nstem_pairs_banded(stemk_ctx* ctx, const stemk_nstem_params* params, uint32_t band, const stemk_nstem_set* x,
                             const stemk_nstem_set* y, size_t n_pairs, const uint32_t* xi, const uint32_t* yi, double* out);

/* The same partial_dp under caller-supplied alignment constraints -- what StemKernel::alignment_constraints
 * (stem_kernel/stem_kernel.cpp:14-67) produces from the pair-HMM posteriors when ali_bound > 0, optionally narrowed by
 * the band (:57-66): row i (0..lx) of x[xi[k]] may pair with columns c_low[win_off[k] + i] .. c_high[win_off[k] + i]
 * of y[yi[k]].  win_off has n_pairs + 1 entries, every pair owns lx + 1 of each array, c_low <= c_high <= ly.  The
 * pair HMM itself (phmm.cpp) stays CPU code in front of this call. */
int stemk_nstem_pairs_windows(stemk_ctx* ctx, const stemk_nstem_params* params, const stemk_nstem_set* x, const stemk_nstem_set* y,
                              size_t n_pairs, const uint32_t* xi, const uint32_t* yi, const uint32_t* win_off,
                              const uint32_t* c_low, const uint32_t* c_high, double* out);

/* ---- Front end: base-pair probabilities -------------------------------------------------------
 * Replaces the per-sequence ViennaRNA call of the reference's front end -- fold / init_pf_fold / pf_fold and the copy
 * bp(i,j) = pr[iindx[i]-j] under a process-wide mutex (common/bpmatrix.cpp:141-177) -- by a batched McCaskill
 * partition function on the device: one CTA per sequence, inside and outside recursions swept diagonal by diagonal.
 * ViennaRNA is an external, unversioned dependency that is absent from this image, so the ENERGY MODEL IS AN INPUT:
 * a nearest-neighbour loop model in the shape of the Vienna 1.8 pf_fold recursions with dangles on both sides of
 * every multiloop / exterior stem (-d2): stacking, hairpin / bulge / interior initiation by size (logarithmic
 * extrapolation beyond 30), terminal mismatches, asymmetry (Ninio) term, terminal A-U / G-U penalty, linear
 * multiloop.  Not modelled: the tabulated 1x1 / 2x1 / 2x2 interior loops and special tri/tetraloop bonuses (they take
 * the generic formulas), no_closingGU, noLonelyPairs, alifold.  PARITY WITH ViennaRNA IS UNPINNED: the checker is
 * this repo's own CPU restatement (oracle/stemk_fold_oracle.c), itself checked against an exhaustive enumeration of
 * all secondary structures of short sequences.
 *
 * Pair types: 1 CG, 2 GC, 3 GU, 4 UG, 5 AU, 6 UA (0 = cannot pair); base codes 1 A, 2 C, 3 G, 4 U (0 = anything else:
 * never pairs).  Energies in kcal/mol.  Pairs need at least 3 unpaired bases between them; interior loops have at most
 * 30 unpaired bases. */
typedef struct stemk_fold_model {
  double temperature;          /* degrees Celsius: kT = (temperature + 273.15) * 1.98717e-3 kcal/mol */
  double pf_scale;             /* per-nucleotide scaling of the partition function (Vienna::pf_scale); <= 0: Vienna's
                                  default exp(-(-185 + (temperature - 37) * 7.27) / (1000 kT)) */
  int32_t no_gu;               /* Vienna::noGU */
  int32_t pad_;
  double stack[8][8];          /* [type(i,j)][reversed type of the inner pair] like Vienna's stack37 */
  double hairpin[31], bulge[31], interior[31];   /* initiation by loop size; sizes > 30: [30] + lxc * ln(size / 30) */
  double lxc;
  double mismatch_h[8][5][5];  /* hairpin terminal mismatch [type][base after i][base before j] (loops of > 3) */
  double mismatch_i[8][5][5];  /* interior-loop terminal mismatch, both ends */
  double dangle5[8][5], dangle3[8][5];   /* [type][neighbouring base]; applied on both sides wherever a neighbour exists */
  double ninio, max_ninio;     /* asymmetry: min(max_ninio, |u1 - u2| * ninio) */
  double terminal_au;          /* pair types > 2 closing a bulge of > 1, a hairpin of 3 or an exterior stem */
  double ml_closing, ml_intern[8], ml_base;
} stemk_fold_model;

/* A stand-in parameter set so that tests, the benchmark and the front end have something to fold with: stacking,
 * initiation and multiloop terms of Turner-1999 magnitude, synthetic mismatch / dangle tables.  NOT a published
 * parameter file; real parameters are the caller's. */
void stemk_fold_model_default(stemk_fold_model* model);

/* Base-pair probabilities of n_seqs sequences: sequence k is text[seq_off[k] .. seq_off[k+1]) (case-insensitive acgu/t;
 * anything else never pairs).  The result stays in the context until the next call: *n_pairs_total = number of pairs
 * (i < j, 1-based like BPMatrix) with probability >= cutoff over all sequences (cutoff <= 0: every pair with a
 * non-zero probability).  ensemble (optional, [n_seqs]): -kT ln Z in kcal/mol.  dense (optional): for every sequence
 * an (L+1) x (L+1) row-major table with dense[i*(L+1)+j] = P(i,j) for 1 <= i < j <= L, 0 elsewhere -- the layout of
 * BPMatrix::table_ as the reference reads it -- concatenated in sequence order. */
int stemk_fold_bpp(stemk_ctx* ctx, const stemk_fold_model* model, uint32_t n_seqs, const uint64_t* seq_off, const char* text,
                   double cutoff, uint64_t* n_pairs_total, double* ensemble, double* dense);
/* The pair lists of the last stemk_fold_bpp: pair_off[n_seqs + 1], then bi / bj / bp of *n_pairs_total entries, per
 * sequence in ascending (i, j) -- the sparse per-row lists the front end (host/frontend.cpp: Profiler + DAGBuilder)
 * consumes -- and unpaired (optional, one entry per character of text): max(0, 1 - sum_j P(i,j)) over ALL pairs, not
 * only the listed ones (Profiler's nbp_, stem_kernel_lite/data.cpp:94-123). */
int stemk_fold_fetch(stemk_ctx* ctx, uint64_t* pair_off, uint32_t* bi, uint32_t* bj, double* bp, double* unpaired);
/* Device time (ms, CUDA events on the context's stream) of the kernel of the last stemk_fold_bpp. */
double stemk_fold_last_ms(const stemk_ctx* ctx);

/* Text of kernel-matrix rows in the reference's output format -- KernelMatrix::print (kernel_matrix.cpp:756-770)
 * and Output::kernel_output (framework.cpp:190-204): one line "<label> 0:<cnt> 1:<v> 2:<v> ... \n" per row, every
 * value printed like operator<<(std::ostream&, double) with default flags ("%g").  m: n_rows x n_cols with row
 * stride ld (doubles); labels: n_rows C strings; the row counter of row r is first_cnt + r (1-based in the
 * reference).  Host code only (no device needed), rows are formatted by n_threads threads (<= 0: all cores).
 * Returns the number of bytes of the text and writes it (without a terminating NUL) when it fits into cap. */
size_t stemk_format_rows(const double* m, uint32_t n_rows, uint32_t n_cols, size_t ld, const char* const* labels,
                         uint32_t first_cnt, int n_threads, char* out, size_t cap);

/* Text of the norm file (Output::norm_output, framework.cpp:218-228): one "%g" value per line. */
size_t stemk_format_values(const double* v, size_t n, char* out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* STEMK_H_ */

// stem_kernel_b200/csrc/stemk_internal.h -- internal structures shared by the host-side set
// compiler (compile_set.cpp), the CUDA kernels and the C ABI (stemk_api.cu).
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "../../include/stemk.h"

namespace stemk {

// ------------------------------------------------------------------------------------------
// Compiled record.  The stem DP only ever needs values on (non-leaf x non-leaf) node pairs: a
// leaf row/column of the reference's tables is a constant (DESIGN.md "restated
// recurrence").  So each record is compiled, once per upload, into its non-leaf nodes renumbered
// level by level (level = longest path to a hairpin-closing pair), with per-node constants that
// fold the leaf rows/columns in.  All of it depends only on the record and the loop gap g.
// ------------------------------------------------------------------------------------------
struct RecDev {
  uint32_t N;       // non-leaf nodes
  uint32_t nlev;    // sub-levels: DAG levels cut into runs of at most 32/kFastRows nodes (compile_set.cpp)
  uint32_t node0;   // offset of this record in the per-node arrays
  uint32_t coff0;   // offset of its N+1 child offsets in `coff`
  uint32_t lev0;    // offset of its nlev+1 level offsets in `lev_off`
  uint32_t boff0;   // offset of its N+1 base-pair-profile offsets in `boff`
  uint32_t L;       // columns
  uint32_t col0;    // offset in the per-column arrays
  uint32_t lr;      // roots that are leaves (0 for DAGs made by the front end)
  uint32_t flags;   // REC_*
  double plr;       // sum over roots of #paths root -> any leaf
  double n_rows;    // ProfileSequence::n_seqs()
  // fast-path (separable gap factors) extras
  uint32_t c16_0;   // offset of the record's padded 16-bit child lists in `c16`
  uint32_t e4;      // entries of those lists (every node's list padded to a multiple of 4 with index N)
  uint32_t blk0;    // offset of its row blocks in `blk`
  uint32_t nblk;    // row blocks (<= kFastRows rows of one level each)
  uint32_t sub1;    // first sub-level whose nodes have inner pairs (the sub-levels before it are DAG level 0)
  uint32_t pad_;
};
enum { REC_HAS_WEIGHT = 1u, REC_SIMPLE_COLS = 2u, REC_SIMPLE_BPF = 4u,
       REC_LEN_MONOTONE = 8u,  // every non-leaf child is strictly shorter than its parent (true for front-end DAGs)
       REC_FAST = 16u };       // eligible for the separable fast kernel (see compile_set.cpp)
#ifndef STEMK_ROWS
#define STEMK_ROWS 2
#endif
constexpr uint32_t kFastRows = STEMK_ROWS;  // rows (1, 2 or 4) of one x level a warp of the fast stem kernel sweeps in lockstep
constexpr uint32_t kFastMaxN = 1024;  // largest staged record (non-leaf nodes) of the fast stem kernel

struct __attribute__((aligned(16))) XNode {   // everything the fast kernel needs about a ROW node, one 64-byte line
  double s2, a;         // g^(len-2-B), g*g*weight
  double up, ql;        // g^(B-len), sum_children e*G0(child, leaf column)
  double bfreq, paths;  // base-pair frequency, root->node paths
  uint32_t e0, e1;      // its non-leaf children in cidx (absolute)
  uint32_t len, bcode;
};

struct NodeI {        // integer part of a node for the fast kernel (8 bytes)
  uint32_t e4_bcode;  // (offset of its padded child list inside the record, in entries) << 8 | bcode
  uint16_t deg4;      // padded list length / 4
  uint16_t len;       // last - first
};

// Pointers into one device (or host) allocation holding a compiled set.
struct SetView {
  uint32_t n_recs;
  const RecDev* rec;
  // per non-leaf node (index node0 + k, k in level order)
  const double* a;       // g*g*weight                       score_table.h:26-29
  const double* el;      // sum over leaf children of g^gaps*w
  const double* ql;      // value of sum_children e * G0(child, leaf column); constant per node
  const double* paths;   // number of root->node paths (replaces the K tables)
  const double* gapt;    // gap count at the node's first column / n_rows   score_table.cpp:189-197
  const double* bfreq;   // base-pair frequency when the node has a single (a,b) entry
  const uint32_t* len;   // last - first
  const uint8_t* bcode;  // a*4+b for a single-entry profile, 0xFF otherwise
  // children (non-leaf only), CSR local to the record
  const uint32_t* coff;  // [sum(N+1)]
  const uint32_t* cidx;  // child, in the record's level numbering
  const double* ce;      // g^gaps * edge weight
  const uint32_t* lev_off;
  // separable fast path: e(j,c) = g^(len_j - len_c - 2) = s2[j] * up[c]; dn = 1/up (as its own power)
  const double* up;      // g^(B - len)   B = the record's reference length (half its longest pair)
  const double* dn;      // g^(len - B)
  const double* s2;      // g^(len - 2 - B)
  const NodeI* nodei;
  const XNode* xnode;
  const uint32_t* lperm; // per record: its non-leaf nodes sorted by length, len << 16 | node (the band of a row is a range of it)
  const uint16_t* c16;   // padded child lists as byte offsets into a row (8 * record-local node number), 8N = the all-zero dummy column
  const uint32_t* blk;   // row blocks: first row | count << 16
  // general base-pair profiles (alignments / IUPAC)
  const uint32_t* boff;  // [sum(N+1)]
  const uint8_t* bab;    // a*4+b
  const double* bfq;
  // per column
  const uint8_t* ccode;  // 0..3 one-hot base, 4 all-zero column (gap/unknown), 5 general profile
  const double* cw;      // column weight (1 when the record has none)
  const float* prof;     // 4 floats per column (A,C,G,U)
  const uint8_t* text;   // raw characters
};

constexpr int kBlobArrays = 28;   // arrays of a SetView

// Host-side result of compiling a descriptor: the record headers and what the work model, the scheduler and the
// launch shapes read.  The per-node / per-edge arrays themselves only exist in the device image, which compile_set
// writes straight into the buffer its sink hands out (blob_lay = where each array of the SetView sits in it).
struct CompiledSet {
  std::vector<RecDev> rec;
  std::vector<uint32_t> len, boff;   // per non-leaf node: last - first; base-pair-profile offsets (absolute)
  std::vector<uint8_t> text;         // raw characters per column
  size_t blob_lay[kBlobArrays] = {};
  size_t blob_bytes = 0;
  uint32_t max_E4 = 0, max_fastN = 0, n_fast = 0;  // over fast-eligible records
  uint32_t max_band_cnt = 1;                       // most nodes of one record inside any length window of 2*len_band+1 (all of them without a band)
  uint32_t n_weighted = 0, n_simple_cols = 0;      // records with per-column weights / with one-hot-or-gap columns only
  // host-only statistics for the work model and the scheduler
  std::vector<uint32_t> n_nodes_all;  // nodes incl. leaves (reference's #V)
  std::vector<uint32_t> n_edges_all;  // edges incl. leaf edges (reference's #E)
  std::vector<uint32_t> max_level_rows;
  std::vector<uint32_t> deg_all;      // per non-leaf node (level order): out-degree incl. leaf edges
  std::vector<uint32_t> band_cnt;     // per record: most nodes inside any length window of 2*len_band+1
  // work model: per record, prefix sums over node length of the out-degrees and of the base-pair-profile sizes
  // (cost_pd[cost_off[r] + k] = sum of deg over the record's non-leaf nodes with len < k), cost_off has n+1 entries
  std::vector<uint64_t> cost_off;
  std::vector<double> cost_pd, cost_pb;
  bool has_dag = false;
  uint32_t max_N = 0, max_L = 0, max_E = 0, max_nlev = 0;  // non-leaf nodes / columns / non-leaf edges / levels
};

// Builds the compiled form of every record of `desc` under loop gap `g` and length band `len_band` (0 = none).
// Returns "" or an error.
// `sink(bytes)` returns the buffer the device image is written into (NULL sink: host arrays only).
std::string compile_set(const stemk_seqset_desc& desc, double g, uint32_t len_band, int n_threads, bool timing, CompiledSet* out,
                        const std::function<char*(size_t)>& sink);

// Kernel constants derived from stemk_params on the host with libm (same exp() the reference calls).
struct KernelTables {
  double pair_tab[256];  // exp(beta*ribosum_p) or stack/covar   score_table.cpp:118-134 / :14-41
  double subst[16];      // exp(alpha*ribosum_s) or match/mismatch   string_kernel.cpp:11-34
};
void make_tables(const stemk_params& p, KernelTables* t);

inline bool kind_has_stem(int k) {
  return k == STEMK_SI_STEM || k == STEMK_SU_STEM || k == STEMK_SI_STEM_STR || k == STEMK_SU_STEM_STR ||
         k == STEMK_LSU_STEM || k == STEMK_LSU_STEM_STR;
}
inline bool kind_has_string(int k) {
  return k == STEMK_SI_STEM_STR || k == STEMK_SU_STEM_STR || k == STEMK_LSU_STR || k == STEMK_LSU_STEM_STR ||
         k == STEMK_STR_SUBST || k == STEMK_STR_SIMPLE || k == STEMK_STR_NAIVE;
}

}  // namespace stemk

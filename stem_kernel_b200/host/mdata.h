// stem_kernel_b200/host/mdata.h -- host-side data model of the drop-in.
//
// MData is the structure-annotated sequence record the reference's kernels consume
// (stem_kernel_lite/data.h:26-53: tree, seq, root, max_pa, weight), held here as flat
// SoA/CSR arrays so that it can be handed to the C ABI (include/stemk.h) without a copy
// per node.  build_mdata() is this repo's own implementation of the reference's front
// end *after* the base-pair probabilities (Profiler + DAGBuilder + find_root +
// find_max_parent + fill_weight, stem_kernel_lite/data.cpp:33-345,396-453); the
// probabilities themselves (McCaskill / ViennaRNA, common/bpmatrix.cpp) stay outside and
// arrive as sparse per-row lists.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace stemk {

// Sparse base-pair probabilities of ONE alignment row, over the row with gaps removed,
// 1-based with i < j (the indexing of BPMatrix, common/bpmatrix.h:54-62).
struct BpList {
  std::vector<uint32_t> i, j;
  std::vector<double> p;
};

struct MData {
  // tree: DAG nodes in the reference's DFS post-order (children before parents)
  std::vector<uint32_t> first, last;  // 0-based alignment columns of the pair; first==last => leaf
  std::vector<float> weight;          // P(first unpaired) * P(last unpaired)   (data.cpp:204,218)
  std::vector<uint32_t> edge_off;     // CSR over nodes, size n_nodes+1
  std::vector<uint32_t> edge_to;      // child node index
  std::vector<uint32_t> edge_gaps;    // dag.h:25-26,32
  std::vector<float> edge_weight;     // always 1 (data.cpp:207,224)
  std::vector<uint32_t> bpf_off;      // CSR over nodes, size n_nodes+1
  std::vector<uint8_t> bpf_a, bpf_b;  // base codes 0..3, ascending (a,b)       (data.cpp:246-262)
  std::vector<float> bpf_freq;
  std::vector<uint32_t> root;         // nodes without a parent, ascending      (data.cpp:396-418)
  std::vector<uint32_t> max_pa;       // largest parent index or UINT32_MAX     (data.cpp:420-435)
  // seq: profile columns [A,C,G,U,GAP] per alignment column                    (common/profile.cpp:57-73)
  std::vector<float> profile;         // L*5
  float n_rows = 0;                   // ProfileSequence::n_seqs()
  uint32_t length = 0;                // L
  std::vector<float> seq_weight;      // per-column unpaired probability; empty for sequence-only data
  std::string text;                   // first row, raw characters (used by the naive string kernel)

  uint32_t n_nodes() const { return (uint32_t)first.size(); }
  uint32_t n_edges() const { return (uint32_t)edge_to.size(); }
};

// Data(const IS&, float th, float pf_scale, const BPMatrix::Options&)  (data.cpp:324-345)
// rows: aligned rows of equal length; bp: one list per row.  Throws std::runtime_error on
// ragged rows or on a pair closer than 2 columns (the reference itself cannot represent it).
MData build_mdata(const std::vector<std::string>& rows, const std::vector<BpList>& bp, float th);

// Data(const IS&)  (data.cpp:347-352): profile only, no structure, no weights.
MData build_mdata_seqonly(const std::vector<std::string>& rows);

// common/rna.cpp:43-72: IUPAC-aware character -> code (0..3 bases, 4 gap/unknown, 5..15 IUPAC)
uint8_t char2rna(char c);

}  // namespace stemk

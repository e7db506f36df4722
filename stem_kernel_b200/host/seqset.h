// stem_kernel_b200/host/seqset.h -- flattening adapter: a vector of MData -> stemk_seqset_desc.
// The descriptor's arrays are the reference's MData fields (stem_kernel_lite/data.h:33-37)
// concatenated record after record; FlatSet owns the storage the descriptor points into.
#pragma once
#include <vector>
#include "../../include/stemk.h"
#include "mdata.h"

namespace stemk {

class FlatSet {
 public:
  FlatSet() { clear(); }
  void clear() {
    node_off_.assign(1, 0); edge_off_.assign(1, 0); bpf_off_.assign(1, 0); root_off_.assign(1, 0);
    col_off_.assign(1, 0); weight_off_.assign(1, 0);
    first_.clear(); last_.clear(); nweight_.clear(); eto_.clear(); egaps_.clear(); eweight_.clear();
    ba_.clear(); bb_.clear(); bf_.clear(); root_.clear(); profile_.clear(); nrows_.clear(); cweight_.clear();
    text_.clear();
  }
  uint32_t size() const { return (uint32_t)nrows_.size(); }

  void add(const MData& d) {
    const uint32_t e0 = (uint32_t)eto_.size(), b0 = (uint32_t)ba_.size();
    first_.insert(first_.end(), d.first.begin(), d.first.end());
    last_.insert(last_.end(), d.last.begin(), d.last.end());
    nweight_.insert(nweight_.end(), d.weight.begin(), d.weight.end());
    for (uint32_t u = 0; u < d.n_nodes(); ++u) {
      edge_off_.push_back(e0 + d.edge_off[u + 1]);
      bpf_off_.push_back(b0 + d.bpf_off[u + 1]);
    }
    eto_.insert(eto_.end(), d.edge_to.begin(), d.edge_to.end());
    egaps_.insert(egaps_.end(), d.edge_gaps.begin(), d.edge_gaps.end());
    eweight_.insert(eweight_.end(), d.edge_weight.begin(), d.edge_weight.end());
    ba_.insert(ba_.end(), d.bpf_a.begin(), d.bpf_a.end());
    bb_.insert(bb_.end(), d.bpf_b.begin(), d.bpf_b.end());
    bf_.insert(bf_.end(), d.bpf_freq.begin(), d.bpf_freq.end());
    root_.insert(root_.end(), d.root.begin(), d.root.end());
    profile_.insert(profile_.end(), d.profile.begin(), d.profile.end());
    cweight_.insert(cweight_.end(), d.seq_weight.begin(), d.seq_weight.end());
    for (uint32_t c = 0; c < d.length; ++c) text_.push_back(c < d.text.size() ? (uint8_t)d.text[c] : (uint8_t)'-');
    nrows_.push_back(d.n_rows);
    node_off_.push_back((uint32_t)first_.size());
    root_off_.push_back((uint32_t)root_.size());
    col_off_.push_back(col_off_.back() + d.length);
    weight_off_.push_back((uint32_t)cweight_.size());
  }

  stemk_seqset_desc desc() const {
    stemk_seqset_desc s;
    s.n_seqs = size();
    s.node_off = node_off_.data(); s.node_first = first_.data(); s.node_last = last_.data();
    s.node_weight = nweight_.data(); s.edge_off = edge_off_.data(); s.edge_to = eto_.data();
    s.edge_gaps = egaps_.data(); s.edge_weight = eweight_.data(); s.bpf_off = bpf_off_.data();
    s.bpf_a = ba_.data(); s.bpf_b = bb_.data(); s.bpf_freq = bf_.data(); s.root_off = root_off_.data();
    s.root = root_.data(); s.col_off = col_off_.data(); s.profile = profile_.data(); s.n_rows = nrows_.data();
    s.weight_off = weight_off_.data(); s.col_weight = cweight_.data(); s.text = text_.data();
    return s;
  }

 private:
  std::vector<uint32_t> node_off_, edge_off_, bpf_off_, root_off_, col_off_, weight_off_;
  std::vector<uint32_t> first_, last_, eto_, egaps_, root_;
  std::vector<float> nweight_, eweight_, bf_, profile_, nrows_, cweight_;
  std::vector<uint8_t> ba_, bb_, text_;
};

}  // namespace stemk

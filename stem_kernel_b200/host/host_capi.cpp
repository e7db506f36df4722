// stem_kernel_b200/host/host_capi.cpp -- C entry points of libstemk_host.so for the Python layer
// (ctypes): build MData records from rows + sparse base-pair lists, inspect them, and flatten a
// list of them into the stemk_seqset_desc that include/stemk.h's stemk_upload() takes.
// Pure host code; no CUDA here.
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "mdata.h"
#include "seqset.h"

using namespace stemk;

namespace {
thread_local std::string g_err;
void set_err(const char* m) { g_err = m; }
}  // namespace

extern "C" {

const char* stemk_host_last_error() { return g_err.c_str(); }

// rows: n_rows C strings of equal length; bp_off[n_rows+1] delimits each row's (bi,bj,bp) triples.
void* stemk_host_mdata_build(int n_rows, const char* const* rows, const uint32_t* bp_off, const uint32_t* bi,
                             const uint32_t* bj, const double* bp, float th) {
  try {
    std::vector<std::string> r(rows, rows + n_rows);
    std::vector<BpList> b(n_rows);
    for (int k = 0; k < n_rows; ++k) {
      b[k].i.assign(bi + bp_off[k], bi + bp_off[k + 1]);
      b[k].j.assign(bj + bp_off[k], bj + bp_off[k + 1]);
      b[k].p.assign(bp + bp_off[k], bp + bp_off[k + 1]);
    }
    return new MData(build_mdata(r, b, th));
  } catch (const std::exception& e) {
    set_err(e.what());
    return nullptr;
  }
}

void* stemk_host_mdata_seqonly(int n_rows, const char* const* rows) {
  try {
    std::vector<std::string> r(rows, rows + n_rows);
    return new MData(build_mdata_seqonly(r));
  } catch (const std::exception& e) {
    set_err(e.what());
    return nullptr;
  }
}

// Build many single-row records in parallel (the bench's synthetic sets): record k has row rows[k]
// and triples bp_off[k]..bp_off[k+1].  out[k] receives the handle (NULL on failure).
int stemk_host_mdata_build_many(int n, const char* const* rows, const uint64_t* bp_off, const uint32_t* bi,
                                const uint32_t* bj, const double* bp, float th, int n_threads, void** out) {
  if (n_threads < 1) n_threads = 1;
  std::vector<std::thread> th_;
  std::vector<int> bad(n_threads, 0);
  for (int t = 0; t < n_threads; ++t)
    th_.push_back(std::thread([&, t]() {
      for (int k = t; k < n; k += n_threads) {
        try {
          std::vector<std::string> r(1, rows[k]);
          std::vector<BpList> b(1);
          b[0].i.assign(bi + bp_off[k], bi + bp_off[k + 1]);
          b[0].j.assign(bj + bp_off[k], bj + bp_off[k + 1]);
          b[0].p.assign(bp + bp_off[k], bp + bp_off[k + 1]);
          out[k] = new MData(build_mdata(r, b, th));
        } catch (...) {
          out[k] = nullptr;
          bad[t] = 1;
        }
      }
    }));
  for (auto& x : th_) x.join();
  for (int t = 0; t < n_threads; ++t) if (bad[t]) { set_err("a record failed to build"); return -1; }
  return 0;
}

void stemk_host_mdata_free(void* h) { delete static_cast<MData*>(h); }

// sizes: n_nodes, n_edges, n_bpf, n_roots, length, n_weights
void stemk_host_mdata_sizes(const void* h, uint32_t* sizes) {
  const MData& d = *static_cast<const MData*>(h);
  sizes[0] = d.n_nodes(); sizes[1] = d.n_edges(); sizes[2] = (uint32_t)d.bpf_a.size();
  sizes[3] = (uint32_t)d.root.size(); sizes[4] = d.length; sizes[5] = (uint32_t)d.seq_weight.size();
}

void stemk_host_mdata_export(const void* h, uint32_t* first, uint32_t* last, float* weight, uint32_t* edge_off,
                             uint32_t* edge_to, uint32_t* edge_gaps, float* edge_w, uint32_t* bpf_off,
                             uint8_t* bpf_a, uint8_t* bpf_b, float* bpf_f, uint32_t* root, uint32_t* max_pa,
                             float* profile, float* n_rows, float* seq_weight) {
  const MData& d = *static_cast<const MData*>(h);
  auto cp = [](auto* dst, const auto& v) { if (!v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
  cp(first, d.first); cp(last, d.last); cp(weight, d.weight); cp(edge_off, d.edge_off); cp(edge_to, d.edge_to);
  cp(edge_gaps, d.edge_gaps); cp(edge_w, d.edge_weight); cp(bpf_off, d.bpf_off); cp(bpf_a, d.bpf_a);
  cp(bpf_b, d.bpf_b); cp(bpf_f, d.bpf_freq); cp(root, d.root); cp(max_pa, d.max_pa); cp(profile, d.profile);
  *n_rows = d.n_rows;
  cp(seq_weight, d.seq_weight);
}

// Assemble an MData from explicit arrays (hand-made DAGs in tests; data handed over by a caller
// that already owns reference MData objects -- see INTEGRATION.md).
void* stemk_host_mdata_from_arrays(uint32_t n_nodes, const uint32_t* first, const uint32_t* last, const float* weight,
                                   const uint32_t* edge_off, const uint32_t* edge_to, const uint32_t* edge_gaps,
                                   const float* edge_w, const uint32_t* bpf_off, const uint8_t* bpf_a,
                                   const uint8_t* bpf_b, const float* bpf_f, uint32_t n_roots, const uint32_t* root,
                                   uint32_t length, const float* profile, float n_rows, uint32_t n_weights,
                                   const float* seq_weight, const char* text) {
  MData* d = new MData;
  d->first.assign(first, first + n_nodes); d->last.assign(last, last + n_nodes);
  d->weight.assign(weight, weight + n_nodes); d->edge_off.assign(edge_off, edge_off + n_nodes + 1);
  uint32_t ne = n_nodes ? edge_off[n_nodes] : 0, nb = n_nodes ? bpf_off[n_nodes] : 0;
  d->edge_to.assign(edge_to, edge_to + ne); d->edge_gaps.assign(edge_gaps, edge_gaps + ne);
  d->edge_weight.assign(edge_w, edge_w + ne); d->bpf_off.assign(bpf_off, bpf_off + n_nodes + 1);
  d->bpf_a.assign(bpf_a, bpf_a + nb); d->bpf_b.assign(bpf_b, bpf_b + nb); d->bpf_freq.assign(bpf_f, bpf_f + nb);
  d->root.assign(root, root + n_roots);
  d->max_pa.assign(n_nodes, 0xffffffffu);
  for (uint32_t u = 0; u < n_nodes; ++u)
    for (uint32_t e = edge_off[u]; e < edge_off[u + 1]; ++e) {
      uint32_t& m = d->max_pa[edge_to[e]];
      if (m == 0xffffffffu || m < u) m = u;
    }
  d->length = length; d->profile.assign(profile, profile + (size_t)length * 5); d->n_rows = n_rows;
  d->seq_weight.assign(seq_weight, seq_weight + n_weights);
  d->text = text ? std::string(text) : std::string();
  return d;
}

// ---- flattened sets
void* stemk_host_set_new() { return new FlatSet; }
void stemk_host_set_free(void* s) { delete static_cast<FlatSet*>(s); }
void stemk_host_set_add(void* s, const void* mdata) { static_cast<FlatSet*>(s)->add(*static_cast<const MData*>(mdata)); }
uint32_t stemk_host_set_size(const void* s) { return static_cast<const FlatSet*>(s)->size(); }
void stemk_host_set_desc(const void* s, stemk_seqset_desc* out) { *out = static_cast<const FlatSet*>(s)->desc(); }

}  // extern "C"

// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin C ABI around the UNMODIFIED reference sources, compiled where they lie
// under /root/reference (see oracle/Makefile; output goes to oracle/_ref/ only).
// Nothing here re-implements the hot path: every number returned comes from the
// reference's own StemKernel / StringKernel / KernelMatrix code
// (stem_kernel_lite/{stem_kernel,string_kernel,score_table,ribosum,data}.cpp,
// common/{profile,rna}.cpp, common/kernel_matrix.{h,cpp}, stem_kernel_lite/def_kernel.h).
//
// The only substituted definitions are the ViennaRNA-facing BPMatrix
// constructors (common/bpmatrix.h:46-52; their real bodies in
// common/bpmatrix.cpp:107-139 call libRNA, which is not in this image).  The
// substitutes take base-pair probabilities from caller-supplied matrices and
// then do what bpmatrix.cpp:306-342,399-417 does with per-row matrices
// (gap index map, accumulate, average, add_matrix).
//
// Used by: tests/ (as the checker), bench.py --impl reference / cpu_baseline.

#include <cstring>
#include <cstdio>
#include <chrono>
#include <list>
#include <string>
#include <vector>
#include <sstream>
#include <stdexcept>

#include "stem_kernel_lite/def_kernel.h"
#include "stem_kernel_lite/data.h"
#include "common/kernel_matrix.h"
#include "common/fa.h"
#include "common/aln.h"
#include "common/maf.h"

// ---------------------------------------------------------------- link stubs
namespace Vienna { extern "C" void init_rand(void) {} }

// loaders declared in common/{fa,aln,maf}.h; data.cpp references them but the
// oracle never reads files.
template <> bool load_fa(std::list<std::string>&, BOOST_SPIRIT_CLASSIC_NS::file_iterator<>&) { return false; }
template <> bool load_aln(std::list<std::string>&, BOOST_SPIRIT_CLASSIC_NS::file_iterator<>&) { return false; }
template <> bool load_maf(std::list<std::string>&, BOOST_SPIRIT_CLASSIC_NS::file_iterator<>&) { return false; }

// ------------------------------------------------- bp-probability substitute
namespace {
struct BPFeed {
  int n_rows;
  const double* const* mats;  // per row: (Lr+1)*(Lr+1) row-major, 1-based (i<j used), Lr = ungapped length
  int next;
};
thread_local BPFeed* g_feed = nullptr;
}  // namespace

// per-row matrix: bpmatrix.cpp:115-121 + :141-177 with Vienna replaced by the feed
BPMatrix::BPMatrix(const std::string& s, float, const Options&)
    : sz_(s.size()), table_(sz_ + 1) {
  table_.fill(0.0);
  if (!g_feed || g_feed->next >= g_feed->n_rows) throw std::runtime_error("bp feed exhausted");
  const double* m = g_feed->mats[g_feed->next++];
  const uint n = sz_ + 1;
  for (uint j = 2; j < n; ++j)
    for (uint i = 1; i < j; ++i) table_(i, j) = m[(size_t)i * n + j];
}

// alignment / FASTA-record matrix: bpmatrix.cpp:123-128 -> :395-417 -> :306-342
BPMatrix::BPMatrix(const std::list<std::string>& ma, float pf_scale, const Options& opts)
    : sz_(ma.begin()->size()), table_(sz_ + 1) {
  table_.fill(0.0);
  const uint n_seq = ma.size();
  for (std::list<std::string>::const_iterator x = ma.begin(); x != ma.end(); ++x) {
    std::string s(*x);
    for (size_t k = 0; k < s.size(); ++k) s[k] = tolower(s[k]);
    std::string ungapped = erase_gap(s);
    boost::shared_ptr<BPMatrix> b(new BPMatrix(ungapped, pf_scale, opts));
    std::vector<uint> idxmap(s.size(), static_cast<uint>(-1));
    for (uint i = 0, j = 0; i != s.size(); ++i)
      if (s[i] != '-') idxmap[i] = j++;
    for (uint j = 1; j != sz_; ++j) {
      if (idxmap[j] == static_cast<uint>(-1)) continue;
      for (uint i = j - 1;; --i) {
        if (idxmap[i] != static_cast<uint>(-1)) table_(i + 1, j + 1) += (*b)(idxmap[i] + 1, idxmap[j] + 1);
        if (i == 0) break;
      }
    }
    add_matrix(b, idxmap);
  }
  for (uint j = 1; j != sz_; ++j)
    for (uint i = j - 1;; --i) {
      table_(i + 1, j + 1) = table_(i + 1, j + 1) / n_seq;
      if (i == 0) break;
    }
}

// --------------------------------------------------- static gap table reset
// SimpleEdgeScore keeps g^k in a process-wide static that is filled with the
// gap of whichever kernel object touched it first and is grown without a lock
// (score_table.cpp:56-77).  A harness that evaluates several loop-gap values
// in one process, or uses threads, must therefore (a) empty it between calls
// and (b) let one untimed single-threaded call size it before threads start.
// The member is private; an explicit instantiation may name it legally.
namespace {
template <class Tag, typename Tag::type M>
struct Expose { friend typename Tag::type expose(Tag) { return M; } };
struct GapVecTag { typedef std::vector<double>* type; friend type expose(GapVecTag); };
template struct Expose<GapVecTag, &SimpleEdgeScore<double, MData>::gap_vec_>;
void reset_gap_table() { expose(GapVecTag())->clear(); }
}  // namespace

// ------------------------------------------------------------------ handles
typedef std::pair<std::string, MData> Example;
typedef std::vector<Example> ExampleSet;

enum RefKind {
  REF_SI_STEM = 0,      // SiStemKernel       (def_kernel.h:12)
  REF_SU_STEM = 1,      // SuStemKernel       (def_kernel.h:36)
  REF_SI_STEM_STR = 2,  // SiStemStrKernel    (def_kernel.h:59)
  REF_SU_STEM_STR = 3,  // SuStemStrKernel    (def_kernel.h:87)
  REF_LSU_STEM = 4,     // LSuStemKernel      (def_kernel.h:114)
  REF_LSU_STR = 5,      // LSuStrKernel       (def_kernel.h:140)
  REF_LSU_STEM_STR = 6, // LSuStemStrKernel   (def_kernel.h:166)
  REF_STR_SUBST = 7,    // StringKernel(gap, alpha)             (string_kernel.cpp:11-21)
  REF_STR_SIMPLE = 8,   // StringKernel(gap, match, mismatch)   (string_kernel.cpp:24-34)
};

struct RefParams {
  int kind;
  double loop_gap, beta, stack, covar;  // stem
  double gap, alpha, match, mismatch;   // string
  unsigned len_band;
};

struct RefKernel {
  RefParams p;
};

namespace {

// Run f(kernel) with the reference kernel object that p describes.
template <class F>
void with_kernel(const RefParams& p, F& f) {
  typedef double V;
  switch (p.kind) {
    case REF_SI_STEM: { SiStemKernel<V, MData> k(p.loop_gap, p.stack, p.covar, p.len_band); f(k); break; }
    case REF_SU_STEM: { SuStemKernel<V, MData> k(p.loop_gap, p.beta, p.len_band); f(k); break; }
    case REF_SI_STEM_STR: { SiStemStrKernel<V, MData> k(p.loop_gap, p.stack, p.covar, p.gap, p.match, p.mismatch, p.len_band); f(k); break; }
    case REF_SU_STEM_STR: { SuStemStrKernel<V, MData> k(p.alpha, p.beta, p.loop_gap, p.gap, p.len_band); f(k); break; }
    case REF_LSU_STEM: { LSuStemKernel<V, MData> k(p.loop_gap, p.beta, p.len_band); f(k); break; }
    case REF_LSU_STR: { LSuStrKernel<V, MData> k(p.gap, p.alpha); f(k); break; }
    case REF_LSU_STEM_STR: { LSuStemStrKernel<V, MData> k(p.alpha, p.beta, p.loop_gap, p.gap, p.len_band); f(k); break; }
    case REF_STR_SUBST: { StringKernel<V, MData> k(p.gap, p.alpha); f(k); break; }
    case REF_STR_SIMPLE: { StringKernel<V, MData> k(p.gap, p.match, p.mismatch); f(k); break; }
    default: throw std::runtime_error("unknown kernel kind");
  }
}

struct PairOp {
  const MData& x; const MData& y; double out;
  template <class K> void operator()(const K& k) { out = k(x, y); }
};

// untimed single-threaded call on the longest sequence: sizes the static gap table
template <class K>
void prewarm(const K& k, MData* const* d, int n) {
  int best = -1;
  for (int i = 0; i < n; ++i) if (best < 0 || d[i]->seq.size() > d[best]->seq.size()) best = i;
  if (best >= 0) (void)k(*d[best], *d[best]);
}

ExampleSet make_set(int n, MData* const* d, const int* labels) {
  ExampleSet ex;
  ex.reserve(n);
  for (int i = 0; i < n; ++i) {
    char buf[32];
    if (labels) snprintf(buf, sizeof buf, "%+d", labels[i]); else snprintf(buf, sizeof buf, "0");
    ex.push_back(Example(std::string(buf), *d[i]));
  }
  return ex;
}

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct GramOp {
  const ExampleSet& ex; MData* const* dd; int n; int normalize; unsigned n_th; double* out; double secs;
  char* text; long cap; long* len;
  template <class K> void operator()(const K& k) {
    KernelMatrix<double> m;
    reset_gap_table();
    prewarm(k, dd, n);
    double t0 = now_s();
    m.calculate(ex, k, normalize != 0, n_th);
    secs = now_s() - t0;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) out[(size_t)i * n + j] = m(i, j);
    if (text) {
      std::ostringstream os; m.print(os);
      std::string s = os.str();
      *len = s.size();
      if ((long)s.size() <= cap) memcpy(text, s.data(), s.size());
    }
  }
};

struct CrossOp {
  const ExampleSet& te; const ExampleSet& tr; MData* const* ted; MData* const* trd;
  int norm_test, normalize; unsigned n_th; double* out; double* self_out; double secs;
  template <class K> void operator()(const K& k) {
    // the reference sizes self_ by train.size() but indexes it by test index
    // (kernel_matrix.cpp:713,721): only call with n_test <= n_train.
    KernelMatrix<double> m(te.size(), tr.size());
    reset_gap_table();
    prewarm(k, trd, tr.size()); prewarm(k, ted, te.size());
    double t0 = now_s();
    m.calculate(te, tr, k, norm_test != 0, normalize != 0, n_th);
    secs = now_s() - t0;
    for (size_t i = 0; i < te.size(); ++i) for (size_t j = 0; j < tr.size(); ++j) out[i * tr.size() + j] = m(i, j);
    if (self_out && (norm_test || normalize)) for (size_t i = 0; i < te.size(); ++i) self_out[i] = m(i);
  }
};

struct RowOp {
  const Example& te; const ExampleSet& tr; MData* const* trd; const std::vector<uint>& sv; unsigned n_th;
  double* out; double* self_out; double secs;
  template <class K> void operator()(const K& k) {
    std::vector<double> v(out, out + tr.size());
    reset_gap_table();
    prewarm(k, trd, tr.size()); (void)k(te.second, te.second);
    double t0 = now_s();
    KernelMatrix<double>::calculate(v, te, tr, sv, k, n_th, self_out);
    secs = now_s() - t0;
    for (size_t j = 0; j < tr.size(); ++j) out[j] = v[j];
  }
};

struct DiagOp {
  const ExampleSet& tr; MData* const* trd; const std::vector<uint>& sv; unsigned n_th; double* out; double secs;
  template <class K> void operator()(const K& k) {
    std::vector<double> v(out, out + tr.size());
    reset_gap_table();
    prewarm(k, trd, tr.size());
    double t0 = now_s();
    KernelMatrix<double>::diagonal(v, tr, sv, k, n_th);
    secs = now_s() - t0;
    for (size_t j = 0; j < tr.size(); ++j) out[j] = v[j];
  }
};

struct PairsOp {
  MData* const* dd; int n; int n_pairs; const int* pi; const int* pj; unsigned n_th; double* out; double secs;
  template <class K> void operator()(const K& k) {
    reset_gap_table();
    prewarm(k, dd, n);
    double t0 = now_s();
    std::vector<std::thread> th;
    for (unsigned t = 0; t < n_th; ++t)
      th.push_back(std::thread([this, &k, t]() {
        K kk(k);  // the reference hands each thread a copy of the functor (kernel_matrix.cpp:542)
        for (int q = t; q < n_pairs; q += n_th) out[q] = kk(*dd[pi[q]], *dd[pj[q]]);
      }));
    for (size_t t = 0; t < th.size(); ++t) th[t].join();
    secs = now_s() - t0;
  }
};

}  // namespace

extern "C" {

// Build MData through the reference's own constructor (data.cpp:324-345).
void* ref_mdata_new(int n_rows, const char* const* rows, const double* const* bp_rows, float th) {
  try {
    std::list<std::string> ma;
    for (int r = 0; r < n_rows; ++r) ma.push_back(rows[r]);
    BPFeed feed = {n_rows, bp_rows, 0};
    g_feed = &feed;
    BPMatrix::Options opts;
    MData* d = new MData(ma, th, -1.0f, opts);
    g_feed = nullptr;
    return d;
  } catch (...) {
    g_feed = nullptr;
    return nullptr;
  }
}

// MData without structure (data.cpp:347-352), as the string-only loaders make.
void* ref_mdata_new_seqonly(int n_rows, const char* const* rows) {
  std::list<std::string> ma;
  for (int r = 0; r < n_rows; ++r) ma.push_back(rows[r]);
  return new MData(ma);
}

// MData assembled from explicit arrays (MData is a plain struct, data.h:26-53);
// lets tests feed hand-made DAGs (empty DAG, single hairpin, ...).
// edge_ppos/edge_cpos: 2 uints per edge; a loop edge (to a leaf) has cpos == (p.first,p.first).
void* ref_mdata_from_flat(int n_nodes, const unsigned* first, const unsigned* last, const float* weight,
                          const unsigned* edge_off, const unsigned* edge_to, const unsigned* edge_ppos,
                          const unsigned* edge_cpos, const float* edge_w, const unsigned* bpf_off,
                          const unsigned char* bpf_a, const unsigned char* bpf_b, const float* bpf_f, int L,
                          const float* seq_weight, int n_rows, const char* const* rows) {
  std::list<std::string> ma;
  for (int r = 0; r < n_rows; ++r) ma.push_back(rows[r]);
  MData* d = new MData(ma);
  d->tree.clear();
  for (int i = 0; i < n_nodes; ++i) {
    std::list<DAG::bp_freq_t> bpf;
    for (unsigned k = bpf_off[i]; k < bpf_off[i + 1]; ++k)
      bpf.push_back(std::make_pair(std::make_pair(bpf_a[k], bpf_b[k]), bpf_f[k]));
    MData::Node node(first[i], last[i], weight[i], bpf, edge_off[i + 1] - edge_off[i]);
    for (unsigned k = edge_off[i]; k < edge_off[i + 1]; ++k) {
      Pos pp(edge_ppos[2 * k], edge_ppos[2 * k + 1]), cp(edge_cpos[2 * k], edge_cpos[2 * k + 1]);
      if (cp.first == cp.second) node[k - edge_off[i]] = MData::Edge(edge_to[k], pp, edge_w[k]);
      else node[k - edge_off[i]] = MData::Edge(edge_to[k], pp, cp, edge_w[k]);
    }
    d->tree.push_back(node);
  }
  // root / max_pa exactly as data.cpp:396-435 define them
  std::vector<bool> is_root(n_nodes, true);
  d->max_pa.assign(n_nodes, static_cast<uint>(-1));
  for (int i = 0; i < n_nodes; ++i)
    for (unsigned k = edge_off[i]; k < edge_off[i + 1]; ++k) {
      is_root[edge_to[k]] = false;
      uint& m = d->max_pa[edge_to[k]];
      if (m == static_cast<uint>(-1) || m < (uint)i) m = i;
    }
  d->root.clear();
  for (int i = 0; i < n_nodes; ++i) if (is_root[i]) d->root.push_back(i);
  d->weight.clear();
  if (seq_weight) d->weight.assign(seq_weight, seq_weight + L);
  return d;
}

void ref_mdata_free(void* h) { delete static_cast<MData*>(h); }

void ref_mdata_counts(void* h, int* n_nodes, int* n_edges, int* n_bpf, int* n_root, int* L, int* n_weight) {
  const MData& d = *static_cast<MData*>(h);
  int e = 0, b = 0;
  for (size_t i = 0; i < d.tree.size(); ++i) {
    e += d.tree[i].size();
    for (DAG::bp_freq_iterator it = d.tree[i].bp_freq_begin(); it != d.tree[i].bp_freq_end(); ++it) ++b;
  }
  *n_nodes = d.tree.size(); *n_edges = e; *n_bpf = b; *n_root = d.root.size(); *L = d.seq.size();
  *n_weight = d.weight.size();
}

// Flattened dump of the reference's MData (the GPU input contract).
void ref_mdata_dump(void* h, unsigned* first, unsigned* last, float* weight, unsigned* edge_off,
                    unsigned* edge_to, unsigned* edge_gaps, float* edge_w, unsigned* edge_ppos,
                    unsigned* edge_cpos, unsigned* bpf_off, unsigned char* bpf_a, unsigned char* bpf_b,
                    float* bpf_f, unsigned* root, unsigned* max_pa, float* profile, float* n_seqs,
                    float* seq_weight) {
  const MData& d = *static_cast<MData*>(h);
  unsigned e = 0, b = 0;
  for (size_t i = 0; i < d.tree.size(); ++i) {
    const MData::Node& n = d.tree[i];
    first[i] = n.first(); last[i] = n.last(); weight[i] = n.weight();
    edge_off[i] = e; bpf_off[i] = b;
    for (MData::Node::const_iterator it = n.begin(); it != n.end(); ++it, ++e) {
      edge_to[e] = it->to(); edge_gaps[e] = it->gaps(); edge_w[e] = it->weight();
      edge_ppos[2 * e] = it->p_pos().first; edge_ppos[2 * e + 1] = it->p_pos().second;
      edge_cpos[2 * e] = it->c_pos().first; edge_cpos[2 * e + 1] = it->c_pos().second;
    }
    for (DAG::bp_freq_iterator it = n.bp_freq_begin(); it != n.bp_freq_end(); ++it, ++b) {
      bpf_a[b] = it->first.first; bpf_b[b] = it->first.second; bpf_f[b] = it->second;
    }
  }
  edge_off[d.tree.size()] = e; bpf_off[d.tree.size()] = b;
  for (size_t i = 0; i < d.root.size(); ++i) root[i] = d.root[i];
  for (size_t i = 0; i < d.max_pa.size(); ++i) max_pa[i] = d.max_pa[i];
  for (unsigned i = 0; i < d.seq.size(); ++i)
    for (unsigned k = 0; k < 5; ++k) profile[5 * i + k] = d.seq[i][k];
  *n_seqs = d.seq.n_seqs();
  for (size_t i = 0; i < d.weight.size(); ++i) seq_weight[i] = d.weight[i];
}

void* ref_kernel_new(const RefParams* p) { RefKernel* k = new RefKernel; k->p = *p; return k; }
void ref_kernel_free(void* k) { delete static_cast<RefKernel*>(k); }

double ref_kernel_pair(void* kh, void* xh, void* yh) {
  reset_gap_table();
  PairOp op = {*static_cast<MData*>(xh), *static_cast<MData*>(yh), 0.0};
  with_kernel(static_cast<RefKernel*>(kh)->p, op);
  return op.out;
}

// Square Gram matrix through KernelMatrix::calculate (kernel_matrix.cpp:485-575).
// Returns wall seconds of the calculate() call (the reference's own `elapsed`
// reads 0 under threads).  If text is given, also renders
// KernelMatrix::print (kernel_matrix.cpp:756-770) into it.
double ref_gram(void* kh, int n, void* const* d, const int* labels, int normalize, unsigned n_th, double* out,
                char* text, long text_cap, long* text_len) {
  ExampleSet ex = make_set(n, reinterpret_cast<MData* const*>(d), labels);
  GramOp op = {ex, reinterpret_cast<MData* const*>(d), n, normalize, n_th, out, 0.0, text, text_cap, text_len};
  with_kernel(static_cast<RefKernel*>(kh)->p, op);
  return op.secs;
}

// Rectangular test x train (kernel_matrix.cpp:699-754).  self_out (n_test) gets
// KernelMatrix::self() when norm_test||normalize.
double ref_cross(void* kh, int n_test, void* const* test, int n_train, void* const* train, int norm_test,
                 int normalize, unsigned n_th, double* out, double* self_out) {
  ExampleSet te = make_set(n_test, reinterpret_cast<MData* const*>(test), nullptr);
  ExampleSet tr = make_set(n_train, reinterpret_cast<MData* const*>(train), nullptr);
  CrossOp op = {te, tr, reinterpret_cast<MData* const*>(test), reinterpret_cast<MData* const*>(train),
                norm_test, normalize, n_th, out, self_out, 0.0};
  with_kernel(static_cast<RefKernel*>(kh)->p, op);
  return op.secs;
}

// One test row (kernel_matrix.cpp:635-697); entries outside sv_index are left as
// the caller initialised them.
double ref_row(void* kh, void* test, int n_train, void* const* train, const unsigned* sv_index, int n_sv,
               unsigned n_th, double* out, double* self_out) {
  ExampleSet tr = make_set(n_train, reinterpret_cast<MData* const*>(train), nullptr);
  Example te(std::string("0"), *static_cast<MData*>(test));
  std::vector<uint> sv(sv_index, sv_index + n_sv);
  RowOp op = {te, tr, reinterpret_cast<MData* const*>(train), sv, n_th, out, self_out, 0.0};
  with_kernel(static_cast<RefKernel*>(kh)->p, op);
  return op.secs;
}

// Diagonal (kernel_matrix.cpp:578-633).
double ref_diag(void* kh, int n_train, void* const* train, const unsigned* sv_index, int n_sv, unsigned n_th,
                double* out) {
  ExampleSet tr = make_set(n_train, reinterpret_cast<MData* const*>(train), nullptr);
  std::vector<uint> sv(sv_index, sv_index + n_sv);
  DiagOp op = {tr, reinterpret_cast<MData* const*>(train), sv, n_th, out, 0.0};
  with_kernel(static_cast<RefKernel*>(kh)->p, op);
  return op.secs;
}

// Timed evaluation of an explicit pair list with n_th threads, pairs dealt
// round-robin like CalcTrainMatrix (kernel_matrix.cpp:42-56): the bounded
// CPU-baseline sample of the big configs.
double ref_pairs_timed(void* kh, int n, void* const* d, int n_pairs, const int* pi, const int* pj, unsigned n_th,
                       double* out) {
  PairsOp op = {reinterpret_cast<MData* const*>(d), n, n_pairs, pi, pj, n_th, out, 0.0};
  with_kernel(static_cast<RefKernel*>(kh)->p, op);
  return op.secs;
}

}  // extern "C"

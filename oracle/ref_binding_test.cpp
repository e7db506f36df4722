// oracle/ref_binding_test.cpp -- TEST INFRASTRUCTURE: the reference-side binding of INTEGRATION.md, COMPILED.
//
// The drop-in claim of include/stemk.h is that a maintainer of the reference adds a few lines -- a kernel type that
// carries its parameters and full specialisations of KernelMatrix<double>'s members for it -- and links
// libstemk_b200.so; everything else (Data / DAGBuilder, KernelMatrix::print, LIBSVM, the optimizer) stays as it is.
// This file is exactly that binding, built against the UNMODIFIED reference headers and sources (the same
// translation units and shim headers as oracle/_ref/libstemk_ref.so, see oracle/Makefile), so that a test can run
//     KernelMatrix<double>::calculate(train, ReferenceKernel, ...)          (the reference's own CPU code)
//     KernelMatrix<double>::calculate(train, GpuBound<ReferenceKernel>, ...)   (this binding -> the C ABI -> the GPU)
// on the same reference-built MData and compare matrices and KernelMatrix::print text (tests/test_ref_binding.py).
//
// Reference interfaces replaced: common/kernel_matrix.h:67-107 (calculate square / rectangular / one row, diagonal),
// bodies common/kernel_matrix.cpp:485-575, 635-754, 578-633; kernel classes stem_kernel_lite/def_kernel.h:12-192
// (their parameters are private, so the bound type keeps a copy -- main.cpp:180-215 has them where it builds the
// kernel object).
#include "ref_harness.cpp"   // the reference translation units' harness: MData builders, kernel dispatch, ExampleSet

#include "../include/stemk.h"

namespace {

// ---- what the maintainer adds -------------------------------------------------------------------------------
// The reference functor (kept: operator() still answers single pairs on the CPU) + the parameters it was built from.
template <class K>
struct GpuBound {
  K ref;
  stemk_params p;
  int device;
  typename K::value_type operator()(const MData& x, const MData& y) const { return ref(x, y); }
};

// vector<MData> -> stemk_seqset_desc: the fields of MData (stem_kernel_lite/data.h:33-37) concatenated
struct FlatSet {
  std::vector<uint32_t> node_off{0}, first, last, edge_off{0}, edge_to, edge_gaps, bpf_off{0}, root_off{0}, root, col_off{0},
      weight_off{0};
  std::vector<float> node_w, edge_w, bpf_f, profile, n_rows, col_w;
  std::vector<uint8_t> bpf_a, bpf_b, text;
  void add(const MData& d) {
    const uint32_t e0 = (uint32_t)edge_to.size(), b0 = (uint32_t)bpf_a.size();
    (void)e0; (void)b0;
    for (std::vector<MData::Node>::const_iterator n = d.tree.begin(); n != d.tree.end(); ++n) {   // dag.h:67-157, children first
      first.push_back(n->first()); last.push_back(n->last()); node_w.push_back(n->weight());
      for (MData::Node::const_iterator e = n->begin(); e != n->end(); ++e) {                      // dag.h:17-65
        edge_to.push_back(e->to()); edge_gaps.push_back(e->gaps()); edge_w.push_back(e->weight());
      }
      edge_off.push_back((uint32_t)edge_to.size());
      for (DAG::bp_freq_iterator bf = n->bp_freq_begin(); bf != n->bp_freq_end(); ++bf) { // ((a,b), freq), dag.h:128-129
        bpf_a.push_back((uint8_t)bf->first.first); bpf_b.push_back((uint8_t)bf->first.second); bpf_f.push_back(bf->second);
      }
      bpf_off.push_back((uint32_t)bpf_a.size());
    }
    node_off.push_back((uint32_t)first.size());
    root.insert(root.end(), d.root.begin(), d.root.end()); root_off.push_back((uint32_t)root.size());
    for (uint i = 0; i != d.seq.size(); ++i) {                                                  // profile.h:15-16
      for (uint k = 0; k != 5; ++k) profile.push_back(d.seq[i][k]);
      text.push_back('n');
    }
    col_off.push_back(col_off.back() + (uint32_t)d.seq.size()); n_rows.push_back(d.seq.n_seqs());
    col_w.insert(col_w.end(), d.weight.begin(), d.weight.end()); weight_off.push_back((uint32_t)col_w.size());
  }
  stemk_seqset_desc desc() const {
    stemk_seqset_desc s;
    s.n_seqs = (uint32_t)n_rows.size();
    s.node_off = node_off.data(); s.node_first = first.data(); s.node_last = last.data(); s.node_weight = node_w.data();
    s.edge_off = edge_off.data(); s.edge_to = edge_to.data(); s.edge_gaps = edge_gaps.data(); s.edge_weight = edge_w.data();
    s.bpf_off = bpf_off.data(); s.bpf_a = bpf_a.data(); s.bpf_b = bpf_b.data(); s.bpf_freq = bpf_f.data();
    s.root_off = root_off.data(); s.root = root.data(); s.col_off = col_off.data(); s.profile = profile.data();
    s.n_rows = n_rows.data(); s.weight_off = weight_off.data(); s.col_weight = col_w.data(); s.text = text.data();
    return s;
  }
};

struct Gpu {   // context + uploaded sets of one call; throws the library's message
  stemk_ctx* ctx = nullptr;
  std::vector<stemk_set*> sets;
  Gpu(const stemk_params& p, int device) { if (stemk_create(&ctx, &p, device) != STEMK_OK) throw std::runtime_error(stemk_last_error(NULL)); }
  stemk_set* upload(const ExampleSet& ex) {
    FlatSet flat;
    for (ExampleSet::const_iterator x = ex.begin(); x != ex.end(); ++x) flat.add(x->second);
    const stemk_seqset_desc d = flat.desc();
    stemk_set* s = nullptr;
    check(stemk_upload(ctx, &d, &s));
    sets.push_back(s);
    return s;
  }
  void check(int rc) { if (rc != STEMK_OK) throw std::runtime_error(stemk_last_error(ctx)); }
  ~Gpu() { for (stemk_set* s : sets) stemk_set_free(ctx, s); stemk_destroy(ctx); }
};

}  // namespace

// KernelMatrix::calculate(train, kernel, normalize) -- kernel_matrix.cpp:485-575
#define STEMK_BIND(KERNEL)                                                                                         \
  template <> template <>                                                                                           \
  double KernelMatrix<double>::calculate(const ExampleSet& train, const GpuBound<KERNEL>& k, bool normalize, uint) {  \
    Gpu g(k.p, k.device);                                                                                           \
    stemk_set* s = g.upload(train);                                                                                 \
    const uint n = train.size();                                                                                    \
    resize(n, n);                                                                                                   \
    for (uint i = 0; i != n; ++i) label_[i] = train[i].first;                                                        \
    g.check(stemk_gram(g.ctx, s, normalize ? 1 : 0, matrix_.data()));                                                \
    return 0.0;   /* `elapsed` */                                                                                   \
  }                                                                                                                 \
  /* KernelMatrix::calculate(test, train, kernel, norm_test, normalize) -- kernel_matrix.cpp:699-754 */              \
  template <> template <>                                                                                           \
  double KernelMatrix<double>::calculate(const ExampleSet& test, const ExampleSet& train, const GpuBound<KERNEL>& k, \
                                         bool norm_test, bool normalize, uint) {                                    \
    Gpu g(k.p, k.device);                                                                                           \
    stemk_set* tr = g.upload(train);                                                                                \
    stemk_set* te = g.upload(test);                                                                                 \
    const uint nt = test.size(), n = train.size();                                                                  \
    resize(nt, n);                                                                                                  \
    self_.resize(nt);                                                                                               \
    for (uint i = 0; i != nt; ++i) label_[i] = test[i].first;                                                        \
    g.check(stemk_cross(g.ctx, te, tr, NULL, 0, normalize ? 1 : 0, matrix_.data(), (norm_test || normalize) ? self_.data() : NULL)); \
    return 0.0;                                                                                                     \
  }                                                                                                                 \
  /* KernelMatrix::diagonal -- kernel_matrix.cpp:578-633 */                                                          \
  template <> template <>                                                                                           \
  double KernelMatrix<double>::diagonal(std::vector<double>& diag, const ExampleSet& train, const std::vector<uint>& sv_index, \
                                        const GpuBound<KERNEL>& k, uint) {                                          \
    Gpu g(k.p, k.device);                                                                                           \
    stemk_set* tr = g.upload(train);                                                                                \
    if (diag.size() != train.size()) diag.resize(train.size());   /* the reference leaves a sized vector as it is */    \
    std::vector<uint32_t> sv(sv_index.begin(), sv_index.end());                                                      \
    g.check(stemk_diag(g.ctx, tr, sv.empty() ? NULL : sv.data(), (uint32_t)sv.size(), diag.data()));                  \
    return 0.0;                                                                                                     \
  }

typedef SuStemKernel<double, MData> RefSuStem;
typedef SuStemStrKernel<double, MData> RefSuStemStr;
typedef SiStemStrKernel<double, MData> RefSiStemStr;
typedef StringKernel<double, MData> RefString;
STEMK_BIND(RefSuStem)
STEMK_BIND(RefSuStemStr)
STEMK_BIND(RefSiStemStr)
STEMK_BIND(RefString)

// ---- the test's entry points ----------------------------------------------------------------------------------
namespace {

stemk_params to_stemk(const RefParams& p) {
  stemk_params q;
  memset(&q, 0, sizeof(q));
  q.kind = p.kind; q.len_band = p.len_band; q.loop_gap = p.loop_gap; q.beta = p.beta; q.stack = p.stack; q.covar = p.covar;
  q.gap = p.gap; q.alpha = p.alpha; q.match = p.match; q.mismatch = p.mismatch;
  return q;
}

struct BoundGram {
  const ExampleSet& ex; const RefParams& rp; int normalize, device; double* out; char* text; long cap; long* len;
  template <class K> void run(const K& ref) {
    GpuBound<K> k = {ref, to_stemk(rp), device};
    KernelMatrix<double> m;
    m.calculate(ex, k, normalize != 0, 1);      // overload resolution lands on the specialisation above
    const int n = (int)ex.size();
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) out[(size_t)i * n + j] = m(i, j);
    if (text) {
      std::ostringstream os; m.print(os);       // the reference's own KernelMatrix::print on the bound matrix
      const std::string s = os.str();
      *len = (long)s.size();
      if ((long)s.size() <= cap) memcpy(text, s.data(), s.size());
    }
  }
};

struct BoundCross {
  const ExampleSet& te; const ExampleSet& tr; const RefParams& rp; int norm_test, normalize, device; double* out; double* self_out;
  template <class K> void run(const K& ref) {
    GpuBound<K> k = {ref, to_stemk(rp), device};
    KernelMatrix<double> m((uint)te.size(), (uint)tr.size());
    m.calculate(te, tr, k, norm_test != 0, normalize != 0, 1);
    for (size_t i = 0; i < te.size(); ++i) for (size_t j = 0; j < tr.size(); ++j) out[i * tr.size() + j] = m(i, j);
    if (self_out && (norm_test || normalize)) for (size_t i = 0; i < te.size(); ++i) self_out[i] = m.self()[i];
  }
};

struct BoundDiag {
  const ExampleSet& tr; const RefParams& rp; const std::vector<uint>& sv; int device; double* out;
  template <class K> void run(const K& ref) {
    GpuBound<K> k = {ref, to_stemk(rp), device};
    std::vector<double> v(out, out + tr.size());   // entries outside sv_index keep the caller's values
    KernelMatrix<double>::diagonal(v, tr, sv, k, 1);
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
  }
};

template <class Op>
int bound_dispatch(const RefParams& p, Op& op, char* err, long errcap) {
  try {
    switch (p.kind) {
      case REF_SU_STEM: { RefSuStem k(p.loop_gap, p.beta, p.len_band); op.run(k); break; }
      case REF_SU_STEM_STR: { RefSuStemStr k(p.alpha, p.beta, p.loop_gap, p.gap, p.len_band); op.run(k); break; }
      case REF_SI_STEM_STR: { RefSiStemStr k(p.loop_gap, p.stack, p.covar, p.gap, p.match, p.mismatch, p.len_band); op.run(k); break; }
      case REF_STR_SUBST: { RefString k(p.gap, p.alpha); op.run(k); break; }
      default: throw std::runtime_error("binding test: kernel kind not bound");
    }
  } catch (const std::exception& ex) {
    if (err && errcap > 0) { strncpy(err, ex.what(), (size_t)errcap - 1); err[errcap - 1] = 0; }
    return 1;
  }
  return 0;
}

}  // namespace

extern "C" {

// KernelMatrix<double>::calculate(train, GpuBound<K>, normalize) + KernelMatrix::print
int refbind_gram(void* kh, int n, void* const* d, const int* labels, int normalize, int device, double* out, char* text, long cap,
                 long* len, char* err, long errcap) {
  const RefKernel* k = static_cast<const RefKernel*>(kh);
  const ExampleSet ex = make_set(n, reinterpret_cast<MData* const*>(d), labels);
  BoundGram op = {ex, k->p, normalize, device, out, text, cap, len};
  return bound_dispatch(k->p, op, err, errcap);
}

int refbind_cross(void* kh, int n_test, void* const* test, int n_train, void* const* train, int norm_test, int normalize, int device,
                  double* out, double* self_out, char* err, long errcap) {
  const RefKernel* k = static_cast<const RefKernel*>(kh);
  const ExampleSet te = make_set(n_test, reinterpret_cast<MData* const*>(test), NULL);
  const ExampleSet tr = make_set(n_train, reinterpret_cast<MData* const*>(train), NULL);
  BoundCross op = {te, tr, k->p, norm_test, normalize, device, out, self_out};
  return bound_dispatch(k->p, op, err, errcap);
}

int refbind_diag(void* kh, int n_train, void* const* train, const unsigned* sv_index, int n_sv, int device, double* out, char* err,
                 long errcap) {
  const RefKernel* k = static_cast<const RefKernel*>(kh);
  const ExampleSet tr = make_set(n_train, reinterpret_cast<MData* const*>(train), NULL);
  const std::vector<uint> sv(sv_index, sv_index + n_sv);
  BoundDiag op = {tr, k->p, sv, device, out};
  return bound_dispatch(k->p, op, err, errcap);
}

}  // extern "C"

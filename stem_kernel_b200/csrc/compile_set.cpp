// stem_kernel_b200/csrc/compile_set.cpp -- host side of stemk_upload(): turns the flattened MData
// records of a stemk_seqset_desc into the per-record tables the CUDA kernels read (RecDev /
// SetView in stemk_internal.h).  Pure host C++ (no CUDA), O(nodes + edges + columns) per record.
//
// What is folded in here, and which reference lines it comes from:
//   a      = g*g*node_weight                         node_score(xx,i), score_table.h:26-29
//   ce     = g^gaps * edge_weight, g^k by repeated multiplication like SimpleEdgeScore::initialize
//            (score_table.cpp:61-77) so the powers are the reference's bit for bit
//   el,ql  = what the leaf rows/columns of the reference's G0 table contribute: G0(leaf,leaf)=1,
//            G0(i,leaf)=a_i*ql_i, G0(leaf,j)=0 for non-leaf j   (stem_kernel.cpp:39-42,62-77)
//   paths  = number of root->node paths; sum_roots K0 == sum_ij paths_x(i) paths_y(j) MATCH(i,j)
//            because the K recursion (stem_kernel.cpp:56,65,75) only counts paths
//   levels = longest-path layering of the non-leaf nodes (rows of one level are independent)
#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <functional>
#include <thread>

#include "ribosum85_60.inc"
#include "stemk_internal.h"

namespace stemk {

void make_tables(const stemk_params& p, KernelTables* t) {
  const bool stem_simple = p.kind == STEMK_SI_STEM || p.kind == STEMK_SI_STEM_STR;
  const bool str_simple = p.kind == STEMK_SI_STEM_STR || p.kind == STEMK_STR_SIMPLE;
  for (int ab = 0; ab < 16; ++ab)
    for (int cd = 0; cd < 16; ++cd)
      t->pair_tab[ab * 16 + cd] =
          stem_simple ? (ab == cd ? p.stack : p.covar) : std::exp(kRibosumPair[ab * 16 + cd] * p.beta);
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b)
      t->subst[a * 4 + b] = str_simple ? (a == b ? p.match : p.mismatch) : std::exp(kRibosumSingle[a * 4 + b] * p.alpha);
}

namespace {

struct RecOut {  // one record's share, appended to the CompiledSet in record order afterwards
  RecDev hdr;
  std::vector<double> a, el, ql, paths, gapt, bfreq, ce, bfq, cw;
  std::vector<uint32_t> len, coff, cidx, lev_off, boff;
  std::vector<uint8_t> bcode, bab, ccode, text;
  std::vector<float> prof;
  std::vector<uint32_t> deg_all;
  std::vector<double> up, dn, s2;
  std::vector<NodeI> nodei;
  std::vector<uint16_t> c16;
  std::vector<uint32_t> blk, lperm;
  std::vector<double> pd, pb;  // work-model prefix sums over node length
  uint32_t n_all = 0, e_all = 0, max_rows = 0, band_cnt = 0;
  std::string err;
  void swap_arrays(RecOut& o) {   // takes over (and with a fresh object: releases) o's arrays
    a.swap(o.a); el.swap(o.el); ql.swap(o.ql); paths.swap(o.paths); gapt.swap(o.gapt); bfreq.swap(o.bfreq); ce.swap(o.ce); bfq.swap(o.bfq);
    cw.swap(o.cw); len.swap(o.len); coff.swap(o.coff); cidx.swap(o.cidx); lev_off.swap(o.lev_off); boff.swap(o.boff); bcode.swap(o.bcode);
    bab.swap(o.bab); ccode.swap(o.ccode); text.swap(o.text); prof.swap(o.prof); deg_all.swap(o.deg_all); up.swap(o.up); dn.swap(o.dn);
    s2.swap(o.s2); nodei.swap(o.nodei); c16.swap(o.c16); blk.swap(o.blk); lperm.swap(o.lperm); pd.swap(o.pd); pb.swap(o.pb);
  }
};

void compile_record(const stemk_seqset_desc& s, uint32_t r, double g, uint32_t len_band, RecOut* o) {
  const uint32_t n0 = s.node_off[r], n = s.node_off[r + 1] - n0;
  const uint32_t* first = s.node_first + n0;
  const uint32_t* last = s.node_last + n0;
  const float* w = s.node_weight + n0;
  const uint32_t* eoff = s.edge_off + n0;
  const uint32_t* boff = s.bpf_off + n0;
  const uint32_t c0 = s.col_off[r], L = s.col_off[r + 1] - c0;
  const float* prof = s.profile + (size_t)5 * c0;
  const float nrows = s.n_rows[r];
  RecDev& h = o->hdr;
  h = RecDev();
  h.L = L;
  h.n_rows = nrows;
  o->n_all = n;
  // descriptor sanity first: everything below sizes arrays from these numbers
  if (s.node_off[r + 1] < n0 || s.col_off[r + 1] < c0) { o->err = "record offsets not monotone"; return; }
  if (!(nrows > 0.0f)) { o->err = "n_rows must be positive"; return; }
  for (uint32_t u = 0; u < n; ++u) {
    if (last[u] < first[u] || last[u] >= L) { o->err = "node positions must satisfy first <= last < sequence length"; return; }
    if (eoff[u + 1] < eoff[u]) { o->err = "edge offsets not monotone"; return; }
    if (boff[u + 1] < boff[u]) { o->err = "base-pair-profile offsets not monotone"; return; }
  }
  for (uint32_t e = n ? eoff[0] : 0; n && e < eoff[n]; ++e)
    if (s.edge_gaps[e] > L) { o->err = "edge gap count exceeds the sequence length"; return; }
  o->e_all = n ? eoff[n] - eoff[0] : 0;

  // ---- columns
  o->ccode.resize(L); o->cw.resize(L); o->prof.resize((size_t)4 * L); o->text.resize(L);
  const uint32_t w0 = s.weight_off[r], nw = s.weight_off[r + 1] - w0;
  if (nw != 0 && nw != L) { o->err = "weight vector length differs from the sequence length"; return; }
  bool simple_cols = true;
  for (uint32_t c = 0; c < L; ++c) {
    const float* p = prof + 5 * c;
    int ones = 0, zeros = 0, which = 0;
    for (int k = 0; k < 4; ++k) {
      if (p[k] == 1.0f) { ++ones; which = k; }
      if (p[k] == 0.0f) ++zeros;
      o->prof[4 * c + k] = p[k];
    }
    uint8_t code = 5;
    if (ones == 1 && zeros == 3) code = (uint8_t)which;
    else if (zeros == 4) code = 4;
    if (code == 5) simple_cols = false;
    o->ccode[c] = code;
    o->cw[c] = nw ? (double)s.col_weight[w0 + c] : 1.0;
    o->text[c] = s.text ? s.text[c0 + c] : 0;
  }
  h.flags = (nw ? REC_HAS_WEIGHT : 0u) | (simple_cols ? REC_SIMPLE_COLS : 0u);

  // ---- DAG
  if (n == 0) {
    h.N = 0; h.nlev = 0; h.flags |= REC_SIMPLE_BPF | REC_LEN_MONOTONE;
    o->coff.assign(1, 0); o->lev_off.assign(1, 0); o->boff.assign(1, 0);
    return;
  }
  uint32_t max_gaps = 0;
  for (uint32_t u = 0; u < n; ++u) {
    if (eoff[u + 1] < eoff[u]) { o->err = "edge offsets not monotone"; return; }
    for (uint32_t e = eoff[u]; e < eoff[u + 1]; ++e) {
      if (s.edge_to[e] >= u) { o->err = "DAG nodes must list children before parents"; return; }
      max_gaps = std::max(max_gaps, s.edge_gaps[e]);
    }
  }
  std::vector<double> gpow(max_gaps + 1);
  gpow[0] = 1.0;
  for (uint32_t k = 1; k <= max_gaps; ++k) gpow[k] = gpow[k - 1] * g;

  std::vector<char> leaf(n);
  std::vector<uint32_t> level(n, 0), newidx(n, 0xffffffffu);
  std::vector<double> px(n), ql(n), el(n), av(n), pl(n), paths(n, 0.0);
  uint32_t nlev = 0;
  for (uint32_t u = 0; u < n; ++u) {
    leaf[u] = eoff[u] == eoff[u + 1];
    if (leaf[u]) { px[u] = 1.0; pl[u] = 1.0; continue; }
    av[u] = g * g * (double)w[u];
    double q = 0.0, e_leaf = 0.0, p_l = 0.0;
    uint32_t lv = 0;
    for (uint32_t e = eoff[u]; e < eoff[u + 1]; ++e) {
      const uint32_t c = s.edge_to[e];
      const double ce = gpow[s.edge_gaps[e]] * (double)s.edge_weight[e];
      q += ce * px[c];
      p_l += pl[c];
      if (leaf[c]) e_leaf += ce; else lv = std::max(lv, level[c] + 1);
    }
    ql[u] = q; el[u] = e_leaf; px[u] = av[u] * q; pl[u] = p_l; level[u] = lv;
    nlev = std::max(nlev, lv + 1);
  }
  // root -> node path counts (parents come after children in the record, so walk backwards)
  double plr = 0.0;
  uint32_t lr = 0;
  for (uint32_t k = s.root_off[r]; k < s.root_off[r + 1]; ++k) {
    const uint32_t u = s.root[k];
    if (u >= n) { o->err = "root index out of range"; return; }
    paths[u] += 1.0;
    plr += pl[u];
    if (leaf[u]) ++lr;
  }
  for (uint32_t u = n; u-- > 0;)
    for (uint32_t e = eoff[u]; e < eoff[u + 1]; ++e) paths[s.edge_to[e]] += paths[u];
  h.plr = plr; h.lr = lr;

  // ---- level order
  std::vector<uint32_t> order;
  order.reserve(n);
  o->lev_off.assign(nlev + 1, 0);
  for (uint32_t u = 0; u < n; ++u) if (!leaf[u]) ++o->lev_off[level[u] + 1];
  for (uint32_t l = 0; l < nlev; ++l) {
    o->max_rows = std::max(o->max_rows, o->lev_off[l + 1]);
    o->lev_off[l + 1] += o->lev_off[l];
  }
  {
    // inside a level the longest nodes first: the nodes at or above a row's length window are then a prefix of
    // every level (stem_fast.cu stops a level at the first node below it); ties: many inner pairs first, so that
    // neighbouring lanes see similar trip counts
    std::vector<uint32_t> nl_deg(n, 0);
    for (uint32_t u = 0; u < n; ++u)
      for (uint32_t e = eoff[u]; e < eoff[u + 1]; ++e) if (!leaf[s.edge_to[e]]) ++nl_deg[u];
    order.clear();
    for (uint32_t u = 0; u < n; ++u) if (!leaf[u]) order.push_back(u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
      if (level[a] != level[b]) return level[a] < level[b];
      const uint32_t la = last[a] - first[a], lb = last[b] - first[b];
      if (la != lb) return la > lb;
      return nl_deg[a] > nl_deg[b];
    });
    for (uint32_t k = 0; k < order.size(); ++k) newidx[order[k]] = k;
  }
  const uint32_t N = (uint32_t)order.size();
  {
    // Levels are cut into SUB-LEVELS of at most 32/kFastRows nodes (still a valid children-first partition): the
    // fast kernel then has at most one node per lane and step of its column sweep, and row blocks never straddle
    // one.  sub1 = the first sub-level whose nodes have inner pairs (everything before it is level 0).
    constexpr uint32_t kSlots = 32u / kFastRows;
    std::vector<uint32_t> sub(1, 0u);
    uint32_t sub1 = 0;
    for (uint32_t l = 0; l < nlev; ++l) {
      for (uint32_t g0 = o->lev_off[l]; g0 < o->lev_off[l + 1]; g0 += kSlots) sub.push_back(std::min(o->lev_off[l + 1], g0 + kSlots));
      if (l == 0) sub1 = (uint32_t)sub.size() - 1u;
    }
    o->lev_off.swap(sub);
    nlev = (uint32_t)o->lev_off.size() - 1u;
    h.sub1 = sub1;
  }
  h.N = N; h.nlev = nlev;
  o->a.resize(N); o->el.resize(N); o->ql.resize(N); o->paths.resize(N); o->gapt.resize(N); o->bfreq.resize(N);
  o->len.resize(N); o->bcode.resize(N); o->coff.assign(N + 1, 0); o->boff.assign(N + 1, 0);
  bool simple_bpf = true, len_mono = true;
  for (uint32_t k = 0; k < N; ++k) {
    const uint32_t u = order[k];
    o->a[k] = av[u]; o->el[k] = el[u]; o->ql[k] = ql[u]; o->paths[k] = paths[u];
    if (first[u] >= L) { o->err = "node position outside the sequence"; return; }
    o->gapt[k] = (double)prof[5 * first[u] + 4] / (double)nrows;
    o->len[k] = last[u] - first[u];
    o->deg_all.push_back(eoff[u + 1] - eoff[u]);
    for (uint32_t e = eoff[u]; e < eoff[u + 1]; ++e) {
      const uint32_t c = s.edge_to[e];
      if (leaf[c]) continue;
      if (last[c] - first[c] >= last[u] - first[u]) len_mono = false;
      o->cidx.push_back(newidx[c]);
      o->ce.push_back(gpow[s.edge_gaps[e]] * (double)s.edge_weight[e]);
    }
    o->coff[k + 1] = (uint32_t)o->cidx.size();
    const uint32_t nb = boff[u + 1] - boff[u];
    for (uint32_t b = boff[u]; b < boff[u + 1]; ++b) {
      if (s.bpf_a[b] > 3 || s.bpf_b[b] > 3) { o->err = "base code out of range in a base-pair profile"; return; }
      o->bab.push_back((uint8_t)(s.bpf_a[b] * 4 + s.bpf_b[b]));
      o->bfq.push_back((double)s.bpf_freq[b]);
    }
    o->boff[k + 1] = (uint32_t)o->bab.size();
    if (nb == 1) { o->bcode[k] = o->bab.back(); o->bfreq[k] = o->bfq.back(); }
    else { o->bcode[k] = 0xFF; o->bfreq[k] = 0.0; simple_bpf = false; }
  }
  if (simple_bpf) h.flags |= REC_SIMPLE_BPF;
  if (len_mono) h.flags |= REC_LEN_MONOTONE;
  uint32_t max_len = 0;
  for (uint32_t k = 0; k < N; ++k) max_len = std::max(max_len, o->len[k]);

  // ---- separable fast path.  A stem edge has gaps = len_p - len_c - 2 (dag.h:25-26), so
  // e(p,c) = g^gaps = g^(len_p-2-B) * g^(B-len_c) for any reference B: rows can be stored pre-scaled by
  // up = g^(B-len) and a parent just ADDS its children and multiplies once by s2 = g^(len-2-B).  B = half the
  // longest pair keeps every factor within g^(+-max_len/2).  Eligible records: every non-leaf edge has weight 1
  // and exactly that gap count, single-entry base-pair profiles, no gap columns under a node, and powers that
  // stay far from the fp64 range limits.  Everything else runs on the general kernel.
  {
    const uint32_t B = max_len / 2;
    const uint32_t span = std::max(B, max_len - B) + 2;
    bool fast = simple_bpf && len_mono && N > 0 && N <= kFastMaxN && max_len < 65535u;
    for (uint32_t k = 0; k < N && fast; ++k) if (o->gapt[k] != 0.0) fast = false;
    for (uint32_t k = 0; k < N && fast; ++k) {
      const uint32_t u = order[k];
      for (uint32_t e = eoff[u]; e < eoff[u + 1]; ++e) {
        const uint32_t c = s.edge_to[e];
        if (leaf[c]) continue;
        const uint32_t lp = last[u] - first[u], lc = last[c] - first[c];
        if (s.edge_weight[e] != 1.0f || lc + 2 > lp || s.edge_gaps[e] != lp - lc - 2) { fast = false; break; }
      }
    }
    std::vector<double> pos(span + 1), neg(span + 1);
    pos[0] = neg[0] = 1.0;
    const double ginv = 1.0 / g;
    for (uint32_t k = 1; k <= span; ++k) { pos[k] = pos[k - 1] * g; neg[k] = neg[k - 1] * ginv; }
    auto in_range = [](double v) { return std::isfinite(v) && std::fabs(v) > 1e-125 && std::fabs(v) < 1e125; };
    if (!in_range(pos[span]) || !in_range(neg[span])) fast = false;
    auto pw = [&](long k) { return k >= 0 ? pos[(size_t)k] : neg[(size_t)(-k)]; };
    o->up.resize(N); o->dn.resize(N); o->s2.resize(N); o->nodei.resize(N);
    for (uint32_t k = 0; k < N; ++k) {
      const long l = (long)o->len[k];
      o->up[k] = fast ? pw((long)B - l) : 0.0;
      o->dn[k] = fast ? pw(l - (long)B) : 0.0;
      o->s2[k] = fast ? pw(l - 2 - (long)B) : 0.0;
      NodeI ni;
      ni.e4_bcode = ((uint32_t)o->c16.size() << 8) | o->bcode[k];
      const uint32_t deg = o->coff[k + 1] - o->coff[k];
      ni.deg4 = (uint16_t)((deg + 3u) / 4u);
      ni.len = (uint16_t)o->len[k];
      o->nodei[k] = ni;
      o->c16.resize(o->c16.size() + 4u * ni.deg4, (uint16_t)(8u * N));   // dummy column: reads 0.0
      if ((o->c16.size() >> 24) != 0) fast = false;
    }
    // Order of the children inside every padded list.  The fast kernel sweeps the nodes of a level 32/kFastRows at a
    // time, one lane per node, and in step p every lane gathers the p-th entry of its list: the step costs as many
    // shared-memory wavefronts as the fullest 8-byte bank.  The sum over a list is order-free, so the entries are
    // placed greedily, step by step, on banks no other node of the group uses in that step (a pad entry -- they all
    // read the same dummy word -- may take a child's place when the list has room left).
    {
      constexpr uint32_t kSlots = 32u / kFastRows;
      std::vector<std::vector<uint32_t>> rem;
      for (uint32_t l = 0; l < nlev; ++l)
        for (uint32_t g0 = o->lev_off[l]; g0 < o->lev_off[l + 1]; g0 += kSlots) {
          const uint32_t g1 = std::min(o->lev_off[l + 1], g0 + kSlots);
          rem.assign(g1 - g0, {});
          uint32_t steps = 0;
          for (uint32_t k = g0; k < g1; ++k) {
            rem[k - g0].assign(o->cidx.begin() + o->coff[k], o->cidx.begin() + o->coff[k + 1]);
            steps = std::max<uint32_t>(steps, 4u * o->nodei[k].deg4);
          }
          std::vector<uint32_t> idx(g1 - g0);
          for (uint32_t p = 0; p < steps; ++p) {
            uint32_t used[16];   // bank -> node column holding it in this step (0xffffffff: free)
            std::fill(used, used + 16, 0xffffffffu);
            uint32_t m = 0;
            for (uint32_t k = g0; k < g1; ++k) if (p < 4u * o->nodei[k].deg4) idx[m++] = k - g0;
            std::stable_sort(idx.begin(), idx.begin() + m, [&](uint32_t a, uint32_t b) { return rem[a].size() < rem[b].size(); });
            for (uint32_t q = 0; q < m; ++q) {
              std::vector<uint32_t>& r = rem[idx[q]];
              const uint32_t k = g0 + idx[q];
              uint16_t* slot = o->c16.data() + (o->nodei[k].e4_bcode >> 8) + p;
              if (r.empty()) continue;                                     // stays a pad entry
              const uint32_t left = 4u * o->nodei[k].deg4 - p;              // steps left for this list, this one included
              size_t pick = r.size();
              for (size_t c = 0; c < r.size(); ++c)
                if (used[r[c] & 15u] == 0xffffffffu || used[r[c] & 15u] == r[c]) { pick = c; break; }
              if (pick == r.size()) {
                if (r.size() < left && (used[N & 15u] == 0xffffffffu || used[N & 15u] == N)) { used[N & 15u] = N; continue; }
                pick = 0;
              }
              if (used[r[pick] & 15u] == 0xffffffffu) used[r[pick] & 15u] = r[pick];
              *slot = (uint16_t)(8u * r[pick]);
              r.erase(r.begin() + (long)pick);
            }
          }
        }
    }
    for (uint32_t l = 0; l < nlev; ++l)
      for (uint32_t r = o->lev_off[l]; r < o->lev_off[l + 1]; r += kFastRows)
        o->blk.push_back(r | (std::min(kFastRows, o->lev_off[l + 1] - r) << 16));
    if (fast) h.flags |= REC_FAST;
    // nodes sorted by length (the band of a row is a range of this list) and the most nodes any window
    // [l - band, l + band] holds: the size of the fast kernel's per-warp MATCH buffer
    o->lperm.resize(N);
    for (uint32_t k = 0; k < N; ++k) o->lperm[k] = (std::min(o->len[k], 0xffffu) << 16) | (k & 0xffffu);
    std::sort(o->lperm.begin(), o->lperm.end());
    if (len_band == 0) o->band_cnt = N;
    else {
      uint32_t lo = 0;
      for (uint32_t hi = 0; hi < N; ++hi) {
        while ((o->lperm[hi] >> 16) - (o->lperm[lo] >> 16) > 2 * len_band) ++lo;
        o->band_cnt = std::max(o->band_cnt, hi - lo + 1);
      }
    }
  }
  o->pd.assign(max_len + 2, 0.0); o->pb.assign(max_len + 2, 0.0);
  for (uint32_t k = 0; k < N; ++k) {
    o->pd[o->len[k] + 1] += o->deg_all[k];
    o->pb[o->len[k] + 1] += o->boff[k + 1] - o->boff[k];
  }
  for (uint32_t l = 1; l < max_len + 2; ++l) { o->pd[l] += o->pd[l - 1]; o->pb[l] += o->pb[l - 1]; }
}

template <class T>
void append(std::vector<T>& dst, const std::vector<T>& src) { dst.insert(dst.end(), src.begin(), src.end()); }

}  // namespace

std::string compile_set(const stemk_seqset_desc& s, double g, uint32_t len_band, int n_threads, bool timing, CompiledSet* out,
                        const std::function<char*(size_t)>& sink) {
  const uint32_t n = s.n_seqs;
  auto t0 = std::chrono::steady_clock::now();
  std::vector<RecOut> recs(n);
  if (n_threads < 1) n_threads = 1;
  n_threads = std::min<int>(n_threads, std::max<uint32_t>(1, n / 16));
  // exceptions (std::bad_alloc on a hostile descriptor) must not leave a worker thread: they become the record's error
  auto guarded = [&](uint32_t r) {
    try { compile_record(s, r, g, len_band, &recs[r]); }
    catch (const std::exception& ex) { recs[r].err = std::string("exception: ") + ex.what(); }
    catch (...) { recs[r].err = "unknown exception"; }
  };
  if (n_threads <= 1) {
    for (uint32_t r = 0; r < n; ++r) guarded(r);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
      th.push_back(std::thread([&, t]() { for (uint32_t r = t; r < n; r += n_threads) guarded(r); }));
    for (auto& x : th) x.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  CompiledSet& c = *out;
  c = CompiledSet();
  c.rec.resize(n);
  // ---- merge.  Pass 1 (serial, sizes only): record headers, the offsets that become absolute, set statistics.
  std::vector<uint32_t> e0s(n), b0s(n);
  size_t nn = 0, ncoff = 0, nlev = 0, nboff = 0, ncol = 0, n16 = 0, nblk = 0, ne = 0, nb = 0, ncost = 0;
  c.cost_off.reserve(n + 1); c.n_nodes_all.reserve(n); c.n_edges_all.reserve(n); c.max_level_rows.reserve(n); c.band_cnt.reserve(n);
  for (uint32_t r = 0; r < n; ++r) {
    RecOut& o = recs[r];
    if (!o.err.empty()) return "record " + std::to_string(r) + ": " + o.err;
    RecDev h = o.hdr;
    h.node0 = (uint32_t)nn; h.coff0 = (uint32_t)ncoff; h.lev0 = (uint32_t)nlev; h.boff0 = (uint32_t)nboff; h.col0 = (uint32_t)ncol;
    h.c16_0 = (uint32_t)n16; h.e4 = (uint32_t)o.c16.size();
    h.blk0 = (uint32_t)nblk; h.nblk = (uint32_t)o.blk.size();
    if (h.flags & REC_FAST) {
      c.max_E4 = std::max(c.max_E4, h.e4); c.max_fastN = std::max(c.max_fastN, h.N); ++c.n_fast;
      c.max_band_cnt = std::max(c.max_band_cnt, o.band_cnt);
    }
    e0s[r] = (uint32_t)ne; b0s[r] = (uint32_t)nb;
    if (o.pd.empty()) { o.pd.assign(1, 0.0); o.pb.assign(1, 0.0); }
    c.cost_off.push_back(ncost);
    c.rec[r] = h;
    c.n_nodes_all.push_back(o.n_all);
    c.n_edges_all.push_back(o.e_all);
    c.max_level_rows.push_back(o.max_rows);
    c.band_cnt.push_back(o.band_cnt);
    if (o.n_all) c.has_dag = true;
    if (h.flags & REC_HAS_WEIGHT) ++c.n_weighted;
    if (h.flags & REC_SIMPLE_COLS) ++c.n_simple_cols;
    c.max_N = std::max(c.max_N, h.N);
    c.max_L = std::max(c.max_L, h.L);
    c.max_E = std::max(c.max_E, (uint32_t)o.cidx.size());
    c.max_nlev = std::max(c.max_nlev, h.nlev);
    nn += o.a.size(); ncoff += o.coff.size(); nlev += o.lev_off.size(); nboff += o.boff.size(); ncol += o.ccode.size();
    n16 += o.c16.size(); nblk += o.blk.size(); ne += o.cidx.size(); nb += o.bab.size(); ncost += o.pd.size();
  }
  c.cost_off.push_back(ncost);
  // what the host keeps (work model, scheduler, launch shapes)
  c.len.resize(nn); c.deg_all.resize(nn); c.boff.resize(nboff); c.text.resize(ncol); c.cost_pd.resize(ncost); c.cost_pb.resize(ncost);
  // ---- the device image: every array of the SetView at its place in ONE buffer (the order of stemk_api.cu's
  // make_view), 256-byte aligned; with no sink (host-only context) only the host arrays are filled
  enum { A_REC, A_A, A_EL, A_QL, A_PATHS, A_GAPT, A_BFREQ, A_LEN, A_BCODE, A_COFF, A_CIDX, A_CE, A_LEV, A_BOFF, A_BAB, A_BFQ,
         A_CCODE, A_CW, A_PROF, A_TEXT, A_UP, A_DN, A_S2, A_NODEI, A_C16, A_BLK, A_XNODE, A_LPERM, A_COUNT };
  static_assert(A_COUNT == kBlobArrays, "blob layout");
  size_t off = 0;
  auto place = [&](size_t bytes) { off = (off + 255) & ~size_t(255); const size_t at = off; off += bytes; return at; };
  size_t* L = c.blob_lay;
  L[A_REC] = place(n * sizeof(RecDev));
  L[A_A] = place(nn * 8); L[A_EL] = place(nn * 8); L[A_QL] = place(nn * 8); L[A_PATHS] = place(nn * 8); L[A_GAPT] = place(nn * 8);
  L[A_BFREQ] = place(nn * 8); L[A_LEN] = place(nn * 4); L[A_BCODE] = place(nn); L[A_COFF] = place(ncoff * 4);
  L[A_CIDX] = place(ne * 4); L[A_CE] = place(ne * 8); L[A_LEV] = place(nlev * 4); L[A_BOFF] = place(nboff * 4);
  L[A_BAB] = place(nb); L[A_BFQ] = place(nb * 8); L[A_CCODE] = place(ncol); L[A_CW] = place(ncol * 8);
  L[A_PROF] = place(ncol * 4 * 4); L[A_TEXT] = place(ncol); L[A_UP] = place(nn * 8); L[A_DN] = place(nn * 8); L[A_S2] = place(nn * 8);
  L[A_NODEI] = place(nn * sizeof(NodeI)); L[A_C16] = place(n16 * 2); L[A_BLK] = place(nblk * 4);
  L[A_XNODE] = place(nn * sizeof(XNode)); L[A_LPERM] = place(nn * 4);
  c.blob_bytes = (off + 255) & ~size_t(255);
  char* blob = sink ? sink(c.blob_bytes) : nullptr;
  if (sink && !blob) return "out of memory for the compiled set";
  // Pass 2: the arrays themselves, records in parallel (each thread a contiguous range of records).
  {
    auto copy = [&](int which, size_t elem_off, const auto& v) {
      if (blob && !v.empty()) std::memcpy(blob + L[which] + elem_off * sizeof(v[0]), v.data(), v.size() * sizeof(v[0]));
    };
    auto fill = [&](uint32_t r_lo, uint32_t r_hi) {
      for (uint32_t r = r_lo; r < r_hi; ++r) {
        RecOut& o = recs[r];
        const RecDev& h = c.rec[r];
        // child / profile offsets become absolute so the kernels index cidx/ce/bab/bfq directly
        for (auto& v : o.coff) v += e0s[r];
        for (auto& v : o.boff) v += b0s[r];
        copy(A_A, h.node0, o.a); copy(A_EL, h.node0, o.el); copy(A_QL, h.node0, o.ql); copy(A_PATHS, h.node0, o.paths);
        copy(A_GAPT, h.node0, o.gapt); copy(A_BFREQ, h.node0, o.bfreq); copy(A_LEN, h.node0, o.len); copy(A_BCODE, h.node0, o.bcode);
        copy(A_COFF, h.coff0, o.coff); copy(A_CIDX, e0s[r], o.cidx); copy(A_CE, e0s[r], o.ce); copy(A_LEV, h.lev0, o.lev_off);
        copy(A_BOFF, h.boff0, o.boff); copy(A_BAB, b0s[r], o.bab); copy(A_BFQ, b0s[r], o.bfq); copy(A_CCODE, h.col0, o.ccode);
        copy(A_CW, h.col0, o.cw); copy(A_PROF, (size_t)4 * h.col0, o.prof); copy(A_TEXT, h.col0, o.text);
        copy(A_UP, h.node0, o.up); copy(A_DN, h.node0, o.dn); copy(A_S2, h.node0, o.s2); copy(A_NODEI, h.node0, o.nodei);
        copy(A_C16, h.c16_0, o.c16); copy(A_BLK, h.blk0, o.blk); copy(A_LPERM, h.node0, o.lperm);
        if (blob) {   // packed row records (absolute child ranges)
          XNode* xn = reinterpret_cast<XNode*>(blob + L[A_XNODE]) + h.node0;
          for (uint32_t k = 0; k < h.N; ++k) {
            xn[k].s2 = o.s2[k]; xn[k].a = o.a[k]; xn[k].up = o.up[k]; xn[k].ql = o.ql[k]; xn[k].bfreq = o.bfreq[k]; xn[k].paths = o.paths[k];
            xn[k].e0 = o.coff[k]; xn[k].e1 = o.coff[k + 1]; xn[k].len = o.len[k]; xn[k].bcode = o.bcode[k];
          }
        }
        std::copy(o.len.begin(), o.len.end(), c.len.begin() + h.node0);
        std::copy(o.deg_all.begin(), o.deg_all.end(), c.deg_all.begin() + h.node0);
        std::copy(o.boff.begin(), o.boff.end(), c.boff.begin() + h.boff0);
        std::copy(o.text.begin(), o.text.end(), c.text.begin() + h.col0);
        std::copy(o.pd.begin(), o.pd.end(), c.cost_pd.begin() + (long)c.cost_off[r]);
        std::copy(o.pb.begin(), o.pb.end(), c.cost_pb.begin() + (long)c.cost_off[r]);
        RecOut().swap_arrays(o);   // the record's share is in place: give its memory back early
      }
    };
    if (blob) std::memcpy(blob + L[A_REC], c.rec.data(), n * sizeof(RecDev));
    if (n_threads <= 1) {
      fill(0, n);
    } else {
      std::vector<std::thread> th;
      for (int t = 0; t < n_threads; ++t)
        th.emplace_back(fill, (uint32_t)((uint64_t)n * t / n_threads), (uint32_t)((uint64_t)n * (t + 1) / n_threads));
      for (auto& x : th) x.join();
    }
  }
  if (timing) {
    auto t2 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "compile_set: %u records, per-record %.1f ms (%d threads), merge %.1f ms\n", n,
                 std::chrono::duration<double, std::milli>(t1 - t0).count(), n_threads,
                 std::chrono::duration<double, std::milli>(t2 - t1).count());
  }
  return "";
}

}  // namespace stemk

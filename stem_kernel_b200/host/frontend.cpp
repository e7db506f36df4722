// stem_kernel_b200/host/frontend.cpp -- base-pair lists -> MData (the DAG the stem kernel walks).
//
// Behavioural contract = the reference's front end after the probability matrix:
//   Profiler        stem_kernel_lite/data.cpp:33-132   (per-row unpaired profile, bp profile)
//   DAGBuilder      stem_kernel_lite/data.cpp:141-307  (threshold -> nodes, child lists, DFS order)
//   find_root / find_max_parent / fill_weight          data.cpp:396-453
//   ProfileSequence common/profile.cpp:11-73, char2rna common/rna.cpp:43-72
//   average of per-row matrices                         common/bpmatrix.cpp:306-342
// Own implementation on sparse pair lists and flat arrays; float/double rounding points are kept
// where the reference has them so that the resulting MData is bit-identical (tests compare it
// against dumps of the reference's own constructor).
#include "mdata.h"

#include <algorithm>
#include <cctype>
#include <map>
#include <stdexcept>
#include <unordered_map>

namespace stemk {

static const uint32_t NONE = 0xffffffffu;

uint8_t char2rna(char c) {
  // order of the reference's lookup table: a c g u t - r y m k s w b d h v n
  switch (std::tolower((unsigned char)c)) {
    case 'a': return 0; case 'c': return 1; case 'g': return 2; case 'u': return 3; case 't': return 3;
    case '-': return 4; case 'r': return 5; case 'y': return 6; case 'm': return 7; case 'k': return 8;
    case 's': return 9; case 'w': return 10; case 'b': return 11; case 'd': return 12; case 'h': return 13;
    case 'v': return 14; case 'n': return 15;
    default: return 4;
  }
}

// common/profile.cpp:11-29 (float literals as written there)
static const float kIupac[16][4] = {
    {1.0f, 0.0f, 0.0f, 0.0f}, {0.0f, 1.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 1.0f, 0.0f}, {0.0f, 0.0f, 0.0f, 1.0f},
    {0.0f, 0.0f, 0.0f, 0.0f},
    {(float)(1.0 / 2), 0.0f, (float)(1.0 / 2), 0.0f}, {0.0f, (float)(1.0 / 2), 0.0f, (float)(1.0 / 2)},
    {(float)(1.0 / 2), (float)(1.0 / 2), 0.0f, 0.0f}, {0.0f, 0.0f, (float)(1.0 / 2), (float)(1.0 / 2)},
    {0.0f, (float)(1.0 / 2), (float)(1.0 / 2), 0.0f}, {(float)(1.0 / 2), 0.0f, 0.0f, (float)(1.0 / 2)},
    {0.0f, (float)(1.0 / 3), (float)(1.0 / 3), (float)(1.0 / 3)},
    {(float)(1.0 / 3), 0.0f, (float)(1.0 / 3), (float)(1.0 / 3)},
    {(float)(1.0 / 3), (float)(1.0 / 3), 0.0f, (float)(1.0 / 3)},
    {(float)(1.0 / 3), (float)(1.0 / 3), (float)(1.0 / 3), 0.0f},
    {(float)(1.0 / 4), (float)(1.0 / 4), (float)(1.0 / 4), (float)(1.0 / 4)},
};

// profile of a set of rows, each with weight 1 (ProfileSequence::add_sequence)
static void add_row_profile(std::vector<float>& prof, float& n_rows, const std::string& row) {
  for (size_t c = 0; c < row.size(); ++c) {
    uint8_t r = char2rna(row[c]);
    if (r != 4) {
      for (int k = 0; k < 4; ++k) prof[5 * c + k] += kIupac[r][k] * 1.0f;
    } else {
      prof[5 * c + 4] += 1.0f;
    }
  }
  n_rows += 1.0f;
}

static void check_rows(const std::vector<std::string>& rows) {
  if (rows.empty()) throw std::runtime_error("empty record");
  for (size_t r = 1; r < rows.size(); ++r)
    if (rows[r].size() != rows[0].size()) throw std::runtime_error("wrong alignment");
}

MData build_mdata_seqonly(const std::vector<std::string>& rows) {
  check_rows(rows);
  MData d;
  d.length = (uint32_t)rows[0].size();
  d.profile.assign((size_t)d.length * 5, 0.0f);
  for (size_t r = 0; r < rows.size(); ++r) add_row_profile(d.profile, d.n_rows, rows[r]);
  d.edge_off.assign(1, 0);
  d.bpf_off.assign(1, 0);
  d.text = rows[0];
  return d;
}

namespace {

typedef std::pair<uint32_t, uint32_t> Pos;

inline uint64_t key(uint32_t i, uint32_t j) { return ((uint64_t)i << 32) | j; }

// One alignment row seen the way the reference's Profiler sees it.
struct RowView {
  std::vector<uint32_t> idx;      // alignment column -> ungapped index, NONE at '-'
  std::vector<float> prof;        // L*5 profile of this row alone
  std::vector<float> nbp;         // unpaired probability per alignment column
  // probability this row's Profiler reads for alignment columns (i,j), i<j, 0-based
  std::unordered_map<uint64_t, double> p;
};

struct Builder {
  uint32_t L;
  const std::vector<RowView>& rows;
  // node set and child lists
  std::unordered_map<uint64_t, std::vector<Pos> > inner;  // bp_(i,j): children of node (i,j)
  std::vector<std::vector<Pos> > head;                    // pairs starting at i, ascending j
  std::unordered_map<uint64_t, uint32_t> visited;         // (i,j) or (i,i) -> node index
  MData& out;

  Builder(uint32_t L_, const std::vector<RowView>& r, MData& o) : L(L_), rows(r), head(L_), out(o) {}

  float loop_profile(uint32_t i) const {
    float v = 0.0f, t = 0.0f;
    for (size_t r = 0; r < rows.size(); ++r) {
      if (rows[r].idx[i] != NONE) v += 1.0f * rows[r].nbp[i];
      t += 1.0f;
    }
    return v / t;
  }

  void bp_profile(uint32_t i, uint32_t j) {
    std::map<std::pair<uint8_t, uint8_t>, float> v;
    float t = 0.0f;
    for (size_t r = 0; r < rows.size(); ++r) {
      const RowView& rv = rows[r];
      if (rv.idx[i] != NONE && rv.idx[j] != NONE) {
        std::unordered_map<uint64_t, double>::const_iterator it = rv.p.find(key(i, j));
        float p = it == rv.p.end() ? 0.0f : (float)it->second;
        for (uint8_t a = 0; a < 4; ++a) {
          if (rv.prof[5 * i + a] == 0.0f) continue;
          for (uint8_t b = 0; b < 4; ++b) {
            if (rv.prof[5 * j + b] == 0.0f) continue;
            float term = 1.0f * p * rv.prof[5 * i + a] * rv.prof[5 * j + b];
            std::pair<uint8_t, uint8_t> k(a, b);
            std::map<std::pair<uint8_t, uint8_t>, float>::iterator x = v.find(k);
            if (x == v.end()) v.insert(std::make_pair(k, term)); else x->second += term;
          }
        }
      }
      t += 1.0f;
    }
    for (std::map<std::pair<uint8_t, uint8_t>, float>::const_iterator y = v.begin(); y != v.end(); ++y) {
      out.bpf_a.push_back(y->first.first);
      out.bpf_b.push_back(y->first.second);
      out.bpf_freq.push_back(y->second / t);
    }
  }

  // Emits the node for `pos` (children first) unless already emitted; returns its index.
  uint32_t emit(const Pos& pos) {
    uint64_t k = key(pos.first, pos.second);
    std::unordered_map<uint64_t, uint32_t>::const_iterator it = visited.find(k);
    if (it != visited.end()) return it->second;
    struct E { uint32_t to, gaps; };
    std::vector<E> edges;
    float w = 1.0f;
    std::vector<uint8_t> a, b;
    std::vector<float> f;
    if (pos.first != pos.second) {
      // node payload is computed before the children are visited, as in make_loop/make_stem;
      // stash bp_freq locally because children append to the shared arrays first
      size_t mark = out.bpf_a.size();
      bp_profile(pos.first, pos.second);
      a.assign(out.bpf_a.begin() + mark, out.bpf_a.end());
      b.assign(out.bpf_b.begin() + mark, out.bpf_b.end());
      f.assign(out.bpf_freq.begin() + mark, out.bpf_freq.end());
      out.bpf_a.resize(mark); out.bpf_b.resize(mark); out.bpf_freq.resize(mark);
      w = loop_profile(pos.first) * loop_profile(pos.second);
      const std::vector<Pos>& cur = inner[k];
      if (cur.empty()) {  // hairpin-closing pair: one edge to the leaf (first,first)
        uint32_t to = emit(Pos(pos.first, pos.first));
        edges.push_back(E{to, pos.second - pos.first - 1});
      } else {
        for (size_t c = 0; c < cur.size(); ++c) {
          uint32_t to = emit(cur[c]);
          edges.push_back(E{to, (cur[c].first - pos.first - 1) + (pos.second - cur[c].second - 1)});
        }
      }
    }
    out.first.push_back(pos.first);
    out.last.push_back(pos.second);
    out.weight.push_back(w);
    for (size_t e = 0; e < edges.size(); ++e) {
      out.edge_to.push_back(edges[e].to);
      out.edge_gaps.push_back(edges[e].gaps);
      out.edge_weight.push_back(1.0f);
    }
    out.edge_off.push_back((uint32_t)out.edge_to.size());
    out.bpf_a.insert(out.bpf_a.end(), a.begin(), a.end());
    out.bpf_b.insert(out.bpf_b.end(), b.begin(), b.end());
    out.bpf_freq.insert(out.bpf_freq.end(), f.begin(), f.end());
    out.bpf_off.push_back((uint32_t)out.bpf_a.size());
    uint32_t id = (uint32_t)out.first.size() - 1;
    visited[k] = id;
    return id;
  }
};

}  // namespace

MData build_mdata(const std::vector<std::string>& rows, const std::vector<BpList>& bp, float th) {
  check_rows(rows);
  if (bp.size() != rows.size()) throw std::runtime_error("one base-pair list per row expected");
  const uint32_t L = (uint32_t)rows[0].size();
  const size_t n_rows = rows.size();

  MData d;
  d.length = L;
  d.profile.assign((size_t)L * 5, 0.0f);
  for (size_t r = 0; r < n_rows; ++r) add_row_profile(d.profile, d.n_rows, rows[r]);
  d.text = rows[0];
  d.edge_off.assign(1, 0);
  d.bpf_off.assign(1, 0);

  // ---- per-row views and the averaged matrix (bpmatrix.cpp:306-342)
  std::vector<RowView> rv(n_rows);
  std::map<uint64_t, double> avg;  // alignment columns, 0-based i<j; sums in row order starting at 0.0
  for (size_t r = 0; r < n_rows; ++r) {
    RowView& v = rv[r];
    v.idx.assign(L, NONE);
    std::vector<uint32_t> col_of;  // ungapped index -> alignment column
    for (uint32_t c = 0, u = 0; c < L; ++c)
      if (rows[r][c] != '-') { v.idx[c] = u++; col_of.push_back(c); }
    v.prof.assign((size_t)L * 5, 0.0f);
    float dummy = 0.0f;
    add_row_profile(v.prof, dummy, rows[r]);
    const BpList& b = bp[r];
    if (b.i.size() != b.j.size() || b.i.size() != b.p.size()) throw std::runtime_error("ragged base-pair list");
    for (size_t k = 0; k < b.i.size(); ++k) {
      if (!(b.i[k] >= 1 && b.i[k] < b.j[k] && b.j[k] <= col_of.size()))
        throw std::runtime_error("base-pair index out of range");
      uint32_t ci = col_of[b.i[k] - 1], cj = col_of[b.j[k] - 1];
      v.p[key(ci, cj)] = b.p[k];
      std::map<uint64_t, double>::iterator it = avg.find(key(ci, cj));
      if (it == avg.end()) avg[key(ci, cj)] = 0.0 + b.p[k]; else it->second += b.p[k];
    }
  }
  for (std::map<uint64_t, double>::iterator it = avg.begin(); it != avg.end(); ++it) it->second = it->second / (double)n_rows;
  // A single row is profiled against the averaged matrix (data.cpp:331-337), which for one row is
  // the row matrix divided by 1: identical values, so rv[0].p already holds it.

  // ---- unpaired profile per row (Profiler::non_bp_profile, data.cpp:94-123)
  for (size_t r = 0; r < n_rows; ++r) {
    RowView& v = rv[r];
    v.nbp.assign(L, 1.0f);
    std::vector<std::vector<std::pair<uint32_t, double> > > partners(L);
    for (std::unordered_map<uint64_t, double>::const_iterator it = v.p.begin(); it != v.p.end(); ++it) {
      uint32_t ci = (uint32_t)(it->first >> 32), cj = (uint32_t)(it->first & 0xffffffffu);
      partners[ci].push_back(std::make_pair(cj, it->second));
      partners[cj].push_back(std::make_pair(ci, it->second));
    }
    for (uint32_t c = 0; c < L; ++c) {
      if (v.idx[c] == NONE) continue;
      std::sort(partners[c].begin(), partners[c].end());
      float x = 1.0f;
      for (size_t k = 0; k < partners[c].size(); ++k) x = (float)((double)x - partners[c][k].second);
      if (x < 0.0f) x = 0.0f;
      v.nbp[c] = x;
    }
  }

  // ---- nodes and child lists (DAGBuilder::initialize, data.cpp:165-191)
  // ch(i,j) = candidates for "children of a pair that encloses [i,j]".  Column j only needs column j-1.
  Builder bld(L, rv, d);
  std::vector<std::vector<Pos> > by_col(L);  // nodes (i,j) grouped by j
  const double thd = (double)th;
  for (std::map<uint64_t, double>::const_iterator it = avg.begin(); it != avg.end(); ++it) {
    if (!(it->second >= thd)) continue;
    uint32_t i = (uint32_t)(it->first >> 32), j = (uint32_t)(it->first & 0xffffffffu);
    if (j - i < 2) throw std::runtime_error("pair closer than 2 columns cannot be a DAG node");
    by_col[j].push_back(Pos(i, j));
  }
  std::vector<std::vector<Pos> > prev(L + 1), cur(L + 1);
  std::vector<char> is_node(L);
  for (uint32_t j = 1; j < L; ++j) {
    std::fill(is_node.begin(), is_node.end(), 0);
    for (size_t k = 0; k < by_col[j].size(); ++k) is_node[by_col[j][k].first] = 1;
    cur[j].clear();  // ch(j,j) is empty
    for (uint32_t i = j - 1;; --i) {
      std::vector<Pos>& c = cur[i];
      c.clear();
      if (is_node[i]) {
        bld.inner[key(i, j)].swap(prev[i + 1]);  // ch(i+1,j-1) moves into the node
        c.push_back(Pos(i, j));
        bld.head[i].push_back(Pos(i, j));
      } else {
        const std::vector<Pos>& below = cur[i + 1];
        const std::vector<Pos>& h = bld.head[i];
        if (h.empty()) {
          c = below;  // nothing starts at i: carried over unchanged (SURVEY 8(c) caveat (i))
        } else {
          uint32_t qmax = h.back().second;
          for (size_t k = 0; k < below.size(); ++k)
            if (!(qmax > below[k].second)) c.push_back(below[k]);
          c.insert(c.end(), h.begin(), h.end());
        }
      }
      if (i == 0) break;
    }
    prev.swap(cur);
  }

  // ---- DFS emission (DAGBuilder::build, data.cpp:151-160)
  for (uint32_t i = 0; i < L; ++i)
    for (size_t k = bld.head[i].size(); k-- > 0;) bld.emit(bld.head[i][k]);

  // ---- root / max_pa / weight (data.cpp:396-453)
  const uint32_t n = d.n_nodes();
  std::vector<char> has_parent(n, 0);
  d.max_pa.assign(n, NONE);
  for (uint32_t u = 0; u < n; ++u)
    for (uint32_t e = d.edge_off[u]; e < d.edge_off[u + 1]; ++e) {
      uint32_t c = d.edge_to[e];
      has_parent[c] = 1;
      if (d.max_pa[c] == NONE || d.max_pa[c] < u) d.max_pa[c] = u;
    }
  for (uint32_t u = 0; u < n; ++u) if (!has_parent[u]) d.root.push_back(u);
  d.seq_weight.resize(L);
  for (uint32_t c = 0; c < L; ++c) d.seq_weight[c] = bld.loop_profile(c);
  return d;
}

}  // namespace stemk

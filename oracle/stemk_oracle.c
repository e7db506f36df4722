/* oracle/stemk_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's pair kernels and Gram driver, on the flattened
 * record layout of include/stemk.h.  It follows the reference statement by statement (same loop
 * order, same operand order, K and G tables both kept) so that it reproduces the compiled
 * reference bit for bit; tests pin it against oracle/_ref (the unmodified reference sources) and
 * against the committed golden vectors in tests/golden/ that were generated from oracle/_ref.
 * The reference ships no golden vectors of its own.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this file; the
 * product (stem_kernel_b200/csrc) never does and has no CPU fallback.
 *
 * Reference lines restated:
 *   stem kernel            stem_kernel_lite/stem_kernel.cpp:14-95
 *   node / edge scores     stem_kernel_lite/score_table.cpp:14-53,56-101,118-134,162-201; score_table.h:26-29
 *   lite string kernel     stem_kernel_lite/string_kernel.cpp:11-132
 *   naive string kernel    string_kernel/string_kernel.cpp:11-50
 *   compositions           common/conv_kernel.h:13-98, stem_kernel_lite/def_kernel.h:12-192
 *   Gram / cross / diag    common/kernel_matrix.cpp:42-56,86-109,152-181,485-575,578-633,699-754
 *   text output            common/kernel_matrix.cpp:756-770
 */
#include "stemk_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../stem_kernel_b200/csrc/ribosum85_60.inc"

/* ---------------------------------------------------------------- score tables */
typedef struct {
  double co_subst[256]; /* exp(beta * ribosum_p)          score_table.cpp:124-133 */
  double subst[16];     /* exp(alpha * ribosum_s) or match/mismatch   string_kernel.cpp:16-33 */
  int stem_simple;      /* SimpleNodeScore instead of SubstNodeScore */
} tables_t;

static void make_tables(const stemk_params* p, tables_t* t) {
  int k = p->kind;
  t->stem_simple = (k == STEMK_SI_STEM || k == STEMK_SI_STEM_STR);
  for (int i = 0; i < 256; ++i) t->co_subst[i] = exp(kRibosumPair[i] * p->beta);
  int str_simple = (k == STEMK_SI_STEM_STR || k == STEMK_STR_SIMPLE);
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b)
      t->subst[a * 4 + b] = str_simple ? (a == b ? p->match : p->mismatch) : exp(kRibosumSingle[a * 4 + b] * p->alpha);
}

/* one record viewed through the descriptor */
typedef struct {
  uint32_t n, L, nroot, nw;
  const uint32_t *first, *last, *eoff, *eto, *egaps, *boff, *root;
  const float *w, *ew, *bf, *prof, *cw;
  const uint8_t *ba, *bb, *text;
  float nrows;
} rec_t;

static rec_t view(const stemk_seqset_desc* s, uint32_t i) {
  rec_t r;
  uint32_t n0 = s->node_off[i];
  r.n = s->node_off[i + 1] - n0;
  r.first = s->node_first + n0; r.last = s->node_last + n0; r.w = s->node_weight + n0;
  r.eoff = s->edge_off + n0; r.boff = s->bpf_off + n0;  /* offsets are global into the edge / bpf arrays */
  r.eto = s->edge_to; r.egaps = s->edge_gaps; r.ew = s->edge_weight;
  r.ba = s->bpf_a; r.bb = s->bpf_b; r.bf = s->bpf_freq;
  r.root = s->root + s->root_off[i]; r.nroot = s->root_off[i + 1] - s->root_off[i];
  r.L = s->col_off[i + 1] - s->col_off[i];
  r.prof = s->profile + (size_t)5 * s->col_off[i];
  r.nrows = s->n_rows[i];
  r.nw = s->weight_off[i + 1] - s->weight_off[i];
  r.cw = s->col_weight + s->weight_off[i];
  r.text = s->text ? s->text + s->col_off[i] : NULL;
  return r;
}

/* score_table.h:26-29 / 51-54 */
static double node_skip(double gap, const rec_t* x, uint32_t i) { return gap * gap * x->w[i]; }

/* score_table.cpp:14-53 (simple) and :162-201 (subst) */
static double node_match(const tables_t* t, const stemk_params* p, const rec_t* x, const rec_t* y, uint32_t i,
                         uint32_t j) {
  double v_c = 0.0;
  for (uint32_t ix = x->boff[i]; ix != x->boff[i + 1]; ++ix) {
    uint8_t a = x->ba[ix], b = x->bb[ix];
    double cx = x->bf[ix];
    for (uint32_t iy = y->boff[j]; iy != y->boff[j + 1]; ++iy) {
      uint8_t c = y->ba[iy], d = y->bb[iy];
      double cy = y->bf[iy];
      if (t->stem_simple) {
        double v = (a != c || b != d) ? p->covar : p->stack;
        v_c += v * cx * cy;
      } else {
        v_c += t->co_subst[((a * 4 + b) * 4 + c) * 4 + d] * cx * cy;
      }
    }
  }
  double nbp_x = x->prof[5 * x->first[i] + 4];
  v_c += node_skip(p->loop_gap, y, j) * nbp_x / x->nrows;
  double nbp_y = y->prof[5 * y->first[j] + 4];
  v_c += node_skip(p->loop_gap, x, i) * nbp_y / y->nrows;
  return v_c;
}

/* stem_kernel.cpp:14-95 */
static double stem_pair(const tables_t* t, const stemk_params* p, const rec_t* x, const rec_t* y) {
  /* SimpleEdgeScore::initialize, score_table.cpp:61-77: g^k by repeated multiplication */
  uint32_t gsz = (x->L > y->L ? x->L : y->L) * 2;
  if (gsz < 2) gsz = 2;
  double* g = (double*)malloc(sizeof(double) * gsz);
  g[0] = 1.0;
  for (uint32_t k = 1; k != gsz; ++k) g[k] = g[k - 1] * p->loop_gap;

  size_t nx = x->n, ny = y->n;
  double* K0 = (double*)calloc(nx * ny + 1, sizeof(double));
  double* G0 = (double*)calloc(nx * ny + 1, sizeof(double));
  double* K1 = (double*)calloc(ny + 1, sizeof(double));
  double* G1 = (double*)calloc(ny + 1, sizeof(double));
  for (uint32_t i = 0; i != nx; ++i) {
    int x_leaf = x->eoff[i] == x->eoff[i + 1];
    for (uint32_t j = 0; j != ny; ++j) {
      int y_leaf = y->eoff[j] == y->eoff[j + 1];
      if (x_leaf && y_leaf) {
        K0[i * ny + j] = G0[i * ny + j] = 1.0;
        continue;
      }
      K1[j] = G1[j] = 0.0;
      uint32_t lx = x->last[i] - x->first[i], ly = y->last[j] - y->first[j];
      uint32_t dl = lx > ly ? lx - ly : ly - lx;
      if (!x_leaf && !y_leaf && (p->len_band == 0 || dl <= p->len_band)) {
        double v_s = node_match(t, p, x, y, i, j);
        for (uint32_t ix = x->eoff[i]; ix != x->eoff[i + 1]; ++ix) {
          for (uint32_t iy = y->eoff[j]; iy != y->eoff[j + 1]; ++iy) {
            double e_s = g[x->egaps[ix]] * g[y->egaps[iy]] * x->ew[ix] * y->ew[iy];
            double v = G0[(size_t)x->eto[ix] * ny + y->eto[iy]] * v_s * e_s;
            K1[j] += v;
            G1[j] += v;
          }
        }
      }
      for (uint32_t iy = y->eoff[j]; iy != y->eoff[j + 1]; ++iy) {
        double v_s = node_skip(p->loop_gap, y, j);
        double e_s = g[y->egaps[iy]] * y->ew[iy];
        K1[j] += K1[y->eto[iy]];
        G1[j] += G1[y->eto[iy]] * v_s * e_s;
      }
      K0[i * ny + j] = K1[j];
      G0[i * ny + j] = G1[j];
      for (uint32_t ix = x->eoff[i]; ix != x->eoff[i + 1]; ++ix) {
        double v_s = node_skip(p->loop_gap, x, i);
        double e_s = g[x->egaps[ix]] * x->ew[ix];
        K0[i * ny + j] += K0[(size_t)x->eto[ix] * ny + j];
        G0[i * ny + j] += G0[(size_t)x->eto[ix] * ny + j] * v_s * e_s;
      }
    }
  }
  double ret = 0.0;
  for (uint32_t i = 0; i != x->nroot; ++i)
    for (uint32_t j = 0; j != y->nroot; ++j) ret += K0[(size_t)x->root[i] * ny + y->root[j]];
  free(g); free(K0); free(G0); free(K1); free(G1);
  return ret;
}

/* string_kernel.cpp:46-64: expectation of the substitution score over two profile columns */
static double subst_score(const double* st, const float* x, const float* y) {
  double v_c = 0.0;
  float n = 0;
  for (int i = 0; i != 4; ++i) {
    if (x[i] == 0) continue;
    for (int j = 0; j != 4; ++j) {
      if (y[j] == 0) continue;
      n += x[i] * y[j];
      v_c += st[i * 4 + j] * x[i] * y[j];
    }
  }
  return n == 0 ? 1.0 : v_c / n;
}

/* string_kernel.cpp:66-132 (rows recycled there; full tables here, same arithmetic) */
static double string_pair(const tables_t* t, const stemk_params* p, const rec_t* x, const rec_t* y) {
  int use_weight = x->nw != 0 && y->nw != 0;
  size_t sx = x->L, sy = y->L, W = sy + 1;
  double* K0 = (double*)calloc((sx + 1) * W, sizeof(double));
  double* G0 = (double*)calloc((sx + 1) * W, sizeof(double));
  double* K1 = (double*)calloc(W, sizeof(double));
  double* G1 = (double*)calloc(W, sizeof(double));
  double gap = p->gap;
  K0[0] = G0[0] = 1.0;
  for (size_t j = 1; j != sy + 1; ++j) {
    K0[j] = 1.0;
    G0[j] = G0[j - 1] * gap;
  }
  for (size_t i = 1; i != sx + 1; ++i) {
    K0[i * W] = 1.0;
    G0[i * W] = G0[(i - 1) * W] * gap;
    K1[0] = G1[0] = 0.0;
    for (size_t j = 1; j != sy + 1; ++j) {
      double v = G0[(i - 1) * W + j - 1];
      if (use_weight) v = v * x->cw[i - 1] * y->cw[j - 1];
      v *= subst_score(t->subst, x->prof + 5 * (i - 1), y->prof + 5 * (j - 1));
      K1[j] = v + K1[j - 1];
      G1[j] = v + G1[j - 1] * gap;
      K0[i * W + j] = K1[j] + K0[(i - 1) * W + j];
      G0[i * W + j] = G1[j] + G0[(i - 1) * W + j] * gap;
    }
  }
  double ret = K0[sx * W + sy];
  free(K0); free(G0); free(K1); free(G1);
  return ret;
}

/* string_kernel/string_kernel.cpp:11-50 */
static double naive_pair(const stemk_params* p, const rec_t* x, const rec_t* y) {
  double g = p->gap, g2 = g * g;
  size_t sx = x->L, sy = y->L, W = sy + 1;
  double* K0 = (double*)calloc((sx + 1) * W, sizeof(double));
  double* G0 = (double*)calloc((sx + 1) * W, sizeof(double));
  double* K1 = (double*)calloc(W, sizeof(double));
  double* G1 = (double*)calloc(W, sizeof(double));
  K0[0] = G0[0] = 1.0;
  for (size_t i = 1; i != sx + 1; ++i) { K0[i * W] = 1.0; G0[i * W] = G0[(i - 1) * W] * g; }
  for (size_t j = 1; j != sy + 1; ++j) { K0[j] = 1.0; G0[j] = G0[j - 1] * g; }
  for (size_t i = 1; i != sx + 1; ++i) {
    K1[0] = G1[0] = 0.0;
    for (size_t j = 1; j != sy + 1; ++j) {
      K1[j] = K1[j - 1];
      G1[j] = G1[j - 1] * g;
      if (x->text[i - 1] == y->text[j - 1]) {
        K1[j] += G0[(i - 1) * W + j - 1] * g2;
        G1[j] += G0[(i - 1) * W + j - 1] * g2;
      }
      K0[i * W + j] = K0[(i - 1) * W + j] + K1[j];
      G0[i * W + j] = G0[(i - 1) * W + j] * g + G1[j];
    }
  }
  double ret = K0[sx * W + sy];
  free(K0); free(G0); free(K1); free(G1);
  return ret;
}

/* def_kernel.h / conv_kernel.h compositions */
static double kernel_pair(const tables_t* t, const stemk_params* p, const rec_t* x, const rec_t* y) {
  switch (p->kind) {
    case STEMK_SI_STEM:
    case STEMK_SU_STEM: return stem_pair(t, p, x, y);
    case STEMK_SI_STEM_STR:
    case STEMK_SU_STEM_STR: return stem_pair(t, p, x, y) + string_pair(t, p, x, y);
    case STEMK_LSU_STEM: return p->beta * log(stem_pair(t, p, x, y)) + 0.0;
    case STEMK_LSU_STR: return p->alpha * log(string_pair(t, p, x, y)) + 0.0;
    case STEMK_LSU_STEM_STR:
      return (p->beta * log(stem_pair(t, p, x, y)) + 0.0) + (p->alpha * log(string_pair(t, p, x, y)) + 0.0);
    case STEMK_STR_SUBST:
    case STEMK_STR_SIMPLE: return string_pair(t, p, x, y);
    case STEMK_STR_NAIVE: return naive_pair(p, x, y);
    default: return NAN;
  }
}

double oracle_pair(const stemk_params* p, const stemk_seqset_desc* X, uint32_t xi, const stemk_seqset_desc* Y,
                   uint32_t yi) {
  tables_t t;
  make_tables(p, &t);
  rec_t x = view(X, xi), y = view(Y, yi);
  return kernel_pair(&t, p, &x, &y);
}

void oracle_pairs(const stemk_params* p, const stemk_seqset_desc* X, const stemk_seqset_desc* Y, size_t n_pairs,
                  const uint32_t* xi, const uint32_t* yi, double* out) {
  tables_t t;
  make_tables(p, &t);
  for (size_t k = 0; k < n_pairs; ++k) {
    rec_t x = view(X, xi[k]), y = view(Y, yi[k]);
    out[k] = kernel_pair(&t, p, &x, &y);
  }
}

/* kernel_matrix.cpp:485-575 (single thread) */
void oracle_gram(const stemk_params* p, const stemk_seqset_desc* S, int normalize, double* out) {
  tables_t t;
  make_tables(p, &t);
  size_t n = S->n_seqs;
  for (size_t i = 0; i != n; ++i) {
    rec_t x = view(S, (uint32_t)i);
    for (size_t j = i; j != n; ++j) {
      rec_t y = view(S, (uint32_t)j);
      out[i * n + j] = kernel_pair(&t, p, &x, &y);
      if (i != j) out[j * n + i] = out[i * n + j];
    }
  }
  if (normalize && n > 0) {
    for (size_t i = 0; i != n - 1; ++i)
      for (size_t j = i + 1; j != n; ++j) {
        out[i * n + j] /= sqrt(out[i * n + i] * out[j * n + j]);
        out[j * n + i] = out[i * n + j];
      }
    for (size_t i = 0; i != n; ++i) out[i * n + i] = 1;
  }
}

/* kernel_matrix.cpp:578-633 */
void oracle_diag(const stemk_params* p, const stemk_seqset_desc* S, const uint32_t* sv_index, uint32_t n_sv,
                 double* out) {
  tables_t t;
  make_tables(p, &t);
  if (n_sv == 0) {
    for (uint32_t i = 0; i != S->n_seqs; ++i) { rec_t x = view(S, i); out[i] = kernel_pair(&t, p, &x, &x); }
  } else {
    for (uint32_t k = 0; k != n_sv; ++k) { rec_t x = view(S, sv_index[k]); out[sv_index[k]] = kernel_pair(&t, p, &x, &x); }
  }
}

/* kernel_matrix.cpp:635-754: rows k(train_x, test_i) for all x or the sv subset; self terms; normalisation */
void oracle_cross(const stemk_params* p, const stemk_seqset_desc* T, const stemk_seqset_desc* S,
                  const uint32_t* sv_index, uint32_t n_sv, int normalize, double* out, double* self_out) {
  tables_t t;
  make_tables(p, &t);
  size_t nt = T->n_seqs, ns = S->n_seqs;
  double* self = (double*)calloc(nt + 1, sizeof(double));
  for (size_t i = 0; i != nt; ++i) {
    rec_t te = view(T, (uint32_t)i);
    if (n_sv == 0) {
      for (size_t j = 0; j != ns; ++j) { rec_t tr = view(S, (uint32_t)j); out[i * ns + j] = kernel_pair(&t, p, &tr, &te); }
    } else {
      for (uint32_t k = 0; k != n_sv; ++k) {
        rec_t tr = view(S, sv_index[k]);
        out[i * ns + sv_index[k]] = kernel_pair(&t, p, &tr, &te);
      }
    }
    if (self_out || normalize) self[i] = kernel_pair(&t, p, &te, &te);
  }
  if (normalize) {
    double* diag = (double*)calloc(ns + 1, sizeof(double));
    oracle_diag(p, S, sv_index, n_sv, diag);
    for (size_t i = 0; i != nt; ++i)
      for (size_t j = 0; j != ns; ++j) out[i * ns + j] /= sqrt(self[i] * diag[j]);
    free(diag);
  }
  if (self_out) memcpy(self_out, self, nt * sizeof(double));
  free(self);
}

/* kernel_matrix.cpp:756-770: "label 0:<row#> 1:v 2:v ... \n" with ostream's default 6 significant digits */
long oracle_print(const double* m, size_t rows, size_t cols, const int* labels, char* buf, long cap) {
  long len = 0;
  char tmp[64];
  for (size_t i = 0; i != rows; ++i) {
    int k = labels ? snprintf(tmp, sizeof tmp, "%+d 0:%zu ", labels[i], i + 1) : snprintf(tmp, sizeof tmp, "0 0:%zu ", i + 1);
    if (len + k < cap) memcpy(buf + len, tmp, k);
    len += k;
    for (size_t j = 0; j != cols; ++j) {
      k = snprintf(tmp, sizeof tmp, "%zu:%g ", j + 1, m[i * cols + j]);
      if (len + k < cap) memcpy(buf + len, tmp, k);
      len += k;
    }
    if (len + 1 < cap) buf[len] = '\n';
    len += 1;
  }
  return len;
}

/* Work model of SURVEY 8(d): cells and algorithmic flops of one pair */
void oracle_pair_cost(const stemk_params* p, const stemk_seqset_desc* X, uint32_t xi, const stemk_seqset_desc* Y,
                      uint32_t yi, double* cells, double* flops) {
  rec_t x = view(X, xi), y = view(Y, yi);
  double c = 0, f = 0;
  int k = p->kind;
  int has_stem = (k <= STEMK_SU_STEM_STR) || k == STEMK_LSU_STEM || k == STEMK_LSU_STEM_STR;
  int has_str = k == STEMK_SI_STEM_STR || k == STEMK_SU_STEM_STR || k == STEMK_LSU_STR || k == STEMK_LSU_STEM_STR ||
                k == STEMK_STR_SUBST || k == STEMK_STR_SIMPLE;
  if (has_stem) {
    double ex = x.n ? x.eoff[x.n] - x.eoff[0] : 0, ey = y.n ? y.eoff[y.n] - y.eoff[0] : 0;
    double um = 0, ub = 0;
    for (uint32_t i = 0; i != x.n; ++i) {
      uint32_t dx = x.eoff[i + 1] - x.eoff[i];
      if (!dx) continue;
      for (uint32_t j = 0; j != y.n; ++j) {
        uint32_t dy = y.eoff[j + 1] - y.eoff[j];
        if (!dy) continue;
        uint32_t lx = x.last[i] - x.first[i], ly = y.last[j] - y.first[j];
        uint32_t dl = lx > ly ? lx - ly : ly - lx;
        if (p->len_band == 0 || dl <= p->len_band) {
          um += (double)dx * dy;
          ub += (double)(x.boff[i + 1] - x.boff[i]) * (y.boff[j + 1] - y.boff[j]);
        }
      }
    }
    c += (double)x.n * y.n;
    f += 2 * um + 3 * ub + 3 * ((double)x.n * ey + ex * (double)y.n);
  }
  if (has_str) {
    int w = x.nw != 0 && y.nw != 0;
    if (!has_stem) c += (double)x.L * y.L;
    f += (w ? 9.0 : 7.0) * x.L * y.L;
  }
  if (k == STEMK_STR_NAIVE) {
    double m = 0;
    for (uint32_t i = 0; i < x.L; ++i) for (uint32_t j = 0; j < y.L; ++j) m += x.text[i] == y.text[j];
    c += (double)x.L * y.L;
    f += 4.0 * x.L * y.L + 3.0 * m;
  }
  *cells = c; *flops = f;
}

/* ------------------------------------------------------------------------------------------------------------
 * BPLA / local-alignment kernel -- restatement of bpla_kernel/bpla_kernel.cpp, statement order and the float /
 * double mix of the reference kept (LAScore :16-45: `n` is a float, the products x*y are float, the table term
 * is double; BPLAScore :47-62: the pairing terms are float products summed in float, alpha is double;
 * local_alignment_exp :64-118; local_alignment_max :120-157; dispatch :160-175). */
static double bpla_la_score(const double* tab, const float* x, const float* y) {
  double v = 0.0;
  float n = 0;
  for (int k = 0; k != 4; ++k) {
    if (x[k] == 0) continue;
    for (int l = 0; l != 4; ++l) {
      if (y[l] == 0) continue;
      n += x[k] * y[l];
      v += tab[k * 4 + l] * x[k] * y[l];
    }
  }
  return n == 0 ? 0.0 : v / n;
}

static double bpla_score(const stemk_bpla_params* p, const stemk_bpla_set* X, size_t cx, const stemk_bpla_set* Y, size_t cy) {
  const double la = bpla_la_score(p->score, X->profile + 5 * cx, Y->profile + 5 * cy);
  if (p->no_bp) return la;
  return p->alpha * (X->p_right[cx] * Y->p_right[cy] + X->p_left[cx] * Y->p_left[cy]) + X->p_unpair[cx] * Y->p_unpair[cy] * la;
}

static double bpla_pair(const stemk_bpla_params* p, const stemk_bpla_set* X, uint32_t xr, const stemk_bpla_set* Y, uint32_t yr) {
  const size_t x0 = X->col_off[xr], lx = X->col_off[xr + 1] - x0, y0 = Y->col_off[yr], ly = Y->col_off[yr + 1] - y0;
  const size_t w = ly + 1;
  double* M = calloc(5 * (lx + 1) * w, sizeof(double));
  double *Xt = M + (lx + 1) * w, *Yt = Xt + (lx + 1) * w, *X2 = Yt + (lx + 1) * w, *Y2 = X2 + (lx + 1) * w;
  double res;
  if (!p->sw) {
    const double beta_gap = exp(p->beta * p->gap), beta_ext = exp(p->beta * p->ext);
    for (size_t i = 1; i != lx + 1; ++i)
      for (size_t j = 1; j != ly + 1; ++j) {
        M[i * w + j] = exp(p->beta * bpla_score(p, X, x0 + i - 1, Y, y0 + j - 1)) *
                       (1 + Xt[(i - 1) * w + j - 1] + Yt[(i - 1) * w + j - 1] + M[(i - 1) * w + j - 1]);
        Xt[i * w + j] = beta_gap * M[(i - 1) * w + j] + beta_ext * Xt[(i - 1) * w + j];
        Yt[i * w + j] = beta_gap * (M[i * w + j - 1] + Xt[i * w + j - 1]) + beta_ext * Yt[i * w + j - 1];
        X2[i * w + j] = M[(i - 1) * w + j] + X2[(i - 1) * w + j];
        Y2[i * w + j] = M[i * w + j - 1] + X2[i * w + j - 1] + Y2[i * w + j - 1];
      }
    res = 1 + X2[lx * w + ly] + Y2[lx * w + ly] + M[lx * w + ly];
  } else {
    double mmax = 0;
    for (size_t i = 1; i != lx + 1; ++i)
      for (size_t j = 1; j != ly + 1; ++j) {
        double m = fmax(0.0, M[(i - 1) * w + j - 1]);
        m = fmax(m, Xt[(i - 1) * w + j - 1]);
        m = fmax(m, Yt[(i - 1) * w + j - 1]);
        m += bpla_score(p, X, x0 + i - 1, Y, y0 + j - 1);
        M[i * w + j] = m;
        mmax = fmax(mmax, m);
        Xt[i * w + j] = fmax(M[(i - 1) * w + j] + p->gap, Xt[(i - 1) * w + j] + p->ext);
        Yt[i * w + j] = fmax(fmax(M[i * w + j - 1] + p->gap, Xt[i * w + j - 1] + p->gap), Yt[i * w + j - 1] + p->ext);
      }
    res = mmax;
  }
  free(M);
  return res;
}

void oracle_bpla_pairs(const stemk_bpla_params* p, const stemk_bpla_set* X, const stemk_bpla_set* Y, size_t n_pairs,
                       const uint32_t* xi, const uint32_t* yi, double* out) {
  for (size_t k = 0; k < n_pairs; ++k) out[k] = bpla_pair(p, X, xi[k], Y, yi[k]);
}

/* BPLAKernel::compute_gradients (bpla_kernel.cpp:176-402): the forward tables (BPLA_Forward :178-243), the backward
 * tables (BPLA_Backward :245-314) and the accumulation of the four partial derivatives d/d{alpha, beta, gap, ext}
 * (BPLA_ForwardBackword :335-385), statement by statement.  The score is always the base-pairing-profile score
 * (:210-212); w_pair is a float expression (the profiles are float vectors), w_unpair a double one. */
enum { GM = 0, GIX, GIY, GLX, GLY, GRX, GRY, GN };

static double bpla_grad_pair(const stemk_bpla_params* p, const stemk_bpla_set* X, uint32_t xr, const stemk_bpla_set* Y, uint32_t yr,
                             double* d) {
  const size_t x0 = X->col_off[xr], lx = X->col_off[xr + 1] - x0, y0 = Y->col_off[yr], ly = Y->col_off[yr + 1] - y0;
  const size_t w = ly + 1, plane = (lx + 1) * w;
  const double alpha = p->alpha, beta = p->beta, gap = p->gap, ext = p->ext;
  const double beta_gap = exp(beta * gap), beta_ext = exp(beta * ext);
  double* F = calloc(2 * GN * plane, sizeof(double));
  double* B = F + GN * plane;
#define FT(t, i, j) F[(size_t)(t) * plane + (size_t)(i) * w + (j)]
#define BT(t, i, j) B[(size_t)(t) * plane + (size_t)(i) * w + (j)]
  /* forward */
  FT(GM, 0, 0) = 1; FT(GLX, 0, 0) = 1; FT(GLY, 0, 0) = 1;
  for (size_t i = 1; i != lx + 1; ++i) FT(GLX, i, 0) += FT(GLX, i - 1, 0);
  for (size_t j = 1; j != ly + 1; ++j) FT(GLY, 0, j) += FT(GLY, 0, j - 1);
  for (size_t i = 1; i != lx + 1; ++i)
    for (size_t j = 1; j != ly + 1; ++j) {
      const size_t cx = x0 + i - 1, cy = y0 + j - 1;
      const double s = alpha * (X->p_right[cx] * Y->p_right[cy] + X->p_left[cx] * Y->p_left[cy]) +
                       X->p_unpair[cx] * Y->p_unpair[cy] * bpla_la_score(p->score, X->profile + 5 * cx, Y->profile + 5 * cy);
      const double beta_s = exp(beta * s);
      FT(GM, i, j) += beta_s * FT(GM, i - 1, j - 1);
      FT(GM, i, j) += beta_s * FT(GIX, i - 1, j - 1);
      FT(GM, i, j) += beta_s * FT(GIY, i - 1, j - 1);
      FT(GM, i, j) += beta_s * FT(GLX, i - 1, j - 1);
      FT(GM, i, j) += beta_s * FT(GLY, i - 1, j - 1);
      FT(GIX, i, j) += beta_gap * FT(GM, i - 1, j);
      FT(GIX, i, j) += beta_ext * FT(GIX, i - 1, j);
      FT(GIY, i, j) += beta_gap * FT(GM, i, j - 1);
      FT(GIY, i, j) += beta_gap * FT(GIX, i, j - 1);
      FT(GIY, i, j) += beta_ext * FT(GIY, i, j - 1);
      FT(GLX, i, j) += FT(GLX, i - 1, 0);
      FT(GLY, i, j) += FT(GLX, i, j - 1);
      FT(GLY, i, j) += FT(GLY, i, j - 1);
      FT(GRX, i, j) += FT(GM, i - 1, j);
      FT(GRX, i, j) += FT(GRX, i - 1, j);
      FT(GRY, i, j) += FT(GM, i, j - 1);
      FT(GRY, i, j) += FT(GRX, i, j - 1);
      FT(GRY, i, j) += FT(GRY, i, j - 1);
    }
  /* backward */
  BT(GM, lx, ly) = 1; BT(GRX, lx, ly) = 1; BT(GRY, lx, ly) = 1;
  for (size_t i = lx; i != 0; --i)
    for (size_t j = ly; j != 0; --j) {
      const size_t cx = x0 + i - 1, cy = y0 + j - 1;
      const double s = alpha * (X->p_right[cx] * Y->p_right[cy] + X->p_left[cx] * Y->p_left[cy]) +
                       X->p_unpair[cx] * Y->p_unpair[cy] * bpla_la_score(p->score, X->profile + 5 * cx, Y->profile + 5 * cy);
      const double beta_s = exp(beta * s);
      BT(GM, i - 1, j - 1) += beta_s * BT(GM, i, j);
      BT(GIX, i - 1, j - 1) += beta_s * BT(GM, i, j);
      BT(GIY, i - 1, j - 1) += beta_s * BT(GM, i, j);
      BT(GLX, i - 1, j - 1) += beta_s * BT(GM, i, j);
      BT(GLY, i - 1, j - 1) += beta_s * BT(GM, i, j);
      BT(GM, i - 1, j) += beta_gap * BT(GIX, i, j);
      BT(GIX, i - 1, j) += beta_ext * BT(GIX, i, j);
      BT(GM, i, j - 1) += beta_gap * BT(GIY, i, j);
      BT(GIX, i, j - 1) += beta_gap * BT(GIY, i, j);
      BT(GIY, i, j - 1) += beta_ext * BT(GIY, i, j);
      BT(GLX, i - 1, 0) += BT(GLX, i, j);
      BT(GLX, i, j - 1) += BT(GLY, i, j);
      BT(GLY, i, j - 1) += BT(GLY, i, j);
      BT(GM, i - 1, j) += BT(GRX, i, j);
      BT(GRX, i - 1, j) += BT(GRX, i, j);
      BT(GM, i, j - 1) += BT(GRY, i, j);
      BT(GRX, i, j - 1) += BT(GRY, i, j);
      BT(GRY, i, j - 1) += BT(GRY, i, j);
    }
  /* (the two border loops of BPLA_Backward :306-312 only feed its return value, which compute_gradients drops) */
  /* forward x backward */
  d[0] = d[1] = d[2] = d[3] = 0.0;
  for (size_t i = 1; i != lx + 1; ++i)
    for (size_t j = 1; j != ly + 1; ++j) {
      const size_t cx = x0 + i - 1, cy = y0 + j - 1;
      const double w_pair = X->p_right[cx] * Y->p_right[cy] + X->p_left[cx] * Y->p_left[cy];
      const double w_unpair = X->p_unpair[cx] * Y->p_unpair[cy] * bpla_la_score(p->score, X->profile + 5 * cx, Y->profile + 5 * cy);
      const double beta_s = exp(beta * (alpha * w_pair + w_unpair));
      for (int t = GM; t <= GLY; ++t) {   /* update_alpha_beta :316-322, tables in the order M, IX, IY, LX, LY */
        const double v = FT(t, i - 1, j - 1) * beta_s * BT(GM, i, j);
        d[0] += beta * w_pair * v;
        d[1] += (alpha * w_pair + w_unpair) * v;
      }
      double v;                           /* update_beta_gap_ext :324-331 */
      v = FT(GM, i - 1, j) * beta_gap * BT(GIX, i, j);  d[1] += gap * v; d[2] += beta * v;
      v = FT(GIX, i - 1, j) * beta_ext * BT(GIX, i, j); d[1] += ext * v; d[3] += beta * v;
      v = FT(GM, i, j - 1) * beta_gap * BT(GIY, i, j);  d[1] += gap * v; d[2] += beta * v;
      v = FT(GIX, i, j - 1) * beta_gap * BT(GIY, i, j); d[1] += gap * v; d[2] += beta * v;
      v = FT(GIY, i, j - 1) * beta_ext * BT(GIY, i, j); d[1] += ext * v; d[3] += beta * v;
    }
  const double res = 1 + FT(GM, lx, ly) + FT(GRX, lx, ly) + FT(GRY, lx, ly);
#undef FT
#undef BT
  free(F);
  return res;
}

void oracle_bpla_gradients(const stemk_bpla_params* p, const stemk_bpla_set* X, const stemk_bpla_set* Y, size_t n_pairs,
                           const uint32_t* xi, const uint32_t* yi, double* value, double* grad) {
  for (size_t k = 0; k < n_pairs; ++k) value[k] = bpla_grad_pair(p, X, xi[k], Y, yi[k], grad + 4 * k);
}

/* ------------------------------------------------------------------------------------------------------------
 * Naive stem kernel -- restatement of StemKernel::full_dp (stem_kernel/stem_kernel.cpp:282-351) with dp_init /
 * dp_update (:85-111) and the base-pair classes (:353-420), statement order kept.  The reference frees the
 * planes of column j-1 after column j (:339-344); here two columns of (Lx+1) planes per table are kept. */
enum { NK0 = 0, NK1, NK2, NK3, NG0, NG1, NG2, NG3 };

static float nstem_prob(const stemk_nstem_params* p, const char* s, size_t len, const float* tab, size_t i, size_t j) {
  if (p->bp_mode == 1) return tab[i * len + j];
  const char a = s[i], b = s[j];
  int ok = (a == 'a' && b == 'u') || (a == 'u' && b == 'a') || (a == 'g' && b == 'c') || (a == 'c' && b == 'g');
  if (p->use_gu) ok = ok || (a == 'g' && b == 'u') || (a == 'u' && b == 'g');
  return (i + 1 + p->loop <= j && ok) ? 1.0f : 0.0f;
}

static double nstem_pair(const stemk_nstem_params* p, const stemk_nstem_set* X, uint32_t xr, const stemk_nstem_set* Y, uint32_t yr) {
  const char* x = X->text + X->off[xr];
  const char* y = Y->text + Y->off[yr];
  const size_t lx = X->off[xr + 1] - X->off[xr], ly = Y->off[yr + 1] - Y->off[yr];
  const float* tx = p->bp_mode == 1 ? X->bp + X->bp_off[xr] : NULL;
  const float* ty = p->bp_mode == 1 ? Y->bp + Y->bp_off[yr] : NULL;
  const double g = p->gap;
  const size_t W = ly + 1, plane = W * W, col = (lx + 1) * plane;
  double* mem = calloc(8 * 2 * col, sizeof(double));
#define DP(s, i, j, k, l) mem[((size_t)(s) * 2 + ((j) & 1)) * col + (size_t)(i) * plane + (size_t)(k) * W + (l)]
  for (size_t j = 0; j != lx + 1; ++j) {
    for (size_t k = 0; k != W; ++k)
      for (size_t l = 0; l != W; ++l) {
        DP(NK0, j, j, k, l) = 1.0;
        DP(NK1, j, j, k, l) = DP(NK2, j, j, k, l) = DP(NK3, j, j, k, l) = 0.0;
        DP(NG0, j, j, k, l) = DP(NG1, j, j, k, l) = DP(NG2, j, j, k, l) = DP(NG3, j, j, k, l) = 0.0;
      }
    for (size_t l = 0; l != ly + 1; ++l) {
      DP(NG0, j, j, l, l) = 1.0;
      if (l == 0) continue;
      for (size_t k = l - 1;; --k) {
        DP(NG0, j, j, k, l) = DP(NG0, j, j, k + 1, l) * g;
        if (k == 0) break;
      }
    }
    if (j == 0) continue;
    for (size_t i = j - 1;; --i) {
      const float bp_ij = nstem_prob(p, x, lx, tx, i, j - 1);
      for (int s = 0; s < 8; ++s) memset(&DP(s, i, j, 0, 0), 0, plane * sizeof(double));
      for (size_t l = 0; l != ly + 1; ++l) {
        DP(NK0, i, j, l, l) = 1.0;
        DP(NG0, i, j, l, l) = DP(NG0, i + 1, j, l, l) * g;
        if (l == 0) continue;
        for (size_t k = l - 1;; --k) {
          DP(NK0, i, j, k, l) = DP(NK0, i, j - 1, k, l);
          DP(NG0, i, j, k, l) = DP(NG0, i, j - 1, k, l) * g;
          DP(NK1, i, j, k, l) = DP(NK1, i + 1, j, k, l);
          DP(NG1, i, j, k, l) = DP(NG1, i + 1, j, k, l) * g;
          DP(NK2, i, j, k, l) = DP(NK2, i, j, k, l - 1);
          DP(NG2, i, j, k, l) = DP(NG2, i, j, k, l - 1) * g;
          DP(NK3, i, j, k, l) = DP(NK3, i, j, k + 1, l);
          DP(NG3, i, j, k, l) = DP(NG3, i, j, k + 1, l) * g;
          if (bp_ij > p->bp_bound) {
            const float bp_kl = nstem_prob(p, y, ly, ty, k, l - 1);
            if (bp_kl > p->bp_bound) {
              if (x[i] == y[k] && x[j - 1] == y[l - 1]) {
                DP(NK3, i, j, k, l) += DP(NG0, i + 1, j - 1, k + 1, l - 1) * p->stack * bp_ij * bp_kl;
                DP(NG3, i, j, k, l) += DP(NG0, i + 1, j - 1, k + 1, l - 1);
              } else {
                DP(NK3, i, j, k, l) += DP(NG0, i + 1, j - 1, k + 1, l - 1) * p->stack * p->subst * bp_ij * bp_kl;
              }
            }
          }
          DP(NK2, i, j, k, l) += DP(NK3, i, j, k, l);
          DP(NG2, i, j, k, l) += DP(NG3, i, j, k, l);
          DP(NK1, i, j, k, l) += DP(NK2, i, j, k, l);
          DP(NG1, i, j, k, l) += DP(NG2, i, j, k, l);
          DP(NK0, i, j, k, l) += DP(NK1, i, j, k, l);
          DP(NG0, i, j, k, l) += DP(NG1, i, j, k, l);
          if (k == 0) break;
        }
      }
      if (i == 0) break;
    }
  }
  const double res = DP(NK0, 0, lx, 0, ly);
#undef DP
  free(mem);
  return res;
}

/* StemKernel::partial_dp (stem_kernel/stem_kernel.cpp:113-280) with the band-only alignment constraints of
 * alignment_constraints (:14-83, ali_bound == 0, band > 0): row i of x may pair with columns c_low[i]..c_high[i] of y
 * around the diagonal; cells outside a plane's window keep the plane's fill value (0, or 1 for K0 of the (j,j) planes)
 * and the four neighbour terms fall back to the reference's "approximation" cells at the window border.  Statement
 * order kept; the #if 0 blocks of the reference are dead code and are not restated. */
static double nstem_pair_banded(const stemk_nstem_params* p, unsigned band, const stemk_nstem_set* X, uint32_t xr,
                                const stemk_nstem_set* Y, uint32_t yr, const uint32_t* win_low, const uint32_t* win_high) {
  const char* x = X->text + X->off[xr];
  const char* y = Y->text + Y->off[yr];
  const size_t lx = X->off[xr + 1] - X->off[xr], ly = Y->off[yr + 1] - Y->off[yr];
  const float* tx = p->bp_mode == 1 ? X->bp + X->bp_off[xr] : NULL;
  const float* ty = p->bp_mode == 1 ? Y->bp + Y->bp_off[yr] : NULL;
  const double g = p->gap;
  const size_t W = ly + 1, plane = W * W, col = (lx + 1) * plane;
  double* g_pow = malloc((ly + 1) * sizeof(double));
  g_pow[0] = 1.0;
  for (size_t i = 1; i != ly + 1; ++i) g_pow[i] = g_pow[i - 1] * g;
  /* alignment_constraints :77-83 */
  size_t* c_low = malloc(2 * (lx + 1) * sizeof(size_t));
  size_t* c_high = c_low + (lx + 1);
  for (size_t i = 0; i != lx + 1; ++i) {
    if (win_low) { c_low[i] = win_low[i]; c_high[i] = win_high[i]; continue; }   /* the caller's constraints (:25-67) */
    const unsigned j = (unsigned)((double)i / lx * ly + 0.5);
    c_low[i] = j < band ? 0 : j - band;
    c_high[i] = j + band > ly ? ly : j + band;
  }
  double* mem = calloc(8 * 2 * col, sizeof(double));
#define DP(s, i, j, k, l) mem[((size_t)(s) * 2 + ((j) & 1)) * col + (size_t)(i) * plane + (size_t)(k) * W + (l)]
  for (size_t j = 0; j != lx + 1; ++j) {
    for (size_t k = 0; k != W; ++k)
      for (size_t l = 0; l != W; ++l) {
        DP(NK0, j, j, k, l) = 1.0;
        DP(NK1, j, j, k, l) = DP(NK2, j, j, k, l) = DP(NK3, j, j, k, l) = 0.0;
        DP(NG0, j, j, k, l) = DP(NG1, j, j, k, l) = DP(NG2, j, j, k, l) = DP(NG3, j, j, k, l) = 0.0;
      }
    for (size_t l = 0; l != ly + 1; ++l) {
      DP(NG0, j, j, l, l) = 1.0;
      if (l == 0) continue;
      for (size_t k = l - 1;; --k) {
        DP(NG0, j, j, k, l) = DP(NG0, j, j, k + 1, l) * g;
        if (k == 0) break;
      }
    }
    if (j == 0) continue;
    for (size_t i = j - 1;; --i) {
      const float bp_ij = nstem_prob(p, x, lx, tx, i, j - 1);
      for (int s = 0; s < 8; ++s) memset(&DP(s, i, j, 0, 0), 0, plane * sizeof(double));
      for (size_t l = c_low[j]; l != c_high[j] + 1; ++l) {
        DP(NK0, i, j, l, l) = 1.0;
        DP(NG0, i, j, l, l) = DP(NG0, i + 1, j, l, l) * g;
        if (l == 0) continue;
        size_t k = l - 1 < c_high[i] ? l - 1 : c_high[i];
        if (k < c_low[i]) continue;                       /* the reference's loop condition k >= c_low[i] */
        for (;; --k) {
          if (l <= c_high[j - 1]) {
            DP(NK0, i, j, k, l) = DP(NK0, i, j - 1, k, l);
            DP(NG0, i, j, k, l) = DP(NG0, i, j - 1, k, l) * g;
          } else {
            DP(NK0, i, j, k, l) = DP(NK0, i, j - 1, k, c_high[j - 1]);
            DP(NG0, i, j, k, l) = DP(NG0, i, j - 1, k, c_high[j - 1]) * g * g;
          }
          if (k >= c_low[i + 1]) {
            DP(NK1, i, j, k, l) = DP(NK1, i + 1, j, k, l);
            DP(NG1, i, j, k, l) = DP(NG1, i + 1, j, k, l) * g;
          } else {
            DP(NK1, i, j, k, l) = DP(NK1, i + 1, j, c_low[i + 1], l);
            DP(NG1, i, j, k, l) = DP(NG1, i + 1, j, c_low[i + 1], l) * g * g;
          }
          if (l - 1 >= c_low[j] || k == l - 1) {
            DP(NK2, i, j, k, l) = DP(NK2, i, j, k, l - 1);
            DP(NG2, i, j, k, l) = DP(NG2, i, j, k, l - 1) * g;
          } else {
            DP(NK2, i, j, k, l) = 0.0;
            DP(NG2, i, j, k, l) = 0.0;
            for (size_t ll = k; ll != l; ++ll) {
              DP(NK2, i, j, k, l) += DP(NK3, i, j, ll, ll);
              DP(NG2, i, j, k, l) += DP(NG3, i, j, ll, ll) * g_pow[l - k];
            }
          }
          if (k + 1 <= c_high[i]) {
            DP(NK3, i, j, k, l) = DP(NK3, i, j, k + 1, l);
            DP(NG3, i, j, k, l) = DP(NG3, i, j, k + 1, l) * g;
          } else {
            DP(NK3, i, j, k, l) = DP(NK3, i, j, l, l);
            DP(NG3, i, j, k, l) = DP(NG3, i, j, l, l) * g_pow[l - k];
          }
          if (bp_ij > p->bp_bound) {
            const float bp_kl = nstem_prob(p, y, ly, ty, k, l - 1);
            if (bp_kl > p->bp_bound) {
              if (x[i] == y[k] && x[j - 1] == y[l - 1]) {
                DP(NK3, i, j, k, l) += DP(NG0, i + 1, j - 1, k + 1, l - 1) * p->stack * bp_ij * bp_kl;
                DP(NG3, i, j, k, l) += DP(NG0, i + 1, j - 1, k + 1, l - 1);
              } else {
                DP(NK3, i, j, k, l) += DP(NG0, i + 1, j - 1, k + 1, l - 1) * p->stack * p->subst * bp_ij * bp_kl;
              }
            }
          }
          DP(NK2, i, j, k, l) += DP(NK3, i, j, k, l);
          DP(NG2, i, j, k, l) += DP(NG3, i, j, k, l);
          DP(NK1, i, j, k, l) += DP(NK2, i, j, k, l);
          DP(NG1, i, j, k, l) += DP(NG2, i, j, k, l);
          DP(NK0, i, j, k, l) += DP(NK1, i, j, k, l);
          DP(NG0, i, j, k, l) += DP(NG1, i, j, k, l);
          if (k == c_low[i]) break;
        }
      }
      if (i == 0) break;
    }
  }
  const double res = DP(NK0, 0, lx, 0, ly);
#undef DP
  free(mem);
  free(c_low);
  free(g_pow);
  return res;
}

/* band == 0: full_dp; band > 0: partial_dp with the band-only constraints (stem_kernel.h:51-54 with ali_bound == 0) */
void oracle_nstem_pairs_banded(const stemk_nstem_params* p, unsigned band, const stemk_nstem_set* X, const stemk_nstem_set* Y,
                               size_t n_pairs, const uint32_t* xi, const uint32_t* yi, double* out) {
  for (size_t k = 0; k < n_pairs; ++k)
    out[k] = band > 0 ? nstem_pair_banded(p, band, X, xi[k], Y, yi[k], NULL, NULL) : nstem_pair(p, X, xi[k], Y, yi[k]);
}

/* partial_dp under caller-supplied per-row windows (the c_low / c_high of alignment_constraints, stem_kernel.cpp:14-67) */
void oracle_nstem_pairs_windows(const stemk_nstem_params* p, const stemk_nstem_set* X, const stemk_nstem_set* Y, size_t n_pairs,
                                const uint32_t* xi, const uint32_t* yi, const uint32_t* win_off, const uint32_t* c_low,
                                const uint32_t* c_high, double* out) {
  for (size_t k = 0; k < n_pairs; ++k)
    out[k] = nstem_pair_banded(p, 1, X, xi[k], Y, yi[k], c_low + win_off[k], c_high + win_off[k]);
}

void oracle_nstem_pairs(const stemk_nstem_params* p, const stemk_nstem_set* X, const stemk_nstem_set* Y, size_t n_pairs,
                        const uint32_t* xi, const uint32_t* yi, double* out) {
  for (size_t k = 0; k < n_pairs; ++k) out[k] = nstem_pair(p, X, xi[k], Y, yi[k]);
}

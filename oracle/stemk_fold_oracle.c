/* oracle/stemk_fold_oracle.c -- CPU restatement of the base-pair-probability front end.  TEST INFRASTRUCTURE: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * PARITY UNPINNED.  The reference obtains base-pair probabilities from ViennaRNA (fold / init_pf_fold / pf_fold,
 * bp(i,j) = pr[iindx[i]-j], common/bpmatrix.cpp:141-177), an external dependency ("Vienna RNA package >= 1.6",
 * README.rd:16, configure.ac:74) that is neither vendored nor present in this image, and no reference test pins a
 * value at that boundary.  This file restates the published algorithm -- McCaskill's partition function (Biopolymers
 * 29, 1990) in the arrangement of Vienna 1.8's part_func.c: qb / qm / qm1 / q inside tables, linear multiloops, dangles
 * on both sides of every multiloop and exterior stem -- for the loop model of include/stemk.h (stemk_fold_model), and
 * is itself checked against an exhaustive enumeration of all structures of short sequences (tests/test_fold.py).
 *
 * Every nucleotide's share s^-1 of the scaling (Vienna::pf_scale) is attached to the loop that owns it: a pair's two
 * bases belong to the loop the pair closes, unpaired bases to the loop they sit in.  The recursions are then
 * scale-free and Z~ = Z / s^n; probabilities do not depend on s. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/stemk.h"

#define TURN 3
#define MAXLOOP 30

static const int kRtype[7] = {0, 2, 1, 4, 3, 6, 5};

static int base_code(char c) {
  switch (c) {
    case 'a': case 'A': return 1;
    case 'c': case 'C': return 2;
    case 'g': case 'G': return 3;
    case 'u': case 'U': case 't': case 'T': return 4;
    default: return 0;
  }
}

static int pair_type(int a, int b, int no_gu) {
  if (a == 2 && b == 3) return 1;
  if (a == 3 && b == 2) return 2;
  if (a == 3 && b == 4) return no_gu ? 0 : 3;
  if (a == 4 && b == 3) return no_gu ? 0 : 4;
  if (a == 1 && b == 4) return 5;
  if (a == 4 && b == 1) return 6;
  return 0;
}

typedef struct {
  const stemk_fold_model* m;
  double kT, s;
  int n;
  const int* S;    /* 1-based codes, S[0] = S[n+1] = 0 */
  double* sp;      /* s^-k */
  double* up;      /* (exp(-ml_base/kT) / s)^k */
} Fold;

static double boltz(const Fold* f, double e) { return exp(-e / f->kT); }

static double w_hairpin(const Fold* f, int i, int j, int type) {
  const stemk_fold_model* m = f->m;
  const int u = j - i - 1;
  double e = u <= 30 ? m->hairpin[u] : m->hairpin[30] + m->lxc * log(u / 30.0);
  if (u == 3) e += type > 2 ? m->terminal_au : 0.0;
  else e += m->mismatch_h[type][f->S[i + 1]][f->S[j - 1]];
  return boltz(f, e) * f->sp[u + 2];
}

/* (i,j) closes, (k,l) is the inner pair of type tkl */
static double w_intloop(const Fold* f, int i, int j, int k, int l, int type, int tkl) {
  const stemk_fold_model* m = f->m;
  const int u1 = k - i - 1, u2 = j - l - 1, t2 = kRtype[tkl];
  double e;
  if (u1 == 0 && u2 == 0) e = m->stack[type][t2];
  else if (u1 == 0 || u2 == 0) {
    const int u = u1 + u2;
    e = m->bulge[u];
    if (u == 1) e += m->stack[type][t2];
    else e += (type > 2 ? m->terminal_au : 0.0) + (t2 > 2 ? m->terminal_au : 0.0);
  } else {
    const int d = u1 > u2 ? u1 - u2 : u2 - u1;
    const double as = d * m->ninio;
    e = m->interior[u1 + u2] + (as < m->max_ninio ? as : m->max_ninio) + m->mismatch_i[type][f->S[i + 1]][f->S[j - 1]] +
        m->mismatch_i[t2][f->S[l + 1]][f->S[k - 1]];
  }
  return boltz(f, e) * f->sp[u1 + u2 + 2];
}

static double w_mlclose(const Fold* f, int i, int j, int type) {
  const stemk_fold_model* m = f->m;
  const int tt = kRtype[type];
  return boltz(f, m->ml_closing + m->ml_intern[tt] + m->dangle3[tt][f->S[i + 1]] + m->dangle5[tt][f->S[j - 1]]) * f->sp[2];
}

static double w_mlstem(const Fold* f, int i, int j, int type) {
  const stemk_fold_model* m = f->m;
  double e = m->ml_intern[type];
  if (i > 1) e += m->dangle5[type][f->S[i - 1]];
  if (j < f->n) e += m->dangle3[type][f->S[j + 1]];
  return boltz(f, e);
}

static double w_extstem(const Fold* f, int i, int j, int type) {
  const stemk_fold_model* m = f->m;
  double e = type > 2 ? m->terminal_au : 0.0;
  if (i > 1) e += m->dangle5[type][f->S[i - 1]];
  if (j < f->n) e += m->dangle3[type][f->S[j + 1]];
  return boltz(f, e);
}

double oracle_fold_default_scale(double temperature) {
  const double kT = (temperature + 273.15) * 1.98717e-3;
  return exp(-(-185.0 + (temperature - 37.0) * 7.27) / (1000.0 * kT));
}

/* dense: (n+1) x (n+1), dense[i*(n+1)+j] = P(i,j), 1 <= i < j <= n; unpaired: [n]; either may be NULL.
 * Returns 0, or -1 when the partition function overflows / vanishes under the given scale. */
int oracle_fold_bpp(const stemk_fold_model* m, const char* seq, uint32_t n_, double* dense, double* ensemble, double* unpaired) {
  const int n = (int)n_;
  const size_t W = (size_t)n + 2;
  Fold f;
  f.m = m;
  f.kT = (m->temperature + 273.15) * 1.98717e-3;
  f.s = m->pf_scale > 0 ? m->pf_scale : oracle_fold_default_scale(m->temperature);
  f.n = n;
  int* S = (int*)calloc(W, sizeof(int));
  for (int i = 1; i <= n; ++i) S[i] = base_code(seq[i - 1]);
  f.S = S;
  double* sp = (double*)malloc((W + 2) * sizeof(double));
  double* up = (double*)malloc((W + 2) * sizeof(double));
  sp[0] = up[0] = 1.0;
  const double u1 = exp(-m->ml_base / f.kT) / f.s;
  for (size_t k = 1; k < W + 2; ++k) { sp[k] = sp[k - 1] / f.s; up[k] = up[k - 1] * u1; }
  f.sp = sp; f.up = up;
#define T2(name) double* name = (double*)calloc(W * W, sizeof(double))
#define AT(t, i, j) t[(size_t)(i) * W + (size_t)(j)]
  T2(Qb); T2(Qm); T2(Qm1); T2(Q); T2(Qq); T2(Ob); T2(Wc); T2(A); T2(Bm);
  unsigned char* ty = (unsigned char*)calloc(W * W, 1);
  for (int i = 1; i <= n; ++i)
    for (int j = i + TURN + 1; j <= n; ++j) AT(ty, i, j) = (unsigned char)pair_type(S[i], S[j], m->no_gu);
  /* empty intervals of the exterior table */
  for (int i = 1; i <= n + 1; ++i) AT(Q, i, i - 1) = 1.0;

  /* ---- inside */
  for (int d = 0; d < n; ++d)
    for (int i = 1; i + d <= n; ++i) {
      const int j = i + d, type = AT(ty, i, j);
      double qb = 0.0;
      if (type) {
        qb = w_hairpin(&f, i, j, type);
        for (int k = i + 1; k <= i + MAXLOOP + 1 && k < j - TURN - 1; ++k)
          for (int l = j - 1; l > k + TURN && (k - i - 1) + (j - l - 1) <= MAXLOOP; --l)
            if (AT(ty, k, l)) qb += AT(Qb, k, l) * w_intloop(&f, i, j, k, l, type, AT(ty, k, l));
        double t = 0.0;
        for (int k = i + 2; k <= j - 1; ++k) t += AT(Qm, i + 1, k - 1) * AT(Qm1, k, j - 1);
        qb += t * w_mlclose(&f, i, j, type);
      }
      AT(Qb, i, j) = qb;
      AT(Qm1, i, j) = AT(Qm1, i, j - 1) * up[1] + (type ? qb * w_mlstem(&f, i, j, type) : 0.0);
      AT(Qq, i, j) = AT(Qq, i, j - 1) * sp[1] + (type ? qb * w_extstem(&f, i, j, type) : 0.0);
      double qm = 0.0, q = sp[d + 1];
      for (int k = i; k <= j; ++k) {
        qm += (up[k - i] + (k > i ? AT(Qm, i, k - 1) : 0.0)) * AT(Qm1, k, j);
        q += AT(Q, i, k - 1) * AT(Qq, k, j);
      }
      AT(Qm, i, j) = qm;
      AT(Q, i, j) = q;
    }
  const double Z = n > 0 ? AT(Q, 1, n) : 1.0;
  int rc = 0;
  if (!(Z > 0.0) || !isfinite(Z)) rc = -1;
  if (ensemble) *ensemble = -f.kT * (log(Z) + n * log(f.s));

  /* ---- outside: pairs by decreasing span */
  for (int d = n - 1; d >= 1 && rc == 0; --d)
    for (int i = 1; i + d <= n; ++i) {
      const int j = i + d, type = AT(ty, i, j);
      if (type) {
        double ob = AT(Q, 1, i - 1) * AT(Q, j + 1, n) * w_extstem(&f, i, j, type);
        for (int p = i - 1; p >= 1 && i - p - 1 <= MAXLOOP; --p)
          for (int q = j + 1; q <= n && (i - p - 1) + (q - j - 1) <= MAXLOOP; ++q)
            if (AT(ty, p, q)) ob += AT(Ob, p, q) * w_intloop(&f, p, q, i, j, AT(ty, p, q), type);
        double t = 0.0;
        for (int p = 1; p < i; ++p) t += AT(Qm, p + 1, i - 1) * AT(A, p, j) + up[i - p - 1] * AT(Bm, p, j);
        ob += t * w_mlstem(&f, i, j, type);
        AT(Ob, i, j) = ob;
        AT(Wc, i, j) = ob * w_mlclose(&f, i, j, type);
      }
      /* A(p,j) = sum_{q>j} W(p,q) (up^(q-j-1) + Qm(j+1,q-1)),  Bm(p,j) = sum_{q>j} W(p,q) Qm(j+1,q-1); here p = i */
      double a = 0.0, b = 0.0;
      for (int q = j + 1; q <= n; ++q) {
        const double w = AT(Wc, i, q);
        if (w != 0.0) {
          const double qm = AT(Qm, j + 1, q - 1);
          a += w * (up[q - j - 1] + qm);
          b += w * qm;
        }
      }
      AT(A, i, j) = a;
      AT(Bm, i, j) = b;
    }

  if (dense) memset(dense, 0, (size_t)(n + 1) * (n + 1) * sizeof(double));
  if (unpaired) for (int i = 0; i < n; ++i) unpaired[i] = 0.0;
  if (rc == 0)
    for (int i = 1; i <= n; ++i)
      for (int j = i + TURN + 1; j <= n; ++j)
        if (AT(ty, i, j)) {
          const double p = AT(Qb, i, j) * AT(Ob, i, j) / Z;
          if (dense) dense[(size_t)i * (n + 1) + j] = p;
          if (unpaired) { unpaired[i - 1] += p; unpaired[j - 1] += p; }
        }
  if (unpaired) for (int i = 0; i < n; ++i) { const double v = 1.0 - unpaired[i]; unpaired[i] = v > 0.0 ? v : 0.0; }
  free(S); free(sp); free(up); free(Qb); free(Qm); free(Qm1); free(Q); free(Qq); free(Ob); free(Wc); free(A); free(Bm); free(ty);
  return rc;
}

/* The stand-in parameter set of stemk_fold_model_default (include/stemk.h), restated so that the checker does not
 * link the product: stacking / initiation / multiloop terms of Turner-1999 magnitude, synthetic mismatch and dangle
 * tables that vary with every index. */
void oracle_fold_model_default(stemk_fold_model* m) {
  static const double stack[6][6] = {{-2.4, -3.3, -2.1, -1.4, -2.1, -2.1}, {-3.3, -3.4, -2.5, -1.5, -2.2, -2.4},
                                     {-2.1, -2.5, 1.3, -0.5, -1.4, -1.3},  {-1.4, -1.5, -0.5, 0.3, -0.6, -1.0},
                                     {-2.1, -2.2, -1.4, -0.6, -1.1, -0.9}, {-2.1, -2.4, -1.3, -1.0, -0.9, -1.3}};
  static const double hp[31] = {99, 99, 99, 5.7, 5.6, 5.6, 5.4, 5.9, 5.6, 6.4, 6.5, 6.6, 6.7, 6.78, 6.86, 6.94,
                                7.01, 7.07, 7.13, 7.19, 7.25, 7.3, 7.35, 7.4, 7.44, 7.49, 7.53, 7.57, 7.61, 7.65, 7.69};
  static const double bl[31] = {99, 3.8, 2.8, 3.2, 3.6, 4.0, 4.4, 4.59, 4.7, 4.8, 4.9, 5.0, 5.1, 5.2, 5.3, 5.4,
                                5.5, 5.6, 5.7, 5.8, 5.9, 6.0, 6.1, 6.2, 6.3, 6.4, 6.5, 6.6, 6.7, 6.8, 6.9};
  static const double il[31] = {99, 99, 4.1, 5.1, 1.7, 1.8, 2.0, 2.2, 2.3, 2.4, 2.5, 2.6, 2.7, 2.8, 2.9, 3.0,
                                3.1, 3.2, 3.3, 3.4, 3.5, 3.6, 3.7, 3.8, 3.9, 4.0, 4.1, 4.2, 4.3, 4.4, 4.5};
  memset(m, 0, sizeof(*m));
  m->temperature = 37.0;
  m->pf_scale = -1.0;
  for (int a = 1; a <= 6; ++a) for (int b = 1; b <= 6; ++b) m->stack[a][b] = stack[a - 1][b - 1];
  for (int u = 0; u <= 30; ++u) { m->hairpin[u] = hp[u]; m->bulge[u] = bl[u]; m->interior[u] = il[u]; }
  m->lxc = 1.07856;
  for (int t = 1; t <= 6; ++t)
    for (int a = 0; a < 5; ++a) {
      for (int b = 0; b < 5; ++b) {
        m->mismatch_h[t][a][b] = -(0.3 + 0.1 * ((t + 2 * a + 3 * b) % 9));
        m->mismatch_i[t][a][b] = -(0.1 * ((2 * t + a + 4 * b) % 8)) + (t > 2 ? 0.7 : 0.0);
      }
      m->dangle5[t][a] = a ? -(0.1 + 0.05 * ((t + 3 * a) % 6)) : 0.0;
      m->dangle3[t][a] = a ? -(0.2 + 0.1 * ((2 * t + a) % 7)) : 0.0;
    }
  m->ninio = 0.5; m->max_ninio = 3.0;
  m->terminal_au = 0.5;
  m->ml_closing = 3.4;
  for (int t = 0; t < 8; ++t) m->ml_intern[t] = 0.4 + (t > 2 ? 0.5 : 0.0);
  m->ml_base = 0.0;
}

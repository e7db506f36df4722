// oracle/ref_harness_svm.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// The downstream equality check of SURVEY 8(c): the reference's vendored LIBSVM
// (libsvm/svm.cpp:908-1022 svm_cross_validation, :671 svm_train, :1053 svm_predict)
// run on a precomputed kernel matrix (kernel_type PRECOMPUTED, svm.h:22; node 0 is
// "0:<row#>", libsvm/svm_util.cpp:84-95).  Compiled unmodified from /root/reference.
#include <cstdlib>
#include <vector>
#include "svm.h"

static svm_parameter make_param(double C) {
  svm_parameter p;
  p.svm_type = C_SVC; p.kernel_type = PRECOMPUTED; p.degree = 3; p.gamma = 0; p.coef0 = 0;
  p.nu = 0.5; p.cache_size = 100; p.C = C; p.eps = 1e-3; p.p = 0.1; p.shrinking = 1; p.probability = 0;
  p.nr_weight = 0; p.weight_label = NULL; p.weight = NULL;
  return p;
}

// rows of a precomputed matrix -> svm_node arrays: index 0 holds the 1-based row number
static void make_nodes(int n_rows, int n_cols, const double* K, std::vector<std::vector<svm_node> >& store,
                       std::vector<svm_node*>& ptr) {
  store.assign(n_rows, std::vector<svm_node>(n_cols + 2));
  ptr.resize(n_rows);
  for (int i = 0; i < n_rows; ++i) {
    store[i][0].index = 0; store[i][0].value = i + 1;
    for (int j = 0; j < n_cols; ++j) { store[i][j + 1].index = j + 1; store[i][j + 1].value = K[(size_t)i * n_cols + j]; }
    store[i][n_cols + 1].index = -1; store[i][n_cols + 1].value = 0;
    ptr[i] = store[i].data();
  }
}

extern "C" {

// n x n training Gram matrix, labels y -> cross-validation targets
int refsvm_cv(int n, const double* K, const double* y, double C, int nr_fold, unsigned seed, double* target) {
  std::vector<std::vector<svm_node> > store; std::vector<svm_node*> ptr;
  make_nodes(n, n, K, store, ptr);
  std::vector<double> yy(y, y + n);
  svm_problem prob; prob.l = n; prob.y = yy.data(); prob.x = ptr.data();
  svm_parameter param = make_param(C);
  if (svm_check_parameter(&prob, &param)) return 1;
  srand(seed);
  svm_cross_validation(&prob, &param, nr_fold, target);
  return 0;
}

// train on K_train (n x n), predict rows of K_test (m x n)
int refsvm_train_predict(int n, const double* Ktrain, const double* y, double C, int m, const double* Ktest,
                         double* pred) {
  std::vector<std::vector<svm_node> > store, tstore; std::vector<svm_node*> ptr, tptr;
  make_nodes(n, n, Ktrain, store, ptr);
  std::vector<double> yy(y, y + n);
  svm_problem prob; prob.l = n; prob.y = yy.data(); prob.x = ptr.data();
  svm_parameter param = make_param(C);
  if (svm_check_parameter(&prob, &param)) return 1;
  svm_model* model = svm_train(&prob, &param);
  make_nodes(m, n, Ktest, tstore, tptr);
  for (int i = 0; i < m; ++i) pred[i] = svm_predict(model, tptr[i]);
  svm_destroy_model(model);
  return 0;
}

}  // extern "C"

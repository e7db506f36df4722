// oracle/ref_harness_naive.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// C ABI around the unmodified naive gap-weighted string kernel
// (string_kernel/string_kernel.cpp:11-50, class in string_kernel/string_kernel.h)
// driven by the reference's KernelMatrix (common/kernel_matrix.cpp:485-575).
// Its class name collides with stem_kernel_lite's StringKernel<V,D>, hence a
// separate shared object.  Like string_kernel/main.cpp:26,93 the gap arrives as
// a float and is widened.
#include <chrono>
#include <string>
#include <utility>
#include <vector>
#include "string_kernel/string_kernel.h"
#include "common/kernel_matrix.h"

typedef std::pair<std::string, std::string> Example;  // common/example.h:14-15
typedef std::vector<Example> ExampleSet;

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

extern "C" {

double refn_pair(float gap, const char* x, const char* y) {
  StringKernel<double> k(gap);
  return k(std::string(x), std::string(y));
}

double refn_gram(float gap, int n, const char* const* seqs, int normalize, unsigned n_th, double* out) {
  ExampleSet ex;
  for (int i = 0; i < n; ++i) ex.push_back(Example("0", seqs[i]));
  StringKernel<double> k(gap);
  KernelMatrix<double> m;
  double t0 = now_s();
  m.calculate(ex, k, normalize != 0, n_th);
  double secs = now_s() - t0;
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) out[(size_t)i * n + j] = m(i, j);
  return secs;
}

double refn_cross(float gap, int n_test, const char* const* test, int n_train, const char* const* train,
                  int norm_test, int normalize, unsigned n_th, double* out, double* self_out) {
  ExampleSet te, tr;
  for (int i = 0; i < n_test; ++i) te.push_back(Example("0", test[i]));
  for (int i = 0; i < n_train; ++i) tr.push_back(Example("0", train[i]));
  StringKernel<double> k(gap);
  KernelMatrix<double> m(n_test, n_train);
  double t0 = now_s();
  m.calculate(te, tr, k, norm_test != 0, normalize != 0, n_th);
  double secs = now_s() - t0;
  for (int i = 0; i < n_test; ++i) for (int j = 0; j < n_train; ++j) out[(size_t)i * n_train + j] = m(i, j);
  if (self_out && (norm_test || normalize)) for (int i = 0; i < n_test; ++i) self_out[i] = m(i);
  return secs;
}

}  // extern "C"

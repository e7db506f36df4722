// oracle/ref_harness_bpla.cpp -- C ABI around the UNMODIFIED reference BPLA kernel (TEST INFRASTRUCTURE).
// bpla_kernel/bpla_kernel.cpp is compiled where it lies (oracle/Makefile); this file only builds the reference's
// own `Data` objects (bpla_kernel/data.h:30-53: public members seq, p_left, p_right, p_unpair) from caller-supplied
// rows and base-pairing profiles -- the ViennaRNA-facing constructor (data.cpp:19-56) is the part that is replaced,
// exactly like the BPMatrix substitute of ref_harness.cpp -- and calls BPLAKernel::operator() (bpla_kernel.cpp:160-175).
#include <cstdint>
#include <list>
#include <string>
#include <vector>

#include "bpla_kernel/bpla_kernel.h"

namespace Vienna { extern "C" void init_rand() {} }

namespace {
std::vector<MData> build(int n, const uint32_t* row_off, const char* const* rows, const uint32_t* col_off, const float* pl,
                         const float* pr, const float* pu) {
  std::vector<MData> v;
  for (int r = 0; r < n; ++r) {
    std::list<std::string> ma;
    for (uint32_t k = row_off[r]; k < row_off[r + 1]; ++k) ma.push_back(rows[k]);
    MData d;
    d.seq = ProfileSequence(ma);
    const uint32_t c0 = col_off[r], c1 = col_off[r + 1];
    d.p_left.assign(pl + c0, pl + c1);
    d.p_right.assign(pr + c0, pr + c1);
    d.p_unpair.assign(pu + c0, pu + c1);
    v.push_back(d);
  }
  return v;
}
}  // namespace

extern "C" int refbpla_pairs(int noBP, int SW, double gap, double ext, double alpha, double beta, const double* table16,
                             int nx, const uint32_t* row_off_x, const char* const* rows_x, const uint32_t* col_off_x,
                             const float* pl_x, const float* pr_x, const float* pu_x,
                             int ny, const uint32_t* row_off_y, const char* const* rows_y, const uint32_t* col_off_y,
                             const float* pl_y, const float* pr_y, const float* pu_y,
                             size_t n_pairs, const uint32_t* xi, const uint32_t* yi, double* out) {
  boost::multi_array<double, 2> table(boost::extents[4][4]);
  for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) table[a][b] = table16[a * 4 + b];
  std::vector<MData> X = build(nx, row_off_x, rows_x, col_off_x, pl_x, pr_x, pu_x);
  std::vector<MData> Y = build(ny, row_off_y, rows_y, col_off_y, pl_y, pr_y, pu_y);
  BPLAKernel<double, MData> k(table, noBP != 0, SW != 0, gap, ext, alpha, beta);
  for (size_t p = 0; p < n_pairs; ++p) out[p] = k(X[xi[p]], Y[yi[p]]);
  return 0;
}

// BPLAKernel::compute_gradients (bpla_kernel.cpp:387-402, static): value[p] and grad[4p..] = d/d{alpha, beta, gap, ext}
extern "C" int refbpla_gradients(double gap, double ext, double alpha, double beta, const double* table16,
                                 int nx, const uint32_t* row_off_x, const char* const* rows_x, const uint32_t* col_off_x,
                                 const float* pl_x, const float* pr_x, const float* pu_x,
                                 int ny, const uint32_t* row_off_y, const char* const* rows_y, const uint32_t* col_off_y,
                                 const float* pl_y, const float* pr_y, const float* pu_y,
                                 size_t n_pairs, const uint32_t* xi, const uint32_t* yi, double* value, double* grad) {
  boost::multi_array<double, 2> table(boost::extents[4][4]);
  for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) table[a][b] = table16[a * 4 + b];
  std::vector<MData> X = build(nx, row_off_x, rows_x, col_off_x, pl_x, pr_x, pu_x);
  std::vector<MData> Y = build(ny, row_off_y, rows_y, col_off_y, pl_y, pr_y, pu_y);
  std::vector<double> param(4), d(4);
  param[0] = alpha; param[1] = beta; param[2] = gap; param[3] = ext;
  for (size_t p = 0; p < n_pairs; ++p) {
    value[p] = BPLAKernel<double, MData>::compute_gradients(X[xi[p]], Y[yi[p]], table, param, d);
    for (int k = 0; k < 4; ++k) grad[4 * p + k] = d[k];
  }
  return 0;
}

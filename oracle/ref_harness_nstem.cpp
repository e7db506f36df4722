// oracle/ref_harness_nstem.cpp -- C ABI around the UNMODIFIED reference naive stem kernel (TEST INFRASTRUCTURE).
// stem_kernel/stem_kernel.cpp keeps its member templates in the .cpp and instantiates them only for its own
// base-pair classes (:424-428), so this translation unit includes the file where it lies and adds ONE more
// instantiation whose base-pair class reads caller-supplied probability tables -- the stand-in for the
// ViennaRNA-backed BPMatrix class (:394-420, HAVE_LIBRNA), exactly like the BPMatrix substitute of ref_harness.cpp.
// The canonical-pair classes NormalBasePair / WobbleBasePair (:353-392) are used as they are.
#include <cstdint>
#include <map>
#include <string>
#include <vector>
typedef unsigned int uint;
#include "stem_kernel/stem_kernel.cpp"

namespace {
struct TableBP {   // prob(i, j) of a registered sequence object: dense row-major L x L floats
  static std::map<const std::string*, const float*>& reg() { static std::map<const std::string*, const float*> m; return m; }
  TableBP(const std::string& seq, uint /*loop*/, bool /*useGU*/) : n_(seq.size()), t_(reg()[&seq]) {}
  float prob(uint i, uint j) { return t_[(size_t)i * n_ + j]; }
  size_t n_;
  const float* t_;
};
}  // namespace
template class StemKernel<double, TableBP>;

extern "C" int refnstem_pairs(int bp_mode, int use_gu, unsigned loop, double gap, double stack, double subst, unsigned band,
                              float ali_bound, float bp_bound,
                              int nx, const uint32_t* off_x, const char* text_x, const uint64_t* bp_off_x, const float* bp_x,
                              int ny, const uint32_t* off_y, const char* text_y, const uint64_t* bp_off_y, const float* bp_y,
                              size_t n_pairs, const uint32_t* xi, const uint32_t* yi, double* out) {
  std::vector<std::string> X(nx), Y(ny);
  for (int r = 0; r < nx; ++r) X[r].assign(text_x + off_x[r], text_x + off_x[r + 1]);
  for (int r = 0; r < ny; ++r) Y[r].assign(text_y + off_y[r], text_y + off_y[r + 1]);
  if (bp_mode == 1) {
    for (int r = 0; r < nx; ++r) TableBP::reg()[&X[r]] = bp_x + bp_off_x[r];
    for (int r = 0; r < ny; ++r) TableBP::reg()[&Y[r]] = bp_y + bp_off_y[r];
    StemKernel<double, TableBP> k(use_gu != 0, loop, gap, stack, subst, band, ali_bound, bp_bound);
    for (size_t p = 0; p < n_pairs; ++p) out[p] = k(X[xi[p]], Y[yi[p]]);
    TableBP::reg().clear();
  } else if (use_gu) {
    StemKernel<double, WobbleBasePair> k(true, loop, gap, stack, subst, band, ali_bound, bp_bound);
    for (size_t p = 0; p < n_pairs; ++p) out[p] = k(X[xi[p]], Y[yi[p]]);
  } else {
    StemKernel<double, NormalBasePair> k(false, loop, gap, stack, subst, band, ali_bound, bp_bound);
    for (size_t p = 0; p < n_pairs; ++p) out[p] = k(X[xi[p]], Y[yi[p]]);
  }
  return 0;
}

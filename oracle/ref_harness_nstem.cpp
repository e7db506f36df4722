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

// alignment_constraints is private; an explicit instantiation may name it legally (same device as ref_harness.cpp's gap table)
namespace {
typedef void (StemKernel<double, NormalBasePair>::*ConstraintsFn)(const std::string&, const std::string&, std::vector<uint>&,
                                                                   std::vector<uint>&) const;
template <class Tag, typename Tag::type M>
struct Expose { friend typename Tag::type expose(Tag) { return M; } };
struct ConstraintsTag { typedef ConstraintsFn type; friend type expose(ConstraintsTag); };
template struct Expose<ConstraintsTag, &StemKernel<double, NormalBasePair>::alignment_constraints>;
}  // namespace

// c_low / c_high (lx + 1 entries each) of StemKernel::alignment_constraints for one pair (stem_kernel.cpp:14-83): the
// pair-HMM constraints when ali_bound > 0 (narrowed by the band, :57-66), the band alone otherwise
extern "C" int refnstem_windows(unsigned band, float ali_bound, const char* x, unsigned lx, const char* y, unsigned ly,
                                uint32_t* c_low, uint32_t* c_high) {
  const std::string X(x, x + lx), Y(y, y + ly);
  StemKernel<double, NormalBasePair> k(false, 3, 0.5, 1.0, 0.5, band, ali_bound, 1.0f);   // only band / ali_bound matter here
  std::vector<uint> lo, hi;
  (k.*expose(ConstraintsTag()))(X, Y, lo, hi);
  for (unsigned i = 0; i <= lx; ++i) { c_low[i] = lo[i]; c_high[i] = hi[i]; }
  return 0;
}

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

// Shim: boost::timer (CPU clock since construction), as used around each
// kernel call in common/kernel_matrix.cpp.
#pragma once
#include <ctime>
namespace boost {
class timer {
 public:
  timer() : t0_(std::clock()) {}
  void restart() { t0_ = std::clock(); }
  double elapsed() const { return double(std::clock() - t0_) / CLOCKS_PER_SEC; }
 private:
  std::clock_t t0_;
};
}

// Shim: the reference only names boost::spirit::classic::file_iterator<> in
// loader signatures (common/fa.h, aln.h, maf.h, stem_kernel_lite/data.h); the
// oracle never parses files, so an inert iterator is enough.  Also pulls in the
// std headers the real Boost header used to drag in for those sources.
#pragma once
#include <algorithm>
#include <cassert>
#include <fstream>
#include <functional>
#include <iterator>
#include <map>
#include <memory>
#include <string>
#define BOOST_SPIRIT_CLASSIC_NS boost::spirit::classic
namespace boost { namespace spirit { namespace classic {
template <class C = char>
class file_iterator {
 public:
  file_iterator() : open_(false) {}
  explicit file_iterator(const std::string& name) : open_(std::ifstream(name.c_str()).good()) {}
  operator bool() const { return open_; }
 private:
  bool open_;
};
}}}

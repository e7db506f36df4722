// Shim: the subset of boost::multi_array the reference touches -- construction
// from boost::extents[a][b]..., zero-initialised row-major storage, chained
// operator[], resize(extents), size() of the leading dimension, data(), num_elements() and ::element.
#pragma once
#include <cstddef>
#include <vector>
namespace boost {
namespace shim_detail {
template <std::size_t N>
struct ext_gen {
  std::size_t d[N ? N : 1];
  ext_gen<N + 1> operator[](std::size_t n) const {
    ext_gen<N + 1> r;
    for (std::size_t k = 0; k < N; ++k) r.d[k] = d[k];
    r.d[N] = n;
    return r;
  }
};
template <class T, std::size_t N>
struct view {
  T* p; const std::size_t* dim; const std::size_t* stride;
  view<T, N - 1> operator[](std::size_t i) const { return view<T, N - 1>{p + i * stride[0], dim + 1, stride + 1}; }
  std::size_t size() const { return dim[0]; }
};
template <class T>
struct view<T, 1> {
  T* p; const std::size_t* dim; const std::size_t* stride;
  T& operator[](std::size_t i) const { return p[i]; }
  std::size_t size() const { return dim[0]; }
};
}  // namespace shim_detail
static const shim_detail::ext_gen<0> extents = shim_detail::ext_gen<0>();

template <class T, std::size_t N>
class multi_array {
 public:
  typedef T element;
  multi_array() { for (std::size_t k = 0; k < N; ++k) dim_[k] = stride_[k] = 0; }
  explicit multi_array(const shim_detail::ext_gen<N>& e) { reshape(e); data_.assign(total(), T()); }
  void resize(const shim_detail::ext_gen<N>& e) { reshape(e); data_.assign(total(), T()); }
  std::size_t size() const { return dim_[0]; }
  T* data() { return data_.data(); }
  const T* data() const { return data_.data(); }
  std::size_t num_elements() const { return data_.size(); }
  shim_detail::view<T, N> whole() { return shim_detail::view<T, N>{data_.data(), dim_, stride_}; }
  shim_detail::view<const T, N> whole() const { return shim_detail::view<const T, N>{data_.data(), dim_, stride_}; }
  auto operator[](std::size_t i) -> decltype(this->whole()[i]) { return whole()[i]; }
  auto operator[](std::size_t i) const -> decltype(this->whole()[i]) { return whole()[i]; }
 private:
  void reshape(const shim_detail::ext_gen<N>& e) {
    for (std::size_t k = 0; k < N; ++k) dim_[k] = e.d[k];
    std::size_t s = 1;
    for (std::size_t k = N; k-- > 0;) { stride_[k] = s; s *= dim_[k]; }
  }
  std::size_t total() const { std::size_t s = 1; for (std::size_t k = 0; k < N; ++k) s *= dim_[k]; return s; }
  std::size_t dim_[N], stride_[N];
  std::vector<T> data_;
};
}  // namespace boost

// TEST INFRASTRUCTURE (oracle): stand-in for the un-vendored third-party header
// <spikes/array.hpp> (thomashilke/spikes) that the reference includes at
// src/core/cell.hpp:8, mesh.hpp:11, solver.hpp:11 ...  Written from the call
// sites in the reference (row-major N-d container); contains no arithmetic.
// Only oracle/_ref is built against it. Never included by the product.
#ifndef ORACLE_SHIM_SPIKES_ARRAY_HPP
#define ORACLE_SHIM_SPIKES_ARRAY_HPP

#include <cstddef>
#include <vector>
#include <initializer_list>
#include <algorithm>
#include <string>
#include <memory>
#include <numeric>
#include <iomanip>

template<typename T>
class array {
public:
  array() : extents_(), data_() {}
  array(std::initializer_list<std::size_t> ext) : extents_(ext), data_() {
    std::size_t n(extents_.empty() ? 0 : 1);
    for (auto e : extents_) n *= e;
    data_.resize(n);
  }
  array(const array&) = default;
  array(array&&) = default;
  array& operator=(const array&) = default;
  array& operator=(array&&) = default;

  std::size_t get_rank() const { return extents_.size(); }
  std::size_t get_size(std::size_t d) const { return extents_[d]; }
  std::size_t get_element_number() const { return data_.size(); }

  T* get_data() { return data_.data(); }
  const T* get_data() const { return data_.data(); }
  void set_data(const T* src) { std::copy(src, src + data_.size(), data_.begin()); }
  void fill(const T& v) { std::fill(data_.begin(), data_.end(), v); }

  template<typename ... Is>
  T& at(Is ... is) { return data_[offset(0, 0, static_cast<std::size_t>(is)...)]; }
  template<typename ... Is>
  const T& at(Is ... is) const { return data_[offset(0, 0, static_cast<std::size_t>(is)...)]; }

private:
  std::vector<std::size_t> extents_;
  // vector<bool> would break at(); the reference never instantiates array<bool>::at by reference
  std::vector<T> data_;

  std::size_t offset(std::size_t acc, std::size_t) const { return acc; }
  template<typename ... Is>
  std::size_t offset(std::size_t acc, std::size_t d, std::size_t i, Is ... is) const {
    return offset(acc * extents_[d] + i, d + 1, is...);
  }
};

// array_from<T, Ts...>::parameters(dst, vs...): copy a parameter pack into dst[]
// (call sites: reference src/core/mesh_data.hpp:363,379)
template<typename T, typename ... Ts>
struct array_from;

template<typename T, typename A, typename ... Ts>
struct array_from<T, A, Ts...> {
  static void parameters(T* dst, A a, Ts ... ts) {
    *dst = a;
    array_from<T, Ts...>::parameters(dst + 1, ts...);
  }
};

template<typename T>
struct array_from<T> {
  static void parameters(T*) {}
};

#endif

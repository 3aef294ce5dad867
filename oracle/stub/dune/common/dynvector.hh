// oracle-build stand-in for Dune::DynamicVector (test infrastructure)
#pragma once
#include <vector>
namespace Dune {
template <class T>
class DynamicVector {
  std::vector<T> v_;

public:
  using value_type = T;
  using size_type  = std::size_t;
  DynamicVector()  = default;
  explicit DynamicVector(std::size_t n, const T& val = T()) : v_(n, val) {}
  void resize(std::size_t n, const T& val = T()) { v_.assign(n, val); }
  std::size_t size() const { return v_.size(); }
  T& operator[](std::size_t i) { return v_[i]; }
  const T& operator[](std::size_t i) const { return v_[i]; }
  T& back() { return v_.back(); }
  T& front() { return v_.front(); }
  const T& back() const { return v_.back(); }
  const T& front() const { return v_.front(); }
  auto begin() { return v_.begin(); }
  auto end() { return v_.end(); }
  auto begin() const { return v_.begin(); }
  auto end() const { return v_.end(); }
  std::vector<T>& container() { return v_; }
  const std::vector<T>& container() const { return v_; }
};
}  // namespace Dune

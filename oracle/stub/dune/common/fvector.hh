// Minimal stand-in for dune-common's FieldVector, written for the oracle build only.
// TEST INFRASTRUCTURE: lets the reference's own headers (under /root/reference) compile without DUNE.
// Provides just the names the hot-path headers touch. Not a DUNE copy.
#pragma once
#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <memory>
#include <numeric>
#include <optional>
#include <stdexcept>
#include <tuple>
#include <vector>

namespace Dune {

struct MathError : public std::runtime_error {
  MathError() : std::runtime_error("Dune::MathError") {}
  explicit MathError(const char* m) : std::runtime_error(m) {}
};
#ifndef DUNE_THROW
#define DUNE_THROW(E, m) throw E()
#endif

template <class T, int n>
class FieldVector {
  std::array<T, (n > 0 ? n : 1)> d_{};

public:
  using value_type                = T;
  using size_type                 = std::size_t;
  static constexpr int dimension  = n;
  FieldVector()                   = default;
  FieldVector(const T& s) { d_.fill(s); }
  FieldVector(std::initializer_list<T> l) {
    std::size_t i = 0;
    for (auto& v : l)
      if (i < static_cast<std::size_t>(n)) d_[i++] = v;
  }
  template <class U>
  FieldVector(const std::array<U, static_cast<std::size_t>(n)>& a) {
    for (int i = 0; i < n; ++i) d_[i] = a[i];
  }
  FieldVector& operator=(const T& s) {
    d_.fill(s);
    return *this;
  }
  T& operator[](std::size_t i) { return d_[i]; }
  const T& operator[](std::size_t i) const { return d_[i]; }
  static constexpr std::size_t size() { return n; }
  auto begin() { return d_.begin(); }
  auto end() { return d_.begin() + n; }
  auto begin() const { return d_.begin(); }
  auto end() const { return d_.begin() + n; }
  FieldVector& operator+=(const FieldVector& o) {
    for (int i = 0; i < n; ++i) d_[i] += o.d_[i];
    return *this;
  }
  FieldVector& operator-=(const FieldVector& o) {
    for (int i = 0; i < n; ++i) d_[i] -= o.d_[i];
    return *this;
  }
  FieldVector& operator*=(const T& s) {
    for (int i = 0; i < n; ++i) d_[i] *= s;
    return *this;
  }
  FieldVector& operator/=(const T& s) {
    for (int i = 0; i < n; ++i) d_[i] /= s;
    return *this;
  }
  FieldVector operator-() const {
    FieldVector r;
    for (int i = 0; i < n; ++i) r.d_[i] = -d_[i];
    return r;
  }
  // scalar product, as in dune-common
  T operator*(const FieldVector& o) const {
    T s{};
    for (int i = 0; i < n; ++i) s += d_[i] * o.d_[i];
    return s;
  }
  T dot(const FieldVector& o) const { return (*this) * o; }
  T two_norm2() const { return (*this) * (*this); }
  T two_norm() const { return std::sqrt(two_norm2()); }
  bool operator==(const FieldVector& o) const { return d_ == o.d_; }
  // size-1 vectors convert to their scalar, as in dune-common
  operator T() const requires(n == 1) { return d_[0]; }
};

template <class T, int n>
FieldVector<T, n> operator+(FieldVector<T, n> a, const FieldVector<T, n>& b) { return a += b; }
template <class T, int n>
FieldVector<T, n> operator-(FieldVector<T, n> a, const FieldVector<T, n>& b) { return a -= b; }
template <class T, int n, class S>
requires std::is_arithmetic_v<S>
FieldVector<T, n> operator*(FieldVector<T, n> a, const S& s) { return a *= static_cast<T>(s); }
template <class T, int n, class S>
requires std::is_arithmetic_v<S>
FieldVector<T, n> operator*(const S& s, FieldVector<T, n> a) { return a *= static_cast<T>(s); }
template <class T, int n, class S>
requires std::is_arithmetic_v<S>
FieldVector<T, n> operator/(FieldVector<T, n> a, const S& s) { return a /= static_cast<T>(s); }
template <class T, int n>
T dot(const FieldVector<T, n>& a, const FieldVector<T, n>& b) { return a * b; }

}  // namespace Dune

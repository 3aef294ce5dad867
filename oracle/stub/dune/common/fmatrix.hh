// Minimal stand-in for dune-common's FieldMatrix (oracle build only; test infrastructure).
#pragma once
#include "fvector.hh"
namespace Dune {
template <class T, int r, int c>
class FieldMatrix {
  std::array<FieldVector<T, c>, (r > 0 ? r : 1)> rows_{};

public:
  using value_type               = T;
  using row_type                 = FieldVector<T, c>;
  static constexpr int rows      = r;
  static constexpr int cols      = c;
  FieldMatrix()                  = default;
  FieldMatrix(const T& s) { for (auto& q : rows_) q = s; }
  FieldMatrix(std::initializer_list<std::initializer_list<T>> l) {
    int i = 0;
    for (auto& row : l) {
      int j = 0;
      for (auto& v : row) rows_[i][j++] = v;
      ++i;
    }
  }
  row_type& operator[](std::size_t i) { return rows_[i]; }
  const row_type& operator[](std::size_t i) const { return rows_[i]; }
  static constexpr int N() { return r; }
  static constexpr int M() { return c; }
  FieldMatrix& operator+=(const FieldMatrix& o) {
    for (int i = 0; i < r; ++i) rows_[i] += o.rows_[i];
    return *this;
  }
  FieldMatrix& operator-=(const FieldMatrix& o) {
    for (int i = 0; i < r; ++i) rows_[i] -= o.rows_[i];
    return *this;
  }
  FieldMatrix& operator*=(const T& s) {
    for (int i = 0; i < r; ++i) rows_[i] *= s;
    return *this;
  }
  FieldMatrix& operator/=(const T& s) {
    for (int i = 0; i < r; ++i) rows_[i] /= s;
    return *this;
  }
  // y = A x
  template <class X, class Y>
  void mv(const X& x, Y& y) const {
    for (int i = 0; i < r; ++i) {
      T s{};
      for (int j = 0; j < c; ++j) s += rows_[i][j] * x[j];
      y[i] = s;
    }
  }
  // y = A^T x
  template <class X, class Y>
  void mtv(const X& x, Y& y) const {
    for (int j = 0; j < c; ++j) {
      T s{};
      for (int i = 0; i < r; ++i) s += rows_[i][j] * x[i];
      y[j] = s;
    }
  }
  T determinant() const requires(r == c) {
    if constexpr (r == 1)
      return rows_[0][0];
    else if constexpr (r == 2)
      return rows_[0][0] * rows_[1][1] - rows_[0][1] * rows_[1][0];
    else if constexpr (r == 3)
      return rows_[0][0] * (rows_[1][1] * rows_[2][2] - rows_[1][2] * rows_[2][1]) -
             rows_[0][1] * (rows_[1][0] * rows_[2][2] - rows_[1][2] * rows_[2][0]) +
             rows_[0][2] * (rows_[1][0] * rows_[2][1] - rows_[1][1] * rows_[2][0]);
    else
      static_assert(r <= 3);
  }
  // solve A x = b by Gaussian elimination with partial pivoting
  template <class X, class B>
  void solve(X& x, const B& b) const requires(r == c) {
    T a[r][r + 1];
    for (int i = 0; i < r; ++i) {
      for (int j = 0; j < r; ++j) a[i][j] = rows_[i][j];
      a[i][r] = b[i];
    }
    for (int k = 0; k < r; ++k) {
      int piv = k;
      for (int i = k + 1; i < r; ++i)
        if (std::abs(a[i][k]) > std::abs(a[piv][k])) piv = i;
      if (piv != k)
        for (int j = 0; j <= r; ++j) std::swap(a[k][j], a[piv][j]);
      for (int i = k + 1; i < r; ++i) {
        T f = a[i][k] / a[k][k];
        for (int j = k; j <= r; ++j) a[i][j] -= f * a[k][j];
      }
    }
    for (int i = r - 1; i >= 0; --i) {
      T s = a[i][r];
      for (int j = i + 1; j < r; ++j) s -= a[i][j] * x[j];
      x[i] = s / a[i][i];
    }
  }
};

template <class T, int r, int c>
FieldMatrix<T, r, c> operator+(FieldMatrix<T, r, c> a, const FieldMatrix<T, r, c>& b) { return a += b; }
template <class T, int r, int c>
FieldMatrix<T, r, c> operator-(FieldMatrix<T, r, c> a, const FieldMatrix<T, r, c>& b) { return a -= b; }
template <class T, int r, int k, int c>
FieldMatrix<T, r, c> operator*(const FieldMatrix<T, r, k>& a, const FieldMatrix<T, k, c>& b) {
  FieldMatrix<T, r, c> m;
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) {
      T s{};
      for (int l = 0; l < k; ++l) s += a[i][l] * b[l][j];
      m[i][j] = s;
    }
  return m;
}
template <class T, int r, int c>
FieldMatrix<T, r, c> operator*(FieldMatrix<T, r, c> a, const T& s) { return a *= s; }
template <class T, int r, int c>
FieldMatrix<T, r, c> operator*(const T& s, FieldMatrix<T, r, c> a) { return a *= s; }
}  // namespace Dune

// oracle-build stand-in for Dune::DynamicMatrix (test infrastructure)
#pragma once
#include "dynvector.hh"
namespace Dune {
template <class T>
class DynamicMatrix {
  std::vector<DynamicVector<T>> rows_;

public:
  using value_type = T;
  DynamicMatrix()  = default;
  DynamicMatrix(std::size_t r, std::size_t c, const T& val = T()) : rows_(r, DynamicVector<T>(c, val)) {}
  void resize(std::size_t r, std::size_t c, const T& val = T()) { rows_.assign(r, DynamicVector<T>(c, val)); }
  DynamicVector<T>& operator[](std::size_t i) { return rows_[i]; }
  const DynamicVector<T>& operator[](std::size_t i) const { return rows_[i]; }
  std::size_t N() const { return rows_.size(); }
  std::size_t M() const { return rows_.empty() ? 0 : rows_[0].size(); }
  std::size_t rows() const { return N(); }
  std::size_t cols() const { return M(); }
};
}  // namespace Dune

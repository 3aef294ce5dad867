// oracle-build stand-in (test infrastructure)
#pragma once
#include "fmatrix.hh"
namespace Dune {
template <class T, int r, int c>
FieldMatrix<T, c, r> transpose(const FieldMatrix<T, r, c>& a) {
  FieldMatrix<T, c, r> t;
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) t[j][i] = a[i][j];
  return t;
}
}  // namespace Dune

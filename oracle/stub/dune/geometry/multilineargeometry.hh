// oracle-build stand-in for the pieces of dune-geometry the hot path touches (test infrastructure).
// dune-geometry is NOT vendored in the reference (pinned <= 2.9.0, pyproject.toml:5); sqrtDetAAT / rightInvA
// are restated here in closed form (|det| for square, sqrt(det(A A^T)) otherwise) -- "parity unpinned"
// pointwise by the reference's own tests, pinned only through integral checks.
#pragma once
#include "../common/fmatrix.hh"
#include "../common/math.hh"
namespace Dune {
class GeometryType {
  int dim_  = 0;
  bool none_ = false;

public:
  GeometryType() = default;
  GeometryType(int d, bool none) : dim_(d), none_(none) {}
  int dim() const { return dim_; }
  bool isNone() const { return none_; }
  bool operator==(const GeometryType& o) const { return dim_ == o.dim_ && none_ == o.none_; }
};
namespace GeometryTypes {
inline GeometryType cube(int d) { return GeometryType(d, false); }
inline GeometryType none(int d) { return GeometryType(d, true); }
}  // namespace GeometryTypes

template <class ct>
struct MultiLinearGeometryTraits {
  struct MatrixHelper {
    template <int m, int n>
    static void AAT(const FieldMatrix<ct, m, n>& A, FieldMatrix<ct, m, m>& ret) {
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
          ct s = 0;
          for (int k = 0; k < n; ++k) s += A[i][k] * A[j][k];
          ret[i][j] = s;
        }
    }
    template <int m, int n>
    static ct sqrtDetAAT(const FieldMatrix<ct, m, n>& A) {
      if constexpr (m == n) {
        return std::abs(A.determinant());
      } else if constexpr (m == 1) {
        ct s = 0;
        for (int k = 0; k < n; ++k) s += A[0][k] * A[0][k];
        return std::sqrt(s);
      } else if constexpr (m == 2 && n == 3) {
        const ct v0 = A[0][0] * A[1][1] - A[0][1] * A[1][0];
        const ct v1 = A[0][0] * A[1][2] - A[1][0] * A[0][2];
        const ct v2 = A[0][1] * A[1][2] - A[0][2] * A[1][1];
        return std::sqrt(v0 * v0 + v1 * v1 + v2 * v2);
      } else {
        FieldMatrix<ct, m, m> aat;
        AAT(A, aat);
        return std::sqrt(aat.determinant());
      }
    }
    // ret = A^T (A A^T)^-1  (n x m); returns sqrt(det(A A^T))
    template <int m, int n>
    static ct rightInvA(const FieldMatrix<ct, m, n>& A, FieldMatrix<ct, n, m>& ret) {
      FieldMatrix<ct, m, m> aat;
      AAT(A, aat);
      // invert aat column by column
      FieldMatrix<ct, m, m> inv;
      for (int j = 0; j < m; ++j) {
        FieldVector<ct, m> e(0), x;
        e[j] = 1;
        aat.solve(x, e);
        for (int i = 0; i < m; ++i) inv[i][j] = x[i];
      }
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) {
          ct s = 0;
          for (int k = 0; k < m; ++k) s += A[k][i] * inv[k][j];
          ret[i][j] = s;
        }
      return sqrtDetAAT<m, n>(A);
    }
    // y = (A A^T)^-1 A x ; returns false if not invertible
    template <int m, int n>
    static bool xTRightInvA(const FieldMatrix<ct, m, n>& A, const FieldVector<ct, n>& x, FieldVector<ct, m>& y) {
      FieldMatrix<ct, m, m> aat;
      AAT(A, aat);
      FieldVector<ct, m> rhs;
      A.mv(x, rhs);
      if constexpr (m <= 3) {
        if (aat.determinant() == ct(0)) return false;
      }
      aat.solve(y, rhs);
      return true;
    }
  };
};
}  // namespace Dune

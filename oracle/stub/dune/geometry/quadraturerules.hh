// oracle-build stand-in for dune-geometry's QuadratureRules (test infrastructure).
// Gauss-Legendre tensor rules on [0,1]^dim computed by Newton iteration on Legendre polynomials; dune-geometry's
// tabulated points are not in the reference tree ("parity unpinned") -- the C-ABI therefore takes
// abscissae/weights as inputs.
#pragma once
#include "multilineargeometry.hh"
#include <map>
namespace Dune {
template <class ct, int dim>
class QuadraturePoint {
  FieldVector<ct, dim> x_;
  ct w_;

public:
  QuadraturePoint(const FieldVector<ct, dim>& x, ct w) : x_(x), w_(w) {}
  const FieldVector<ct, dim>& position() const { return x_; }
  const ct& weight() const { return w_; }
};
namespace QuadratureType {
enum Enum { GaussLegendre = 0, GaussLobatto = 4 };
}
template <class ct, int dim>
class QuadratureRule : public std::vector<QuadraturePoint<ct, dim>> {
public:
  QuadratureRule() = default;
};
namespace StubImpl {
inline void gaussLegendre01(int n, std::vector<double>& x, std::vector<double>& w) {
  x.resize(n);
  w.resize(n);
  for (int i = 0; i < n; ++i) {
    double z = std::cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1, p2 = 0;
      for (int j = 1; j <= n; ++j) {
        double p3 = p2;
        p2        = p1;
        p1        = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
      }
      pp        = n * (z * p1 - p2) / (z * z - 1.0);
      double dz = p1 / pp;
      z -= dz;
      if (std::abs(dz) < 1e-16) break;
    }
    x[n - 1 - i] = 0.5 * (z + 1.0);
    w[n - 1 - i] = 1.0 / ((1.0 - z * z) * pp * pp);
  }
}
}  // namespace StubImpl
template <class ct, int dim>
class QuadratureRules {
public:
  static const QuadratureRule<ct, dim>& rule(const GeometryType&, int order,
                                             QuadratureType::Enum = QuadratureType::GaussLegendre) {
    static std::map<int, QuadratureRule<ct, dim>> cache;
    auto it = cache.find(order);
    if (it != cache.end()) return it->second;
    const int n = order / 2 + 1;
    std::vector<double> x, w;
    StubImpl::gaussLegendre01(n, x, w);
    QuadratureRule<ct, dim> r;
    int total = 1;
    for (int d = 0; d < dim; ++d) total *= n;
    for (int f = 0; f < total; ++f) {
      FieldVector<ct, dim> p;
      ct ww = 1;
      int h = f;
      for (int d = 0; d < dim; ++d) {
        p[d] = x[h % n];
        ww *= w[h % n];
        h /= n;
      }
      r.emplace_back(p, ww);
    }
    return cache.emplace(order, r).first->second;
  }
};
}  // namespace Dune

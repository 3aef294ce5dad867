// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, never linked into or called by the product path.
//
// C-ABI window onto the REFERENCE'S OWN headers (compiled where they lie under /root/reference against the
// name stub in oracle/stub/, see oracle/Makefile).  Nothing of the reference is copied into this repository;
// this file only *calls* it:
//   dune/iga/bsplinealgorithms.hh  findSpanCorrected :24-32, BsplineBasis1D::basisFunctions :96-137,
//                                  basisFunctionDerivatives :148-222
//   dune/iga/nurbsalgorithms.hh    netOfSpan :80-89, extractWeights :103-108, extractControlCoordinates :110-114,
//                                  Nurbs<dim>::basisFunctions :160-180, basisFunctionDerivatives :193-250,
//                                  degreeElevate :264-439, knotRefinement :441-528, generateRefinedKnots :49-66,
//                                  makeCircularArc :579-633, makeSurfaceOfRevolution :636-705
//   dune/iga/nurbspatchgeometry.hh global :76-82, jacobianTransposed :163-180, integrationElement :182-185,
//                                  zeroFirstAndSecondDerivativeOfPosition :84-117, volume :49-56
//
// What is RESTATED here because the header that holds it cannot be compiled without full DUNE
// (nurbsgeometry.hh, nurbspatch.hh, nurbsbasis.hh pull dune-grid / dune-functions / Clipper2):
//   * element -> span:  nurbspatch.hh:248-253 / nurbsbasis.hh:355-365  (findSpanCorrected(p, uniqueKnot[e], U, e))
//   * element-local scaling: nurbsgeometry.hh:64-68,189-194,262-289,365-378 and nurbsbasis.hh:88-95,101-107
// The per-point arithmetic in oref_eval_* is the reference's: Nurbs<dim>::basisFunctions /
// basisFunctionDerivatives with the span passed explicitly, netOfSpan, dot -- exactly what NURBSGeometry does after
// construction (nurbsgeometry.hh:133-138,182-203,304-354).  This is the "reference kernel path" CPU baseline.
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstring>
#include <memory>
#include <numeric>
#include <optional>
#include <tuple>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include <dune/common/fmatrix.hh>
#include <dune/common/fvector.hh>
#include <dune/common/math.hh>
#include <dune/iga/nurbsalgorithms.hh>
#include <dune/iga/nurbspatchgeometry.hh>

using namespace Dune;
using namespace Dune::IGA;

namespace {

struct PatchBase {
  int dimRt, dimworldRt;  // run-time copies of the template parameters
  virtual ~PatchBase() = default;
};

template <int dim, int dw>
struct PatchT : PatchBase {
  using Data = NURBSPatchData<(std::size_t)dim, (std::size_t)dw, double>;
  std::shared_ptr<Data> data;
  // caches for the element path (what NURBSGeometry / NurbsPreBasis hold after construction)
  MultiDimensionNet<dim, double> weights;
  MultiDimensionNet<dim, FieldVector<double, dw>> coords;
  std::array<std::vector<double>, dim> uniqueKnots;
  std::array<std::vector<int>, dim> spanOfElement;
  explicit PatchT(const Data& d) : data(std::make_shared<Data>(d)) {
    this->dimRt      = dim;
    this->dimworldRt = dw;
    weights        = extractWeights(data->controlPoints);
    coords         = extractControlCoordinates(data->controlPoints);
    for (int i = 0; i < dim; ++i) {
      // unique knots as in nurbspatch.hh:58-63 / nurbsbasis.hh:511-515
      std::ranges::unique_copy(data->knotSpans[i], std::back_inserter(uniqueKnots[i]),
                               [](auto& l, auto& r) { return Dune::FloatCmp::eq(l, r); });
      const int nel = (int)uniqueKnots[i].size() - 1;
      spanOfElement[i].resize(nel);
      for (int e = 0; e < nel; ++e)
        spanOfElement[i][e] = (int)findSpanCorrected(data->degree[i], uniqueKnots[i][e], data->knotSpans[i], e);
    }
  }
  int nb() const {
    int r = 1;
    for (int i = 0; i < dim; ++i) r *= data->degree[i] + 1;
    return r;
  }
  long nElements() const {
    long r = 1;
    for (int i = 0; i < dim; ++i) r *= (long)spanOfElement[i].size();
    return r;
  }
};

template <class F>
auto dispatch(PatchBase* p, F&& f) {
  const int key = p->dimRt * 10 + p->dimworldRt;
  switch (key) {
    case 11: return f(static_cast<PatchT<1, 1>*>(p));
    case 12: return f(static_cast<PatchT<1, 2>*>(p));
    case 13: return f(static_cast<PatchT<1, 3>*>(p));
    case 22: return f(static_cast<PatchT<2, 2>*>(p));
    case 23: return f(static_cast<PatchT<2, 3>*>(p));
    case 33: return f(static_cast<PatchT<3, 3>*>(p));
    default: throw std::logic_error("oracle: unsupported (dim,dimworld)");
  }
}

template <int dim, int dw>
PatchBase* makePatch(const int* degree, const int* nknots, const double* knots, const int* ncp, const double* cp) {
  using P = PatchT<dim, dw>;
  typename P::Data d;
  std::array<int, dim> sizes;
  long off = 0, total = 1;
  for (int i = 0; i < dim; ++i) {
    d.degree[i] = degree[i];
    d.knotSpans[i].assign(knots + off, knots + off + nknots[i]);
    off += nknots[i];
    sizes[i] = ncp[i];
    total *= ncp[i];
  }
  typename P::Data::ControlPointNetType net(sizes);
  for (long k = 0; k < total; ++k) {
    auto& c = net.directGet((int)k);
    for (int j = 0; j < dw; ++j) c.p[j] = cp[k * (dw + 1) + j];
    c.w = cp[k * (dw + 1) + dw];
  }
  d.controlPoints = net;
  return new P(d);
}

// H row order of the reference: pure second derivatives first, then mixed in next_permutation(greater) order
// (nurbsgeometry.hh:268-289): 2-D [uu,vv,uv]; 3-D [uu,vv,ww,(1,1,0),(1,0,1),(0,1,1)]
template <int dim>
std::vector<std::array<int, dim>> secondOrderMultiIndices() {
  std::vector<std::array<int, dim>> r;
  for (int d = 0; d < dim; ++d) {
    std::array<int, dim> a{};
    a[d] = 2;
    r.push_back(a);
  }
  if constexpr (dim > 1) {
    std::array<int, dim> m{};
    std::ranges::fill_n(m.begin(), 2, 1);
    do {
      r.push_back(m);
    } while (std::ranges::next_permutation(m, std::greater()).found);
  }
  return r;
}

template <int dim, int dw>
void evalOnePoint(const PatchT<dim, dw>* P, const std::array<int, dim>& span, const std::array<double, dim>& scale,
                  const std::array<double, dim>& offs, const MultiDimensionNet<dim, FieldVector<double, dw>>& cpNet,
                  const double* xi, int order, double* N, double* dN, double* d2N, double* x, double* Jt, double* H,
                  double* detJ) {
  const auto& D = *P->data;
  const int nb  = P->nb();
  FieldVector<double, dim> u;
  for (int d = 0; d < dim; ++d) u[d] = xi[d] * scale[d] + offs[d];  // nurbsgeometry.hh:365-378
  // values: A2.2 path, as global()/evaluateFunction use (nurbsgeometry.hh:133-138, nurbsbasis.hh:637-643)
  if (N || x) {
    auto R = Nurbs<dim>::basisFunctions(u, D.knotSpans, D.degree, P->weights, span);
    if (N)
      for (int i = 0; i < nb; ++i) N[i] = R.directGet(i);
    if (x) {
      auto pos = dot(R, cpNet);
      for (int j = 0; j < dw; ++j) x[j] = pos[j];
    }
  }
  if (order >= 1 && (dN || Jt || detJ || d2N || H)) {
    const int k = (order >= 2 && (d2N || H)) ? 2 : 1;
    auto dR     = Nurbs<dim>::basisFunctionDerivatives(u, D.knotSpans, D.degree, P->weights, k, false, span);
    FieldMatrix<double, dim, dw> J;
    for (int d = 0; d < dim; ++d) {
      std::array<int, dim> a{};
      a[d]           = 1;
      const auto& Rd = dR.get(a);
      if (dN)
        for (int i = 0; i < nb; ++i) dN[i * dim + d] = Rd.directGet(i) * scale[d];  // nurbsbasis.hh:93-95
      J[d] = dot(Rd, cpNet);
      J[d] *= scale[d];  // nurbsgeometry.hh:189-194
    }
    if (Jt)
      for (int d = 0; d < dim; ++d)
        for (int j = 0; j < dw; ++j) Jt[d * dw + j] = J[d][j];
    if (detJ) *detJ = MultiLinearGeometryTraits<double>::MatrixHelper::template sqrtDetAAT<dim, dw>(J);
    if (k == 2) {
      static const auto idx2 = secondOrderMultiIndices<dim>();
      for (std::size_t r = 0; r < idx2.size(); ++r) {
        double s = 1;
        for (int d = 0; d < dim; ++d) s *= Dune::power(scale[d], idx2[r][d]);  // nurbsbasis.hh:105-107
        const auto& Rr = dR.get(idx2[r]);
        if (d2N)
          for (int i = 0; i < nb; ++i) d2N[i * (int)idx2.size() + (int)r] = Rr.directGet(i) * s;
        if (H) {
          auto h = dot(Rr, cpNet);
          h *= s;  // nurbsgeometry.hh:262-289
          for (int j = 0; j < dw; ++j) H[r * dw + j] = h[j];
        }
      }
    }
  }
}

template <class PT, std::size_t dimS>
void elementSetup(const PT* P, long e, std::array<int, dimS>& span, std::array<double, dimS>& scale,
                  std::array<double, dimS>& offs) {
  long h = e;
  for (int d = 0; d < (int)dimS; ++d) {
    const int ne = (int)P->spanOfElement[d].size();
    const int ed = (int)(h % ne);  // x-fastest, mdnet.hh:245-257 / nurbsbasis.hh:683-691
    h /= ne;
    span[d]        = P->spanOfElement[d][ed];
    const auto& U  = P->data->knotSpans[d];
    scale[d]       = U[span[d] + 1] - U[span[d]];  // nurbsgeometry.hh:64-68
    offs[d]        = U[span[d]];
  }
}

}  // namespace

extern "C" {

void* oref_patch_create(int dim, int dimworld, const int* degree, const int* nknots, const double* knots,
                        const int* ncp, const double* cp) {
  try {
    switch (dim * 10 + dimworld) {
      case 11: return makePatch<1, 1>(degree, nknots, knots, ncp, cp);
      case 12: return makePatch<1, 2>(degree, nknots, knots, ncp, cp);
      case 13: return makePatch<1, 3>(degree, nknots, knots, ncp, cp);
      case 22: return makePatch<2, 2>(degree, nknots, knots, ncp, cp);
      case 23: return makePatch<2, 3>(degree, nknots, knots, ncp, cp);
      case 33: return makePatch<3, 3>(degree, nknots, knots, ncp, cp);
    }
  } catch (...) {
  }
  return nullptr;
}

void oref_patch_destroy(void* h) { delete static_cast<PatchBase*>(h); }

// sizes: degree[dim], nknots[dim], ncp[dim]
void oref_patch_info(void* h, int* dim, int* dimworld, int* degree, int* nknots, int* ncp, long* nelem) {
  dispatch(static_cast<PatchBase*>(h), [&](auto* P) {
    *dim      = P->dimRt;
    *dimworld = P->dimworldRt;
    auto ss   = P->data->controlPoints.strideSizes();
    for (int i = 0; i < P->dimRt; ++i) {
      degree[i] = P->data->degree[i];
      nknots[i] = (int)P->data->knotSpans[i].size();
      ncp[i]    = ss[i];
    }
    *nelem = P->nElements();
    return 0;
  });
}

void oref_patch_get(void* h, double* knots, double* cp) {
  dispatch(static_cast<PatchBase*>(h), [&](auto* P) {
    long off = 0;
    for (int i = 0; i < P->dimRt; ++i)
      for (double v : P->data->knotSpans[i]) knots[off++] = v;
    const auto& all = P->data->controlPoints.directGetAll();
    const int dw    = P->dimworldRt;
    for (std::size_t k = 0; k < all.size(); ++k) {
      for (int j = 0; j < dw; ++j) cp[k * (dw + 1) + j] = all[k].p[j];
      cp[k * (dw + 1) + dw] = all[k].w;
    }
    return 0;
  });
}

void oref_patch_spans(void* h, int* span /* [nelem][dim] */) {
  dispatch(static_cast<PatchBase*>(h), [&](auto* P) {
    constexpr int dim = std::remove_pointer_t<decltype(P)>::Data::patchDim;
    const long n      = P->nElements();
    for (long e = 0; e < n; ++e) {
      std::array<int, dim> s;
      std::array<double, dim> sc, of;
      elementSetup(P, e, s, sc, of);
      for (int d = 0; d < dim; ++d) span[e * dim + d] = s[d];
    }
    return 0;
  });
}

void* oref_patch_knot_refine(void* h, int dir, const double* newKnots, int n) {
  return dispatch(static_cast<PatchBase*>(h), [&](auto* P) -> void* {
    using PT = std::remove_pointer_t<decltype(P)>;
    std::vector<double> nk(newKnots, newKnots + n);
    return static_cast<PatchBase*>(new PT(knotRefinement(*P->data, nk, dir)));
  });
}

// generateRefinedKnots + knotRefinement, as NURBSGrid::globalRefineInDirection does (nurbsgrid.hh:143-149)
void* oref_patch_refine_uniform(void* h, int dir, int level) {
  return dispatch(static_cast<PatchBase*>(h), [&](auto* P) -> void* {
    using PT = std::remove_pointer_t<decltype(P)>;
    auto nk  = generateRefinedKnots(P->data->knotSpans, dir, level);
    return static_cast<PatchBase*>(new PT(knotRefinement(*P->data, nk, dir)));
  });
}

void* oref_patch_degree_elevate(void* h, int dir, int t) {
  return dispatch(static_cast<PatchBase*>(h), [&](auto* P) -> void* {
    using PT = std::remove_pointer_t<decltype(P)>;
    return static_cast<PatchBase*>(new PT(degreeElevate(*P->data, dir, t)));
  });
}

void* oref_make_circular_arc(double radius, double startAngle, double endAngle) {
  return static_cast<PatchBase*>(new PatchT<1, 3>(makeCircularArc(radius, startAngle, endAngle)));
}

void* oref_make_surface_of_revolution(void* generatrix, const double* point, const double* axis, double angle) {
  auto* g = static_cast<PatchT<1, 3>*>(static_cast<PatchBase*>(generatrix));
  FieldVector<double, 3> p{point[0], point[1], point[2]}, a{axis[0], axis[1], axis[2]};
  return static_cast<PatchBase*>(new PatchT<2, 3>(makeSurfaceOfRevolution(*g->data, p, a, angle)));
}

long oref_find_span(int p, double u, const double* U, int n, int offset) {
  std::vector<double> k(U, U + n);
  return findSpanCorrected(p, u, k, offset);
}

void oref_bspline_basis(double u, const double* U, int n, int p, int span, double* out) {
  std::vector<double> k(U, U + n);
  auto N = BsplineBasis1D<double>::basisFunctions(u, k, p, span >= 0 ? std::optional<int>(span) : std::nullopt);
  for (int i = 0; i <= p; ++i) out[i] = N[i];
}

void oref_bspline_ders(double u, const double* U, int n, int p, int k, int span, double* out /* [(k+1)][p+1] */) {
  std::vector<double> kn(U, U + n);
  auto dN =
      BsplineBasis1D<double>::basisFunctionDerivatives(u, kn, p, k, span >= 0 ? std::optional<int>(span) : std::nullopt);
  for (int r = 0; r <= k; ++r)
    for (int i = 0; i <= p; ++i) out[r * (p + 1) + i] = dN[r][i];
}

// patch-coordinate NURBS basis (span searched when span==nullptr)
void oref_nurbs_basis(void* h, const double* u, const int* span, double* out) {
  dispatch(static_cast<PatchBase*>(h), [&](auto* P) {
    constexpr int dim = std::remove_pointer_t<decltype(P)>::Data::patchDim;
    FieldVector<double, dim> uu;
    std::array<int, dim> s{};
    for (int d = 0; d < dim; ++d) {
      uu[d] = u[d];
      if (span) s[d] = span[d];
    }
    auto R = Nurbs<dim>::basisFunctions(uu, P->data->knotSpans, P->data->degree, P->weights,
                                        span ? std::optional(s) : std::nullopt);
    for (int i = 0; i < R.size(); ++i) out[i] = R.directGet(i);
    return 0;
  });
}

// all mixed derivatives alpha in {0..k}^dim; out[(k+1)^dim][nb], alpha flat index first-direction-fastest
void oref_nurbs_ders(void* h, const double* u, int k, const int* span, double* out) {
  dispatch(static_cast<PatchBase*>(h), [&](auto* P) {
    constexpr int dim = std::remove_pointer_t<decltype(P)>::Data::patchDim;
    FieldVector<double, dim> uu;
    std::array<int, dim> s{};
    for (int d = 0; d < dim; ++d) {
      uu[d] = u[d];
      if (span) s[d] = span[d];
    }
    auto R = Nurbs<dim>::basisFunctionDerivatives(uu, P->data->knotSpans, P->data->degree, P->weights, k, false,
                                                  span ? std::optional(s) : std::nullopt);
    const int nb = P->nb();
    for (int a = 0; a < R.size(); ++a)
      for (int i = 0; i < nb; ++i) out[(long)a * nb + i] = R.directGet(a).directGet(i);
    return 0;
  });
}

void oref_patchgeo_global(void* h, const double* u, double* out) {
  dispatch(static_cast<PatchBase*>(h), [&](auto* P) {
    constexpr int dim = std::remove_pointer_t<decltype(P)>::Data::patchDim;
    constexpr int dw  = std::remove_pointer_t<decltype(P)>::Data::dimworld;
    NURBSPatchGeometry<dim, dw> geo(P->data);
    FieldVector<double, dim> uu;
    for (int d = 0; d < dim; ++d) uu[d] = u[d];
    auto x = geo.global(uu);
    for (int j = 0; j < dw; ++j) out[j] = x[j];
    return 0;
  });
}

void oref_patchgeo_jt(void* h, const double* u, double* out) {
  dispatch(static_cast<PatchBase*>(h), [&](auto* P) {
    constexpr int dim = std::remove_pointer_t<decltype(P)>::Data::patchDim;
    constexpr int dw  = std::remove_pointer_t<decltype(P)>::Data::dimworld;
    NURBSPatchGeometry<dim, dw> geo(P->data);
    FieldVector<double, dim> uu;
    for (int d = 0; d < dim; ++d) uu[d] = u[d];
    auto J = geo.jacobianTransposed(uu);
    for (int d = 0; d < dim; ++d)
      for (int j = 0; j < dw; ++j) out[d * dw + j] = J[d][j];
    return 0;
  });
}

double oref_patchgeo_detj(void* h, const double* u) {
  return dispatch(static_cast<PatchBase*>(h), [&](auto* P) {
    constexpr int dim = std::remove_pointer_t<decltype(P)>::Data::patchDim;
    constexpr int dw  = std::remove_pointer_t<decltype(P)>::Data::dimworld;
    NURBSPatchGeometry<dim, dw> geo(P->data);
    FieldVector<double, dim> uu;
    for (int d = 0; d < dim; ++d) uu[d] = u[d];
    return geo.integrationElement(uu);
  });
}

void oref_patchgeo_012(void* h, const double* u, double* pos, double* Jt, double* H) {
  dispatch(static_cast<PatchBase*>(h), [&](auto* P) {
    constexpr int dim = std::remove_pointer_t<decltype(P)>::Data::patchDim;
    constexpr int dw  = std::remove_pointer_t<decltype(P)>::Data::dimworld;
    NURBSPatchGeometry<dim, dw> geo(P->data);
    FieldVector<double, dim> uu;
    for (int d = 0; d < dim; ++d) uu[d] = u[d];
    auto [x, J, Hm] = geo.zeroFirstAndSecondDerivativeOfPosition(uu);
    for (int j = 0; j < dw; ++j) pos[j] = x[j];
    for (int d = 0; d < dim; ++d)
      for (int j = 0; j < dw; ++j) Jt[d * dw + j] = J[d][j];
    for (int r = 0; r < dim * (dim + 1) / 2; ++r)
      for (int j = 0; j < dw; ++j) H[r * dw + j] = Hm[r][j];
    return 0;
  });
}

// NURBSPatchGeometry::local (dune/iga/nurbspatchgeometry.hh:126-158): Newton for dim == dimworld, the trust-region
// closest-point projection (dune/iga/geometry/closestpointprojection.hh) for curves and surfaces in a larger space.
// Returns 0, or 1 when the reference threw (Dune::MathError: no energy-decreasing direction / singular matrix).
int oref_patchgeo_local(void* h, const double* xglobal, double* u) {
  return dispatch(static_cast<PatchBase*>(h), [&](auto* P) {
    constexpr int dim = std::remove_pointer_t<decltype(P)>::Data::patchDim;
    constexpr int dw  = std::remove_pointer_t<decltype(P)>::Data::dimworld;
    if constexpr (dim == dw || dim <= 2) {
      NURBSPatchGeometry<dim, dw> geo(P->data);
      FieldVector<double, dw> g;
      for (int j = 0; j < dw; ++j) g[j] = xglobal[j];
      try {
        auto r = geo.local(g);
        for (int d = 0; d < dim; ++d) u[d] = r[d];
      } catch (...) {
        return 1;
      }
      return 0;
    } else
      return 2;
  });
}

double oref_patchgeo_volume(void* h, int scaleOrder) {
  return dispatch(static_cast<PatchBase*>(h), [&](auto* P) {
    constexpr int dim = std::remove_pointer_t<decltype(P)>::Data::patchDim;
    constexpr int dw  = std::remove_pointer_t<decltype(P)>::Data::dimworld;
    NURBSPatchGeometry<dim, dw> geo(P->data);
    return geo.volume(scaleOrder);
  });
}

// Element-batched evaluation at a tensor rule: the reference kernel path, one scalar call chain per point,
// OpenMP over elements.  Layouts (all i-fastest, MultiDimensionNet order):
//   N[pt][nb], dN[pt][nb][dim], d2N[pt][nb][nh], x[pt][dw], Jt[pt][dim][dw], H[pt][nh][dw], detJ[pt]
//   pt = (e - elem_begin) * nq + q,  q flat over the tensor rule, first direction fastest;  nh = dim(dim+1)/2
// Any output pointer may be null.  Returns number of threads used.
int oref_eval_tensor(void* h, long elem_begin, long elem_end, int order, const int* nq1, const double* xi1d,
                     double* N, double* dN, double* d2N, double* x, double* Jt, double* H, double* detJ,
                     int nthreads) {
  int used = 1;
  dispatch(static_cast<PatchBase*>(h), [&](auto* P) {
    constexpr int dim = std::remove_pointer_t<decltype(P)>::Data::patchDim;
    constexpr int dw  = std::remove_pointer_t<decltype(P)>::Data::dimworld;
    constexpr int nh  = dim * (dim + 1) / 2;
    const int nb      = P->nb();
    long nq           = 1;
    std::array<const double*, dim> xi;
    {
      long off = 0;
      for (int d = 0; d < dim; ++d) {
        xi[d] = xi1d + off;
        off += nq1[d];
        nq *= nq1[d];
      }
    }
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    used = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(static)
    for (long e = elem_begin; e < elem_end; ++e) {
      std::array<int, dim> span;
      std::array<double, dim> scale, offs;
      elementSetup(P, e, span, scale, offs);
      // per-element control sub-net, as cached by the NURBSGeometry ctor (nurbsgeometry.hh:83-84)
      const auto cpNet = netOfSpan(span, P->data->degree, P->coords);
      for (long q = 0; q < nq; ++q) {
        double loc[dim];
        long hq = q;
        for (int d = 0; d < dim; ++d) {
          loc[d] = xi[d][hq % nq1[d]];
          hq /= nq1[d];
        }
        const long pt = (e - elem_begin) * nq + q;
        evalOnePoint<dim, dw>(P, span, scale, offs, cpNet, loc, order, N ? N + pt * nb : nullptr,
                              dN ? dN + pt * nb * dim : nullptr, d2N ? d2N + pt * nb * nh : nullptr,
                              x ? x + pt * dw : nullptr, Jt ? Jt + pt * dim * dw : nullptr,
                              H ? H + pt * nh * dw : nullptr, detJ ? detJ + pt : nullptr);
      }
    }
    return 0;
  });
  return used;
}

// Arbitrary element-local points (trimmed batches): elem[npts], xi[npts][dim]
int oref_eval_points(void* h, long npts, const int* elem, const double* xi, int order, double* N, double* dN,
                     double* d2N, double* x, double* Jt, double* H, double* detJ, int nthreads) {
  int used = 1;
  dispatch(static_cast<PatchBase*>(h), [&](auto* P) {
    constexpr int dim = std::remove_pointer_t<decltype(P)>::Data::patchDim;
    constexpr int dw  = std::remove_pointer_t<decltype(P)>::Data::dimworld;
    constexpr int nh  = dim * (dim + 1) / 2;
    const int nb      = P->nb();
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    used = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(static)
    for (long pt = 0; pt < npts; ++pt) {
      std::array<int, dim> span;
      std::array<double, dim> scale, offs;
      elementSetup(P, elem[pt], span, scale, offs);
      const auto cpNet = netOfSpan(span, P->data->degree, P->coords);
      evalOnePoint<dim, dw>(P, span, scale, offs, cpNet, xi + pt * dim, order, N ? N + pt * nb : nullptr,
                            dN ? dN + pt * nb * dim : nullptr, d2N ? d2N + pt * nb * nh : nullptr,
                            x ? x + pt * dw : nullptr, Jt ? Jt + pt * dim * dw : nullptr,
                            H ? H + pt * nh * dw : nullptr, detJ ? detJ + pt : nullptr);
    }
    return 0;
  });
  return used;
}

// partial(alpha) at element-local points, as NurbsLocalBasis::partial (nurbsbasis.hh:99-108,667-677):
// derivative order passed to the reference = sum(alpha); result scaled by prod scale_d^alpha_d.
void oref_partial(void* h, long npts, const int* elem, const double* xi, const int* alpha, double* out /*[npts][nb]*/) {
  dispatch(static_cast<PatchBase*>(h), [&](auto* P) {
    constexpr int dim = std::remove_pointer_t<decltype(P)>::Data::patchDim;
    const int nb      = P->nb();
    int k             = 0;
    std::array<int, dim> a;
    for (int d = 0; d < dim; ++d) {
      a[d] = alpha[d];
      k += alpha[d];
    }
    for (long pt = 0; pt < npts; ++pt) {
      std::array<int, dim> span;
      std::array<double, dim> scale, offs;
      elementSetup(P, elem[pt], span, scale, offs);
      FieldVector<double, dim> u;
      for (int d = 0; d < dim; ++d) u[d] = xi[pt * dim + d] * scale[d] + offs[d];
      auto dR  = Nurbs<dim>::basisFunctionDerivatives(u, P->data->knotSpans, P->data->degree, P->weights, k, false, span);
      double s = 1;
      for (int d = 0; d < dim; ++d) s *= Dune::power(scale[d], a[d]);
      const auto& R = dR.get(a);
      for (int i = 0; i < nb; ++i) out[pt * nb + i] = R.directGet(i) * s;
    }
    return 0;
  });
}

// The per-point path AS SHIPPED: what a dune-functions / dune-grid client pays per quadrature point when
// it calls localBasis.evaluateFunction + evaluateJacobian (nurbsbasis.hh:637-663: extractWeights of the WHOLE patch per
// call) and the patch geometry's global + jacobianTransposed + integrationElement (nurbspatchgeometry.hh:76-82,163-185:
// extractControlCoordinates of the whole patch per call).  O(N_cp) copies per point; OpenMP over points.
// u: [npts][dim] patch coordinates.  acc: a checksum so that nothing is optimised away.  Returns threads used.
int oref_as_shipped_points(void* h, long npts, const double* u, double* acc, int nthreads) {
  int used = 1;
  dispatch(static_cast<PatchBase*>(h), [&](auto* P) {
    constexpr int dim = std::remove_pointer_t<decltype(P)>::Data::patchDim;
    constexpr int dw  = std::remove_pointer_t<decltype(P)>::Data::dimworld;
    const auto& D     = *P->data;
    double total      = 0.0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel reduction(+ : total)
#endif
    {
#ifdef _OPENMP
#pragma omp single
      used = omp_get_num_threads();
#pragma omp for schedule(static)
#endif
      for (long k = 0; k < npts; ++k) {
        FieldVector<double, dim> uu;
        for (int d = 0; d < dim; ++d) uu[d] = u[k * dim + d];
        const auto span = findSpanCorrected(D.degree, uu, D.knotSpans);
        const auto N    = Nurbs<dim>::basisFunctions(uu, D.knotSpans, D.degree, extractWeights(D.controlPoints), span);
        const auto dN   = Nurbs<dim>::basisFunctionDerivatives(uu, D.knotSpans, D.degree,
                                                               extractWeights(D.controlPoints), 1, false, span);
        NURBSPatchGeometry<dim, dw> geo(P->data);
        const auto x  = geo.global(uu);
        const auto J  = geo.jacobianTransposed(uu);
        const double dj = geo.integrationElement(uu);
        std::array<int, dim> e0{};
        e0[0] = 1;
        total += N.directGet(0) + dN.get(e0).directGet(0) + x[0] + J[0][0] + dj;
      }
    }
    if (acc) *acc = total;
    return 0;
  });
  return used;
}

int oref_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"

// local.cu -- batched inverse of the element map, NURBSGeometry::local (dune/iga/nurbsgeometry.hh:145-176), one thread
// per query point.
//   dim == dimworld : Newton from xi = 0.5 with the step dx = (Jt Jt^T)^-1 Jt (x(xi) - X) (MatrixHelper::xTRightInvA),
//                     clamp to [0,1]^dim, stop when every coordinate sits on the boundary
//                     (Utilities::clampToBoundaryAndCheckIfIsAtAllBoundaries, dune/iga/geometry/geohelper.hh:30-45) or
//                     |dx|^2 <= 16 eps; a singular Jacobian yields numeric_limits<double>::max() in every coordinate.
//   dim <  dimworld : closestPointProjectionByTrustRegion (dune/iga/geometry/closestpointprojection.hh:13-170) for
//                     curves and surfaces: Newton on E(u) = |x(u) - X|^2 / 2 with H = J J^T + sum H_ab . (x - X), step
//                     limited to a quarter of the domain, halved (or reversed when H is not positive along it) while E
//                     grows, <= 100 iterations, |grad E| < 1e-10.
// The geometry at the iterate (x, J^T scaled by the span lengths, H) comes from the homogeneous sums A^(a) = sum N^(a) w P,
// W^(a) = sum N^(a) w and the quotient rule on the geometry (same arithmetic as the tensor kernels, sf_geometry.cuh),
// with the univariate rows from A2.3 (bspline_device.cuh).  Flags: 0 ok, 1 singular matrix, 2 no energy-decreasing
// direction (the reference throws Dune::MathError), 3 iteration cap of the Newton branch.
#include <cfloat>

#include "bspline_device.cuh"
#include "igab_internal.cuh"

namespace igab {

namespace {

template <int DIM>
struct Ctx {
  int sp[DIM];
  double scale[DIM], lo[DIM];
  long long cpBase, cpStride[DIM];
};

// multi-index component: 0 zero; 1..DIM e_d; DIM+1..2DIM 2e_d; then mixed (0,1),(0,2),(1,2)
template <int DIM>
__device__ __forceinline__ int acomp(int a, int d) {
  if (a == 0) return 0;
  if (a <= DIM) return (a - 1 == d) ? 1 : 0;
  if (a <= 2 * DIM) return (a - 1 - DIM == d) ? 2 : 0;
  const int m = a - 2 * DIM - 1;
  const int c0 = (m == 2) ? 1 : 0, c1 = (m == 0) ? 1 : 2;
  return (d == c0 || d == c1) ? 1 : 0;
}

template <int DIM, int DW, int ORDER>
__device__ void geom_at(const PatchDev& P, const Ctx<DIM>& c, const double* xi, double* X, double (*J)[DW],
                        double (*H)[DW]) {
  constexpr int NH = DIM * (DIM + 1) / 2;
  constexpr int NA = (ORDER == 0) ? 1 : (ORDER == 1 ? 1 + DIM : 1 + DIM + NH);
  double ders[DIM][(ORDER + 1) * kPM];
#pragma unroll
  for (int d = 0; d < DIM; ++d)
    ders1d_any(xi[d] * c.scale[d] + c.lo[d], P.knots[d], P.degree[d], c.sp[d], ORDER, ders[d]);
  double A[NA][DW + 1];
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int k = 0; k <= DW; ++k) A[a][k] = 0.0;
  int l[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) l[d] = 0;
  for (int i = 0; i < P.nb; ++i) {
    long long g = c.cpBase;
#pragma unroll
    for (int d = 0; d < DIM; ++d) g += l[d] * c.cpStride[d];
    double Pc[DW + 1];
#pragma unroll
    for (int k = 0; k <= DW; ++k) Pc[k] = P.cp[g * (DW + 1) + k];
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      double f = Pc[DW];
#pragma unroll
      for (int d = 0; d < DIM; ++d) f *= ders[d][acomp<DIM>(a, d) * kPM + l[d]];
#pragma unroll
      for (int k = 0; k < DW; ++k) A[a][k] += f * Pc[k];
      A[a][DW] += f;
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      if (++l[d] <= P.degree[d]) break;
      l[d] = 0;
    }
  }
  const double iW = 1.0 / A[0][DW];
#pragma unroll
  for (int k = 0; k < DW; ++k) X[k] = A[0][k] * iW;
  if constexpr (ORDER >= 1) {
#pragma unroll
    for (int d = 0; d < DIM; ++d)
#pragma unroll
      for (int k = 0; k < DW; ++k) J[d][k] = (A[1 + d][k] - A[1 + d][DW] * X[k]) * iW;
    if constexpr (ORDER >= 2) {
#pragma unroll
      for (int d = 0; d < DIM; ++d)
#pragma unroll
        for (int k = 0; k < DW; ++k)
          H[d][k] = (A[1 + DIM + d][k] - 2.0 * A[1 + d][DW] * J[d][k] - A[1 + DIM + d][DW] * X[k]) * iW * c.scale[d] * c.scale[d];
      int m = 0;
#pragma unroll
      for (int c0 = 0; c0 < DIM; ++c0)
#pragma unroll
        for (int c1 = c0 + 1; c1 < DIM; ++c1) {
          const int a = 1 + 2 * DIM + m;
#pragma unroll
          for (int k = 0; k < DW; ++k)
            H[DIM + m][k] = (A[a][k] - A[1 + c0][DW] * J[c1][k] - A[1 + c1][DW] * J[c0][k] - A[a][DW] * X[k]) * iW *
                            c.scale[c0] * c.scale[c1];
          ++m;
        }
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d)
#pragma unroll
      for (int k = 0; k < DW; ++k) J[d][k] *= c.scale[d];
  }
}

// y = A^-1 b for the symmetric / general small systems of the two branches; false when det == 0 or not finite
template <int N>
__device__ __forceinline__ bool solve_small(const double (*A)[N], const double* b, double* y) {
  if constexpr (N == 1) {
    if (A[0][0] == 0.0 || !isfinite(A[0][0])) return false;
    y[0] = b[0] / A[0][0];
  } else if constexpr (N == 2) {
    const double det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
    if (det == 0.0 || !isfinite(det)) return false;
    y[0] = (A[1][1] * b[0] - A[0][1] * b[1]) / det;
    y[1] = (A[0][0] * b[1] - A[1][0] * b[0]) / det;
  } else {
    const double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1], c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2],
                 c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
    const double det = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02;
    if (det == 0.0 || !isfinite(det)) return false;
    y[0] = (c00 * b[0] + (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * b[1] + (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * b[2]) / det;
    y[1] = (c01 * b[0] + (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * b[1] + (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * b[2]) / det;
    y[2] = (c02 * b[0] + (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * b[1] + (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * b[2]) / det;
  }
  return true;
}

template <int DIM>
__device__ __forceinline__ bool clamp_all_boundaries(double* u) {
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < DIM; ++j) {
    u[j] = fmin(fmax(u[j], 0.0), 1.0);
    if (u[j] == 0.0 || u[j] == 1.0) ++cnt;
  }
  return cnt == DIM;
}

template <int DIM, int DW>
__global__ void __launch_bounds__(64) local_kernel(PatchDev P, long long npts, const int* __restrict__ elem,
                                                   const double* __restrict__ xg, double* __restrict__ xiOut,
                                                   int* __restrict__ flagOut) {
  const long long pt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= npts) return;
  constexpr int NH = DIM * (DIM + 1) / 2;
  Ctx<DIM> c;
  {
    long long h = elem[pt], stride = 1;
    c.cpBase = 0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const int ed = (int)(h % P.nel[d]);
      h /= P.nel[d];
      c.sp[d] = P.spanTab[d][ed];
      c.lo[d] = P.knots[d][c.sp[d]];
      c.scale[d] = P.knots[d][c.sp[d] + 1] - c.lo[d];
      c.cpStride[d] = stride;
      c.cpBase += (long long)(c.sp[d] - P.degree[d]) * stride;
      stride *= P.ncp[d];
    }
  }
  double X[DW];
#pragma unroll
  for (int k = 0; k < DW; ++k) X[k] = xg[pt * DW + k];
  double u[DIM], x[DW], J[DIM][DW], H[NH][DW];
#pragma unroll
  for (int d = 0; d < DIM; ++d) u[d] = 0.5;
  int fl = 0;

  if constexpr (DIM == DW) {
    const double tol = 16.0 * DBL_EPSILON;
    for (int it = 0;; ++it) {
      geom_at<DIM, DW, 1>(P, c, u, x, J, H);
      double G[DIM][DIM], rhs[DIM], dx[DIM];
#pragma unroll
      for (int a = 0; a < DIM; ++a) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < DW; ++k) s += J[a][k] * (x[k] - X[k]);
        rhs[a] = s;
#pragma unroll
        for (int b = 0; b < DIM; ++b) {
          double q = 0.0;
#pragma unroll
          for (int k = 0; k < DW; ++k) q += J[a][k] * J[b][k];
          G[a][b] = q;
        }
      }
      if (!solve_small<DIM>(G, rhs, dx)) {
        fl = 1;
#pragma unroll
        for (int d = 0; d < DIM; ++d) u[d] = DBL_MAX;
        break;
      }
      double n2 = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        u[d] -= dx[d];
        n2 += dx[d] * dx[d];
      }
      if (clamp_all_boundaries<DIM>(u)) break;
      if (!(n2 > tol)) break;
      if (it >= 99) {
        fl = 3;
        break;
      }
    }
  } else {
    const double tol = 1e-10, domainFraction = 0.25;
    double R[DIM], HV[DIM][DIM];
    auto energy = [&](const double* uu) {
      geom_at<DIM, DW, 0>(P, c, uu, x, J, H);
      double e2 = 0.0;
#pragma unroll
      for (int k = 0; k < DW; ++k) e2 += (x[k] - X[k]) * (x[k] - X[k]);
      return 0.5 * e2;
    };
    auto gradHess = [&](const double* uu) {
      geom_at<DIM, DW, 2>(P, c, uu, x, J, H);
      double dist[DW];
#pragma unroll
      for (int k = 0; k < DW; ++k) dist[k] = x[k] - X[k];
#pragma unroll
      for (int a = 0; a < DIM; ++a) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < DW; ++k) s += J[a][k] * dist[k];
        R[a] = s;
#pragma unroll
        for (int b = 0; b < DIM; ++b) {
          double q = 0.0;
#pragma unroll
          for (int k = 0; k < DW; ++k) q += J[a][k] * J[b][k];
          HV[a][b] = q;
        }
      }
      double dd[NH];
#pragma unroll
      for (int r = 0; r < NH; ++r) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < DW; ++k) s += H[r][k] * dist[k];
        dd[r] = s;
      }
#pragma unroll
      for (int a = 0; a < DIM; ++a) HV[a][a] += dd[a];
      if constexpr (DIM == 2) {
        HV[0][1] += dd[2];
        HV[1][0] += dd[2];
      }
    };
    gradHess(u);
    for (int i = 0; i < 100; ++i) {
      double du[DIM], mR[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d) mR[d] = -R[d];
      if (!solve_small<DIM>(HV, mR, du)) {
        fl = 1;
        break;
      }
      double duNorm = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) duNorm += du[d] * du[d];
      duNorm = sqrt(duNorm);
      if (duNorm > domainFraction) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) du[d] *= domainFraction / duNorm;
      }
      double kappa = 0.0;
#pragma unroll
      for (int a = 0; a < DIM; ++a) {
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < DIM; ++b) s += HV[a][b] * du[b];
        kappa += du[a] * s;
      }
      const bool isPD = kappa > 0.0;
      double uold[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d) uold[d] = u[d];
      const double oldEnergy = energy(u);
#pragma unroll
      for (int d = 0; d < DIM; ++d) u[d] += du[d];
      if (clamp_all_boundaries<DIM>(u)) break;
      for (int j = 0; j < 20; ++j) {
        const double e = energy(u);
        if (e > oldEnergy && duNorm > 1e-8) {
          const double sf = ldexp(1.0, j + 1);
#pragma unroll
          for (int d = 0; d < DIM; ++d) u[d] = isPD ? uold[d] + du[d] / sf : uold[d] - du[d] / sf;
        } else
          break;
        if (j == 19) fl = 2;
      }
      if (fl) break;
      gradHess(u);
      double rn = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) rn += R[d] * R[d];
      if (sqrt(rn) < tol) break;
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d) u[d] = fl ? nan("") : fmin(fmax(u[d], 0.0), 1.0);
  }
#pragma unroll
  for (int d = 0; d < DIM; ++d) xiOut[pt * DIM + d] = u[d];
  if (flagOut) flagOut[pt] = fl;
}

template <int DIM, int DW>
void launch(const PatchDev& P, long long npts, const int* elem, const double* xg, double* xi, int* flag,
            cudaStream_t stream) {
  local_kernel<DIM, DW><<<(unsigned)((npts + 63) / 64), 64, 0, stream>>>(P, npts, elem, xg, xi, flag);
  count_launch();
}

}  // namespace

igab_status geometry_local_device(const PatchDev& P, long long npts, const int* elem, const double* xg, double* xi,
                                  int* flag, cudaStream_t stream) {
  if (npts == 0) return IGAB_OK;
  if (P.dim == 1 && P.dw == 1) launch<1, 1>(P, npts, elem, xg, xi, flag, stream);
  else if (P.dim == 1 && P.dw == 2) launch<1, 2>(P, npts, elem, xg, xi, flag, stream);
  else if (P.dim == 1 && P.dw == 3) launch<1, 3>(P, npts, elem, xg, xi, flag, stream);
  else if (P.dim == 2 && P.dw == 2) launch<2, 2>(P, npts, elem, xg, xi, flag, stream);
  else if (P.dim == 2 && P.dw == 3) launch<2, 3>(P, npts, elem, xg, xi, flag, stream);
  else if (P.dim == 3 && P.dw == 3) launch<3, 3>(P, npts, elem, xg, xi, flag, stream);
  else {
    set_error("geometry_local: unsupported (dim, dimworld)");
    return IGAB_UNSUPPORTED;
  }
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

}  // namespace igab

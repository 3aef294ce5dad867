// surface.cu -- differential geometry of surfaces from the J^T / H arrays the evaluation kernels write.
// Replaces, batched over points,
//   NURBSGeometry::normal                 dune/iga/nurbsgeometry.hh:223-228   N = J0 x J1
//   NURBSGeometry::unitNormal             dune/iga/nurbsgeometry.hh:216-221   n = N / |N|
//   NURBSGeometry::metric                 dune/iga/nurbsgeometry.hh:238-243   g = J J^T   (MatrixHelper::AAT)
//   NURBSGeometry::secondFundamentalForm  dune/iga/nurbsgeometry.hh:245-255   b_ab = H_ab . n, H rows [uu, vv, uv]
//   NURBSGeometry::gaussianCurvature      dune/iga/nurbsgeometry.hh:230-236   K = det b / det g
// One thread per point; the pass is a pure stream over (dim + dim(dim+1)/2) * dimworld doubles in and up to
// 2*dimworld + 2*dim*dim + 1 doubles out per point.
#include "igab_internal.cuh"

namespace igab {

namespace {

template <int DIM, int DW>
__global__ void __launch_bounds__(256) surface_forms_kernel(long long npts, const double* __restrict__ Jt,
                                                            const double* __restrict__ H, double* __restrict__ normal,
                                                            double* __restrict__ unitNormal, double* __restrict__ metric,
                                                            double* __restrict__ secondFF, double* __restrict__ gaussK) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npts) return;
  double J[DIM][DW];
#pragma unroll
  for (int d = 0; d < DIM; ++d)
#pragma unroll
    for (int c = 0; c < DW; ++c) J[d][c] = Jt[(t * DIM + d) * DW + c];
  double g[DIM][DIM];
#pragma unroll
  for (int a = 0; a < DIM; ++a)
#pragma unroll
    for (int b = a; b < DIM; ++b) {
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < DW; ++c) s += J[a][c] * J[b][c];
      g[a][b] = g[b][a] = s;
    }
  if (metric) {
#pragma unroll
    for (int a = 0; a < DIM; ++a)
#pragma unroll
      for (int b = 0; b < DIM; ++b) metric[(t * DIM + a) * DIM + b] = g[a][b];
  }
  if constexpr (DIM == 2 && DW == 3) {
    const double N[3] = {J[0][1] * J[1][2] - J[0][2] * J[1][1], J[0][2] * J[1][0] - J[0][0] * J[1][2],
                         J[0][0] * J[1][1] - J[0][1] * J[1][0]};
    const double len = sqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
    const double n[3] = {N[0] / len, N[1] / len, N[2] / len};
    if (normal) {
#pragma unroll
      for (int c = 0; c < 3; ++c) normal[t * 3 + c] = N[c];
    }
    if (unitNormal) {
#pragma unroll
      for (int c = 0; c < 3; ++c) unitNormal[t * 3 + c] = n[c];
    }
    if (H && (secondFF || gaussK)) {
      double b[3];  // uu, vv, uv
#pragma unroll
      for (int r = 0; r < 3; ++r) b[r] = H[(t * 3 + r) * 3] * n[0] + H[(t * 3 + r) * 3 + 1] * n[1] + H[(t * 3 + r) * 3 + 2] * n[2];
      if (secondFF) {
        secondFF[t * 4 + 0] = b[0];
        secondFF[t * 4 + 1] = b[2];
        secondFF[t * 4 + 2] = b[2];
        secondFF[t * 4 + 3] = b[1];
      }
      if (gaussK) gaussK[t] = (b[0] * b[1] - b[2] * b[2]) / (g[0][0] * g[1][1] - g[0][1] * g[1][0]);
    }
  }
}

template <int DIM, int DW>
void launch(long long npts, const double* Jt, const double* H, double* normal, double* unitNormal, double* metric,
            double* secondFF, double* gaussK, cudaStream_t stream) {
  const unsigned grid = (unsigned)((npts + 255) / 256);
  surface_forms_kernel<DIM, DW><<<grid, 256, 0, stream>>>(npts, Jt, H, normal, unitNormal, metric, secondFF, gaussK);
  count_launch();
}

}  // namespace

igab_status surface_forms_device(int dim, int dw, long long npts, const double* Jt, const double* H, double* normal,
                                 double* unitNormal, double* metric, double* secondFF, double* gaussK,
                                 cudaStream_t stream) {
  if (npts == 0) return IGAB_OK;
  const bool surf = normal || unitNormal || secondFF || gaussK;
  if (surf && !(dim == 2 && dw == 3)) {
    set_error("surface_forms: normal / second fundamental form / curvature need a surface in R^3 (dim 2, dimworld 3)");
    return IGAB_UNSUPPORTED;
  }
  if ((secondFF || gaussK) && !H) {
    set_error("surface_forms: second fundamental form and curvature need H");
    return IGAB_INVALID_ARGUMENT;
  }
#define IGAB_SURF_CASE(D, W) \
  if (dim == D && dw == W) { launch<D, W>(npts, Jt, H, normal, unitNormal, metric, secondFF, gaussK, stream); } else
  IGAB_SURF_CASE(1, 1) IGAB_SURF_CASE(1, 2) IGAB_SURF_CASE(1, 3) IGAB_SURF_CASE(2, 2) IGAB_SURF_CASE(2, 3)
  IGAB_SURF_CASE(3, 3) {
    set_error("surface_forms: unsupported (dim, dimworld)");
    return IGAB_UNSUPPORTED;
  }
#undef IGAB_SURF_CASE
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

}  // namespace igab

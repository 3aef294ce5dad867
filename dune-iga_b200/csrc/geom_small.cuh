// geom_small.cuh -- closed forms for the small matrices of a point's geometry, shared by the per-point kernels.
//   sqrt_det_aat: NURBSGeometry::integrationElement  dune/iga/nurbsgeometry.hh:200-203 (MatrixHelper::sqrtDetAAT of
//                 dune-geometry: |det J| for square J, norm of the minors for 2x3, row norm for 1xn)
//   right_inv:    NURBSGeometry::jacobianInverseTransposed  dune/iga/nurbsgeometry.hh:205-210 (MatrixHelper::rightInvA)
#pragma once

namespace igab {

template <int DIM, int DW>
__device__ __forceinline__ double sqrt_det_aat(const double (*J)[DW]) {
  if constexpr (DIM == 1) {
    double s = 0;
    for (int c = 0; c < DW; ++c) s += J[0][c] * J[0][c];
    return sqrt(s);
  } else if constexpr (DIM == 2 && DW == 2) {
    return fabs(J[0][0] * J[1][1] - J[0][1] * J[1][0]);
  } else if constexpr (DIM == 2 && DW == 3) {
    const double v0 = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double v1 = J[0][0] * J[1][2] - J[1][0] * J[0][2];
    const double v2 = J[0][1] * J[1][2] - J[0][2] * J[1][1];
    return sqrt(v0 * v0 + v1 * v1 + v2 * v2);
  } else {
    return fabs(J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]));
  }
}

// JinvT = J^T (J J^T)^-1, [DW][DIM]
template <int DIM, int DW>
__device__ __forceinline__ void right_inv(const double (*J)[DW], double* out) {
  double G[DIM][DIM], Gi[DIM][DIM];
  for (int i = 0; i < DIM; ++i)
    for (int j = 0; j < DIM; ++j) {
      double s = 0;
      for (int c = 0; c < DW; ++c) s += J[i][c] * J[j][c];
      G[i][j] = s;
    }
  if constexpr (DIM == 1) {
    Gi[0][0] = 1.0 / G[0][0];
  } else if constexpr (DIM == 2) {
    const double det = G[0][0] * G[1][1] - G[0][1] * G[1][0];
    Gi[0][0] = G[1][1] / det; Gi[0][1] = -G[0][1] / det; Gi[1][0] = -G[1][0] / det; Gi[1][1] = G[0][0] / det;
  } else {
    const double c00 = G[1][1] * G[2][2] - G[1][2] * G[2][1];
    const double c01 = G[1][2] * G[2][0] - G[1][0] * G[2][2];
    const double c02 = G[1][0] * G[2][1] - G[1][1] * G[2][0];
    const double det = G[0][0] * c00 + G[0][1] * c01 + G[0][2] * c02;
    Gi[0][0] = c00 / det; Gi[0][1] = (G[0][2] * G[2][1] - G[0][1] * G[2][2]) / det; Gi[0][2] = (G[0][1] * G[1][2] - G[0][2] * G[1][1]) / det;
    Gi[1][0] = c01 / det; Gi[1][1] = (G[0][0] * G[2][2] - G[0][2] * G[2][0]) / det; Gi[1][2] = (G[0][2] * G[1][0] - G[0][0] * G[1][2]) / det;
    Gi[2][0] = c02 / det; Gi[2][1] = (G[0][1] * G[2][0] - G[0][0] * G[2][1]) / det; Gi[2][2] = (G[0][0] * G[1][1] - G[0][1] * G[1][0]) / det;
  }
  for (int c = 0; c < DW; ++c)
    for (int j = 0; j < DIM; ++j) {
      double s = 0;
      for (int k = 0; k < DIM; ++k) s += J[k][c] * Gi[k][j];
      out[c * DIM + j] = s;
    }
}

}  // namespace igab

// Univariate B-spline device routines shared by the kernels.
//   ders1d_any: The NURBS Book A2.3 as the reference implements it, BsplineBasis1D::basisFunctionDerivatives,
//               dune/iga/bsplinealgorithms.hh:148-222 (ndu triangle :161-173, a-rows :178-213, factors :216-220).
#pragma once
#include "igab_internal.cuh"

namespace igab {

constexpr int kPM = kMaxOrder1D;  // 9: row length of every univariate work array

__device__ __forceinline__ bool feq8(double a, double b) {
  // Dune::FloatCmp::eq default (relativeWeak, 8 eps) -- bsplinealgorithms.hh:104-108
  const double eps = 8.0 * 2.220446049250313e-16;
  return fabs(a - b) <= eps * fmax(fabs(a), fabs(b));
}

// A2.2 with the reference's clamp and end-point early exits
__device__ inline void basis1d(double u, const double* __restrict__ U, int n, int p, int sp, double* N) {
  for (int i = 0; i <= p; ++i) N[i] = 0.0;
  u = fmin(fmax(u, U[0]), U[n - 1]);
  if (feq8(u, U[n - 1])) { N[p] = 1.0; return; }
  if (feq8(u, U[0])) { N[0] = 1.0; return; }
  N[0] = 1.0;
  for (int j = 0; j < p; ++j) {
    double saved = 0.0;
    for (int r = 0; r <= j; ++r) {
      const double rD = U[sp + 1 + r] - u;
      const double lD = u - U[sp - (j - r)];
      const double tmp = N[r] / (rD + lD);
      N[r] = saved + rD * tmp;
      saved = lD * tmp;
    }
    N[j + 1] = saved;
  }
}

// values + derivatives 0..K of the p+1 non-zero B-splines of span sp at u.  dN[k*kPM + j].  K may exceed p (rows = 0).
__device__ inline void ders1d_any(double u, const double* __restrict__ U, int p, int sp, int K, double* dN) {
  double left[kPM], right[kPM], ndu[kPM][kPM], a[2][kPM];
  ndu[0][0] = 1.0;
  for (int j = 1; j <= p; ++j) {
    left[j] = u - U[sp + 1 - j];
    right[j] = U[sp + j] - u;
    double saved = 0.0;
    for (int r = 0; r < j; ++r) {
      ndu[j][r] = right[r + 1] + left[j - r];
      const double temp = ndu[r][j - 1] / ndu[j][r];
      ndu[r][j] = saved + right[r + 1] * temp;
      saved = left[j - r] * temp;
    }
    ndu[j][j] = saved;
  }
  for (int j = 0; j <= p; ++j) dN[j] = ndu[j][p];
  for (int r = 0; r <= p; ++r) {
    int s1 = 0, s2 = 1;
    a[0][0] = 1.0;
    for (int k = 1; k <= K; ++k) {
      double d = 0.0;
      const int rk = r - k, pk = p - k;
      if (k <= p) {
        if (r >= k) {
          a[s2][0] = a[s1][0] / ndu[pk + 1][rk];
          d = a[s2][0] * ndu[rk][pk];
        }
        const int j1 = (rk >= -1) ? 1 : -rk;
        const int j2 = (r - 1 <= pk) ? k - 1 : p - r;
        for (int j = j1; j <= j2; ++j) {
          a[s2][j] = (a[s1][j] - a[s1][j - 1]) / ndu[pk + 1][rk + j];
          d += a[s2][j] * ndu[rk + j][pk];
        }
        if (r <= pk) {
          a[s2][k] = -a[s1][k - 1] / ndu[pk + 1][r];
          d += a[s2][k] * ndu[r][pk];
        }
      }
      dN[k * kPM + r] = d;
      const int t = s1; s1 = s2; s2 = t;
    }
  }
  int fac = p;
  for (int k = 1; k <= K; ++k) {
    for (int j = 0; j <= p; ++j) dN[k * kPM + j] *= fac;
    fac *= (p - k);
  }
}

// A2.3 (bsplinealgorithms.hh:148-222) with compile-time degree and order: fully unrolled, everything in registers.
// out[k][i] = k-th derivative w.r.t. the ELEMENT-LOCAL coordinate (scaled by span length^k).
template <int P, int K>
__device__ __forceinline__ void ders1d_reg(double u, const double* __restrict__ U, int sp, double scale,
                                           double (&out)[K + 1][P + 1]) {
  double left[P + 1], right[P + 1], ndu[P + 1][P + 1];
  ndu[0][0] = 1.0;
#pragma unroll
  for (int j = 1; j <= P; ++j) {
    left[j] = u - U[sp + 1 - j];
    right[j] = U[sp + j] - u;
    double saved = 0.0;
#pragma unroll
    for (int r = 0; r < j; ++r) {
      ndu[j][r] = right[r + 1] + left[j - r];
      const double temp = ndu[r][j - 1] / ndu[j][r];
      ndu[r][j] = saved + right[r + 1] * temp;
      saved = left[j - r] * temp;
    }
    ndu[j][j] = saved;
  }
#pragma unroll
  for (int j = 0; j <= P; ++j) out[0][j] = ndu[j][P];
  if constexpr (K >= 1) {
#pragma unroll
    for (int r = 0; r <= P; ++r) {
      double a[2][P + 1];
      a[0][0] = 1.0;
#pragma unroll
      for (int k = 1; k <= K; ++k) {
        const int s1 = (k - 1) & 1, s2 = k & 1;
        double d = 0.0;
        const int rk = r - k, pk = P - k;
        if (k <= P) {
          if (r >= k) {
            a[s2][0] = a[s1][0] / ndu[pk + 1][rk];
            d = a[s2][0] * ndu[rk][pk];
          }
          const int j1 = (rk >= -1) ? 1 : -rk;
          const int j2 = (r - 1 <= pk) ? k - 1 : P - r;
#pragma unroll
          for (int j = 1; j <= K; ++j)
            if (j >= j1 && j <= j2) {
              a[s2][j] = (a[s1][j] - a[s1][j - 1]) / ndu[pk + 1][rk + j];
              d += a[s2][j] * ndu[rk + j][pk];
            }
          if (r <= pk) {
            a[s2][k] = -a[s1][k - 1] / ndu[pk + 1][r];
            d += a[s2][k] * ndu[r][pk];
          }
        }
        out[k][r] = d;
      }
    }
    double fac = P, sc = scale;
#pragma unroll
    for (int k = 1; k <= K; ++k) {
#pragma unroll
      for (int j = 0; j <= P; ++j) out[k][j] *= fac * sc;
      fac *= (P - k);
      sc *= scale;
    }
  }
}

// The same A2.3 with the divisions replaced by multiplications: every denominator ndu[j][r] (r < j) is the knot
// difference U[sp+r+1] - U[sp+1-j+r] of the span, independent of u, so its reciprocal rd[j*(j-1)/2 + r] is tabulated per
// element of the direction at patch creation (PatchDev::rden).  Differs from the dividing form by <= 1 ulp per quotient.
template <int P, int K>
__device__ __forceinline__ void ders1d_reg_rd(double u, const double* __restrict__ U, int sp, double scale,
                                              const double* __restrict__ rdp, double (&out)[K + 1][P + 1]) {
  double left[P + 1], right[P + 1], nd[P + 1][P + 1];  // nd[r][j], r <= j: the upper triangle (basis values)
  constexpr int TRI = P * (P + 1) / 2;
  double rd[TRI > 0 ? TRI : 1];
#pragma unroll
  for (int t = 0; t < TRI; ++t) rd[t] = __ldg(rdp + t);
  nd[0][0] = 1.0;
#pragma unroll
  for (int j = 1; j <= P; ++j) {
    left[j] = u - U[sp + 1 - j];
    right[j] = U[sp + j] - u;
    double saved = 0.0;
#pragma unroll
    for (int r = 0; r < j; ++r) {
      const double temp = nd[r][j - 1] * rd[j * (j - 1) / 2 + r];
      nd[r][j] = saved + right[r + 1] * temp;
      saved = left[j - r] * temp;
    }
    nd[j][j] = saved;
  }
#pragma unroll
  for (int j = 0; j <= P; ++j) out[0][j] = nd[j][P];
  if constexpr (K >= 1) {
#pragma unroll
    for (int r = 0; r <= P; ++r) {
      double a[2][P + 1];
      a[0][0] = 1.0;
#pragma unroll
      for (int k = 1; k <= K; ++k) {
        const int s1 = (k - 1) & 1, s2 = k & 1;
        double d = 0.0;
        const int rk = r - k, pk = P - k;
        if (k <= P) {
          // 1 / ndu[pk+1][c] = rd[(pk+1) pk / 2 + c]
          if (r >= k) {
            a[s2][0] = a[s1][0] * rd[(pk + 1) * pk / 2 + rk];
            d = a[s2][0] * nd[rk][pk];
          }
          const int j1 = (rk >= -1) ? 1 : -rk;
          const int j2 = (r - 1 <= pk) ? k - 1 : P - r;
#pragma unroll
          for (int j = 1; j <= K; ++j)
            if (j >= j1 && j <= j2) {
              a[s2][j] = (a[s1][j] - a[s1][j - 1]) * rd[(pk + 1) * pk / 2 + rk + j];
              d += a[s2][j] * nd[rk + j][pk];
            }
          if (r <= pk) {
            a[s2][k] = -a[s1][k - 1] * rd[(pk + 1) * pk / 2 + r];
            d += a[s2][k] * nd[r][pk];
          }
        }
        out[k][r] = d;
      }
    }
    double fac = P, sc = scale;
#pragma unroll
    for (int k = 1; k <= K; ++k) {
#pragma unroll
      for (int j = 0; j <= P; ++j) out[k][j] *= fac * sc;
      fac *= (P - k);
      sc *= scale;
    }
  }
}

}  // namespace igab

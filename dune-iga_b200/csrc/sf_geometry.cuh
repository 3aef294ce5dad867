// sf_geometry.cuh -- the sum-factorised geometry front end shared by the 3-D tensor-rule kernels
// (eval_tensor_sf.cu: one warp per element, eval_tensor_blk.cu / assemble_fused.cu: one block per element).
//
// For one element it (1) copies the element's slice of the K1 univariate tables and its (p+1)^3 control points
// (netOfSpan: start = span - degree, dune/iga/nurbsalgorithms.hh:80-89) into shared memory, the control points in
// homogeneous form (w x, w y, w z, w), and (2) contracts that net with the three univariate tables one direction at a
// time:   G[alpha][c][pt] = sum_i dN^(alpha)_i(pt) H_i[c],   alpha in {0, e0, e1, e2}, c in {wx, wy, wz, w},
// from which x = C/W and J_d = (C_d - W_d x)/W follow (quotient rule of nurbsalgorithms.hh:226-248 applied to the
// geometry).  O((p+1)^4) per element instead of O((p+1)^3 Q^3).
//
// NT = number of cooperating threads: 32 (a warp, synchronised with __syncwarp) or a whole block (__syncthreads).
// tid = index of the calling thread inside that group.
#pragma once
#include "igab_internal.cuh"

namespace igab {

template <int P, int Q>
struct SfLayout {
  static constexpr int NB1 = P + 1, NB = NB1 * NB1 * NB1, NQ = Q * Q * Q, Q2 = Q * Q;
  static constexpr int TAB = Q * NB1 * 2;     // doubles per direction: [q][i][a], a = 0 value, 1 d/dxi
  static constexpr int HS = 4 * NB1 + 1;      // stride of H over i12 = i1 + NB1 i2 (odd: conflict-free stage-1 reads)
  static constexpr int S1C = NB1 * NB1 + 1;   // stride of S1 over c
  static constexpr int S2S = NB1 | 1;         // stride of S2 over q01 (odd)
  static constexpr int ev(int x) { return (x + 1) & ~1; }  // regions stay 16-byte aligned
  static constexpr int SZ_T = 3 * TAB;
  static constexpr int SZ_H = ev(NB1 * NB1 * HS);
  static constexpr int SZ_S1 = 2 * Q * 4 * S1C;
  static constexpr int SZ_G = 16 * NQ;        // [alpha(4)][c(4)][pt]
  static constexpr int SZ_A = ev((SZ_H + SZ_S1) > SZ_G ? (SZ_H + SZ_S1) : SZ_G);  // H + S1, later G
  static constexpr int SZ_S2 = ev(3 * 4 * Q2 * S2S);
};

template <int NT>
__device__ __forceinline__ void sf_sync() {
  if constexpr (NT == 32)
    __syncwarp();
  else
    __syncthreads();
}

__device__ __forceinline__ double2 sf_lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }

// (1) tables and control points of element (e0, e1, e2) -> sT, sH (in region A), sW.  No trailing sync.
template <int P, int Q, int NT>
__device__ __forceinline__ void sf_load(const PatchDev& Pd, int tid, int e0, int e1, int e2,
                                        const double* __restrict__ tab0, const double* __restrict__ tab1,
                                        const double* __restrict__ tab2, double* sT, double* sW, double* rA) {
  using L = SfLayout<P, Q>;
  constexpr int NB1 = L::NB1, NB = L::NB;
  const double* t0 = tab0 + (size_t)e0 * L::TAB;
  const double* t1 = tab1 + (size_t)e1 * L::TAB;
  const double* t2 = tab2 + (size_t)e2 * L::TAB;
  for (int j = tid; j < L::TAB; j += NT) {
    sT[j] = __ldg(t0 + j);
    sT[L::TAB + j] = __ldg(t1 + j);
    sT[2 * L::TAB + j] = __ldg(t2 + j);
  }
  const int b0 = Pd.spanTab[0][e0] - P, b1 = Pd.spanTab[1][e1] - P, b2 = Pd.spanTab[2][e2] - P;
  const long long n0 = Pd.ncp[0], n1 = Pd.ncp[1];
  const double2* cp2 = reinterpret_cast<const double2*>(Pd.cp);
  double* sH = rA;
#pragma unroll
  for (int k = 0; k < (NB + NT - 1) / NT; ++k) {
    const int i = tid + NT * k;
    if (i < NB) {
      const int i0 = i % NB1, i12 = i / NB1, i1 = i12 % NB1, i2 = i12 / NB1;
      const long long g = (b0 + i0) + n0 * ((b1 + i1) + n1 * (long long)(b2 + i2));
      const double2 xy = __ldg(cp2 + 2 * g), zw = __ldg(cp2 + 2 * g + 1);
      double* h = sH + i12 * L::HS + 4 * i0;
      h[0] = xy.x * zw.y;
      h[1] = xy.y * zw.y;
      h[2] = zw.x * zw.y;
      h[3] = zw.y;
      sW[i] = zw.y;
    }
  }
}

// (2) the three 1-D contractions.  In: sT, H (region A).  Scratch: S1 (region A), S2.  Out: G (region A, overwrites
// H / S1).  Starts and ends with a group sync.
template <int P, int Q, int NT>
__device__ __forceinline__ void sf_contract(int tid, const double* sT, double* rA, double* sS2) {
  using L = SfLayout<P, Q>;
  constexpr int NB1 = L::NB1, NQ = L::NQ, Q2 = L::Q2;
  double* const sH = rA;
  double* const sS1 = rA + L::SZ_H;
  double* const sG = rA;
  sf_sync<NT>();
  // stage 1: S1[a0][q0][c][i12] = sum_i0 T0[q0][i0][a0] H[i12][i0][c]
  for (int col = tid; col < 4 * NB1 * NB1; col += NT) {
    const int i12 = col % (NB1 * NB1), c = col / (NB1 * NB1);
    double h[NB1];
#pragma unroll
    for (int i0 = 0; i0 < NB1; ++i0) h[i0] = sH[i12 * L::HS + 4 * i0 + c];
#pragma unroll
    for (int q0 = 0; q0 < Q; ++q0) {
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int i0 = 0; i0 < NB1; ++i0) {
        const double2 t = sf_lds2(sT + (q0 * NB1 + i0) * 2);
        a0 = fma(t.x, h[i0], a0);
        a1 = fma(t.y, h[i0], a1);
      }
      sS1[((0 * Q + q0) * 4 + c) * L::S1C + i12] = a0;
      sS1[((1 * Q + q0) * 4 + c) * L::S1C + i12] = a1;
    }
  }
  sf_sync<NT>();
  // stage 2: S2[a01][c][q0 + Q q1][i2] = sum_i1 T1[q1][i1][a1] S1[a0][q0][c][i1 + NB1 i2];  a01: 0=(0,0) 1=(1,0) 2=(0,1)
  for (int col = tid; col < 2 * Q * 4 * NB1; col += NT) {
    int r = col;
    const int i2 = r % NB1; r /= NB1;
    const int c = r % 4; r /= 4;
    const int q0 = r % Q;
    const int a0 = r / Q;
    double s[NB1];
#pragma unroll
    for (int i1 = 0; i1 < NB1; ++i1) s[i1] = sS1[((a0 * Q + q0) * 4 + c) * L::S1C + i1 + NB1 * i2];
#pragma unroll
    for (int q1 = 0; q1 < Q; ++q1) {
      double v0 = 0.0, v1 = 0.0;
#pragma unroll
      for (int i1 = 0; i1 < NB1; ++i1) {
        const double2 t = sf_lds2(sT + L::TAB + (q1 * NB1 + i1) * 2);
        v0 = fma(t.x, s[i1], v0);
        v1 = fma(t.y, s[i1], v1);
      }
      const int q01 = q0 + Q * q1;
      if (a0 == 0) {
        sS2[((0 * 4 + c) * Q2 + q01) * L::S2S + i2] = v0;
        sS2[((2 * 4 + c) * Q2 + q01) * L::S2S + i2] = v1;
      } else {
        sS2[((1 * 4 + c) * Q2 + q01) * L::S2S + i2] = v0;
      }
    }
  }
  sf_sync<NT>();
  // stage 3: G[alpha][c][q01 + Q^2 q2] = sum_i2 T2[q2][i2][a2] S2[a01][c][q01][i2];  alpha: 0 value, 1..3 d/dxi_0..2
  for (int col = tid; col < 3 * 4 * Q2; col += NT) {
    int r = col;
    const int q01 = r % Q2; r /= Q2;
    const int c = r % 4;
    const int a01 = r / 4;
    double s[NB1];
#pragma unroll
    for (int i2 = 0; i2 < NB1; ++i2) s[i2] = sS2[((a01 * 4 + c) * Q2 + q01) * L::S2S + i2];
#pragma unroll
    for (int q2 = 0; q2 < Q; ++q2) {
      double v0 = 0.0, v1 = 0.0;
#pragma unroll
      for (int i2 = 0; i2 < NB1; ++i2) {
        const double2 t = sf_lds2(sT + 2 * L::TAB + (q2 * NB1 + i2) * 2);
        v0 = fma(t.x, s[i2], v0);
        v1 = fma(t.y, s[i2], v1);
      }
      const int pt = q01 + Q2 * q2;
      if (a01 == 0) {
        sG[(0 * 4 + c) * NQ + pt] = v0;
        sG[(3 * 4 + c) * NQ + pt] = v1;
      } else {
        sG[(a01 * 4 + c) * NQ + pt] = v0;  // a01 = 1 -> alpha 1 ; a01 = 2 -> alpha 2
      }
    }
  }
  sf_sync<NT>();
}

// x, J^T (rows d = d/dxi_d), from G at point pt
template <int NQ>
__device__ __forceinline__ void sf_point(const double* sG, int pt, double& invW, double (&Wd)[3], double (&x)[3],
                                         double (&J)[3][3]) {
  const double W = sG[3 * NQ + pt];
  invW = 1.0 / W;
#pragma unroll
  for (int c = 0; c < 3; ++c) x[c] = sG[c * NQ + pt] * invW;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    Wd[d] = sG[((d + 1) * 4 + 3) * NQ + pt];
#pragma unroll
    for (int c = 0; c < 3; ++c) J[d][c] = (sG[((d + 1) * 4 + c) * NQ + pt] - Wd[d] * x[c]) * invW;
  }
}

// det J and J^-T[c][d] = cof(J)[d][c] / det  (square case of rightInvA, dune/iga/nurbsgeometry.hh:205-210)
__device__ __forceinline__ double sf_det_cof(const double (&J)[3][3], double (&cof)[3][3]) {
  cof[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
  cof[0][1] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
  cof[0][2] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  cof[1][0] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
  cof[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
  cof[1][2] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
  cof[2][0] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
  cof[2][1] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
  cof[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
  return J[0][0] * cof[0][0] + J[0][1] * cof[0][1] + J[0][2] * cof[0][2];
}

}  // namespace igab

// norms.cu -- patch-level reductions of a discrete field u_h = sum_i R_i u_i (BASELINE north_star: "patch-level
// reductions (volume, L2/energy norms)"), fused: nothing is written to HBM but 2 x 1024 partial sums.
//   result[0] = sum_e sum_q w detJ |u_h(x_q)|^2                      (squared L2 norm            = u^T M u)
//   result[1] = sum_e sum_q w detJ grad u_h . grad u_h                (squared H1 seminorm/energy = u^T K u, Poisson)
// with R_i, dR_i/dxi as NurbsLocalBasis::evaluateFunction / evaluateJacobian (dune/iga/nurbsbasis.hh:77-96), detJ =
// integrationElement = sqrt(det J J^T) (nurbsgeometry.hh:200-203) and grad = J^T (J J^T)^-1 d/dxi
// (jacobianInverseTransposed, :205-210), i.e. the quantities the loop of dune/python/test/poisson.py:47-79 forms per
// point.  u is indexed like the global basis: g*ncomp + c with g the tensor-product DOF index (nurbsbasis.hh:572-584 =
// control-point index for an untrimmed patch; flat-interleaved components as PowerPreBasis).
//
// One thread per quadrature point (grid-stride over a fixed contiguous slice per block): univariate rows by A2.3, the
// field rides along with the geometry as extra homogeneous coordinates, A^(a)[k] = sum_i N_i^(a) w_i {P_i, 1, u_i},
// then the quotient rule.  Deterministic: sequential per thread, fixed tree per block, fixed-order final sum.
#include <algorithm>
#include <cstdlib>

#include "bspline_device.cuh"
#include "igab_internal.cuh"

namespace igab {

namespace {

constexpr int kMaxComp = 3;

// where the points come from besides the tensor rule, and the numbering of u on trimmed patches
struct NormSrc {
  const uint8_t* trim;    // element flags (tensor mode: only IGAB_TRIM_FULL elements contribute) or nullptr
  const int32_t* map;     // original scalar DOF -> compacted (prepareForTrim) or nullptr
  const int32_t* lelem;   // list mode: element of point k
  const double* lxi;      //            element-local coordinates [npts][dim]
  const double* lw;       //            weights
};

template <int DIM, int DW, bool LIST>
__global__ void __launch_bounds__(128) norms_partial_kernel(PatchDev P, NormSrc S, long long elem_begin, long long npts, long long nq,
                                                            int nq0, int nq1_, int nq2, const double* __restrict__ x0,
                                                            const double* __restrict__ x1, const double* __restrict__ x2,
                                                            const double* __restrict__ w0, const double* __restrict__ w1,
                                                            const double* __restrict__ w2, const double* __restrict__ u,
                                                            int ncomp, double* __restrict__ partial, int npartial) {
  __shared__ double sh[2][128];
  const long long per = (npts + gridDim.x - 1) / gridDim.x;
  const long long b0 = (long long)blockIdx.x * per, b1 = min(npts, b0 + per);
  constexpr int NK = DW + 1 + kMaxComp;
  double accL2 = 0.0, accEn = 0.0;
  for (long long pt = b0 + threadIdx.x; pt < b1; pt += 128) {
    const long long e = LIST ? (long long)S.lelem[pt] : elem_begin + pt / nq;
    if (!LIST && S.trim && S.trim[e] != IGAB_TRIM_FULL) continue;  // trimmed elements bring their own points
    long long q = LIST ? 0 : pt % nq;
    double xi[DIM], wq;
    if constexpr (LIST) {
#pragma unroll
      for (int d = 0; d < DIM; ++d) xi[d] = S.lxi[pt * DIM + d];
      wq = S.lw[pt];
    } else {
      const int a = (int)(q % nq0);
      xi[0] = x0[a];
      wq = w0[a];
      q /= nq0;
      if constexpr (DIM > 1) {
        const int b = (int)(q % nq1_);
        xi[1] = x1[b];
        wq *= w1[b];
        q /= nq1_;
      }
      if constexpr (DIM > 2) {
        xi[2] = x2[q % nq2];
        wq *= w2[q % nq2];
      }
    }
    int sp[DIM];
    double scale[DIM], ders[DIM][2 * kPM];
    long long cpBase = 0, cpStride[DIM];
    {
      long long h = e, stride = 1;
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        const int ed = (int)(h % P.nel[d]);
        h /= P.nel[d];
        sp[d] = P.spanTab[d][ed];
        const double lo = P.knots[d][sp[d]];
        scale[d] = P.knots[d][sp[d] + 1] - lo;
        ders1d_any(xi[d] * scale[d] + lo, P.knots[d], P.degree[d], sp[d], 1, ders[d]);
        cpStride[d] = stride;
        cpBase += (long long)(sp[d] - P.degree[d]) * stride;
        stride *= P.ncp[d];
      }
    }
    double A[1 + DIM][NK];
#pragma unroll
    for (int a = 0; a <= DIM; ++a)
#pragma unroll
      for (int k = 0; k < NK; ++k) A[a][k] = 0.0;
    int l[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) l[d] = 0;
    for (int i = 0; i < P.nb; ++i) {
      long long g = cpBase;
#pragma unroll
      for (int d = 0; d < DIM; ++d) g += l[d] * cpStride[d];
      double V[NK];
#pragma unroll
      for (int k = 0; k < DW; ++k) V[k] = P.cp[g * (DW + 1) + k];
      V[DW] = 1.0;
      const double w = P.cp[g * (DW + 1) + DW];
#pragma unroll
      const long long gu = S.map ? (long long)S.map[g] : g;  // compacted numbering on trimmed patches
#pragma unroll
      for (int c = 0; c < kMaxComp; ++c) V[DW + 1 + c] = (c < ncomp && gu >= 0) ? u[gu * ncomp + c] : 0.0;
#pragma unroll
      for (int a = 0; a <= DIM; ++a) {
        double f = w;
#pragma unroll
        for (int d = 0; d < DIM; ++d) f *= ders[d][((a == d + 1) ? kPM : 0) + l[d]];
#pragma unroll
        for (int k = 0; k < NK; ++k) A[a][k] += f * V[k];
      }
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        if (++l[d] <= P.degree[d]) break;
        l[d] = 0;
      }
    }
    // quotient rule on geometry and field; derivatives scaled to the element's [0,1]^dim
    const double iW = 1.0 / A[0][DW];
    double F[NK], dF[DIM][NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) F[k] = A[0][k] * iW;
#pragma unroll
    for (int d = 0; d < DIM; ++d)
#pragma unroll
      for (int k = 0; k < NK; ++k) dF[d][k] = (A[1 + d][k] - A[1 + d][DW] * F[k]) * iW * scale[d];
    double G[DIM][DIM];
#pragma unroll
    for (int a = 0; a < DIM; ++a)
#pragma unroll
      for (int b = 0; b < DIM; ++b) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < DW; ++k) s += dF[a][k] * dF[b][k];
        G[a][b] = s;
      }
    double Gi[DIM][DIM], detG;
    if constexpr (DIM == 1) {
      detG = G[0][0];
      Gi[0][0] = 1.0 / detG;
    } else if constexpr (DIM == 2) {
      detG = G[0][0] * G[1][1] - G[0][1] * G[1][0];
      Gi[0][0] = G[1][1] / detG; Gi[0][1] = -G[0][1] / detG; Gi[1][0] = -G[1][0] / detG; Gi[1][1] = G[0][0] / detG;
    } else {
      const double c00 = G[1][1] * G[2][2] - G[1][2] * G[2][1], c01 = G[1][2] * G[2][0] - G[1][0] * G[2][2],
                   c02 = G[1][0] * G[2][1] - G[1][1] * G[2][0];
      detG = G[0][0] * c00 + G[0][1] * c01 + G[0][2] * c02;
      Gi[0][0] = c00 / detG; Gi[0][1] = (G[0][2] * G[2][1] - G[0][1] * G[2][2]) / detG; Gi[0][2] = (G[0][1] * G[1][2] - G[0][2] * G[1][1]) / detG;
      Gi[1][0] = c01 / detG; Gi[1][1] = (G[0][0] * G[2][2] - G[0][2] * G[2][0]) / detG; Gi[1][2] = (G[0][2] * G[1][0] - G[0][0] * G[1][2]) / detG;
      Gi[2][0] = c02 / detG; Gi[2][1] = (G[0][1] * G[2][0] - G[0][0] * G[2][1]) / detG; Gi[2][2] = (G[0][0] * G[1][1] - G[0][1] * G[1][0]) / detG;
    }
    const double wd = wq * sqrt(fabs(detG));
    double l2 = 0.0, en = 0.0;
#pragma unroll
    for (int c = 0; c < kMaxComp; ++c) {
      const int k = DW + 1 + c;
      l2 += F[k] * F[k];
#pragma unroll
      for (int a = 0; a < DIM; ++a)
#pragma unroll
        for (int b = 0; b < DIM; ++b) en += dF[a][k] * Gi[a][b] * dF[b][k];
    }
    accL2 += wd * l2;
    accEn += wd * en;
  }
  sh[0][threadIdx.x] = accL2;
  sh[1][threadIdx.x] = accEn;
  __syncthreads();
  for (int s = 64; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + s];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = sh[0][0];
    partial[npartial + blockIdx.x] = sh[1][0];
  }
}

// result[k] = fixed-order tree sum of partial[k*npartial .. +n)
__global__ void __launch_bounds__(1024) norms_final_kernel(int n, int npartial, const double* __restrict__ partial,
                                                           double* __restrict__ result, int accumulate) {
  __shared__ double sh[1024];
  for (int k = 0; k < 2; ++k) {
    sh[threadIdx.x] = threadIdx.x < n ? partial[k * npartial + threadIdx.x] : 0.0;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
      if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) result[k] = accumulate ? result[k] + sh[0] : sh[0];
    __syncthreads();
  }
}

template <int DIM, int DW>
void launch(const PatchDev& P, const NormSrc& S, bool list, long long eb, long long npts, long long nq, const int* nq1,
            const double* const* xi, const double* const* w, const double* u, int ncomp, double* partial, int npartial,
            int nblk, cudaStream_t stream) {
  const int d1 = DIM > 1 ? 1 : 0, d2 = DIM > 2 ? 2 : 0;
  if (list)
    norms_partial_kernel<DIM, DW, true><<<nblk, 128, 0, stream>>>(P, S, eb, npts, 1, 1, 1, 1, nullptr, nullptr, nullptr,
                                                                  nullptr, nullptr, nullptr, u, ncomp, partial, npartial);
  else
    norms_partial_kernel<DIM, DW, false><<<nblk, 128, 0, stream>>>(P, S, eb, npts, nq, nq1[0], nq1[d1], nq1[d2], xi[0],
                                                                   xi[d1], xi[d2], w[0], w[d1], w[d2], u, ncomp, partial,
                                                                   npartial);
}

igab_status run(const PatchDev& P, const NormSrc& S, bool list, long long eb, long long npts, long long nq, const int* nq1,
                const double* const* xi, const double* const* w, const double* u, int ncomp, double* partial, int npartial,
                double* result_dev, int accumulate, cudaStream_t stream) {
  const int nblk = (int)std::min<long long>(npartial, (npts + 127) / 128);
  switch (P.dim * 10 + P.dw) {
    case 11: launch<1, 1>(P, S, list, eb, npts, nq, nq1, xi, w, u, ncomp, partial, npartial, nblk, stream); break;
    case 12: launch<1, 2>(P, S, list, eb, npts, nq, nq1, xi, w, u, ncomp, partial, npartial, nblk, stream); break;
    case 13: launch<1, 3>(P, S, list, eb, npts, nq, nq1, xi, w, u, ncomp, partial, npartial, nblk, stream); break;
    case 22: launch<2, 2>(P, S, list, eb, npts, nq, nq1, xi, w, u, ncomp, partial, npartial, nblk, stream); break;
    case 23: launch<2, 3>(P, S, list, eb, npts, nq, nq1, xi, w, u, ncomp, partial, npartial, nblk, stream); break;
    case 33: launch<3, 3>(P, S, list, eb, npts, nq, nq1, xi, w, u, ncomp, partial, npartial, nblk, stream); break;
    default:
      set_error("reduce_norms: unsupported (dim, dimworld)");
      return IGAB_UNSUPPORTED;
  }
  norms_final_kernel<<<1, 1024, 0, stream>>>(nblk, npartial, partial, result_dev, accumulate);
  count_launch(2);
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

// ---- K6c: the same reduction for 3-D tensor rules with equal degrees, SUM-FACTORISED --------------------------------------
// One block per element (fixed grid of `npartial` blocks, grid-stride): the field rides along with the geometry as extra
// homogeneous coordinates of the element's control net, H_i = w_i (P_i, 1, u_i), which is contracted with the three
// univariate K1 tables one direction at a time (as sf_geometry.cuh does for the geometry alone): O((p+1)^4) per element
// instead of O((p+1)^3 Q^3) for the thread-per-point kernel above (11x fewer FMAs at p=3 with one component).  Then the
// quotient rule per point, a fixed tree over the block and a sequential sum over the block's elements: deterministic.
template <int P, int Q, int NCF>
struct NCfg {
  static constexpr int NB1 = P + 1, NB = NB1 * NB1 * NB1, NQ = Q * Q * Q, Q2 = Q * Q, NC = 4 + NCF;
  static constexpr int TAB = Q * NB1 * 2;
  static constexpr int HS = NC * NB1 + 1, S1C = NB1 * NB1 + 1, S2S = NB1 | 1;
  static constexpr int ev(int x) { return (x + 1) & ~1; }
  static constexpr int SZ_T = 3 * TAB;
  static constexpr int SZ_H = ev(NB1 * NB1 * HS), SZ_S1 = 2 * Q * NC * S1C, SZ_G = 4 * NC * NQ;
  static constexpr int SZ_A = ev((SZ_H + SZ_S1) > SZ_G ? (SZ_H + SZ_S1) : SZ_G);
  static constexpr int SZ_S2 = ev(3 * NC * Q2 * S2S);
  static constexpr int SZ_TOTAL = SZ_T + SZ_A + SZ_S2 + 2 * 128;
};

template <int P, int Q, int NCF>
__global__ void __launch_bounds__(128) norms_sf3d_kernel(PatchDev Pd, NormSrc S, long long elem_begin, long long nelem,
                                                         const double* __restrict__ tab0, const double* __restrict__ tab1,
                                                         const double* __restrict__ tab2, const double* __restrict__ w0,
                                                         const double* __restrict__ w1, const double* __restrict__ w2,
                                                         const double* __restrict__ u, int ncomp,
                                                         double* __restrict__ partial, int npartial) {
  using C = NCfg<P, Q, NCF>;
  constexpr int NB1 = C::NB1, NB = C::NB, NQ = C::NQ, Q2 = C::Q2, NC = C::NC, TAB = C::TAB, NT = 128;
  extern __shared__ __align__(16) double nsm[];
  double* const sT = nsm;
  double* const rA = sT + C::SZ_T;
  double* const sH = rA;
  double* const sS1 = rA + C::SZ_H;
  double* const sG = rA;
  double* const sS2 = rA + C::SZ_A;
  double* const red = sS2 + C::SZ_S2;
  const int tid = threadIdx.x;
  double accL2 = 0.0, accEn = 0.0;  // thread 0: running sums over this block's elements
  for (long long el = blockIdx.x; el < nelem; el += gridDim.x) {
    const long long e = elem_begin + el;
    if (S.trim && S.trim[e] != IGAB_TRIM_FULL) continue;  // block-uniform
    int e0, e1, e2;
    {
      long long h = e;
      e0 = (int)(h % Pd.nel[0]); h /= Pd.nel[0];
      e1 = (int)(h % Pd.nel[1]); h /= Pd.nel[1];
      e2 = (int)h;
    }
    __syncthreads();  // previous element done with every buffer
    for (int j = tid; j < TAB; j += NT) {
      sT[j] = __ldg(tab0 + (size_t)e0 * TAB + j);
      sT[TAB + j] = __ldg(tab1 + (size_t)e1 * TAB + j);
      sT[2 * TAB + j] = __ldg(tab2 + (size_t)e2 * TAB + j);
    }
    {
      const int b0 = Pd.spanTab[0][e0] - P, b1 = Pd.spanTab[1][e1] - P, b2 = Pd.spanTab[2][e2] - P;
      const long long n0 = Pd.ncp[0], n1 = Pd.ncp[1];
      for (int i = tid; i < NB; i += NT) {
        const int i0 = i % NB1, i12 = i / NB1, i1 = i12 % NB1, i2 = i12 / NB1;
        const long long g = (b0 + i0) + n0 * ((b1 + i1) + n1 * (long long)(b2 + i2));
        const double* cp = Pd.cp + g * 4;
        const double w = cp[3];
        double* h = sH + i12 * C::HS + NC * i0;
        h[0] = cp[0] * w;
        h[1] = cp[1] * w;
        h[2] = cp[2] * w;
        h[3] = w;
        const long long gu = S.map ? (long long)S.map[g] : g;  // compacted numbering on trimmed patches
#pragma unroll
        for (int c = 0; c < NCF; ++c) h[4 + c] = (c < ncomp && gu >= 0) ? w * u[gu * ncomp + c] : 0.0;
      }
    }
    __syncthreads();
    // stage 1: S1[a0][q0][c][i12] = sum_i0 T0[q0][i0][a0] H[i12][i0][c]
    for (int col = tid; col < NC * NB1 * NB1; col += NT) {
      const int i12 = col % (NB1 * NB1), c = col / (NB1 * NB1);
      double h[NB1];
#pragma unroll
      for (int i0 = 0; i0 < NB1; ++i0) h[i0] = sH[i12 * C::HS + NC * i0 + c];
#pragma unroll
      for (int q0 = 0; q0 < Q; ++q0) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int i0 = 0; i0 < NB1; ++i0) {
          a0 = fma(sT[(q0 * NB1 + i0) * 2], h[i0], a0);
          a1 = fma(sT[(q0 * NB1 + i0) * 2 + 1], h[i0], a1);
        }
        sS1[((0 * Q + q0) * NC + c) * C::S1C + i12] = a0;
        sS1[((1 * Q + q0) * NC + c) * C::S1C + i12] = a1;
      }
    }
    __syncthreads();
    // stage 2: S2[a01][c][q0 + Q q1][i2] = sum_i1 T1[q1][i1][a1] S1[a0][q0][c][i1 + NB1 i2];  a01: 0=(0,0) 1=(1,0) 2=(0,1)
    for (int col = tid; col < 2 * Q * NC * NB1; col += NT) {
      int r = col;
      const int i2 = r % NB1; r /= NB1;
      const int c = r % NC; r /= NC;
      const int q0 = r % Q;
      const int a0 = r / Q;
      double sv[NB1];
#pragma unroll
      for (int i1 = 0; i1 < NB1; ++i1) sv[i1] = sS1[((a0 * Q + q0) * NC + c) * C::S1C + i1 + NB1 * i2];
#pragma unroll
      for (int q1 = 0; q1 < Q; ++q1) {
        double v0 = 0.0, v1 = 0.0;
#pragma unroll
        for (int i1 = 0; i1 < NB1; ++i1) {
          v0 = fma(sT[TAB + (q1 * NB1 + i1) * 2], sv[i1], v0);
          v1 = fma(sT[TAB + (q1 * NB1 + i1) * 2 + 1], sv[i1], v1);
        }
        const int q01 = q0 + Q * q1;
        if (a0 == 0) {
          sS2[((0 * NC + c) * Q2 + q01) * C::S2S + i2] = v0;
          sS2[((2 * NC + c) * Q2 + q01) * C::S2S + i2] = v1;
        } else {
          sS2[((1 * NC + c) * Q2 + q01) * C::S2S + i2] = v0;
        }
      }
    }
    __syncthreads();
    // stage 3: G[alpha][c][q01 + Q^2 q2] = sum_i2 T2[q2][i2][a2] S2[a01][c][q01][i2];  alpha: 0 value, 1..3 d/dxi_0..2
    for (int col = tid; col < 3 * NC * Q2; col += NT) {
      int r = col;
      const int q01 = r % Q2; r /= Q2;
      const int c = r % NC;
      const int a01 = r / NC;
      double sv[NB1];
#pragma unroll
      for (int i2 = 0; i2 < NB1; ++i2) sv[i2] = sS2[((a01 * NC + c) * Q2 + q01) * C::S2S + i2];
#pragma unroll
      for (int q2 = 0; q2 < Q; ++q2) {
        double v0 = 0.0, v1 = 0.0;
#pragma unroll
        for (int i2 = 0; i2 < NB1; ++i2) {
          v0 = fma(sT[2 * TAB + (q2 * NB1 + i2) * 2], sv[i2], v0);
          v1 = fma(sT[2 * TAB + (q2 * NB1 + i2) * 2 + 1], sv[i2], v1);
        }
        const int pt = q01 + Q2 * q2;
        if (a01 == 0) {
          sG[(0 * NC + c) * NQ + pt] = v0;
          sG[(3 * NC + c) * NQ + pt] = v1;
        } else {
          sG[(a01 * NC + c) * NQ + pt] = v0;
        }
      }
    }
    __syncthreads();
    // per point: quotient rule on geometry and field, |u|^2 and grad u . grad u = d_xi u^T (J J^T)^-1 d_xi u
    double l2 = 0.0, en = 0.0;
    for (int pt = tid; pt < NQ; pt += NT) {
      const double iW = 1.0 / sG[3 * NQ + pt];
      double F[NC], dF[3][NC];
#pragma unroll
      for (int k = 0; k < NC; ++k) F[k] = sG[k * NQ + pt] * iW;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double Wd = sG[((d + 1) * NC + 3) * NQ + pt];
#pragma unroll
        for (int k = 0; k < NC; ++k) dF[d][k] = (sG[((d + 1) * NC + k) * NQ + pt] - Wd * F[k]) * iW;
      }
      double G[3][3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) G[a][b] = dF[a][0] * dF[b][0] + dF[a][1] * dF[b][1] + dF[a][2] * dF[b][2];
      const double c00 = G[1][1] * G[2][2] - G[1][2] * G[2][1], c01 = G[1][2] * G[2][0] - G[1][0] * G[2][2],
                   c02 = G[1][0] * G[2][1] - G[1][1] * G[2][0];
      const double detG = G[0][0] * c00 + G[0][1] * c01 + G[0][2] * c02, id = 1.0 / detG;
      double Gi[3][3];
      Gi[0][0] = c00 * id; Gi[0][1] = (G[0][2] * G[2][1] - G[0][1] * G[2][2]) * id; Gi[0][2] = (G[0][1] * G[1][2] - G[0][2] * G[1][1]) * id;
      Gi[1][0] = c01 * id; Gi[1][1] = (G[0][0] * G[2][2] - G[0][2] * G[2][0]) * id; Gi[1][2] = (G[0][2] * G[1][0] - G[0][0] * G[1][2]) * id;
      Gi[2][0] = c02 * id; Gi[2][1] = (G[0][1] * G[2][0] - G[0][0] * G[2][1]) * id; Gi[2][2] = (G[0][0] * G[1][1] - G[0][1] * G[1][0]) * id;
      int q = pt;
      double wq = w0[q % Q];
      q /= Q;
      wq *= w1[q % Q];
      wq *= w2[q / Q];
      const double wd = wq * sqrt(fabs(detG));
      double a2 = 0.0, g2 = 0.0;
#pragma unroll
      for (int c = 0; c < NCF; ++c) {
        const int k = 4 + c;
        a2 += F[k] * F[k];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) g2 += dF[a][k] * Gi[a][b] * dF[b][k];
      }
      l2 += wd * a2;
      en += wd * g2;
    }
    red[tid] = l2;
    red[128 + tid] = en;
    __syncthreads();
    for (int s = 64; s > 0; s >>= 1) {
      if (tid < s) {
        red[tid] += red[tid + s];
        red[128 + tid] += red[128 + tid + s];
      }
      __syncthreads();
    }
    if (tid == 0) {
      accL2 += red[0];
      accEn += red[128];
    }
  }
  if (tid == 0) {
    partial[blockIdx.x] = accL2;
    partial[npartial + blockIdx.x] = accEn;
  }
}

template <int P, int Q, int NCF>
igab_status launch_sfn(const PatchDev& Pd, SfWorkspace& ws, const NormSrc& S, long long eb, long long nel,
                       const double* const* w, const double* u, int ncomp, double* partial, int npartial,
                       double* result_dev, cudaStream_t stream) {
  using C = NCfg<P, Q, NCF>;
  const size_t smem = (size_t)C::SZ_TOTAL * sizeof(double);
  KernelCfg kc;
  if (igab_status cst = kernel_config((const void*)norms_sf3d_kernel<P, Q, NCF>, 128, smem, &kc)) return cst;
  const int nblk = (int)std::min<long long>(npartial, nel);  // fixed by the element count alone: reproducible on any GPU
  norms_sf3d_kernel<P, Q, NCF><<<nblk, 128, smem, stream>>>(Pd, S, eb, nel, ws.tab[0], ws.tab[1], ws.tab[2], w[0], w[1], w[2],
                                                           u, ncomp, partial, npartial);
  norms_final_kernel<<<1, 1024, 0, stream>>>(nblk, npartial, partial, result_dev, 0);
  count_launch(2);
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

}  // namespace

// Sum-factorised tensor part for 3-D patches with equal degrees 2..4 and p+1 / p+2 points per direction; IGAB_UNSUPPORTED
// (no error message) otherwise -- the caller then takes launch_norms.  Writes result_dev[0..1] (no list points).
igab_status launch_norms_sf(const PatchDev& P, SfWorkspace& ws, const double* const* xi1d_host, long long elem_begin,
                            long long nelem, const int* nq1, const double* const* xi1d_dev,
                            const double* const* w1d_dev, const double* u_dev, int ncomp, double* partial_dev,
                            int npartial, double* result_dev, cudaStream_t stream, const uint8_t* trim_dev,
                            const int32_t* map_dev) {
  if (P.dim != 3 || P.dw != 3 || ncomp < 1 || ncomp > kMaxComp || nelem <= 0 || getenv("IGAB_NORMS_SF_OFF")) return IGAB_UNSUPPORTED;
  const int p = P.degree[0], q = nq1[0];
  if (P.degree[1] != p || P.degree[2] != p || nq1[1] != q || nq1[2] != q) return IGAB_UNSUPPORTED;
  if (!((p == 2 && (q == 3 || q == 4)) || (p == 3 && (q == 4 || q == 5)) || (p == 4 && (q == 5 || q == 6)))) return IGAB_UNSUPPORTED;
  igab_status st = ensure_tables3d(P, ws, q, xi1d_dev, xi1d_host, stream);
  if (st) return st;
  NormSrc S{trim_dev, map_dev, nullptr, nullptr, nullptr};
#define IGAB_SFN(PP, QQ)                                                                                              \
  if (p == PP && q == QQ)                                                                                             \
    return ncomp == 1 ? launch_sfn<PP, QQ, 1>(P, ws, S, elem_begin, nelem, w1d_dev, u_dev, ncomp, partial_dev, npartial, \
                                              result_dev, stream)                                                    \
                      : launch_sfn<PP, QQ, 3>(P, ws, S, elem_begin, nelem, w1d_dev, u_dev, ncomp, partial_dev, npartial, \
                                              result_dev, stream);
  IGAB_SFN(2, 3) IGAB_SFN(2, 4) IGAB_SFN(3, 4) IGAB_SFN(3, 5) IGAB_SFN(4, 5) IGAB_SFN(4, 6)
#undef IGAB_SFN
  return IGAB_UNSUPPORTED;
}

// Tensor rule over the (full) elements [elem_begin, elem_begin + nelem), then -- npts_list > 0 -- the trimmed elements' own
// points (device arrays) added on.  trim_dev / map_dev: nullptr on untrimmed patches.
igab_status launch_norms(const PatchDev& P, long long elem_begin, long long nelem, const int* nq1,
                         const double* const* xi1d_dev, const double* const* w1d_dev, const double* u_dev, int ncomp,
                         double* partial_dev, int npartial, double* result_dev, cudaStream_t stream,
                         const uint8_t* trim_dev, const int32_t* map_dev, long long npts_list, const int32_t* lelem_dev,
                         const double* lxi_dev, const double* lw_dev, bool tensor_done) {
  if (ncomp < 1 || ncomp > kMaxComp) {
    set_error("reduce_norms: ncomp must be 1, 2 or 3");
    return IGAB_INVALID_ARGUMENT;
  }
  long long nq = 1;
  for (int d = 0; d < P.dim; ++d) nq *= nq1[d];
  const long long npts = nelem * nq;
  NormSrc S{trim_dev, map_dev, lelem_dev, lxi_dev, lw_dev};
  if (!tensor_done) IGAB_CUDA_TRY(cudaMemsetAsync(result_dev, 0, 2 * sizeof(double), stream));
  if (npts > 0 && !tensor_done) {
    const igab_status st = run(P, S, false, elem_begin, npts, nq, nq1, xi1d_dev, w1d_dev, u_dev, ncomp, partial_dev, npartial,
                               result_dev, 0, stream);
    if (st) return st;
  }
  if (npts_list > 0)
    return run(P, S, true, 0, npts_list, 1, nq1, xi1d_dev, w1d_dev, u_dev, ncomp, partial_dev, npartial, result_dev, 1, stream);
  return IGAB_OK;
}

}  // namespace igab

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

}  // namespace

// Tensor rule over the (full) elements [elem_begin, elem_begin + nelem), then -- npts_list > 0 -- the trimmed elements' own
// points (device arrays) added on.  trim_dev / map_dev: nullptr on untrimmed patches.
igab_status launch_norms(const PatchDev& P, long long elem_begin, long long nelem, const int* nq1,
                         const double* const* xi1d_dev, const double* const* w1d_dev, const double* u_dev, int ncomp,
                         double* partial_dev, int npartial, double* result_dev, cudaStream_t stream,
                         const uint8_t* trim_dev, const int32_t* map_dev, long long npts_list, const int32_t* lelem_dev,
                         const double* lxi_dev, const double* lw_dev) {
  if (ncomp < 1 || ncomp > kMaxComp) {
    set_error("reduce_norms: ncomp must be 1, 2 or 3");
    return IGAB_INVALID_ARGUMENT;
  }
  long long nq = 1;
  for (int d = 0; d < P.dim; ++d) nq *= nq1[d];
  const long long npts = nelem * nq;
  NormSrc S{trim_dev, map_dev, lelem_dev, lxi_dev, lw_dev};
  IGAB_CUDA_TRY(cudaMemsetAsync(result_dev, 0, 2 * sizeof(double), stream));
  if (npts > 0) {
    const igab_status st = run(P, S, false, elem_begin, npts, nq, nq1, xi1d_dev, w1d_dev, u_dev, ncomp, partial_dev, npartial,
                               result_dev, 0, stream);
    if (st) return st;
  }
  if (npts_list > 0)
    return run(P, S, true, 0, npts_list, 1, nq1, xi1d_dev, w1d_dev, u_dev, ncomp, partial_dev, npartial, result_dev, 1, stream);
  return IGAB_OK;
}

}  // namespace igab

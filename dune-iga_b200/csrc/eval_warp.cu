// eval_warp.cu -- K4w: the general tensor-rule evaluation kernel.  One warp per element, degrees, rule and dimension
// taken at run time (any dim <= 3, degree <= 8 per direction, mixed degrees, any points per direction, derivative
// order <= 2).  It is what a (dim, degree, rule) without a tuned instantiation (eval_tensor_sf*.cu, eval_tensor_2d.cu,
// eval_tensor_blk.cu) runs on; the thread-per-point kernel of eval_generic.cu remains for explicit point lists.
//
// Per element (element iteration of nurbsleafgridview.hh:147-155, batched):
//   1. lanes <-> (direction, abscissa): univariate values and derivatives by A2.3
//      (BsplineBasis1D::basisFunctionDerivatives, dune/iga/bsplinealgorithms.hh:148-222) at u = xi*scaling + offset
//      (nurbsgeometry.hh:365-378), rows pre-multiplied by scaling^k (nurbsbasis.hh:93-95,105-107) -> shared memory;
//   2. lanes <-> local control points: netOfSpan (nurbsalgorithms.hh:80-89) -> shared memory, homogeneous (w x, w);
//   3. 32 points at a time, lanes <-> points: homogeneous sums  G[alpha][c] = sum_i prod_d N_d^(alpha_d) (w P)_i[c]
//      (control points are broadcast reads), then x, J^T, H, detJ, J^-T by the quotient rule
//      (nurbsalgorithms.hh:226-248 applied to the geometry; nurbsgeometry.hh:133-138,182-210,257-292);
//   4. one point at a time, lanes <-> basis functions: R^(alpha)_i by the same quotient rule with the point's W^(alpha)
//      broadcast by shuffles; N rows leave as coalesced 256 B stores, the [i][d] derivative rows are transposed
//      through a 32-function slab in shared memory so that they leave as contiguous runs too.
// HBM-bound on its stores once the local basis has >= 32 functions.
#include "bspline_device.cuh"
#include "geom_small.cuh"
#include "igab_internal.cuh"

namespace igab {

namespace {

// derivative multi-indices in the order  0 | e_d | 2 e_d | (0,1) (0,2) (1,2)   (H rows as nurbsgeometry.hh:268-289)
template <int DIM>
struct Al {
  static constexpr int NH = DIM * (DIM + 1) / 2;
  __host__ __device__ static constexpr int comp(int a, int d) {
    if (a == 0) return 0;
    if (a <= DIM) return (a - 1 == d) ? 1 : 0;
    if (a <= 2 * DIM) return (a - 1 - DIM == d) ? 2 : 0;
    const int m = a - 2 * DIM - 1;
    const int c0 = (m == 2) ? 1 : 0;
    const int c1 = (m == 0) ? 1 : 2;
    return (d == c0 || d == c1) ? 1 : 0;
  }
};

struct WarpLayout {
  int tabOff[kMaxDim];   // doubles, start of direction d's table inside the warp's region
  int strideQ[kMaxDim];  // doubles between abscissae (odd)
  int nq1[kMaxDim];
  int nqTot;             // sum nq1
  int packDoubles;       // block-level local-index table
  int offC, offW, offS1, offS2, perWarp;
};

template <int DIM, int DW, int ORDER>
__global__ void __launch_bounds__(128)
    eval_warp_kernel(PatchDev P, PointSource src, long long nelem, EvalOutDev out, WarpLayout L) {
  extern __shared__ __align__(16) double smem[];
  constexpr int NH = Al<DIM>::NH;
  constexpr int NA = (ORDER == 0) ? 1 : (ORDER == 1 ? 1 + DIM : 1 + DIM + NH);
  constexpr int C1 = DW + 1;
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int nb = P.nb;
  int n1[kMaxDim];
#pragma unroll
  for (int d = 0; d < kMaxDim; ++d) n1[d] = (d < DIM) ? P.degree[d] + 1 : 1;

  // local index i -> (l0, l1, l2), first direction fastest (mdnet.hh:245-257); the same for every element
  unsigned* pack = reinterpret_cast<unsigned*>(smem);
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    const int l0 = i % n1[0], r = i / n1[0];
    const int l1 = r % n1[1], l2 = r / n1[1];
    pack[i] = (unsigned)l0 | ((unsigned)l1 << 8) | ((unsigned)l2 << 16);
  }
  __syncthreads();

  double* sT = smem + L.packDoubles + (size_t)warp * L.perWarp;
  double* sC = sT + L.offC;
  double* sW = sT + L.offW;
  double* sS1 = sT + L.offS1;  // two slabs [32][DIM]
  double* sS2 = sT + L.offS2;  // two slabs [32][NH]
  const long long nq = src.nq;

  for (long long we = (long long)blockIdx.x * nwarp + warp; we < nelem; we += (long long)gridDim.x * nwarp) {
    // ---- element -> span, offset, scaling (nurbspatch.hh:248-253, nurbsgeometry.hh:64-68)
    int sp[DIM];
    long long cpStride[DIM], cpBase = 0;
    {
      long long h = src.elem_begin + we, stride = 1;
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        const int ed = (int)(h % P.nel[d]);
        h /= P.nel[d];
        sp[d] = P.spanTab[d][ed];
        cpStride[d] = stride;
        cpBase += (long long)(sp[d] - P.degree[d]) * stride;
        stride *= P.ncp[d];
      }
    }
    __syncwarp();  // the previous element's readers are done with sT / sC
    // ---- 1. univariate tables
    for (int item = lane; item < L.nqTot; item += 32) {
      int d = 0, q = item;
#pragma unroll
      for (int dd = 0; dd < DIM - 1; ++dd)
        if (d == dd && q >= L.nq1[dd]) {
          q -= L.nq1[dd];
          d = dd + 1;
        }
      const double* U = P.knots[d];
      const int p = P.degree[d];
      const double lo = U[sp[d]], scale = U[sp[d] + 1] - lo;
      const double u = src.xi1d[d][q] * scale + lo;
      double loc[(ORDER + 1) * kPM];
      ders1d_any(u, U, p, sp[d], ORDER, loc);
      double* t = sT + L.tabOff[d] + q * L.strideQ[d];
      double sc = 1.0;
      for (int k = 0; k <= ORDER; ++k) {
        for (int j = 0; j <= p; ++j) t[k * (p + 1) + j] = loc[k * kPM + j] * sc;
        sc *= scale;
      }
    }
    // ---- 2. control points of the span, homogeneous
    for (int i = lane; i < nb; i += 32) {
      const unsigned pk = pack[i];
      long long g = cpBase + (long long)(pk & 255u) * cpStride[0];
      if constexpr (DIM >= 2) g += (long long)((pk >> 8) & 255u) * cpStride[1];
      if constexpr (DIM >= 3) g += (long long)(pk >> 16) * cpStride[2];
      const double* c = P.cp + g * C1;
      const double w = c[DW];
#pragma unroll
      for (int k = 0; k < DW; ++k) sC[i * C1 + k] = c[k] * w;
      sC[i * C1 + DW] = w;
      sW[i] = w;
    }
    __syncwarp();

    const long long pt0 = we * nq;
    for (long long qb = 0; qb < nq; qb += 32) {
      const int cnt = (int)((nq - qb < 32) ? (nq - qb) : 32);
      // ---- 3. geometry of 32 points, lane <-> point
      double G[NA][C1];
#pragma unroll
      for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int c = 0; c < C1; ++c) G[a][c] = 0.0;
      if (lane < cnt) {
        long long q = qb + lane;
        const double* t[kMaxDim];
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          t[d] = sT + L.tabOff[d] + (int)(q % L.nq1[d]) * L.strideQ[d];
          q /= L.nq1[d];
        }
        int i = 0;
        for (int l2 = 0; l2 < n1[2]; ++l2) {
          for (int l1 = 0; l1 < n1[1]; ++l1) {
            double g12[NA];
#pragma unroll
            for (int a = 0; a < NA; ++a) {
              double f = 1.0;
              if constexpr (DIM >= 2) f = t[1][Al<DIM>::comp(a, 1) * n1[1] + l1];
              if constexpr (DIM >= 3) f *= t[2][Al<DIM>::comp(a, 2) * n1[2] + l2];
              g12[a] = f;
            }
            for (int l0 = 0; l0 < n1[0]; ++l0, ++i) {
              double hc[C1];
#pragma unroll
              for (int c = 0; c < C1; ++c) hc[c] = sC[i * C1 + c];
              double t0[ORDER + 1];
#pragma unroll
              for (int k = 0; k <= ORDER; ++k) t0[k] = t[0][k * n1[0] + l0];
#pragma unroll
              for (int a = 0; a < NA; ++a) {
                const double f = t0[Al<DIM>::comp(a, 0)] * g12[a];
#pragma unroll
                for (int c = 0; c < C1; ++c) G[a][c] += f * hc[c];
              }
            }
          }
        }
        const double W0 = G[0][DW];
        const long long pt = pt0 + qb + lane;
        double X[DW];
#pragma unroll
        for (int c = 0; c < DW; ++c) X[c] = G[0][c] / W0;
        if (out.x)
#pragma unroll
          for (int c = 0; c < DW; ++c) out.x[pt * DW + c] = X[c];
        if constexpr (ORDER >= 1) {
          double J[DIM][DW];
#pragma unroll
          for (int d = 0; d < DIM; ++d)
#pragma unroll
            for (int c = 0; c < DW; ++c) J[d][c] = (G[1 + d][c] - G[1 + d][DW] * X[c]) / W0;
          if (out.Jt)
#pragma unroll
            for (int d = 0; d < DIM; ++d)
#pragma unroll
              for (int c = 0; c < DW; ++c) out.Jt[pt * DIM * DW + d * DW + c] = J[d][c];
          if (out.detJ) out.detJ[pt] = sqrt_det_aat<DIM, DW>(J);
          if (out.JinvT) right_inv<DIM, DW>(J, out.JinvT + pt * DIM * DW);
          if constexpr (ORDER >= 2) {
            if (out.H) {
#pragma unroll
              for (int d = 0; d < DIM; ++d)
#pragma unroll
                for (int c = 0; c < DW; ++c)
                  out.H[pt * NH * DW + d * DW + c] =
                      (G[1 + DIM + d][c] - 2.0 * G[1 + d][DW] * J[d][c] - G[1 + DIM + d][DW] * X[c]) / W0;
              int m = 0;
#pragma unroll
              for (int c0 = 0; c0 < DIM; ++c0)
#pragma unroll
                for (int c1 = c0 + 1; c1 < DIM; ++c1) {
                  const int a = 1 + 2 * DIM + m;
#pragma unroll
                  for (int c = 0; c < DW; ++c)
                    out.H[pt * NH * DW + (DIM + m) * DW + c] =
                        (G[a][c] - G[1 + c0][DW] * J[c1][c] - G[1 + c1][DW] * J[c0][c] - G[a][DW] * X[c]) / W0;
                  ++m;
                }
            }
          }
        }
      }
      if (!(out.N || out.dN || out.d2N)) continue;
      // ---- 4. basis rows, one point at a time, lane <-> basis function
      int buf = 0;
      for (int tp = 0; tp < cnt; ++tp) {
        double Wt[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) Wt[a] = __shfl_sync(FULL, G[a][DW], tp);
        long long q = qb + tp;
        const double* t[kMaxDim];
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          t[d] = sT + L.tabOff[d] + (int)(q % L.nq1[d]) * L.strideQ[d];
          q /= L.nq1[d];
        }
        const long long pt = pt0 + qb + tp;
        for (int ib = 0; ib < nb; ib += 32, buf ^= 1) {
          const int i = ib + lane;
          double R[NA];
          if (i < nb) {
            const unsigned pk = pack[i];
            const int l0 = pk & 255u, l1 = (pk >> 8) & 255u, l2 = pk >> 16;
            const double w = sW[i];
#pragma unroll
            for (int a = 0; a < NA; ++a) {
              double f = t[0][Al<DIM>::comp(a, 0) * n1[0] + l0];
              if constexpr (DIM >= 2) f *= t[1][Al<DIM>::comp(a, 1) * n1[1] + l1];
              if constexpr (DIM >= 3) f *= t[2][Al<DIM>::comp(a, 2) * n1[2] + l2];
              R[a] = f * w;
            }
            // quotient rule, nurbsalgorithms.hh:226-248
            R[0] = R[0] / Wt[0];
            if constexpr (ORDER >= 1) {
#pragma unroll
              for (int d = 0; d < DIM; ++d) R[1 + d] = (R[1 + d] - Wt[1 + d] * R[0]) / Wt[0];
            }
            if constexpr (ORDER >= 2) {
#pragma unroll
              for (int d = 0; d < DIM; ++d)
                R[1 + DIM + d] = (R[1 + DIM + d] - 2.0 * Wt[1 + d] * R[1 + d] - Wt[1 + DIM + d] * R[0]) / Wt[0];
              int m = 0;
#pragma unroll
              for (int c0 = 0; c0 < DIM; ++c0)
#pragma unroll
                for (int c1 = c0 + 1; c1 < DIM; ++c1) {
                  const int a = 1 + 2 * DIM + m;
                  R[a] = (R[a] - Wt[1 + c0] * R[1 + c1] - Wt[1 + c1] * R[1 + c0] - Wt[a] * R[0]) / Wt[0];
                  ++m;
                }
            }
            if (out.N) out.N[pt * nb + i] = R[0];
          }
          if constexpr (ORDER >= 1) {
            const int rows = (nb - ib < 32) ? (nb - ib) : 32;
            const bool wd = out.dN != nullptr;
            const bool wh = (ORDER >= 2) && out.d2N != nullptr;
            double* s1 = sS1 + buf * 32 * DIM;
            double* s2 = sS2 + buf * 32 * NH;
            if (i < nb) {
              if (wd)
#pragma unroll
                for (int d = 0; d < DIM; ++d) s1[lane * DIM + d] = R[1 + d];
              if constexpr (ORDER >= 2) {
                if (wh)
#pragma unroll
                  for (int r = 0; r < NH; ++r) s2[lane * NH + r] = R[1 + DIM + r];
              }
            }
            if (wd || wh) __syncwarp();
            if (wd) {
              double* o = out.dN + (pt * nb + ib) * DIM;
#pragma unroll
              for (int k = 0; k < DIM; ++k) {
                const int j = k * 32 + lane;
                if (j < rows * DIM) o[j] = s1[j];
              }
            }
            if constexpr (ORDER >= 2) {
              if (wh) {
                double* o = out.d2N + (pt * nb + ib) * NH;
#pragma unroll
                for (int k = 0; k < NH; ++k) {
                  const int j = k * 32 + lane;
                  if (j < rows * NH) o[j] = s2[j];
                }
              }
            }
          }
        }
      }
      __syncwarp();  // slabs are reused by the next chunk from buffer 0
    }
  }
}

template <int DIM, int DW, int ORDER>
igab_status launch_o(const PatchDev& P, const PointSource& src, long long nelem, const EvalOutDev& out,
                     cudaStream_t stream) {
  constexpr int NH = Al<DIM>::NH;
  WarpLayout L{};
  int off = 0;
  L.nqTot = 0;
  for (int d = 0; d < DIM; ++d) {
    L.tabOff[d] = off;
    L.strideQ[d] = ((ORDER + 1) * (P.degree[d] + 1)) | 1;
    L.nq1[d] = src.nq1[d];
    L.nqTot += src.nq1[d];
    off += L.strideQ[d] * src.nq1[d];
  }
  auto ev = [](int x) { return (x + 1) & ~1; };
  off = ev(off);
  L.offC = off;
  off += ev(P.nb * (DW + 1));
  L.offW = off;
  off += ev(P.nb);
  L.offS1 = off;
  off += 2 * 32 * DIM;
  L.offS2 = off;
  off += 2 * 32 * NH;
  L.perWarp = ev(off);
  L.packDoubles = ev((P.nb + 1) / 2);
  auto kern = eval_warp_kernel<DIM, DW, ORDER>;
  const size_t kMaxSmem = 227 * 1024;
  int warps = 4;
  while (warps > 1 && ((size_t)L.packDoubles + (size_t)warps * L.perWarp) * sizeof(double) > kMaxSmem / 2) warps >>= 1;
  const size_t smem = ((size_t)L.packDoubles + (size_t)warps * L.perWarp) * sizeof(double);
  if (smem > kMaxSmem) return IGAB_UNSUPPORTED;  // caller falls back to the thread-per-point kernel
  KernelCfg kc;  // raises the dynamic shared-memory limit of this instantiation once
  if (igab_status st = kernel_config((const void*)kern, 128, kMaxSmem, &kc)) return st;
  int bps = 1;
  IGAB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, warps * 32, smem));
  if (bps < 1) bps = 1;
  long long blocks = (long long)kc.sms * bps;
  const long long need = (nelem + warps - 1) / warps;
  if (blocks > need) blocks = need;
  kern<<<(unsigned)blocks, warps * 32, smem, stream>>>(P, src, nelem, out, L);
  count_launch();
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

template <int DIM, int DW>
igab_status launch_dd(const PatchDev& P, const PointSource& src, long long nelem, int order, const EvalOutDev& out,
                      cudaStream_t stream) {
  if (order == 0) return launch_o<DIM, DW, 0>(P, src, nelem, out, stream);
  if (order == 1) return launch_o<DIM, DW, 1>(P, src, nelem, out, stream);
  return launch_o<DIM, DW, 2>(P, src, nelem, out, stream);
}

}  // namespace

// IGAB_UNSUPPORTED (without touching the error string) when the element's working set does not fit shared memory.
igab_status launch_eval_warp(const PatchDev& P, const PointSource& src, long long nelem, int order,
                             const EvalOutDev& out, cudaStream_t stream) {
  if (!src.tensor) return IGAB_UNSUPPORTED;
  if (nelem == 0) return IGAB_OK;
  switch (P.dim * 10 + P.dw) {
    case 11: return launch_dd<1, 1>(P, src, nelem, order, out, stream);
    case 12: return launch_dd<1, 2>(P, src, nelem, order, out, stream);
    case 13: return launch_dd<1, 3>(P, src, nelem, order, out, stream);
    case 22: return launch_dd<2, 2>(P, src, nelem, order, out, stream);
    case 23: return launch_dd<2, 3>(P, src, nelem, order, out, stream);
    case 33: return launch_dd<3, 3>(P, src, nelem, order, out, stream);
  }
  return IGAB_UNSUPPORTED;
}

}  // namespace igab

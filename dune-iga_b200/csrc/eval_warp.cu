// eval_warp.cu -- K4w: the general tensor-rule evaluation kernel.  One warp per element, degrees, rule and dimension
// taken at run time (any dim <= 3, degree <= 8 per direction, mixed degrees, any points per direction, derivative
// order <= 2).  It is what a (dim, degree, rule) without a tuned instantiation (eval_tensor_sf*.cu, eval_tensor_2d.cu,
// eval_tensor_blk.cu) runs on; the thread-per-point kernel of eval_generic.cu remains for explicit point lists.
//
// Per element (element iteration of nurbsleafgridview.hh:147-155, batched):
//   1. the element's slices of the univariate tables -> shared memory (tables_any_kernel builds them once per launch
//      for every element of every direction: A2.3 values and derivatives, BsplineBasis1D::basisFunctionDerivatives,
//      dune/iga/bsplinealgorithms.hh:148-222, at u = xi*scaling + offset, nurbsgeometry.hh:365-378, rows pre-multiplied
//      by scaling^k, nurbsbasis.hh:93-95,105-107);
//   2. lanes <-> local control points: netOfSpan (nurbsalgorithms.hh:80-89) -> shared memory, homogeneous (w x, w);
//   3. 32 points at a time, lanes <-> points: homogeneous sums  G[alpha][c] = sum_i prod_d N_d^(alpha_d) (w P)_i[c]
//      (control points are broadcast reads), then x, J^T, H, detJ, J^-T by the quotient rule
//      (nurbsalgorithms.hh:226-248 applied to the geometry; nurbsgeometry.hh:133-138,182-210,257-292);
//   4. lanes <-> basis functions (32 at a time), the chunk's points one after the other: R^(alpha)_i by the same
//      quotient rule with the point's 1/W and W^(alpha) read back from shared memory (broadcast); the products over the
//      directions >= 1 are refreshed only when the first direction's abscissa wraps; N rows leave as coalesced 256 B
//      stores, the [i][d] derivative rows are transposed through a double-buffered 32-function slab in shared memory so
//      that they leave as contiguous runs too.  One division per point (1/W), none per basis function.
// HBM-bound on its stores once the local basis has >= 32 functions.
#include <algorithm>

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
  int offC, offW, offS1, offS2, offQ, perWarp;
  long long tabGOff[kMaxDim];  // start of direction d's table in the global table buffer
  long long tabGTotal;
};

// K1 for run-time degrees: one thread per (direction, element of that direction, abscissa): A2.3 values and derivatives
// (BsplineBasis1D::basisFunctionDerivatives, dune/iga/bsplinealgorithms.hh:148-222) at u = xi*scaling + offset
// (nurbsgeometry.hh:365-378), row k pre-multiplied by scaling^k (nurbsbasis.hh:93-95,105-107).
// Layout per direction: [e_d][q][k][j], abscissa stride strideQ (odd), so an element's slice is one contiguous run.
__global__ void tables_any_kernel(PatchDev P, PointSource src, WarpLayout L, int order, double* __restrict__ tabG) {
  const int d = blockIdx.y;
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= P.nel[d] * L.nq1[d]) return;
  const int e = item / L.nq1[d], q = item % L.nq1[d];
  const double* U = P.knots[d];
  const int p = P.degree[d], sp = P.spanTab[d][e];
  const double lo = U[sp], scale = U[sp + 1] - lo;
  const double u = src.xi1d[d][q] * scale + lo;
  double loc[3 * kPM];
  ders1d_any(u, U, p, sp, order, loc);
  double* t = tabG + L.tabGOff[d] + (size_t)item * L.strideQ[d];
  double sc = 1.0;
  for (int k = 0; k <= order; ++k) {
    for (int j = 0; j <= p; ++j) t[k * (p + 1) + j] = loc[k * kPM + j] * sc;
    sc *= scale;
  }
  if ((((order + 1) * (p + 1)) & 1) == 0) t[(order + 1) * (p + 1)] = 0.0;  // the padding word is copied, keep it defined
}

template <int DIM, int DW, int ORDER>
__global__ void __launch_bounds__(128, (ORDER <= 1 ? 5 : 2))
    eval_warp_kernel(PatchDev P, PointSource src, long long nelem, EvalOutDev out, WarpLayout L,
                     const double* __restrict__ tabG) {
  extern __shared__ __align__(16) double smem[];
  constexpr int NH = Al<DIM>::NH;
  constexpr int NA = (ORDER == 0) ? 1 : (ORDER == 1 ? 1 + DIM : 1 + DIM + NH);
  constexpr int C1 = DW + 1;
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
  double* sQ = sT + L.offQ;    // [32][NA]: 1/W, W^(alpha) of the chunk's points
  const long long nq = src.nq;

  for (long long we = (long long)blockIdx.x * nwarp + warp; we < nelem; we += (long long)gridDim.x * nwarp) {
    // ---- element -> span, offset, scaling (nurbspatch.hh:248-253, nurbsgeometry.hh:64-68)
    int sp[DIM], ed[DIM];
    long long cpStride[DIM], cpBase = 0;
    {
      long long h = src.elem_begin + we, stride = 1;
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        ed[d] = (int)(h % P.nel[d]);
        h /= P.nel[d];
        sp[d] = P.spanTab[d][ed[d]];
        cpStride[d] = stride;
        cpBase += (long long)(sp[d] - P.degree[d]) * stride;
        stride *= P.ncp[d];
      }
    }
    __syncwarp();  // the previous element's readers are done with sT / sC
    // ---- 1. the element's slices of the univariate tables (built once per launch by tables_any_kernel)
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const int len = L.nq1[d] * L.strideQ[d];
      const double* g = tabG + L.tabGOff[d] + (size_t)ed[d] * len;
      double* t = sT + L.tabOff[d];
      for (int j = lane; j < len; j += 32) t[j] = __ldg(g + j);
    }
    // ---- 2. control points of the span, homogeneous and RELATIVE to the span's first control point: the sums of
    //         step 3 then carry magnitudes of the element's size instead of the patch's, which keeps the cancellation
    //         in the quotient rule (S_a - W_a x) harmless on strongly refined patches; x gets the offset back
    double pref[DW];
#pragma unroll
    for (int k = 0; k < DW; ++k) pref[k] = __ldg(P.cp + cpBase * C1 + k);
    for (int i = lane; i < nb; i += 32) {
      const unsigned pk = pack[i];
      long long g = cpBase + (long long)(pk & 255u) * cpStride[0];
      if constexpr (DIM >= 2) g += (long long)((pk >> 8) & 255u) * cpStride[1];
      if constexpr (DIM >= 3) g += (long long)(pk >> 16) * cpStride[2];
      const double* c = P.cp + g * C1;
      const double w = c[DW];
#pragma unroll
      for (int k = 0; k < DW; ++k) sC[i * C1 + k] = (c[k] - pref[k]) * w;
      sC[i * C1 + DW] = w;
      sW[i] = w;
    }
    __syncwarp();

    const long long pt0 = we * nq;
    const int nqi = (int)nq;
    for (int qb = 0; qb < nqi; qb += 32) {
      const int cnt = (nqi - qb < 32) ? (nqi - qb) : 32;
      // ---- 3. geometry of 32 points, lane <-> point
      if (lane < cnt) {
        double G[NA][C1];
#pragma unroll
        for (int a = 0; a < NA; ++a)
#pragma unroll
          for (int c = 0; c < C1; ++c) G[a][c] = 0.0;
        int q = qb + lane;
        const double* t[kMaxDim];
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          t[d] = sT + L.tabOff[d] + (q % L.nq1[d]) * L.strideQ[d];
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
              if constexpr (C1 % 2 == 0) {  // 16-byte aligned rows: broadcast LDS.128
#pragma unroll
                for (int c = 0; c < C1; c += 2) {
                  const double2 v = *reinterpret_cast<const double2*>(sC + i * C1 + c);
                  hc[c] = v.x;
                  hc[c + 1] = v.y;
                }
              } else {
#pragma unroll
                for (int c = 0; c < C1; ++c) hc[c] = sC[i * C1 + c];
              }
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
        // quotient rule on the geometry; the point's 1/W and W^(alpha) go to shared memory for phase 4
        const double inv = 1.0 / G[0][DW];
        sQ[lane * NA] = inv;
#pragma unroll
        for (int a = 1; a < NA; ++a) sQ[lane * NA + a] = G[a][DW];
        const long long pt = pt0 + qb + lane;
        double X[DW];
#pragma unroll
        for (int c = 0; c < DW; ++c) X[c] = __dmul_rn(G[0][c], inv);  // no FMA contraction: same bits for every order
        if (out.x)
#pragma unroll
          for (int c = 0; c < DW; ++c) out.x[pt * DW + c] = __dadd_rn(X[c], pref[c]);  // X stays relative below
        if constexpr (ORDER >= 1) {
          double J[DIM][DW];
#pragma unroll
          for (int d = 0; d < DIM; ++d)
#pragma unroll
            for (int c = 0; c < DW; ++c) J[d][c] = (G[1 + d][c] - G[1 + d][DW] * X[c]) * inv;
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
                      (G[1 + DIM + d][c] - 2.0 * G[1 + d][DW] * J[d][c] - G[1 + DIM + d][DW] * X[c]) * inv;
              int m = 0;
#pragma unroll
              for (int c0 = 0; c0 < DIM; ++c0)
#pragma unroll
                for (int c1 = c0 + 1; c1 < DIM; ++c1) {
                  const int a = 1 + 2 * DIM + m;
#pragma unroll
                  for (int c = 0; c < DW; ++c)
                    out.H[pt * NH * DW + (DIM + m) * DW + c] =
                        (G[a][c] - G[1 + c0][DW] * J[c1][c] - G[1 + c1][DW] * J[c0][c] - G[a][DW] * X[c]) * inv;
                  ++m;
                }
            }
          }
        }
      }
      if (!(out.N || out.dN || out.d2N)) continue;
      __syncwarp();  // sQ visible
      // ---- 4. basis rows: lane <-> basis function (32 at a time), then the chunk's points one after the other.  The
      //         points run first-direction-fastest, so the products over the directions >= 1 change only when q0 wraps.
      const bool wn = out.N != nullptr;
      const bool wd = (ORDER >= 1) && out.dN != nullptr;
      const bool wh = (ORDER >= 2) && out.d2N != nullptr;
      const int s0 = L.strideQ[0], s1q = L.strideQ[1], s2q = L.strideQ[2];
      const int nq0 = L.nq1[0], nq1b = L.nq1[1];
      int bo = 0;  // slab buffer toggle (0 / 1)
      const long long ptb = pt0 + qb;
      for (int ib = 0; ib < nb; ib += 32) {
        const int i = ib + lane;
        const bool act = i < nb;
        const int rows = (nb - ib < 32) ? (nb - ib) : 32;
        const unsigned pk = pack[act ? i : 0];
        const double w = sW[act ? i : 0];
        const double* b0 = sT + L.tabOff[0] + (int)(pk & 255u);
        const double* b1 = sT + L.tabOff[1] + (int)((pk >> 8) & 255u);
        const double* b2 = sT + L.tabOff[2] + (int)(pk >> 16);
        int q0 = qb % nq0, q1 = 0, q2 = 0;
        if constexpr (DIM >= 2) {
          const int r = qb / nq0;
          q1 = r % nq1b;
          q2 = r / nq1b;
        }
        const double* p0 = b0 + q0 * s0;
        const double* p1 = b1 + q1 * s1q;
        const double* p2 = b2 + q2 * s2q;
        const double* wq = sQ;
        double* pN = out.N + ptb * nb + i;
        double* pD = nullptr;
        double* pH = nullptr;
        if constexpr (ORDER >= 1) pD = out.dN + (ptb * nb + ib) * DIM + lane;
        if constexpr (ORDER >= 2) pH = out.d2N + (ptb * nb + ib) * NH + lane;
        bool okD[DIM > 0 ? DIM : 1], okH[NH];
#pragma unroll
        for (int k = 0; k < DIM; ++k) okD[k] = wd && (k * 32 + lane < rows * DIM);
#pragma unroll
        for (int k = 0; k < NH; ++k) okH[k] = wh && (k * 32 + lane < rows * NH);
        bool fresh = true;
        double g12[NA];
        for (int tp = 0; tp < cnt; ++tp) {
          if (fresh) {
            fresh = false;
#pragma unroll
            for (int a = 0; a < NA; ++a) {
              double f = w;
              if constexpr (DIM >= 2) f *= p1[Al<DIM>::comp(a, 1) * n1[1]];
              if constexpr (DIM >= 3) f *= p2[Al<DIM>::comp(a, 2) * n1[2]];
              g12[a] = f;
            }
          }
          double t0[ORDER + 1];
#pragma unroll
          for (int k = 0; k <= ORDER; ++k) t0[k] = p0[k * n1[0]];
          const double inv = wq[0];  // broadcast reads
          double R[NA];
#pragma unroll
          for (int a = 0; a < NA; ++a) R[a] = t0[Al<DIM>::comp(a, 0)] * g12[a];
          // quotient rule, nurbsalgorithms.hh:226-248
          R[0] *= inv;
          if constexpr (ORDER >= 1) {
#pragma unroll
            for (int d = 0; d < DIM; ++d) R[1 + d] = (R[1 + d] - wq[1 + d] * R[0]) * inv;
          }
          if constexpr (ORDER >= 2) {
#pragma unroll
            for (int d = 0; d < DIM; ++d)
              R[1 + DIM + d] = (R[1 + DIM + d] - 2.0 * wq[1 + d] * R[1 + d] - wq[1 + DIM + d] * R[0]) * inv;
            int m = 0;
#pragma unroll
            for (int c0 = 0; c0 < DIM; ++c0)
#pragma unroll
              for (int c1 = c0 + 1; c1 < DIM; ++c1) {
                const int a = 1 + 2 * DIM + m;
                R[a] = (R[a] - wq[1 + c0] * R[1 + c1] - wq[1 + c1] * R[1 + c0] - wq[a] * R[0]) * inv;
                ++m;
              }
          }
          if (act && wn) *pN = R[0];
          if constexpr (ORDER >= 1) {
            double* s1 = sS1 + bo * (32 * DIM);
            double* s2 = sS2 + bo * (32 * NH);
            if (act) {
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
#pragma unroll
            for (int k = 0; k < DIM; ++k)
              if (okD[k]) pD[k * 32] = s1[k * 32 + lane];
            if constexpr (ORDER >= 2) {
#pragma unroll
              for (int k = 0; k < NH; ++k)
                if (okH[k]) pH[k * 32] = s2[k * 32 + lane];
            }
            bo ^= 1;
            pD += (size_t)nb * DIM;
            if constexpr (ORDER >= 2) pH += (size_t)nb * NH;
          }
          pN += nb;
          wq += NA;
          p0 += s0;
          if (++q0 == nq0) {
            q0 = 0;
            p0 = b0;
            fresh = true;
            p1 += s1q;
            if (++q1 == nq1b) {
              q1 = 0;
              p1 = b1;
              p2 += s2q;
            }
          }
        }
      }
      __syncwarp();  // sQ and the slabs are rewritten by the next chunk
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same scheme for explicit point lists (igab_eval_points in 1-D / 3-D and for degrees without a tuned list kernel):
// a warp takes 32 consecutive points of the list.  Every lane builds the univariate rows of ITS point in shared memory
// -- derivative rows by A2.3 at the unclamped u, and the value row exactly as the reference's value path does it (A2.2
// with the clamp to the patch and the end-point exits, bsplinealgorithms.hh:96-137), kept as an extra row: N and x come
// from that row with its own weight sum, the derivative formulas use the A2.3 values, as NurbsLocalBasis /
// NURBSGeometry do.  Control points and weights are read from global memory (points of a list are grouped by element in
// practice: the reads hit L1).  Phases 3 and 4 as above.
struct ListLayout {
  int tabOff[kMaxDim];  // start of direction d's rows inside a lane's table
  int laneStride;       // doubles per lane (odd)
  int offS1, offS2, offQ, perWarp, packDoubles;
};

template <int DIM, int DW, int ORDER>
__global__ void __launch_bounds__(128)
    eval_warp_list_kernel(PatchDev P, PointSource src, long long npts, EvalOutDev out, ListLayout L) {
  extern __shared__ __align__(16) double smem[];
  constexpr int NH = Al<DIM>::NH;
  constexpr int NA = (ORDER == 0) ? 1 : (ORDER == 1 ? 1 + DIM : 1 + DIM + NH);
  constexpr int C1 = DW + 1;
  constexpr int KV = ORDER + 1;  // table row of the reference's value path
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int nb = P.nb;
  int n1[kMaxDim];
#pragma unroll
  for (int d = 0; d < kMaxDim; ++d) n1[d] = (d < DIM) ? P.degree[d] + 1 : 1;
  long long cpStride[kMaxDim] = {1, 0, 0};
  if constexpr (DIM >= 2) cpStride[1] = P.ncp[0];
  if constexpr (DIM >= 3) cpStride[2] = (long long)P.ncp[0] * P.ncp[1];

  unsigned* pack = reinterpret_cast<unsigned*>(smem);
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    const int l0 = i % n1[0], r = i / n1[0];
    const int l1 = r % n1[1], l2 = r / n1[1];
    pack[i] = (unsigned)l0 | ((unsigned)l1 << 8) | ((unsigned)l2 << 16);
  }
  __syncthreads();
  double* sT = smem + L.packDoubles + (size_t)warp * L.perWarp;
  double* sS1 = sT + L.offS1;
  double* sS2 = sT + L.offS2;
  double* sQ = sT + L.offQ;  // [32][NA + 1]: 1/W, W^(alpha) (A2.3 rows), 1/W of the value row

  const long long nchunk = (npts + 31) / 32;
  for (long long ch = (long long)blockIdx.x * nwarp + warp; ch < nchunk; ch += (long long)gridDim.x * nwarp) {
    const long long c0 = ch * 32;
    const int cnt = (int)((npts - c0 < 32) ? (npts - c0) : 32);
    __syncwarp();
    // ---- 1. this lane's point: element, span, u; univariate rows -> its slice of shared memory
    long long cpBase = 0;
    double* myT = sT + (size_t)lane * L.laneStride;
    if (lane < cnt) {
      long long h = src.elem[c0 + lane];
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        const int ed = (int)(h % P.nel[d]);
        h /= P.nel[d];
        const int sp = P.spanTab[d][ed], p = P.degree[d];
        cpBase += (long long)(sp - p) * cpStride[d];
        const double* U = P.knots[d];
        const double lo = U[sp], scale = U[sp + 1] - lo;
        const double u = src.xi[(c0 + lane) * DIM + d] * scale + lo;  // nurbsgeometry.hh:365-378
        double loc[(ORDER + 1) * kPM], val[kPM];
        ders1d_any(u, U, p, sp, ORDER, loc);
        basis1d(u, U, P.nknots[d], p, sp, val);
        double* t = myT + L.tabOff[d];
        double sc = 1.0;
        for (int k = 0; k <= ORDER; ++k) {
          for (int j = 0; j <= p; ++j) t[k * (p + 1) + j] = loc[k * kPM + j] * sc;  // nurbsbasis.hh:93-95,105-107
          sc *= scale;
        }
        for (int j = 0; j <= p; ++j) t[KV * (p + 1) + j] = val[j];
      }
    }
    __syncwarp();
    // ---- 3. geometry, lane <-> point (control points from global memory, relative to the span's first one)
    if (lane < cnt) {
      double G[NA][C1], GV[C1];
#pragma unroll
      for (int c = 0; c < C1; ++c) GV[c] = 0.0;
#pragma unroll
      for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int c = 0; c < C1; ++c) G[a][c] = 0.0;
      double pref[DW];
#pragma unroll
      for (int k = 0; k < DW; ++k) pref[k] = __ldg(P.cp + cpBase * C1 + k);
      const double* t0p = myT + L.tabOff[0];
      const double* t1p = myT + L.tabOff[1];
      const double* t2p = myT + L.tabOff[2];
      for (int l2 = 0; l2 < n1[2]; ++l2) {
        for (int l1 = 0; l1 < n1[1]; ++l1) {
          double g12[NA], gv = 1.0;
          if constexpr (DIM >= 2) gv = t1p[KV * n1[1] + l1];
          if constexpr (DIM >= 3) gv *= t2p[KV * n1[2] + l2];
#pragma unroll
          for (int a = 0; a < NA; ++a) {
            double f = 1.0;
            if constexpr (DIM >= 2) f = t1p[Al<DIM>::comp(a, 1) * n1[1] + l1];
            if constexpr (DIM >= 3) f *= t2p[Al<DIM>::comp(a, 2) * n1[2] + l2];
            g12[a] = f;
          }
          const double* crow = P.cp + (cpBase + l1 * cpStride[1] + l2 * cpStride[2]) * C1;
          for (int l0 = 0; l0 < n1[0]; ++l0) {
            double hc[C1];
            const double w = __ldg(crow + l0 * C1 + DW);
#pragma unroll
            for (int c = 0; c < DW; ++c) hc[c] = (__ldg(crow + l0 * C1 + c) - pref[c]) * w;
            hc[DW] = w;
            const double fv = t0p[KV * n1[0] + l0] * gv;
#pragma unroll
            for (int c = 0; c < C1; ++c) GV[c] += fv * hc[c];
            if constexpr (ORDER >= 1) {
              double t0[ORDER + 1];
#pragma unroll
              for (int k = 0; k <= ORDER; ++k) t0[k] = t0p[k * n1[0] + l0];
#pragma unroll
              for (int a = 0; a < NA; ++a) {
                const double f = t0[Al<DIM>::comp(a, 0)] * g12[a];
#pragma unroll
                for (int c = 0; c < C1; ++c) G[a][c] += f * hc[c];
              }
            }
          }
        }
      }
      const long long pt = c0 + lane;
      const double invV = 1.0 / GV[DW];
      sQ[lane * (NA + 1) + NA] = invV;
      if (out.x)
#pragma unroll
        for (int c = 0; c < DW; ++c) out.x[pt * DW + c] = __dadd_rn(__dmul_rn(GV[c], invV), pref[c]);
      if constexpr (ORDER >= 1) {
        const double inv = 1.0 / G[0][DW];
        sQ[lane * (NA + 1)] = inv;
#pragma unroll
        for (int a = 1; a < NA; ++a) sQ[lane * (NA + 1) + a] = G[a][DW];
        double X[DW], J[DIM][DW];
#pragma unroll
        for (int c = 0; c < DW; ++c) X[c] = G[0][c] * inv;
#pragma unroll
        for (int d = 0; d < DIM; ++d)
#pragma unroll
          for (int c = 0; c < DW; ++c) J[d][c] = (G[1 + d][c] - G[1 + d][DW] * X[c]) * inv;
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
                    (G[1 + DIM + d][c] - 2.0 * G[1 + d][DW] * J[d][c] - G[1 + DIM + d][DW] * X[c]) * inv;
            int m = 0;
#pragma unroll
            for (int c0d = 0; c0d < DIM; ++c0d)
#pragma unroll
              for (int c1d = c0d + 1; c1d < DIM; ++c1d) {
                const int a = 1 + 2 * DIM + m;
#pragma unroll
                for (int c = 0; c < DW; ++c)
                  out.H[pt * NH * DW + (DIM + m) * DW + c] =
                      (G[a][c] - G[1 + c0d][DW] * J[c1d][c] - G[1 + c1d][DW] * J[c0d][c] - G[a][DW] * X[c]) * inv;
                ++m;
              }
          }
        }
      }
    }
    if (!(out.N || out.dN || out.d2N)) continue;
    __syncwarp();
    // ---- 4. basis rows: lanes <-> basis functions, the chunk's points one after the other
    const bool wn = out.N != nullptr;
    const bool wd = (ORDER >= 1) && out.dN != nullptr;
    const bool wh = (ORDER >= 2) && out.d2N != nullptr;
    int bo = 0;
    for (int tp = 0; tp < cnt; ++tp) {
      const long long base = __shfl_sync(FULL, cpBase, tp);
      const double* tb = sT + (size_t)tp * L.laneStride;
      const double* wq = sQ + tp * (NA + 1);
      const long long pt = c0 + tp;
      for (int ib = 0; ib < nb; ib += 32, bo ^= 1) {
        const int i = ib + lane;
        const bool act = i < nb;
        const int rows = (nb - ib < 32) ? (nb - ib) : 32;
        double R[NA], RV = 0.0;
        if (act) {
          const unsigned pk = pack[i];
          const int l0 = pk & 255u, l1 = (pk >> 8) & 255u, l2 = pk >> 16;
          const double w = __ldg(P.cp + (base + l0 + l1 * cpStride[1] + l2 * cpStride[2]) * C1 + DW);
          const double* b0 = tb + L.tabOff[0] + l0;
          const double* b1 = tb + L.tabOff[1] + l1;
          const double* b2 = tb + L.tabOff[2] + l2;
          double fv = b0[KV * n1[0]];
          if constexpr (DIM >= 2) fv *= b1[KV * n1[1]];
          if constexpr (DIM >= 3) fv *= b2[KV * n1[2]];
          RV = (fv * w) * wq[NA];  // R = N-hat w / W on the value path (nurbsalgorithms.hh:176-179)
          if constexpr (ORDER >= 1) {
#pragma unroll
            for (int a = 0; a < NA; ++a) {
              double f = b0[Al<DIM>::comp(a, 0) * n1[0]];
              if constexpr (DIM >= 2) f *= b1[Al<DIM>::comp(a, 1) * n1[1]];
              if constexpr (DIM >= 3) f *= b2[Al<DIM>::comp(a, 2) * n1[2]];
              R[a] = f * w;
            }
            const double inv = wq[0];
            R[0] *= inv;
#pragma unroll
            for (int d = 0; d < DIM; ++d) R[1 + d] = (R[1 + d] - wq[1 + d] * R[0]) * inv;
            if constexpr (ORDER >= 2) {
#pragma unroll
              for (int d = 0; d < DIM; ++d)
                R[1 + DIM + d] = (R[1 + DIM + d] - 2.0 * wq[1 + d] * R[1 + d] - wq[1 + DIM + d] * R[0]) * inv;
              int m = 0;
#pragma unroll
              for (int c0d = 0; c0d < DIM; ++c0d)
#pragma unroll
                for (int c1d = c0d + 1; c1d < DIM; ++c1d) {
                  const int a = 1 + 2 * DIM + m;
                  R[a] = (R[a] - wq[1 + c0d] * R[1 + c1d] - wq[1 + c1d] * R[1 + c0d] - wq[a] * R[0]) * inv;
                  ++m;
                }
            }
          }
          if (wn) out.N[pt * nb + i] = RV;
        }
        if constexpr (ORDER >= 1) {
          double* s1 = sS1 + bo * (32 * DIM);
          double* s2 = sS2 + bo * (32 * NH);
          if (act) {
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
    __syncwarp();
  }
}

template <int DIM, int DW, int ORDER>
igab_status launch_list_o(const PatchDev& P, const PointSource& src, long long npts, const EvalOutDev& out,
                          cudaStream_t stream) {
  constexpr int NH = Al<DIM>::NH;
  constexpr int NA = (ORDER == 0) ? 1 : (ORDER == 1 ? 1 + DIM : 1 + DIM + NH);
  ListLayout L{};
  int off = 0;
  for (int d = 0; d < DIM; ++d) {
    L.tabOff[d] = off;
    off += (ORDER + 2) * (P.degree[d] + 1);
  }
  L.laneStride = off | 1;
  auto ev = [](int x) { return (x + 1) & ~1; };
  off = ev(32 * L.laneStride);
  L.offS1 = off;
  off += 2 * 32 * DIM;
  L.offS2 = off;
  off += 2 * 32 * NH;
  L.offQ = off;
  off += 32 * (NA + 1);
  L.perWarp = ev(off);
  L.packDoubles = ev((P.nb + 1) / 2);
  auto kern = eval_warp_list_kernel<DIM, DW, ORDER>;
  const size_t kMaxSmem = 227 * 1024;
  int warps = 4;
  while (warps > 1 && ((size_t)L.packDoubles + (size_t)warps * L.perWarp) * sizeof(double) > kMaxSmem / 2) warps >>= 1;
  const size_t smem = ((size_t)L.packDoubles + (size_t)warps * L.perWarp) * sizeof(double);
  if (smem > kMaxSmem) return IGAB_UNSUPPORTED;
  KernelCfg kc;
  if (igab_status st = kernel_config((const void*)kern, 128, kMaxSmem, &kc)) return st;
  int bps = 1;
  IGAB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, warps * 32, smem));
  if (bps < 1) bps = 1;
  long long blocks = (long long)kc.sms * bps;
  const long long need = ((npts + 31) / 32 + warps - 1) / warps;
  if (blocks > need) blocks = need;
  kern<<<(unsigned)blocks, warps * 32, smem, stream>>>(P, src, npts, out, L);
  count_launch();
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

template <int DIM, int DW>
igab_status launch_list_dd(const PatchDev& P, const PointSource& src, long long npts, int order, const EvalOutDev& out,
                           cudaStream_t stream) {
  if (order == 0) return launch_list_o<DIM, DW, 0>(P, src, npts, out, stream);
  if (order == 1) return launch_list_o<DIM, DW, 1>(P, src, npts, out, stream);
  return launch_list_o<DIM, DW, 2>(P, src, npts, out, stream);
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
  L.offQ = off;
  off += 32 * (1 + DIM + NH);
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
  // univariate tables of this rule for all elements of every direction: a few hundred KB, stream-ordered scratch
  long long tot = 0;
  int maxItems = 0;
  for (int d = 0; d < DIM; ++d) {
    L.tabGOff[d] = tot;
    tot += (long long)P.nel[d] * src.nq1[d] * L.strideQ[d];
    maxItems = std::max(maxItems, P.nel[d] * src.nq1[d]);
  }
  L.tabGTotal = tot;
  double* tabG = nullptr;
  IGAB_CUDA_TRY(cudaMallocAsync(&tabG, sizeof(double) * tot, stream));
  tables_any_kernel<<<dim3((maxItems + 127) / 128, DIM), 128, 0, stream>>>(P, src, L, ORDER, tabG);
  kern<<<(unsigned)blocks, warps * 32, smem, stream>>>(P, src, nelem, out, L, tabG);
  count_launch(2);
  IGAB_CUDA_TRY(cudaGetLastError());
  IGAB_CUDA_TRY(cudaFreeAsync(tabG, stream));
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
  if (nelem == 0) return IGAB_OK;
  if (!src.tensor) {  // explicit point list: nelem is the number of points
    switch (P.dim * 10 + P.dw) {
      case 11: return launch_list_dd<1, 1>(P, src, nelem, order, out, stream);
      case 12: return launch_list_dd<1, 2>(P, src, nelem, order, out, stream);
      case 13: return launch_list_dd<1, 3>(P, src, nelem, order, out, stream);
      case 22: return launch_list_dd<2, 2>(P, src, nelem, order, out, stream);
      case 23: return launch_list_dd<2, 3>(P, src, nelem, order, out, stream);
      case 33: return launch_list_dd<3, 3>(P, src, nelem, order, out, stream);
    }
    return IGAB_UNSUPPORTED;
  }
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

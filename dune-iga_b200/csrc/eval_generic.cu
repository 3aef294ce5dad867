// eval_generic.cu -- general per-point NURBS evaluation kernel (K4 in SURVEY.md 2: arbitrary element-local points,
// any degree <= 8 per direction, any rule), one thread per point.  It is the path for trimmed-element batches
// (variable points per element, no tensor structure to share) and for every (dim, degree, rule) that has no
// sum-factorised instantiation in eval_tensor_sf.cu.  Follows the reference's per-point algorithm step by step so
// that it agrees with it to a few ulps:
//   1-D values      BsplineBasis1D::basisFunctions            dune/iga/bsplinealgorithms.hh:96-137   (A2.2 + end exits)
//   1-D derivatives BsplineBasis1D::basisFunctionDerivatives  dune/iga/bsplinealgorithms.hh:148-222  (A2.3)
//   rational basis  Nurbs::basisFunctions                     dune/iga/nurbsalgorithms.hh:160-180
//   rational ders   Nurbs::basisFunctionDerivatives           dune/iga/nurbsalgorithms.hh:193-250 (|alpha| <= 2 only)
//   geometry        NURBSGeometry::global/jacobianTransposed/secondDerivativeOfPosition/integrationElement/
//                   jacobianInverseTransposed                 dune/iga/nurbsgeometry.hh:133-138,182-210,257-292
//   local scaling   nurbsbasis.hh:88-95,101-107; nurbsgeometry.hh:365-378
// HBM-bound on its stores (thread-per-point stores are sector-complete but not line-coalesced); the tuned tensor
// kernel is the one measured against the roofline.
#include "bspline_device.cuh"
#include "geom_small.cuh"
#include "igab_internal.cuh"

namespace igab {

namespace {

constexpr int PM = kPM;  // 9

// A2.3, rows 0..K (K <= 2 here), dN[k][PM]
template <int K>
__device__ void ders1d(double u, const double* __restrict__ U, int p, int sp, double (*dN)[PM]) {
  double left[PM], right[PM], ndu[PM][PM], a[2][PM];
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
  for (int j = 0; j <= p; ++j) dN[0][j] = ndu[j][p];
  if (K == 0) return;
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
      dN[k][r] = d;
      const int t = s1; s1 = s2; s2 = t;
    }
  }
  int fac = p;
  for (int k = 1; k <= K; ++k) {
    for (int j = 0; j <= p; ++j) dN[k][j] *= fac;
    fac *= (p - k);
  }
}

template <int DIM>
struct Alpha {
  // index 0: zero; 1..DIM: e_d; DIM+1..2DIM: 2 e_d; then mixed (0,1),(0,2),(1,2)
  static constexpr int NH = DIM * (DIM + 1) / 2;
  static constexpr int N2 = 1 + DIM + NH;
  __device__ static __forceinline__ int comp(int a, int d) {
    if (a == 0) return 0;
    if (a <= DIM) return (a - 1 == d) ? 1 : 0;
    if (a <= 2 * DIM) return (a - 1 - DIM == d) ? 2 : 0;
    const int m = a - 2 * DIM - 1;  // 0:(0,1) 1:(0,2) 2:(1,2)
    const int c0 = (m == 2) ? 1 : 0;
    const int c1 = (m == 0) ? 1 : 2;
    return (d == c0 || d == c1) ? 1 : 0;
  }
};

template <int DIM, int DW, int ORDER>
__global__ void __launch_bounds__(128) eval_generic_kernel(PatchDev P, PointSource src, long long npts, EvalOutDev out) {
  const long long pt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= npts) return;
  constexpr int NH = Alpha<DIM>::NH;
  constexpr int NA = (ORDER == 0) ? 1 : (ORDER == 1 ? 1 + DIM : 1 + DIM + NH);

  // ---- which element, which local coordinates
  long long e;
  double xi[DIM];
  if (src.tensor) {
    e = src.elem_begin + pt / src.nq;
    long long q = pt % src.nq;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      xi[d] = src.xi1d[d][q % src.nq1[d]];
      q /= src.nq1[d];
    }
  } else {
    e = src.elem[pt];
#pragma unroll
    for (int d = 0; d < DIM; ++d) xi[d] = src.xi[pt * DIM + d];
  }

  // ---- element -> span, scaling, u  (nurbspatch.hh:248-253, nurbsgeometry.hh:64-68,365-378)
  int sp[DIM], p[DIM];
  double scale[DIM];
  double vals[DIM][PM];
  double ders[DIM][ORDER + 1][PM];
  long long cpBase = 0, cpStride[DIM];
  {
    long long h = e, stride = 1;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const int ed = (int)(h % P.nel[d]);
      h /= P.nel[d];
      p[d] = P.degree[d];
      sp[d] = P.spanTab[d][ed];
      const double* U = P.knots[d];
      const double lo = U[sp[d]];
      scale[d] = U[sp[d] + 1] - lo;
      const double u = xi[d] * scale[d] + lo;
      basis1d(u, U, P.nknots[d], p[d], sp[d], vals[d]);
      ders1d<ORDER>(u, U, p[d], sp[d], ders[d]);
      cpStride[d] = stride;
      cpBase += (long long)(sp[d] - p[d]) * stride;
      stride *= P.ncp[d];
    }
  }
  const int nb = P.nb;
  const double* __restrict__ cp = P.cp;

  // ---- pass 1: W^(alpha) = sum_i w_i prod_d N_d^(alpha_d)  (sequential, i-fastest, as mdnet.hh:480-487)
  double W[NA], Wv = 0.0;
#pragma unroll
  for (int a = 0; a < NA; ++a) W[a] = 0.0;
  {
    int l[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) l[d] = 0;
    for (int i = 0; i < nb; ++i) {
      long long g = cpBase;
#pragma unroll
      for (int d = 0; d < DIM; ++d) g += l[d] * cpStride[d];
      const double w = cp[g * (DW + 1) + DW];
      double v = vals[0][l[0]];
#pragma unroll
      for (int d = 1; d < DIM; ++d) v *= vals[d][l[d]];
      Wv += v * w;
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        double f = ders[0][Alpha<DIM>::comp(a, 0)][l[0]];
#pragma unroll
        for (int d = 1; d < DIM; ++d) f *= ders[d][Alpha<DIM>::comp(a, d)][l[d]];
        W[a] += f * w;
      }
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        if (++l[d] <= p[d]) break;
        l[d] = 0;
      }
    }
  }

  // ---- pass 2: quotient rule (nurbsalgorithms.hh:226-248), outputs, geometry sums
  double X[DW], J[DIM][DW], H[NH][DW];
#pragma unroll
  for (int c = 0; c < DW; ++c) X[c] = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d)
#pragma unroll
    for (int c = 0; c < DW; ++c) J[d][c] = 0.0;
#pragma unroll
  for (int r = 0; r < NH; ++r)
#pragma unroll
    for (int c = 0; c < DW; ++c) H[r][c] = 0.0;

  double* __restrict__ oN = out.N ? out.N + pt * nb : nullptr;
  double* __restrict__ odN = (ORDER >= 1 && out.dN) ? out.dN + pt * nb * DIM : nullptr;
  double* __restrict__ od2N = (ORDER >= 2 && out.d2N) ? out.d2N + pt * nb * NH : nullptr;
  {
    int l[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) l[d] = 0;
    for (int i = 0; i < nb; ++i) {
      long long g = cpBase;
#pragma unroll
      for (int d = 0; d < DIM; ++d) g += l[d] * cpStride[d];
      double Pc[DW + 1];
#pragma unroll
      for (int c = 0; c <= DW; ++c) Pc[c] = cp[g * (DW + 1) + c];
      const double w = Pc[DW];
      // values: R = (Nhat w) / W   (nurbsalgorithms.hh:176-179)
      double v = vals[0][l[0]];
#pragma unroll
      for (int d = 1; d < DIM; ++d) v *= vals[d][l[d]];
      const double Rv = (v * w) / Wv;
      if (oN) oN[i] = Rv;
#pragma unroll
      for (int c = 0; c < DW; ++c) X[c] += Rv * Pc[c];
      if constexpr (ORDER >= 1) {
        double R[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          double f = ders[0][Alpha<DIM>::comp(a, 0)][l[0]];
#pragma unroll
          for (int d = 1; d < DIM; ++d) f *= ders[d][Alpha<DIM>::comp(a, d)][l[d]];
          R[a] = f * w;
        }
        R[0] = R[0] / W[0];
#pragma unroll
        for (int d = 0; d < DIM; ++d) R[1 + d] = (R[1 + d] - W[1 + d] * R[0]) / W[0];
        if constexpr (ORDER >= 2) {
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            R[1 + DIM + d] = (R[1 + DIM + d] - 2.0 * W[1 + d] * R[1 + d] - W[1 + DIM + d] * R[0]) / W[0];
          int m = 0;
#pragma unroll
          for (int c0 = 0; c0 < DIM; ++c0)
#pragma unroll
            for (int c1 = c0 + 1; c1 < DIM; ++c1) {
              const int a = 1 + 2 * DIM + m;
              R[a] = (R[a] - W[1 + c0] * R[1 + c1] - W[1 + c1] * R[1 + c0] - W[a] * R[0]) / W[0];
              ++m;
            }
        }
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          if (odN) odN[i * DIM + d] = R[1 + d] * scale[d];  // nurbsbasis.hh:93-95
#pragma unroll
          for (int c = 0; c < DW; ++c) J[d][c] += R[1 + d] * Pc[c];
        }
        if constexpr (ORDER >= 2) {
#pragma unroll
          for (int r = 0; r < NH; ++r) {
            double s = 1.0;
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
              const int ad = Alpha<DIM>::comp(1 + DIM + r, d);
              if (ad >= 1) s *= scale[d];
              if (ad == 2) s *= scale[d];
            }
            if (od2N) od2N[i * NH + r] = R[1 + DIM + r] * s;  // nurbsbasis.hh:105-107
#pragma unroll
            for (int c = 0; c < DW; ++c) H[r][c] += R[1 + DIM + r] * Pc[c];
          }
        }
      }
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        if (++l[d] <= p[d]) break;
        l[d] = 0;
      }
    }
  }

  if (out.x)
#pragma unroll
    for (int c = 0; c < DW; ++c) out.x[pt * DW + c] = X[c];
  if constexpr (ORDER >= 1) {
#pragma unroll
    for (int d = 0; d < DIM; ++d)
#pragma unroll
      for (int c = 0; c < DW; ++c) J[d][c] *= scale[d];  // nurbsgeometry.hh:189-194
    if (out.Jt)
#pragma unroll
      for (int d = 0; d < DIM; ++d)
#pragma unroll
        for (int c = 0; c < DW; ++c) out.Jt[pt * DIM * DW + d * DW + c] = J[d][c];
    if (out.detJ) out.detJ[pt] = sqrt_det_aat<DIM, DW>(J);
    if (out.JinvT) right_inv<DIM, DW>(J, out.JinvT + pt * DIM * DW);
    if constexpr (ORDER >= 2) {
      if (out.H)
#pragma unroll
        for (int r = 0; r < NH; ++r) {
          double s = 1.0;
#pragma unroll
          for (int d = 0; d < DIM; ++d) {
            const int ad = Alpha<DIM>::comp(1 + DIM + r, d);
            if (ad >= 1) s *= scale[d];
            if (ad == 2) s *= scale[d];
          }
#pragma unroll
          for (int c = 0; c < DW; ++c) out.H[pt * NH * DW + r * DW + c] = H[r][c] * s;  // nurbsgeometry.hh:262-289
        }
    }
  }
}

template <int DIM, int DW>
igab_status launch_dd(const PatchDev& P, const PointSource& src, long long npts, int order, const EvalOutDev& out,
                      cudaStream_t stream) {
  const int threads = 128;
  const long long blocks = (npts + threads - 1) / threads;
  if (blocks > 0x7fffffffLL) {
    set_error("eval_generic: too many points for one launch");
    return IGAB_INVALID_ARGUMENT;
  }
  if (npts == 0) return IGAB_OK;
  if (order == 0)
    eval_generic_kernel<DIM, DW, 0><<<(unsigned)blocks, threads, 0, stream>>>(P, src, npts, out);
  else if (order == 1)
    eval_generic_kernel<DIM, DW, 1><<<(unsigned)blocks, threads, 0, stream>>>(P, src, npts, out);
  else
    eval_generic_kernel<DIM, DW, 2><<<(unsigned)blocks, threads, 0, stream>>>(P, src, npts, out);
  count_launch();
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

}  // namespace

igab_status launch_eval_generic(const PatchDev& P, const PointSource& src, long long npts, int order,
                                const EvalOutDev& out, cudaStream_t stream) {
  switch (P.dim * 10 + P.dw) {
    case 11: return launch_dd<1, 1>(P, src, npts, order, out, stream);
    case 12: return launch_dd<1, 2>(P, src, npts, order, out, stream);
    case 13: return launch_dd<1, 3>(P, src, npts, order, out, stream);
    case 22: return launch_dd<2, 2>(P, src, npts, order, out, stream);
    case 23: return launch_dd<2, 3>(P, src, npts, order, out, stream);
    case 33: return launch_dd<3, 3>(P, src, npts, order, out, stream);
  }
  set_error("eval: unsupported (dim, dimworld)");
  return IGAB_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------------------------------
// partial(alpha) for an arbitrary multi-index: one warp per point, lanes stride over the local basis functions.
// R^(g) for all g <= alpha by the generalised quotient rule (nurbsalgorithms.hh:226-248; formula as written there,
// see oracle/igab_oracle.c igo_nurbs_ders for the one deviation from the reference's loop).
namespace {
constexpr int kMaxG = 64;

__device__ __forceinline__ double binom_d(int n, int k) {
  double r = 1.0;
  for (int i = 0; i < k; ++i) r = r * (n - i) / (i + 1);
  return r;
}

struct AlphaArg {
  int a[kMaxDim];
};

// smem per warp: ders [dim][maxA+1][PM] + W[kMaxG]
__global__ void __launch_bounds__(128) partial_kernel(PatchDev P, long long npts, const int* __restrict__ elem,
                                                      const double* __restrict__ xi, AlphaArg al,
                                                      double* __restrict__ out) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long pt = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  const int dim = P.dim;
  int maxA = 0, ng = 1, gsz[kMaxDim];
  for (int d = 0; d < dim; ++d) {
    maxA = max(maxA, al.a[d]);
    gsz[d] = al.a[d] + 1;
    ng *= gsz[d];
  }
  const int perWarp = kMaxDim * (maxA + 1) * PM + kMaxG;
  double* ders = smem + (size_t)warp * perWarp;
  double* Wg = ders + kMaxDim * (maxA + 1) * PM;
  if (pt >= npts) return;  // whole warp leaves together

  int sp[kMaxDim];
  double scl = 1.0;
  long long cpBase = 0, cpStride[kMaxDim];
  {
    long long h = elem[pt], stride = 1;
    for (int d = 0; d < dim; ++d) {
      const int ed = (int)(h % P.nel[d]);
      h /= P.nel[d];
      sp[d] = P.spanTab[d][ed];
      const double* U = P.knots[d];
      const double sc = U[sp[d] + 1] - U[sp[d]];
      const double u = xi[pt * dim + d] * sc + U[sp[d]];
      for (int q = 0; q < al.a[d]; ++q) scl *= sc;
      if (lane == d) ders1d_any(u, U, P.degree[d], sp[d], al.a[d], ders + (size_t)d * (maxA + 1) * PM);
      cpStride[d] = stride;
      cpBase += (long long)(sp[d] - P.degree[d]) * stride;
      stride *= P.ncp[d];
    }
  }
  __syncwarp();
  const int nb = P.nb;
  const int DW = P.dw;
  // W^(g): warp reduction over i for every g
  for (int g = 0; g < ng; ++g) {
    int ga[kMaxDim], hg = g;
    for (int d = 0; d < dim; ++d) { ga[d] = hg % gsz[d]; hg /= gsz[d]; }
    double s = 0.0;
    for (int i = lane; i < nb; i += 32) {
      int hh = i;
      long long gi = cpBase;
      double f = 1.0;
      for (int d = 0; d < dim; ++d) {
        const int l = hh % (P.degree[d] + 1);
        hh /= (P.degree[d] + 1);
        gi += l * cpStride[d];
        f *= ders[((size_t)d * (maxA + 1) + ga[d]) * PM + l];
      }
      s += f * P.cp[gi * (DW + 1) + DW];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) Wg[g] = s;
  }
  __syncwarp();
  // per basis function: recursion over g ascending
  for (int i = lane; i < nb; i += 32) {
    double R[kMaxG];
    int l[kMaxDim], hh = i;
    long long gi = cpBase;
    for (int d = 0; d < dim; ++d) {
      l[d] = hh % (P.degree[d] + 1);
      hh /= (P.degree[d] + 1);
      gi += l[d] * cpStride[d];
    }
    const double w = P.cp[gi * (DW + 1) + DW];
    for (int g = 0; g < ng; ++g) {
      int ga[kMaxDim], hg = g;
      double f = 1.0;
      for (int d = 0; d < dim; ++d) {
        ga[d] = hg % gsz[d];
        hg /= gsz[d];
        f *= ders[((size_t)d * (maxA + 1) + ga[d]) * PM + l[d]];
      }
      double r = f * w;
      for (int b = 1; b <= g; ++b) {
        int hb = b, amb = 0, st = 1;
        bool ok = true;
        double c = 1.0;
        for (int d = 0; d < dim; ++d) {
          const int bd = hb % gsz[d];
          hb /= gsz[d];
          if (bd > ga[d]) ok = false;
          c *= binom_d(ga[d], bd);
          amb += (ga[d] - bd) * st;
          st *= gsz[d];
        }
        if (ok) r -= c * Wg[b] * R[amb];
      }
      R[g] = r / Wg[0];
    }
    out[pt * nb + i] = R[ng - 1] * scl;  // nurbsbasis.hh:105-107
  }
}
}  // namespace

igab_status launch_partial(const PatchDev& P, long long npts, const int* elem, const double* xi, const int* alpha,
                           double* out, cudaStream_t stream) {
  AlphaArg al{};
  int ng = 1, maxA = 0;
  for (int d = 0; d < P.dim; ++d) {
    if (alpha[d] < 0) {
      set_error("partial: negative derivative order");
      return IGAB_INVALID_ARGUMENT;
    }
    al.a[d] = alpha[d];
    ng *= alpha[d] + 1;
    maxA = alpha[d] > maxA ? alpha[d] : maxA;
  }
  if (ng > kMaxG) {
    set_error("partial: prod(alpha_d+1) > 64 not supported");
    return IGAB_UNSUPPORTED;
  }
  if (npts == 0) return IGAB_OK;
  const int warps = 4;
  const size_t smem = (size_t)warps * (kMaxDim * (maxA + 1) * PM + kMaxG) * sizeof(double);
  const long long blocks = (npts + warps - 1) / warps;
  partial_kernel<<<(unsigned)blocks, warps * 32, smem, stream>>>(P, npts, elem, xi, al, out);
  count_launch();
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

}  // namespace igab

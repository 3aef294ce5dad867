// eval_tensor_blk.cu -- evaluation at tensor rules for LARGE local bases in 3-D (p = 4: 125 functions, 4 lane-slots):
// one BLOCK per element instead of one warp (eval_tensor_sf.cu), so that the element's geometry scratch is shared by the
// KS = ceil(nb/32) warps instead of being replicated per warp (30 KB -> 37 KB per FOUR warps: 24 resident warps per SM
// instead of 6).  Same algorithm and layouts as eval_sf3d_kernel:
//   tables + control points -> shared; geometry by block-wide sum factorisation; per point x, J, detJ, J^-T;
//   warp k owns basis slot k (functions 32k .. 32k+31) and loops over the Q^3 points with its univariate rows in
//   registers; N rows leave as coalesced 256-byte stores, dN rows are transposed through a per-warp shared slab.
// Replaces the same reference chain: nurbsbasis.hh:77-96, nurbsgeometry.hh:133-138,182-210, nurbsalgorithms.hh:193-250.
#include <algorithm>
#include <cstdlib>

#include "igab_internal.cuh"
#include "sf_geometry.cuh"

namespace igab {

namespace {

template <int P, int Q>
struct BCfg {
  static constexpr int NB1 = P + 1, NB = NB1 * NB1 * NB1, NQ = Q * Q * Q, Q2 = Q * Q;
  static constexpr int KS = (NB + 31) / 32, NT = KS * 32;
  static constexpr int TAB = Q * NB1 * 2;
  static constexpr int HS = 4 * NB1 + 1, S1C = NB1 * NB1 + 1, S2S = NB1 | 1;
  static constexpr int ev(int x) { return (x + 1) & ~1; }
  static constexpr int SZ_T = 3 * TAB;
  static constexpr int SZ_W = ev(NB);
  static constexpr int SZ_H = ev(NB1 * NB1 * HS);
  static constexpr int SZ_S1 = 2 * Q * 4 * S1C;
  static constexpr int SZ_G = 16 * NQ;
  static constexpr int SZ_A = ev((SZ_H + SZ_S1) > SZ_G ? (SZ_H + SZ_S1) : SZ_G);
  static constexpr int SZ_S2 = ev(3 * 4 * Q2 * S2S);
  static constexpr int SZ_C = 4 * NQ;
  static constexpr int SZ_ST = ev(9 * NQ);
  static constexpr int SZ_D = KS * 2 * 96;
  static constexpr int SZ_B = ev(SZ_S2 > (SZ_C + SZ_ST + SZ_D) ? SZ_S2 : (SZ_C + SZ_ST + SZ_D));
  static constexpr int SZ_TOTAL = SZ_T + SZ_W + SZ_A + SZ_B;
};

__device__ __forceinline__ double2 lds2b(const double* p) { return *reinterpret_cast<const double2*>(p); }

template <int P, int Q>
__global__ void __launch_bounds__(BCfg<P, Q>::NT) eval_sf3d_blk_kernel(PatchDev Pd, long long elem_begin, long long nelem,
                                                                       const double* __restrict__ tab0,
                                                                       const double* __restrict__ tab1,
                                                                       const double* __restrict__ tab2, EvalOutDev out,
                                                                       unsigned int* __restrict__ counter) {
  using C = BCfg<P, Q>;
  constexpr int NB1 = C::NB1, NB = C::NB, NQ = C::NQ, Q2 = C::Q2, NT = C::NT;
  extern __shared__ __align__(16) double bsm[];
  __shared__ unsigned int sNext;
  double* const sT = bsm;
  double* const sW = sT + C::SZ_T;
  double* const rA = sW + C::SZ_W;
  double* const rB = rA + C::SZ_A;
  double* const sG = rA;  // output of sf_contract
  double* const sS2 = rB;
  double* const sC = rB;
  double* const sSt = rB + C::SZ_C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* const sD = sSt + C::SZ_ST + warp * 2 * 96;

  for (;;) {
    __syncthreads();  // previous element finished with every buffer (and with sNext)
    if (tid == 0) sNext = atomicAdd(counter, 1u);
    __syncthreads();
    const long long el = sNext;
    if (el >= nelem) break;
    const long long e = elem_begin + el;
    int e0, e1, e2;
    {
      long long h = e;
      e0 = (int)(h % Pd.nel[0]); h /= Pd.nel[0];
      e1 = (int)(h % Pd.nel[1]); h /= Pd.nel[1];
      e2 = (int)h;
    }
    sf_load<P, Q, NT>(Pd, tid, e0, e1, e2, tab0, tab1, tab2, sT, sW, rA);
    sf_contract<P, Q, NT>(tid, sT, rA, sS2);
    // ---- per point (one per thread when NQ <= NT, else strided): geometry in registers, constants for phase 4
    const long long ptBase = el * NQ;
    constexpr int PPT = (NQ + NT - 1) / NT;
    double gx[PPT][3], gJ[PPT][9], gdet[PPT];
#pragma unroll
    for (int m = 0; m < PPT; ++m) {
      const int pt = tid + NT * m;
      gdet[m] = 1.0;
      if (pt < NQ) {
        const double W = sG[3 * NQ + pt], invW = 1.0 / W;
#pragma unroll
        for (int c = 0; c < 3; ++c) gx[m][c] = sG[c * NQ + pt] * invW;
        double Wd[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          Wd[d] = sG[((d + 1) * 4 + 3) * NQ + pt];
#pragma unroll
          for (int c = 0; c < 3; ++c) gJ[m][3 * d + c] = (sG[((d + 1) * 4 + c) * NQ + pt] - Wd[d] * gx[m][c]) * invW;
        }
        const double* J = gJ[m];
        gdet[m] = J[0] * (J[4] * J[8] - J[5] * J[7]) - J[1] * (J[3] * J[8] - J[5] * J[6]) + J[2] * (J[3] * J[7] - J[4] * J[6]);
        // sC lives in region B (stage-2 scratch, dead): written after the sync above
        double* cst = sC + 4 * pt;
        cst[0] = invW; cst[1] = Wd[0]; cst[2] = Wd[1]; cst[3] = Wd[2];
      }
    }
    // field by field through the block-wide staging area, flushed as one contiguous run per field
    auto flush = [&](double* g, int width) {
      __syncthreads();
      for (int j = tid; j < width * NQ; j += NT) g[j] = sSt[j];
      __syncthreads();
    };
    if (out.x) {
#pragma unroll
      for (int m = 0; m < PPT; ++m) {
        const int pt = tid + NT * m;
        if (pt < NQ)
          for (int c = 0; c < 3; ++c) sSt[3 * pt + c] = gx[m][c];
      }
      flush(out.x + ptBase * 3, 3);
    }
    if (out.Jt) {
#pragma unroll
      for (int m = 0; m < PPT; ++m) {
        const int pt = tid + NT * m;
        if (pt < NQ)
          for (int c = 0; c < 9; ++c) sSt[9 * pt + c] = gJ[m][c];
      }
      flush(out.Jt + ptBase * 9, 9);
    }
    if (out.JinvT) {
#pragma unroll
      for (int m = 0; m < PPT; ++m) {
        const int pt = tid + NT * m;
        if (pt < NQ) {
          const double* J = gJ[m];
          const double id = 1.0 / gdet[m];
          double* s = sSt + 9 * pt;  // J^-T[c][d] = cof[d][c] / det
          s[0] = (J[4] * J[8] - J[5] * J[7]) * id; s[1] = (J[2] * J[7] - J[1] * J[8]) * id; s[2] = (J[1] * J[5] - J[2] * J[4]) * id;
          s[3] = (J[5] * J[6] - J[3] * J[8]) * id; s[4] = (J[0] * J[8] - J[2] * J[6]) * id; s[5] = (J[2] * J[3] - J[0] * J[5]) * id;
          s[6] = (J[3] * J[7] - J[4] * J[6]) * id; s[7] = (J[1] * J[6] - J[0] * J[7]) * id; s[8] = (J[0] * J[4] - J[1] * J[3]) * id;
        }
      }
      flush(out.JinvT + ptBase * 9, 9);
    }
    if (out.detJ) {
#pragma unroll
      for (int m = 0; m < PPT; ++m) {
        const int pt = tid + NT * m;
        if (pt < NQ) out.detJ[ptBase + pt] = fabs(gdet[m]);
      }
    }
    __syncthreads();  // sC complete for every warp
    // ---- basis functions: warp k owns slot k
    if (out.N || out.dN) {
      const int k = warp;
      const int i = lane + 32 * k;
      const bool valid = i < NB;
      const int ii = valid ? i : 0;
      const int i0 = ii % NB1, i1 = (ii / NB1) % NB1, i2 = ii / (NB1 * NB1);
      const double w = valid ? sW[ii] : 0.0;
      double n0[Q], d0[Q], n1[Q], d1[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const double2 a = lds2b(sT + (q * NB1 + i0) * 2), b = lds2b(sT + C::TAB + (q * NB1 + i1) * 2);
        n0[q] = a.x; d0[q] = a.y; n1[q] = b.x; d1[q] = b.y;
      }
      const int nrow = min(32, NB - 32 * k);
      double* gN = out.N ? out.N + ptBase * NB + i : nullptr;
      double* gD = out.dN ? out.dN + (ptBase * NB + 32 * k) * 3 : nullptr;
      int buf = 0;
#pragma unroll 1
      for (int q2 = 0; q2 < Q; ++q2) {
        const double2 t2 = lds2b(sT + 2 * C::TAB + (q2 * NB1 + i2) * 2);
        const double wn2 = w * t2.x, wd2 = w * t2.y;
#pragma unroll
        for (int q1 = 0; q1 < Q; ++q1) {
#pragma unroll
          for (int q0 = 0; q0 < Q; ++q0) {
            const int pt = q0 + Q * q1 + Q2 * q2;
            const double2 ca = lds2b(sC + 4 * pt), cb = lds2b(sC + 4 * pt + 2);
            const double t00 = n0[q0] * n1[q1];
            const double R = t00 * wn2 * ca.x;
            if (gN && valid) gN[(size_t)pt * NB] = R;
            if (gD) {
              double* sd = sD + buf * 96;
              sd[3 * lane + 0] = ((d0[q0] * n1[q1]) * wn2 - ca.y * R) * ca.x;
              sd[3 * lane + 1] = ((n0[q0] * d1[q1]) * wn2 - cb.x * R) * ca.x;
              sd[3 * lane + 2] = (t00 * wd2 - cb.y * R) * ca.x;
              __syncwarp();
              double* g = gD + (size_t)pt * NB * 3;
#pragma unroll
              for (int r = 0; r < 3; ++r) {
                const int j = lane + 32 * r;
                if (j < 3 * nrow) g[j] = sd[j];
              }
              buf ^= 1;
            }
          }
        }
      }
    }
  }
}

template <int P, int Q>
igab_status launch_blk(const PatchDev& Pd, SfWorkspace& ws, const PointSource& src, const double* const* xi1d_host,
                       long long nelem, const EvalOutDev& out, cudaStream_t stream) {
  using C = BCfg<P, Q>;
  if (nelem > 0x7fffffffLL) {
    set_error("eval_tensor: more than 2^31 elements in one launch");
    return IGAB_INVALID_ARGUMENT;
  }
  igab_status st = ensure_tables3d(Pd, ws, Q, src.xi1d, xi1d_host, stream);
  if (st) return st;
  const size_t smem = (size_t)C::SZ_TOTAL * sizeof(double);
  KernelCfg kc;
  if (igab_status cst = kernel_config((const void*)eval_sf3d_blk_kernel<P, Q>, C::NT, smem, &kc)) return cst;
  const int bps = kc.blocksPerSm, sms = kc.sms;
  const long long blocks = std::min<long long>(nelem, (long long)bps * sms);
  unsigned int* ctr = ws.counter + (ws.counterSlot++ % 64);
  IGAB_CUDA_TRY(cudaMemsetAsync(ctr, 0, sizeof(unsigned int), stream));
  eval_sf3d_blk_kernel<P, Q><<<(unsigned)blocks, C::NT, smem, stream>>>(Pd, src.elem_begin, nelem, ws.tab[0], ws.tab[1],
                                                                       ws.tab[2], out, ctr);
  count_launch();
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

}  // namespace

bool eval_tensor_blk_supported(const PatchDev& P, const int* nq1, int order, const EvalOutDev& out) {
  (void)order;
  if (P.dim != 3 || P.dw != 3 || out.d2N || out.H) return false;
  if (P.degree[0] != P.degree[1] || P.degree[0] != P.degree[2]) return false;
  if (nq1[0] != nq1[1] || nq1[0] != nq1[2]) return false;
  // 7 points per direction is the reference's default rule for p = 4 (order dim * p = 12, nurbsgridentity.hh:123-124)
  return P.degree[0] == 4 && (nq1[0] == 5 || nq1[0] == 6 || nq1[0] == 7);
}

igab_status launch_eval_tensor_blk(const PatchDev& P, SfWorkspace& ws, const PointSource& src,
                                   const double* const* xi1d_host, long long nelem, int order, const EvalOutDev& out,
                                   cudaStream_t stream) {
  if (!eval_tensor_blk_supported(P, src.nq1, order, out)) {
    set_error("eval_tensor: no block kernel for this configuration");
    return IGAB_UNSUPPORTED;
  }
  EvalOutDev o = out;
  if (order < 1) o.dN = o.Jt = o.detJ = o.JinvT = nullptr;
  if (src.nq1[0] == 5) return launch_blk<4, 5>(P, ws, src, xi1d_host, nelem, o, stream);
  if (src.nq1[0] == 6) return launch_blk<4, 6>(P, ws, src, xi1d_host, nelem, o, stream);
  return launch_blk<4, 7>(P, ws, src, xi1d_host, nelem, o, stream);
}

}  // namespace igab

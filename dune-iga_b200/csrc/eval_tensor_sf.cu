// eval_tensor_sf.cu -- the tuned evaluation path for tensor quadrature rules (K1 + K2 of SURVEY.md 2).
//
// What it replaces: the user loop "for element: for gauss point: evaluateFunction / evaluateJacobian / global /
// jacobianTransposed / integrationElement" (dune/iga/nurbsbasis.hh:77-96, dune/iga/nurbsgeometry.hh:133-138,182-203),
// i.e. per point Nurbs<dim>::basisFunctionDerivatives (dune/iga/nurbsalgorithms.hh:193-250) + dim dots.
//
// B200 design (not a port of the per-point algorithm):
//  K1  univariate tables: for every direction d, element e_d and abscissa q the p+1 non-zero B-splines and their first
//      derivative (A2.3, bsplinealgorithms.hh:148-222), derivative pre-multiplied by the span length so that all later
//      arithmetic is in element-local (xi) coordinates (nurbsbasis.hh:93-95, nurbsgeometry.hh:189-194).
//      n_el_d * Q * (p+1) * 2 doubles per direction -- computed once per call, lives in L2.
//  K2  one WARP per element, persistent over elements:
//      (1) the element's (p+1)^3 control points (x,y,z,w) are fetched with 16-byte vector loads (rows of p+1
//          consecutive points are contiguous) and kept in shared memory in homogeneous form (wx, wy, wz, w);
//      (2) geometry by SUM FACTORISATION of the homogeneous net: G^(a)(q) = sum_i dN^(a)_i(q) (wP_i, w_i) for
//          a in {0, e0, e1, e2}, three 1-D contractions through shared memory (O(p^4) instead of O(p^6) per element);
//          then x = C/W, J_d = (C_d - W_d x)/W (quotient rule on the geometry), detJ = |det J|, J^-T;
//      (3) basis functions: lane <-> basis function i (i = lane + 32 k), loop over the Q^3 points with the three
//          univariate rows of the lane held in registers:  R_i = w_i N0 N1 N2 / W,
//          dR_i/dxi_d = (w_i d_d(N0 N1 N2) - W_d R_i)/W  (nurbsalgorithms.hh:226-248 for |alpha| = 1);
//      (4) stores: N as 256-byte coalesced rows straight from registers; dN [i][d]-interleaved rows are transposed
//          through a 768-byte shared-memory slab so that every store instruction writes full 32-byte sectors of
//          consecutive addresses; x / Jt / detJ / JinvT staged per 32 points and flushed as contiguous runs.
//  The pass is bound by HBM writes (2,152 B per point at p=3, 4^3 points); FP64 work is ~15 DP instructions per
//  (point, basis function), i.e. < 25 % of the DP pipe at the HBM roofline.
#include <algorithm>
#include <cstdlib>

#include "bspline_device.cuh"
#include "igab_internal.cuh"
#include "sf_geometry.cuh"

namespace igab {

namespace {

// ---- K1 ---------------------------------------------------------------------------------------------------------------
// tab[d] layout: [e_d][q][i][a], a = 0 value, a = 1 d/dxi ; one thread per (e_d, q)
__global__ void tables_kernel(PatchDev P, int d, int nq, const double* __restrict__ xi, double* __restrict__ tab) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.nel[d] * nq) return;
  const int e = t / nq, q = t % nq;
  const int p = P.degree[d];
  const int sp = P.spanTab[d][e];
  const double* U = P.knots[d];
  const double lo = U[sp], scale = U[sp + 1] - lo;
  const double u = xi[q] * scale + lo;  // nurbsgeometry.hh:365-378
  double dN[2 * kPM];
  ders1d_any(u, U, p, sp, 1, dN);
  double* o = tab + ((size_t)t * (p + 1)) * 2;
  for (int i = 0; i <= p; ++i) {
    o[2 * i] = dN[i];
    o[2 * i + 1] = dN[kPM + i] * scale;
  }
}

// weighted copy: wtab[e][q][i][a] = tab[e][q][i][a] * wfac[span(e) - p + i]
__global__ void wtables_kernel(PatchDev P, int d, int nq, const double* __restrict__ tab, double* __restrict__ wtab) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int nb1 = P.degree[d] + 1;
  if (t >= P.nel[d] * nq * nb1) return;
  const int i = t % nb1, e = t / (nb1 * nq);
  const double w = P.wfac[d][P.spanTab[d][e] - P.degree[d] + i];
  wtab[2 * t] = tab[2 * t] * w;
  wtab[2 * t + 1] = tab[2 * t + 1] * w;
}

template <int P, int Q>
struct Cfg {
  static constexpr int NB1 = P + 1;
  static constexpr int NB = NB1 * NB1 * NB1;
  static constexpr int NQ = Q * Q * Q;
  static constexpr int KS = (NB + 31) / 32;     // basis slots per lane
  static constexpr int PS = (NQ + 31) / 32;     // point slots per lane
  static constexpr int TAB = Q * NB1 * 2;       // doubles per direction
  // shared-memory layout per warp (in doubles)
  static constexpr int HS = 4 * NB1 + 1;        // stride of H over i12 = i1 + NB1 i2 (odd: conflict-free stage-1 reads)
  static constexpr int S1C = NB1 * NB1 + 1;     // stride of S1 over c
  static constexpr int S2S = NB1 | 1;           // stride of S2 over q01 (odd)
  static constexpr int ev(int x) { return (x + 1) & ~1; }  // regions stay 16-byte aligned
  static constexpr int SZ_T = 3 * TAB;
  static constexpr int SZ_W = ev(NB);           // weights by local index
  static constexpr int SZ_H = ev(NB1 * NB1 * HS);
  static constexpr int SZ_S1 = 2 * Q * 4 * S1C;
  static constexpr int SZ_G = 16 * NQ;          // [alpha(4)][c(4)][pt]
  // k-inner variant: all basis slots of a lane share (i0, i1) (32 % (p+1)^2 == 0, i.e. p = 3 or 1), so one point's
  // complete dN row [nb][3] is produced at once and leaves through ONE async bulk store (TMA, UBLKCP) per point.
  static constexpr bool KINNER = (32 % (NB1 * NB1) == 0) && (NB % 32 == 0);
  static constexpr int NBUF = 4;                // ring depth of bulk-store slabs
  static constexpr int SZ_D = KINNER ? NBUF * NB * 3 : 2 * 96;  // dN slabs (bulk ring | transposition double buffer)
  static constexpr int SZ_A = ev((SZ_H + SZ_S1) > (SZ_G) ? (SZ_H + SZ_S1) : (SZ_G));
  static constexpr int SZ_S2 = 3 * 4 * Q * Q * S2S;
  static constexpr int SZ_C = 4 * NQ;           // invW, W_0, W_1, W_2 per point
  static constexpr int SZ_ST = 32 * 9;          // staging of one geometry field for 32 points
  static constexpr int SZ_STD = SZ_ST > SZ_D ? SZ_ST : SZ_D;  // geometry staging (phase 3) aliases the dN slabs (phase 4)
  static constexpr int SZ_B = ev(SZ_S2 > (SZ_C + SZ_STD) ? SZ_S2 : (SZ_C + SZ_STD));
  static constexpr int SZ_WARP = ev(SZ_T + SZ_W + SZ_A + SZ_B);
};

__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }

// ---- async bulk copy shared -> global (TMA engine; SASS: UBLKCP), bulk-group completion
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_hint(void* gdst, const void* ssrc, unsigned bytes, unsigned long long pol) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(s),
               "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// coalesced copy of n doubles from shared memory to global memory by one warp
__device__ __forceinline__ void warp_flush(double* __restrict__ dst, const double* __restrict__ src, int n, int lane) {
  for (int j = lane; j < n; j += 32) dst[j] = src[j];
}

template <int P, int Q>
__global__ void __launch_bounds__(128) eval_sf3d_kernel(PatchDev Pd, long long elem_begin, long long nelem,
                                                        const double* __restrict__ tab0,
                                                        const double* __restrict__ tab1,
                                                        const double* __restrict__ tab2, EvalOutDev out, int tune,
                                                        unsigned int* __restrict__ counter) {
  using C = Cfg<P, Q>;
  constexpr int NB1 = C::NB1, NB = C::NB, NQ = C::NQ, Q2 = Q * Q;
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31;
  const int warpInBlock = threadIdx.x >> 5;
  double* const sT = smem + (size_t)warpInBlock * C::SZ_WARP;
  double* const sW = sT + C::SZ_T;
  double* const rA = sW + C::SZ_W;
  double* const rB = rA + C::SZ_A;
  double* const sH = rA;               // stage 1 input
  double* const sS1 = rA + C::SZ_H;    // stage 1 output
  double* const sG = rA;               // stage 3 output (H, S1 dead by then)
  double* const sS2 = rB;              // stage 2 output
  double* const sC = rB;               // after stage 3 (S2 dead)
  double* const sSt = rB + C::SZ_C;
  double* const sD = sSt;              // phase 4 only (staging dead)

  // dynamic scheduling: every warp draws the next element from a global counter, so elements in flight stay
  // consecutive (output streams stay compact) and slow SMs (far-die L2, DRAM channel contention) do not leave a tail
  for (;;) {
    unsigned int nxt = 0;
    if (lane == 0) nxt = atomicAdd(counter, 1u);
    const long long el = (long long)__shfl_sync(0xffffffffu, nxt, 0);
    if (el >= nelem) break;
    const long long e = elem_begin + el;
    int e0, e1, e2;
    {
      long long h = e;
      e0 = (int)(h % Pd.nel[0]); h /= Pd.nel[0];
      e1 = (int)(h % Pd.nel[1]); h /= Pd.nel[1];
      e2 = (int)h;
    }
    // ---- (0)-(2): tables + control points -> shared, geometry by sum factorisation (sf_geometry.cuh)
    sf_load<P, Q, 32>(Pd, lane, e0, e1, e2, tab0, tab1, tab2, sT, sW, rA);
    sf_contract<P, Q, 32>(lane, sT, rA, sS2);
    // ---- (3) per point: x = C/W, J_d = (C_d - W_d x)/W, detJ, J^-T ; constants for (4)
    const long long ptBase = el * NQ;
#pragma unroll 1
    for (int m = 0; m < C::PS; ++m) {
      const int pt = lane + 32 * m;
      const int nvalid = min(32, NQ - 32 * m);
      double x[3], J[3][3], det = 0.0;
      if (pt < NQ) {
        const double W = sG[3 * NQ + pt];
        const double invW = 1.0 / W;
#pragma unroll
        for (int c = 0; c < 3; ++c) x[c] = sG[c * NQ + pt] * invW;
        double Wd[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          Wd[d] = sG[((d + 1) * 4 + 3) * NQ + pt];
#pragma unroll
          for (int c = 0; c < 3; ++c) J[d][c] = (sG[((d + 1) * 4 + c) * NQ + pt] - Wd[d] * x[c]) * invW;
        }
        double* cst = sC + 4 * pt;
        cst[0] = invW; cst[1] = Wd[0]; cst[2] = Wd[1]; cst[3] = Wd[2];
        det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
              J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
      }
      const long long p0 = ptBase + 32 * m;
      if (out.x) {
        if (pt < NQ) { sSt[3 * lane] = x[0]; sSt[3 * lane + 1] = x[1]; sSt[3 * lane + 2] = x[2]; }
        __syncwarp();
        warp_flush(out.x + p0 * 3, sSt, 3 * nvalid, lane);
        __syncwarp();
      }
      if (out.Jt) {
        if (pt < NQ) {
#pragma unroll
          for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int c = 0; c < 3; ++c) sSt[9 * lane + 3 * d + c] = J[d][c];
        }
        __syncwarp();
        warp_flush(out.Jt + p0 * 9, sSt, 9 * nvalid, lane);
        __syncwarp();
      }
      if (out.JinvT) {
        if (pt < NQ) {
          // J^-T [c][d] = cof(J)[d][c] / det  (square case of rightInvA, nurbsgeometry.hh:205-210)
          const double id = 1.0 / det;
          const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
                       c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
          const double c10 = J[0][2] * J[2][1] - J[0][1] * J[2][2], c11 = J[0][0] * J[2][2] - J[0][2] * J[2][0],
                       c12 = J[0][1] * J[2][0] - J[0][0] * J[2][1];
          const double c20 = J[0][1] * J[1][2] - J[0][2] * J[1][1], c21 = J[0][2] * J[1][0] - J[0][0] * J[1][2],
                       c22 = J[0][0] * J[1][1] - J[0][1] * J[1][0];
          // inverse of J (as a matrix with rows d) is adj/det with adj[i][j] = cof[j][i]; JinvT[c][d] = inv(J)^T[c][d]
          // = inv(J)[d][c] ... with Jt rows = d: JinvT = (Jt)^-1 transposed appropriately: JinvT[c][d] = cof[d][c]/det
          double* s = sSt + 9 * lane;
          s[0] = c00 * id; s[1] = c10 * id; s[2] = c20 * id;
          s[3] = c01 * id; s[4] = c11 * id; s[5] = c21 * id;
          s[6] = c02 * id; s[7] = c12 * id; s[8] = c22 * id;
        }
        __syncwarp();
        warp_flush(out.JinvT + p0 * 9, sSt, 9 * nvalid, lane);
        __syncwarp();
      }
      if (out.detJ && pt < NQ) out.detJ[p0 + lane] = fabs(det);
    }
    __syncwarp();
    // ---- (4) basis functions
    if constexpr (C::KINNER) {
      if (out.N || out.dN) {
        constexpr int KS = C::KS;
        const int i0 = lane % NB1, i1 = (lane / NB1) % NB1, i2b = lane / (NB1 * NB1);
        constexpr int I2STEP = 32 / (NB1 * NB1);  // i2 of slot k = i2b + k * I2STEP
        double w[KS];
#pragma unroll
        for (int k = 0; k < KS; ++k) w[k] = sW[lane + 32 * k];
        double n0[Q], d0[Q], n1[Q], d1[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const double2 a = lds2(sT + (q * NB1 + i0) * 2), b = lds2(sT + C::TAB + (q * NB1 + i1) * 2);
          n0[q] = a.x; d0[q] = a.y; n1[q] = b.x; d1[q] = b.y;
        }
        double* gN = out.N ? out.N + ptBase * NB + lane : nullptr;
        double* gD = out.dN ? out.dN + ptBase * NB * 3 : nullptr;
        // the bulk path needs 16-byte aligned global rows (row = NB*24 bytes, a multiple of 16 here)
        const bool bulk = gD && ((reinterpret_cast<unsigned long long>(gD) & 15ull) == 0);
        const unsigned long long pol = policy_evict_first();
#pragma unroll 1
        for (int q2 = 0; q2 < Q; ++q2) {
          double wn2[KS], wd2[KS];
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            const double2 t2 = lds2(sT + 2 * C::TAB + (q2 * NB1 + i2b + k * I2STEP) * 2);
            wn2[k] = w[k] * t2.x;
            wd2[k] = w[k] * t2.y;
          }
#pragma unroll
          for (int q1 = 0; q1 < Q; ++q1) {
#pragma unroll
            for (int q0 = 0; q0 < Q; ++q0) {
              const int pt = q0 + Q * q1 + Q2 * q2;
              const double2 ca = lds2(sC + 4 * pt), cb = lds2(sC + 4 * pt + 2);  // invW, W0 | W1, W2
              const double t00 = n0[q0] * n1[q1], t10 = d0[q0] * n1[q1], t01 = n0[q0] * d1[q1];
              double* sd = sD + ((q0 + Q * q1) % C::NBUF) * (NB * 3);
              if (gD) {
                if (bulk) {
                  if (lane == 0) bulk_wait_read<C::NBUF - 1>();  // the slab written NBUF points ago has been read
                }
                __syncwarp();
              }
#pragma unroll
              for (int k = 0; k < KS; ++k) {
                const double R = (t00 * wn2[k]) * ca.x;
                if (gN) { if (tune & 1) __stcs(gN + (size_t)pt * NB + 32 * k, R); else gN[(size_t)pt * NB + 32 * k] = R; }
                if (gD) {
                  double* o = sd + 3 * (lane + 32 * k);
                  o[0] = (t10 * wn2[k] - ca.y * R) * ca.x;
                  o[1] = (t01 * wn2[k] - cb.x * R) * ca.x;
                  o[2] = (t00 * wd2[k] - cb.y * R) * ca.x;
                }
              }
              if (gD) {
                double* g = gD + (size_t)pt * NB * 3;
                if (bulk) {
                  fence_proxy_async_smem();  // generic-proxy writes -> visible to the async proxy
                  __syncwarp();
                  if (lane == 0) {
                    if (tune & 1)
                      bulk_store_hint(g, sd, NB * 3 * 8, pol);
                    else
                      bulk_store(g, sd, NB * 3 * 8);
                    bulk_commit();
                  }
                } else {
                  __syncwarp();
                  for (int j = lane; j < NB * 3; j += 32) g[j] = sd[j];
                }
              }
            }
          }
        }
        if (bulk && lane == 0) bulk_wait_read<0>();  // slabs alias stage-2 scratch of the next element
        __syncwarp();
      }
    } else if (out.N || out.dN) {
#pragma unroll 1
      for (int k = 0; k < C::KS; ++k) {
        const int i = lane + 32 * k;
        const bool valid = i < NB;
        const int ii = valid ? i : 0;
        const int i0 = ii % NB1, i1 = (ii / NB1) % NB1, i2 = ii / (NB1 * NB1);
        const double w = valid ? sW[ii] : 0.0;
        double n0[Q], d0[Q], n1[Q], d1[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const double2 a = lds2(sT + (q * NB1 + i0) * 2), b = lds2(sT + C::TAB + (q * NB1 + i1) * 2);
          n0[q] = a.x; d0[q] = a.y; n1[q] = b.x; d1[q] = b.y;
        }
        const int nrow = min(32, NB - 32 * k);  // basis functions in this slot
        double* gN = out.N ? out.N + ptBase * NB + i : nullptr;
        double* gD = out.dN ? out.dN + (ptBase * NB + 32 * k) * 3 : nullptr;
        int buf = 0;
#pragma unroll 1
        for (int q2 = 0; q2 < Q; ++q2) {
          const double2 t2 = lds2(sT + 2 * C::TAB + (q2 * NB1 + i2) * 2);
          const double wn2 = w * t2.x, wd2 = w * t2.y;
#pragma unroll
          for (int q1 = 0; q1 < Q; ++q1) {
#pragma unroll
            for (int q0 = 0; q0 < Q; ++q0) {
              const int pt = q0 + Q * q1 + Q2 * q2;
              const double2 ca = lds2(sC + 4 * pt), cb = lds2(sC + 4 * pt + 2);  // invW, W0 | W1, W2
              const double t00 = n0[q0] * n1[q1];
              const double A0 = t00 * wn2;
              const double R = A0 * ca.x;
              if (gN && valid) gN[(size_t)pt * NB] = R;
              if (gD) {
                const double A1 = (d0[q0] * n1[q1]) * wn2;
                const double A2 = (n0[q0] * d1[q1]) * wn2;
                const double A3 = t00 * wd2;
                double* sd = sD + buf * 96;
                sd[3 * lane + 0] = (A1 - ca.y * R) * ca.x;
                sd[3 * lane + 1] = (A2 - cb.x * R) * ca.x;
                sd[3 * lane + 2] = (A3 - cb.y * R) * ca.x;
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
        __syncwarp();
      }
    }
    __syncwarp();
  }
}

constexpr int kCounterRing = 64;

}  // namespace

// K1 for 3-D patches (cached per handle: rebuilt only when the rule or the stream changes); table layout
// [e_d][q][i][a], a = 0 value, a = 1 d/dxi.  Shared by the evaluation kernel and the fused assembly kernel.
igab_status ensure_tables3d(const PatchDev& Pd, SfWorkspace& ws, int Q, const double* const* xi1d_dev,
                            const double* const* xi1d_host, cudaStream_t stream) {
  bool same = ws.valid && ws.stream == stream;
  for (int d = 0; d < 3 && same; ++d) {
    same = ws.nq1[d] == Q && (int)ws.xi[d].size() == Q;
    for (int q = 0; q < Q && same; ++q) same = ws.xi[d][q] == xi1d_host[d][q];
  }
  if (!ws.counter) {
    IGAB_CUDA_TRY(cudaMalloc(&ws.counter, sizeof(unsigned int) * kCounterRing));
  }
  if (same) return IGAB_OK;
  for (int d = 0; d < 3; ++d) {
    const size_t n = (size_t)Pd.nel[d] * Q * (Pd.degree[d] + 1) * 2;
    if (ws.tabCap[d] < n) {
      // stream-ordered: earlier launches on this stream may still read the old tables
      if (ws.tab[d]) IGAB_CUDA_TRY(cudaFreeAsync(ws.tab[d], stream));
      ws.tab[d] = nullptr;
      ws.tabCap[d] = 0;
      IGAB_CUDA_TRY(cudaMallocAsync(&ws.tab[d], n * sizeof(double), stream));
      ws.tabCap[d] = n;
    }
    const int threads = 128, total = Pd.nel[d] * Q;
    tables_kernel<<<(total + threads - 1) / threads, threads, 0, stream>>>(Pd, d, Q, xi1d_dev[d], ws.tab[d]);
    ws.nq1[d] = Q;
    ws.xi[d].assign(xi1d_host[d], xi1d_host[d] + Q);
  }
  count_launch(3);
  ws.valid = true;
  ws.wvalid = false;
  ws.stream = stream;
  return IGAB_OK;
}

igab_status ensure_wtables3d(const PatchDev& Pd, SfWorkspace& ws, int Q, cudaStream_t stream) {
  if (!Pd.wfac[0] || ws.wvalid) return IGAB_OK;
  for (int d = 0; d < 3; ++d) {
    const size_t n = (size_t)Pd.nel[d] * Q * (Pd.degree[d] + 1) * 2;
    if (ws.wtabCap[d] < n) {
      if (ws.wtab[d]) IGAB_CUDA_TRY(cudaFreeAsync(ws.wtab[d], stream));
      ws.wtab[d] = nullptr;
      ws.wtabCap[d] = 0;
      IGAB_CUDA_TRY(cudaMallocAsync(&ws.wtab[d], n * sizeof(double), stream));
      ws.wtabCap[d] = n;
    }
    const int threads = 128, total = (int)(n / 2);
    wtables_kernel<<<(total + threads - 1) / threads, threads, 0, stream>>>(Pd, d, Q, ws.tab[d], ws.wtab[d]);
  }
  count_launch(3);
  IGAB_CUDA_TRY(cudaGetLastError());
  ws.wvalid = true;
  return IGAB_OK;
}

namespace {

template <int P, int Q>
igab_status launch_cfg(const PatchDev& Pd, SfWorkspace& ws, const PointSource& src, const double* const* xi1d_host,
                       long long nelem, const EvalOutDev& out, cudaStream_t stream) {
  using C = Cfg<P, Q>;
  if (nelem > 0x7fffffffLL) {
    set_error("eval_tensor: more than 2^31 elements in one launch");
    return IGAB_INVALID_ARGUMENT;
  }
  igab_status st = ensure_tables3d(Pd, ws, Q, src.xi1d, xi1d_host, stream);
  if (st) return st;
  // K2: warps per block chosen so that at least three blocks fit the 227 KB of shared memory of an SM
  constexpr int WARPS = (C::SZ_WARP * 8 * 4 * 3 <= 225 * 1024) ? 4 : ((C::SZ_WARP * 8 * 2 * 3 <= 225 * 1024) ? 2 : 1);
  const size_t smem = (size_t)WARPS * C::SZ_WARP * sizeof(double);
  KernelCfg kc;
  if (igab_status cst = kernel_config((const void*)eval_sf3d_kernel<P, Q>, WARPS * 32, smem, &kc)) return cst;
  const int blocksPerSm = kc.blocksPerSm, numSms = kc.sms;
  int bps = blocksPerSm, tune = 0;
  if (const char* e = getenv("IGAB_SF_BLOCKS_PER_SM")) bps = std::max(1, std::min(blocksPerSm, atoi(e)));  // tuning knob
  if (const char* e = getenv("IGAB_SF_TUNE")) tune = atoi(e);
  long long blocks = (long long)numSms * bps;  // persistent: grid = SMs x resident blocks
  const long long need = (nelem + WARPS - 1) / WARPS;
  if (blocks > need) blocks = need;
  unsigned int* ctr = ws.counter + (ws.counterSlot++ % kCounterRing);
  IGAB_CUDA_TRY(cudaMemsetAsync(ctr, 0, sizeof(unsigned int), stream));
  eval_sf3d_kernel<P, Q><<<(unsigned)blocks, WARPS * 32, smem, stream>>>(Pd, src.elem_begin, nelem, ws.tab[0],
                                                                         ws.tab[1], ws.tab[2], out, tune, ctr);
  count_launch();
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

struct Key {
  int p, q;
};
bool uniform_cfg(const PatchDev& P, const int* nq1, Key* k) {
  if (P.dim != 3 || P.dw != 3) return false;
  if (P.degree[0] != P.degree[1] || P.degree[0] != P.degree[2]) return false;
  if (nq1[0] != nq1[1] || nq1[0] != nq1[2]) return false;
  k->p = P.degree[0];
  k->q = nq1[0];
  return true;
}

}  // namespace

bool eval_tensor_sf_supported(const PatchDev& P, const int* nq1, int order, const EvalOutDev& out) {
  Key k;
  if (!uniform_cfg(P, nq1, &k)) return false;
  if (out.d2N || out.H) return false;  // second derivatives: general kernel
  (void)order;
  return (k.p == 3 && (k.q == 4 || k.q == 5)) || (k.p == 2 && (k.q == 3 || k.q == 4)) || (k.p == 4 && k.q == 5);
}

void free_sf_workspace(SfWorkspace& ws) {
  for (int d = 0; d < kMaxDim; ++d)
    if (ws.tab[d]) cudaFree(ws.tab[d]);
  for (int d = 0; d < kMaxDim; ++d)
    if (ws.tabO2[d]) cudaFree(ws.tabO2[d]);
  for (int d = 0; d < kMaxDim; ++d)
    if (ws.wtab[d]) cudaFree(ws.wtab[d]);
  if (ws.counter) cudaFree(ws.counter);
  ws = SfWorkspace();
}

igab_status launch_eval_tensor_sf(const PatchDev& P, SfWorkspace& ws, const PointSource& src,
                                  const double* const* xi1d_host, long long nelem, int order, const EvalOutDev& out,
                                  cudaStream_t stream) {
  Key k;
  if (!uniform_cfg(P, src.nq1, &k) || !eval_tensor_sf_supported(P, src.nq1, order, out)) {
    set_error("eval_tensor: no sum-factorised kernel for this configuration");
    return IGAB_UNSUPPORTED;
  }
  EvalOutDev o = out;
  if (order < 1) o.dN = o.Jt = o.detJ = o.JinvT = nullptr;
  if (k.p == 3 && k.q == 4) return launch_cfg<3, 4>(P, ws, src, xi1d_host, nelem, o, stream);
  if (k.p == 3 && k.q == 5) return launch_cfg<3, 5>(P, ws, src, xi1d_host, nelem, o, stream);
  if (k.p == 2 && k.q == 3) return launch_cfg<2, 3>(P, ws, src, xi1d_host, nelem, o, stream);
  if (k.p == 2 && k.q == 4) return launch_cfg<2, 4>(P, ws, src, xi1d_host, nelem, o, stream);
  return launch_cfg<4, 5>(P, ws, src, xi1d_host, nelem, o, stream);
}

}  // namespace igab

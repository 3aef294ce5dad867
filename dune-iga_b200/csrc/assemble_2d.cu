// assemble_2d.cu -- fused element matrices for 2-D patches (plane problems and surfaces in R^3), p = 1, 2, 3 with p, p+1
// or p+2 Gauss points per direction: evaluation and contraction in one kernel, only K_e / f_e leave the SM.  For these small local bases
// (n = 4 .. 32) the contraction is FMA work, not tensor-core work (BASELINE north_star: DMMA only for p >= 3 in 3-D);
// the general B-stack path (assemble.cu) makes an HBM round trip of ~8x the size of K_e, this kernel makes none.
//
// Functional spec: localAssembler of dune/python/test/poisson.py:37-81 generalised to B^T D B (same as assemble_fused.cu):
//   Poisson     K[i][j] = sum_q w detJ grad R_i . grad R_j   (grad = J^T (J J^T)^-1 d/dxi: surfaces in R^3 included)
//   mass        K[i][j] = sum_q w detJ R_i R_j
//   elasticity  K[(i,a),(j,b)] = sum_{c,c'} Cf[a][c][b][c'] M^{cc'}[i][j],  M^{cc'} = sum_q w detJ G_c[i] G_c'[j],
//               Cf = S^T D S for the plane Voigt operator (xx, yy, xy), DOFs i*2 + a
//   f_e[i] = sum_q w detJ R_i fconst
//
// One warp owns EPW = 32 / Q^2 consecutive elements.
//  (1) lanes <-> (element, quadrature point): univariate rows from the cached K1 tables (eval_tensor_2d.cu), homogeneous
//      geometry sums, J, detJ = sqrtDetAAT, J^T (J J^T)^-1; the point's physical gradients of all (p+1)^2 functions,
//      pre-scaled by sqrt(w detJ), go to shared memory  G[element][point][i][c];
//  (2) lanes <-> rows (element, i): each lane accumulates one row of K_e over the points in registers; the column
//      operands are shared-memory broadcasts.  Rows are staged in shared memory;
//  (3) the EPW matrices are one contiguous run in global memory: coalesced flush.
#include <algorithm>

#include "bspline_device.cuh"
#include "igab_internal.cuh"

namespace igab {

namespace {

struct Elas2Coef {
  double cf[16];  // [a][c][b][c']
};

template <int DW, int P, int Q, int KIND>
struct A2Cfg {
  static constexpr int NB1 = P + 1, NB = NB1 * NB1, NQ = Q * Q;
  static constexpr int EPW = 32 / NQ;                          // elements per warp
  static constexpr int NR = KIND == IGAB_ASM_MASS ? 1 : DW;    // rows of the gradient stack per point
  static constexpr int NRP = NR == 3 ? 4 : NR;                 // padded (16-byte rows)
  static constexpr int NCOMP = KIND == IGAB_ASM_ELASTICITY ? 2 : 1;
  static constexpr int N = NB * NCOMP;
  // stride of a point's gradient block: odd for scalar rows, an odd number of 16-byte units otherwise, so that the
  // point-lanes' vector stores of phase (1) spread over all banks
  static constexpr int PG = NRP == 1 ? (NB | 1) : NB * NRP + (((NB * NRP / 2) % 2 == 0) ? 2 : 0);
  static constexpr int PR = NB | 1;                            // stride of a point's row of w detJ R_i (odd)
  static constexpr int KS = N | 1;                             // row stride of the staged matrices (odd)
  static constexpr int SZ_G = EPW * NQ * PG;
  static constexpr int SZ_R = EPW * NQ * PR;                   // w detJ R_i (load vector)
  static constexpr int SZ_K = EPW * N * KS;
  static constexpr int SZ_WARP = (SZ_G + SZ_R + SZ_K + 1) & ~1;
  static constexpr int WARPS = 4;
};

__device__ __forceinline__ double2 lds2a(const double* p) { return *reinterpret_cast<const double2*>(p); }

// One quadrature point: homogeneous geometry sums, J, detJ = sqrtDetAAT, J^T (J J^T)^-1 and the physical gradients of all
// (p+1)^2 functions pre-scaled by sqrt(w detJ) -> G[i][NRP] (mass: R_i sqrt(w detJ) -> G[i]); Rw[i] = w detJ R_i fconst.
// n0 / n1: the point's univariate rows (value, d/dxi scaled to the element).  wq: quadrature weight.
template <int DW, int P, int KIND>
__device__ __forceinline__ void a2_point(const PatchDev& Pd, long long cpBase, const double (&n0)[2][P + 1],
                                         const double (&n1)[2][P + 1], double wq, double fconst, double* __restrict__ G,
                                         double* __restrict__ Rw) {
  constexpr int NB1 = P + 1, NB = NB1 * NB1;
  constexpr int NR = KIND == IGAB_ASM_MASS ? 1 : DW, NRP = NR == 3 ? 4 : NR;
  const double* __restrict__ cp = Pd.cp;
  double Wa[3] = {0.0, 0.0, 0.0}, Ca[3][DW], wv[NB];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < DW; ++c) Ca[a][c] = 0.0;
#pragma unroll
  for (int i1 = 0; i1 < NB1; ++i1)
#pragma unroll
    for (int i0 = 0; i0 < NB1; ++i0) {
      const double* pc = cp + (cpBase + i0 + (long long)Pd.ncp[0] * i1) * (DW + 1);
      double Pc[DW + 1];
#pragma unroll
      for (int c = 0; c <= DW; ++c) Pc[c] = __ldg(pc + c);
      const double w = Pc[DW];
      wv[i0 + NB1 * i1] = w;
      const double A[3] = {n0[0][i0] * n1[0][i1] * w, n0[1][i0] * n1[0][i1] * w, n0[0][i0] * n1[1][i1] * w};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        Wa[a] += A[a];
#pragma unroll
        for (int c = 0; c < DW; ++c) Ca[a][c] = fma(A[a], Pc[c], Ca[a][c]);
      }
    }
  const double invW = 1.0 / Wa[0];
  double J[2][DW];
#pragma unroll
  for (int d = 0; d < 2; ++d)
#pragma unroll
    for (int c = 0; c < DW; ++c) J[d][c] = (Ca[1 + d][c] - Wa[1 + d] * (Ca[0][c] * invW)) * invW;
  double g00 = 0.0, g01 = 0.0, g11 = 0.0;
#pragma unroll
  for (int c = 0; c < DW; ++c) {
    g00 += J[0][c] * J[0][c];
    g01 += J[0][c] * J[1][c];
    g11 += J[1][c] * J[1][c];
  }
  double det;
  if constexpr (DW == 2)
    det = fabs(J[0][0] * J[1][1] - J[0][1] * J[1][0]);
  else
    det = sqrt(g00 * g11 - g01 * g01);
  const double wdet = wq * det, sq = sqrt(wdet);
  if constexpr (KIND == IGAB_ASM_MASS) {
    const double f = sq * invW;
#pragma unroll
    for (int i1 = 0; i1 < NB1; ++i1)
#pragma unroll
      for (int i0 = 0; i0 < NB1; ++i0) G[i0 + NB1 * i1] = n0[0][i0] * n1[0][i1] * wv[i0 + NB1 * i1] * f;
  } else {
    // G_c[i] = sum_d M[c][d] A_d[i] - v[c] Ahat[i],  M = s/W J^T (J J^T)^-1 [c][d],  v = (1/W) M dW/dxi
    const double idet = 1.0 / (g00 * g11 - g01 * g01), f = sq * invW * idet;
    double M[DW][2], v[DW];
#pragma unroll
    for (int c = 0; c < DW; ++c) {
      M[c][0] = (J[0][c] * g11 - J[1][c] * g01) * f;
      M[c][1] = (J[1][c] * g00 - J[0][c] * g01) * f;
      v[c] = (M[c][0] * Wa[1] + M[c][1] * Wa[2]) * invW;
    }
#pragma unroll
    for (int i1 = 0; i1 < NB1; ++i1)
#pragma unroll
      for (int i0 = 0; i0 < NB1; ++i0) {
        const int i = i0 + NB1 * i1;
        const double w = wv[i];
        const double Ah = n0[0][i0] * n1[0][i1] * w, A0 = n0[1][i0] * n1[0][i1] * w, A1 = n0[0][i0] * n1[1][i1] * w;
        double gc[DW];
#pragma unroll
        for (int c = 0; c < DW; ++c) gc[c] = M[c][0] * A0 + M[c][1] * A1 - v[c] * Ah;
        *reinterpret_cast<double2*>(G + i * NRP) = make_double2(gc[0], gc[1]);
        if constexpr (DW == 3) *reinterpret_cast<double2*>(G + i * NRP + 2) = make_double2(gc[2], 0.0);
      }
  }
  if (Rw) {
    const double f = wdet * invW * fconst;
#pragma unroll
    for (int i1 = 0; i1 < NB1; ++i1)
#pragma unroll
      for (int i0 = 0; i0 < NB1; ++i0) Rw[i0 + NB1 * i1] = n0[0][i0] * n1[0][i1] * wv[i0 + NB1 * i1] * f;
  }
}

template <int DW, int P, int Q, int KIND>
__global__ void __launch_bounds__(128) assemble2d_kernel(PatchDev Pd, long long elem_begin, long long nelem,
                                                         const double* __restrict__ tab0, const double* __restrict__ tab1,
                                                         const double* __restrict__ w0, const double* __restrict__ w1,
                                                         Elas2Coef coef, double fconst, double* __restrict__ Ke,
                                                         double* __restrict__ fe) {
  using C = A2Cfg<DW, P, Q, KIND>;
  constexpr int NB1 = C::NB1, NB = C::NB, NQ = C::NQ, EPW = C::EPW, NR = C::NR, NRP = C::NRP, N = C::N,
                NCOMP = C::NCOMP, PG = C::PG, PR = C::PR, KS = C::KS;
  extern __shared__ __align__(16) double a2smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* const sG = a2smem + (size_t)warp * C::SZ_WARP;
  double* const sR = sG + C::SZ_G;
  double* const sK = sR + C::SZ_R;
  const long long ngroups = (nelem + EPW - 1) / EPW;
  for (long long grp = (long long)blockIdx.x * C::WARPS + warp; grp < ngroups; grp += (long long)gridDim.x * C::WARPS) {
    const long long el0 = grp * EPW;                         // first element of the group (relative)
    const int nel = (int)min((long long)EPW, nelem - el0);   // elements in this group
    // ---- (1) lanes <-> (element, point)
    {
      const int es = lane / NQ, q = lane % NQ;
      const bool valid = es < nel;  // lanes >= EPW*NQ have es >= EPW >= nel
      if (valid) {
        const long long e = elem_begin + el0 + es;
        const int e0 = (int)(e % Pd.nel[0]), e1 = (int)(e / Pd.nel[0]);
        const int q0 = q % Q, q1 = q / Q;
        const int s0 = Pd.spanTab[0][e0], s1 = Pd.spanTab[1][e1];
        // univariate rows (value, d/dxi) of this point from the K1 tables [e_d][q][i][2]
        double n0[2][NB1], n1[2][NB1];
        {
          const double2* t0 = reinterpret_cast<const double2*>(tab0) + ((size_t)e0 * Q + q0) * NB1;
          const double2* t1 = reinterpret_cast<const double2*>(tab1) + ((size_t)e1 * Q + q1) * NB1;
#pragma unroll
          for (int i = 0; i < NB1; ++i) {
            const double2 a = __ldg(t0 + i), b = __ldg(t1 + i);
            n0[0][i] = a.x; n0[1][i] = a.y; n1[0][i] = b.x; n1[1][i] = b.y;
          }
        }
        const long long cpBase = (s0 - P) + (long long)Pd.ncp[0] * (s1 - P);
        a2_point<DW, P, KIND>(Pd, cpBase, n0, n1, __ldg(w0 + q0) * __ldg(w1 + q1), fconst,
                              sG + (size_t)(es * NQ + q) * PG, fe ? sR + (size_t)(es * NQ + q) * PR : nullptr);
      }
    }
    __syncwarp();
    // ---- (2) lanes <-> rows (element, i): the lane keeps row i of K_e in registers; the columns' gradients G[.][j] are
    //      the same address for all lanes of an element (shared-memory broadcast), the lane's own G[.][i] is read once
    //      per point
    for (int row = lane; row < nel * NB; row += 32) {  // more than 32 rows only for reduced rules (EPW * NB > 32)
      const int es = row / NB, i = row - es * NB;
      const double* Ge = sG + (size_t)es * NQ * PG;
      double* Krow = sK + (size_t)es * N * KS;
      if constexpr (KIND == IGAB_ASM_ELASTICITY) {
        double m[NB][4];  // M^{cc'}[i][j], [c][c']
#pragma unroll
        for (int j = 0; j < NB; ++j) m[j][0] = m[j][1] = m[j][2] = m[j][3] = 0.0;
#pragma unroll 1
        for (int q = 0; q < NQ; ++q) {
          const double2 a = lds2a(Ge + q * PG + i * NRP);
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            const double2 b = lds2a(Ge + q * PG + j * NRP);
            m[j][0] = fma(a.x, b.x, m[j][0]); m[j][1] = fma(a.x, b.y, m[j][1]);
            m[j][2] = fma(a.y, b.x, m[j][2]); m[j][3] = fma(a.y, b.y, m[j][3]);
          }
        }
#pragma unroll
        for (int j = 0; j < NB; ++j)
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
              double d = 0.0;
#pragma unroll
              for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int cp = 0; cp < 2; ++cp) d = fma(coef.cf[((a * 2 + c) * 2 + b) * 2 + cp], m[j][c * 2 + cp], d);
              Krow[(2 * i + a) * KS + 2 * j + b] = d;
            }
      } else {
        double acc[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[j] = 0.0;
#pragma unroll 1
        for (int q = 0; q < NQ; ++q) {
          const double* Gq = Ge + q * PG;
          if constexpr (NR == 1) {
            const double a = Gq[i];
#pragma unroll
            for (int j = 0; j < NB; ++j) acc[j] = fma(a, Gq[j], acc[j]);
          } else if constexpr (NR == 2) {
            const double2 a = lds2a(Gq + i * NRP);
#pragma unroll
            for (int j = 0; j < NB; ++j) {
              const double2 b = lds2a(Gq + j * NRP);
              acc[j] = fma(a.y, b.y, fma(a.x, b.x, acc[j]));
            }
          } else {
            const double2 a = lds2a(Gq + i * NRP), a2 = lds2a(Gq + i * NRP + 2);
#pragma unroll
            for (int j = 0; j < NB; ++j) {
              const double2 b = lds2a(Gq + j * NRP), b2 = lds2a(Gq + j * NRP + 2);
              acc[j] = fma(a2.x, b2.x, fma(a.y, b.y, fma(a.x, b.x, acc[j])));
            }
          }
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) Krow[i * KS + j] = acc[j];
      }
    }
    __syncwarp();
    // ---- (3) coalesced flush of the group's matrices (one contiguous run) and load vectors
    {
      double* g = Ke + el0 * (long long)(N * N);
      const int tot = nel * N * N;
      for (int k = lane; k < tot; k += 32) {
        const int r = k / N, c = k - r * N;  // r = element-major row index
        g[k] = sK[r * KS + c];
      }
      if (fe) {
        for (int s = lane; s < nel * NB; s += 32) {
          const int es = s / NB, i = s - es * NB;
          double acc = 0.0;
#pragma unroll
          for (int q = 0; q < NQ; ++q) acc += sR[(size_t)(es * NQ + q) * PR + i];
#pragma unroll
          for (int a = 0; a < NCOMP; ++a) fe[(el0 + es) * N + i * NCOMP + a] = acc;
        }
      }
    }
    __syncwarp();
  }
}

// ---- element matrices of elements that carry their own quadrature (trimmed elements: the rule of
// dune/iga/utils/fillquadraturerule.hh:8-19, variable number of points per element) --------------------------------------
// One warp per list element.  Its points are processed in chunks of 32 (lane <-> point, univariate rows by A2.3 in
// registers since arbitrary points share no table); then lanes <-> (row i, part s): part s of the S = 32 / nb parts
// accumulates the points q = s (mod S) of the chunk into row i of K_e (registers, kept across chunks); at the end the S
// partial rows are added through shared memory and the matrix leaves as one contiguous run.
template <int DW, int P, int KIND>
struct A2PCfg {
  static constexpr int NB1 = P + 1, NB = NB1 * NB1;
  static constexpr int NR = KIND == IGAB_ASM_MASS ? 1 : DW, NRP = NR == 3 ? 4 : NR;
  static constexpr int NCOMP = KIND == IGAB_ASM_ELASTICITY ? 2 : 1, N = NB * NCOMP;
  static constexpr int S = 32 / NB;  // parts
  static constexpr int PG = NRP == 1 ? (NB | 1) : NB * NRP + (((NB * NRP / 2) % 2 == 0) ? 2 : 0);
  static constexpr int PR = NB | 1;
  static constexpr int SZ_G = 32 * PG, SZ_R = 32 * PR;
  static constexpr int SZ_K = S * N * N;
  static constexpr int SZ_WARP = ((SZ_G + SZ_R > SZ_K ? SZ_G + SZ_R : SZ_K) + 1) & ~1;  // K staging aliases G / R
  static constexpr int WARPS = 4;
};

template <int DW, int P, int KIND>
__global__ void __launch_bounds__(128) assemble2d_points_kernel(PatchDev Pd, long long nlist, const int* __restrict__ elem,
                                                                const long long* __restrict__ offset,
                                                                const double* __restrict__ xi, const double* __restrict__ wq,
                                                                Elas2Coef coef, double fconst, double* __restrict__ Ke,
                                                                double* __restrict__ fe) {
  using C = A2PCfg<DW, P, KIND>;
  constexpr int NB1 = C::NB1, NB = C::NB, NRP = C::NRP, NR = C::NR, N = C::N, NCOMP = C::NCOMP, S = C::S, PG = C::PG, PR = C::PR;
  extern __shared__ __align__(16) double a2psmem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* const sG = a2psmem + (size_t)warp * C::SZ_WARP;
  double* const sR = sG + C::SZ_G;
  double* const sK = sG;  // after the last chunk
  for (long long k = (long long)blockIdx.x * C::WARPS + warp; k < nlist; k += (long long)gridDim.x * C::WARPS) {
    const long long e = elem[k];
    const long long p0 = offset[k], p1 = offset[k + 1];
    const int e0 = (int)(e % Pd.nel[0]), e1 = (int)(e / Pd.nel[0]);
    const int s0 = Pd.spanTab[0][e0], s1 = Pd.spanTab[1][e1];
    const double* U0 = Pd.knots[0];
    const double* U1 = Pd.knots[1];
    const double lo0 = U0[s0], sc0 = U0[s0 + 1] - lo0, lo1 = U1[s1], sc1 = U1[s1 + 1] - lo1;
    const long long cpBase = (s0 - P) + (long long)Pd.ncp[0] * (s1 - P);
    const int ri = lane % NB, part = lane / NB;  // row and part of this lane in phase (2); part >= S: idle
    constexpr int NACC = KIND == IGAB_ASM_ELASTICITY ? 4 * NB : NB;
    double acc[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) acc[j] = 0.0;
    double facc = 0.0;
    for (long long c0 = p0; c0 < p1; c0 += 32) {
      const int np = (int)min(32LL, p1 - c0);
      __syncwarp();
      if (lane < np) {
        const long long pt = c0 + lane;
        double n0[2][NB1], n1[2][NB1];
        constexpr int TRI = P * (P + 1) / 2;
        ders1d_reg_rd<P, 1>(xi[pt * 2] * sc0 + lo0, U0, s0, sc0, Pd.rden[0] + (size_t)e0 * TRI, n0);
        ders1d_reg_rd<P, 1>(xi[pt * 2 + 1] * sc1 + lo1, U1, s1, sc1, Pd.rden[1] + (size_t)e1 * TRI, n1);
        a2_point<DW, P, KIND>(Pd, cpBase, n0, n1, wq[pt], fconst, sG + (size_t)lane * PG, fe ? sR + (size_t)lane * PR : nullptr);
      }
      __syncwarp();
      if (part < S) {
        for (int q = part; q < np; q += S) {
          const double* Gq = sG + (size_t)q * PG;
          if constexpr (KIND == IGAB_ASM_ELASTICITY) {
            const double2 a = lds2a(Gq + ri * NRP);
#pragma unroll
            for (int j = 0; j < NB; ++j) {
              const double2 b = lds2a(Gq + j * NRP);
              acc[4 * j + 0] = fma(a.x, b.x, acc[4 * j + 0]); acc[4 * j + 1] = fma(a.x, b.y, acc[4 * j + 1]);
              acc[4 * j + 2] = fma(a.y, b.x, acc[4 * j + 2]); acc[4 * j + 3] = fma(a.y, b.y, acc[4 * j + 3]);
            }
          } else if constexpr (NR == 1) {
            const double a = Gq[ri];
#pragma unroll
            for (int j = 0; j < NB; ++j) acc[j] = fma(a, Gq[j], acc[j]);
          } else if constexpr (NR == 2) {
            const double2 a = lds2a(Gq + ri * NRP);
#pragma unroll
            for (int j = 0; j < NB; ++j) {
              const double2 b = lds2a(Gq + j * NRP);
              acc[j] = fma(a.y, b.y, fma(a.x, b.x, acc[j]));
            }
          } else {
            const double2 a = lds2a(Gq + ri * NRP), a2 = lds2a(Gq + ri * NRP + 2);
#pragma unroll
            for (int j = 0; j < NB; ++j) {
              const double2 b = lds2a(Gq + j * NRP), b2 = lds2a(Gq + j * NRP + 2);
              acc[j] = fma(a2.x, b2.x, fma(a.y, b.y, fma(a.x, b.x, acc[j])));
            }
          }
          if (fe) facc += sR[(size_t)q * PR + ri];
        }
      }
    }
    __syncwarp();
    // partial rows -> shared memory [part][row][col], then the sum over the parts leaves as one contiguous run
    if (part < S) {
      double* Kp = sK + (size_t)part * N * N;
      if constexpr (KIND == IGAB_ASM_ELASTICITY) {
#pragma unroll
        for (int j = 0; j < NB; ++j)
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
              double d = 0.0;
#pragma unroll
              for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int cp = 0; cp < 2; ++cp) d = fma(coef.cf[((a * 2 + c) * 2 + b) * 2 + cp], acc[4 * j + c * 2 + cp], d);
              Kp[(2 * ri + a) * N + 2 * j + b] = d;
            }
      } else {
#pragma unroll
        for (int j = 0; j < NB; ++j) Kp[ri * N + j] = acc[j];
      }
    }
    // load vector: sum of the parts by shuffles (lanes ri + NB * part)
    if (fe) {
      double f = part < S ? facc : 0.0;
      double tot = 0.0;
#pragma unroll
      for (int s = 0; s < S; ++s) tot += __shfl_sync(0xffffffffu, f, (ri + NB * s) & 31);
      if (lane < NB) {
#pragma unroll
        for (int a = 0; a < NCOMP; ++a) fe[k * N + lane * NCOMP + a] = tot;
      }
    }
    __syncwarp();
    double* g = Ke + k * (long long)(N * N);
    for (int idx = lane; idx < N * N; idx += 32) {
      double v = 0.0;
#pragma unroll
      for (int s = 0; s < S; ++s) v += sK[(size_t)s * N * N + idx];
      g[idx] = v;
    }
    __syncwarp();
  }
}

template <int DW, int P, int KIND>
igab_status launch_a2p(const PatchDev& Pd, const Elas2Coef& coef, double fconst, long long nlist, const int* elem,
                       const long long* offset, const double* xi, const double* w, double* Ke, double* fe,
                       cudaStream_t stream) {
  using C = A2PCfg<DW, P, KIND>;
  const size_t smem = (size_t)C::WARPS * C::SZ_WARP * sizeof(double);
  KernelCfg kc;
  if (igab_status cst = kernel_config((const void*)assemble2d_points_kernel<DW, P, KIND>, C::WARPS * 32, smem, &kc)) return cst;
  const long long need = (nlist + C::WARPS - 1) / C::WARPS;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(need, (long long)kc.sms * kc.blocksPerSm * 4));
  assemble2d_points_kernel<DW, P, KIND><<<grid, C::WARPS * 32, smem, stream>>>(Pd, nlist, elem, offset, xi, w, coef, fconst, Ke, fe);
  count_launch();
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

template <int DW, int KIND>
igab_status dispatch_a2p(int p, const PatchDev& Pd, const Elas2Coef& coef, double fconst, long long nlist, const int* elem,
                         const long long* offset, const double* xi, const double* w, double* Ke, double* fe,
                         cudaStream_t stream) {
  if (p == 1) return launch_a2p<DW, 1, KIND>(Pd, coef, fconst, nlist, elem, offset, xi, w, Ke, fe, stream);
  if (p == 2) return launch_a2p<DW, 2, KIND>(Pd, coef, fconst, nlist, elem, offset, xi, w, Ke, fe, stream);
  return launch_a2p<DW, 3, KIND>(Pd, coef, fconst, nlist, elem, offset, xi, w, Ke, fe, stream);
}

template <int DW, int P, int Q, int KIND>
igab_status launch_a2(const PatchDev& Pd, const Elas2Coef& coef, double fconst, long long eb, long long nel,
                      const double* const* tab, const double* const* w, double* Ke, double* fe, cudaStream_t stream) {
  using C = A2Cfg<DW, P, Q, KIND>;
  const size_t smem = (size_t)C::WARPS * C::SZ_WARP * sizeof(double);
  KernelCfg kc;
  if (igab_status cst = kernel_config((const void*)assemble2d_kernel<DW, P, Q, KIND>, C::WARPS * 32, smem, &kc)) return cst;
  const int sms = kc.sms, bps = kc.blocksPerSm;
  const long long groups = (nel + C::EPW - 1) / C::EPW;
  const long long need = (groups + C::WARPS - 1) / C::WARPS;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(need, (long long)sms * bps * 4));
  assemble2d_kernel<DW, P, Q, KIND><<<grid, C::WARPS * 32, smem, stream>>>(Pd, eb, nel, tab[0], tab[1], w[0], w[1], coef,
                                                                           fconst, Ke, fe);
  count_launch();
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

template <int DW, int KIND>
igab_status dispatch_a2(int p, int q, const PatchDev& Pd, const Elas2Coef& coef, double fconst, long long eb, long long nel,
                        const double* const* xi, const double* const* w, double* Ke, double* fe, cudaStream_t stream) {
#define IGAB_A2_CASE(PP, QQ) \
  if (p == PP && q == QQ) return launch_a2<DW, PP, QQ, KIND>(Pd, coef, fconst, eb, nel, xi, w, Ke, fe, stream);
  IGAB_A2_CASE(1, 2) IGAB_A2_CASE(1, 3)
  IGAB_A2_CASE(2, 2) IGAB_A2_CASE(2, 3) IGAB_A2_CASE(2, 4)
  IGAB_A2_CASE(3, 3) IGAB_A2_CASE(3, 4) IGAB_A2_CASE(3, 5)
  if constexpr (KIND != IGAB_ASM_ELASTICITY) {  // p = 4: 25 functions, one element per warp; scalar problems only
    IGAB_A2_CASE(4, 4) IGAB_A2_CASE(4, 5)
  }
#undef IGAB_A2_CASE
  set_error("assemble: no fused 2-D kernel for this (degree, rule)");
  return IGAB_UNSUPPORTED;
}

}  // namespace

bool assemble2d_supported(const PatchDev& P, int kind, const int* nq1) {
  if (P.dim != 2 || (P.dw != 2 && P.dw != 3)) return false;
  if (P.degree[0] != P.degree[1] || P.degree[0] < 1 || P.degree[0] > 4) return false;
  // rules with p, p+1 or p+2 points per direction (p+1 = full Gauss; dune/python/test/poisson.py:44 uses order 3 = 2 points)
  const int p = P.degree[0], q = nq1[0];
  if (nq1[1] != q || q < p || q > p + 2 || q < 2) return false;
  if (p == 4 && (q > 5 || kind == IGAB_ASM_ELASTICITY)) return false;  // q^2 <= 32 lanes; scalar problems
  if (kind == IGAB_ASM_ELASTICITY) return P.dw == 2;
  return kind == IGAB_ASM_POISSON || kind == IGAB_ASM_MASS;
}

igab_status launch_assemble2d(const PatchDev& P, SfWorkspace& ws, const double* const* xi1d_host, int kind,
                              const double* D_host, double fconst, long long eb, long long nel,
                              const int* nq1, const double* const* xi1d_dev_in, const double* const* w1d_dev, double* Ke,
                              double* fe, cudaStream_t stream) {
  const int nq = nq1[0];
  if (nel == 0) return IGAB_OK;
  igab_status tst = ensure_tables2d(P, ws, nq, 1, xi1d_dev_in, xi1d_host, stream);
  if (tst) return tst;
  const double* const xi1d_dev[2] = {ws.tab[0], ws.tab[1]};  // the kernels read the tables, not the abscissae
  Elas2Coef coef;
  for (double& v : coef.cf) v = 0.0;
  const int p = P.degree[0];
  if (kind == IGAB_ASM_ELASTICITY) {
    if (!D_host) {
      set_error("assemble: elasticity needs the 3x3 D matrix");
      return IGAB_INVALID_ARGUMENT;
    }
    // Cf[a][c][b][c'] = sum_{r,t} S[r][a][c] D[r][t] S[t][b][c'], plane Voigt selector (xx, yy, xy; engineering shear)
    static const int S[4][3] = {{0, 0, 0}, {1, 1, 1}, {2, 0, 1}, {2, 1, 0}};
    for (auto& s1 : S)
      for (auto& s2 : S) coef.cf[((s1[1] * 2 + s1[2]) * 2 + s2[1]) * 2 + s2[2]] += D_host[s1[0] * 3 + s2[0]];
    return dispatch_a2<2, IGAB_ASM_ELASTICITY>(p, nq, P, coef, fconst, eb, nel, xi1d_dev, w1d_dev, Ke, fe, stream);
  }
  if (kind == IGAB_ASM_MASS) {
    if (P.dw == 2) return dispatch_a2<2, IGAB_ASM_MASS>(p, nq, P, coef, fconst, eb, nel, xi1d_dev, w1d_dev, Ke, fe, stream);
    return dispatch_a2<3, IGAB_ASM_MASS>(p, nq, P, coef, fconst, eb, nel, xi1d_dev, w1d_dev, Ke, fe, stream);
  }
  if (P.dw == 2) return dispatch_a2<2, IGAB_ASM_POISSON>(p, nq, P, coef, fconst, eb, nel, xi1d_dev, w1d_dev, Ke, fe, stream);
  return dispatch_a2<3, IGAB_ASM_POISSON>(p, nq, P, coef, fconst, eb, nel, xi1d_dev, w1d_dev, Ke, fe, stream);
}

static bool elas2_coef(const double* D_host, Elas2Coef* coef) {
  for (double& v : coef->cf) v = 0.0;
  if (!D_host) return false;
  static const int S[4][3] = {{0, 0, 0}, {1, 1, 1}, {2, 0, 1}, {2, 1, 0}};
  for (auto& s1 : S)
    for (auto& s2 : S) coef->cf[((s1[1] * 2 + s1[2]) * 2 + s2[1]) * 2 + s2[2]] += D_host[s1[0] * 3 + s2[0]];
  return true;
}

bool assemble2d_points_supported(const PatchDev& P, int kind) {
  if (P.dim != 2 || (P.dw != 2 && P.dw != 3)) return false;
  if (P.degree[0] != P.degree[1] || P.degree[0] < 1 || P.degree[0] > 3) return false;
  if (kind == IGAB_ASM_ELASTICITY) return P.dw == 2;
  return kind == IGAB_ASM_POISSON || kind == IGAB_ASM_MASS;
}

// all pointers device
igab_status launch_assemble2d_points(const PatchDev& P, int kind, const double* D_host, double fconst, long long nlist,
                                     const int* elem, const long long* offset, const double* xi, const double* w,
                                     double* Ke, double* fe, cudaStream_t stream) {
  if (nlist == 0) return IGAB_OK;
  if (!assemble2d_points_supported(P, kind)) {
    set_error("assemble_points: 2-D patches with equal degrees 1..3 only (trimming is a 2-D feature of the reference)");
    return IGAB_UNSUPPORTED;
  }
  Elas2Coef coef;
  const bool haveD = elas2_coef(D_host, &coef);
  const int p = P.degree[0];
  if (kind == IGAB_ASM_ELASTICITY) {
    if (!haveD) {
      set_error("assemble_points: elasticity needs the 3x3 D matrix");
      return IGAB_INVALID_ARGUMENT;
    }
    return dispatch_a2p<2, IGAB_ASM_ELASTICITY>(p, P, coef, fconst, nlist, elem, offset, xi, w, Ke, fe, stream);
  }
  if (kind == IGAB_ASM_MASS) {
    if (P.dw == 2) return dispatch_a2p<2, IGAB_ASM_MASS>(p, P, coef, fconst, nlist, elem, offset, xi, w, Ke, fe, stream);
    return dispatch_a2p<3, IGAB_ASM_MASS>(p, P, coef, fconst, nlist, elem, offset, xi, w, Ke, fe, stream);
  }
  if (P.dw == 2) return dispatch_a2p<2, IGAB_ASM_POISSON>(p, P, coef, fconst, nlist, elem, offset, xi, w, Ke, fe, stream);
  return dispatch_a2p<3, IGAB_ASM_POISSON>(p, P, coef, fconst, nlist, elem, offset, xi, w, Ke, fe, stream);
}

}  // namespace igab

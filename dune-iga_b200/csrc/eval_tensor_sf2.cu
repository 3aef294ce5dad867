// eval_tensor_sf2.cu -- tuned 3-D evaluation WITH SECOND DERIVATIVES at tensor rules (p = 2, 3; Q = p + 1, p + 2).
//
// What it replaces per point: NurbsLocalBasis::evaluateFunction / evaluateJacobian / partial(|alpha| = 2)
// (dune/iga/nurbsbasis.hh:77-108), i.e. Nurbs<3>::basisFunctionDerivatives with k = 2 (dune/iga/nurbsalgorithms.hh:193-250,
// which allocates 27 nets per point although 10 are read), and NURBSGeometry::global / jacobianTransposed /
// secondDerivativeOfPosition / integrationElement / jacobianInverseTransposed (dune/iga/nurbsgeometry.hh:133-138,182-210,
// 257-292; H rows [uu, vv, ww, uv, uw, vw]).
//
// Design (one warp per element, dynamic element scheduling as eval_tensor_sf.cu):
//  K1' univariate tables with value, first and second derivative pre-scaled by the span length ([e][q][i][4], 32 B rows);
//  (A) tables and the (p+1)^3 control points in homogeneous form -> shared memory;
//  (B) lanes <-> quadrature points: the ten homogeneous sums A^(a) = sum_i N_i^(a) (wP_i, w_i), a in {0, e_d, 2e_d,
//      e_c+e_d}, with the point's three univariate rows in registers, then the quotient rule on the geometry ->
//      x, J^T, H, detJ, J^-T staged per 32 points and flushed as contiguous runs; 1/W and the nine W^(a) kept per point;
//  (C) lanes <-> basis functions (i = lane + 32 k): R_i and its 3 + 6 derivatives by the quotient rule
//      (nurbsalgorithms.hh:226-248); N rows leave as coalesced stores from registers, the dN [nb][3] and d2N [nb][6] rows
//      of a point are assembled in a ring of shared-memory slabs and leave through ONE async bulk store each
//      (cp.async.bulk, SASS UBLKCP) when rows are 16-byte multiples (p = 3), else as coalesced warp copies.
// HBM-write bound: 8 (10 nb + 31) B per point = 5,368 B at p = 3.
#include <algorithm>

#include "bspline_device.cuh"
#include "igab_internal.cuh"

namespace igab {

namespace {

__global__ void tables_o2_kernel(PatchDev P, int d, int nq, const double* __restrict__ xi, double* __restrict__ tab) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.nel[d] * nq) return;
  const int e = t / nq, q = t % nq;
  const int p = P.degree[d];
  const int sp = P.spanTab[d][e];
  const double* U = P.knots[d];
  const double lo = U[sp], scale = U[sp + 1] - lo;
  double dN[3 * kPM];
  ders1d_any(xi[q] * scale + lo, U, p, sp, 2, dN);
  double* o = tab + ((size_t)t * (p + 1)) * 4;
  for (int i = 0; i <= p; ++i) {
    o[4 * i] = dN[i];
    o[4 * i + 1] = dN[kPM + i] * scale;
    o[4 * i + 2] = dN[2 * kPM + i] * scale * scale;
    o[4 * i + 3] = 0.0;
  }
}

template <int P, int Q>
struct Cfg2 {
  static constexpr int NB1 = P + 1, NB = NB1 * NB1 * NB1, NQ = Q * Q * Q, Q2 = Q * Q;
  static constexpr int KS = (NB + 31) / 32, PS = (NQ + 31) / 32;
  static constexpr int TAB = Q * NB1 * 4;
  static constexpr bool SHARED01 = (32 % (NB1 * NB1) == 0);  // all slots of a lane share (i0, i1)
  static constexpr bool BULK = (NB * 3 * 8) % 16 == 0;       // dN / d2N rows are 16-byte multiples
  static constexpr int NBUF = 3;
  static constexpr int ev(int x) { return (x + 1) & ~1; }
  static constexpr int SZ_T = 3 * TAB;
  static constexpr int SZ_CP = 4 * NB;
  static constexpr int CST = 12;                              // per point: 1/W, W_d (3), W_dd (3), W_cd (3), pad
  static constexpr int SZ_C = CST * NQ;
  static constexpr int SLAB = ev(9 * NB);                     // dN row (3 nb) followed by d2N row (6 nb)
  static constexpr int SZ_ST = 32 * 19;                       // staging of one geometry field for 32 points (odd stride)
  static constexpr int SZ_D = (NBUF * SLAB) > SZ_ST ? (NBUF * SLAB) : SZ_ST;
  static constexpr int SZ_WARP = ev(SZ_T + SZ_CP + SZ_C + SZ_D);
};

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_st(void* gdst, const void* ssrc, unsigned bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_rd() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

template <int P, int Q>
__global__ void __launch_bounds__(128) eval_sf3d_o2_kernel(PatchDev Pd, long long elem_begin, long long nelem,
                                                           const double* __restrict__ tab0,
                                                           const double* __restrict__ tab1,
                                                           const double* __restrict__ tab2, EvalOutDev out,
                                                           unsigned int* __restrict__ counter) {
  using C = Cfg2<P, Q>;
  constexpr int NB1 = C::NB1, NB = C::NB, NQ = C::NQ, Q2 = C::Q2, KS = C::KS, CST = C::CST;
  extern __shared__ __align__(16) double smem2[];
  const int lane = threadIdx.x & 31;
  double* const sT = smem2 + (size_t)(threadIdx.x >> 5) * C::SZ_WARP;
  double* const sCP = sT + C::SZ_T;
  double* const sC = sCP + C::SZ_CP;
  double* const sD = sC + C::SZ_C;   // slabs (phase C) / geometry staging (phase B)

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
    // ---- (A) tables + control points (homogeneous) -> shared
    {
      const double* t0 = tab0 + (size_t)e0 * C::TAB;
      const double* t1 = tab1 + (size_t)e1 * C::TAB;
      const double* t2 = tab2 + (size_t)e2 * C::TAB;
      for (int j = lane; j < C::TAB; j += 32) {
        sT[j] = __ldg(t0 + j);
        sT[C::TAB + j] = __ldg(t1 + j);
        sT[2 * C::TAB + j] = __ldg(t2 + j);
      }
      const int b0 = Pd.spanTab[0][e0] - P, b1 = Pd.spanTab[1][e1] - P, b2 = Pd.spanTab[2][e2] - P;
      const long long n0 = Pd.ncp[0], n1 = Pd.ncp[1];
      const double2* cp2 = reinterpret_cast<const double2*>(Pd.cp);
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        const int i = lane + 32 * k;
        if (i < NB) {
          const int i0 = i % NB1, i12 = i / NB1, i1 = i12 % NB1, i2 = i12 / NB1;
          const long long g = (b0 + i0) + n0 * ((b1 + i1) + n1 * (long long)(b2 + i2));
          const double2 xy = __ldg(cp2 + 2 * g), zw = __ldg(cp2 + 2 * g + 1);
          double* h = sCP + 4 * i;
          h[0] = xy.x * zw.y; h[1] = xy.y * zw.y; h[2] = zw.x * zw.y; h[3] = zw.y;
        }
      }
    }
    __syncwarp();
    const long long ptBase = el * NQ;
    // ---- (B) geometry, lanes <-> points
#pragma unroll 1
    for (int m = 0; m < C::PS; ++m) {
      const int pt = lane + 32 * m;
      const int nvalid = min(32, NQ - 32 * m);
      const bool valid = pt < NQ;
      const int pv = valid ? pt : 0;
      const int q0 = pv % Q, q1 = (pv / Q) % Q, q2 = pv / Q2;
      // A[a][c]: a = 0 | e0 e1 e2 | 2e0 2e1 2e2 | e0+e1 e0+e2 e1+e2 ; c = wx wy wz w
      double A[10][4];
#pragma unroll
      for (int a = 0; a < 10; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) A[a][c] = 0.0;
      double ra[NB1][3], rb[NB1][3];
#pragma unroll
      for (int i = 0; i < NB1; ++i) {
        const double2 a01 = ld2(sT + (q0 * NB1 + i) * 4), a2 = ld2(sT + (q0 * NB1 + i) * 4 + 2);
        const double2 b01 = ld2(sT + C::TAB + (q1 * NB1 + i) * 4), b2 = ld2(sT + C::TAB + (q1 * NB1 + i) * 4 + 2);
        ra[i][0] = a01.x; ra[i][1] = a01.y; ra[i][2] = a2.x;
        rb[i][0] = b01.x; rb[i][1] = b01.y; rb[i][2] = b2.x;
      }
#pragma unroll 1
      for (int i2 = 0; i2 < NB1; ++i2) {
        const double2 c01 = ld2(sT + 2 * C::TAB + (q2 * NB1 + i2) * 4), c2v = ld2(sT + 2 * C::TAB + (q2 * NB1 + i2) * 4 + 2);
        const double c0 = c01.x, c1 = c01.y, c2 = c2v.x;
#pragma unroll
        for (int i1 = 0; i1 < NB1; ++i1) {
          const double bc00 = rb[i1][0] * c0, bc10 = rb[i1][1] * c0, bc01 = rb[i1][0] * c1, bc20 = rb[i1][2] * c0,
                       bc02 = rb[i1][0] * c2, bc11 = rb[i1][1] * c1;
#pragma unroll
          for (int i0 = 0; i0 < NB1; ++i0) {
            const double* h = sCP + 4 * (i0 + NB1 * (i1 + NB1 * i2));
            const double2 hxy = ld2(h), hzw = ld2(h + 2);
            const double H4[4] = {hxy.x, hxy.y, hzw.x, hzw.y};
            const double f[10] = {ra[i0][0] * bc00, ra[i0][1] * bc00, ra[i0][0] * bc10, ra[i0][0] * bc01,
                                  ra[i0][2] * bc00, ra[i0][0] * bc20, ra[i0][0] * bc02,
                                  ra[i0][1] * bc10, ra[i0][1] * bc01, ra[i0][0] * bc11};
#pragma unroll
            for (int a = 0; a < 10; ++a)
#pragma unroll
              for (int c = 0; c < 4; ++c) A[a][c] = fma(f[a], H4[c], A[a][c]);
          }
        }
      }
      // quotient rule on the geometry
      const double iW = 1.0 / A[0][3];
      double x[3], J[3][3], Hm[6][3];
#pragma unroll
      for (int c = 0; c < 3; ++c) x[c] = A[0][c] * iW;
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int c = 0; c < 3; ++c) J[d][c] = (A[1 + d][c] - A[1 + d][3] * x[c]) * iW;
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int c = 0; c < 3; ++c) Hm[d][c] = (A[4 + d][c] - 2.0 * A[1 + d][3] * J[d][c] - A[4 + d][3] * x[c]) * iW;
      {
        constexpr int M0[3] = {0, 0, 1}, M1[3] = {1, 2, 2};
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            Hm[3 + r][c] = (A[7 + r][c] - A[1 + M0[r]][3] * J[M1[r]][c] - A[1 + M1[r]][3] * J[M0[r]][c] - A[7 + r][3] * x[c]) * iW;
      }
      if (valid) {
        double* cst = sC + CST * pt;
        cst[0] = iW;
#pragma unroll
        for (int a = 1; a < 10; ++a) cst[a] = A[a][3];
      }
      const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                         J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
      const long long p0 = ptBase + 32 * m;
      double* sSt = sD;
      auto flush = [&](double* dst, int width) {
        __syncwarp();
        for (int j = lane; j < width * nvalid; j += 32) dst[j] = sSt[(j / width) * 19 + j % width];
        __syncwarp();
      };
      if (out.x) {
        if (valid) { sSt[19 * lane] = x[0]; sSt[19 * lane + 1] = x[1]; sSt[19 * lane + 2] = x[2]; }
        flush(out.x + p0 * 3, 3);
      }
      if (out.Jt) {
        if (valid) {
#pragma unroll
          for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int c = 0; c < 3; ++c) sSt[19 * lane + 3 * d + c] = J[d][c];
        }
        flush(out.Jt + p0 * 9, 9);
      }
      if (out.H) {
        if (valid) {
#pragma unroll
          for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) sSt[19 * lane + 3 * r + c] = Hm[r][c];
        }
        flush(out.H + p0 * 18, 18);
      }
      if (out.JinvT) {
        if (valid) {
          const double id = 1.0 / det;
          const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
                       c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
          const double c10 = J[0][2] * J[2][1] - J[0][1] * J[2][2], c11 = J[0][0] * J[2][2] - J[0][2] * J[2][0],
                       c12 = J[0][1] * J[2][0] - J[0][0] * J[2][1];
          const double c20 = J[0][1] * J[1][2] - J[0][2] * J[1][1], c21 = J[0][2] * J[1][0] - J[0][0] * J[1][2],
                       c22 = J[0][0] * J[1][1] - J[0][1] * J[1][0];
          double* s = sSt + 19 * lane;
          s[0] = c00 * id; s[1] = c10 * id; s[2] = c20 * id;
          s[3] = c01 * id; s[4] = c11 * id; s[5] = c21 * id;
          s[6] = c02 * id; s[7] = c12 * id; s[8] = c22 * id;
        }
        flush(out.JinvT + p0 * 9, 9);
      }
      if (out.detJ && valid) out.detJ[p0 + lane] = fabs(det);
    }
    __syncwarp();
    // ---- (C) basis functions, lanes <-> basis functions
    if (out.N || out.dN || out.d2N) {
      int i0s[KS], i1s[KS], i2s[KS];
      double w[KS];
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        const int i = min(lane + 32 * k, NB - 1);
        i0s[k] = i % NB1; i1s[k] = (i / NB1) % NB1; i2s[k] = i / (NB1 * NB1);
        w[k] = (lane + 32 * k < NB) ? sCP[4 * i + 3] : 0.0;
      }
      constexpr int KR = C::SHARED01 ? 1 : KS;  // distinct (i0, i1) rows per lane
      double n0[KR][Q], d0[KR][Q], s0[KR][Q];
#pragma unroll
      for (int k = 0; k < KR; ++k)
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const double* a = sT + (q * NB1 + i0s[k]) * 4;
          const double2 a01 = ld2(a), a2 = ld2(a + 2);
          n0[k][q] = a01.x; d0[k][q] = a01.y; s0[k][q] = a2.x;
        }
      double* gN = out.N ? out.N + ptBase * NB : nullptr;
      double* gD = out.dN ? out.dN + ptBase * NB * 3 : nullptr;
      double* gD2 = out.d2N ? out.d2N + ptBase * NB * 6 : nullptr;
      const bool anyD = gD || gD2;
      const bool bulk = C::BULK && anyD && (((reinterpret_cast<unsigned long long>(gD) | reinterpret_cast<unsigned long long>(gD2)) & 15ull) == 0);
      int slot = 0;
#pragma unroll 1
      for (int q2 = 0; q2 < Q; ++q2) {
        double wn2[KS], wd2[KS], ws2[KS];
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          const double* c = sT + 2 * C::TAB + (q2 * NB1 + i2s[k]) * 4;
          const double2 c01 = ld2(c), c2 = ld2(c + 2);
          wn2[k] = w[k] * c01.x; wd2[k] = w[k] * c01.y; ws2[k] = w[k] * c2.x;
        }
#pragma unroll 1
        for (int q1 = 0; q1 < Q; ++q1) {
          double n1[KR], d1[KR], s1[KR];  // direction-1 row of this q1 (re-read from shared memory: keeps registers low)
#pragma unroll
          for (int k = 0; k < KR; ++k) {
            const double* b = sT + C::TAB + (q1 * NB1 + i1s[k]) * 4;
            const double2 b01 = ld2(b), b2 = ld2(b + 2);
            n1[k] = b01.x; d1[k] = b01.y; s1[k] = b2.x;
          }
#pragma unroll
          for (int q0 = 0; q0 < Q; ++q0) {
            const int pt = q0 + Q * q1 + Q2 * q2;
            const double* cst = sC + CST * pt;
            const double2 k01 = ld2(cst), k23 = ld2(cst + 2), k45 = ld2(cst + 4), k67 = ld2(cst + 6), k89 = ld2(cst + 8);
            const double iW = k01.x, W0 = k01.y, W1 = k23.x, W2 = k23.y, W00 = k45.x, W11 = k45.y, W22 = k67.x,
                         W01 = k67.y, W02 = k89.x, W12 = k89.y;
            double* sd = sD + slot * C::SLAB;
            if (anyD) {
              if (bulk && lane == 0) bulk_wait_rd<C::NBUF - 1>();  // the slab used NBUF points ago has been read
              __syncwarp();
            }
#pragma unroll
            for (int k = 0; k < KS; ++k) {
              const int kr = C::SHARED01 ? 0 : k;
              const double t00 = n0[kr][q0] * n1[kr], t10 = d0[kr][q0] * n1[kr], t01 = n0[kr][q0] * d1[kr];
              const double t20 = s0[kr][q0] * n1[kr], t02 = n0[kr][q0] * s1[kr], t11 = d0[kr][q0] * d1[kr];
              const int i = lane + 32 * k;
              const double R = (t00 * wn2[k]) * iW;
              const double R0 = (t10 * wn2[k] - W0 * R) * iW;
              const double R1 = (t01 * wn2[k] - W1 * R) * iW;
              const double R2 = (t00 * wd2[k] - W2 * R) * iW;
              if (i < NB) {
                if (gN) gN[(size_t)pt * NB + i] = R;
                if (gD) {
                  double* o = sd + 3 * i;
                  o[0] = R0; o[1] = R1; o[2] = R2;
                }
                if (gD2) {
                  double* o = sd + 3 * NB + 6 * i;
                  o[0] = (t20 * wn2[k] - 2.0 * W0 * R0 - W00 * R) * iW;
                  o[1] = (t02 * wn2[k] - 2.0 * W1 * R1 - W11 * R) * iW;
                  o[2] = (t00 * ws2[k] - 2.0 * W2 * R2 - W22 * R) * iW;
                  o[3] = (t11 * wn2[k] - W0 * R1 - W1 * R0 - W01 * R) * iW;
                  o[4] = (t10 * wd2[k] - W0 * R2 - W2 * R0 - W02 * R) * iW;
                  o[5] = (t01 * wd2[k] - W1 * R2 - W2 * R1 - W12 * R) * iW;
                }
              }
            }
            if (anyD) {
              if (bulk) {
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                  if (gD) bulk_st(gD + (size_t)pt * NB * 3, sd, NB * 3 * 8);
                  if (gD2) bulk_st(gD2 + (size_t)pt * NB * 6, sd + 3 * NB, NB * 6 * 8);
                  bulk_commit_group();
                }
              } else {
                __syncwarp();
                if (gD) {
                  double* g = gD + (size_t)pt * NB * 3;
                  for (int j = lane; j < NB * 3; j += 32) g[j] = sd[j];
                }
                if (gD2) {
                  double* g = gD2 + (size_t)pt * NB * 6;
                  for (int j = lane; j < NB * 6; j += 32) g[j] = sd[3 * NB + j];
                }
              }
              slot = (slot + 1 == C::NBUF) ? 0 : slot + 1;
            }
          }
        }
      }
      if (bulk && lane == 0) bulk_wait_rd<0>();  // slabs alias the geometry staging of the next element
      __syncwarp();
    }
    __syncwarp();
  }
}

constexpr int kCounterRing2 = 64;

template <int P, int Q>
igab_status launch_o2(const PatchDev& Pd, SfWorkspace& ws, const PointSource& src, long long nelem, const EvalOutDev& out,
                      cudaStream_t stream) {
  using C = Cfg2<P, Q>;
  if (nelem > 0x7fffffffLL) {
    set_error("eval_tensor: more than 2^31 elements in one launch");
    return IGAB_INVALID_ARGUMENT;
  }
  constexpr int WARPS = (C::SZ_WARP * 8 * 4 * 2 <= 225 * 1024) ? 4 : 2;
  const size_t smem = (size_t)WARPS * C::SZ_WARP * sizeof(double);
  KernelCfg kc;
  if (igab_status cst = kernel_config((const void*)eval_sf3d_o2_kernel<P, Q>, WARPS * 32, smem, &kc)) return cst;
  const int blocksPerSm = kc.blocksPerSm, numSms = kc.sms;
  long long blocks = (long long)numSms * blocksPerSm;
  const long long need = (nelem + WARPS - 1) / WARPS;
  if (blocks > need) blocks = need;
  if (!ws.counter) IGAB_CUDA_TRY(cudaMalloc(&ws.counter, sizeof(unsigned int) * kCounterRing2));
  unsigned int* ctr = ws.counter + (ws.counterSlot++ % kCounterRing2);
  IGAB_CUDA_TRY(cudaMemsetAsync(ctr, 0, sizeof(unsigned int), stream));
  eval_sf3d_o2_kernel<P, Q><<<(unsigned)blocks, WARPS * 32, smem, stream>>>(Pd, src.elem_begin, nelem, ws.tabO2[0],
                                                                            ws.tabO2[1], ws.tabO2[2], out, ctr);
  count_launch();
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

igab_status ensure_tables_o2(const PatchDev& Pd, SfWorkspace& ws, int Q, const double* const* xi1d_dev,
                             const double* const* xi1d_host, cudaStream_t stream) {
  bool same = ws.validO2 && ws.streamO2 == stream;
  for (int d = 0; d < 3 && same; ++d) {
    same = (int)ws.xiO2[d].size() == Q;
    for (int q = 0; q < Q && same; ++q) same = ws.xiO2[d][q] == xi1d_host[d][q];
  }
  if (same) return IGAB_OK;
  for (int d = 0; d < 3; ++d) {
    const size_t n = (size_t)Pd.nel[d] * Q * (Pd.degree[d] + 1) * 4;
    if (ws.tabO2Cap[d] < n) {
      if (ws.tabO2[d]) IGAB_CUDA_TRY(cudaFreeAsync(ws.tabO2[d], stream));
      ws.tabO2[d] = nullptr;
      ws.tabO2Cap[d] = 0;
      IGAB_CUDA_TRY(cudaMallocAsync(&ws.tabO2[d], n * sizeof(double), stream));
      ws.tabO2Cap[d] = n;
    }
    const int threads = 128, total = Pd.nel[d] * Q;
    tables_o2_kernel<<<(total + threads - 1) / threads, threads, 0, stream>>>(Pd, d, Q, xi1d_dev[d], ws.tabO2[d]);
    ws.xiO2[d].assign(xi1d_host[d], xi1d_host[d] + Q);
  }
  count_launch(3);
  ws.validO2 = true;
  ws.streamO2 = stream;
  return IGAB_OK;
}

bool cfg_of(const PatchDev& P, const int* nq1, int* p, int* q) {
  if (P.dim != 3 || P.dw != 3) return false;
  if (P.degree[0] != P.degree[1] || P.degree[0] != P.degree[2]) return false;
  if (nq1[0] != nq1[1] || nq1[0] != nq1[2]) return false;
  *p = P.degree[0];
  *q = nq1[0];
  return (*p == 3 && (*q == 4 || *q == 5)) || (*p == 2 && (*q == 3 || *q == 4));
}

}  // namespace

bool eval_tensor_sf2_supported(const PatchDev& P, const int* nq1, int order, const EvalOutDev& out) {
  int p, q;
  return order >= 2 && (out.d2N || out.H) && cfg_of(P, nq1, &p, &q);
}

igab_status launch_eval_tensor_sf2(const PatchDev& P, SfWorkspace& ws, const PointSource& src,
                                   const double* const* xi1d_host, long long nelem, const EvalOutDev& out,
                                   cudaStream_t stream) {
  int p, q;
  if (!cfg_of(P, src.nq1, &p, &q)) {
    set_error("eval_tensor: no second-derivative tensor kernel for this configuration");
    return IGAB_UNSUPPORTED;
  }
  igab_status st = ensure_tables_o2(P, ws, q, src.xi1d, xi1d_host, stream);
  if (st) return st;
  if (p == 3 && q == 4) return launch_o2<3, 4>(P, ws, src, nelem, out, stream);
  if (p == 3 && q == 5) return launch_o2<3, 5>(P, ws, src, nelem, out, stream);
  if (p == 2 && q == 3) return launch_o2<2, 3>(P, ws, src, nelem, out, stream);
  return launch_o2<2, 4>(P, ws, src, nelem, out, stream);
}

}  // namespace igab

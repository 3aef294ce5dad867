// gram_common.cuh -- pieces shared by the fused element-matrix kernels (assemble_fused.cu, assemble_sym.cu): the FP64
// tensor-core instruction, the shared-memory configuration of the fused front end, element_prepare (tables + control
// points + sum-factorised geometry + per-point constants) and the group layout of the sum-factorised contraction.
#pragma once
#include <utility>

#include "igab_internal.cuh"
#include "sf_geometry.cuh"

namespace igab {
namespace {

// Optional phase accounting (tools/phase_timing.py builds a separate library with -DIGAB_PHASE_TIMING): clock64 deltas of
// lane 0 of every warp, summed per phase.  Compiled out of the product library.
#ifdef IGAB_PHASE_TIMING
__device__ unsigned long long g_phase_ticks[16];
#define IGAB_PT_BEGIN long long pt_last = clock64();
#define IGAB_PT(k)                                                                                   \
  {                                                                                                  \
    const long long pt_now = clock64();                                                              \
    if ((threadIdx.x & 31) == 0) atomicAdd(&g_phase_ticks[k], (unsigned long long)(pt_now - pt_last)); \
    pt_last = pt_now;                                                                                \
  }
#define IGAB_WPT_BEGIN long long wpt_last = clock64();
#define IGAB_WPT(k)                                                                                    \
  {                                                                                                    \
    const long long wpt_now = clock64();                                                               \
    if ((threadIdx.x & 127) == 0) atomicAdd(&g_phase_ticks[k], (unsigned long long)(wpt_now - wpt_last)); \
    wpt_last = wpt_now;                                                                                \
  }
#else
#define IGAB_PT_BEGIN
#define IGAB_PT(k)
#define IGAB_WPT_BEGIN
#define IGAB_WPT(k)
#endif

__device__ __forceinline__ void dmma8x8x4_f(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 lds2f(const double* p) { return *reinterpret_cast<const double2*>(p); }

template <int P, int Q>
struct FCfg {
  static constexpr int NB1 = P + 1, NB = NB1 * NB1 * NB1, NQ = Q * Q * Q, Q2 = Q * Q;
  static constexpr int NBP = (NB + 31) / 32 * 32;
  static constexpr int NBLK = NBP / 32;            // 32-wide blocks (elasticity kernel)
  static constexpr int NTR = NBP / 8, NW = NTR / 2;  // 8-wide tile rows; warp w owns tile rows w and NTR-1-w
  static constexpr int NT = NW * 32;
  static constexpr int TAB = Q * NB1 * 2;
  static constexpr int HS = 4 * NB1 + 1, S1C = NB1 * NB1 + 1, S2S = NB1 | 1;
  static constexpr int ev(int x) { return (x + 1) & ~1; }
  static constexpr int SZ_T = 3 * TAB;
  static constexpr int SZ_W = ev(NBP);
  static constexpr int SZ_H = ev(NB1 * NB1 * HS);
  static constexpr int SZ_S1 = 2 * Q * 4 * S1C;
  static constexpr int SZ_G = 16 * NQ;
  static constexpr int SZ_A = ev((SZ_H + SZ_S1) > SZ_G ? (SZ_H + SZ_S1) : SZ_G);
  static constexpr int SZ_S2 = ev(3 * 4 * Q2 * S2S);
  static constexpr int CS = 16;  // doubles per point: 1/W, v[3], M[3][3], w detJ, s = sqrt(w detJ), s/W (element_prepare)
  static constexpr int SZ_C = CS * NQ;
  static constexpr int KC = 24, LD = NBP + 4;  // k-rows per chunk: 8 points x 3 gradient components (24 points for mass)
  static constexpr int SZ_SA = 2 * KC * LD;
  static constexpr int SZ_TOTAL = SZ_T + SZ_W + SZ_A + SZ_S2 + SZ_C + SZ_SA;
};

// Steps (1)+(2) for one element, executed by the whole block (NT threads): tables + control points -> shared, geometry
// of all quadrature points by sum factorisation, per-point constants sC[pt][16] (layout below).  Ends with a
// __syncthreads().
template <int P, int Q, int NT, int CS = FCfg<P, Q>::CS>
__device__ __forceinline__ void element_prepare(const PatchDev& Pd, long long e, const double* __restrict__ tab0,
                                                const double* __restrict__ tab1, const double* __restrict__ tab2,
                                                const double* __restrict__ w0, const double* __restrict__ w1,
                                                const double* __restrict__ w2, double* sT, double* sW, double* rA,
                                                double* sS2, double* sC) {
  using C = FCfg<P, Q>;
  constexpr int NB1 = C::NB1, NB = C::NB, NQ = C::NQ, Q2 = C::Q2;
  double* const sH = rA;
  double* const sS1 = rA + C::SZ_H;
  double* const sG = rA;
  const int tid = threadIdx.x;
  int e0, e1, e2;
  {
    long long h = e;
    e0 = (int)(h % Pd.nel[0]); h /= Pd.nel[0];
    e1 = (int)(h % Pd.nel[1]); h /= Pd.nel[1];
    e2 = (int)h;
  }
  __syncthreads();  // previous element done with every buffer
  sf_load<P, Q, NT>(Pd, tid, e0, e1, e2, tab0, tab1, tab2, sT, sW, rA);
  sf_contract<P, Q, NT>(tid, sT, rA, sS2);
  // per point: x = C/W, J_d = (C_d - W_d x)/W, detJ, J^-T (cofactors / det), w detJ
  for (int pt = tid; pt < NQ; pt += NT) {
    const double W = sG[3 * NQ + pt], invW = 1.0 / W;
    double x[3], J[3][3], Wd[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) x[c] = sG[c * NQ + pt] * invW;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      Wd[d] = sG[((d + 1) * 4 + 3) * NQ + pt];
#pragma unroll
      for (int c = 0; c < 3; ++c) J[d][c] = (sG[((d + 1) * 4 + c) * NQ + pt] - Wd[d] * x[c]) * invW;
    }
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
                 c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    const double c10 = J[0][2] * J[2][1] - J[0][1] * J[2][2], c11 = J[0][0] * J[2][2] - J[0][2] * J[2][0],
                 c12 = J[0][1] * J[2][0] - J[0][0] * J[2][1];
    const double c20 = J[0][1] * J[1][2] - J[0][2] * J[1][1], c21 = J[0][2] * J[1][0] - J[0][0] * J[1][2],
                 c22 = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02, id = 1.0 / det;
    int q = pt;
    double w = w0[q % Q];
    q /= Q;
    w *= w1[q % Q];
    w *= w2[q / Q];
    double* o = sC + CS * pt;
    // per point: physical gradient of R_i (pre-scaled by s = sqrt(w detJ), so that K = A^T A needs no scaling in the MMA
    // loop) as G_c[i] = sum_d M[c][d] A_d[i] - v[c] Ahat[i] with Ahat = w_i Nhat_i, A_d = w_i d_d Nhat_i and
    //   M[c][d] = s/W J^-T[c][d],  v[c] = (1/W) sum_d M[c][d] dW/dxi_d        (quotient rule folded into the constants)
    const double wdet = w * fabs(det), sq = sqrt(wdet), f = sq * invW * id;
    const double M[3][3] = {{c00 * f, c10 * f, c20 * f}, {c01 * f, c11 * f, c21 * f}, {c02 * f, c12 * f, c22 * f}};
    o[0] = invW;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o[1 + c] = (M[c][0] * Wd[0] + M[c][1] * Wd[1] + M[c][2] * Wd[2]) * invW;
#pragma unroll
      for (int d = 0; d < 3; ++d) o[4 + 3 * c + d] = M[c][d];
    }
    o[13] = wdet;
    o[14] = sq;
    o[15] = sq * invW;
  }
  __syncthreads();

}


// Weighted univariate tables (tensor-product weights, PatchDev::wfac): wtab_d[e][q][i][a] = w_d(i) * tab_d[e][q][i][a].
// With them the contraction runs on w_i Phi_i directly and the stored entries need no w_i w_j scaling; null otherwise.
struct WTabs {
  const double *t0, *t1, *t2;
};

template <int P, int Q, int KIND, int NW = 8>
struct SCfg {
  static constexpr bool MASS = KIND == 1, ELAS = KIND == 2;
  using F = FCfg<P, Q>;
  static constexpr int NB1 = P + 1, NP = NB1 * NB1, NB = F::NB, NQ = F::NQ, NWARP = NW, NT = 32 * NW;
  static constexpr int r4(int x) { return (x + 3) & ~3; }
  static constexpr int r8(int x) { return (x + 7) & ~7; }
  static constexpr int NG = MASS ? 1 : 4;
  static constexpr int NCOMBO = MASS ? 1 : 16;
  static constexpr int NU = MASS ? 1 : (ELAS ? 16 : 10);  // stored entries of M4 (symmetric for Poisson)
  static constexpr int ncombo(int g) { return MASS ? 1 : (g == 0 ? 9 : (g == 3 ? 1 : 3)); }
  static constexpr int goff(int g) {
    int o = 0;
    for (int h = 0; h < g; ++h) o += r4(ncombo(h) * Q);
    return o;
  }
  static constexpr int KTOT = goff(NG);
  static constexpr int M2 = NB1 * Q, M2P = r8(M2);      // stage-2 rows (j2, q0)
  static constexpr int NPP = r8(NP);                    // pair index, padded
  static constexpr int MT2 = M2P / 8, NT2 = NPP / 8;
  static constexpr int TPW = (MT2 * NT2 + NWARP - 1) / NWARP;  // stage-2 tiles per warp (same tile row)
  static constexpr int K3 = r4(NG * Q);
  static constexpr int M3 = NB1 * NP, M3P = r8(M3);     // stage-3 rows (j1, j2, i1)
  static constexpr int MT3 = M3P / 8, NT3 = NPP / 8;
  // leading dimensions: A operands = 4 mod 8 doubles, B operand = 8 mod 16 (conflict-free fragment loads)
  static constexpr int LDA2 = (KTOT + 3) / 8 * 8 + 4;
  // B operand: lane (lr, lc) reads [4 ks + lc][8 nt + lr]; a 64-bit shared load is served per half-warp (lr 0..3, lc 0..3), so
  // the four rows must start 8 banks apart: LDB2 = 4 mod 16 doubles (8 mod 16 put rows lc and lc + 2 on the same banks: half
  // of the stage-2 B loads were bank conflicts, profiles/README.md)
  static constexpr int LDB2 = (NPP + 11) / 16 * 16 + 4;
  // per-point constants of element_prepare with an ODD stride: a thread per point reads them (M4, load vector), and the
  // even stride 16 put every lane of a half-warp on the same banks
  static constexpr int CSS = 17, SZ_CS = F::ev(CSS * NQ);
  static constexpr int LDA3 = (K3 + 3) / 8 * 8 + 4;
  static constexpr int SZ_T1 = M2P * LDA2;
  static constexpr int SZ_X = F::ev((F::SZ_A + F::SZ_S2) > SZ_T1 ? (F::SZ_A + F::SZ_S2) : SZ_T1);
  static constexpr int SZ_M = F::ev(NU * NQ);
  static constexpr int SZ_T2 = M3P * LDA3;
  static constexpr int SZ_B2 = KTOT * LDB2;
  static constexpr int SZ_TOTAL = F::SZ_T + F::SZ_W + SZ_CS + SZ_X + SZ_M + SZ_T2 + SZ_B2 + F::SZ_T;  // + weighted tables
  // tiles of a warp share a tile row (one A fragment for TPW tiles) when that divides evenly; otherwise (GEN2) every
  // warp walks single tiles round-robin
  static constexpr bool GEN2 = !(TPW == 1 || NT2 % TPW == 0);
  static_assert(TPW * NWARP >= MT2 * NT2, "stage-2 tile distribution");
  // Stage-2 column order.  Column position c of the B operand / the accumulator tile stands for the index pair
  // n = cperm(c) = (j1, i1): within a group of eight, positions 2 lc + h hold the pairs lc + 4 h, so that the four lanes of
  // a quad store CONSECUTIVE pairs -> consecutive rows of T2 (the natural order 2 lc + h put them two rows apart: 2.8x
  // the ideal number of shared-memory wavefronts for the T2 stores, 1.4x now).
  __host__ __device__ static constexpr int cperm(int c) { return 8 * (c / 8) + ((c % 8) / 2) + 4 * ((c % 8) % 2); }
  __host__ __device__ static constexpr int cpos(int n) { return 8 * (n / 8) + 2 * ((n % 8) % 4) + ((n % 8) / 4); }
  // combo index (group order) -> (beta, gamma)
  __host__ __device__ static constexpr int set0(int k) { return k == 0 ? 0 : k + 1; }  // {0, 2, 3}
  __host__ __device__ static constexpr int beta_of(int g, int c) { return MASS ? 0 : (g == 0 ? set0(c / 3) : (g == 2 ? set0(c) : 1)); }
  __host__ __device__ static constexpr int gamma_of(int g, int c) { return MASS ? 0 : (g == 0 ? set0(c % 3) : (g == 1 ? set0(c) : 1)); }
  __host__ __device__ static constexpr int usym(int b, int c) {
    if (ELAS) return b * 4 + c;
    const int lo = b < c ? b : c, hi = b < c ? c : b;
    return lo * 4 - lo * (lo - 1) / 2 + (hi - lo);
  }
};


// compile-time loop: f(std::integral_constant<int, 0>{}), ..., f(std::integral_constant<int, N - 1>{}) -- the loop index is a
// constant EXPRESSION inside f (an unrolled run-time loop leaves the constexpr table look-ups below to the optimiser,
// which kept them as code: 12 k warp instructions per element)
template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) {
  (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  static_for_impl(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
}

// Stage 1 of the sum-factorised contraction on the tensor cores (gram_sym_kernel, gram_sf_kernel, gram_ws_kernel):
//   T1[(slot, q0)][(combo, q1)] = sum_q2 Phi2[q2][i2][b2(beta)] Phi2[q2][j2][b2(gamma)] M4_combo[q2][(q1, q0)],   b2(.) = (. == 3)
// per combo (beta, gamma) a small product [slot] x [q2] x [(q1, q0)] whose A operand exists in four variants
// v = b2(beta) + 2 b2(gamma); a slot is a pair (i2, j2) -- all j2 of one i2 in the kernels that contract every block.
// Eight warps: warp <-> column tile nt of (q1, q0) and a share wg of the combos; lane (lr, lc) holds
// C[slot = lr][n = 8 nt + 2 lc + h], n = q1 Q + q0.  Needs the number of column tiles to divide 8 (Q = 4, 5).
template <int P, int Q, int KIND>
struct S1Cfg {
  using S = SCfg<P, Q, KIND>;
  static constexpr bool MASS = KIND == 1;
  static constexpr int NQ = S::NQ;
  static constexpr int g_of(int cc) { return MASS ? 0 : (cc >= 15 ? 3 : (cc >= 12 ? 2 : (cc >= 9 ? 1 : 0))); }
  static constexpr int c_of(int cc) { return MASS ? 0 : (cc >= 15 ? 0 : (cc >= 12 ? cc - 12 : (cc >= 9 ? cc - 9 : cc))); }
  static constexpr int var_of(int cc) {
    return (S::beta_of(g_of(cc), c_of(cc)) == 3 ? 1 : 0) + (S::gamma_of(g_of(cc), c_of(cc)) == 3 ? 2 : 0);
  }
  static constexpr int var_count(int v) {
    int n = 0;
    for (int cc = 0; cc < S::NCOMBO; ++cc) n += var_of(cc) == v ? 1 : 0;
    return n;
  }
  static constexpr int var_item(int v, int k) {  // k-th combo of variant v, -1 past the end
    for (int cc = 0; cc < S::NCOMBO; ++cc)
      if (var_of(cc) == v && k-- == 0) return cc;
    return -1;
  }
  static constexpr int moff_of(int cc) { return cc < 0 ? 0 : S::usym(S::beta_of(g_of(cc), c_of(cc)), S::gamma_of(g_of(cc), c_of(cc))) * NQ; }
  static constexpr int coff_of(int cc) { return cc < 0 ? 0 : S::goff(g_of(cc)) + c_of(cc) * Q; }
  static constexpr int NW = 8;                   // warps that share the stage
  static constexpr int NTQ = S::r8(Q * Q) / 8;   // 8-column tiles over (q1, q0)
  static constexpr bool OK = NW % NTQ == 0;
  static constexpr int WG = OK ? NW / NTQ : 1;
  // the combos of a variant are dealt to the WG warp groups round-robin, starting where the previous variant stopped:
  // 9 + 3 + 3 + 1 combos over two groups are 8 + 8 (started at group 0 every time they were 10 + 6 rounds)
  static constexpr int var_first_group(int v) {
    int n = 0;
    for (int u = 0; u < v; ++u) n += var_count(u);
    return n % WG;
  }
  // gw = index of the warp among the eight; rows lr >= nslot_max are never stored
  __device__ static __forceinline__ void store_offsets(int gw, int lane, int nslot_max, int lda2, int (&s1off)[2]) {
    const int lr = lane >> 2, lc = lane & 3, nt = gw % NTQ;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = 8 * nt + 2 * lc + h;
      s1off[h] = (n < Q * Q && lr < nslot_max) ? (lr * Q + n % Q) * lda2 + n / Q : -1;
    }
  }
};

// (i2, j2) = the pair of this lane's slot lr, sv = slot in use; sTc = the (weighted) tables, sM = point matrices [u][q2][q1][q0]
template <int P, int Q, int KIND>
__device__ __forceinline__ void stage1_dmma(int gw, int lane, int i2, int j2, bool sv, const int (&s1off)[2],
                                            const double* sTc, const double* sM, double* sT1) {
  using Y = S1Cfg<P, Q, KIND>;
  constexpr bool MASS = Y::MASS;
  constexpr int NB1 = P + 1, TAB = FCfg<P, Q>::TAB, WG1 = Y::WG, KS1 = (Q + 3) / 4;
  const int lr = lane >> 2, lc = lane & 3;
  const int nt1 = gw % Y::NTQ, wg1 = gw / Y::NTQ;
  // the table entries every variant is made of, one round of loads: tv[ks][0..3] = Phi2[q2][i2][0 / 1], Phi2[q2][j2][0 / 1]
  double tv[KS1][4];
#pragma unroll
  for (int ks = 0; ks < KS1; ++ks) {
    const int q2 = min(4 * ks + lc, Q - 1);
    const bool on = sv && 4 * ks + lc < Q;
    const double2 ti = lds2f(sTc + 2 * TAB + (q2 * NB1 + i2) * 2), tj = lds2f(sTc + 2 * TAB + (q2 * NB1 + j2) * 2);
    tv[ks][0] = on ? ti.x : 0.0;
    tv[ks][1] = on ? ti.y : 0.0;
    tv[ks][2] = tj.x;
    tv[ks][3] = tj.y;
  }
  static_for<(MASS ? 1 : 4)>([&](auto vc) {
    constexpr int v = decltype(vc)::value;
    constexpr int NIT = (Y::var_count(v) + WG1 - 1) / WG1;
    double a[KS1];
#pragma unroll
    for (int ks = 0; ks < KS1; ++ks) a[ks] = tv[ks][v & 1] * tv[ks][2 + (v >> 1)];
    // all rounds' products first, the T1 stores afterwards: a shared store between two rounds would pin the next
    // round's operand loads behind it (the compiler cannot tell T1 from the point matrices)
    double cacc[NIT][2];
    int coffs[NIT];
    static_for<NIT>([&](auto itc) {
      constexpr int it = decltype(itc)::value;
      int moff = -1, coff = 0;
      static_for<WG1>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        // group k takes item (k - first group of the variant) mod WG of this round
        constexpr int cc = Y::var_item(v, it * WG1 + (k - Y::var_first_group(v) + WG1) % WG1);
        if constexpr (cc >= 0) {
          constexpr int mo = Y::moff_of(cc), co = Y::coff_of(cc);
          if (wg1 == k) {
            moff = mo;
            coff = co;
          }
        }
      });
      coffs[it] = moff >= 0 ? coff : -1;
      cacc[it][0] = cacc[it][1] = 0.0;
      if (moff >= 0) {
        const double* brow = sM + moff + 8 * nt1 + lr;
#pragma unroll
        for (int ks = 0; ks < KS1; ++ks) dmma8x8x4_f(cacc[it][0], cacc[it][1], a[ks], brow[Q * Q * min(4 * ks + lc, Q - 1)]);
      }
    });
    static_for<NIT>([&](auto itc) {
      constexpr int it = decltype(itc)::value;
      if (coffs[it] >= 0 && sv) {
        if (s1off[0] >= 0) sT1[s1off[0] + coffs[it]] = cacc[it][0];
        if (s1off[1] >= 0) sT1[s1off[1] + coffs[it]] = cacc[it][1];
      }
    });
  });
}

inline WTabs wtabs_of(const PatchDev& Pd, const SfWorkspace& ws) {
  if (Pd.wfac[0] && ws.wvalid && ws.wuse) return WTabs{ws.wtab[0], ws.wtab[1], ws.wtab[2]};
  return WTabs{nullptr, nullptr, nullptr};
}

}  // namespace
}  // namespace igab

// assemble_fused.cu -- fully fused element-matrix kernel for 3-D Poisson / mass at p = 2, 3, 4 (BASELINE config C3: p = 4):
// evaluation AND contraction in one kernel, nothing but K_e (and f_e) leaves the SM.
//
// Functional spec: localAssembler of dune/python/test/poisson.py:37-81 (per point: evaluateJacobian, geometry
// jacobianInverseTransposed / integrationElement, grad_i = J^-T dphi_i, A[i][j] += w detJ grad_i . grad_j).
//
// One block per element (grid-stride, two blocks resident per SM):
//   (1) univariate K1 tables of the element and its (p+1)^3 control points -> shared memory (homogeneous form);
//   (2) geometry of ALL quadrature points by sum factorisation with the whole block (as eval_tensor_sf.cu, O(p^4));
//       per point: 1/W, dW/dxi, J^-T, w detJ kept in shared memory;
//   (3) the quadrature points are consumed in chunks of 16 k-rows (5 points x 3 gradient components + 1 zero row; 16
//       points for the mass matrix): physical gradients G_q[c][i] of the chunk are built in shared memory straight from
//       the tables (rational basis + quotient rule, nurbsalgorithms.hh:226-248), double-buffered, ONE __syncthreads
//       per chunk, and contracted on the FP64 tensor cores (mma.sync.m8n8k4.f64, SASS DMMA) into register
//       accumulators.  K_e is symmetric: only the upper 32x32 blocks are contracted (one warp each) and mirrored
//       on the way out.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "gram_common.cuh"

namespace igab {

namespace {

// Warp w of the Gram kernel owns the upper-triangular parts of tile rows RA = w and RB = NTR-1-w of the symmetric
// matrix (8x8 DMMA tiles): (NTR - w) + (w + 1) = NTR + 1 tiles for EVERY warp -- perfectly balanced over the four
// SM sub-partitions.  Slot s of a warp is tile (RA, w + s) for s < NTR - w and tile (RB, s - 1) after that; one code
// path for all warps (row / column of a slot are warp-uniform run-time values, the accumulator index is static).
template <int NTR>
struct TileSlot {
  int w, na, ra, rb;
  __device__ __forceinline__ explicit TileSlot(int warp) : w(warp), na(NTR - warp), ra(warp), rb(NTR - 1 - warp) {}
  __device__ __forceinline__ int row(int s) const { return s < na ? ra : rb; }
  __device__ __forceinline__ int col(int s) const { return s < na ? w + s : s - 1; }
};

template <int P, int Q, int NROW>
__global__ void __launch_bounds__(FCfg<P, Q>::NT, 2) gram_fused_kernel(
    PatchDev Pd, long long elem_begin, long long nelem, const double* __restrict__ tab0,
    const double* __restrict__ tab1, const double* __restrict__ tab2, const double* __restrict__ w0,
    const double* __restrict__ w1, const double* __restrict__ w2, double fconst, double* __restrict__ Ke,
    double* __restrict__ fe) {
  using C = FCfg<P, Q>;
  constexpr int NB1 = C::NB1, NB = C::NB, NQ = C::NQ, Q2 = C::Q2, NBP = C::NBP, NT = C::NT, KC = C::KC, LD = C::LD;
  constexpr int CS = C::CS;
  // k-rows per chunk and number of chunks: see the build below
  constexpr int KCX = NROW == 3 ? (3 * Q + 3) / 4 * 4 : KC;
  constexpr int pts = KC / NROW, nchunk = NROW == 3 ? Q * Q : (NQ + pts - 1) / pts;
  static_assert(KCX <= KC, "chunk buffer too small");
  extern __shared__ __align__(16) double fsm[];
  double* const sT = fsm;
  double* const sW = sT + C::SZ_T;
  double* const rA = sW + C::SZ_W;
  double* const sS2 = rA + C::SZ_A;
  double* const sC = sS2 + C::SZ_S2;
  double* const sA = sC + C::SZ_C;  // [2][KC][LD]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  for (int idx = tid; idx < C::SZ_SA; idx += NT) sA[idx] = 0.0;  // zero rows / padding columns stay zero
  for (int idx = tid; idx < C::SZ_W; idx += NT) sW[idx] = 0.0;

  for (long long el = blockIdx.x; el < nelem; el += gridDim.x) {
    const long long e = elem_begin + el;
    IGAB_PT_BEGIN
    element_prepare<P, Q, NT>(Pd, e, tab0, tab1, tab2, w0, w1, w2, sT, sW, rA, sS2, sC);
    IGAB_PT(0)

    // ---- (3) chunked Gram contraction
    // Poisson (NROW == 3): a chunk is one line (q1, q2) of the rule, its Q points x 3 gradient components are the k-rows
    // (padded to a multiple of 4 with rows that stay zero).  Thread <-> (basis function i, half of the line): the
    // function's indices, weight and direction-0 row live in registers for the whole element, the (q1, q2) factor is
    // formed once per chunk, so that an item costs 4 multiplications + 12 FMAs and no index arithmetic.
    // Mass (NROW == 1): 24 points per chunk, generic item mapping.
    constexpr int QH = (Q + 1) / 2;
    [[maybe_unused]] const int bi_ = tid % NBP, half_ = tid / NBP;
    [[maybe_unused]] const bool bvalid = bi_ < NB;
    [[maybe_unused]] const int bic = bvalid ? bi_ : 0;
    [[maybe_unused]] const int bi1 = (bic / NB1) % NB1, bi2 = bic / (NB1 * NB1);
    [[maybe_unused]] double2 areg[QH];
    [[maybe_unused]] double bw = 0.0;
    if constexpr (NROW == 3) {
      bw = sW[bic];
#pragma unroll
      for (int u = 0; u < QH; ++u) {
        const int q0 = min(half_ * QH + u, Q - 1);
        areg[u] = lds2f(sT + (q0 * NB1 + bic % NB1) * 2);
      }
    }
    auto build = [&](int ch) {
      double* A = sA + (ch & 1) * KC * LD;
      if constexpr (NROW == 3) {
        if (!bvalid) return;  // padding columns stay zero
        const int q1 = ch % Q, q2 = ch / Q;
        const double2 b = lds2f(sT + C::TAB + (q1 * NB1 + bi1) * 2), c = lds2f(sT + 2 * C::TAB + (q2 * NB1 + bi2) * 2);
        const double cw = c.x * bw;
        const double BC0 = b.x * cw, BC1 = b.y * cw, BC2 = b.x * (c.y * bw);
#pragma unroll
        for (int u = 0; u < QH; ++u) {
          const int q0 = half_ * QH + u;
          if (q0 < Q) {
            const double* k = sC + CS * (q0 + Q * ch);
            const double2 k01 = lds2f(k), k23 = lds2f(k + 2), k45 = lds2f(k + 4), k67 = lds2f(k + 6),
                          k89 = lds2f(k + 8), kab = lds2f(k + 10), kc = lds2f(k + 12);
            const double Ah = areg[u].x * BC0, A0 = areg[u].y * BC0, A1 = areg[u].x * BC1, A2 = areg[u].x * BC2;
            A[(3 * q0 + 0) * LD + bi_] = k45.x * A0 + k45.y * A1 + k67.x * A2 - k01.y * Ah;
            A[(3 * q0 + 1) * LD + bi_] = k67.y * A0 + k89.x * A1 + k89.y * A2 - k23.x * Ah;
            A[(3 * q0 + 2) * LD + bi_] = kab.x * A0 + kab.y * A1 + kc.x * A2 - k23.y * Ah;
          }
        }
      } else {
#pragma unroll 2
        for (int item = tid; item < pts * NBP; item += NT) {
          const int ql = item / NBP, i = item % NBP;
          const int pt = ch * pts + ql;
          double r0 = 0.0;
          if (pt < NQ && i < NB) {
            const int i0 = i % NB1, i1 = (i / NB1) % NB1, i2 = i / (NB1 * NB1);
            const int q0 = pt % Q, q1 = (pt / Q) % Q, q2 = pt / Q2;
            const double2 a = lds2f(sT + (q0 * NB1 + i0) * 2), b = lds2f(sT + C::TAB + (q1 * NB1 + i1) * 2),
                          c = lds2f(sT + 2 * C::TAB + (q2 * NB1 + i2) * 2);
            r0 = (a.x * b.x) * (sW[i] * c.x) * sC[CS * pt + 15];
          }
          A[ql * LD + i] = r0;
        }
      }
    };
    constexpr int NS = C::NTR + 1;  // tiles per warp
    const TileSlot<C::NTR> slot(warp);
    double acc[NS][2];
    int colOff[NS];
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) {
      acc[sl][0] = acc[sl][1] = 0.0;
      colOff[sl] = 8 * slot.col(sl);
    }
    build(0);
    IGAB_PT(1)
    __syncthreads();
    IGAB_PT(3)
#pragma unroll 1
    for (int ch = 0; ch < nchunk; ++ch) {
      const double* A = sA + (ch & 1) * KC * LD;
#pragma unroll
      for (int ks = 0; ks < KCX; ks += 4) {
        const double* Ak = A + (ks + t) * LD + g;  // rows beyond the chunk's points are zero
        const double aA = Ak[8 * slot.ra], aB = Ak[8 * slot.rb];
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) dmma8x8x4_f(acc[sl][0], acc[sl][1], sl < slot.na ? aA : aB, Ak[colOff[sl]]);
      }
      IGAB_PT(2)
      if (ch + 1 < nchunk) build(ch + 1);
      IGAB_PT(1)
      __syncthreads();
      IGAB_PT(3)
    }
    double* K = Ke + el * (long long)NB * NB;
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) {
      const int r = slot.row(sl), c = slot.col(sl);
      const int i = 8 * r + g, j = 8 * c + 2 * t;
      if (i < NB) {
        if (j < NB) K[(long long)i * NB + j] = acc[sl][0];
        if (j + 1 < NB) K[(long long)i * NB + j + 1] = acc[sl][1];
        if (r != c) {  // mirror the off-diagonal tile
          if (j < NB) K[(long long)j * NB + i] = acc[sl][0];
          if (j + 1 < NB) K[(long long)(j + 1) * NB + i] = acc[sl][1];
        }
      }
    }
    // ---- load vector f_e[i] = sum_q w detJ R_i fconst  (poisson.py:79)
    if (fe) {
      for (int i = tid; i < NB; i += NT) {
        const int i0 = i % NB1, i1 = (i / NB1) % NB1, i2 = i / (NB1 * NB1);
        const double w = sW[i];
        double s = 0.0;
        for (int pt = 0; pt < NQ; ++pt) {
          const int q0 = pt % Q, q1 = (pt / Q) % Q, q2 = pt / Q2;
          const double R = sT[(q0 * NB1 + i0) * 2] * sT[C::TAB + (q1 * NB1 + i1) * 2] *
                           (w * sT[2 * C::TAB + (q2 * NB1 + i2) * 2]) * sC[CS * pt];
          s += sC[CS * pt + 13] * R * fconst;
        }
        fe[el * NB + i] = s;
      }
    }
    IGAB_PT(4)
  }
}

// ---- fused linear-elasticity element matrices (3-D, vector-valued basis, DOFs interleaved i*3 + a) ---------------------
// K[(i,a),(j,b)] = sum_q w detJ sum_{r,t} B[r][(i,a)] D[r][t] B[t][(j,b)] with the Voigt strain operator B.  Every column
// of B is a selection of the physical gradient G_c[i], so
//     K[(i,a),(j,b)] = sum_{c,c'} Cf[a][c][b][c'] M^{cc'}[i][j],   M^{cc'}[i][j] = sum_q w detJ G_c[q][i] G_c'[q][j],
// with Cf = S^T D S folded on the host (81 numbers).  The nine Gram blocks M^{cc'} are what the tensor cores contract:
// 9 x 2 n_b^2 n_q flops instead of the 2 (6 n_q)(3 n_b)^2 of the dense B^T D B -- six times less work for the same K_e.
// One block (9 warps) per element; for every upper 32x32 block pair (bi <= bj) of the (i,j) plane warp (c,c') owns
// M^{cc'}; the gradients of 16 points at a time are built in shared memory from the tables; the nine blocks are then
// combined with Cf through shared memory and written as contiguous rows, the mirror block with the transposed
// coefficients (no symmetry of D is assumed).
struct ElasCoef {
  double cf[81];   // [a][c][b][c']
  unsigned nz[9];  // per (a, b): bit 3 c + c' set where cf[a][c][b][c'] != 0 (fill_nz; read by elas_point_matrix only)
  int ortho;       // every non-zero coefficient has (a==c, b==c') or (a==b, c==c') or (a==c', b==c): orthotropic / isotropic D
  void fill_nz() {
    ortho = 1;
    for (int ab = 0; ab < 9; ++ab) {
      const int a = ab / 3, b = ab % 3;
      nz[ab] = 0;
      for (int c = 0; c < 3; ++c)
        for (int cp = 0; cp < 3; ++cp)
          if (cf[((a * 3 + c) * 3 + b) * 3 + cp] != 0.0) {
            nz[ab] |= 1u << (3 * c + cp);
            if (!((a == c && b == cp) || (a == b && c == cp) || (a == cp && b == c))) ortho = 0;
          }
    }
  }
};

// elasticity: M4 = C^T Cf[a][.][b][.] C of one point, entry (beta, gamma) stored at dst[(4 beta + gamma) * stride].
// The helpers' FP64 FMAs queue behind the DMMAs of the other warps on the shared FP64 datapath, and this step is on the
// critical path of the pass pipeline (tools/ws_phase.py), so an orthotropic / isotropic D (the usual case; ElasCoef::ortho)
// takes its structure literally instead of the 84 FMAs of the general triple product:
//   a != b:  M4 = n C[a] (x) C[b] + t C[b] (x) C[a],  n = Cf[a][a][b][b], t = Cf[a][b][b][a]   (rows a, b of C only)
//   a == b:  M4 = sum_c k_c C[c] (x) C[c],            k_c = Cf[a][c][a][c]
// (tests on the coefficient mask inside the general loops were tried first: ptxas predicates them, all 84 still issue).
// The terms left out are exact zeros of the general sums, which are taken in the same order: same bits.
template <int CS>
__device__ __forceinline__ void elas_point_matrix(const double* sC, const ElasCoef& coef, int ca, int cb, int pt,
                                                  double* dst, int stride) {
  const double* o = sC + CS * pt;
  if (coef.ortho) {
    if (ca != cb) {
      const int lo = min(ca, cb), hi = max(ca, cb);  // rows in ascending order, as the general loop meets them
      double Cl[4], Ch[4], Tl[4], Th[4];
      Cl[0] = -o[1 + lo];
      Ch[0] = -o[1 + hi];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        Cl[1 + d] = o[4 + 3 * lo + d];
        Ch[1 + d] = o[4 + 3 * hi + d];
      }
      // Tm[c][.] = Cf[a][c][b][c'] C[c'][.]: row c = a pairs with c' = b, row c = b with c' = a
      const double fl = coef.cf[((ca * 3 + lo) * 3 + cb) * 3 + hi], fh = coef.cf[((ca * 3 + hi) * 3 + cb) * 3 + lo];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        Tl[g] = fl * Ch[g];
        Th[g] = fh * Cl[g];
      }
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int g = 0; g < 4; ++g) dst[(size_t)(b * 4 + g) * stride] = fma(Ch[b], Th[g], Cl[b] * Tl[g]);
    } else {
      double Cm[3][4], Tm[3][4];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        Cm[c][0] = -o[1 + c];
#pragma unroll
        for (int d = 0; d < 3; ++d) Cm[c][1 + d] = o[4 + 3 * c + d];
        const double k = coef.cf[((ca * 3 + c) * 3 + ca) * 3 + c];
#pragma unroll
        for (int g = 0; g < 4; ++g) Tm[c][g] = k * Cm[c][g];
      }
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int g = 0; g < 4; ++g)
          dst[(size_t)(b * 4 + g) * stride] = fma(Cm[2][b], Tm[2][g], fma(Cm[1][b], Tm[1][g], Cm[0][b] * Tm[0][g]));
    }
    return;
  }
  double Cm[3][4], Tm[3][4];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    Cm[c][0] = -o[1 + c];
#pragma unroll
    for (int d = 0; d < 3; ++d) Cm[c][1 + d] = o[4 + 3 * c + d];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      double t = 0.0;
#pragma unroll
      for (int cp = 0; cp < 3; ++cp) t = fma(coef.cf[((ca * 3 + c) * 3 + cb) * 3 + cp], Cm[cp][g], t);
      Tm[c][g] = t;
    }
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int g = 0; g < 4; ++g)
      dst[(size_t)(b * 4 + g) * stride] = fma(Cm[2][b], Tm[2][g], fma(Cm[1][b], Tm[1][g], Cm[0][b] * Tm[0][g]));
}

template <int P, int Q>
struct ECfg {
  using F = FCfg<P, Q>;
  static constexpr int NT = 8 * 32;             // 2 x 4 warp grid over the 12 x 12 tiles of the nine blocks: 2 warps per SM sub-partition
  static constexpr int KC = 16, LDC = 64 + 4;   // chunk rows: 32 columns of block bi, 32 of block bj, padding
  static constexpr int SZ_GS = 3 * KC * LDC;    // one chunk buffer [3][KC][LDC]
  static constexpr int MS = 33;                 // row stride of a staged 32x32 block
  static constexpr int SZ_M = 9 * 32 * MS;
  static constexpr int mx(int a, int b) { return a > b ? a : b; }
  // the preparation scratch, the two chunk buffers and the staging of the nine blocks share one region
  static constexpr int SZ_R2 = mx(mx(F::SZ_A + F::SZ_S2, 2 * SZ_GS), SZ_M);
  static constexpr int SZ_TOTAL = F::SZ_T + F::SZ_W + F::SZ_C + SZ_R2;
};

// Cf[a][c][b][c'] of an orthotropic material aligned with the axes (isotropic included) is non-zero only for
// (a==c, b==c') [normal-normal], (a==b, c==c') and (a==c', b==c) [shear-shear]: 21 of the 81 coefficients.
__host__ __device__ constexpr bool ortho_pattern(int a, int c, int b, int cp) {
  return (a == c && b == cp) || (a == b && c == cp) || (a == cp && b == c);
}

template <int P, int Q, bool ORTHO>
__global__ void __launch_bounds__(ECfg<P, Q>::NT, 2) elasticity_fused_kernel(
    PatchDev Pd, long long elem_begin, long long nelem, const double* __restrict__ tab0,
    const double* __restrict__ tab1, const double* __restrict__ tab2, const double* __restrict__ w0,
    const double* __restrict__ w1, const double* __restrict__ w2, ElasCoef coef, int symd, double fconst,
    double* __restrict__ Ke, double* __restrict__ fe) {
  using C = FCfg<P, Q>;
  using E = ECfg<P, Q>;
  constexpr int NB1 = C::NB1, NB = C::NB, NQ = C::NQ, Q2 = C::Q2, CS = C::CS;
  constexpr int NT = E::NT, KC = E::KC, LDC = E::LDC, MS = E::MS, NBLK = C::NBLK;
  constexpr int nchunk = (NQ + KC - 1) / KC;
  constexpr int N3 = 3 * NB;
  extern __shared__ __align__(16) double esm[];
  double* const sT = esm;
  double* const sW = sT + C::SZ_T;
  double* const sC = sW + C::SZ_W;
  double* const rA = sC + C::SZ_C;
  double* const sS2 = rA + C::SZ_A;
  double* const sGs = rA;  // [2][3][KC][LDC], after element_prepare
  double* const sM = rA;   // [9][32][MS], after the contraction of a block pair
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  // The nine 32x32 blocks M^{cc'} of a block pair form a 96 x 96 matrix (rows (c, i), columns (c', j)) = 12 x 12 DMMA
  // tiles; warp (wr, wq) of the 2 x 4 grid owns tile rows 6 wr .. 6 wr + 5 and tile columns 3 wq .. 3 wq + 2.
  const int wr = warp >> 2, wq = warp & 3;
  int aoff[6], boffs[3];
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    const int tr = 6 * wr + a;
    aoff[a] = (tr >> 2) * KC * LDC + (tr & 3) * 8 + g;
  }
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    const int tc = 3 * wq + b;
    boffs[b] = (tc >> 2) * KC * LDC + (tc & 3) * 8 + g;
  }
  for (int idx = tid; idx < C::SZ_W; idx += NT) sW[idx] = 0.0;

  for (long long el = blockIdx.x; el < nelem; el += gridDim.x) {
    const long long e = elem_begin + el;
    IGAB_PT_BEGIN
    element_prepare<P, Q, NT>(Pd, e, tab0, tab1, tab2, w0, w1, w2, sT, sW, rA, sS2, sC);
    IGAB_PT(0)
    double* K = Ke + el * (long long)N3 * N3;
    for (int bi = 0; bi < NBLK; ++bi)
      for (int bj = bi; bj < NBLK; ++bj) {
        const int ncol = (bi == bj) ? 32 : 64;
        const int boff = (bi == bj) ? 0 : 32;
        // physical gradients (pre-scaled by sqrt(w detJ)) of the 16 points of chunk ch for the columns of blocks bi, bj
        auto build = [&](int ch) {
          double* buf = sGs + (ch & 1) * E::SZ_GS;
#pragma unroll 2
          for (int item = tid; item < KC * ncol; item += NT) {
            const int ql = item / ncol, r = item % ncol;
            const int i = (r < 32 ? bi * 32 : bj * 32 - 32) + r;
            const int pt = ch * KC + ql;
            double r0 = 0.0, r1 = 0.0, r2 = 0.0;
            if (pt < NQ && i < NB) {
              const int i0 = i % NB1, i1 = (i / NB1) % NB1, i2 = i / (NB1 * NB1);
              const int q0 = pt % Q, q1 = (pt / Q) % Q, q2 = pt / Q2;
              const double2 a = lds2f(sT + (q0 * NB1 + i0) * 2), b = lds2f(sT + C::TAB + (q1 * NB1 + i1) * 2),
                            c = lds2f(sT + 2 * C::TAB + (q2 * NB1 + i2) * 2);
              const double* k = sC + CS * pt;
              const double2 k01 = lds2f(k), k23 = lds2f(k + 2), k45 = lds2f(k + 4), k67 = lds2f(k + 6),
                            k89 = lds2f(k + 8), kab = lds2f(k + 10), kc = lds2f(k + 12);
              const double w = sW[i];
              const double t00 = a.x * b.x, wn2 = w * c.x;
              const double Ah = t00 * wn2;
              const double A0 = (a.y * b.x) * wn2, A1 = (a.x * b.y) * wn2, A2 = t00 * (w * c.y);
              r0 = k45.x * A0 + k45.y * A1 + k67.x * A2 - k01.y * Ah;
              r1 = k67.y * A0 + k89.x * A1 + k89.y * A2 - k23.x * Ah;
              r2 = kab.x * A0 + kab.y * A1 + kc.x * A2 - k23.y * Ah;
            }
            buf[(0 * KC + ql) * LDC + r] = r0;
            buf[(1 * KC + ql) * LDC + r] = r1;
            buf[(2 * KC + ql) * LDC + r] = r2;
          }
        };
        double acc[6][3][2];
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
        build(0);
        IGAB_PT(1)
        __syncthreads();
        IGAB_PT(3)
#pragma unroll 1
        for (int ch = 0; ch < nchunk; ++ch) {
          const double* Gk = sGs + (ch & 1) * E::SZ_GS + t * LDC;
#pragma unroll
          for (int ks = 0; ks < KC; ks += 4) {  // rows of points beyond the rule are zero
            double af[6], bf[3];
#pragma unroll
            for (int a = 0; a < 6; ++a) af[a] = Gk[ks * LDC + aoff[a]];
#pragma unroll
            for (int b = 0; b < 3; ++b) bf[b] = Gk[ks * LDC + boff + boffs[b]];
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
              for (int b = 0; b < 3; ++b) dmma8x8x4_f(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
          }
          IGAB_PT(2)
          if (ch + 1 < nchunk) build(ch + 1);  // into the other buffer, overlapping the other warps' MMAs
          IGAB_PT(1)
          __syncthreads();
          IGAB_PT(3)
        }
        // all warps are past the last chunk: the buffers are reused as staging, block (c, c') at sM[(3c + c') * 32 * MS]
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            const int tr = 6 * wr + a, tc = 3 * wq + b;
            double* M = sM + (((tr >> 2) * 3 + (tc >> 2)) * 32 + (tr & 3) * 8 + g) * MS + (tc & 3) * 8 + 2 * t;
            M[0] = acc[a][b][0];
            M[1] = acc[a][b][1];
          }
        __syncthreads();
        // combine: direct block rows (i,a), columns (j,b); lanes over j -> contiguous 3-double groups per row
        for (int item = tid; item < 32 * 32; item += NT) {
          const int il = item / 32, jl = item % 32;
          const int i = bi * 32 + il, j = bj * 32 + jl;
          if (i < NB && j < NB) {
            double m[9];
#pragma unroll
            for (int pr = 0; pr < 9; ++pr) m[pr] = sM[(pr * 32 + il) * MS + jl];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
              for (int b = 0; b < 3; ++b) {
                double v = 0.0;
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                  for (int cp = 0; cp < 3; ++cp)
                    if (!ORTHO || ortho_pattern(a, c, b, cp)) v = fma(coef.cf[((a * 3 + c) * 3 + b) * 3 + cp], m[c * 3 + cp], v);
                K[(long long)(3 * i + a) * N3 + 3 * j + b] = v;
                // symmetric D: the mirror block is the transpose of this one -- keep the value in the slot of M^{ab}
                if (symd) sM[((a * 3 + b) * 32 + il) * MS + jl] = v;
              }
          }
        }
        if (bi != bj && symd) {
          __syncthreads();
          // mirror block by transposition through shared memory (no arithmetic): lanes over i -> contiguous rows
          for (int item = tid; item < 32 * 32; item += NT) {
            const int jl = item / 32, il = item % 32;
            const int i = bi * 32 + il, j = bj * 32 + jl;
            if (i < NB && j < NB) {
#pragma unroll
              for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int a = 0; a < 3; ++a) K[(long long)(3 * j + b) * N3 + 3 * i + a] = sM[((a * 3 + b) * 32 + il) * MS + jl];
            }
          }
        } else if (bi != bj) {
          // mirror block: K[(j,b),(i,a)] = sum Cf[b][c'][a][c] M^{cc'}[i][j]; lanes over i -> contiguous rows
          for (int item = tid; item < 32 * 32; item += NT) {
            const int jl = item / 32, il = item % 32;
            const int i = bi * 32 + il, j = bj * 32 + jl;
            if (i < NB && j < NB) {
              double m[9];
#pragma unroll
              for (int pr = 0; pr < 9; ++pr) m[pr] = sM[(pr * 32 + il) * MS + jl];
#pragma unroll
              for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                  double v = 0.0;
#pragma unroll
                  for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int cp = 0; cp < 3; ++cp)
                      if (!ORTHO || ortho_pattern(b, cp, a, c)) v = fma(coef.cf[((b * 3 + cp) * 3 + a) * 3 + c], m[c * 3 + cp], v);
                  K[(long long)(3 * j + b) * N3 + 3 * i + a] = v;
                }
            }
          }
        }
        __syncthreads();  // the staging area becomes chunk buffer 0 of the next block pair
        IGAB_PT(4)
      }
    if (fe) {
      for (int i = tid; i < NB; i += NT) {
        const int i0 = i % NB1, i1 = (i / NB1) % NB1, i2 = i / (NB1 * NB1);
        const double w = sW[i];
        double s = 0.0;
        for (int pt = 0; pt < NQ; ++pt) {
          const int q0 = pt % Q, q1 = (pt / Q) % Q, q2 = pt / Q2;
          const double R = sT[(q0 * NB1 + i0) * 2] * sT[C::TAB + (q1 * NB1 + i1) * 2] *
                           (w * sT[2 * C::TAB + (q2 * NB1 + i2) * 2]) * sC[CS * pt];
          s += sC[CS * pt + 13] * R * fconst;
        }
        fe[el * N3 + 3 * i] = fe[el * N3 + 3 * i + 1] = fe[el * N3 + 3 * i + 2] = s;
      }
    }
  }
}

template <int P, int Q>
igab_status launch_elas(const PatchDev& Pd, SfWorkspace& ws, const double* D_host, double fconst, long long eb,
                        long long nel, const double* const* w_dev, double* Ke, double* fe, cudaStream_t stream) {
  using E = ECfg<P, Q>;
  // Cf[a][c][b][c'] = sum_{r,t} S[r][a][c] D[r][t] S[t][b][c'], S = Voigt selector (xx,yy,zz,yz,xz,xy; engineering shear)
  static const int S[9][3] = {{0, 0, 0}, {1, 1, 1}, {2, 2, 2}, {3, 1, 2}, {3, 2, 1}, {4, 0, 2}, {4, 2, 0}, {5, 0, 1}, {5, 1, 0}};
  ElasCoef coef;
  for (double& v : coef.cf) v = 0.0;
  for (auto& s1 : S)
    for (auto& s2 : S) coef.cf[((s1[1] * 3 + s1[2]) * 3 + s2[1]) * 3 + s2[2]] += D_host[s1[0] * 6 + s2[0]];
  coef.fill_nz();
  // orthotropic / isotropic D (the usual case): only 21 coefficients can be non-zero, the combination is specialised
  bool ortho = true;
  for (int a = 0; a < 3; ++a)
    for (int c = 0; c < 3; ++c)
      for (int b = 0; b < 3; ++b)
        for (int cp = 0; cp < 3; ++cp)
          if (!ortho_pattern(a, c, b, cp) && coef.cf[((a * 3 + c) * 3 + b) * 3 + cp] != 0.0) ortho = false;
  const size_t smem = (size_t)E::SZ_TOTAL * sizeof(double);
  KernelCfg kc;
  if (igab_status cst = kernel_config(ortho ? (const void*)elasticity_fused_kernel<P, Q, true>
                                            : (const void*)elasticity_fused_kernel<P, Q, false>, E::NT, smem, &kc))
    return cst;
  const int sms = kc.sms, bps = kc.blocksPerSm;
  const unsigned grid = (unsigned)std::min<long long>(nel, (long long)bps * sms);
  // symmetric D (Cf[a][c][b][c'] == Cf[b][c'][a][c]): mirror blocks are written by transposition instead of recomputation
  int symd = 1;
  for (int a = 0; a < 3; ++a)
    for (int c = 0; c < 3; ++c)
      for (int b = 0; b < 3; ++b)
        for (int cp = 0; cp < 3; ++cp)
          if (coef.cf[((a * 3 + c) * 3 + b) * 3 + cp] != coef.cf[((b * 3 + cp) * 3 + a) * 3 + c]) symd = 0;
  if (ortho)
    elasticity_fused_kernel<P, Q, true><<<grid, E::NT, smem, stream>>>(Pd, eb, nel, ws.tab[0], ws.tab[1], ws.tab[2], w_dev[0],
                                                                       w_dev[1], w_dev[2], coef, symd, fconst, Ke, fe);
  else
    elasticity_fused_kernel<P, Q, false><<<grid, E::NT, smem, stream>>>(Pd, eb, nel, ws.tab[0], ws.tab[1], ws.tab[2], w_dev[0],
                                                                        w_dev[1], w_dev[2], coef, symd, fconst, Ke, fe);
  count_launch();
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

// =====================================================================================================================
// Sum-factorised Gram kernel (3-D Poisson / mass).  The physical gradient of a rational basis function is LINEAR in the
// four tensor-product B-spline quantities Phi_i^(beta), beta = 0 (value), 1..3 (d/dxi_0..2):
//     grad_c R_i (x sqrt(w detJ)) = w_i sum_beta C[c][beta] Phi_i^(beta),   C[c][0] = -v[c],  C[c][1+d] = M[c][d]
// (element_prepare's per-point constants), so
//     K_ij = w_i w_j sum_q sum_{beta,gamma} Phi_i^(beta)(q) M4_{beta gamma}(q) Phi_j^(gamma)(q),   M4 = C^T C (4x4, per point)
// and every Phi is a product of univariate table entries: the sum over q factorises direction by direction over PAIRS
// (i_d, j_d).  Per element 0.6 M multiply-adds instead of 3.5 M (upper tiles of the dense B^T B).  For each i2:
//   stage 1 (FMA)   T1[(j2,q0)][(beta gamma, q1)] = sum_q2 Phi2[q2][i2] Phi2[q2][j2] M4[q0,q1,q2]
//   stage 2 (DMMA)  T2[(j1,j2,i1)][(g,q0)]        = sum_{(beta gamma) in g, q1} T1 . Phi1[q1][i1] Phi1[q1][j1]
//                   g = (is beta the d/dxi_0 index, is gamma) -- the only thing stage 3 needs to know about (beta, gamma)
//   stage 3 (DMMA)  K[(j1,j2,i1)][(j0,i0)]        = sum_{g,q0} T2 . Phi0[q0][i0] Phi0[q0][j0]
// and the 8x8 accumulator tiles go straight to global memory: rows of a tile step j by NB1, columns by 1, so a warp's
// stores form contiguous runs of the 25 complete K_e rows that belong to this i2.
// KIND: 0 Poisson, 1 mass, 2 elasticity (nine (a, b) blocks, each a Poisson-like contraction with its own 4x4 M4)
template <int P, int Q, int KIND>
__global__ void __launch_bounds__(256, (SCfg<P, Q, KIND>::SZ_TOTAL * 8 > 113 * 1024 ? 1 : 2)) gram_sf_kernel(
    PatchDev Pd, long long elem_begin, long long nelem, const double* __restrict__ tab0,
    const double* __restrict__ tab1, const double* __restrict__ tab2, const double* __restrict__ w0,
    const double* __restrict__ w1, const double* __restrict__ w2, ElasCoef coef, double fconst,
    double* __restrict__ Ke, double* __restrict__ fe, WTabs wt) {
  using C = FCfg<P, Q>;
  using S = SCfg<P, Q, KIND>;
  constexpr bool MASS = S::MASS, ELAS = S::ELAS;
  constexpr int NCP = ELAS ? 3 : 1, NROWK = NCP * S::NB;  // components per basis function, rows of K_e
  constexpr int NB1 = S::NB1, NP = S::NP, NB = S::NB, NQ = S::NQ, NT = S::NT, CS = S::CSS, TAB = C::TAB;
  constexpr int LDA2 = S::LDA2, LDB2 = S::LDB2, LDA3 = S::LDA3;
  extern __shared__ __align__(16) double ssm[];
  double* const sT = ssm;
  double* const sW = sT + C::SZ_T;
  double* const sC = sW + C::SZ_W;
  double* const rX = sC + S::SZ_CS;       // element_prepare's scratch (rA | sS2), then T1
  double* const sT1 = rX;
  double* const sM = rX + S::SZ_X;
  double* const sT2 = sM + S::SZ_M;
  double* const sB2 = sT2 + S::SZ_T2;
  double* const sTw = sB2 + S::SZ_B2;  // weighted tables (tensor-product weights)
  const bool useW = wt.t0 != nullptr;
  const double* const sTc = useW ? sTw : sT;  // the tables the contraction stages read
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lr = lane >> 2, lc = lane & 3;
  for (int idx = tid; idx < C::SZ_W; idx += NT) sW[idx] = 0.0;
  for (int idx = tid; idx < S::SZ_B2; idx += NT) sB2[idx] = 0.0;  // padding rows stay zero

  // ---- per-thread constants of the fragment <-> index maps (the same for every element)
  // stage 2: this warp's tiles (tile row mt2, tiles nt2 .. nt2 + TPW - 1); where C[lr][2 lc + h] of tile s goes in T2
  const int tile0 = warp * S::TPW;
  const bool has2 = tile0 < S::MT2 * S::NT2;
  const int mt2 = has2 ? tile0 / S::NT2 : 0, nt2 = has2 ? tile0 % S::NT2 : 0;
  int t2off[S::TPW][2];
  {
    const int m = 8 * mt2 + lr;
#pragma unroll
    for (int s = 0; s < S::TPW; ++s)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = S::cperm(8 * (nt2 + s) + 2 * lc + h);
        t2off[s][h] = (has2 && m < S::M2 && n < NP) ? ((n % NB1) + NB1 * (m / Q) + NP * (n / NB1)) * LDA3 + (m % Q) : -1;
      }
  }
  // stage 3: column position 8 nt + 2 lc + h of a tile holds the pair n = cperm(position) = 8 nt + lc + 4 h, (j0, i0) =
  // (n % NB1, n / NB1): offset i0 NB + j0 | i0 << 16 | j0 << 20.  The four lanes of a quad then store four CONSECUTIVE
  // entries of a row of K_e (every other one in the natural order: twice the sectors per store instruction)
  int c3[S::NT3][2];
#pragma unroll
  for (int nt = 0; nt < S::NT3; ++nt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = S::cperm(8 * nt + 2 * lc + h);
      c3[nt][h] = n < NP ? ((n / NB1) * NB + (n % NB1)) | ((n / NB1) << 16) | ((n % NB1) << 20) : -1;
    }

  // stage 3: row lr of this warp's u-th tile (tile row warp + 8 u) is (j1, j2, i1): NB1 i1 | (NB1 j1 + NP j2) << 8, or -1
  constexpr int NMT = (S::MT3 + S::NWARP - 1) / S::NWARP;
  int r3[NMT];
#pragma unroll
  for (int u = 0; u < NMT; ++u) {
    const int mp = 8 * (warp + S::NWARP * u) + lr;
    r3[u] = mp < S::M3 ? (NB1 * (mp / NP)) | ((NB1 * (mp % NB1) + NP * ((mp / NB1) % NB1)) << 8) : -1;
  }

  // a block takes a CONTIGUOUS run of elements (as gram_sym_kernel): consecutive elements share e1, so the stage-2 operand
  // B2, which depends on direction 1 only, is rebuilt once per row of elements instead of once per element
  // (not for the mass matrix, whose B2 is tiny and which is bound by its stores: contiguous runs measured 0.89-0.99x there,
  //  1.02-1.17x for Poisson / elasticity)
  constexpr bool RUNS = !MASS;
  const long long chunk = (nelem + gridDim.x - 1) / gridDim.x;
  const long long el_lo = RUNS ? blockIdx.x * chunk : blockIdx.x;
  const long long el_hi = RUNS ? (el_lo + chunk < nelem ? el_lo + chunk : nelem) : nelem;
  const long long el_step = RUNS ? 1 : gridDim.x;
  int b2_e1 = -1;
  for (long long el = el_lo; el < el_hi; el += el_step) {
    const long long e = elem_begin + el;
    if (useW) {
      // nobody reads sTw any more: the previous element's last stage 1 lies two barriers back; element_prepare's barriers
      // order these writes before the operand builds below
      const int e0 = (int)(e % Pd.nel[0]), e1 = (int)((e / Pd.nel[0]) % Pd.nel[1]), e2 = (int)(e / ((long long)Pd.nel[0] * Pd.nel[1]));
      for (int j = tid; j < TAB; j += NT) {
        sTw[j] = __ldg(wt.t0 + (size_t)e0 * TAB + j);
        sTw[TAB + j] = __ldg(wt.t1 + (size_t)e1 * TAB + j);
        sTw[2 * TAB + j] = __ldg(wt.t2 + (size_t)e2 * TAB + j);
      }
    }
    element_prepare<P, Q, NT, CS>(Pd, e, tab0, tab1, tab2, w0, w1, w2, sT, sW, rX, rX + C::SZ_A, sC);
    // ---- T1 zeroed (its padding columns must stay zero; the scratch it aliases is free after element_prepare)
    for (int idx = tid; idx < S::SZ_T1; idx += NT) sT1[idx] = 0.0;
    // ---- B2[(g, c, q1)][pi1 = j1 + NB1 i1] = Phi1[q1][i1][b1(beta)] Phi1[q1][j1][b1(gamma)],  b1(beta) = (beta == 2)
    const int e1cur = (int)((e / Pd.nel[0]) % Pd.nel[1]);
    if (e1cur != b2_e1)
#pragma unroll
    for (int g = 0; g < S::NG; ++g) {
      for (int idx = tid; idx < S::ncombo(g) * Q * NP; idx += NT) {
        const int n = idx % NP, r = idx / NP, q1 = r % Q, c = r / Q;
        const int j1 = n % NB1, i1 = n / NB1;
        const int bb = S::beta_of(g, c) == 2, bg = S::gamma_of(g, c) == 2;
        sB2[(S::goff(g) + c * Q + q1) * LDB2 + S::cpos(n)] =
            sTc[TAB + (q1 * NB1 + i1) * 2 + bb] * sTc[TAB + (q1 * NB1 + j1) * 2 + bg];
      }
    }
    b2_e1 = e1cur;
    // ---- B3[(g, q0)][pi0 = j0 + NB1 i0] = Phi0[q0][i0][b0(beta)] Phi0[q0][j0][b0(gamma)], group g = (b0(beta), b0(gamma)):
    //      0 (0,0), 1 (1,0), 2 (0,1), 3 (1,1).  Built once in the (still unused) T2 region, then held as fragments.
    for (int idx = tid; idx < S::K3 * S::NPP; idx += NT) {
      const int n = idx % S::NPP, k = idx / S::NPP;
      double v = 0.0;
      if (k < S::NG * Q && n < NP) {
        const int g = k / Q, q0 = k % Q, j0 = n % NB1, i0 = n / NB1;
        const int bb = (g == 1 || g == 3) ? 1 : 0, bg = (g >= 2) ? 1 : 0;
        v = sTc[(q0 * NB1 + i0) * 2 + bb] * sTc[(q0 * NB1 + j0) * 2 + bg];
      }
      sT2[k * LDB2 + S::cpos(n)] = v;
    }
    __syncthreads();
    double b3[S::K3 / 4][S::NT3];
#pragma unroll
    for (int ks = 0; ks < S::K3 / 4; ++ks)
#pragma unroll
      for (int nt = 0; nt < S::NT3; ++nt) b3[ks][nt] = sT2[(4 * ks + lc) * LDB2 + 8 * nt + lr];
    __syncthreads();
    for (int idx = tid; idx < S::SZ_T2; idx += NT) sT2[idx] = 0.0;  // padding rows / columns of T2 must be zero

    double* const K = Ke + el * (long long)NROWK * NROWK;
    // elasticity: the (a, b) component blocks, b innermost INSIDE a pass, so that the three stores that complete a
    // 32-byte sector of the interleaved (j, b) columns reach L2 close together (b outermost: 4x the DRAM traffic)
    for (int ca = 0; ca < NCP; ++ca)
    for (int i2 = 0; i2 < NB1; ++i2)
    for (int cb = 0; cb < NCP; ++cb) {
    // ---- per point: M4 = C^T C (symmetric 4x4; mass: (sqrt(w detJ)/W)^2; elasticity: C^T Cf[a][.][b][.] C)
    if (ELAS || i2 == 0) {
    for (int pt = tid; pt < NQ; pt += NT) {
      const double* o = sC + CS * pt;
      if constexpr (MASS) {
        sM[pt] = o[15] * o[15];
      } else if constexpr (ELAS) {
        elas_point_matrix<CS>(sC, coef, ca, cb, pt, sM + pt, NQ);
      } else {
        double Cm[3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          Cm[c][0] = -o[1 + c];
#pragma unroll
          for (int d = 0; d < 3; ++d) Cm[c][1 + d] = o[4 + 3 * c + d];
        }
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int g = b; g < 4; ++g)
            sM[S::usym(b, g) * NQ + pt] = Cm[0][b] * Cm[0][g] + Cm[1][b] * Cm[1][g] + Cm[2][b] * Cm[2][g];
      }
    }
    __syncthreads();
    }
    {
      // ---- stage 1: thread <-> (combo, q1, half of the q0 range), all j2: the M4 line and the i2 factor in registers
      // the q0 range in pieces of two: NCOMBO Q ceil(Q / 2) items (240 of the block's 256 threads at p = 4, Poisson)
      constexpr int QH = 2, NSPL = (Q + QH - 1) / QH;
      for (int item = tid; item < S::NCOMBO * Q * NSPL; item += NT) {
        const int half = item % NSPL, r = item / NSPL, q1 = r % Q, cc = r / Q;
        int g = 0, c = cc;
        if constexpr (!MASS) {
          if (cc >= 15) { g = 3; c = 0; }
          else if (cc >= 12) { g = 2; c = cc - 12; }
          else if (cc >= 9) { g = 1; c = cc - 9; }
        }
        const int be = S::beta_of(g, c), ga = S::gamma_of(g, c);
        const int bb = be == 3, bg = ga == 3;
        const int qa = half * QH;                       // q0 in [qa, qa + QH) (clipped to Q)
        const double* mrow = sM + S::usym(be, ga) * NQ + Q * q1;
        double am[QH][Q];
#pragma unroll
        for (int q2 = 0; q2 < Q; ++q2) {
          const double a2 = sTc[2 * TAB + (q2 * NB1 + i2) * 2 + bb];
#pragma unroll
          for (int u = 0; u < QH; ++u) am[u][q2] = a2 * mrow[min(qa + u, Q - 1) + Q * Q * q2];
        }
        double* trow = sT1 + S::goff(g) + c * Q + q1;
        // all sums first, the T1 stores afterwards: shared stores may alias the table loads as far as the compiler can
        // tell, so a store inside the j2 loop pins the next iteration's loads behind it
        double sums[NB1][QH];
#pragma unroll
        for (int j2 = 0; j2 < NB1; ++j2) {
          double f[Q];
#pragma unroll
          for (int q2 = 0; q2 < Q; ++q2) f[q2] = sTc[2 * TAB + (q2 * NB1 + j2) * 2 + bg];
#pragma unroll
          for (int u = 0; u < QH; ++u) {
            double sum = 0.0;
#pragma unroll
            for (int q2 = 0; q2 < Q; ++q2) sum = fma(am[u][q2], f[q2], sum);
            sums[j2][u] = sum;
          }
        }
#pragma unroll
        for (int j2 = 0; j2 < NB1; ++j2)
#pragma unroll
          for (int u = 0; u < QH; ++u)
            if (qa + u < Q) trow[(j2 * Q + qa + u) * LDA2] = sums[j2][u];
      }
      __syncthreads();
      // ---- stage 2: C_g[(j2,q0)][pi1] = sum_{k in g} T1[.][k] B2[k][.]  ->  T2[j1 + NB1 j2 + NP i1][g Q + q0]
      if constexpr (S::GEN2) {
        for (int tile = warp; tile < S::MT2 * S::NT2; tile += S::NWARP) {
          const int mt = tile / S::NT2, nt = tile % S::NT2;
          const double* arow = sT1 + (8 * mt + lr) * LDA2 + lc;
          const double* brow = sB2 + lc * LDB2 + 8 * nt + lr;
          const int m = 8 * mt + lr;
          int off[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int n = S::cperm(8 * nt + 2 * lc + h);
            off[h] = (m < S::M2 && n < NP) ? ((n % NB1) + NB1 * (m / Q) + NP * (n / NB1)) * LDA3 + (m % Q) : -1;
          }
#pragma unroll
          for (int g = 0; g < S::NG; ++g) {
            double a0 = 0.0, a1 = 0.0;
            const int kb = S::goff(g), nks = S::r4(S::ncombo(g) * Q) / 4;
#pragma unroll 4
            for (int ks = 0; ks < nks; ++ks) dmma8x8x4_f(a0, a1, arow[kb + 4 * ks], brow[(kb + 4 * ks) * LDB2]);
            if (off[0] >= 0) sT2[off[0] + g * Q] = a0;
            if (off[1] >= 0) sT2[off[1] + g * Q] = a1;
          }
        }
      } else if (has2) {
        const double* arow = sT1 + (8 * mt2 + lr) * LDA2 + lc;
        const double* brow = sB2 + lc * LDB2 + 8 * nt2 + lr;
        // all groups' accumulators first, the T2 stores afterwards: the stores (which may alias the operand loads as far as
        // the compiler can tell) no longer separate the groups, so their NG x TPW independent DMMA chains interleave
        double acc[S::NG][S::TPW][2];
#pragma unroll
        for (int g = 0; g < S::NG; ++g) {
#pragma unroll
          for (int s = 0; s < S::TPW; ++s) acc[g][s][0] = acc[g][s][1] = 0.0;
          const int kb = S::goff(g), nks = S::r4(S::ncombo(g) * Q) / 4;  // constants after unrolling over g
#pragma unroll
          for (int ks = 0; ks < nks; ++ks) {
            const double a = arow[kb + 4 * ks];
#pragma unroll
            for (int s = 0; s < S::TPW; ++s) dmma8x8x4_f(acc[g][s][0], acc[g][s][1], a, brow[(kb + 4 * ks) * LDB2 + 8 * s]);
          }
        }
#pragma unroll
        for (int g = 0; g < S::NG; ++g)
#pragma unroll
          for (int s = 0; s < S::TPW; ++s)
#pragma unroll
            for (int h = 0; h < 2; ++h)
              if (t2off[s][h] >= 0) sT2[t2off[s][h] + g * Q] = acc[g][s][h];
      }
      __syncthreads();
      // ---- stage 3: K[(j1,j2,i1)][pi0] = sum_{(g,q0)} T2 B3, straight to global memory with the weights w_i w_j
#pragma unroll
      for (int u = 0; u < NMT; ++u) {
        const int mt = warp + S::NWARP * u;
        if (mt >= S::MT3) break;
        double acc[S::NT3][2];
#pragma unroll
        for (int nt = 0; nt < S::NT3; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
        const double* arow = sT2 + (8 * mt + lr) * LDA3 + lc;
#pragma unroll
        for (int ks = 0; ks < S::K3 / 4; ++ks) {
          const double a = arow[4 * ks];
#pragma unroll
          for (int nt = 0; nt < S::NT3; ++nt) dmma8x8x4_f(acc[nt][0], acc[nt][1], a, b3[ks][nt]);
        }
        if (r3[u] >= 0) {
          const int ib = (r3[u] & 255) + NP * i2, jb = r3[u] >> 8;
          const double* const wI = sW + ib;
          const double* const wJ = sW + jb;
          if constexpr (ELAS) {
            double* const Kb = K + (long long)(ib * 3 + ca) * NROWK + jb * 3 + cb;
            if (useW) {
#pragma unroll
              for (int nt = 0; nt < S::NT3; ++nt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const int pk = c3[nt][h];
                  if (pk >= 0) Kb[((pk >> 16) & 15) * 3 * NROWK + (pk >> 20) * 3] = acc[nt][h];
                }
            } else {
#pragma unroll
              for (int nt = 0; nt < S::NT3; ++nt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const int pk = c3[nt][h];
                  if (pk >= 0) {
                    const int i0 = (pk >> 16) & 15, j0 = pk >> 20;
                    Kb[i0 * 3 * NROWK + j0 * 3] = acc[nt][h] * (wI[i0] * wJ[j0]);
                  }
                }
            }
          } else {
            double* const Kb = K + ib * NB + jb;
            if (useW) {  // the weights are inside the tables: the accumulators are the entries of K_e
#pragma unroll
              for (int nt = 0; nt < S::NT3; ++nt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const int pk = c3[nt][h];
                  if (pk >= 0) Kb[pk & 0xffff] = acc[nt][h];
                }
            } else {
#pragma unroll
              for (int nt = 0; nt < S::NT3; ++nt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const int pk = c3[nt][h];
                  if (pk >= 0) Kb[pk & 0xffff] = acc[nt][h] * (wI[(pk >> 16) & 15] * wJ[pk >> 20]);
                }
            }
          }
        }
      }
      // no barrier here: the next pass writes T1 (stage 2 is done with it) and T2 only after its own first barrier
    }
    }  // (a, i2, b): the next block rewrites sM, last read two barriers ago
    // ---- load vector f_e[i] = sum_q w detJ R_i fconst  (poisson.py:79)
    if (fe) {
      // sum-factorised: f_i = w_i sum_q0 N0[q0][i0] sum_q1 N1[q1][i1] sum_q2 N2[q2][i2] (w detJ / W)(q) fconst, one direction
      // per step (the plain loop over all points per function cost a quarter of the kernel's time); scratch in the T1
      // region, which stage 2 of the last pass was the last to read
      double* const fA = sT1;                // [q0 + Q q1][i2]
      double* const fB = sT1 + Q * Q * NB1;  // [q0][i1 + NB1 i2]
      for (int idx = tid; idx < Q * Q * NB1; idx += NT) {
        const int q01 = idx % (Q * Q), i2 = idx / (Q * Q);
        double a = 0.0;
#pragma unroll
        for (int q2 = 0; q2 < Q; ++q2) {
          const double* o = sC + CS * (q01 + Q * Q * q2);
          a = fma(sT[2 * TAB + (q2 * NB1 + i2) * 2], o[13] * o[0], a);
        }
        fA[idx] = a;
      }
      __syncthreads();
      for (int idx = tid; idx < Q * NB1 * NB1; idx += NT) {
        const int q0 = idx % Q, i12 = idx / Q, i1 = i12 % NB1, i2 = i12 / NB1;
        double b = 0.0;
#pragma unroll
        for (int q1 = 0; q1 < Q; ++q1) b = fma(sT[TAB + (q1 * NB1 + i1) * 2], fA[q0 + Q * q1 + Q * Q * i2], b);
        fB[idx] = b;
      }
      __syncthreads();
      for (int i = tid; i < NB; i += NT) {
        const int i0 = i % NB1, i12 = i / NB1;
        double f = 0.0;
#pragma unroll
        for (int q0 = 0; q0 < Q; ++q0) f = fma(sT[(q0 * NB1 + i0) * 2], fB[q0 + Q * i12], f);
        f *= sW[i] * fconst;
        if constexpr (ELAS)
          fe[el * NROWK + 3 * i] = fe[el * NROWK + 3 * i + 1] = fe[el * NROWK + 3 * i + 2] = f;
        else
          fe[el * NB + i] = f;
      }
    }
  }
}

// =====================================================================================================================
// Warp-specialised version of the sum-factorised kernel: ONE 512-thread block per SM, the same three stages per pass
// (a, i2, b), but run as a pipeline between two kinds of warps instead of barrier-separated phases of one block:
//   * 4 DMMA warps (one per SM sub-partition): stage 2 and stage 3 of pass k, nothing else -- their instruction stream is
//     fragment loads, DMMAs and the unscaled 8x8 accumulator tiles written to a shared-memory tile buffer;
//   * 12 helper warps: stage 1 (FMA) of pass k+1 while the DMMA warps are in stage 3 of pass k, and the epilogue of pass
//     k -- scaling by the rational weights w_i w_j and the stores to global memory -- while they are in stage 2 of pass
//     k+1.  A pass's 25 rows of K_e are ONE contiguous run, so the epilogue is a flat, fully coalesced scaled copy; for
//     elasticity the three b-blocks of a row meet in the tile buffer and leave together as complete rows of 3 n_b doubles
//     (the interleaved (j, b) stores of gram_sf_kernel cost it 2.2x the algorithmic DRAM traffic).
// Hand-offs are named barriers (bar.arrive on the producing side, bar.sync on the consuming side): T1 full / free, tile
// buffer full / free, plus one barrier inside each group.  The element preparation (tables, control points, geometry,
// stage operands) is done by all 16 warps together between two elements.
__device__ __forceinline__ void nb_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nb_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// ---- async bulk copy shared -> global (TMA engine; SASS: UBLKCP), bulk-group completion (as in eval_tensor_sf.cu)
__device__ __forceinline__ void ws_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void ws_bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ws_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ws_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

template <int P, int Q, int KIND>
struct WCfg {
  static constexpr int NWM = 8, NWH = 8, NT = 32 * (NWM + NWH), NTH = 32 * NWH, NTM = 32 * NWM;
  using S = SCfg<P, Q, KIND, NWM>;
  using F = FCfg<P, Q>;
  static constexpr int NCP = S::ELAS ? 3 : 1;
  static constexpr int SZ_B3 = S::K3 * S::LDB2;
  // tile buffer: the rows (i0, i1) of one i2 (x the three b-blocks).  Row slots of RSK doubles (even: 16-byte aligned slots);
  // a row starts at slot + 0 or + 1 -- whichever gives it the 16-byte phase of its destination in global memory (rows of
  // 125 / 375 doubles alternate there) -- so that it can leave through ONE async bulk store of the TMA engine
  static constexpr int RSK = F::ev(S::NB * NCP + 1);
  static constexpr int SZ_K = S::NP * RSK;
  static constexpr int SZ_TOTAL =
      F::SZ_T + F::SZ_W + S::SZ_CS + S::SZ_X + S::SZ_M + S::SZ_T2 + S::SZ_T1 + S::SZ_B2 + SZ_B3 + SZ_K + F::SZ_T;
  static constexpr bool FITS = SZ_TOTAL * 8 <= 227 * 1024;
  // registers per thread after the role split (8 + 8 warps x 32 lanes: REG_MMA + REG_HLP = 256 fills the register file)
  // 160 / 96: the DMMA branch uses 140 registers; the out-of-line slow paths of the FP64 division / square root that
  // ptxas shares between the two branches use registers up to R93, so the helpers must keep at least 96 (checked on the
  // SASS by tools/ws_regs_check.py, part of the CPU test suite).  Measured: 152 / 104 0.97x, 168 / 88 the same speed.
#ifndef IGAB_WS_REG_MMA
#define IGAB_WS_REG_MMA 160
#endif
  static constexpr int REG_MMA = IGAB_WS_REG_MMA, REG_HLP = 256 - IGAB_WS_REG_MMA;
};

template <int P, int Q, int KIND>
__global__ void __launch_bounds__(512, 1) gram_ws_kernel(
    PatchDev Pd, long long elem_begin, long long nelem, const double* __restrict__ tab0,
    const double* __restrict__ tab1, const double* __restrict__ tab2, const double* __restrict__ w0,
    const double* __restrict__ w1, const double* __restrict__ w2, ElasCoef coef, double fconst,
    double* __restrict__ Ke, double* __restrict__ fe, WTabs wt, int bulk_rows) {
  using C = FCfg<P, Q>;
  using WC = WCfg<P, Q, KIND>;
  using S = typename WC::S;
  constexpr bool MASS = S::MASS, ELAS = S::ELAS;
  constexpr int NCP = WC::NCP, NROWK = NCP * S::NB;
  constexpr int NB1 = S::NB1, NP = S::NP, NB = S::NB, NQ = S::NQ, CS = S::CSS, TAB = C::TAB;
  constexpr int NT = WC::NT, NTH = WC::NTH, NTM = WC::NTM;
  constexpr int LDA2 = S::LDA2, LDB2 = S::LDB2, LDA3 = S::LDA3;
  constexpr int NPASS = NCP * NB1 * NCP;
  constexpr int RSK = WC::RSK;
  constexpr bool ODDK = (NROWK & 1) != 0;  // rows of K_e alternate between the two 16-byte phases
  static_assert((C::SZ_T + C::SZ_W + S::SZ_CS + S::SZ_X + S::SZ_M + S::SZ_T2 + S::SZ_T1 + S::SZ_B2 + WC::SZ_B3) % 2 == 0 &&
                    RSK % 2 == 0, "tile-buffer row slots must be 16-byte aligned");
  // T1 is double-buffered (FULL / FREE barriers per buffer: + 0 / + 1): stage 1 of pass k + 1 (helper warps) overlaps stage 2
  // of pass k (DMMA warps).  With one buffer the two were serialised -- pass period = stage 1 + stage 2, 6.0 k cycles at C3
  // elasticity (tools/ws_phase.py) -- although they run on different warps.  T2 needs one buffer only: stages 2 and 3 of a
  // pass run on the same warps, one after the other.
  enum { BAR_T1_FULL = 1, BAR_T1_FREE = 3, BAR_K_FULL = 5, BAR_K_FREE = 6, BAR_MMA = 7, BAR_HLP = 8, BAR_MMA2 = 9 };
  extern __shared__ __align__(16) double ssm[];
  double* const sT = ssm;
  double* const sW = sT + C::SZ_T;
  double* const sC = sW + C::SZ_W;
  double* const rX = sC + S::SZ_CS;  // element_prepare's scratch (rA | sS2), then T1
  double* const sT1 = rX;
  double* const sM = rX + S::SZ_X;
  double* const sT2 = sM + S::SZ_M;
  double* const sT1b = sT2 + S::SZ_T2;  // second T1 buffer (odd passes)
  double* const sB2 = sT1b + S::SZ_T1;
  double* const sB3 = sB2 + S::SZ_B2;
  double* const sK = sB3 + WC::SZ_B3;
  double* const sTw = sK + WC::SZ_K;  // weighted tables (tensor-product weights)
  const bool useW = wt.t0 != nullptr;
  const bool useB = useW && bulk_rows;  // rows leave by async bulk stores
  const double* const sTc = useW ? sTw : sT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_mma = warp < WC::NWM;
  const int htid = tid - NTM;  // helper thread index (negative for the DMMA warps)
  const int lr = lane >> 2, lc = lane & 3;
  for (int idx = tid; idx < C::SZ_W; idx += NT) sW[idx] = 0.0;
  for (int idx = tid; idx < S::SZ_B2; idx += NT) sB2[idx] = 0.0;      // padding rows stay zero
  for (int idx = tid; idx < S::SZ_T2 + S::SZ_T1; idx += NT) sT2[idx] = 0.0;  // padding of T2 and of the second T1 buffer stays zero

  // ---- DMMA warps: fragment <-> index maps, the same for every element (see gram_sf_kernel)
  const int tile0 = warp * S::TPW;
  const bool has2 = is_mma && tile0 < S::MT2 * S::NT2;
  const int mt2 = has2 ? tile0 / S::NT2 : 0, nt2 = has2 ? tile0 % S::NT2 : 0;
  int t2off[S::TPW][2];
  {
    const int m = 8 * mt2 + lr;
#pragma unroll
    for (int s = 0; s < S::TPW; ++s)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = S::cperm(8 * (nt2 + s) + 2 * lc + h);
        t2off[s][h] = (has2 && m < S::M2 && n < NP) ? ((n % NB1) + NB1 * (m / Q) + NP * (n / NB1)) * LDA3 + (m % Q) : -1;
      }
  }
  // stage 3: column n of a tile is (j0, i0) = (n % NB1, n / NB1): offset into the tile buffer i0 * RSK + j0 * NCP; the
  // buffer row is i0 + NB1 i1, its start shifted by (row + group parity) & 1 when the rows of K_e alternate in phase:
  // bit 2 nt + h of pc3 = parity of i0, bit u of pr3 = parity of NB1 i1
  int c3[S::NT3][2];
  unsigned pc3 = 0, pr3 = 0;
#pragma unroll
  for (int nt = 0; nt < S::NT3; ++nt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = 8 * nt + 2 * lc + h;
      c3[nt][h] = n < NP ? (n / NB1) * RSK + (n % NB1) * NCP : -1;
      if (ODDK && n < NP) pc3 |= (unsigned)((n / NB1) & 1) << (2 * nt + h);
    }
  // stage 3: row lr of this warp's u-th tile row is (j1, j2, i1): offset NB1 i1 * RSK + (NB1 j1 + NP j2) * NCP
  constexpr int NMT = (S::MT3 + WC::NWM - 1) / WC::NWM;
  int r3[NMT];
#pragma unroll
  for (int u = 0; u < NMT; ++u) {
    const int mp = 8 * (warp + WC::NWM * u) + lr;
    r3[u] = (is_mma && mp < S::M3) ? (NB1 * (mp / NP)) * RSK + (NB1 * (mp % NB1) + NP * ((mp / NB1) % NB1)) * NCP : -1;
    if (ODDK && is_mma && mp < S::M3) pr3 |= (unsigned)((NB1 * (mp / NP)) & 1) << u;
  }

  IGAB_WPT_BEGIN
  // The roles never rejoin (each walks the elements in its own loop, the common preparation is a lambda instantiated in
  // both), so that setmaxnreg can move registers from the helper warps to the DMMA warps for the whole kernel.
  auto element_common = [&](long long el) {
    const long long e = elem_begin + el;
    __syncthreads();  // the helpers' last stage 1 of the previous element is done with sTw
    if (useW) {
      const int e0 = (int)(e % Pd.nel[0]), e1 = (int)((e / Pd.nel[0]) % Pd.nel[1]), e2 = (int)(e / ((long long)Pd.nel[0] * Pd.nel[1]));
      for (int j = tid; j < TAB; j += NT) {
        sTw[j] = __ldg(wt.t0 + (size_t)e0 * TAB + j);
        sTw[TAB + j] = __ldg(wt.t1 + (size_t)e1 * TAB + j);
        sTw[2 * TAB + j] = __ldg(wt.t2 + (size_t)e2 * TAB + j);
      }
    }
    element_prepare<P, Q, NT, CS>(Pd, e, tab0, tab1, tab2, w0, w1, w2, sT, sW, rX, rX + C::SZ_A, sC);
    IGAB_WPT(is_mma ? 8 : 9)
    // ---- between two elements, all 16 warps: T1 zeroed, the B operands of stages 2 / 3, the point matrices M4
    for (int idx = tid; idx < S::SZ_T1; idx += NT) sT1[idx] = 0.0;
#pragma unroll
    for (int g = 0; g < S::NG; ++g) {
      for (int idx = tid; idx < S::ncombo(g) * Q * NP; idx += NT) {
        const int n = idx % NP, r = idx / NP, q1 = r % Q, c = r / Q;
        const int j1 = n % NB1, i1 = n / NB1;
        const int bb = S::beta_of(g, c) == 2, bg = S::gamma_of(g, c) == 2;
        sB2[(S::goff(g) + c * Q + q1) * LDB2 + S::cpos(n)] =
            sTc[TAB + (q1 * NB1 + i1) * 2 + bb] * sTc[TAB + (q1 * NB1 + j1) * 2 + bg];
      }
    }
    for (int idx = tid; idx < S::K3 * S::NPP; idx += NT) {
      const int n = idx % S::NPP, k = idx / S::NPP;
      double v = 0.0;
      if (k < S::NG * Q && n < NP) {
        const int g = k / Q, q0 = k % Q, j0 = n % NB1, i0 = n / NB1;
        const int bb = (g == 1 || g == 3) ? 1 : 0, bg = (g >= 2) ? 1 : 0;
        v = sTc[(q0 * NB1 + i0) * 2 + bb] * sTc[(q0 * NB1 + j0) * 2 + bg];
      }
      sB3[k * LDB2 + n] = v;
    }
    if constexpr (!ELAS) {
      for (int pt = tid; pt < NQ; pt += NT) {
        const double* o = sC + CS * pt;
        if constexpr (MASS) {
          sM[pt] = o[15] * o[15];
        } else {
          double Cm[3][4];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            Cm[c][0] = -o[1 + c];
#pragma unroll
            for (int d = 0; d < 3; ++d) Cm[c][1 + d] = o[4 + 3 * c + d];
          }
#pragma unroll
          for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int g = b; g < 4; ++g)
              sM[S::usym(b, g) * NQ + pt] = Cm[0][b] * Cm[0][g] + Cm[1][b] * Cm[1][g] + Cm[2][b] * Cm[2][g];
        }
      }
    }
    __syncthreads();
    IGAB_WPT(is_mma ? 10 : 11)
  };
  if (is_mma) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WC::REG_MMA));
    for (long long el = blockIdx.x; el < nelem; el += gridDim.x) {
    element_common(el);
    double* const K = Ke + el * (long long)NROWK * NROWK;
    // 16-byte phase (0 / 1, in doubles) of row `row` of the group (i2, a) of this element in global memory:
    // (kpar + ((row + NP i2) NCP + a) NROWK) & 1 = gpar(i2, a) ^ (row & 1) for odd NROWK, kpar otherwise
    const unsigned kpar = (unsigned)(reinterpret_cast<unsigned long long>(K) >> 3) & 1u;
    auto gpar = [&](int i2, int ca) -> unsigned { return ODDK ? (kpar ^ (unsigned)((NP * i2 * NCP + ca) & 1)) : kpar; };

    {
      // ================================ DMMA warps: stage 2, stage 3 ================================================
      double b3[S::K3 / 4][S::NT3];
#pragma unroll
      for (int ks = 0; ks < S::K3 / 4; ++ks)
#pragma unroll
        for (int nt = 0; nt < S::NT3; ++nt) b3[ks][nt] = sB3[(4 * ks + lc) * LDB2 + 8 * nt + lr];
      for (int k = 0; k < NPASS; ++k) {
        const int cb = k % NCP;
        const unsigned gp = gpar((k / NCP) % NB1, k / (NCP * NB1));
        double* const T2 = sT2;
        const double* const T1 = (k & 1) ? sT1b : sT1;
        // T2 may be rewritten once every DMMA warp has left stage 3 of the previous pass: the FULL barrier below completes
        // only when all of them (and the helpers' arrivals) are in, so it orders that as well -- no barrier of its own
        nb_sync(BAR_T1_FULL + (k & 1), NT);
        IGAB_WPT(0)
        // ---- stage 2: C_g[(j2,q0)][pi1] = sum_{k in g} T1[.][k] B2[k][.]  ->  T2[j1 + NB1 j2 + NP i1][g Q + q0]
        if constexpr (S::GEN2) {
          for (int tile = warp; tile < S::MT2 * S::NT2; tile += WC::NWM) {
            const int mt = tile / S::NT2, nt = tile % S::NT2;
            const double* arow = T1 + (8 * mt + lr) * LDA2 + lc;
            const double* brow = sB2 + lc * LDB2 + 8 * nt + lr;
            const int m = 8 * mt + lr;
            int off[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int n = S::cperm(8 * nt + 2 * lc + h);
              off[h] = (m < S::M2 && n < NP) ? ((n % NB1) + NB1 * (m / Q) + NP * (n / NB1)) * LDA3 + (m % Q) : -1;
            }
#pragma unroll
            for (int g = 0; g < S::NG; ++g) {
              double a0 = 0.0, a1 = 0.0;
              const int kb = S::goff(g), nks = S::r4(S::ncombo(g) * Q) / 4;
#pragma unroll 4
              for (int ks = 0; ks < nks; ++ks) dmma8x8x4_f(a0, a1, arow[kb + 4 * ks], brow[(kb + 4 * ks) * LDB2]);
              if (off[0] >= 0) T2[off[0] + g * Q] = a0;
              if (off[1] >= 0) T2[off[1] + g * Q] = a1;
            }
          }
        } else if (has2) {
          const double* arow = T1 + (8 * mt2 + lr) * LDA2 + lc;
          const double* brow = sB2 + lc * LDB2 + 8 * nt2 + lr;
          // all groups' accumulators first, the T2 stores afterwards: NG x TPW independent DMMA chains interleave (the
          // DMMA warps run with REG_MMA registers after the role split; with 128 this cost 6 %)
          double acc[S::NG][S::TPW][2];
#pragma unroll
          for (int g = 0; g < S::NG; ++g) {
#pragma unroll
            for (int s = 0; s < S::TPW; ++s) acc[g][s][0] = acc[g][s][1] = 0.0;
            const int kb = S::goff(g), nks = S::r4(S::ncombo(g) * Q) / 4;
#pragma unroll
            for (int ks = 0; ks < nks; ++ks) {
              const double a = arow[kb + 4 * ks];
#pragma unroll
              for (int s = 0; s < S::TPW; ++s) dmma8x8x4_f(acc[g][s][0], acc[g][s][1], a, brow[(kb + 4 * ks) * LDB2 + 8 * s]);
            }
          }
#pragma unroll
          for (int g = 0; g < S::NG; ++g)
#pragma unroll
            for (int s = 0; s < S::TPW; ++s)
#pragma unroll
              for (int h = 0; h < 2; ++h)
                if (t2off[s][h] >= 0) T2[t2off[s][h] + g * Q] = acc[g][s][h];
        }
        IGAB_WPT(1)
        if (k + 2 < NPASS) nb_arrive(BAR_T1_FREE + (k & 1), NT);  // buffer consumed: the helpers may write pass k+2 into it
        if (useB && cb == 0 && k > 0) {  // the engine has read this warp's rows of the previous group out of the tile buffer
          if (lane == 0) ws_bulk_wait_read();
          __syncwarp();
        }
        nb_sync(BAR_MMA, NTM);                                     // T2 of this pass complete (and the tile buffer free)
        if (!useB && cb == 0 && k > 0) nb_sync(BAR_K_FREE, NT);    // the epilogue of the previous rows has left the tile buffer
        IGAB_WPT(2)
        // ---- stage 3: G[(j1,j2,i1)][pi0] = sum_{(g,q0)} T2 B3, unscaled, into the tile buffer
#pragma unroll
        for (int u = 0; u < NMT; ++u) {
          const int mt = warp + WC::NWM * u;
          if (mt >= S::MT3) break;
          double acc[S::NT3][2];
#pragma unroll
          for (int nt = 0; nt < S::NT3; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
          const double* arow = T2 + (8 * mt + lr) * LDA3 + lc;
#pragma unroll
          for (int ks = 0; ks < S::K3 / 4; ++ks) {
            const double a = arow[4 * ks];
#pragma unroll
            for (int nt = 0; nt < S::NT3; ++nt) dmma8x8x4_f(acc[nt][0], acc[nt][1], a, b3[ks][nt]);
          }
          if (r3[u] >= 0) {
            double* const kb = sK + r3[u] + cb;
            const unsigned t = gp ^ (pr3 >> u);
#pragma unroll
            for (int nt = 0; nt < S::NT3; ++nt)
#pragma unroll
              for (int h = 0; h < 2; ++h)
                if (c3[nt][h] >= 0) kb[c3[nt][h] + (int)(((pc3 >> (2 * nt + h)) ^ t) & 1u)] = acc[nt][h];
          }
        }
        IGAB_WPT(3)
        if (cb == NCP - 1) {
          __threadfence_block();
          if (useB) {
            // ---- the NP rows (i0, i1) of this i2 (x component a) are complete.  Tensor-product weights are inside the
            //      tables, so a row is a plain copy: it leaves through ONE async bulk store (lane 0 of the warp that takes
            //      the row; the odd double before / after the 16-byte aligned run by lanes 1 / 2), issued by the DMMA
            //      warps themselves -- three instructions per warp -- and read out of the tile buffer by the TMA engine
            //      while they run stage 2 of the next pass.  The helpers have no epilogue at all: their copy loop
            //      (1.1 MB of shared loads / global stores per element) and the FULL / FREE hand-offs put them behind
            //      with stage 1, the critical path of the pass pipeline (tools/ws_phase.py).
            ws_fence_proxy_async();  // generic-proxy writes of the tile buffer -> visible to the async proxy
            nb_sync(BAR_MMA2, NTM);
            const int i2 = (k / NCP) % NB1, ca = k / (NCP * NB1);
            for (int row = warp; row < NP; row += WC::NWM) {
              const int i = row + NP * i2;
              const int sh = (int)(ODDK ? (gp ^ (unsigned)(row & 1)) : gp);
              const double* src = sK + row * RSK + sh;
              double* dst = ELAS ? K + (long long)(i * 3 + ca) * NROWK : K + (long long)i * NB;
              const int nbulk = (NROWK - sh) & ~1;  // src + sh and dst + sh are 16-byte aligned
              if (lane == 0) ws_bulk_store(dst + sh, src + sh, nbulk * 8);
              else if (lane == 1 && sh) dst[0] = src[0];
              else if (lane == 2 && sh + nbulk < NROWK) dst[NROWK - 1] = src[NROWK - 1];
            }
            if (lane == 0) ws_bulk_commit();
            __syncwarp();
          } else {
            nb_arrive(BAR_K_FULL, NT);
          }
        }
      }
      if (useB) {  // the next element's first pass writes the tile buffer without a hand-off
        if (lane == 0) ws_bulk_wait_read();
        __syncwarp();
        if (fe) nb_arrive(BAR_K_FULL, NT);  // stage 2 of the last pass has read T1: the helpers' load vector may use it
      }
    }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WC::REG_HLP));
    for (long long el = blockIdx.x; el < nelem; el += gridDim.x) {
    element_common(el);
    double* const K = Ke + el * (long long)NROWK * NROWK;
    // 16-byte phase (0 / 1, in doubles) of row `row` of the group (i2, a) of this element in global memory:
    // (kpar + ((row + NP i2) NCP + a) NROWK) & 1 = gpar(i2, a) ^ (row & 1) for odd NROWK, kpar otherwise
    const unsigned kpar = (unsigned)(reinterpret_cast<unsigned long long>(K) >> 3) & 1u;
    auto gpar = [&](int i2, int ca) -> unsigned { return ODDK ? (kpar ^ (unsigned)((NP * i2 * NCP + ca) & 1)) : kpar; };

    {
      // ================================ helper warps: stage 1, epilogue ============================================
      auto point_matrices = [&](int ca, int cb) {  // elasticity: M4 = C^T Cf[a][.][b][.] C per point
        if constexpr (ELAS) {
          for (int pt = htid; pt < NQ; pt += NTH) elas_point_matrix<CS>(sC, coef, ca, cb, pt, sM + pt, NQ);
          nb_sync(BAR_HLP, NTH);
        }
      };
      auto stage1 = [&](int i2, int buf) {
        double* const T1w = buf ? sT1b : sT1;
        // thread <-> (combo, q1, pair of q0), all j2: the M4 line and the i2 factor in registers
        constexpr int QH = 2, NSPL = (Q + QH - 1) / QH;
        for (int item = htid; item < S::NCOMBO * Q * NSPL; item += NTH) {
          const int part = item % NSPL, r = item / NSPL, q1 = r % Q, cc = r / Q;
          int g = 0, c = cc;
          if constexpr (!MASS) {
            if (cc >= 15) { g = 3; c = 0; }
            else if (cc >= 12) { g = 2; c = cc - 12; }
            else if (cc >= 9) { g = 1; c = cc - 9; }
          }
          const int be = S::beta_of(g, c), ga = S::gamma_of(g, c);
          const int bb = be == 3, bg = ga == 3;
          const int qa = part * QH;
          const double* mrow = sM + S::usym(be, ga) * NQ + Q * q1;
          double am[QH][Q];
#pragma unroll
          for (int q2 = 0; q2 < Q; ++q2) {
            const double a2 = sTc[2 * TAB + (q2 * NB1 + i2) * 2 + bb];
#pragma unroll
            for (int u = 0; u < QH; ++u) am[u][q2] = a2 * mrow[min(qa + u, Q - 1) + Q * Q * q2];
          }
          double* trow = T1w + S::goff(g) + c * Q + q1;
          // all sums first, the T1 stores afterwards (as in gram_sf_kernel: a shared store inside the j2 loop pins the next
          // iteration's table loads behind it; the helpers' stage 1 is the critical path of the pass pipeline)
          double sums[NB1][QH];
#pragma unroll
          for (int j2 = 0; j2 < NB1; ++j2) {
            double f[Q];
#pragma unroll
            for (int q2 = 0; q2 < Q; ++q2) f[q2] = sTc[2 * TAB + (q2 * NB1 + j2) * 2 + bg];
#pragma unroll
            for (int u = 0; u < QH; ++u) {
              double sum = 0.0;
#pragma unroll
              for (int q2 = 0; q2 < Q; ++q2) sum = fma(am[u][q2], f[q2], sum);
              sums[j2][u] = sum;
            }
          }
#pragma unroll
          for (int j2 = 0; j2 < NB1; ++j2)
#pragma unroll
            for (int u = 0; u < QH; ++u)
              if (qa + u < Q) trow[(j2 * Q + qa + u) * LDA2] = sums[j2][u];
        }
        __threadfence_block();
        nb_arrive(BAR_T1_FULL + buf, NT);
      };
      point_matrices(0, 0);
      stage1(0, 0);
      IGAB_WPT(5)
      for (int k = 0; k < NPASS; ++k) {
        const int cb = k % NCP, i2 = (k / NCP) % NB1, ca = k / (NCP * NB1);
        if (k + 1 < NPASS) {
          const int k1 = k + 1;
          if (k1 >= 2) nb_sync(BAR_T1_FREE + (k1 & 1), NT);  // stage 2 of pass k1 - 2 is done with this buffer
          IGAB_WPT(4)
          point_matrices(k1 / (NCP * NB1), k1 % NCP);
          stage1((k1 / NCP) % NB1, k1 & 1);
          IGAB_WPT(5)
        }
        if (cb == NCP - 1 && !useB) {
          nb_sync(BAR_K_FULL, NT);
          IGAB_WPT(6)
          // ---- epilogue without the bulk path (general weights: scaled by w_i w_j; or IGAB_WS_BULK=0): the NP rows (i0, i1)
          //      of this i2 (x component a) as complete rows; a warp takes whole rows, its lanes consecutive columns
          //      (coalesced 256-byte stores)
          {
            const int hw = warp - WC::NWM;
            const unsigned gp = gpar(i2, ca);
            for (int row = hw; row < NP; row += WC::NWH) {
              const int i = row + NP * i2;
              const int sh = (int)(ODDK ? (gp ^ (unsigned)(row & 1)) : gp);
              const double* src = sK + row * RSK + sh;
              double* dst = ELAS ? K + (long long)(i * 3 + ca) * NROWK : K + (long long)i * NB;
              if (useW) {
                for (int c = lane; c < NROWK; c += 32) dst[c] = src[c];
              } else {
                const double wi = sW[i];
                for (int c = lane; c < NROWK; c += 32) dst[c] = src[c] * (wi * sW[ELAS ? c / 3 : c]);
              }
            }
          }
          IGAB_WPT(7)
          if (k + 1 < NPASS) nb_arrive(BAR_K_FREE, NT);
        }
      }
      IGAB_WPT(12)
      // ---- load vector f_e[i] = sum_q w detJ R_i fconst  (poisson.py:79), sum-factorised as in gram_sf_kernel; the
      //      DMMA warps are done with T1 (every T1_FULL they waited for has been consumed)
      if (fe) {
        if (useB) nb_sync(BAR_K_FULL, NT);  // (without the bulk path the last epilogue has waited for the DMMA warps)
        double* const fA = sT1;
        double* const fB = sT1 + Q * Q * NB1;
        for (int idx = htid; idx < Q * Q * NB1; idx += NTH) {
          const int q01 = idx % (Q * Q), i2 = idx / (Q * Q);
          double a = 0.0;
#pragma unroll
          for (int q2 = 0; q2 < Q; ++q2) {
            const double* o = sC + CS * (q01 + Q * Q * q2);
            a = fma(sT[2 * TAB + (q2 * NB1 + i2) * 2], o[13] * o[0], a);
          }
          fA[idx] = a;
        }
        nb_sync(BAR_HLP, NTH);
        for (int idx = htid; idx < Q * NB1 * NB1; idx += NTH) {
          const int q0 = idx % Q, i12 = idx / Q, i1 = i12 % NB1, i2 = i12 / NB1;
          double b = 0.0;
#pragma unroll
          for (int q1 = 0; q1 < Q; ++q1) b = fma(sT[TAB + (q1 * NB1 + i1) * 2], fA[q0 + Q * q1 + Q * Q * i2], b);
          fB[idx] = b;
        }
        nb_sync(BAR_HLP, NTH);
        for (int i = htid; i < NB; i += NTH) {
          const int i0 = i % NB1, i12 = i / NB1;
          double f = 0.0;
#pragma unroll
          for (int q0 = 0; q0 < Q; ++q0) f = fma(sT[(q0 * NB1 + i0) * 2], fB[q0 + Q * i12], f);
          f *= sW[i] * fconst;
          if constexpr (ELAS)
            fe[el * NROWK + 3 * i] = fe[el * NROWK + 3 * i + 1] = fe[el * NROWK + 3 * i + 2] = f;
          else
            fe[el * NB + i] = f;
        }
      }
    }
    }
  }
}

template <int P, int Q, int KIND>
igab_status launch_ws(const PatchDev& Pd, SfWorkspace& ws, const ElasCoef& coef, double fconst, long long eb,
                      long long nel, const double* const* w_dev, double* Ke, double* fe, cudaStream_t stream) {
  using WC = WCfg<P, Q, KIND>;
  const size_t sm = (size_t)WC::SZ_TOTAL * sizeof(double);
  KernelCfg kc;
  if (igab_status cst = kernel_config((const void*)gram_ws_kernel<P, Q, KIND>, WC::NT, sm, &kc)) return cst;
  const unsigned grid = (unsigned)std::min<long long>(nel, (long long)kc.blocksPerSm * kc.sms);
  const char* const benv = getenv("IGAB_WS_BULK");  // 0: the helpers copy the rows themselves (measurement switch)
  const int bulk_rows = !(benv && benv[0] == '0');
  gram_ws_kernel<P, Q, KIND><<<grid, WC::NT, sm, stream>>>(Pd, eb, nel, ws.tab[0], ws.tab[1], ws.tab[2], w_dev[0], w_dev[1],
                                                          w_dev[2], coef, fconst, Ke, fe, wtabs_of(Pd, ws), bulk_rows);
  count_launch();
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}
// Which configurations take the warp-specialised kernel.  Measured on B200 (tools/ws_bench.py, profiles/README.md): for
// elasticity at p = 4 it runs at the speed of gram_sf_kernel and writes K_e with 0.98x the algorithmic DRAM traffic
// instead of 2.2x (complete rows leave the tile buffer) -> default there.  For Poisson / mass it is 0.7-0.9x: with one
// block per SM the element preparation is exposed (a quarter of the element's time) and the shared-memory pipe, not the
// DMMA pipe, is the busiest unit -> compiled, tested, selectable with IGAB_GRAM_WS=1, not the default.
template <int P, int Q, int KIND>
constexpr bool ws_compiled() {
  return (P == 4 && Q == 5) || (P == 3 && Q == 4);
}
template <int P, int Q, int KIND>
inline bool ws_wanted() {
  const char* env = getenv("IGAB_GRAM_WS");
  return env ? env[0] != '0' : (KIND == 2 && P == 4);
}

// Poisson / mass are symmetric forms: gram_sym_kernel (assemble_sym.cu) contracts the upper block triangle only and mirrors
// it -- the default where it is faster (gram_sym_default); IGAB_GRAM_SYM=0 / 1 forces gram_sf_kernel / gram_sym_kernel
inline bool sym_wanted(int p, int q, bool mass) {
  const char* env = getenv("IGAB_GRAM_SYM");
  const int kind = mass ? IGAB_ASM_MASS : IGAB_ASM_POISSON;
  if (!gram_sym_supported(p, q, kind)) return false;
  return env ? env[0] != '0' : gram_sym_default(p, q, kind);
}

template <int P, int Q>
igab_status launch_elas_sf(const PatchDev& Pd, SfWorkspace& ws, const double* D_host, double fconst, long long eb,
                           long long nel, const double* const* w_dev, double* Ke, double* fe, cudaStream_t stream) {
  // Cf[a][c][b][c'] = sum_{r,t} S[r][a][c] D[r][t] S[t][b][c'], S = Voigt selector (xx,yy,zz,yz,xz,xy; engineering shear)
  static const int S[9][3] = {{0, 0, 0}, {1, 1, 1}, {2, 2, 2}, {3, 1, 2}, {3, 2, 1}, {4, 0, 2}, {4, 2, 0}, {5, 0, 1}, {5, 1, 0}};
  ElasCoef coef;
  for (double& v : coef.cf) v = 0.0;
  for (auto& s1 : S)
    for (auto& s2 : S) coef.cf[((s1[1] * 3 + s1[2]) * 3 + s2[1]) * 3 + s2[2]] += D_host[s1[0] * 6 + s2[0]];
  coef.fill_nz();
  if constexpr (ws_compiled<P, Q, 2>()) {
    if (ws_wanted<P, Q, 2>()) return launch_ws<P, Q, 2>(Pd, ws, coef, fconst, eb, nel, w_dev, Ke, fe, stream);
  }
  const size_t sm = (size_t)SCfg<P, Q, 2>::SZ_TOTAL * sizeof(double);
  KernelCfg kc;
  if (igab_status cst = kernel_config((const void*)gram_sf_kernel<P, Q, 2>, 256, sm, &kc)) return cst;
  const unsigned grid = (unsigned)std::min<long long>(nel, (long long)kc.blocksPerSm * kc.sms);
  gram_sf_kernel<P, Q, 2><<<grid, 256, sm, stream>>>(Pd, eb, nel, ws.tab[0], ws.tab[1], ws.tab[2], w_dev[0], w_dev[1],
                                                    w_dev[2], coef, fconst, Ke, fe, wtabs_of(Pd, ws));
  count_launch();
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

// degrees the dense-Gram kernels were never built for (p = 5): sum-factorised contraction only
template <int P, int Q>
igab_status launch_sf_only(const PatchDev& Pd, SfWorkspace& ws, bool mass, double fconst, long long eb, long long nel,
                           const double* const* w_dev, double* Ke, double* fe, cudaStream_t stream) {
  if (sym_wanted(P, Q, mass)) return launch_gram_sym(Pd, ws, mass ? IGAB_ASM_MASS : IGAB_ASM_POISSON, Q, fconst, eb, nel, w_dev, Ke, fe, stream);
  const void* fn = mass ? (const void*)gram_sf_kernel<P, Q, 1> : (const void*)gram_sf_kernel<P, Q, 0>;
  const size_t sm = (size_t)(mass ? SCfg<P, Q, 1>::SZ_TOTAL : SCfg<P, Q, 0>::SZ_TOTAL) * sizeof(double);
  const ElasCoef nocoef{};
  KernelCfg kc;
  if (igab_status cst = kernel_config(fn, 256, sm, &kc)) return cst;
  const unsigned grid = (unsigned)std::min<long long>(nel, (long long)kc.blocksPerSm * kc.sms);
  if (mass)
    gram_sf_kernel<P, Q, 1><<<grid, 256, sm, stream>>>(Pd, eb, nel, ws.tab[0], ws.tab[1], ws.tab[2], w_dev[0], w_dev[1],
                                                      w_dev[2], nocoef, fconst, Ke, fe, wtabs_of(Pd, ws));
  else
    gram_sf_kernel<P, Q, 0><<<grid, 256, sm, stream>>>(Pd, eb, nel, ws.tab[0], ws.tab[1], ws.tab[2], w_dev[0], w_dev[1],
                                                      w_dev[2], nocoef, fconst, Ke, fe, wtabs_of(Pd, ws));
  count_launch();
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

template <int P, int Q>
igab_status launch_pq(const PatchDev& Pd, SfWorkspace& ws, bool mass, double fconst, long long eb, long long nel,
                      const double* const* w_dev, double* Ke, double* fe, cudaStream_t stream) {
  using C = FCfg<P, Q>;
  {
    // the sum-factorised contraction wins from p = 3 on (measured: p = 3 1.5x, p = 4 2.4-3.8x; p = 2 0.5-0.7x);
    // IGAB_GRAM_SF=0 / 1 forces the dense-Gram / the sum-factorised kernel
    const char* env = getenv("IGAB_GRAM_SF");
    if (env ? env[0] != '0' : P >= 3) {
      if constexpr (ws_compiled<P, Q, 0>()) {
        if (ws_wanted<P, Q, 0>()) {
          const ElasCoef nc{};
          return mass ? launch_ws<P, Q, 1>(Pd, ws, nc, fconst, eb, nel, w_dev, Ke, fe, stream)
                      : launch_ws<P, Q, 0>(Pd, ws, nc, fconst, eb, nel, w_dev, Ke, fe, stream);
        }
      }
      if (sym_wanted(P, Q, mass))
        return launch_gram_sym(Pd, ws, mass ? IGAB_ASM_MASS : IGAB_ASM_POISSON, Q, fconst, eb, nel, w_dev, Ke, fe, stream);
      const void* fn = mass ? (const void*)gram_sf_kernel<P, Q, 1> : (const void*)gram_sf_kernel<P, Q, 0>;
      const size_t sm = (size_t)(mass ? SCfg<P, Q, 1>::SZ_TOTAL : SCfg<P, Q, 0>::SZ_TOTAL) * sizeof(double);
      const ElasCoef nocoef{};
      KernelCfg kc;
      if (igab_status cst = kernel_config(fn, 256, sm, &kc)) return cst;
      const unsigned grid = (unsigned)std::min<long long>(nel, (long long)kc.blocksPerSm * kc.sms);
      if (mass)
        gram_sf_kernel<P, Q, 1><<<grid, 256, sm, stream>>>(Pd, eb, nel, ws.tab[0], ws.tab[1], ws.tab[2], w_dev[0],
                                                          w_dev[1], w_dev[2], nocoef, fconst, Ke, fe, wtabs_of(Pd, ws));
      else
        gram_sf_kernel<P, Q, 0><<<grid, 256, sm, stream>>>(Pd, eb, nel, ws.tab[0], ws.tab[1], ws.tab[2], w_dev[0],
                                                          w_dev[1], w_dev[2], nocoef, fconst, Ke, fe, wtabs_of(Pd, ws));
      count_launch();
      IGAB_CUDA_TRY(cudaGetLastError());
      return IGAB_OK;
    }
  }
  const size_t smem = (size_t)C::SZ_TOTAL * sizeof(double);
  KernelCfg kc;
  if (igab_status cst = kernel_config(mass ? (const void*)gram_fused_kernel<P, Q, 1> : (const void*)gram_fused_kernel<P, Q, 3>,
                                      C::NT, smem, &kc))
    return cst;
  const int sms = kc.sms, bps = kc.blocksPerSm;
  const unsigned grid = (unsigned)std::min<long long>(nel, (long long)bps * sms);  // persistent, grid-stride
  if (mass)
    gram_fused_kernel<P, Q, 1><<<grid, C::NT, smem, stream>>>(Pd, eb, nel, ws.tab[0], ws.tab[1], ws.tab[2], w_dev[0],
                                                             w_dev[1], w_dev[2], fconst, Ke, fe);
  else
    gram_fused_kernel<P, Q, 3><<<grid, C::NT, smem, stream>>>(Pd, eb, nel, ws.tab[0], ws.tab[1], ws.tab[2], w_dev[0],
                                                             w_dev[1], w_dev[2], fconst, Ke, fe);
  count_launch();
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

}  // namespace

bool gram_fused_supported(const PatchDev& P, int kind, const int* nq1) {
  if (P.dim != 3 || P.dw != 3) return false;
  if (P.degree[0] != P.degree[1] || P.degree[0] != P.degree[2]) return false;
  if (nq1[0] != nq1[1] || nq1[0] != nq1[2]) return false;
  // full Gauss rule (p+1 points per direction), the over-integrated p+2 rule, and at p = 4 the reference's default rule
  // (order dim * max(p) = 12 -> 7 points per direction, nurbsgridentity.hh:123-124)
  const int p = P.degree[0], q = nq1[0];
  // p = 5: sum-factorised kernel only; 7 points fit the 227 KB of shared memory for the mass matrix only (Poisson: 233 KB)
  if (p == 5) return (kind == IGAB_ASM_POISSON && q == 6) || (kind == IGAB_ASM_MASS && (q == 6 || q == 7));
  return p >= 2 && p <= 4 && (q == p + 1 || q == p + 2 || (p == 4 && q == 7));
}

igab_status launch_gram_fused(const PatchDev& P, SfWorkspace& ws, const double* const* xi1d_host,
                              const double* const* xi1d_dev, const double* const* w1d_dev, int kind, double fconst,
                              long long eb, long long nel, const int* nq1, double* Ke, double* fe, cudaStream_t stream,
                              const double* D_host) {
  if (!gram_fused_supported(P, kind, nq1)) {
    set_error("assemble: no fused kernel for this configuration");
    return IGAB_UNSUPPORTED;
  }
  igab_status st = ensure_tables3d(P, ws, nq1[0], xi1d_dev, xi1d_host, stream);
  if (st) return st;
  {
    // Inside a CUDA graph capture the general path is recorded: a graph replayed after igab_patch_update_control_points
    // must not read weighted tables that were built from the previous weights.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (stream) cudaStreamIsCapturing(stream, &cap);
    ws.wuse = cap == cudaStreamCaptureStatusNone;
    if (ws.wuse && (st = ensure_wtables3d(P, ws, nq1[0], stream))) return st;
  }
  if (kind == IGAB_ASM_ELASTICITY) {
    if (!D_host) {
      set_error("assemble: elasticity needs the 6x6 D matrix");
      return IGAB_INVALID_ARGUMENT;
    }
    // elasticity: the sum-factorised kernels from p = 3 on (tools/elas_sf_bench.py: p = 3 1.13x with 4 and 5 points, p = 4
    // 1.40x; before the stage-3 column permutation and the contiguous element runs p = 3 was 0.9x), the nine-Gram-block
    // kernel at p = 2 (2.6x / 1.8x ahead there); IGAB_ELAS_SF=0 / 1 forces one of them
    const char* esf = getenv("IGAB_ELAS_SF");
    const bool use_sf = esf ? esf[0] != '0' : P.degree[0] >= 3;
#define IGAB_ELAS_CASE(PP, QQ)                                                                                   \
  if (P.degree[0] == PP && nq1[0] == QQ)                                                                         \
    return use_sf ? launch_elas_sf<PP, QQ>(P, ws, D_host, fconst, eb, nel, w1d_dev, Ke, fe, stream)              \
                  : launch_elas<PP, QQ>(P, ws, D_host, fconst, eb, nel, w1d_dev, Ke, fe, stream);
    IGAB_ELAS_CASE(2, 3) IGAB_ELAS_CASE(2, 4) IGAB_ELAS_CASE(3, 4) IGAB_ELAS_CASE(3, 5) IGAB_ELAS_CASE(4, 5) IGAB_ELAS_CASE(4, 6)
    IGAB_ELAS_CASE(4, 7)
#undef IGAB_ELAS_CASE
    set_error("assemble: no fused kernel for this configuration");
    return IGAB_UNSUPPORTED;
  }
  const bool mass = kind == IGAB_ASM_MASS;
  if (P.degree[0] == 5 && nq1[0] == 6) return launch_sf_only<5, 6>(P, ws, mass, fconst, eb, nel, w1d_dev, Ke, fe, stream);
  if (P.degree[0] == 5 && nq1[0] == 7) return launch_sf_only<5, 7>(P, ws, mass, fconst, eb, nel, w1d_dev, Ke, fe, stream);
#define IGAB_PQ_CASE(PP, QQ) \
  if (P.degree[0] == PP && nq1[0] == QQ) return launch_pq<PP, QQ>(P, ws, mass, fconst, eb, nel, w1d_dev, Ke, fe, stream);
  IGAB_PQ_CASE(2, 3) IGAB_PQ_CASE(2, 4) IGAB_PQ_CASE(3, 4) IGAB_PQ_CASE(3, 5) IGAB_PQ_CASE(4, 5) IGAB_PQ_CASE(4, 6)
  IGAB_PQ_CASE(4, 7)
#undef IGAB_PQ_CASE
  set_error("assemble: no fused kernel for this configuration");
  return IGAB_UNSUPPORTED;
}

}  // namespace igab

#ifdef IGAB_PHASE_TIMING
extern "C" int igab_debug_phase_ticks(unsigned long long* out, int reset) {
  unsigned long long z[16] = {0};
  if (out && cudaMemcpyFromSymbol(out, igab::g_phase_ticks, sizeof(z)) != cudaSuccess) return 1;
  if (reset && cudaMemcpyToSymbol(igab::g_phase_ticks, z, sizeof(z)) != cudaSuccess) return 1;
  return 0;
}
#endif

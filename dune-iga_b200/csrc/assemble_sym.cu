// assemble_sym.cu -- sum-factorised element matrices for the SYMMETRIC forms (3-D Poisson / mass), upper block triangle only.
//
// Functional spec: localAssembler of dune/python/test/poisson.py:37-81 (A[i][j] += w detJ grad_i . grad_j), as
// gram_sf_kernel (assemble_fused.cu) whose three contraction stages this kernel keeps.  What changes is WHICH index pairs
// are contracted.  K_e is symmetric, and with i = (i0, i1, i2), j = (j0, j1, j2) the 25 x 25 blocks (i2, j2) satisfy
// block(j2, i2) = block(i2, j2)^T: only the pairs j2 >= i2 are computed -- 15 of 25 at p = 4 -- and the strictly upper
// blocks are written a second time, transposed, from a shared-memory mirror buffer as complete contiguous rows.
//
// gram_sf_kernel makes one pass per i2 with the (p+1) values of j2 as "slots" of its stage-2 / stage-3 row index.  Here a
// pass takes TWO values of i2, a = pass and b = p - pass, whose j2 >= i2 ranges have (p+1-a) + (a+1) = p + 2 slots
// together -- every pass has the same size (the middle pass of an odd p+1 stands alone with (p+2)/2 slots):
//   p = 4:  passes {0,4}, {1,3}, {2}  instead of five; stage-2 rows 30 (-> 32) per pass, stage-3 rows 150 (-> 152);
//           1,840 DMMA per element instead of 3,360, nine barriers + three for the mirror buffer instead of fifteen.
// Slot s of a pass: s < na = p+1-a  ->  (i2, j2) = (a, a + s);  otherwise (b, b + s - na).
#include <algorithm>
#include <cstdlib>

#include "gram_common.cuh"

namespace igab {

namespace {

// Phase accounting (tools/sym_phase.py builds with -DIGAB_PHASE_TIMING): clock64 deltas accumulated in registers, one
// atomicAdd per phase and warp at the END of the kernel (an atomic per phase and element distorts the picture)
#ifdef IGAB_PHASE_TIMING
#define SYM_PT_DECL unsigned long long pt_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pt_last = 0;
#define SYM_PT_BEGIN pt_last = clock64();
#define SYM_PT(k)                                  \
  {                                                \
    const long long pt_now = clock64();            \
    pt_acc[k] += (unsigned long long)(pt_now - pt_last); \
    pt_last = pt_now;                              \
  }
#define SYM_PT_FLUSH                                                                 \
  if ((threadIdx.x & 31) == 0) {                                                     \
    for (int k = 0; k < 8; ++k) atomicAdd(&g_phase_ticks[k], pt_acc[k]);              \
  }
#else
#define SYM_PT_DECL
#define SYM_PT_BEGIN
#define SYM_PT(k)
#define SYM_PT_FLUSH
#endif

#ifdef IGAB_SYM_EXPERIMENT
__constant__ int c_skip;   // timing experiments only (results are wrong): bit 0 stage 1, 1 B2 build, 2 mirror, 3 K_e stores, 4 stage 2, 5 stage 3 DMMA
#define SKIP(b) (c_skip & (1 << (b)))
#else
#define SKIP(b) false
#endif

template <int P, int Q, int KIND>
struct YCfg {
  using S = SCfg<P, Q, KIND>;
  using F = FCfg<P, Q>;
  static constexpr bool MASS = KIND == 1;
  static constexpr int imax(int a, int b) { return a > b ? a : b; }
  static constexpr int NB1 = P + 1, NP = NB1 * NB1, NB = F::NB, NQ = F::NQ, NWARP = 8, NT = 256;
  static constexpr int NSLOT = NB1 + 1, NPASS = (NB1 + 1) / 2;
  static constexpr int NG = S::NG, KTOT = S::KTOT, NPP = S::NPP, K3 = S::K3;
  static constexpr int M2 = NSLOT * Q, M2P = S::r8(M2), MT2 = M2P / 8, NT2 = NPP / 8;  // stage-2 rows (slot, q0)
  static constexpr int TPW = (MT2 * NT2 + NWARP - 1) / NWARP;
  static constexpr bool GEN2 = !(TPW == 1 || NT2 % TPW == 0);
  static constexpr int M3 = NSLOT * NP, M3P = S::r8(M3), MT3 = M3P / 8, NT3 = NPP / 8;  // stage-3 rows (j1, i1, slot)
  static constexpr int NMT = (MT3 + NWARP - 1) / NWARP;
  // stage 3 split by column halves: warp w takes the column tiles of half w % 2 and the row tiles w / 2 + 4 v -- 19 row tiles
  // over 8 warps were 3 / 2 whole rows per warp (12 against 8 tiles); now 5 / 4 half rows (10 against 8), and a warp
  // holds only its half of the B3 fragments
  // (only where a half still has two tiles per A fragment and both halves are equal: p = 4.  Measured against whole rows:
  //  p = 4 +3-8 %, p = 3 with one tile per half -5 %, p = 5 with halves of 3 + 2 tiles -3 %)
  static constexpr bool SPLIT3 = NT3 >= 4 && NT3 % 2 == 0;
  static constexpr int NH0 = SPLIT3 ? (NT3 + 1) / 2 : NT3, NH1 = NT3 - NH0;
  static constexpr int NMT3 = SPLIT3 ? (MT3 + 3) / 4 : NMT;
  // stage 2 in the middle pass of an odd p + 1 (half the slots): one tile per warp when they are at most eight
  static constexpr int MIDROWS = ((NB1 + 1) / 2) * Q, MTMID = (MIDROWS + 7) / 8;
  static constexpr bool ALT2 = (NB1 % 2 == 1) && !GEN2 && TPW > 1 && MTMID * NT2 <= NWARP;
  static constexpr int LDA2 = S::LDA2, LDB2 = S::LDB2, LDA3 = S::LDA3, CSS = S::CSS;
  static constexpr int SZ_T1 = M2P * LDA2;
  static constexpr int npadc() {  // padding columns of T1: the tail of every group's k range and the tail of the row
    int n = LDA2 - KTOT;
    for (int g = 0; g < NG; ++g) n += S::r4(S::ncombo(g) * Q) - S::ncombo(g) * Q;
    return n;
  }
  static constexpr int NPADC = npadc();
  using S1 = S1Cfg<P, Q, KIND>;
  static constexpr bool S1_DMMA = S1::OK && NSLOT <= 8;
  static constexpr int SZ_X = F::ev(imax(F::SZ_A + F::SZ_S2, SZ_T1));
  static constexpr int SZ_M = S::SZ_M;
  static constexpr int SZ_T2 = F::ev(imax(M3P * LDA3, K3 * LDB2));
  static constexpr int SZ_B2 = S::SZ_B2;
  // mirror buffer [mirror slot][j0 + NB1 j1][i0 + NB1 i1]: the strictly upper blocks of a pass, transposed; it takes the
  // place of element_prepare's per-point constants, which are dead once the point matrices are built
  static constexpr int NMS = NSLOT - 2, SZ_MB = NMS * NP * NP;
  static constexpr int SZ_CM = F::ev(imax(S::SZ_CS, SZ_MB));
  static constexpr int SZ_F = F::ev(NQ);  // w detJ / W per point, kept for the load vector
  static constexpr int SZ_TOTAL = F::SZ_T + F::SZ_W + SZ_CM + SZ_X + SZ_M + SZ_T2 + SZ_B2 + F::SZ_T + SZ_F;
  static constexpr bool FITS = (size_t)SZ_TOTAL * 8 <= 232448;
  static constexpr int BLOCKS = (size_t)SZ_TOTAL * 8 > 113 * 1024 ? 1 : 2;
};

template <int P, int Q, int KIND>
__global__ void __launch_bounds__(256, (YCfg<P, Q, KIND>::BLOCKS)) gram_sym_kernel(
    PatchDev Pd, long long elem_begin, long long nelem, const double* __restrict__ tab0,
    const double* __restrict__ tab1, const double* __restrict__ tab2, const double* __restrict__ w0,
    const double* __restrict__ w1, const double* __restrict__ w2, double fconst, double* __restrict__ Ke,
    double* __restrict__ fe, WTabs wt) {
  using C = FCfg<P, Q>;
  using S = SCfg<P, Q, KIND>;
  using Y = YCfg<P, Q, KIND>;
  constexpr bool MASS = Y::MASS;
  constexpr int NB1 = Y::NB1, NP = Y::NP, NB = Y::NB, NQ = Y::NQ, NT = Y::NT, CS = Y::CSS, TAB = C::TAB;
  constexpr int LDA2 = Y::LDA2, LDB2 = Y::LDB2, LDA3 = Y::LDA3, NSLOT = Y::NSLOT, NMT = Y::NMT;
  extern __shared__ __align__(16) double ssm[];
  double* const sT = ssm;
  double* const sW = sT + C::SZ_T;
  double* const sC = sW + C::SZ_W;        // per-point constants of element_prepare, then the mirror buffer
  double* const sMB = sC;
  double* const rX = sC + Y::SZ_CM;       // element_prepare's scratch (rA | sS2), then T1
  double* const sT1 = rX;
  double* const sM = rX + Y::SZ_X;
  double* const sT2 = sM + Y::SZ_M;
  double* const sB2 = sT2 + Y::SZ_T2;
  double* const sTw = sB2 + Y::SZ_B2;     // weighted tables (tensor-product weights)
  double* const sF = sTw + C::SZ_T;
  const bool useW = wt.t0 != nullptr;
  const double* const sTc = useW ? sTw : sT;  // the tables the contraction stages read
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lr = lane >> 2, lc = lane & 3;
  for (int idx = tid; idx < C::SZ_W; idx += NT) sW[idx] = 0.0;
  for (int idx = tid; idx < Y::SZ_B2; idx += NT) sB2[idx] = 0.0;  // padding rows stay zero
  // T2 is zeroed ONCE: afterwards it only ever holds finite numbers (the B3 build and stage 2 write products of table
  // entries), and whatever sits in a padding column meets a zero row of B3; rows are independent of each other
  for (int idx = tid; idx < Y::SZ_T2; idx += NT) sT2[idx] = 0.0;

  // ---- per-thread constants of the fragment <-> index maps (the same for every element and pass)
  // stage 2: this warp's tiles (tile row mt2, tiles nt2 .. nt2 + TPW - 1); C[lr][2 lc + h] of tile s is the entry
  // (m, n) = ((slot, q0), (j1, i1)) and goes to T2[n + NP slot][g Q + q0]
  const int tile0 = warp * Y::TPW;
  const bool has2 = tile0 < Y::MT2 * Y::NT2;
  const int mt2 = has2 ? tile0 / Y::NT2 : 0, nt2 = has2 ? tile0 % Y::NT2 : 0;
  const int m2 = 8 * mt2 + lr;
  int t2off[Y::TPW][2];
#pragma unroll
  for (int s = 0; s < Y::TPW; ++s)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = S::cperm(8 * (nt2 + s) + 2 * lc + h);
      t2off[s][h] = (has2 && m2 < Y::M2 && n < NP) ? (n + NP * (m2 / Q)) * LDA3 + (m2 % Q) : -1;
    }
  // stage 3: column position 8 nt + 2 lc + h of a tile holds the pair n = cperm(position) = 8 nt + lc + 4 h, (j0, i0) =
  // (n % NB1, n / NB1): the four lanes of a quad store four CONSECUTIVE entries of a row of K_e, and with the rows of the
  // tile (j1 = lr) a store instruction fills runs of 4 out of every 5 doubles instead of every other one (sectors written
  // per instruction: ~11 instead of ~25)
  //   bits 0-15 i0 NB + j0 (offset in K_e), 16-19 i0, 20-23 j0, 24-31 j0 NP + i0 (offset in the mirror buffer); -1 = padding
  const int hw3 = Y::SPLIT3 ? (warp & 1) : 0, ntBase3 = hw3 ? Y::NH0 : 0, ntCnt3 = hw3 ? Y::NH1 : Y::NH0;
  int c3[Y::NH0][2];
#pragma unroll
  for (int t = 0; t < Y::NH0; ++t)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = S::cperm(8 * (ntBase3 + t) + 2 * lc + h);
      const int i0 = n / NB1, j0 = n % NB1;
      c3[t][h] = (t < ntCnt3 && n < NP) ? (int)((unsigned)(i0 * NB + j0) | ((unsigned)i0 << 16) | ((unsigned)j0 << 20) |
                                                ((unsigned)(j0 * NP + i0) << 24))
                                        : -1;
    }
  // stage 3: row lr of this warp's v-th row tile is (j1, i1, slot): NB1 i1 | NB1 j1 << 8 | slot << 16, or -1
  constexpr int NMT3 = Y::NMT3;
  const int mtFirst3 = Y::SPLIT3 ? (warp >> 1) : warp, mtStep3 = Y::SPLIT3 ? 4 : Y::NWARP;
  int r3[NMT3];
#pragma unroll
  for (int u = 0; u < NMT3; ++u) {
    const int mp = 8 * (mtFirst3 + mtStep3 * u) + lr;
    const int slot = mp / NP, n1 = mp % NP;
    r3[u] = mp < Y::M3 ? (NB1 * (n1 / NB1)) | ((NB1 * (n1 % NB1)) << 8) | (slot << 16) : -1;
  }
  // stage 2, middle pass: tile (warp / NT2, warp % NT2)
  [[maybe_unused]] int t2alt[2] = {-1, -1};
  if constexpr (Y::ALT2) {
    const int mtA = warp / Y::NT2, ntA = warp % Y::NT2, m = 8 * mtA + lr;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = S::cperm(8 * ntA + 2 * lc + h);
      t2alt[h] = (mtA < Y::MTMID && m < Y::MIDROWS && n < NP) ? (n + NP * (m / Q)) * LDA3 + (m % Q) : -1;
    }
  }
  // stage 1 (tensor-core version, stage1_dmma in gram_common.cuh): where this lane's two accumulator columns go in T1
  int s1off[2];
  S1Cfg<P, Q, KIND>::store_offsets(warp, lane, NSLOT, LDA2, s1off);

  // mirror copy-out: entry NT k + tid of a (p+1)^2 x (p+1)^2 block is (row, c) -> offset row NB + c in K_e
  constexpr int NCO = (NP * NP + NT - 1) / NT;
  int cooff[NCO];
#pragma unroll
  for (int k = 0; k < NCO; ++k) cooff[k] = ((NT * k + tid) / NP) * NB + (NT * k + tid) % NP;

  // a block takes a CONTIGUOUS run of elements: consecutive elements share e1 (and mostly e2), so the stage-2 operand B2,
  // which depends on direction 1 only, is rebuilt once per row of elements instead of once per element
  SYM_PT_DECL
  const long long chunk = (nelem + gridDim.x - 1) / gridDim.x;
  const long long el_lo = blockIdx.x * chunk, el_hi = el_lo + chunk < nelem ? el_lo + chunk : nelem;
  int b2_e1 = -1;
  for (long long el = el_lo; el < el_hi; ++el) {
    const long long e = elem_begin + el;
    SYM_PT_BEGIN
    if (useW) {
      // nobody reads sTw any more: the previous element's last stage 1 lies three barriers back; element_prepare's
      // barriers order these writes before the operand builds below
      const int e0 = (int)(e % Pd.nel[0]), e1 = (int)((e / Pd.nel[0]) % Pd.nel[1]), e2 = (int)(e / ((long long)Pd.nel[0] * Pd.nel[1]));
      for (int j = tid; j < TAB; j += NT) {
        sTw[j] = __ldg(wt.t0 + (size_t)e0 * TAB + j);
        sTw[TAB + j] = __ldg(wt.t1 + (size_t)e1 * TAB + j);
        sTw[2 * TAB + j] = __ldg(wt.t2 + (size_t)e2 * TAB + j);
      }
    }
    // (the first barrier inside also ends the previous element's mirror copy-out, which reads the region sC aliases)
    if (!SKIP(6)) element_prepare<P, Q, NT, CS>(Pd, e, tab0, tab1, tab2, w0, w1, w2, sT, sW, rX, rX + C::SZ_A, sC);
    else __syncthreads();
    SYM_PT(0)
    // ---- T1: the padding columns of every group must be zero (they meet zero rows of B2, but the scratch T1 aliases may
    //      hold anything); the other columns of the rows in use are written by stage 1, and rows are independent
    for (int idx = tid; idx < Y::M2P * Y::NPADC; idx += NT) {
      const int i = idx % Y::NPADC, row = idx / Y::NPADC;
      int col = Y::KTOT + i, cnt = 0;   // columns KTOT .. LDA2 - 1 come last in the enumeration
#pragma unroll
      for (int g = 0; g < S::NG; ++g) {
        const int lo = S::goff(g) + S::ncombo(g) * Q, n = S::r4(S::ncombo(g) * Q) - S::ncombo(g) * Q;
        if (i >= cnt && i < cnt + n) col = lo + i - cnt;
        cnt += n;
      }
      if (i >= cnt) col = Y::KTOT + i - cnt;
      sT1[row * LDA2 + col] = 0.0;
    }
    // ---- B2[(g, c, q1)][pi1 = j1 + NB1 i1] = Phi1[q1][i1][b1(beta)] Phi1[q1][j1][b1(gamma)],  b1(beta) = (beta == 2)
    const int e1cur = (int)((e / Pd.nel[0]) % Pd.nel[1]);
    if (e1cur != b2_e1 && !SKIP(1))
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
    // ---- per point: M4 = C^T C (symmetric 4x4; mass: (sqrt(w detJ)/W)^2), and w detJ / W for the load vector
    for (int pt = tid; pt < NQ; pt += NT) {
      const double* o = sC + CS * pt;
      if (fe) sF[pt] = o[13] * o[0];
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
    __syncthreads();
    double b3[S::K3 / 4][Y::NH0];  // this warp's half of the column tiles (zero tile past the end of a shorter half)
#pragma unroll
    for (int ks = 0; ks < S::K3 / 4; ++ks)
#pragma unroll
      for (int t = 0; t < Y::NH0; ++t)
        b3[ks][t] = t < ntCnt3 ? sT2[(4 * ks + lc) * LDB2 + 8 * (ntBase3 + t) + lr] : 0.0;
    __syncthreads();

    double* const K = Ke + el * (long long)NB * NB;
    SYM_PT(1)
    for (int pass = 0; pass < Y::NPASS; ++pass) {
      const int i2b = NB1 - 1 - pass, na = NB1 - pass;   // side a: i2 = pass, na slots; side b: i2 = i2b, pass + 1 slots
      const bool paired = i2b > pass;
      const int nslots = paired ? NSLOT : na;
      const int rows2 = nslots * Q, rows3 = nslots * NP;
      if constexpr (Y::S1_DMMA) {
        // ---- stage 1 on the tensor cores: T1[(slot, q0)][(combo, q1)] = sum_q2 A_v[slot][q2] M4_combo[q2][(q1, q0)]
        const bool sv = lr < nslots, sa1 = lr < na;
        const int i2 = sv ? (sa1 ? pass : i2b) : 0, j2 = sv ? i2 + (sa1 ? lr : lr - na) : 0;
        if (!SKIP(0)) stage1_dmma<P, Q, KIND>(warp, lane, i2, j2, sv, s1off, sTc, sM, sT1);
      } else {
        // ---- stage 1: thread <-> (combo, q1, piece of the q0 range); per side the M4 line times the i2 factor in
        // registers, then all slots of that side: T1[(slot, q0)][(combo, q1)] = sum_q2 Phi2[q2][i2] Phi2[q2][j2] M4
        constexpr int QH = 2, NSPL = (Q + QH - 1) / QH;
        if (!SKIP(0))
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
          double* trow = sT1 + S::goff(g) + c * Q + q1;
#pragma unroll 1
          for (int side = 0; side < 2; ++side) {
            if (side && !paired) break;
            const int i2 = side ? i2b : pass, nj = side ? pass + 1 : na, slot0 = side ? na : 0;  // j2 = i2 + s
            double am[QH][Q];
#pragma unroll
            for (int q2 = 0; q2 < Q; ++q2) {
              const double a2 = sTc[2 * TAB + (q2 * NB1 + i2) * 2 + bb];
#pragma unroll
              for (int u = 0; u < QH; ++u) am[u][q2] = a2 * mrow[min(qa + u, Q - 1) + Q * Q * q2];
            }
            // all sums first, the T1 stores afterwards (a shared store pins the next iteration's loads behind it)
            double sums[NB1][QH];
#pragma unroll
            for (int s = 0; s < NB1; ++s) {
              if (s < nj) {
                double f[Q];
#pragma unroll
                for (int q2 = 0; q2 < Q; ++q2) f[q2] = sTc[2 * TAB + (q2 * NB1 + i2 + s) * 2 + bg];
#pragma unroll
                for (int u = 0; u < QH; ++u) {
                  double sum = 0.0;
#pragma unroll
                  for (int q2 = 0; q2 < Q; ++q2) sum = fma(am[u][q2], f[q2], sum);
                  sums[s][u] = sum;
                }
              }
            }
#pragma unroll
            for (int s = 0; s < NB1; ++s)
              if (s < nj) {
#pragma unroll
                for (int u = 0; u < QH; ++u)
                  if (qa + u < Q) trow[((slot0 + s) * Q + qa + u) * LDA2] = sums[s][u];
              }
          }
        }
      }
      SYM_PT(2)
      __syncthreads();
      SYM_PT(6)
      // ---- stage 2: C_g[(slot,q0)][pi1] = sum_{k in g} T1[.][k] B2[k][.]  ->  T2[pi1 + NP slot][g Q + q0]
      if constexpr (Y::GEN2) {
        for (int tile = warp; tile < Y::MT2 * Y::NT2; tile += Y::NWARP) {
          const int mt = tile / Y::NT2, nt = tile % Y::NT2;
          if (8 * mt >= rows2) break;
          const double* arow = sT1 + (8 * mt + lr) * LDA2 + lc;
          const double* brow = sB2 + lc * LDB2 + 8 * nt + lr;
          const int m = 8 * mt + lr;
          int off[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int n = S::cperm(8 * nt + 2 * lc + h);
            off[h] = (m < rows2 && n < NP) ? (n + NP * (m / Q)) * LDA3 + (m % Q) : -1;
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
      } else if (Y::ALT2 && !paired) {
        // middle pass: half the slots, one tile per warp
        const int mtA = warp / Y::NT2, ntA = warp % Y::NT2;
        if (mtA < Y::MTMID && !SKIP(4)) {
          const double* arow = sT1 + (8 * mtA + lr) * LDA2 + lc;
          const double* brow = sB2 + lc * LDB2 + 8 * ntA + lr;
          double acc[S::NG][2];
#pragma unroll
          for (int g = 0; g < S::NG; ++g) {
            acc[g][0] = acc[g][1] = 0.0;
            const int kb = S::goff(g), nks = S::r4(S::ncombo(g) * Q) / 4;
#pragma unroll
            for (int ks = 0; ks < nks; ++ks) dmma8x8x4_f(acc[g][0], acc[g][1], arow[kb + 4 * ks], brow[(kb + 4 * ks) * LDB2]);
          }
#pragma unroll
          for (int g = 0; g < S::NG; ++g)
#pragma unroll
            for (int h = 0; h < 2; ++h)
              if (t2alt[h] >= 0) sT2[t2alt[h] + g * Q] = acc[g][h];
        }
      } else if (has2 && 8 * mt2 < rows2 && !SKIP(4)) {
        const double* arow = sT1 + (8 * mt2 + lr) * LDA2 + lc;
        const double* brow = sB2 + lc * LDB2 + 8 * nt2 + lr;
        // all groups' accumulators first, the T2 stores afterwards: NG x TPW independent DMMA chains interleave
        double acc[S::NG][Y::TPW][2];
#pragma unroll
        for (int g = 0; g < S::NG; ++g) {
#pragma unroll
          for (int s = 0; s < Y::TPW; ++s) acc[g][s][0] = acc[g][s][1] = 0.0;
          const int kb = S::goff(g), nks = S::r4(S::ncombo(g) * Q) / 4;  // constants after unrolling over g
#pragma unroll
          for (int ks = 0; ks < nks; ++ks) {
            const double a = arow[kb + 4 * ks];
#pragma unroll
            for (int s = 0; s < Y::TPW; ++s) dmma8x8x4_f(acc[g][s][0], acc[g][s][1], a, brow[(kb + 4 * ks) * LDB2 + 8 * s]);
          }
        }
        if (m2 < rows2) {
#pragma unroll
          for (int g = 0; g < S::NG; ++g)
#pragma unroll
            for (int s = 0; s < Y::TPW; ++s)
#pragma unroll
              for (int h = 0; h < 2; ++h)
                if (t2off[s][h] >= 0) sT2[t2off[s][h] + g * Q] = acc[g][s][h];
        }
      }
      SYM_PT(3)
      __syncthreads();
      SYM_PT(6)
      // ---- stage 3: K[(j1,i1,slot)][pi0] = sum_{(g,q0)} T2 B3: to global memory (with the weights w_i w_j on the general
      //      path), and the strictly upper blocks (j2 > i2) also into the mirror buffer, transposed
#pragma unroll
      for (int u = 0; u < NMT3; ++u) {
        const int mt = mtFirst3 + mtStep3 * u;
        if (mt >= Y::MT3 || 8 * mt >= rows3) break;
        double acc[Y::NH0][2];
#pragma unroll
        for (int t = 0; t < Y::NH0; ++t) acc[t][0] = acc[t][1] = 0.0;
        const double* arow = sT2 + (8 * mt + lr) * LDA3 + lc;
        if (!SKIP(5))
#pragma unroll
        for (int ks = 0; ks < S::K3 / 4; ++ks) {
          const double a = arow[4 * ks];
#pragma unroll
          for (int t = 0; t < Y::NH0; ++t) dmma8x8x4_f(acc[t][0], acc[t][1], a, b3[ks][t]);
        }
        const int slot = r3[u] >> 16;
        if (r3[u] >= 0 && slot < nslots) {
          const bool sa = slot < na;
          const int i2 = sa ? pass : i2b, j2 = i2 + (sa ? slot : slot - na);
          const int i1b = r3[u] & 255, j1b = (r3[u] >> 8) & 255;
          const int ib = i1b + NP * i2, jb = j1b + NP * j2;
          double* const Kb = K + ib * NB + jb;
          if (!useW) {
            const double* const wI = sW + ib;
            const double* const wJ = sW + jb;
#pragma unroll
            for (int nt = 0; nt < Y::NH0; ++nt)
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int pk = c3[nt][h];
                if (pk != -1) acc[nt][h] *= wI[(pk >> 16) & 15] * wJ[(pk >> 20) & 15];
              }
          }
          if (!SKIP(3))
#pragma unroll
          for (int nt = 0; nt < Y::NH0; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int pk = c3[nt][h];
              if (pk != -1) Kb[pk & 0xffff] = acc[nt][h];
            }
          // (the transposed entries straight from the registers -- scattered 8-byte stores, merged in L2 -- were measured:
          //  0.94x at p = 4 Poisson, 0.74x for the mass matrix; profiles/README.md)
          if (j2 > i2 && !SKIP(2)) {
            double* const Mb = sMB + (sa ? slot - 1 : slot - 2) * (NP * NP) + j1b * NP + i1b;
#pragma unroll
            for (int nt = 0; nt < Y::NH0; ++nt)
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int pk = c3[nt][h];
                if (pk != -1) Mb[(unsigned)pk >> 24] = acc[nt][h];
              }
          }
        }
      }
      SYM_PT(4)
      __syncthreads();
      SYM_PT(6)
      // ---- mirror: block (j2, i2) = block (i2, j2)^T, complete runs of (p+1)^2 doubles per row of K_e.  No barrier after
      //      it: the next pass writes the buffer in its stage 3, two barriers away (T1 and T2 are not touched here)
      {
        const int nm = SKIP(2) ? 0 : (paired ? NSLOT - 2 : na - 1);
        // block by block (warp-uniform i2, j2); inside a block the thread's entries sit at offsets fixed for the whole kernel
        for (int ms = 0; ms < nm; ++ms) {
          const bool sa = ms < na - 1;
          const int i2 = sa ? pass : i2b, j2 = i2 + 1 + (sa ? ms : ms - (na - 1));
          double* const Kd = K + (NP * j2) * NB + NP * i2;
          const double* const src = sMB + ms * (NP * NP);
#pragma unroll
          for (int k = 0; k < NCO; ++k)
            if (NT * k + tid < NP * NP) Kd[cooff[k]] = src[NT * k + tid];
        }
      }
      SYM_PT(5)
    }
    // ---- load vector f_e[i] = sum_q w detJ R_i fconst  (poisson.py:79)
    if (fe) {
      // sum-factorised: f_i = w_i sum_q0 N0[q0][i0] sum_q1 N1[q1][i1] sum_q2 N2[q2][i2] (w detJ / W)(q) fconst, one direction
      // per step; scratch in the T1 region, which stage 2 of the last pass was the last to read
      double* const fA = sT1;                // [q0 + Q q1][i2]
      double* const fB = sT1 + Q * Q * NB1;  // [q0][i1 + NB1 i2]
      for (int idx = tid; idx < Q * Q * NB1; idx += NT) {
        const int q01 = idx % (Q * Q), i2 = idx / (Q * Q);
        double a = 0.0;
#pragma unroll
        for (int q2 = 0; q2 < Q; ++q2) a = fma(sT[2 * TAB + (q2 * NB1 + i2) * 2], sF[q01 + Q * Q * q2], a);
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
        fe[el * NB + i] = f;
      }
    }
    SYM_PT(7)
  }
  SYM_PT_FLUSH
}

template <int P, int Q, int KIND>
igab_status launch_sym_t(const PatchDev& Pd, SfWorkspace& ws, double fconst, long long eb, long long nel,
                         const double* const* w_dev, double* Ke, double* fe, cudaStream_t stream) {
  using Y = YCfg<P, Q, KIND>;
  if constexpr (!Y::FITS) {
    return IGAB_UNSUPPORTED;
  } else {
    const size_t sm = (size_t)Y::SZ_TOTAL * sizeof(double);
    KernelCfg kc;
    if (igab_status cst = kernel_config((const void*)gram_sym_kernel<P, Q, KIND>, Y::NT, sm, &kc)) return cst;
    const unsigned grid = (unsigned)std::min<long long>(nel, (long long)kc.blocksPerSm * kc.sms);
#ifdef IGAB_SYM_EXPERIMENT
    {
      const char* sk = getenv("IGAB_SYM_SKIP");
      const int v = sk ? atoi(sk) : 0;
      cudaMemcpyToSymbolAsync(c_skip, &v, sizeof(int), 0, cudaMemcpyHostToDevice, stream);
    }
#endif
    gram_sym_kernel<P, Q, KIND><<<grid, Y::NT, sm, stream>>>(Pd, eb, nel, ws.tab[0], ws.tab[1], ws.tab[2], w_dev[0],
                                                            w_dev[1], w_dev[2], fconst, Ke, fe, wtabs_of(Pd, ws));
    count_launch();
    IGAB_CUDA_TRY(cudaGetLastError());
    return IGAB_OK;
  }
}

template <int P, int Q>
constexpr bool sym_fits(bool mass) {
  return mass ? YCfg<P, Q, 1>::FITS : YCfg<P, Q, 0>::FITS;
}

}  // namespace

#define IGAB_SYM_LIST(X) X(3, 4) X(3, 5) X(4, 5) X(4, 6) X(4, 7) X(5, 6) X(5, 7)

// Where the symmetric kernel is the default (measured on B200 against gram_sf_kernel, tools/sym_bench.py, element matrices/s):
// Poisson p = 4: 1.32x (5 points), 1.23x (7 points); p = 3 with 5 points 1.17x; p = 5 1.31x.  Elsewhere the contraction is too
// small a share of the element's time to pay for the mirror pass (Poisson p = 3 / 4 points 0.89x, p = 4 / 6 points 0.99x;
// mass 0.74-0.92x) and gram_sf_kernel stays.  IGAB_GRAM_SYM=1 forces it wherever it is compiled.
bool gram_sym_default(int p, int q, int kind) {
  if (kind != IGAB_ASM_POISSON) return false;
  return (p == 4 && (q == 5 || q == 7)) || (p == 3 && q == 5) || (p == 5 && q == 6);
}

bool gram_sym_supported(int p, int q, int kind) {
  if (kind != IGAB_ASM_POISSON && kind != IGAB_ASM_MASS) return false;
  const bool mass = kind == IGAB_ASM_MASS;
#define IGAB_SYM_CASE(PP, QQ) \
  if (p == PP && q == QQ) return sym_fits<PP, QQ>(mass);
  IGAB_SYM_LIST(IGAB_SYM_CASE)
#undef IGAB_SYM_CASE
  return false;
}

igab_status launch_gram_sym(const PatchDev& Pd, SfWorkspace& ws, int kind, int q, double fconst, long long eb,
                            long long nel, const double* const* w_dev, double* Ke, double* fe, cudaStream_t stream) {
  const int p = Pd.degree[0];
  const bool mass = kind == IGAB_ASM_MASS;
#define IGAB_SYM_CASE(PP, QQ)                                                                         \
  if (p == PP && q == QQ)                                                                             \
    return mass ? launch_sym_t<PP, QQ, 1>(Pd, ws, fconst, eb, nel, w_dev, Ke, fe, stream)             \
                : launch_sym_t<PP, QQ, 0>(Pd, ws, fconst, eb, nel, w_dev, Ke, fe, stream);
  IGAB_SYM_LIST(IGAB_SYM_CASE)
#undef IGAB_SYM_CASE
  set_error("assemble: no symmetric sum-factorised kernel for this configuration");
  return IGAB_UNSUPPORTED;
}

}  // namespace igab

#ifdef IGAB_PHASE_TIMING
extern "C" int igab_debug_phase_ticks_sym(unsigned long long* out, int reset) {
  unsigned long long z[16] = {0};
  if (out && cudaMemcpyFromSymbol(out, igab::g_phase_ticks, sizeof(z)) != cudaSuccess) return 1;
  if (reset && cudaMemcpyToSymbol(igab::g_phase_ticks, z, sizeof(z)) != cudaSuccess) return 1;
  return 0;
}
#endif

// assemble.cu -- element matrices K_e = sum_q w detJ B^T D B (K5) and patch reductions (K6).
// Functional spec: the Python loop dune/python/test/poisson.py:37-81 (localAssembler), generalised to a B-operator
// with `nrow` rows per quadrature point:
//     Poisson     B = J^-T grad_xi R_i              (dimworld rows)   poisson.py:58,64-77
//     mass        B = R_i                            (1 row)
//     elasticity  B = Voigt strain operator          (3 | 6 rows), D from the caller, DOFs interleaved i*dim+c
// Written as a batched FP64 contraction  K_e = A1_e^T A2_e  with A1 = B-stack [(q,row)][n], A2 = w detJ D B-stack.
// Round-1 structure: (1) evaluation kernel (tensor sum-factorised or general) -> N, dN, detJ, JinvT in stream-ordered
// scratch, (2) B-stack kernel, (3) contraction kernel: FP64 tensor-core DMMA (mma.sync.m8n8k4.f64) for n >= 32,
// shared-memory FMA tiles below that.
#include <cstdlib>

#include "igab_internal.cuh"

namespace igab {

namespace {

// ---- (2) B-stacks ---------------------------------------------------------------------------------------------------
// one thread per (point, basis function)
template <int DIM, int DW>
__global__ void bstack_kernel(int kind, long long npts, int nb, long long nq, int nq0, int nq1_, int nq2,
                              const double* __restrict__ w0, const double* __restrict__ w1, const double* __restrict__ w2,
                              const double* __restrict__ N, const double* __restrict__ dN,
                              const double* __restrict__ detJ, const double* __restrict__ JinvT,
                              const double* __restrict__ Dm, double* __restrict__ A1, double* __restrict__ A2, int nrow,
                              int n) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npts * nb) return;
  const long long pt = t / nb;
  const int i = (int)(t % nb);
  long long q = pt % nq;
  const long long e = pt / nq;
  double w = w0[q % nq0];
  q /= nq0;
  if (DIM > 1) { w *= w1[q % nq1_]; q /= nq1_; }
  if (DIM > 2) { w *= w2[q % nq2]; }
  const double s = w * detJ[pt];
  // rows of this point inside the element's stack: [(q,row)][n]
  const long long rowBase = (e * nq + (pt % nq)) * nrow;
  if (kind == IGAB_ASM_MASS) {
    const double v = N[pt * nb + i];
    A1[rowBase * n + i] = v;
    A2[rowBase * n + i] = s * v;
    return;
  }
  double g[DW];
#pragma unroll
  for (int c = 0; c < DW; ++c) {
    double acc = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) acc += JinvT[pt * DIM * DW + c * DIM + d] * dN[(pt * nb + i) * DIM + d];
    g[c] = acc;
  }
  if (kind == IGAB_ASM_POISSON) {
#pragma unroll
    for (int c = 0; c < DW; ++c) {
      A1[(rowBase + c) * n + i] = g[c];
      A2[(rowBase + c) * n + i] = s * g[c];
    }
    return;
  }
  // elasticity: B columns i*DIM + c
  constexpr int NS = (DIM == 2) ? 3 : 6;
  double Bc[NS][DIM];
#pragma unroll
  for (int r = 0; r < NS; ++r)
#pragma unroll
    for (int c = 0; c < DIM; ++c) Bc[r][c] = 0.0;
  if constexpr (DIM == 2) {
    Bc[0][0] = g[0]; Bc[1][1] = g[1]; Bc[2][0] = g[1]; Bc[2][1] = g[0];
  } else if constexpr (DIM == 3) {
    Bc[0][0] = g[0]; Bc[1][1] = g[1]; Bc[2][2] = g[2];
    Bc[3][1] = g[2]; Bc[3][2] = g[1];
    Bc[4][0] = g[2]; Bc[4][2] = g[0];
    Bc[5][0] = g[1]; Bc[5][1] = g[0];
  }
#pragma unroll
  for (int r = 0; r < NS; ++r)
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
      A1[(rowBase + r) * n + i * DIM + c] = Bc[r][c];
      double acc = 0.0;
#pragma unroll
      for (int t2 = 0; t2 < NS; ++t2) acc += Dm[r * NS + t2] * Bc[t2][c];
      A2[(rowBase + r) * n + i * DIM + c] = s * acc;
    }
}

// f_e[i] = sum_q w detJ N_i fconst  (poisson.py:79); one thread per (element, dof)
template <int DIM>
__global__ void load_kernel(long long nel, int nb, int ncomp, long long nq, int nq0, int nq1_, int nq2,
                            const double* __restrict__ w0, const double* __restrict__ w1, const double* __restrict__ w2,
                            const double* __restrict__ N, const double* __restrict__ detJ, double fconst,
                            double* __restrict__ fe) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nel * nb) return;
  const long long e = t / nb;
  const int i = (int)(t % nb);
  double acc = 0.0;
  for (long long q = 0; q < nq; ++q) {
    long long h = q;
    double w = w0[h % nq0];
    h /= nq0;
    if (DIM > 1) { w *= w1[h % nq1_]; h /= nq1_; }
    if (DIM > 2) { w *= w2[h % nq2]; }
    const long long pt = e * nq + q;
    acc += w * detJ[pt] * N[pt * nb + i] * fconst;
  }
  for (int c = 0; c < ncomp; ++c) fe[(e * nb + i) * ncomp + c] = acc;
}

// ---- (3a) contraction, FMA path: K_e = A1^T A2, 16x16 thread tile, 2x2 outputs per thread ---------------------------
__global__ void __launch_bounds__(256) contract_fma_kernel(int n, int kdim, const double* __restrict__ A1,
                                                           const double* __restrict__ A2, double* __restrict__ Ke) {
  constexpr int T = 32, KT = 16;
  __shared__ double sa[KT][T + 1], sb[KT][T + 1];
  const long long e = blockIdx.z;
  const double* a1 = A1 + e * (long long)kdim * n;
  const double* a2 = A2 + e * (long long)kdim * n;
  const int i0 = blockIdx.y * T, j0 = blockIdx.x * T;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[2][2] = {{0, 0}, {0, 0}};
  for (int k0 = 0; k0 < kdim; k0 += KT) {
    for (int idx = threadIdx.x; idx < KT * T; idx += 256) {
      const int kk = idx / T, c = idx % T;
      const int k = k0 + kk;
      sa[kk][c] = (k < kdim && i0 + c < n) ? a1[(long long)k * n + i0 + c] : 0.0;
      sb[kk][c] = (k < kdim && j0 + c < n) ? a2[(long long)k * n + j0 + c] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      const double x0 = sa[kk][ty], x1 = sa[kk][ty + 16], y0 = sb[kk][tx], y1 = sb[kk][tx + 16];
      acc[0][0] += x0 * y0; acc[0][1] += x0 * y1; acc[1][0] += x1 * y0; acc[1][1] += x1 * y1;
    }
    __syncthreads();
  }
  double* K = Ke + e * (long long)n * n;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
      if (i < n && j < n) K[(long long)i * n + j] = acc[a][b];
    }
}

// ---- (3b) contraction, FP64 tensor cores: mma.sync.aligned.m8n8k4.row.col.f64 ---------------------------------------
// Block = 4 warps computes a 64x64 tile of K_e; each warp a 32x32 sub-tile = 4x4 DMMA tiles (2 doubles per lane each).
// C[i][j] = sum_k A1[k][i] A2[k][j]:  A-operand (row-major M x K) element (i,k) = A1[k][i]; B-operand (K x N) = A2[k][j].
// Fragment layout of m8n8k4.f64: lane = 4*g + t (g = 0..7, t = 0..3): a = A[g][t], b = B[t][g], c0 = C[g][2t], c1 = C[g][2t+1].
__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(128) contract_dmma_kernel(int n, int kdim, const double* __restrict__ A1,
                                                            const double* __restrict__ A2, double* __restrict__ Ke) {
  constexpr int T = 64, KT = 16, LD = T + 4;  // row stride 68 doubles: the 4 k-rows of a fragment land in 4 distinct bank groups
  __shared__ double sa[KT][LD], sb[KT][LD];
  const long long e = blockIdx.z;
  const double* a1 = A1 + e * (long long)kdim * n;
  const double* a2 = A2 + e * (long long)kdim * n;
  const int i0 = blockIdx.y * T, j0 = blockIdx.x * T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wi = (warp >> 1) * 32, wj = (warp & 1) * 32;
  double c[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) c[a][b][0] = c[a][b][1] = 0.0;
  for (int k0 = 0; k0 < kdim; k0 += KT) {
    for (int idx = threadIdx.x; idx < KT * T; idx += 128) {
      const int kk = idx / T, col = idx % T;
      const int k = k0 + kk;
      sa[kk][col] = (k < kdim && i0 + col < n) ? a1[(long long)k * n + i0 + col] : 0.0;
      sb[kk][col] = (k < kdim && j0 + col < n) ? a2[(long long)k * n + j0 + col] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < KT; ks += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) af[a] = sa[ks + t][wi + 8 * a + g];
#pragma unroll
      for (int b = 0; b < 4; ++b) bf[b] = sb[ks + t][wj + 8 * b + g];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dmma8x8x4(c[a][b][0], c[a][b][1], af[a], bf[b]);
    }
    __syncthreads();
  }
  double* K = Ke + e * (long long)n * n;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = i0 + wi + 8 * a + g, j = j0 + wj + 8 * b + 2 * t;
      if (i < n) {
        if (j < n) K[(long long)i * n + j] = c[a][b][0];
        if (j + 1 < n) K[(long long)i * n + j + 1] = c[a][b][1];
      }
    }
}

// ---- (3c) fused Gram kernel for 3-D Poisson / mass at nb >= 33 (p >= 3): no B-stack in HBM ---------------------------
// K_e[i][j] = sum_q w_q detJ_q sum_c G_q[c][i] G_q[c][j],  G_q = J_q^-T grad_xi R(q)   (poisson.py:58,64-77)
// One block per element (persistent, grid-stride).  The quadrature points are consumed in chunks of 16 k-rows
// (5 points x 3 gradient components + 1 zero row, or 16 points for the mass matrix).  Software pipeline, ONE
// __syncthreads per chunk:   phase c = { issue the global loads of chunk c+1 (dN rows, coalesced) into registers;
//                                        contract chunk c on the FP64 tensor cores (mma.sync.m8n8k4.f64, SASS DMMA)
//                                        from shared buffer c&1 into register accumulators;
//                                        turn the prefetched rows into physical gradients with J^-T and store them
//                                        into buffer (c+1)&1;  stage J^-T / w detJ of chunk c+2 }.
// The per-row scale w detJ is applied to the B-operand fragment, so only ONE stack lives in shared memory.
// Each warp owns a 32 x (NBP/WARPS_N) block of K_e.
template <int NBP, int NROW, int MINB>
__global__ void __launch_bounds__((NBP / 32) * (NBP / 32 + 1) / 2 * 32, MINB) gram_dmma_kernel(
    int nb, int nq, long long nelem, int nq0, int nq1_, int nq2, const double* __restrict__ w0,
    const double* __restrict__ w1, const double* __restrict__ w2, const double* __restrict__ N,
    const double* __restrict__ dN, const double* __restrict__ detJ, const double* __restrict__ JinvT,
    double* __restrict__ Ke) {
  // K_e is symmetric: only the NBLK (NBLK+1)/2 upper 32 x 32 blocks are contracted, one warp each; the strictly upper
  // blocks are mirrored on the way out.
  constexpr int NBLK = NBP / 32, NUP = NBLK * (NBLK + 1) / 2;
  constexpr int KC = 16, LD = NBP + 4, NT = NUP * 32;
  constexpr int TM = 4, TN = 4;  // 8x8 tiles per warp
  constexpr int nrow = NROW, NV = NROW == 3 ? 3 : 1;
  constexpr int MAXIT = ((KC / NROW) * NBP + NT - 1) / NT;
  __shared__ double sA[2][KC][LD];
  __shared__ double sS[3][KC];
  __shared__ double sJ[3][KC][9];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  int bi = 0, bj = 0;
  {
    int w = warp;  // row-major enumeration of the upper blocks: (0,0) (0,1) .. (0,NBLK-1) (1,1) ..
    while (w >= NBLK - bi) {
      w -= NBLK - bi;
      ++bi;
    }
    bj = bi + w;
  }
  const int wi = bi * 32, wj = bj * 32;
  constexpr int pts = KC / NROW;  // points per chunk (5 | 16)
  const int nchunk = (nq + pts - 1) / pts;
  constexpr int nitem = pts * NBP;
  for (int idx = threadIdx.x; idx < 2 * KC * LD; idx += NT) (&sA[0][0][0])[idx] = 0.0;

  // stage J^-T and w detJ of chunk ch of element e into ring slot ch % 3
  auto stage_consts = [&](long long e, int ch) {
    const int slot = ch % 3, q0c = ch * pts, npt = min(pts, nq - q0c);
    if (nrow == 3 && threadIdx.x < pts * 9) {
      const int ql = threadIdx.x / 9, k = threadIdx.x % 9;
      sJ[slot][ql][k] = ql < npt ? JinvT[(e * nq + q0c + ql) * 9 + k] : 0.0;
    } else if (threadIdx.x >= 64 && threadIdx.x < 64 + KC) {
      const int k = threadIdx.x - 64, ql = k / nrow;
      double sc = 0.0;
      if (ql < npt && k < pts * nrow) {
        int q = q0c + ql;
        double w = w0[q % nq0];
        q /= nq0;
        w *= w1[q % nq1_];
        q /= nq1_;
        w *= w2[q % nq2];
        sc = w * detJ[e * nq + q0c + ql];
      }
      sS[slot][k] = sc;
    }
  };
  // global loads of one chunk into registers (3 doubles per item for gradients, 1 for the mass matrix)
  auto prefetch = [&](long long e, int ch, double (&r)[MAXIT][NV]) {
    const int q0c = ch * pts, npt = min(pts, nq - q0c);
#pragma unroll
    for (int m = 0; m < MAXIT; ++m) {
      const int item = threadIdx.x + m * NT;
      const int ql = item / NBP, i = item % NBP;
      const bool live = item < nitem && ql < npt && i < nb;
      const long long pt = e * nq + q0c + ql;
      if constexpr (NROW == 3) {
        const double* d = dN + (pt * nb + i) * 3;
        r[m][0] = live ? d[0] : 0.0;
        r[m][1] = live ? d[1] : 0.0;
        r[m][2] = live ? d[2] : 0.0;
      } else {
        r[m][0] = live ? N[pt * nb + i] : 0.0;
      }
    }
  };
  auto build = [&](int ch, const double (&r)[MAXIT][NV]) {
    const int buf = ch & 1, slot = ch % 3;
#pragma unroll
    for (int m = 0; m < MAXIT; ++m) {
      const int item = threadIdx.x + m * NT;
      if (item < nitem) {
        const int ql = item / NBP, i = item % NBP;
        if constexpr (NROW == 3) {
          const double* J = sJ[slot][ql];
          sA[buf][3 * ql + 0][i] = J[0] * r[m][0] + J[1] * r[m][1] + J[2] * r[m][2];
          sA[buf][3 * ql + 1][i] = J[3] * r[m][0] + J[4] * r[m][1] + J[5] * r[m][2];
          sA[buf][3 * ql + 2][i] = J[6] * r[m][0] + J[7] * r[m][1] + J[8] * r[m][2];
        } else {
          sA[buf][ql][i] = r[m][0];
        }
      }
    }
  };

  for (long long e = blockIdx.x; e < nelem; e += gridDim.x) {
    double c[TM][TN][2];
#pragma unroll
    for (int a = 0; a < TM; ++a)
#pragma unroll
      for (int b = 0; b < TN; ++b) c[a][b][0] = c[a][b][1] = 0.0;
    double r[MAXIT][NV];
    // prologue: constants of chunks 0, 1; stack of chunk 0
    stage_consts(e, 0);
    __syncthreads();  // also orders the previous element's last DMMA reads before buffers are rewritten
    if (nchunk > 1) stage_consts(e, 1);
    prefetch(e, 0, r);
    build(0, r);
    __syncthreads();
    for (int ch = 0; ch < nchunk; ++ch) {
      const bool more = ch + 1 < nchunk;
      if (more) prefetch(e, ch + 1, r);
      if (ch + 2 < nchunk) stage_consts(e, ch + 2);
      const int buf = ch & 1, slot = ch % 3;
#pragma unroll
      for (int ks = 0; ks < KC; ks += 4) {
        double af[TM], bf[TN];
        const double sc = sS[slot][ks + t];
#pragma unroll
        for (int a = 0; a < TM; ++a) af[a] = sA[buf][ks + t][wi + 8 * a + g];
#pragma unroll
        for (int b = 0; b < TN; ++b) bf[b] = sA[buf][ks + t][wj + 8 * b + g] * sc;
#pragma unroll
        for (int a = 0; a < TM; ++a)
#pragma unroll
          for (int b = 0; b < TN; ++b) dmma8x8x4(c[a][b][0], c[a][b][1], af[a], bf[b]);
      }
      if (more) build(ch + 1, r);
      __syncthreads();
    }
    double* K = Ke + e * (long long)nb * nb;
#pragma unroll
    for (int a = 0; a < TM; ++a)
#pragma unroll
      for (int b = 0; b < TN; ++b) {
        const int i = wi + 8 * a + g, j = wj + 8 * b + 2 * t;
        if (i < nb) {
          if (j < nb) K[(long long)i * nb + j] = c[a][b][0];
          if (j + 1 < nb) K[(long long)i * nb + j + 1] = c[a][b][1];
          if (bi != bj) {  // mirror
            if (j < nb) K[(long long)j * nb + i] = c[a][b][0];
            if (j + 1 < nb) K[(long long)(j + 1) * nb + i] = c[a][b][1];
          }
        }
      }
  }
}

// ---- K6: deterministic volume reduction -----------------------------------------------------------------------------
// partial[b] = sum over a fixed contiguous slice of points of detJ*w (sequential per thread, fixed tree per block);
// result = fixed-order sum of the partials.  Same launch geometry => bitwise reproducible.
template <int DIM>
__global__ void __launch_bounds__(256) volume_partial_kernel(long long npts, long long nq, int nq0, int nq1_, int nq2,
                                                             const double* __restrict__ w0, const double* __restrict__ w1,
                                                             const double* __restrict__ w2,
                                                             const double* __restrict__ detJ, double* __restrict__ partial,
                                                             const uint8_t* __restrict__ trim, long long elem_begin) {
  __shared__ double sh[256];
  const long long per = (npts + gridDim.x - 1) / gridDim.x;
  const long long b0 = (long long)blockIdx.x * per, b1 = min(npts, b0 + per);
  double acc = 0.0;
  for (long long pt = b0 + threadIdx.x; pt < b1; pt += 256) {
    // trimmed patches: only full elements carry the tensor rule (trimmed ones bring their own points, empty ones nothing)
    if (trim && trim[elem_begin + pt / nq] != IGAB_TRIM_FULL) continue;
    long long q = pt % nq;
    double w = w0[q % nq0];
    q /= nq0;
    if (DIM > 1) { w *= w1[q % nq1_]; q /= nq1_; }
    if (DIM > 2) { w *= w2[q % nq2]; }
    acc += detJ[pt] * w;
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(1024) volume_final_kernel(int n, const double* __restrict__ partial, double* __restrict__ result) {
  __shared__ double sh[1024];
  sh[threadIdx.x] = threadIdx.x < n ? partial[threadIdx.x] : 0.0;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *result = sh[0];
}

igab_status eval_for(const PatchDev& P, SfWorkspace& ws, const double* const* xi1d_host, long long eb, long long nel,
                     const int* nq1, const double* const* xi1d_dev, const EvalOutDev& out, cudaStream_t stream) {
  PointSource src{};
  src.tensor = 1;
  src.elem_begin = eb;
  src.nq = 1;
  for (int d = 0; d < P.dim; ++d) {
    src.nq1[d] = nq1[d];
    src.nq *= nq1[d];
    src.xi1d[d] = xi1d_dev[d];
  }
  if (eval_tensor_blk_supported(P, nq1, 1, out)) return launch_eval_tensor_blk(P, ws, src, xi1d_host, nel, 1, out, stream);
  if (eval_tensor_sf_supported(P, nq1, 1, out)) return launch_eval_tensor_sf(P, ws, src, xi1d_host, nel, 1, out, stream);
  if (eval_tensor_2d_supported(P, nq1, 1, out)) return launch_eval_tensor_2d(P, ws, src, xi1d_host, nel, 1, out, stream);
  {
    const igab_status st = launch_eval_warp(P, src, nel, 1, out, stream);  // general warp-per-element kernel
    if (st != IGAB_UNSUPPORTED) return st;
  }
  return launch_eval_generic(P, src, nel * src.nq, 1, out, stream);
}

template <int DIM, int DW>
igab_status assemble_dd(const PatchDev& P, SfWorkspace& ws, const double* const* xi1d_host, int kind, const double* D_host, const double* D_dev, double fconst, long long eb, long long nel,
                        const int* nq1, const double* const* xi, const double* const* w, double* Ke, double* fe,
                        cudaStream_t stream) {
  if (!getenv("IGAB_NO_FUSED_ASSEMBLY") && gram_fused_supported(P, kind, nq1))
    return launch_gram_fused(P, ws, xi1d_host, xi, w, kind, fconst, eb, nel, nq1, Ke, fe, stream, D_host);
  if (!getenv("IGAB_NO_FUSED_ASSEMBLY") && assemble2d_supported(P, kind, nq1))
    return launch_assemble2d(P, ws, xi1d_host, kind, D_host, fconst, eb, nel, nq1, xi, w, Ke, fe, stream);
  long long nq = 1;
  for (int d = 0; d < DIM; ++d) nq *= nq1[d];
  const int nb = P.nb;
  const int ncomp = kind == IGAB_ASM_ELASTICITY ? DIM : 1;
  const int n = nb * ncomp;
  const int nrow = kind == IGAB_ASM_MASS ? 1 : (kind == IGAB_ASM_POISSON ? DW : (DIM == 2 ? 3 : 6));
  const int kdim = (int)nq * nrow;
  // element chunks bounded by scratch (A1 + A2 dominate)
  // fused Gram path: 3-D Poisson / mass with a local basis of 33..128 functions (p = 3, 4): no B-stack in HBM
  const bool gram = DIM == 3 && DW == 3 && kind != IGAB_ASM_ELASTICITY && nb > 32 && nb <= 128;
  const long long perEl = 8LL * ((gram ? 0LL : 2LL * kdim * n) + nq * (nb + nb * DIM + 1 + DIM * DW));
  // the contraction kernels index elements with gridDim.z (<= 65535)
  const long long chunk = std::max(1LL, std::min(std::min(nel, (4LL << 30) / perEl), gram ? nel : 65535LL));
  const int q0 = nq1[0], q1 = DIM > 1 ? nq1[1] : 1, q2 = DIM > 2 ? nq1[2] : 1;
  for (long long c0 = 0; c0 < nel; c0 += chunk) {
    const long long m = std::min(chunk, nel - c0);
    const long long npts = m * nq;
    double *dNv = nullptr, *ddN = nullptr, *ddet = nullptr, *dJi = nullptr, *A1 = nullptr, *A2 = nullptr;
    IGAB_CUDA_TRY(cudaMallocAsync(&dNv, 8 * npts * nb, stream));
    IGAB_CUDA_TRY(cudaMallocAsync(&ddN, 8 * npts * nb * DIM, stream));
    IGAB_CUDA_TRY(cudaMallocAsync(&ddet, 8 * npts, stream));
    IGAB_CUDA_TRY(cudaMallocAsync(&dJi, 8 * npts * DIM * DW, stream));
    if (!gram) {
      IGAB_CUDA_TRY(cudaMallocAsync(&A1, 8LL * m * kdim * n, stream));
      IGAB_CUDA_TRY(cudaMallocAsync(&A2, 8LL * m * kdim * n, stream));
    }
    EvalOutDev o{};
    // N is only needed for the mass matrix and the load vector
    o.N = (kind == IGAB_ASM_MASS || fe) ? dNv : nullptr; o.dN = (kind == IGAB_ASM_MASS) ? nullptr : ddN; o.detJ = ddet; o.JinvT = (kind == IGAB_ASM_MASS) ? nullptr : dJi;
    igab_status st = eval_for(P, ws, xi1d_host, eb + c0, m, nq1, xi, o, stream);
    if (st) return st;
    double* K = Ke + (long long)c0 * n * n;
    if (gram) {
      int sms = 148, dev = 0;
      IGAB_CUDA_TRY(cudaGetDevice(&dev));
      IGAB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      const bool mass = kind == IGAB_ASM_MASS;
      if (nb > 64) {
        const unsigned grid = (unsigned)std::min<long long>(m, 2LL * sms);
        if (mass)
          gram_dmma_kernel<128, 1, 2><<<grid, 320, 0, stream>>>(nb, (int)nq, m, q0, q1, q2, w[0], w[1], w[2], dNv, ddN, ddet, dJi, K);
        else
          gram_dmma_kernel<128, 3, 2><<<grid, 320, 0, stream>>>(nb, (int)nq, m, q0, q1, q2, w[0], w[1], w[2], dNv, ddN, ddet, dJi, K);
      } else {
        const unsigned grid = (unsigned)std::min<long long>(m, 4LL * sms);
        if (mass)
          gram_dmma_kernel<64, 1, 4><<<grid, 96, 0, stream>>>(nb, (int)nq, m, q0, q1, q2, w[0], w[1], w[2], dNv, ddN, ddet, dJi, K);
        else
          gram_dmma_kernel<64, 3, 4><<<grid, 96, 0, stream>>>(nb, (int)nq, m, q0, q1, q2, w[0], w[1], w[2], dNv, ddN, ddet, dJi, K);
      }
      count_launch();
    } else {
      const long long tot = npts * nb;
      bstack_kernel<DIM, DW><<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(
          kind, npts, nb, nq, q0, q1, q2, w[0], w[DIM > 1 ? 1 : 0], w[DIM > 2 ? 2 : 0], dNv, ddN, ddet, dJi, D_dev, A1, A2,
          nrow, n);
      count_launch();
      if (n >= 32) {
        dim3 grid((n + 63) / 64, (n + 63) / 64, (unsigned)m);
        contract_dmma_kernel<<<grid, 128, 0, stream>>>(n, kdim, A1, A2, K);
      } else {
        dim3 grid((n + 31) / 32, (n + 31) / 32, (unsigned)m);
        contract_fma_kernel<<<grid, 256, 0, stream>>>(n, kdim, A1, A2, K);
      }
      count_launch();
    }
    if (fe) {
      const long long tot = m * nb;
      load_kernel<DIM><<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(m, nb, ncomp, nq, q0, q1, q2, w[0],
                                                                           w[DIM > 1 ? 1 : 0], w[DIM > 2 ? 2 : 0], dNv,
                                                                           ddet, fconst, fe + (long long)c0 * n);
      count_launch();
    }
    IGAB_CUDA_TRY(cudaGetLastError());
    cudaFreeAsync(dNv, stream); cudaFreeAsync(ddN, stream); cudaFreeAsync(ddet, stream);
    cudaFreeAsync(dJi, stream);
    if (A1) cudaFreeAsync(A1, stream);
    if (A2) cudaFreeAsync(A2, stream);
  }
  return IGAB_OK;
}

}  // namespace

igab_status launch_assemble(const PatchDev& P, SfWorkspace& ws, const double* const* xi1d_host, int kind, const double* D_host, const double* D_dev, double fconst, long long elem_begin,
                            long long nelem, const int* nq1, const double* const* xi1d_dev,
                            const double* const* w1d_dev, double* Ke, double* fe, cudaStream_t stream) {
  switch (P.dim * 10 + P.dw) {
    case 11: return assemble_dd<1, 1>(P, ws, xi1d_host, kind, D_host, D_dev, fconst, elem_begin, nelem, nq1, xi1d_dev, w1d_dev, Ke, fe, stream);
    case 22: return assemble_dd<2, 2>(P, ws, xi1d_host, kind, D_host, D_dev, fconst, elem_begin, nelem, nq1, xi1d_dev, w1d_dev, Ke, fe, stream);
    case 23: return assemble_dd<2, 3>(P, ws, xi1d_host, kind, D_host, D_dev, fconst, elem_begin, nelem, nq1, xi1d_dev, w1d_dev, Ke, fe, stream);
    case 33: return assemble_dd<3, 3>(P, ws, xi1d_host, kind, D_host, D_dev, fconst, elem_begin, nelem, nq1, xi1d_dev, w1d_dev, Ke, fe, stream);
  }
  set_error("assemble: unsupported (dim, dimworld)");
  return IGAB_UNSUPPORTED;
}

igab_status launch_volume(const PatchDev& P, SfWorkspace& ws, const double* const* xi1d_host, long long elem_begin, long long nelem, const int* nq1,
                          const double* const* xi1d_dev, const double* const* w1d_dev, double* partial_dev,
                          int npartial, double* result_dev, cudaStream_t stream, const uint8_t* trim_dev) {
  long long nq = 1;
  for (int d = 0; d < P.dim; ++d) nq *= nq1[d];
  const long long npts = nelem * nq;
  if (npts == 0) {
    IGAB_CUDA_TRY(cudaMemsetAsync(result_dev, 0, sizeof(double), stream));
    return IGAB_OK;
  }
  double* ddet = nullptr;
  IGAB_CUDA_TRY(cudaMallocAsync(&ddet, 8 * npts, stream));
  EvalOutDev o{};
  o.detJ = ddet;
  igab_status st = eval_for(P, ws, xi1d_host, elem_begin, nelem, nq1, xi1d_dev, o, stream);
  if (st) return st;
  const int q0 = nq1[0], q1 = P.dim > 1 ? nq1[1] : 1, q2 = P.dim > 2 ? nq1[2] : 1;
  const int nblk = (int)std::min<long long>(npartial, (npts + 255) / 256);
  const double* w0 = w1d_dev[0];
  const double* w1 = w1d_dev[P.dim > 1 ? 1 : 0];
  const double* w2 = w1d_dev[P.dim > 2 ? 2 : 0];
  if (P.dim == 1)
    volume_partial_kernel<1><<<nblk, 256, 0, stream>>>(npts, nq, q0, q1, q2, w0, w1, w2, ddet, partial_dev, trim_dev, elem_begin);
  else if (P.dim == 2)
    volume_partial_kernel<2><<<nblk, 256, 0, stream>>>(npts, nq, q0, q1, q2, w0, w1, w2, ddet, partial_dev, trim_dev, elem_begin);
  else
    volume_partial_kernel<3><<<nblk, 256, 0, stream>>>(npts, nq, q0, q1, q2, w0, w1, w2, ddet, partial_dev, trim_dev, elem_begin);
  volume_final_kernel<<<1, 1024, 0, stream>>>(nblk, partial_dev, result_dev);
  count_launch(2);
  IGAB_CUDA_TRY(cudaGetLastError());
  cudaFreeAsync(ddet, stream);
  return IGAB_OK;
}

}  // namespace igab

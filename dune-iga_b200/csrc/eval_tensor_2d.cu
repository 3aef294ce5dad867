// eval_tensor_2d.cu -- tuned evaluation for 2-D patches (surfaces in R^2 / shells in R^3) at tensor rules: the small
// local bases (nb = (p+1)^2 <= 25) of BASELINE configs C1 (p=2), C4 (p=2 in R^3, second derivatives) and C5 (p=3).
//
// Replaces the same per-point reference chain as eval_tensor_sf.cu (nurbsbasis.hh:77-108, nurbsgeometry.hh:133-138,
// 182-210,257-292, nurbsalgorithms.hh:193-250), restructured for the GPU:
//  * univariate values / derivatives come from the K1 tables (A2.3 once per (direction, element, abscissa));
//    derivative rows are pre-multiplied by span length^k, so everything below is in element-local coordinates;
//  * ONE THREAD PER POINT, whole local basis in registers (compile-time p, Q, order): pass 1 accumulates the
//    homogeneous sums  C^(a) = sum_i w_i dN^(a)_i P_i  and  W^(a) = sum_i w_i dN^(a)_i  for every needed multi-index a,
//    geometry follows from the quotient rule applied to C (x = C/W, J_d = (C_d - W_d x)/W, H from C_cd ...);
//    pass 2 forms R_i, dR_i, d2R_i (generalised Piegl-Tiller 4.20 for |a| <= 2);
//  * 32 consecutive points of a warp own contiguous output runs in every field: rows are staged in a per-warp
//    shared-memory slab (odd row stride => conflict-free) and leave as full-sector coalesced stores.
// HBM-write bound: 296 B/point (C1), 602 B/point (C4), 464 B/point (C5).
#include <algorithm>
#include <cstdlib>

#include "bspline_device.cuh"
#include "igab_internal.cuh"

namespace igab {

namespace {

// tab layout: [e][q][i][a], a = 0..ORDER, derivative a scaled by (span length)^a ; one thread per (e, q)
__global__ void tables2d_kernel(PatchDev P, int d, int nq, int order, const double* __restrict__ xi,
                                double* __restrict__ tab) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.nel[d] * nq) return;
  const int e = t / nq, q = t % nq;
  const int p = P.degree[d];
  const int sp = P.spanTab[d][e];
  const double* U = P.knots[d];
  const double lo = U[sp], scale = U[sp + 1] - lo;
  const double u = xi[q] * scale + lo;
  double dN[3 * kPM];
  ders1d_any(u, U, p, sp, order, dN);
  double* o = tab + ((size_t)t * (p + 1)) * (order + 1);
  double s = 1.0;
  for (int a = 0; a <= order; ++a) {
    for (int i = 0; i <= p; ++i) o[i * (order + 1) + a] = dN[a * kPM + i] * s;
    s *= scale;
  }
}

}  // namespace

// K1 for 2-D patches: tab[d] = [e_d][q][i][a], a = 0..order, cached per handle (rebuilt when the rule, the derivative
// order or the stream changes; the order is part of the key through nq1[2], the unused direction).  Shared by the
// evaluation kernels and the fused 2-D assembly kernel.
igab_status ensure_tables2d(const PatchDev& Pd, SfWorkspace& ws, int Q, int order, const double* const* xi1d_dev,
                            const double* const* xi1d_host, cudaStream_t stream) {
  bool same = ws.valid && ws.stream == stream && ws.nq1[2] == -(order + 1);
  for (int d = 0; d < 2 && same; ++d) {
    same = ws.nq1[d] == Q && (int)ws.xi[d].size() == Q;
    for (int q = 0; q < Q && same; ++q) same = ws.xi[d][q] == xi1d_host[d][q];
  }
  ws.built2d = !same;  // a tables kernel precedes the caller's launch in the stream
  if (same) return IGAB_OK;
  for (int d = 0; d < 2; ++d) {
    const size_t n = (size_t)Pd.nel[d] * Q * (Pd.degree[d] + 1) * (order + 1);
    if (ws.tabCap[d] < n) {
      if (ws.tab[d]) IGAB_CUDA_TRY(cudaFreeAsync(ws.tab[d], stream));
      ws.tab[d] = nullptr;
      ws.tabCap[d] = 0;
      IGAB_CUDA_TRY(cudaMallocAsync(&ws.tab[d], n * sizeof(double), stream));
      ws.tabCap[d] = n;
    }
    const int threads = 128, total = Pd.nel[d] * Q;
    tables2d_kernel<<<(total + threads - 1) / threads, threads, 0, stream>>>(Pd, d, Q, order, xi1d_dev[d], ws.tab[d]);
    ws.nq1[d] = Q;
    ws.xi[d].assign(xi1d_host[d], xi1d_host[d] + Q);
  }
  count_launch(2);
  ws.nq1[2] = -(order + 1);
  ws.valid = true;
  ws.stream = stream;
  return IGAB_OK;
}

namespace {

template <int DW, int P, int Q, int ORDER>
struct Cfg2 {
  static constexpr int NB1 = P + 1, NB = NB1 * NB1, NQ = Q * Q, NH = 3;
  static constexpr int NA = ORDER == 0 ? 1 : (ORDER == 1 ? 3 : 6);  // 0 | e0 e1 | 2e0 2e1 e0+e1
  static constexpr int A1 = ORDER + 1;
  static constexpr int TAB = Q * NB1 * A1;
  static constexpr int WMAX = ORDER >= 2 ? NB * NH : (ORDER >= 1 ? NB * 2 : NB);  // widest row (doubles per point)
  static constexpr int STRIDE = WMAX | 1;
  // p <= 2: fields leave through the TMA engine (one async bulk store per field and chunk) from a ring of two slabs
  static constexpr bool BULK = P <= 2 && ORDER <= 1;  // (order 2 at p = 2, config C4: measured 0.85 -> 0.80 with the bulk path)
  static constexpr int SLAB = 32 * STRIDE;
  static constexpr int SLABS = BULK ? 2 : 1;
};

// alpha index -> (a0, a1): 0:(0,0) 1:(1,0) 2:(0,1) 3:(2,0) 4:(0,2) 5:(1,1)   (H row order of nurbsgeometry.hh:262-289)
__device__ __forceinline__ constexpr int al0(int a) { return a == 1 ? 1 : (a == 3 ? 2 : (a == 5 ? 1 : 0)); }
__device__ __forceinline__ constexpr int al1(int a) { return a == 2 ? 1 : (a == 4 ? 2 : (a == 5 ? 1 : 0)); }

// ---- async bulk copy shared -> global (TMA engine; SASS: UBLKCP), bulk-group completion (as in eval_tensor_sf.cu)
__device__ __forceinline__ void fence_proxy_async_smem2() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store2(void* gdst, const void* ssrc, unsigned bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit2() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read2() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// stage one field row-by-row (thread = point) and flush the warp's contiguous run.
// BULK (p <= 2): the fields take turns on two slabs (fld counts them); every field commits ONE bulk group -- an empty one
// on the path that copies with ordinary stores -- so "at most one group pending" always means that the copy which last
// used the slab about to be written has finished reading it.  Rows whose width is odd or 2 mod 4 are staged DENSE (stride
// W: conflict-free / two-way conflicts on the staging stores) and leave through one async bulk store of the TMA engine:
// the slab-load + global-store round of the copy loop -- two of the three trips of every stored double through the L1
// data pipe, and a third of the kernel's instructions -- disappears.
template <int W, int STRIDE, bool FAST, bool BULK>
__device__ __forceinline__ void flush_field(double* __restrict__ slab0, int& fld, const double (&row)[W],
                                            double* __restrict__ g, int lane, int nvalid) {
  double* __restrict__ slab = slab0;
  if constexpr (BULK) {
    slab = slab0 + (fld & 1) * (32 * STRIDE);
    ++fld;
    if (lane == 0) bulk_wait_read2<1>();
    __syncwarp();
    constexpr bool DENSE_OK = (W % 2 == 1) || (W % 4 == 2);
    if constexpr (DENSE_OK) {
      if (nvalid == 32 && (reinterpret_cast<unsigned long long>(g) & 15ull) == 0) {
#pragma unroll
        for (int k = 0; k < W; ++k) slab[lane * W + k] = row[k];
        fence_proxy_async_smem2();  // generic-proxy writes -> visible to the async proxy
        __syncwarp();
        if (lane == 0) {
          bulk_store2(g, slab, 32 * W * 8);
          bulk_commit2();
        }
        return;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < W; ++k) slab[lane * STRIDE + k] = row[k];
  __syncwarp();
  // (odd widths have STRIDE = W: the slab is dense, entry j sits at slab[j]; even widths need j / W per entry and are unrolled
  //  for FAST = p >= 3 only: in the 96-register p = 2 instantiations that address arithmetic spilled and cost 5-10 %)
  if ((STRIDE == W || FAST) && nvalid == 32) {
    // full chunk (all but the last one of a launch): W fully unrolled 256-byte stores, entry j = lane + 32 it of the run
    // sits at slab[j + (j / W) (STRIDE - W)] -- no loop counter, no division per entry (the general loop below spent 12
    // instructions per stored double: 39 % of the kernel's instructions at config C5, ncu source view)
#pragma unroll
    for (int it = 0; it < W; ++it) {
      const int j = lane + 32 * it;
      g[j] = slab[j + (j / W) * (STRIDE - W)];
      // four loads in flight, not W: the fully hoisted version spilled in the 96-register instantiations (p = 2)
      if (it % 4 == 3) asm volatile("" ::: "memory");
    }
  } else {
    const int n = nvalid * W;
    for (int j = lane; j < n; j += 32) {
      const int r = j / W, k = j - r * W;
      g[j] = slab[r * STRIDE + k];
    }
  }
  __syncwarp();
  if constexpr (BULK) {
    if (lane == 0) bulk_commit2();  // empty group: keeps "one group per field"
  }
}

// LIST = false: tensor rule, univariate rows from the K1 tables.  LIST = true: arbitrary element-local points (trimmed
// batches, utils/fillquadraturerule.hh:8-19): rows computed per point with ders1d_reg_rd (A2.3 in registers, tabulated reciprocals).
template <int DW, int P, int Q, int ORDER, bool LIST>
__global__ void __launch_bounds__(128, (ORDER >= 2 ? 3 : (P <= 2 ? 5 : (P >= 4 ? 2 : 4)))) eval_2d_kernel(PatchDev Pd, long long elem_begin, long long npts,
                                                      const double* __restrict__ tab0, const double* __restrict__ tab1,
                                                      const int* __restrict__ elemList, const double* __restrict__ xiList,
                                                      EvalOutDev out) {
  using C = Cfg2<DW, P, Q, ORDER>;
  constexpr int NB1 = C::NB1, NB = C::NB, NQ = C::NQ, NA = C::NA, A1 = C::A1, NH = C::NH;
  extern __shared__ __align__(16) double smem2[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* slab = smem2 + (size_t)warp * (C::SLAB * C::SLABS);
  int fld = 0;  // fields flushed so far (slab ring of the bulk path)
  // Programmatic dependent launch: the next launch of this kernel in the stream may become resident while this one drains.
  // Tensor rules read only handle-owned buffers (tables, control net, spans) that no kernel preceding this launch wrote --
  // launch_2d drops the attribute when it has just rebuilt the tables, and control-net updates are copies, which keep the
  // plain stream order -- so the kernel does everything up to its first store (loads, the homogeneous sums of its first
  // chunk) before it waits for its predecessor to complete (griddepcontrol.wait: outputs may alias the predecessor's).
  // Point lists read caller arrays that the preceding kernel may have produced: they wait before their first load, and
  // only the launch itself overlaps.  At config C1 (31 us per launch) launch gap and ramp are a tenth of the time.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  bool pdl_waited = false;
  if constexpr (LIST) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    pdl_waited = true;
  }
  // persistent warps: the grid is a fixed number of resident blocks per SM (no partial last wave: at config C1 the
  // one-block-per-128-points grid ran 6.2 waves), every warp walks chunks of 32 points with the grid's stride
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5) * 32;
  for (long long pt0 = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * 32; pt0 < npts; pt0 += wstride) {
  const long long pt = pt0 + lane;
  const bool valid = pt < npts;
  const int nvalid = (int)min(32LL, npts - pt0);
  const long long ptc = valid ? pt : npts - 1;
  const long long e = LIST ? (long long)elemList[ptc] : elem_begin + ptc / (NQ > 0 ? NQ : 1);
  const int e0 = (int)(e % Pd.nel[0]), e1 = (int)(e / Pd.nel[0]);

  // univariate rows of this point
  double n0[A1][NB1], n1[A1][NB1];
  if constexpr (LIST) {
    const int s0 = Pd.spanTab[0][e0], s1 = Pd.spanTab[1][e1];
    const double* U0 = Pd.knots[0];
    const double* U1 = Pd.knots[1];
    const double lo0 = U0[s0], sc0 = U0[s0 + 1] - lo0, lo1 = U1[s1], sc1 = U1[s1 + 1] - lo1;
    constexpr int TRI = P * (P + 1) / 2 > 0 ? P * (P + 1) / 2 : 1;
    ders1d_reg_rd<P, ORDER>(xiList[ptc * 2] * sc0 + lo0, U0, s0, sc0, Pd.rden[0] + (size_t)e0 * TRI, n0);
    ders1d_reg_rd<P, ORDER>(xiList[ptc * 2 + 1] * sc1 + lo1, U1, s1, sc1, Pd.rden[1] + (size_t)e1 * TRI, n1);
  } else {
    const int q = (int)(ptc % (NQ > 0 ? NQ : 1)), q0 = q % (Q > 0 ? Q : 1), q1 = q / (Q > 0 ? Q : 1);
    const double* t0 = tab0 + ((size_t)e0 * Q + q0) * NB1 * A1;
    const double* t1 = tab1 + ((size_t)e1 * Q + q1) * NB1 * A1;
#pragma unroll
    for (int i = 0; i < NB1; ++i)
#pragma unroll
      for (int a = 0; a < A1; ++a) {
        n0[a][i] = __ldg(t0 + i * A1 + a);
        n1[a][i] = __ldg(t1 + i * A1 + a);
      }
  }
  const long long cpBase = (Pd.spanTab[0][e0] - P) + (long long)Pd.ncp[0] * (Pd.spanTab[1][e1] - P);
  const double* __restrict__ cp = Pd.cp;

  // ---- pass 1: homogeneous sums
  double Wa[NA], Ca[NA][DW], wv[NB];
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    Wa[a] = 0.0;
#pragma unroll
    for (int c = 0; c < DW; ++c) Ca[a][c] = 0.0;
  }
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
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        const double A = n0[al0(a)][i0] * n1[al1(a)][i1] * w;
        Wa[a] += A;
#pragma unroll
        for (int c = 0; c < DW; ++c) Ca[a][c] = fma(A, Pc[c], Ca[a][c]);
      }
    }
  const double invW = 1.0 / Wa[0];

  // ---- geometry by the quotient rule on the homogeneous sums
  double x[DW], J[2][DW], H[NH][DW];
#pragma unroll
  for (int c = 0; c < DW; ++c) x[c] = Ca[0][c] * invW;
  if constexpr (ORDER >= 1) {
#pragma unroll
    for (int d = 0; d < 2; ++d)
#pragma unroll
      for (int c = 0; c < DW; ++c) J[d][c] = (Ca[1 + d][c] - Wa[1 + d] * x[c]) * invW;
  }
  if constexpr (ORDER >= 2) {
#pragma unroll
    for (int c = 0; c < DW; ++c) {
      H[0][c] = (Ca[3][c] - 2.0 * Wa[1] * J[0][c] - Wa[3] * x[c]) * invW;
      H[1][c] = (Ca[4][c] - 2.0 * Wa[2] * J[1][c] - Wa[4] * x[c]) * invW;
      H[2][c] = (Ca[5][c] - Wa[1] * J[1][c] - Wa[2] * J[0][c] - Wa[5] * x[c]) * invW;
    }
  }

  // ---- pass 2 + stores, field by field
  if (!pdl_waited) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    pdl_waited = true;
  }
  if (out.N) {
    double row[NB];
#pragma unroll
    for (int i1 = 0; i1 < NB1; ++i1)
#pragma unroll
      for (int i0 = 0; i0 < NB1; ++i0) row[i0 + NB1 * i1] = n0[0][i0] * n1[0][i1] * wv[i0 + NB1 * i1] * invW;
    flush_field<NB, C::STRIDE, (P >= 3), C::BULK>(slab, fld, row, out.N + pt0 * NB, lane, nvalid);
  }
  if constexpr (ORDER >= 1) {
    if (out.dN || (ORDER >= 2 && out.d2N)) {
      double rowD[NB * 2];
      [[maybe_unused]] double rowH[ORDER >= 2 ? NB * NH : 1];
#pragma unroll
      for (int i1 = 0; i1 < NB1; ++i1)
#pragma unroll
        for (int i0 = 0; i0 < NB1; ++i0) {
          const int i = i0 + NB1 * i1;
          const double w = wv[i];
          const double R = n0[0][i0] * n1[0][i1] * w * invW;
          const double R0 = (n0[1][i0] * n1[0][i1] * w - Wa[1] * R) * invW;
          const double R1 = (n0[0][i0] * n1[1][i1] * w - Wa[2] * R) * invW;
          rowD[2 * i] = R0;
          rowD[2 * i + 1] = R1;
          if constexpr (ORDER >= 2) {
            rowH[NH * i + 0] = (n0[2][i0] * n1[0][i1] * w - 2.0 * Wa[1] * R0 - Wa[3] * R) * invW;
            rowH[NH * i + 1] = (n0[0][i0] * n1[2][i1] * w - 2.0 * Wa[2] * R1 - Wa[4] * R) * invW;
            rowH[NH * i + 2] = (n0[1][i0] * n1[1][i1] * w - Wa[1] * R1 - Wa[2] * R0 - Wa[5] * R) * invW;
          }
        }
      if (out.dN) flush_field<NB * 2, C::STRIDE, (P >= 3), C::BULK>(slab, fld, rowD, out.dN + pt0 * NB * 2, lane, nvalid);
      if constexpr (ORDER >= 2) {
        if (out.d2N) flush_field<NB * NH, C::STRIDE, (P >= 3), C::BULK>(slab, fld, rowH, out.d2N + pt0 * NB * NH, lane, nvalid);
      }
    }
  }
  if (out.x) {
    double row[DW];
#pragma unroll
    for (int c = 0; c < DW; ++c) row[c] = x[c];
    flush_field<DW, C::STRIDE, (P >= 3), C::BULK>(slab, fld, row, out.x + pt0 * DW, lane, nvalid);
  }
  if constexpr (ORDER >= 1) {
    if (out.Jt) {
      double row[2 * DW];
#pragma unroll
      for (int d = 0; d < 2; ++d)
#pragma unroll
        for (int c = 0; c < DW; ++c) row[d * DW + c] = J[d][c];
      flush_field<2 * DW, C::STRIDE, (P >= 3), C::BULK>(slab, fld, row, out.Jt + pt0 * 2 * DW, lane, nvalid);
    }
    double det;
    if constexpr (DW == 2) {
      det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    } else {
      const double v0 = J[0][0] * J[1][1] - J[0][1] * J[1][0], v1 = J[0][0] * J[1][2] - J[1][0] * J[0][2],
                   v2 = J[0][1] * J[1][2] - J[0][2] * J[1][1];
      det = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
    }
    if (out.detJ && valid) out.detJ[pt] = fabs(det);
    if (out.JinvT) {
      // J^T (J J^T)^-1, [DW][2]  (rightInvA, nurbsgeometry.hh:205-210)
      double g00 = 0, g01 = 0, g11 = 0;
#pragma unroll
      for (int c = 0; c < DW; ++c) {
        g00 += J[0][c] * J[0][c];
        g01 += J[0][c] * J[1][c];
        g11 += J[1][c] * J[1][c];
      }
      const double idet = 1.0 / (g00 * g11 - g01 * g01);
      double row[2 * DW];
#pragma unroll
      for (int c = 0; c < DW; ++c) {
        row[c * 2 + 0] = (J[0][c] * g11 - J[1][c] * g01) * idet;
        row[c * 2 + 1] = (J[1][c] * g00 - J[0][c] * g01) * idet;
      }
      flush_field<2 * DW, C::STRIDE, (P >= 3), C::BULK>(slab, fld, row, out.JinvT + pt0 * 2 * DW, lane, nvalid);
    }
  }
  if constexpr (ORDER >= 2) {
    if (out.H) {
      double row[NH * DW];
#pragma unroll
      for (int r = 0; r < NH; ++r)
#pragma unroll
        for (int c = 0; c < DW; ++c) row[r * DW + c] = H[r][c];
      flush_field<NH * DW, C::STRIDE, (P >= 3), C::BULK>(slab, fld, row, out.H + pt0 * NH * DW, lane, nvalid);
    }
  }
  }  // chunk loop
  if constexpr (C::BULK) {
    if (lane == 0) bulk_wait_read2<0>();  // the slabs must outlive the copies that read them
    __syncwarp();
  }
}

// launch with the programmatic-stream-serialization attribute (see the kernel); allow = false or IGAB_NO_PDL=1: plain launch
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(bool allow, void (*kernel)(KArgs...), unsigned blocks, unsigned threads, size_t smem,
                       cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  const char* env = getenv("IGAB_NO_PDL");
  const bool pdl = allow && pdl_allowed() && !(env && env[0] == '1');
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <int DW, int P, int Q, int ORDER>
igab_status launch_2d(const PatchDev& Pd, SfWorkspace& ws, const PointSource& src, const double* const* xi1d_host,
                      long long nelem, const EvalOutDev& out, cudaStream_t stream) {
  using C = Cfg2<DW, P, Q, ORDER>;
  igab_status tst = ensure_tables2d(Pd, ws, Q, ORDER, src.xi1d, xi1d_host, stream);
  if (tst) return tst;
  const long long npts = nelem * C::NQ;
  constexpr int WARPS = 4;
  const size_t smem = (size_t)WARPS * C::SLAB * C::SLABS * sizeof(double);
  KernelCfg kc;
  if (igab_status cst = kernel_config((const void*)eval_2d_kernel<DW, P, Q, ORDER, false>, WARPS * 32, smem, &kc)) return cst;
  long long blocks = (npts + WARPS * 32 - 1) / (WARPS * 32);
  // persistent grid where the tail of the last wave matters (measured: +7 % at 6 waves, +5 % at 17, -2 % at 177)
  if (!getenv("IGAB_2D_NONPERSISTENT") && blocks <= 64LL * kc.sms * kc.blocksPerSm)
    blocks = std::min(blocks, (long long)kc.sms * kc.blocksPerSm);
  if (blocks > 0x7fffffffLL) {
    set_error("eval_tensor: too many points for one launch");
    return IGAB_INVALID_ARGUMENT;
  }
  IGAB_CUDA_TRY(launch_pdl(!ws.built2d, eval_2d_kernel<DW, P, Q, ORDER, false>, (unsigned)blocks, WARPS * 32, smem, stream, Pd,
                           (long long)src.elem_begin, npts, (const double*)ws.tab[0], (const double*)ws.tab[1],
                           (const int*)nullptr, (const double*)nullptr, out));
  count_launch();
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

template <int DW, int ORDER>
igab_status dispatch_pq(int p, int q, const PatchDev& Pd, SfWorkspace& ws, const PointSource& src,
                        const double* const* xi, long long nelem, const EvalOutDev& out, cudaStream_t stream) {
  if (p == 1 && q == 2) return launch_2d<DW, 1, 2, ORDER>(Pd, ws, src, xi, nelem, out, stream);
  if (p == 2 && q == 3) return launch_2d<DW, 2, 3, ORDER>(Pd, ws, src, xi, nelem, out, stream);
  if (p == 2 && q == 4) return launch_2d<DW, 2, 4, ORDER>(Pd, ws, src, xi, nelem, out, stream);
  if (p == 3 && q == 4) return launch_2d<DW, 3, 4, ORDER>(Pd, ws, src, xi, nelem, out, stream);
  if (p == 3 && q == 5) return launch_2d<DW, 3, 5, ORDER>(Pd, ws, src, xi, nelem, out, stream);
  if constexpr (ORDER <= 1) {
    if (p == 4 && q == 5) return launch_2d<DW, 4, 5, ORDER>(Pd, ws, src, xi, nelem, out, stream);
  }
  set_error("eval_tensor: no 2-D kernel for this (degree, rule)");
  return IGAB_UNSUPPORTED;
}

template <int DW, int P, int ORDER>
igab_status launch_list(const PatchDev& Pd, const PointSource& src, long long npts, const EvalOutDev& out,
                        cudaStream_t stream) {
  using C = Cfg2<DW, P, 0, ORDER>;
  constexpr int WARPS = 4;
  const size_t smem = (size_t)WARPS * C::SLAB * C::SLABS * sizeof(double);
  KernelCfg kc;
  if (igab_status cst = kernel_config((const void*)eval_2d_kernel<DW, P, 0, ORDER, true>, WARPS * 32, smem, &kc)) return cst;
  long long blocks = (npts + WARPS * 32 - 1) / (WARPS * 32);
  // persistent grid where the tail of the last wave matters (measured: +7 % at 6 waves, +5 % at 17, -2 % at 177)
  if (!getenv("IGAB_2D_NONPERSISTENT") && blocks <= 64LL * kc.sms * kc.blocksPerSm)
    blocks = std::min(blocks, (long long)kc.sms * kc.blocksPerSm);
  if (blocks > 0x7fffffffLL) {
    set_error("eval_points: too many points for one launch");
    return IGAB_INVALID_ARGUMENT;
  }
  IGAB_CUDA_TRY(launch_pdl(true, eval_2d_kernel<DW, P, 0, ORDER, true>, (unsigned)blocks, WARPS * 32, smem, stream, Pd, 0LL, npts,
                           (const double*)nullptr, (const double*)nullptr, (const int*)src.elem, (const double*)src.xi, out));
  count_launch();
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

template <int DW, int ORDER>
igab_status dispatch_list(int p, const PatchDev& Pd, const PointSource& src, long long npts, const EvalOutDev& out,
                          cudaStream_t stream) {
  if (p == 1) return launch_list<DW, 1, ORDER>(Pd, src, npts, out, stream);
  if (p == 2) return launch_list<DW, 2, ORDER>(Pd, src, npts, out, stream);
  if (p == 3) return launch_list<DW, 3, ORDER>(Pd, src, npts, out, stream);
  set_error("eval_points: no 2-D list kernel for this degree");
  return IGAB_UNSUPPORTED;
}

}  // namespace

bool eval_points_2d_supported(const PatchDev& P) {
  return P.dim == 2 && (P.dw == 2 || P.dw == 3) && P.degree[0] == P.degree[1] && P.degree[0] >= 1 && P.degree[0] <= 3;
}

igab_status launch_eval_points_2d(const PatchDev& P, const PointSource& src, long long npts, int order,
                                  const EvalOutDev& out, cudaStream_t stream) {
  if (!eval_points_2d_supported(P)) {
    set_error("eval_points: no 2-D list kernel for this configuration");
    return IGAB_UNSUPPORTED;
  }
  EvalOutDev o = out;
  int ord = 0;
  if (order >= 1 && (o.dN || o.Jt || o.detJ || o.JinvT)) ord = 1;
  if (order >= 2 && (o.d2N || o.H)) ord = 2;
  if (ord < 2) o.d2N = o.H = nullptr;
  if (ord < 1) o.dN = o.Jt = o.detJ = o.JinvT = nullptr;
  const int p = P.degree[0];
  if (P.dw == 2) {
    if (ord == 0) return dispatch_list<2, 0>(p, P, src, npts, o, stream);
    if (ord == 1) return dispatch_list<2, 1>(p, P, src, npts, o, stream);
    return dispatch_list<2, 2>(p, P, src, npts, o, stream);
  }
  if (ord == 0) return dispatch_list<3, 0>(p, P, src, npts, o, stream);
  if (ord == 1) return dispatch_list<3, 1>(p, P, src, npts, o, stream);
  return dispatch_list<3, 2>(p, P, src, npts, o, stream);
}

bool eval_tensor_2d_supported(const PatchDev& P, const int* nq1, int order, const EvalOutDev& out) {
  (void)out;
  if (P.dim != 2 || (P.dw != 2 && P.dw != 3)) return false;
  if (P.degree[0] != P.degree[1] || nq1[0] != nq1[1]) return false;
  if (order < 0 || order > 2) return false;
  const int p = P.degree[0], q = nq1[0];
  return (p == 1 && q == 2) || (p == 2 && q == 3) || (p == 2 && q == 4) || (p == 3 && q == 4) || (p == 3 && q == 5) ||
         (p == 4 && q == 5 && order <= 1 && !out.d2N && !out.H);
}

igab_status launch_eval_tensor_2d(const PatchDev& P, SfWorkspace& ws, const PointSource& src,
                                  const double* const* xi1d_host, long long nelem, int order, const EvalOutDev& out,
                                  cudaStream_t stream) {
  if (!eval_tensor_2d_supported(P, src.nq1, order, out)) {
    set_error("eval_tensor: no 2-D kernel for this configuration");
    return IGAB_UNSUPPORTED;
  }
  EvalOutDev o = out;
  // compile only what is asked for: the derivative order of the kernel follows the requested outputs
  int ord = 0;
  if (order >= 1 && (o.dN || o.Jt || o.detJ || o.JinvT)) ord = 1;
  if (order >= 2 && (o.d2N || o.H)) ord = 2;
  if (ord < 2) o.d2N = o.H = nullptr;
  if (ord < 1) o.dN = o.Jt = o.detJ = o.JinvT = nullptr;
  const int p = P.degree[0], q = src.nq1[0];
  if (P.dw == 2) {
    if (ord == 0) return dispatch_pq<2, 0>(p, q, P, ws, src, xi1d_host, nelem, o, stream);
    if (ord == 1) return dispatch_pq<2, 1>(p, q, P, ws, src, xi1d_host, nelem, o, stream);
    return dispatch_pq<2, 2>(p, q, P, ws, src, xi1d_host, nelem, o, stream);
  }
  if (ord == 0) return dispatch_pq<3, 0>(p, q, P, ws, src, xi1d_host, nelem, o, stream);
  if (ord == 1) return dispatch_pq<3, 1>(p, q, P, ws, src, xi1d_host, nelem, o, stream);
  return dispatch_pq<3, 2>(p, q, P, ws, src, xi1d_host, nelem, o, stream);
}

}  // namespace igab

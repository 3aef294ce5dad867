// igab_api.cu -- the C-ABI of include/igab200.h: handle management, integer kernels (spans, DOF map), dispatch of the
// evaluation / assembly / reduction kernels, and host<->device staging for IGAB_HOST buffers.
#include <dlfcn.h>
#include <sched.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include "igab_internal.cuh"

namespace igab {

static thread_local std::string g_last_error;

igab_status kernel_config(const void* func, int threads, size_t smem, KernelCfg* out) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, KernelCfg> cache;
  int dev = 0;
  IGAB_CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find({func, dev});
  if (it == cache.end()) {
    KernelCfg c;
    if (smem > 48 * 1024) IGAB_CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    IGAB_CUDA_TRY(cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev));
    IGAB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.blocksPerSm, func, threads, smem));
    if (c.blocksPerSm < 1) c.blocksPerSm = 1;
    it = cache.emplace(std::make_pair(func, dev), c).first;
  }
  *out = it->second;
  return IGAB_OK;
}
static std::atomic<long long> g_launches{0};

void set_error(const std::string& msg) { g_last_error = msg; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
static thread_local int t_no_pdl = 0;
NoPdlScope::NoPdlScope() { ++t_no_pdl; }
NoPdlScope::~NoPdlScope() { --t_no_pdl; }
bool pdl_allowed() { return t_no_pdl == 0; }
igab_status cuda_fail(cudaError_t e, const char* what) {
  g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  return e == cudaErrorMemoryAllocation ? IGAB_OUT_OF_MEMORY : IGAB_CUDA_ERROR;
}

// ---- integer kernels ------------------------------------------------------------------------------------------------
// K7a: spans of all elements [nelem][dim]  (nurbspatch.hh:248-253; element index first-direction-fastest)
__global__ void spans_kernel(PatchDev P, int32_t* __restrict__ span) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P.nelem) return;
  long long h = e;
  for (int d = 0; d < P.dim; ++d) {
    span[e * P.dim + d] = P.spanTab[d][h % P.nel[d]];
    h /= P.nel[d];
  }
}

// K7b: element-to-global DOF map (nurbsbasis.hh:552-587): one thread per (element, local basis function);
// g_d = max(s_d - p_d, 0) + l_d ; flat = g_0 + n_0 (g_1 + n_1 g_2), n_d = |U_d| - p_d - 1.
__global__ void dofmap_kernel(PatchDev P, int32_t* __restrict__ dof) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.nelem * P.nb) return;
  const long long e = t / P.nb;
  int i = (int)(t % P.nb);
  long long h = e, flat = 0, stride = 1;
  for (int d = 0; d < P.dim; ++d) {
    const int s = P.spanTab[d][h % P.nel[d]];
    h /= P.nel[d];
    const int l = i % (P.degree[d] + 1);
    i /= (P.degree[d] + 1);
    const int g = max(s - P.degree[d], 0) + l;
    flat += g * stride;
    stride *= (P.nknots[d] - P.degree[d] - 1);
  }
  dof[t] = (int32_t)flat;
}

// trim compaction (nurbsbasis.hh:599-615): mark DOFs touched by non-empty elements, exclusive scan, remap.
__global__ void dof_mark_kernel(PatchDev P, const uint8_t* __restrict__ trim, const int32_t* __restrict__ dof,
                                int32_t* __restrict__ used) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.nelem * P.nb) return;
  if (trim[t / P.nb] != IGAB_TRIM_EMPTY) used[dof[t]] = 1;
}
// Ordered exclusive scan of a 0/1 predicate over n entries in three phases (tile sums, scan of the tile sums by one
// block, tile-local scan + write): exact integers, any n, every SM busy.  The predicate is used[i] != 0 (DOF marks) or
// trim[i] != IGAB_TRIM_EMPTY (element flags).  map[i] = exclusive rank of a kept entry, -1 otherwise; inv (nullable):
// inv[rank] = i -- the list of kept entries in ascending order.
constexpr int kScanThreads = 256, kScanPer = 16, kScanTile = kScanThreads * kScanPer;
__device__ __forceinline__ int scan_pred(const int32_t* __restrict__ used, const uint8_t* __restrict__ trim, long long i) {
  return used ? (used[i] != 0) : (trim[i] != IGAB_TRIM_EMPTY);
}
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const int32_t* __restrict__ used,
                                                                      const uint8_t* __restrict__ trim, long long n,
                                                                      long long* __restrict__ tileSum) {
  __shared__ int wsum[kScanThreads / 32];
  const long long base = (long long)blockIdx.x * kScanTile;
  int c = 0;
  for (int k = 0; k < kScanPer; ++k) {
    const long long i = base + (long long)k * kScanThreads + threadIdx.x;
    if (i < n) c += scan_pred(used, trim, i);
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) t += wsum[w];
    tileSum[blockIdx.x] = t;
  }
}
// exclusive scan of the tile sums in place; tileSum[ntiles] = total
__global__ void __launch_bounds__(1024) scan_tiles_kernel(long long* __restrict__ tileSum, long long ntiles) {
  __shared__ long long carry;
  __shared__ long long wsum[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long base = 0; base < ntiles; base += blockDim.x) {
    const long long i = base + threadIdx.x;
    const long long v = i < ntiles ? tileSum[i] : 0;
    long long s = v;
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane == 31) wsum[warp] = s;
    __syncthreads();
    if (warp == 0) {
      long long w = wsum[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      wsum[lane] = w;
    }
    __syncthreads();
    if (i < ntiles) tileSum[i] = carry + (warp ? wsum[warp - 1] : 0) + s - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += wsum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) tileSum[ntiles] = carry;
}
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int32_t* __restrict__ used,
                                                                  const uint8_t* __restrict__ trim, long long n,
                                                                  const long long* __restrict__ tileSum,
                                                                  int32_t* __restrict__ map, int32_t* __restrict__ inv,
                                                                  long long* __restrict__ total) {
  __shared__ int wsum[kScanThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // thread t owns the kScanPer consecutive entries base + t*kScanPer .. so that ranks follow the index order
  const long long first = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanPer;
  int v[kScanPer], c = 0;
#pragma unroll
  for (int k = 0; k < kScanPer; ++k) {
    v[k] = (first + k < n) ? scan_pred(used, trim, first + k) : 0;
    c += v[k];
  }
  int s = c;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, s, o);
    if (lane >= o) s += t;
  }
  if (lane == 31) wsum[warp] = s;
  __syncthreads();
  if (warp == 0) {
    int w = lane < kScanThreads / 32 ? wsum[lane] : 0;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    if (lane < kScanThreads / 32) wsum[lane] = w;
  }
  __syncthreads();
  long long r = tileSum[blockIdx.x] + (warp ? wsum[warp - 1] : 0) + s - c;
#pragma unroll
  for (int k = 0; k < kScanPer; ++k) {
    const long long i = first + k;
    if (i >= n) break;
    if (map) map[i] = v[k] ? (int32_t)r : -1;
    if (v[k] && inv) inv[r] = (int32_t)i;
    r += v[k];
  }
  if (total && blockIdx.x == 0 && threadIdx.x == 0) *total = tileSum[gridDim.x];
}
__global__ void dof_remap_kernel(long long n, const int32_t* __restrict__ map, int32_t* __restrict__ dof) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) dof[t] = map[dof[t]];
}

igab_status launch_flag_scan(const int32_t* used, const uint8_t* trim, long long n, int32_t* map, int32_t* inv,
                             long long* total_dev, cudaStream_t stream) {
  if (n <= 0) {
    if (total_dev) IGAB_CUDA_TRY(cudaMemsetAsync(total_dev, 0, sizeof(long long), stream));
    return IGAB_OK;
  }
  const long long ntiles = (n + kScanTile - 1) / kScanTile;
  long long* tileSum = nullptr;
  IGAB_CUDA_TRY(cudaMallocAsync(&tileSum, sizeof(long long) * (ntiles + 1), stream));
  scan_tile_sums_kernel<<<(unsigned)ntiles, kScanThreads, 0, stream>>>(used, trim, n, tileSum);
  scan_tiles_kernel<<<1, 1024, 0, stream>>>(tileSum, ntiles);
  scan_apply_kernel<<<(unsigned)ntiles, kScanThreads, 0, stream>>>(used, trim, n, tileSum, map, inv, total_dev);
  count_launch(3);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(tileSum, stream);
  if (e != cudaSuccess) return cuda_fail(e, "flag scan");
  return IGAB_OK;
}

igab_status launch_spans(const PatchDev& P, int32_t* span, cudaStream_t stream) {
  const int th = 256;
  spans_kernel<<<(unsigned)((P.nelem + th - 1) / th), th, 0, stream>>>(P, span);
  count_launch();
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}
igab_status launch_dofmap(const PatchDev& P, int32_t* dof, cudaStream_t stream) {
  const int th = 256;
  const long long n = P.nelem * P.nb;
  dofmap_kernel<<<(unsigned)((n + th - 1) / th), th, 0, stream>>>(P, dof);
  count_launch();
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}
igab_status launch_dof_compact(const PatchDev& P, const uint8_t* trim_dev, int32_t* dof, int32_t* used, int32_t* map,
                               long long ndof, long long* total_dev, cudaStream_t stream) {
  const int th = 256;
  const long long n = P.nelem * P.nb;
  IGAB_CUDA_TRY(cudaMemsetAsync(used, 0, sizeof(int32_t) * ndof, stream));
  dof_mark_kernel<<<(unsigned)((n + th - 1) / th), th, 0, stream>>>(P, trim_dev, dof, used);
  igab_status st = launch_flag_scan(used, nullptr, ndof, map, nullptr, total_dev, stream);
  if (st) return st;
  dof_remap_kernel<<<(unsigned)((n + th - 1) / th), th, 0, stream>>>(n, map, dof);
  count_launch(2);
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

// ---- host helpers ---------------------------------------------------------------------------------------------------
static bool feq8(double a, double b) {
  const double eps = 8.0 * 2.220446049250313e-16;
  return std::fabs(a - b) <= eps * std::max(std::fabs(a), std::fabs(b));
}
// findSpanCorrected (bsplinealgorithms.hh:24-32): integer work done once per direction on the host at handle creation
static int find_span_corrected(int p, double u, const std::vector<double>& U, int offset) {
  if (u <= U.front()) return p;
  if (u >= U.back()) return (int)U.size() - p - 2;
  auto it = std::upper_bound(U.begin() + p - 1 + offset, U.end(), u);
  return (int)(it - U.begin()) - 1;
}

constexpr int kMaxRule = 64;  // points per direction of a tensor rule

// Handle-owned device scratch (staged rule, univariate tables, scheduler counters, reduction buffers, the control net
// on update) is shared by all calls on a handle.  When the caller moves to ANOTHER stream, the work it queued on the
// previous one may still be reading that scratch: order the new stream behind it with an event recorded on the old
// stream at the moment of the switch (no per-call cost on the steady path: one pointer compare).
igab_status enter_stream(igab_patch* p, cudaStream_t stream) {
  if (!p) {
    set_error("null patch handle");
    return IGAB_INVALID_ARGUMENT;
  }
  if (p->pending_host) {  // an igab_eval_tensor_begin is in flight: its staging buffers are handle scratch
    if (igab_status wst = igab_patch_wait(p)) return wst;
  }
  if (p->have_last_stream && p->last_stream != stream) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (p->last_stream) cudaStreamIsCapturing(p->last_stream, &cap);
    if (cap != cudaStreamCaptureStatusNone) {
      set_error("this handle is inside a CUDA graph capture on another stream: end the capture before switching streams");
      return IGAB_INVALID_ARGUMENT;
    }
    if (!p->stream_switch_ev) IGAB_CUDA_TRY(cudaEventCreateWithFlags(&p->stream_switch_ev, cudaEventDisableTiming));
    if (cudaEventRecord(p->stream_switch_ev, p->last_stream) == cudaSuccess) {
      IGAB_CUDA_TRY(cudaStreamWaitEvent(stream, p->stream_switch_ev, 0));
    } else {
      cudaGetLastError();  // the caller destroyed the old stream: its work has been released with it
    }
  }
  p->last_stream = stream;
  p->have_last_stream = true;
  return IGAB_OK;
}

struct RuleDev {
  const double* xi[kMaxDim];
  const double* w[kMaxDim];
};
// copies the rule to the handle's device scratch on `stream`
static igab_status stage_rule(igab_patch* p, const int* nq1, const double* const* xi1d, const double* const* w1d,
                              RuleDev* r, cudaStream_t stream) {
  double host[2 * kMaxDim * kMaxRule];
  std::memset(host, 0, sizeof(host));
  for (int d = 0; d < p->dev.dim; ++d) {
    if (nq1[d] < 1 || nq1[d] > kMaxRule) {
      set_error("quadrature rule: 1..64 points per direction supported");
      return IGAB_INVALID_ARGUMENT;
    }
    if (!xi1d || !xi1d[d]) {
      set_error("quadrature rule: null abscissae");
      return IGAB_INVALID_ARGUMENT;
    }
    std::memcpy(host + d * kMaxRule, xi1d[d], sizeof(double) * nq1[d]);
    if (w1d && w1d[d]) std::memcpy(host + (kMaxDim + d) * kMaxRule, w1d[d], sizeof(double) * nq1[d]);
  }
  // skip the copy when the same rule is already resident (one launch per call on the steady-state path).  A call without
  // weights (evaluation) matches a resident rule with the same abscissae whatever weights it carries, so that alternating
  // evaluation / assembly calls with one rule never restage.
  const size_t nd = sizeof(host) / sizeof(double), nx = (size_t)kMaxDim * kMaxRule;
  bool resident = p->rule_host.size() == nd && p->rule_stream == stream &&
                  std::memcmp(p->rule_host.data(), host, nx * sizeof(double)) == 0;
  if (resident && w1d) resident = std::memcmp(p->rule_host.data() + nx, host + nx, nx * sizeof(double)) == 0;
  if (!resident) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (stream) cudaStreamIsCapturing(stream, &cap);
    if (cap != cudaStreamCaptureStatusNone) {
      // a copy node would read host memory at replay time: stage the rule once outside the capture instead
      set_error("quadrature rule is not resident on this stream: call once with this rule before capturing a CUDA graph");
      return IGAB_INVALID_ARGUMENT;
    }
    // pageable source: the copy is staged by the runtime before the call returns, so `host` may go out of scope
    IGAB_CUDA_TRY(cudaMemcpyAsync(p->d_rule, host, sizeof(host), cudaMemcpyHostToDevice, stream));
    p->rule_host.assign(host, host + nd);
    p->rule_stream = stream;
  }
  for (int d = 0; d < kMaxDim; ++d) {
    r->xi[d] = p->d_rule + d * kMaxRule;
    r->w[d] = p->d_rule + (kMaxDim + d) * kMaxRule;
  }
  return IGAB_OK;
}

static void out_sizes(const PatchDev& P, long long npts, long long sz[8]) {
  const int nh = P.dim * (P.dim + 1) / 2;
  sz[0] = npts * P.nb;
  sz[1] = npts * P.nb * P.dim;
  sz[2] = npts * P.nb * nh;
  sz[3] = npts * P.dw;
  sz[4] = npts * P.dim * P.dw;
  sz[5] = npts * nh * P.dw;
  sz[6] = npts;
  sz[7] = npts * P.dim * P.dw;
}
static double** out_field(igab_eval_out* o, int k) {
  double** f[8] = {&o->N, &o->dN, &o->d2N, &o->x, &o->Jt, &o->H, &o->detJ, &o->JinvT};
  return f[k];
}
static igab_status check_out(const igab_eval_out* out, int order, igab_eval_out* eff) {
  if (!out) {
    set_error("eval: null output set");
    return IGAB_INVALID_ARGUMENT;
  }
  if (order < 0 || order > 2) {
    set_error("eval: deriv_order must be 0, 1 or 2");
    return IGAB_INVALID_ARGUMENT;
  }
  *eff = *out;
  if (order < 2) eff->d2N = eff->H = nullptr;
  if (order < 1) eff->dN = eff->Jt = eff->detJ = eff->JinvT = nullptr;
  return IGAB_OK;
}
static EvalOutDev to_dev(const igab_eval_out& o) { return EvalOutDev{o.N, o.dN, o.d2N, o.x, o.Jt, o.H, o.detJ, o.JinvT}; }

static igab_status run_eval_tensor_device(igab_patch* p, long long eb, long long ee, int order, const int* nq1,
                                          const double* const* xi1d_host, const RuleDev& rule,
                                          const igab_eval_out& out, cudaStream_t stream) {
  PointSource src{};
  src.tensor = 1;
  src.elem_begin = eb;
  src.nq = 1;
  for (int d = 0; d < p->dev.dim; ++d) {
    src.nq1[d] = nq1[d];
    src.nq *= nq1[d];
    src.xi1d[d] = rule.xi[d];
  }
  const EvalOutDev od = to_dev(out);
  const long long nel = ee - eb;
  if (nel == 0) return IGAB_OK;
  const bool tuned = p->kernel_mode == IGAB_KERNEL_AUTO || p->kernel_mode == IGAB_KERNEL_TENSOR;
  bool sf = tuned && eval_tensor_sf_supported(p->dev, nq1, order, od);
  const bool t2 = tuned && eval_tensor_2d_supported(p->dev, nq1, order, od);
  if (t2) return launch_eval_tensor_2d(p->dev, p->sf, src, xi1d_host, nel, order, od, stream);
  if (tuned && !getenv("IGAB_NO_BLK") && eval_tensor_blk_supported(p->dev, nq1, order, od))
    return launch_eval_tensor_blk(p->dev, p->sf, src, xi1d_host, nel, order, od, stream);
  if (tuned && eval_tensor_sf2_supported(p->dev, nq1, order, od))
    return launch_eval_tensor_sf2(p->dev, p->sf, src, xi1d_host, nel, od, stream);
  if (p->kernel_mode == IGAB_KERNEL_TENSOR && !sf) {
    set_error("eval_tensor: no sum-factorised kernel compiled for this (dim, dimworld, degree, rule, outputs)");
    return IGAB_UNSUPPORTED;
  }
  if (sf) return launch_eval_tensor_sf(p->dev, p->sf, src, xi1d_host, nel, order, od, stream);
  // no tuned instantiation: the general warp-per-element kernel (run-time degrees and rule) ...
  if (p->kernel_mode != IGAB_KERNEL_GENERIC) {
    EvalOutDev ow = od;
    if (order < 2) ow.d2N = ow.H = nullptr;
    if (order < 1) ow.dN = ow.Jt = ow.detJ = ow.JinvT = nullptr;
    const igab_status st = launch_eval_warp(p->dev, src, nel, order, ow, stream);
    if (st != IGAB_UNSUPPORTED) return st;
    if (p->kernel_mode == IGAB_KERNEL_WARP) {
      set_error("eval_tensor: the element's working set does not fit the warp-per-element kernel");
      return IGAB_UNSUPPORTED;
    }
  }
  // ... and, where even that does not fit (or on request), the thread-per-point kernel in bounded launches
  const long long maxPts = 1LL << 30;
  const long long elPer = std::max(1LL, maxPts / src.nq);
  for (long long b = 0; b < nel; b += elPer) {
    const long long n = std::min(elPer, nel - b);
    PointSource s2 = src;
    s2.elem_begin = eb + b;
    igab_eval_out o2 = out;
    long long sz[8];
    out_sizes(p->dev, b * src.nq, sz);
    for (int k = 0; k < 8; ++k)
      if (*out_field(&o2, k)) *out_field(&o2, k) += sz[k];
    igab_status st = launch_eval_generic(p->dev, s2, n * src.nq, order, to_dev(o2), stream);
    if (st) return st;
  }
  return IGAB_OK;
}

// point lists: tuned 2-D list kernel (p <= 3), else the warp-per-32-points kernel, else thread per point
static igab_status run_eval_points_device(igab_patch* p, const PointSource& src, long long npts, int order,
                                          const EvalOutDev& out, cudaStream_t stream) {
  const bool tuned = p->kernel_mode == IGAB_KERNEL_AUTO || p->kernel_mode == IGAB_KERNEL_TENSOR;
  if (tuned && eval_points_2d_supported(p->dev)) return launch_eval_points_2d(p->dev, src, npts, order, out, stream);
  if (p->kernel_mode != IGAB_KERNEL_GENERIC) {
    const igab_status st = launch_eval_warp(p->dev, src, npts, order, out, stream);
    if (st != IGAB_UNSUPPORTED) return st;
  }
  return launch_eval_generic(p->dev, src, npts, order, out, stream);
}

// Are the weights of a 3-D net a tensor product w(i,j,k) = a_i b_j c_k (to 1e-14 relative)?  Then the factors stay on the
// device (PatchDev::wfac) and the assembly kernels fold them into their univariate tables.  Checked ON THE DEVICE and
// lazily -- by the first assembly call after the net was set -- so that evaluation-only callers and
// igab_patch_update_control_points pay nothing (a host pass over the 2.2 M weights of config C2 costs 10 ms).
__global__ void wfac_extract_kernel(PatchDev P, double* a, double* b, double* c) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n0 = P.ncp[0], n1 = P.ncp[1], n2 = P.ncp[2];
  const double w000 = P.cp[3];
  if (t < n0) a[t] = P.cp[(size_t)t * 4 + 3];
  if (t < n1) b[t] = P.cp[(size_t)t * n0 * 4 + 3] / w000;
  if (t < n2) c[t] = P.cp[(size_t)t * n0 * n1 * 4 + 3] / w000;
}
__global__ void wfac_check_kernel(PatchDev P, const double* __restrict__ a, const double* __restrict__ b,
                                  const double* __restrict__ c, long long n, int* __restrict__ bad) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const int i = (int)(g % P.ncp[0]), j = (int)((g / P.ncp[0]) % P.ncp[1]), k = (int)(g / ((long long)P.ncp[0] * P.ncp[1]));
  const double w = P.cp[g * 4 + 3];
  if (!(w > 0.0) || !(fabs(w - a[i] * (b[j] * c[k])) <= 1e-14 * fabs(w))) *bad = 1;
}
igab_status detect_tensor_product_weights(igab_patch* p, cudaStream_t stream) {
  PatchDev& D = p->dev;
  p->wfac_known = true;
  for (int d = 0; d < kMaxDim; ++d) D.wfac[d] = nullptr;
  if (D.dim != 3 || D.dw != 3 || getenv("IGAB_NO_WTAB")) return IGAB_OK;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (stream) cudaStreamIsCapturing(stream, &cap);
  if (cap != cudaStreamCaptureStatusNone) {  // no synchronisation inside a capture: general path, decide after the capture
    p->wfac_known = false;
    return IGAB_OK;
  }
  for (int d = 0; d < 3; ++d)
    if (!p->d_wfac[d]) IGAB_CUDA_TRY(cudaMalloc(&p->d_wfac[d], sizeof(double) * D.ncp[d]));
  int* d_bad = nullptr;
  IGAB_CUDA_TRY(cudaMallocAsync(&d_bad, sizeof(int), stream));
  IGAB_CUDA_TRY(cudaMemsetAsync(d_bad, 0, sizeof(int), stream));
  const int nmax = std::max(D.ncp[0], std::max(D.ncp[1], D.ncp[2]));
  wfac_extract_kernel<<<(nmax + 127) / 128, 128, 0, stream>>>(D, p->d_wfac[0], p->d_wfac[1], p->d_wfac[2]);
  wfac_check_kernel<<<(unsigned)((p->ncp_total + 255) / 256), 256, 0, stream>>>(D, p->d_wfac[0], p->d_wfac[1], p->d_wfac[2],
                                                                              p->ncp_total, d_bad);
  count_launch(2);
  int bad = 1;
  cudaError_t e = cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cudaFreeAsync(d_bad, stream);
  if (e != cudaSuccess) return cuda_fail(e, "tensor-product weight check");
  if (!bad)
    for (int d = 0; d < 3; ++d) D.wfac[d] = p->d_wfac[d];
  return IGAB_OK;
}

}  // namespace igab

using namespace igab;

// =====================================================================================================================
extern "C" {

const char* igab_last_error_string(void) { return g_last_error.c_str(); }
const char* igab_version(void) { return "igab200 0.1 (sm_100a)"; }
int64_t igab_launch_count(void) { return g_launches.load(); }

namespace {
// CPUs of the NUMA node the current CUDA device hangs off (sysfs), empty when unknown / single-node
bool device_numa_cpus(cpu_set_t* set, int* node_out) {
  int dev = 0;
  char bus[32] = {0};
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
  char path[128];
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE* f = fopen(path, "r");
  if (!f) return false;
  int node = -1;
  const int got = fscanf(f, "%d", &node);
  fclose(f);
  if (got != 1 || node < 0) return false;
  snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
  f = fopen(path, "r");
  if (!f) return false;
  char list[4096] = {0};
  const bool ok = fgets(list, sizeof(list), f) != nullptr;
  fclose(f);
  if (!ok) return false;
  CPU_ZERO(set);
  int n = 0;
  for (char* tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
    int a = 0, b = 0;
    const int k = sscanf(tok, "%d-%d", &a, &b);
    if (k == 1) b = a;
    if (k < 1) continue;
    for (int c = a; c <= b && c < CPU_SETSIZE; ++c) {
      CPU_SET(c, set);
      ++n;
    }
  }
  if (node_out) *node_out = node;
  return n > 0;
}
}  // namespace

igab_status igab_host_alloc(void** ptr, uint64_t bytes) {
  if (!ptr) return IGAB_INVALID_ARGUMENT;
  // Page-locked memory is populated at allocation time on the NUMA node of the CALLING thread.  With one process per GPU
  // on a two-socket box that is whatever node the launcher left the process on (usually node 0 for every rank), and the
  // D2H streams of the GPUs on the other socket then cross the inter-socket link.  Allocate from a CPU of the node the
  // current device hangs off, then restore the caller's affinity.
  cpu_set_t want, old;
  const bool rebind = !getenv("IGAB_NO_NUMA_BIND") && device_numa_cpus(&want, nullptr) &&
                      sched_getaffinity(0, sizeof(old), &old) == 0 && sched_setaffinity(0, sizeof(want), &want) == 0;
  const cudaError_t e = cudaMallocHost(ptr, bytes);
  if (rebind) sched_setaffinity(0, sizeof(old), &old);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocHost");
  return IGAB_OK;
}
int igab_device_numa_node(void) {
  cpu_set_t s;
  int node = -1;
  return device_numa_cpus(&s, &node) ? node : -1;
}
igab_status igab_host_free(void* ptr) {
  IGAB_CUDA_TRY(cudaFreeHost(ptr));
  return IGAB_OK;
}

static igab_status patch_create_impl(int dim, int dimworld, const int* degree, const int* nknots,
                                     const double* const* knots, const double* cp, igab_memspace cp_space,
                                     const uint8_t* trimflags, igab_patch** out);
igab_status igab_patch_create(int dim, int dimworld, const int* degree, const int* nknots, const double* const* knots,
                              const double* cp, const uint8_t* trimflags, igab_patch** out) {
  return patch_create_impl(dim, dimworld, degree, nknots, knots, cp, IGAB_HOST, trimflags, out);
}
igab_status igab_patch_create_from_device(int dim, int dimworld, const int* degree, const int* nknots,
                                          const double* const* knots, const double* cp_dev, const uint8_t* trimflags,
                                          igab_patch** out) {
  return patch_create_impl(dim, dimworld, degree, nknots, knots, cp_dev, IGAB_DEVICE, trimflags, out);
}
static igab_status patch_create_impl(int dim, int dimworld, const int* degree, const int* nknots,
                                     const double* const* knots, const double* cp, igab_memspace cp_space,
                                     const uint8_t* trimflags, igab_patch** out) {
  if (!out) return IGAB_INVALID_ARGUMENT;
  *out = nullptr;
  if (dim < 1 || dim > kMaxDim || dimworld < dim || dimworld > 3 || !degree || !nknots || !knots || !cp) {
    set_error("patch_create: need 1 <= dim <= dimworld <= 3 and non-null degree/nknots/knots/cp");
    return IGAB_INVALID_ARGUMENT;
  }
  if (!((dim == 1) || (dim == 2 && dimworld >= 2) || (dim == 3 && dimworld == 3))) {
    set_error("patch_create: unsupported (dim, dimworld)");
    return IGAB_UNSUPPORTED;
  }
  igab_patch* p = new (std::nothrow) igab_patch();
  if (!p) return IGAB_OUT_OF_MEMORY;
  PatchDev& D = p->dev;
  D.dim = dim;
  D.dw = dimworld;
  D.nb = 1;
  D.nelem = 1;
  p->ncp_total = 1;
  for (int d = 0; d < dim; ++d) {
    const int pd = degree[d], n = nknots[d];
    if (pd < 1 || pd > kMaxDegree || n < 2 * (pd + 1) || !knots[d]) {
      set_error("patch_create: degree must be 1..8 and the knot vector must hold at least 2(p+1) knots");
      delete p;
      return IGAB_INVALID_ARGUMENT;
    }
    std::vector<double>& U = p->hknots[d];
    U.assign(knots[d], knots[d] + n);
    bool ok = std::is_sorted(U.begin(), U.end());
    for (int i = 0; i <= pd && ok; ++i) ok = (U[i] == U[0]) && (U[n - 1 - i] == U[n - 1]);
    if (!ok || !(U[0] < U[n - 1])) {
      // the reference asserts open knot vectors (bsplinealgorithms.hh:98-99, nurbsgrid.hh:112-122)
      set_error("patch_create: knot vector must be sorted and open (degree+1 equal end knots)");
      delete p;
      return IGAB_INVALID_ARGUMENT;
    }
    // unique knots with FloatCmp::eq (nurbspatch.hh:58-63) and the per-direction element->span table
    std::vector<double> uq;
    for (double v : U)
      if (uq.empty() || !feq8(uq.back(), v)) uq.push_back(v);
    const int nel = (int)uq.size() - 1;
    p->hspan[d].resize(nel);
    for (int e = 0; e < nel; ++e) p->hspan[d][e] = find_span_corrected(pd, uq[e], U, e);
    D.degree[d] = pd;
    D.nknots[d] = n;
    D.ncp[d] = n - pd - 1;
    D.nel[d] = nel;
    D.nb *= pd + 1;
    D.nelem *= nel;
    p->ncp_total *= D.ncp[d];
  }
  p->ndof_untrimmed = p->ncp_total;
  if (trimflags) p->htrim.assign(trimflags, trimflags + D.nelem);

  cudaError_t e = cudaGetDevice(&p->device);
  if (e != cudaSuccess) {
    igab_status st = cuda_fail(e, "cudaGetDevice (no CUDA device: this library has no CPU fallback)");
    delete p;
    return st;
  }
  {
    // keep stream-ordered scratch (cudaMallocAsync) cached in the pool instead of returning it to the driver at every
    // synchronisation: the assembly / reduction paths allocate per call
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, p->device) == cudaSuccess) {
      unsigned long long thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  auto fail = [&](cudaError_t err, const char* what) {
    igab_status st = cuda_fail(err, what);
    igab_patch_destroy(p);
    return st;
  };
  for (int d = 0; d < dim; ++d) {
    if ((e = cudaMalloc(&p->d_knots[d], sizeof(double) * D.nknots[d])) != cudaSuccess) return fail(e, "cudaMalloc knots");
    if ((e = cudaMalloc(&p->d_span[d], sizeof(int) * D.nel[d])) != cudaSuccess) return fail(e, "cudaMalloc spans");
    if ((e = cudaMemcpy(p->d_knots[d], p->hknots[d].data(), sizeof(double) * D.nknots[d], cudaMemcpyHostToDevice)) != cudaSuccess)
      return fail(e, "copy knots");
    if ((e = cudaMemcpy(p->d_span[d], p->hspan[d].data(), sizeof(int) * D.nel[d], cudaMemcpyHostToDevice)) != cudaSuccess)
      return fail(e, "copy spans");
    D.knots[d] = p->d_knots[d];
    D.spanTab[d] = p->d_span[d];
    // every division of A2.3 (bsplinealgorithms.hh:161-213) is by a knot difference U[sp+r+1] - U[sp+1-j+r] of the
    // point's span: their reciprocals, per element of the direction, row j*(j-1)/2 + r  (j = 1..p, r < j)
    const int pd = D.degree[d], tri = std::max(1, pd * (pd + 1) / 2);
    std::vector<double> rd((size_t)D.nel[d] * tri, 0.0);
    for (int el = 0; el < D.nel[d]; ++el) {
      const int sp = p->hspan[d][el];
      for (int j = 1; j <= pd; ++j)
        for (int r = 0; r < j; ++r)
          rd[(size_t)el * tri + j * (j - 1) / 2 + r] = 1.0 / (p->hknots[d][sp + r + 1] - p->hknots[d][sp + 1 - j + r]);
    }
    if ((e = cudaMalloc(&p->d_rden[d], sizeof(double) * rd.size())) != cudaSuccess) return fail(e, "cudaMalloc rden");
    if ((e = cudaMemcpy(p->d_rden[d], rd.data(), sizeof(double) * rd.size(), cudaMemcpyHostToDevice)) != cudaSuccess)
      return fail(e, "copy rden");
    D.rden[d] = p->d_rden[d];
  }
  const size_t cpBytes = sizeof(double) * p->ncp_total * (dimworld + 1);
  if ((e = cudaMalloc(&p->d_cp, cpBytes)) != cudaSuccess) return fail(e, "cudaMalloc control net");
  if ((e = cudaMemcpy(p->d_cp, cp, cpBytes, cp_space == IGAB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice)) !=
      cudaSuccess)
    return fail(e, "copy control net");
  D.cp = p->d_cp;
  for (int d = 0; d < kMaxDim; ++d) D.wfac[d] = nullptr;  // decided lazily by the first assembly call
  if (trimflags) {
    if ((e = cudaMalloc(&p->d_trim, D.nelem)) != cudaSuccess) return fail(e, "cudaMalloc trim flags");
    if ((e = cudaMemcpy(p->d_trim, trimflags, D.nelem, cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e, "copy trim");
  }
  if ((e = cudaMalloc(&p->d_rule, sizeof(double) * 2 * kMaxDim * kMaxRule)) != cudaSuccess) return fail(e, "cudaMalloc rule");
  *out = p;
  return IGAB_OK;
}

void igab_patch_destroy(igab_patch* p) {
  if (!p) return;
  for (int d = 0; d < kMaxDim; ++d) {
    if (p->d_knots[d]) cudaFree(p->d_knots[d]);
    if (p->d_span[d]) cudaFree(p->d_span[d]);
    if (p->d_rden[d]) cudaFree(p->d_rden[d]);
  }
  if (p->d_cp) cudaFree(p->d_cp);
  for (int d = 0; d < kMaxDim; ++d)
    if (p->d_wfac[d]) cudaFree(p->d_wfac[d]);
  if (p->d_trim) cudaFree(p->d_trim);
  if (p->d_rule) cudaFree(p->d_rule);
  if (p->d_red) cudaFree(p->d_red);
  if (p->d_norm) cudaFree(p->d_norm);
  for (int b = 0; b < 2; ++b) {
    if (p->stage_buf[b]) cudaFree(p->stage_buf[b]);
    if (p->stage_done[b]) cudaEventDestroy(p->stage_done[b]);
    if (p->stage_copied[b]) cudaEventDestroy(p->stage_copied[b]);
  }
  if (p->stage_stream) cudaStreamDestroy(p->stage_stream);
  for (int d = 0; d < kMaxDim; ++d)
    if (p->d_elem_of_first[d]) cudaFree(p->d_elem_of_first[d]);
  if (p->d_csr_rowptr) cudaFree(p->d_csr_rowptr);
  if (p->d_csr_colidx) cudaFree(p->d_csr_colidx);
  if (p->ga_scratch) cudaFree(p->ga_scratch);
  if (p->xch_buf) cudaFree(p->xch_buf);
  if (p->stream_switch_ev) cudaEventDestroy(p->stream_switch_ev);
  if (p->d_D) cudaFree(p->d_D);
  for (int b = 0; b < 2; ++b)
    if (p->asm_buf[b]) cudaFree(p->asm_buf[b]);
  if (p->d_dofc_map) cudaFree(p->d_dofc_map);
  if (p->d_dofc_inv) cudaFree(p->d_dofc_inv);
  free_sf_workspace(p->sf);
  delete p;
}

igab_status igab_patch_sizes(const igab_patch* p, int64_t* n_elem, int32_t* n_elem_dir, int32_t* nb,
                             int64_t* ndof_untrimmed) {
  if (!p) return IGAB_INVALID_ARGUMENT;
  if (n_elem) *n_elem = p->dev.nelem;
  if (n_elem_dir)
    for (int d = 0; d < p->dev.dim; ++d) n_elem_dir[d] = p->dev.nel[d];
  if (nb) *nb = p->dev.nb;
  if (ndof_untrimmed) *ndof_untrimmed = p->ndof_untrimmed;
  return IGAB_OK;
}

igab_status igab_patch_set_kernel(igab_patch* p, int kernel) {
  if (!p || kernel < 0 || kernel > 3) return IGAB_INVALID_ARGUMENT;
  p->kernel_mode = kernel;
  return IGAB_OK;
}

igab_status igab_patch_update_control_points(igab_patch* p, const double* cp) {
  if (!p || !cp) return IGAB_INVALID_ARGUMENT;
  // kernels still in flight on the stream of the last call (possibly a non-blocking one) read the control net
  if (p->have_last_stream && cudaStreamSynchronize(p->last_stream) != cudaSuccess) cudaGetLastError();
  IGAB_CUDA_TRY(cudaMemcpy(p->d_cp, cp, sizeof(double) * p->ncp_total * (p->dev.dw + 1), cudaMemcpyHostToDevice));
  p->sf.wvalid = false;  // the weighted tables carry the old weights
  p->wfac_known = false;
  for (int d = 0; d < kMaxDim; ++d) p->dev.wfac[d] = nullptr;
  return IGAB_OK;
}

igab_status igab_patch_spans(igab_patch* p, int32_t* span, igab_memspace space, void* stream_) {
  if (!p || !span) return IGAB_INVALID_ARGUMENT;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (igab_status est = enter_stream(p, stream)) return est;
  if (space == IGAB_DEVICE) return launch_spans(p->dev, span, stream);
  int32_t* d = nullptr;
  const size_t bytes = sizeof(int32_t) * p->dev.nelem * p->dev.dim;
  IGAB_CUDA_TRY(cudaMalloc(&d, bytes));
  igab_status st = launch_spans(p->dev, d, stream);
  if (!st) {
    cudaError_t e = cudaMemcpyAsync(span, d, bytes, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) st = cuda_fail(e, "spans D2H");
  }
  cudaFree(d);
  return st;
}

igab_status igab_patch_dofmap(igab_patch* p, int32_t* dof, int64_t* ndof, igab_memspace space, void* stream_) {
  if (!p || !dof) return IGAB_INVALID_ARGUMENT;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (igab_status est = enter_stream(p, stream)) return est;
  const long long n = p->dev.nelem * p->dev.nb;
  if (n > 0x7fffffffLL * 256LL) {
    set_error("dofmap: too many entries");
    return IGAB_INVALID_ARGUMENT;
  }
  int32_t* d = dof;
  if (space == IGAB_HOST) IGAB_CUDA_TRY(cudaMalloc(&d, sizeof(int32_t) * n));
  igab_status st = launch_dofmap(p->dev, d, stream);
  long long total = p->ndof_untrimmed;
  int32_t *used = nullptr, *map = nullptr;
  long long* total_dev = nullptr;
  if (!st && p->d_trim) {
    cudaError_t e = cudaMalloc(&used, sizeof(int32_t) * p->ndof_untrimmed);
    if (e == cudaSuccess) e = cudaMalloc(&map, sizeof(int32_t) * p->ndof_untrimmed);
    if (e == cudaSuccess) e = cudaMalloc(&total_dev, sizeof(long long));
    if (e != cudaSuccess) st = cuda_fail(e, "dofmap scratch");
    if (!st) st = launch_dof_compact(p->dev, p->d_trim, d, used, map, p->ndof_untrimmed, total_dev, stream);
    if (!st) {
      e = cudaMemcpyAsync(&total, total_dev, sizeof(long long), cudaMemcpyDeviceToHost, stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
      if (e != cudaSuccess) st = cuda_fail(e, "dofmap total D2H");
    }
  }
  if (!st && space == IGAB_HOST) {
    cudaError_t e = cudaMemcpyAsync(dof, d, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) st = cuda_fail(e, "dofmap D2H");
  }
  if (used) cudaFree(used);
  if (map) cudaFree(map);
  if (total_dev) cudaFree(total_dev);
  if (space == IGAB_HOST && d) cudaFree(d);
  if (!st && ndof) *ndof = total;
  return st;
}

igab_status igab_patch_real_index_maps(igab_patch* p, int32_t* direct_of_real, int32_t* real_of_direct, int64_t* n_real,
                                       igab_memspace space, void* stream_) {
  if (!p) return IGAB_INVALID_ARGUMENT;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (igab_status est = enter_stream(p, stream)) return est;
  const long long n = p->dev.nelem;
  if (!p->d_trim) {  // untrimmed patch: identity (nurbspatch.hh:599-606)
    if (n_real) *n_real = n;
    if (space == IGAB_HOST) {
      for (long long e = 0; e < n; ++e) {
        if (direct_of_real) direct_of_real[e] = (int32_t)e;
        if (real_of_direct) real_of_direct[e] = (int32_t)e;
      }
      return IGAB_OK;
    }
    std::vector<int32_t> id(n);
    for (long long e = 0; e < n; ++e) id[e] = (int32_t)e;
    if (direct_of_real) IGAB_CUDA_TRY(cudaMemcpyAsync(direct_of_real, id.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, stream));
    if (real_of_direct) IGAB_CUDA_TRY(cudaMemcpyAsync(real_of_direct, id.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, stream));
    IGAB_CUDA_TRY(cudaStreamSynchronize(stream));
    return IGAB_OK;
  }
  int32_t *dMap = real_of_direct, *dInv = direct_of_real;
  long long* dTotal = nullptr;
  int32_t* tmp = nullptr;
  if (space == IGAB_HOST) {
    IGAB_CUDA_TRY(cudaMalloc(&tmp, sizeof(int32_t) * 2 * std::max(1LL, n)));
    dMap = tmp;
    dInv = tmp + n;
  }
  cudaError_t e = cudaMalloc(&dTotal, sizeof(long long));
  igab_status st = e == cudaSuccess ? IGAB_OK : cuda_fail(e, "real_index_maps scratch");
  if (!st) st = launch_flag_scan(nullptr, p->d_trim, n, dMap, dInv, dTotal, stream);
  long long total = 0;
  if (!st) {
    e = cudaMemcpyAsync(&total, dTotal, sizeof(long long), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess && space == IGAB_HOST) {
      if (real_of_direct) e = cudaMemcpy(real_of_direct, dMap, sizeof(int32_t) * n, cudaMemcpyDeviceToHost);
      if (e == cudaSuccess && direct_of_real) e = cudaMemcpy(direct_of_real, dInv, sizeof(int32_t) * total, cudaMemcpyDeviceToHost);
    }
    if (e != cudaSuccess) st = cuda_fail(e, "real_index_maps D2H");
  }
  if (dTotal) cudaFree(dTotal);
  if (tmp) cudaFree(tmp);
  if (!st && n_real) *n_real = total;
  return st;
}

// ---- evaluation -----------------------------------------------------------------------------------------------------
static igab_status eval_tensor_impl(igab_patch* p, int64_t elem_begin, int64_t elem_end, int deriv_order, const int* nq1,
                                    const double* const* xi1d, const igab_eval_out* out, igab_memspace out_space,
                                    void* stream_, bool async_host) {
  if (!p || !nq1 || !xi1d) {
    set_error("eval_tensor: null argument");
    return IGAB_INVALID_ARGUMENT;
  }
  if (elem_begin < 0 || elem_end < elem_begin || elem_end > p->dev.nelem) {
    set_error("eval_tensor: element range outside [0, n_elem]");
    return IGAB_INVALID_ARGUMENT;
  }
  igab_eval_out eff;
  igab_status st = check_out(out, deriv_order, &eff);
  if (st) return st;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (igab_status est = enter_stream(p, stream)) return est;
  RuleDev rule;
  if ((st = stage_rule(p, nq1, xi1d, nullptr, &rule, stream))) return st;
  if (out_space == IGAB_DEVICE) return run_eval_tensor_device(p, elem_begin, elem_end, deriv_order, nq1, xi1d, rule, eff, stream);

  // HOST outputs: element chunks through two sets of device buffers; the D2H copy of chunk k overlaps the kernel of
  // chunk k+1 (copy stream + events).  Full PCIe rate needs pinned destination memory (igab_host_alloc).
  NoPdlScope plain_launches;
  long long nq = 1;
  for (int d = 0; d < p->dev.dim; ++d) nq *= nq1[d];
  long long per_pt[8], bytes_pt = 0;
  out_sizes(p->dev, 1, per_pt);
  for (int k = 0; k < 8; ++k)
    if (*out_field(&eff, k)) bytes_pt += per_pt[k] * 8;
  const long long nel = elem_end - elem_begin;
  if (nel == 0 || bytes_pt == 0) return IGAB_OK;
  const long long chunkBytes = 512LL << 20;  // per buffer set
  const long long elChunk = std::max(1LL, std::min(nel, chunkBytes / (bytes_pt * nq)));
  const size_t needBytes = (size_t)(bytes_pt * nq * elChunk) + 8 * 16;
  cudaError_t e = cudaSuccess;
  if (!p->stage_stream) {
    e = cudaStreamCreateWithFlags(&p->stage_stream, cudaStreamNonBlocking);
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
      e = cudaEventCreateWithFlags(&p->stage_done[b], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->stage_copied[b], cudaEventDisableTiming);
    }
  }
  if (e == cudaSuccess && p->stage_cap < needBytes) {
    for (int b = 0; b < 2; ++b) {
      if (p->stage_buf[b]) cudaFree(p->stage_buf[b]);
      p->stage_buf[b] = nullptr;
    }
    p->stage_cap = 0;
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) e = cudaMalloc(&p->stage_buf[b], needBytes);
    if (e == cudaSuccess) p->stage_cap = needBytes;
  }
  if (e != cudaSuccess) return cuda_fail(e, "eval_tensor host staging allocation");
  cudaStream_t cstream = p->stage_stream;
  int b = 0;
  for (long long c0 = 0; c0 < nel && !st; c0 += elChunk, b ^= 1) {
    const long long n = std::min(elChunk, nel - c0);
    igab_eval_out dv{};
    {
      double* cur = p->stage_buf[b];  // fields packed back to back (each a multiple of 8 bytes; keep 16-byte alignment)
      for (int k = 0; k < 8; ++k)
        if (*out_field(&eff, k)) {
          *out_field(&dv, k) = cur;
          cur += ((per_pt[k] * nq * elChunk + 1) & ~1LL);
        }
    }
    if ((e = cudaStreamWaitEvent(stream, p->stage_copied[b], 0)) != cudaSuccess) { st = cuda_fail(e, "wait copied"); break; }
    st = run_eval_tensor_device(p, elem_begin + c0, elem_begin + c0 + n, deriv_order, nq1, xi1d, rule, dv, stream);
    if (st) break;
    if ((e = cudaEventRecord(p->stage_done[b], stream)) != cudaSuccess) { st = cuda_fail(e, "record done"); break; }
    if ((e = cudaStreamWaitEvent(cstream, p->stage_done[b], 0)) != cudaSuccess) { st = cuda_fail(e, "wait done"); break; }
    for (int k = 0; k < 8; ++k) {
      double* h = *out_field(&eff, k);
      if (!h) continue;
      e = cudaMemcpyAsync(h + per_pt[k] * nq * c0, *out_field(&dv, k), per_pt[k] * 8 * nq * n, cudaMemcpyDeviceToHost, cstream);
      if (e != cudaSuccess) { st = cuda_fail(e, "eval_tensor D2H"); break; }
    }
    if (st) break;
    if ((e = cudaEventRecord(p->stage_copied[b], cstream)) != cudaSuccess) { st = cuda_fail(e, "record copied"); break; }
  }
  if (async_host && !st) {  // igab_eval_tensor_begin: the caller collects the result with igab_patch_wait
    p->pending_host_stream = stream;
    p->pending_host = true;
    return IGAB_OK;
  }
  e = cudaStreamSynchronize(cstream);
  if (e != cudaSuccess && !st) st = cuda_fail(e, "sync copy stream");
  e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess && !st) st = cuda_fail(e, "sync stream");
  return st;
}

igab_status igab_eval_tensor(igab_patch* p, int64_t elem_begin, int64_t elem_end, int deriv_order, const int* nq1,
                             const double* const* xi1d, const igab_eval_out* out, igab_memspace out_space,
                             void* stream_) {
  if (p && p->pending_host) {
    if (igab_status st = igab_patch_wait(p)) return st;
  }
  return eval_tensor_impl(p, elem_begin, elem_end, deriv_order, nq1, xi1d, out, out_space, stream_, false);
}

igab_status igab_eval_tensor_begin(igab_patch* p, int64_t elem_begin, int64_t elem_end, int deriv_order, const int* nq1,
                                   const double* const* xi1d, const igab_eval_out* out, void* stream_) {
  if (p && p->pending_host) {
    if (igab_status st = igab_patch_wait(p)) return st;
  }
  return eval_tensor_impl(p, elem_begin, elem_end, deriv_order, nq1, xi1d, out, IGAB_HOST, stream_, true);
}

igab_status igab_patch_wait(igab_patch* p) {
  if (!p) return IGAB_INVALID_ARGUMENT;
  if (!p->pending_host) return IGAB_OK;
  p->pending_host = false;
  cudaError_t e = p->stage_stream ? cudaStreamSynchronize(p->stage_stream) : cudaSuccess;
  if (e == cudaSuccess) e = cudaStreamSynchronize(p->pending_host_stream);
  if (e != cudaSuccess) return cuda_fail(e, "patch_wait");
  return IGAB_OK;
}

igab_status igab_eval_points(igab_patch* p, int64_t npts, const int32_t* elem, const double* xi, int deriv_order,
                             const igab_eval_out* out, igab_memspace space, void* stream_) {
  if (!p || npts < 0 || (npts > 0 && (!elem || !xi))) {
    set_error("eval_points: null argument");
    return IGAB_INVALID_ARGUMENT;
  }
  igab_eval_out eff;
  igab_status st = check_out(out, deriv_order, &eff);
  if (st) return st;
  if (npts == 0) return IGAB_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (igab_status est = enter_stream(p, stream)) return est;
  PointSource src{};
  src.tensor = 0;
  if (space == IGAB_DEVICE) {
    src.elem = elem;
    src.xi = xi;
    return run_eval_points_device(p, src, npts, deriv_order, to_dev(eff), stream);
  }
  for (long long k = 0; k < npts; ++k)
    if (elem[k] < 0 || elem[k] >= p->dev.nelem) {
      set_error("eval_points: element index out of range");
      return IGAB_INVALID_ARGUMENT;
    }
  int* d_elem = nullptr;
  double* d_xi = nullptr;
  double* dbuf[8] = {};
  long long sz[8];
  out_sizes(p->dev, npts, sz);
  cudaError_t e = cudaMalloc(&d_elem, sizeof(int) * npts);
  if (e == cudaSuccess) e = cudaMalloc(&d_xi, sizeof(double) * npts * p->dev.dim);
  for (int k = 0; k < 8 && e == cudaSuccess; ++k)
    if (*out_field(&eff, k)) e = cudaMalloc(&dbuf[k], sz[k] * 8);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_elem, elem, sizeof(int) * npts, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_xi, xi, sizeof(double) * npts * p->dev.dim, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) st = cuda_fail(e, "eval_points staging");
  if (!st) {
    src.elem = d_elem;
    src.xi = d_xi;
    igab_eval_out dv{};
    for (int k = 0; k < 8; ++k) *out_field(&dv, k) = dbuf[k];
    st = run_eval_points_device(p, src, npts, deriv_order, to_dev(dv), stream);
  }
  for (int k = 0; k < 8 && !st; ++k)
    if (dbuf[k]) {
      e = cudaMemcpyAsync(*out_field(&eff, k), dbuf[k], sz[k] * 8, cudaMemcpyDeviceToHost, stream);
      if (e != cudaSuccess) st = cuda_fail(e, "eval_points D2H");
    }
  e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess && !st) st = cuda_fail(e, "eval_points sync");
  if (d_elem) cudaFree(d_elem);
  if (d_xi) cudaFree(d_xi);
  for (int k = 0; k < 8; ++k)
    if (dbuf[k]) cudaFree(dbuf[k]);
  return st;
}

igab_status igab_partial(igab_patch* p, int64_t npts, const int32_t* elem, const double* xi, const int* alpha,
                         double* out, igab_memspace space, void* stream_) {
  if (!p || !alpha || npts < 0 || (npts > 0 && (!elem || !xi || !out))) {
    set_error("partial: null argument");
    return IGAB_INVALID_ARGUMENT;
  }
  for (int d = 0; d < p->dev.dim; ++d)
    if (alpha[d] < 0 || alpha[d] > 15) {
      set_error("partial: derivative order per direction must be 0..15");
      return IGAB_INVALID_ARGUMENT;
    }
  if (npts == 0) return IGAB_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (igab_status est = enter_stream(p, stream)) return est;
  if (space == IGAB_DEVICE) return launch_partial(p->dev, npts, elem, xi, alpha, out, stream);
  for (long long k = 0; k < npts; ++k)
    if (elem[k] < 0 || elem[k] >= p->dev.nelem) {
      set_error("partial: element index out of range");
      return IGAB_INVALID_ARGUMENT;
    }
  int* d_elem = nullptr;
  double *d_xi = nullptr, *d_out = nullptr;
  igab_status st = IGAB_OK;
  cudaError_t e = cudaMalloc(&d_elem, sizeof(int) * npts);
  if (e == cudaSuccess) e = cudaMalloc(&d_xi, sizeof(double) * npts * p->dev.dim);
  if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(double) * npts * p->dev.nb);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_elem, elem, sizeof(int) * npts, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_xi, xi, sizeof(double) * npts * p->dev.dim, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) st = cuda_fail(e, "partial staging");
  if (!st) st = launch_partial(p->dev, npts, d_elem, d_xi, alpha, d_out, stream);
  if (!st) {
    e = cudaMemcpyAsync(out, d_out, sizeof(double) * npts * p->dev.nb, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) st = cuda_fail(e, "partial D2H");
  }
  if (d_elem) cudaFree(d_elem);
  if (d_xi) cudaFree(d_xi);
  if (d_out) cudaFree(d_out);
  return st;
}

// ---- element matrices -----------------------------------------------------------------------------------------------
igab_status igab_assemble(igab_patch* p, int kind, const double* D, double fconst, int64_t elem_begin,
                          int64_t elem_end, const int* nq1, const double* const* xi1d, const double* const* w1d,
                          double* Ke, double* fe, igab_memspace out_space, void* stream_) {
  if (!p || !nq1 || !xi1d || !w1d || !Ke) {
    set_error("assemble: null argument");
    return IGAB_INVALID_ARGUMENT;
  }
  if (kind < 0 || kind > 2) {
    set_error("assemble: unknown kind");
    return IGAB_INVALID_ARGUMENT;
  }
  if (elem_begin < 0 || elem_end < elem_begin || elem_end > p->dev.nelem) {
    set_error("assemble: element range outside [0, n_elem]");
    return IGAB_INVALID_ARGUMENT;
  }
  if (kind == IGAB_ASM_ELASTICITY && (p->dev.dim != p->dev.dw || p->dev.dim < 2 || !D)) {
    set_error("assemble: elasticity needs dim == dimworld in {2,3} and a D matrix");
    return IGAB_INVALID_ARGUMENT;
  }
  if (kind == IGAB_ASM_POISSON && p->dev.dim != p->dev.dw && p->dev.dim == 3) {
    set_error("assemble: unsupported dims");
    return IGAB_UNSUPPORTED;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  if (igab_status est = enter_stream(p, stream)) return est;
  RuleDev rule;
  igab_status st = stage_rule(p, nq1, xi1d, w1d, &rule, stream);
  if (st) return st;
  const long long nel = elem_end - elem_begin;
  if (nel == 0) return IGAB_OK;
  if (!p->wfac_known && (st = detect_tensor_product_weights(p, stream))) return st;
  const int n = p->dev.nb * (kind == IGAB_ASM_ELASTICITY ? p->dev.dim : 1);
  const int ns = p->dev.dim == 2 ? 3 : 6;
  double* d_D = nullptr;
  cudaError_t e;
  if (kind == IGAB_ASM_ELASTICITY) {
    // D lives in the handle and is restaged only when it (or the stream) changes: no allocation, no synchronisation,
    // capturable in a CUDA graph once resident -- the same contract as the quadrature rule
    if (!p->d_D) IGAB_CUDA_TRY(cudaMalloc(&p->d_D, sizeof(double) * 36));
    const bool resident = p->D_host.size() == (size_t)(ns * ns) && p->D_stream == stream &&
                          std::memcmp(p->D_host.data(), D, sizeof(double) * ns * ns) == 0;
    if (!resident) {
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      if (stream) cudaStreamIsCapturing(stream, &cap);
      if (cap != cudaStreamCaptureStatusNone) {
        set_error("material matrix D is not resident on this stream: call once with this D before capturing a CUDA graph");
        return IGAB_INVALID_ARGUMENT;
      }
      p->D_host.assign(D, D + ns * ns);
      // pageable source: staged by the runtime before the call returns
      IGAB_CUDA_TRY(cudaMemcpyAsync(p->d_D, p->D_host.data(), sizeof(double) * ns * ns, cudaMemcpyHostToDevice, stream));
      p->D_stream = stream;
    }
    d_D = p->d_D;
  }
  if (out_space == IGAB_DEVICE)
    return launch_assemble(p->dev, p->sf, xi1d, kind, D, d_D, fconst, elem_begin, nel, nq1, rule.xi, rule.w, Ke, fe, stream);
  // HOST outputs: element chunks through two device buffers kept in the handle; the D2H copy of chunk k (copy stream)
  // overlaps the kernel of chunk k+1.  Full PCIe rate needs pinned destination memory (igab_host_alloc).
  const long long keD = (long long)n * n, feD = fe ? n : 0;
  long long elChunk = std::max(1LL, std::min(nel, (512LL << 20) / (8 * (keD + feD))));
  if (const char* env = getenv("IGAB_ASM_CHUNK_ELEMS")) elChunk = std::max(1L, atol(env));  // test knob
  const size_t need = (size_t)((keD + feD) * elChunk);
  e = cudaSuccess;
  if (!p->stage_stream) {
    e = cudaStreamCreateWithFlags(&p->stage_stream, cudaStreamNonBlocking);
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
      e = cudaEventCreateWithFlags(&p->stage_done[b], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->stage_copied[b], cudaEventDisableTiming);
    }
  } else if (!p->stage_done[0]) {
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
      e = cudaEventCreateWithFlags(&p->stage_done[b], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->stage_copied[b], cudaEventDisableTiming);
    }
  }
  if (e == cudaSuccess && p->asm_cap < need) {
    cudaStreamSynchronize(stream);
    for (int b = 0; b < 2; ++b) {
      if (p->asm_buf[b]) cudaFree(p->asm_buf[b]);
      p->asm_buf[b] = nullptr;
    }
    p->asm_cap = 0;
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) e = cudaMalloc(&p->asm_buf[b], sizeof(double) * need);
    if (e == cudaSuccess) p->asm_cap = need;
  }
  if (e != cudaSuccess) return cuda_fail(e, "assemble host staging allocation");
  cudaStream_t cstream = p->stage_stream;
  int b = 0;
  for (long long c0 = 0; c0 < nel && !st; c0 += elChunk, b ^= 1) {
    const long long m = std::min(elChunk, nel - c0);
    double* dK = p->asm_buf[b];
    double* dF = fe ? dK + keD * elChunk : nullptr;
    if ((e = cudaStreamWaitEvent(stream, p->stage_copied[b], 0)) != cudaSuccess) { st = cuda_fail(e, "assemble: wait copied"); break; }
    st = launch_assemble(p->dev, p->sf, xi1d, kind, D, d_D, fconst, elem_begin + c0, m, nq1, rule.xi, rule.w, dK, dF, stream);
    if (st) break;
    if ((e = cudaEventRecord(p->stage_done[b], stream)) != cudaSuccess) { st = cuda_fail(e, "assemble: record done"); break; }
    if ((e = cudaStreamWaitEvent(cstream, p->stage_done[b], 0)) != cudaSuccess) { st = cuda_fail(e, "assemble: wait done"); break; }
    e = cudaMemcpyAsync(Ke + keD * c0, dK, sizeof(double) * keD * m, cudaMemcpyDeviceToHost, cstream);
    if (e == cudaSuccess && fe) e = cudaMemcpyAsync(fe + (long long)n * c0, dF, sizeof(double) * n * m, cudaMemcpyDeviceToHost, cstream);
    if (e == cudaSuccess) e = cudaEventRecord(p->stage_copied[b], cstream);
    if (e != cudaSuccess) st = cuda_fail(e, "assemble D2H");
  }
  e = cudaStreamSynchronize(cstream);
  if (e != cudaSuccess && !st) st = cuda_fail(e, "assemble: sync copy stream");
  e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess && !st) st = cuda_fail(e, "assemble: sync stream");
  return st;
}

// ---- refinement -------------------------------------------------------------------------------------------------------
igab_status igab_refine_apply(int dim, int dimworld, const int* ncp_old, int direction, int n_new, int maxlen,
                              const int32_t* first, const int32_t* len, const double* coef, const double* cp_in,
                              double* cp_out) {
  if (dim < 1 || dim > kMaxDim || dimworld < 1 || dimworld > 3 || !ncp_old || direction < 0 || direction >= dim ||
      n_new < 1 || maxlen < 1 || !first || !len || !coef || !cp_in || !cp_out) {
    set_error("refine_apply: invalid argument");
    return IGAB_INVALID_ARGUMENT;
  }
  long long nin = 1, nout = 1;
  for (int d = 0; d < dim; ++d) {
    nin *= ncp_old[d];
    nout *= (d == direction) ? n_new : ncp_old[d];
  }
  for (int i = 0; i < n_new; ++i)
    if (len[i] < 0 || len[i] > maxlen || first[i] < 0 || first[i] + len[i] > ncp_old[direction]) {
      set_error("refine_apply: operator row outside the old net");
      return IGAB_INVALID_ARGUMENT;
    }
  int *d_first = nullptr, *d_len = nullptr;
  double *d_coef = nullptr, *d_in = nullptr, *d_out = nullptr;
  const size_t row = sizeof(double) * (dimworld + 1);
  igab_status st = IGAB_OK;
  cudaError_t e = cudaMalloc(&d_first, sizeof(int) * n_new);
  if (e == cudaSuccess) e = cudaMalloc(&d_len, sizeof(int) * n_new);
  if (e == cudaSuccess) e = cudaMalloc(&d_coef, sizeof(double) * n_new * maxlen);
  if (e == cudaSuccess) e = cudaMalloc(&d_in, row * nin);
  if (e == cudaSuccess) e = cudaMalloc(&d_out, row * nout);
  if (e == cudaSuccess) e = cudaMemcpy(d_first, first, sizeof(int) * n_new, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_len, len, sizeof(int) * n_new, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_coef, coef, sizeof(double) * n_new * maxlen, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_in, cp_in, row * nin, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) st = cuda_fail(e, "refine_apply staging");
  if (!st) st = refine_apply(dim, dimworld, ncp_old, direction, n_new, maxlen, d_first, d_len, d_coef, d_in, d_out, nullptr);
  if (!st) {
    e = cudaMemcpy(cp_out, d_out, row * nout, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) st = cuda_fail(e, "refine_apply D2H");
  }
  if (d_first) cudaFree(d_first);
  if (d_len) cudaFree(d_len);
  if (d_coef) cudaFree(d_coef);
  if (d_in) cudaFree(d_in);
  if (d_out) cudaFree(d_out);
  return st;
}

// A chain of banded 1-D operators (knot insertion, degree elevation, any linear refinement) applied direction by direction
// with every intermediate net in HBM; only the operators (n_new x maxlen doubles each) cross PCIe.
igab_status igab_refine_operators(int dim, int dimworld, const int* ncp_old, int nops, const int* direction,
                                  const int* n_new, const int* maxlen, const int32_t* const* first,
                                  const int32_t* const* len, const double* const* coef, const double* cp_in,
                                  igab_memspace in_space, double* cp_out, igab_memspace out_space, void* stream_) {
  if (dim < 1 || dim > kMaxDim || dimworld < 1 || dimworld > 3 || !ncp_old || nops < 0 || !cp_in || !cp_out ||
      (nops > 0 && (!direction || !n_new || !maxlen || !first || !len || !coef))) {
    set_error("refine_operators: invalid argument");
    return IGAB_INVALID_ARGUMENT;
  }
  int ncp[kMaxDim] = {1, 1, 1};
  long long nin = 1;
  for (int d = 0; d < dim; ++d) {
    if (ncp_old[d] < 1) {
      set_error("refine_operators: empty net");
      return IGAB_INVALID_ARGUMENT;
    }
    ncp[d] = ncp_old[d];
    nin *= ncp[d];
  }
  long long nmax = nin, nout = nin;
  size_t ibytes = 0, cbytes = 0;
  for (int o = 0; o < nops; ++o) {
    const int d = direction[o];
    if (d < 0 || d >= dim || n_new[o] < 1 || maxlen[o] < 1 || !first[o] || !len[o] || !coef[o]) {
      set_error("refine_operators: invalid operator");
      return IGAB_INVALID_ARGUMENT;
    }
    for (int i = 0; i < n_new[o]; ++i)
      if (len[o][i] < 0 || len[o][i] > maxlen[o] || first[o][i] < 0 || first[o][i] + len[o][i] > ncp[d]) {
        set_error("refine_operators: operator row outside the net it is applied to");
        return IGAB_INVALID_ARGUMENT;
      }
    nout = nout / ncp[d] * n_new[o];
    ncp[d] = n_new[o];
    nmax = std::max(nmax, nout);
    ibytes = std::max(ibytes, sizeof(int) * (size_t)n_new[o]);
    cbytes = std::max(cbytes, sizeof(double) * (size_t)n_new[o] * maxlen[o]);
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t row = sizeof(double) * (dimworld + 1);
  const cudaMemcpyKind kin = in_space == IGAB_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  const cudaMemcpyKind kout = out_space == IGAB_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  double *buf[2] = {nullptr, nullptr}, *dCoef = nullptr;
  int *dFirst = nullptr, *dLen = nullptr;
  cudaError_t e = cudaMallocAsync(&buf[0], row * nmax, stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&buf[1], row * nmax, stream);
  if (e == cudaSuccess && nops > 0) e = cudaMallocAsync(&dFirst, ibytes, stream);
  if (e == cudaSuccess && nops > 0) e = cudaMallocAsync(&dLen, ibytes, stream);
  if (e == cudaSuccess && nops > 0) e = cudaMallocAsync(&dCoef, cbytes, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(buf[0], cp_in, row * nin, kin, stream);
  igab_status st = e == cudaSuccess ? IGAB_OK : cuda_fail(e, "refine_operators: buffers");
  for (int d = 0; d < dim; ++d) ncp[d] = ncp_old[d];
  int cur = 0;
  for (int o = 0; o < nops && !st; ++o) {
    // the operator arrays are pageable host memory and the staging buffers are reused: one operator in flight at a time
    e = cudaMemcpyAsync(dFirst, first[o], sizeof(int) * n_new[o], cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dLen, len[o], sizeof(int) * n_new[o], cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(dCoef, coef[o], sizeof(double) * n_new[o] * maxlen[o], cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) st = cuda_fail(e, "refine_operators: operator H2D");
    if (!st)
      st = refine_apply(dim, dimworld, ncp, direction[o], n_new[o], maxlen[o], dFirst, dLen, dCoef, buf[cur], buf[cur ^ 1],
                        stream);
    if (!st && (e = cudaStreamSynchronize(stream)) != cudaSuccess) st = cuda_fail(e, "refine_operators: apply");
    ncp[direction[o]] = n_new[o];
    cur ^= 1;
  }
  if (!st) {
    e = cudaMemcpyAsync(cp_out, buf[cur], row * nout, kout, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) st = cuda_fail(e, "refine_operators: output");
  }
  if (buf[0]) cudaFreeAsync(buf[0], stream);
  if (buf[1]) cudaFreeAsync(buf[1], stream);
  if (dFirst) cudaFreeAsync(dFirst, stream);
  if (dLen) cudaFreeAsync(dLen, stream);
  if (dCoef) cudaFreeAsync(dCoef, stream);
  return st;
}

igab_status igab_refine_knots(int dim, int dimworld, const int* degree, const int* nknots_old,
                              const double* const* knots_old, const int* nknots_new, const double* const* knots_new,
                              const double* cp_in, igab_memspace in_space, double* cp_out, igab_memspace out_space,
                              void* stream_) {
  if (dim < 1 || dim > kMaxDim || dimworld < 1 || dimworld > 3 || !degree || !nknots_old || !knots_old || !nknots_new ||
      !knots_new || !cp_in || !cp_out) {
    set_error("refine_knots: invalid argument");
    return IGAB_INVALID_ARGUMENT;
  }
  long long nin = 1, nout = 1;
  for (int d = 0; d < dim; ++d) {
    const int p = degree[d];
    if (p < 1 || p > kMaxDegree || nknots_old[d] < 2 * (p + 1) || nknots_new[d] < nknots_old[d] || !knots_old[d] || !knots_new[d]) {
      set_error("refine_knots: degree 1..8, open knot vectors, and the new knot vector must contain the old one");
      return IGAB_INVALID_ARGUMENT;
    }
    // the new knot vector must be a refinement: every old knot appears (with at least its multiplicity), sorted
    int j = 0;
    for (int i = 0; i < nknots_new[d]; ++i) {
      if (i > 0 && knots_new[d][i] < knots_new[d][i - 1]) {
        set_error("refine_knots: You passed a non-sorted vector of new knots.");  // the reference's assert, :444-445
        return IGAB_INVALID_ARGUMENT;
      }
      if (j < nknots_old[d] && knots_new[d][i] == knots_old[d][j]) ++j;
    }
    if (j != nknots_old[d]) {
      set_error("refine_knots: the new knot vector does not contain the old one");
      return IGAB_INVALID_ARGUMENT;
    }
    nin *= nknots_old[d] - p - 1;
    nout *= nknots_new[d] - p - 1;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t row = sizeof(double) * (dimworld + 1);
  double *dIn = nullptr, *dOut = nullptr, *dScr = nullptr;
  cudaError_t e = cudaSuccess;
  if (in_space == IGAB_HOST) {
    e = cudaMallocAsync(&dIn, row * nin, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dIn, cp_in, row * nin, cudaMemcpyHostToDevice, stream);
  }
  if (e == cudaSuccess && out_space == IGAB_HOST) e = cudaMallocAsync(&dOut, row * nout, stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&dScr, row * nout, stream);
  igab_status st = e == cudaSuccess ? IGAB_OK : cuda_fail(e, "refine_knots: buffers");
  if (!st)
    st = refine_knots_device(dim, dimworld, degree, nknots_old, knots_old, nknots_new, knots_new, dIn ? dIn : cp_in,
                             dOut ? dOut : cp_out, dScr, stream);
  if (!st && out_space == IGAB_HOST) {
    e = cudaMemcpyAsync(cp_out, dOut, row * nout, cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) st = cuda_fail(e, "refine_knots: D2H");
  }
  // the knot vectors are pageable host arrays read by the asynchronous copies above
  e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess && !st) st = cuda_fail(e, "refine_knots: sync");
  if (dIn) cudaFreeAsync(dIn, stream);
  if (dOut) cudaFreeAsync(dOut, stream);
  if (dScr) cudaFreeAsync(dScr, stream);
  return st;
}

// ---- faces -----------------------------------------------------------------------------------------------------------
igab_status igab_eval_faces(igab_patch* p, int64_t nfaces, const int32_t* elem, const int32_t* face, const int* nqf1,
                            const double* const* xi1d, double* x, double* normal, double* dS, igab_memspace space,
                            void* stream_) {
  if (!p || nfaces < 0 || !nqf1 || !xi1d || (nfaces > 0 && (!elem || !face))) {
    set_error("eval_faces: null argument");
    return IGAB_INVALID_ARGUMENT;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  if (igab_status est = enter_stream(p, stream)) return est;
  const int dim = p->dev.dim, dw = p->dev.dw;
  if (dim < 2) {
    set_error("eval_faces: needs a 2-D or 3-D patch");
    return IGAB_UNSUPPORTED;
  }
  // the face rule is staged like an element rule (unused last direction padded with one point)
  int nq1[kMaxDim] = {1, 1, 1};
  const double one = 0.0;
  const double* xs[kMaxDim] = {&one, &one, &one};
  for (int d = 0; d < dim - 1; ++d) {
    nq1[d] = nqf1[d];
    xs[d] = xi1d[d];
  }
  RuleDev rule;
  igab_status st = stage_rule(p, nq1, xs, nullptr, &rule, stream);
  if (st) return st;
  if (nfaces == 0) return IGAB_OK;
  const int nqf = nq1[0] * (dim == 3 ? nq1[1] : 1);
  const long long npts = nfaces * nqf;
  if (space == IGAB_DEVICE) return eval_faces_device(p, nfaces, elem, face, nq1, rule.xi, x, normal, dS, stream);
  for (long long k = 0; k < nfaces; ++k)
    if (elem[k] < 0 || elem[k] >= p->dev.nelem || face[k] < 0 || face[k] >= 2 * dim) {
      set_error("eval_faces: element or face index out of range");
      return IGAB_INVALID_ARGUMENT;
    }
  int *d_el = nullptr, *d_fc = nullptr;
  double *d_x = nullptr, *d_n = nullptr, *d_s = nullptr;
  cudaError_t e = cudaMalloc(&d_el, sizeof(int) * nfaces);
  if (e == cudaSuccess) e = cudaMalloc(&d_fc, sizeof(int) * nfaces);
  if (e == cudaSuccess && x) e = cudaMalloc(&d_x, sizeof(double) * npts * dw);
  if (e == cudaSuccess && normal) e = cudaMalloc(&d_n, sizeof(double) * npts * dw);
  if (e == cudaSuccess && dS) e = cudaMalloc(&d_s, sizeof(double) * npts);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_el, elem, sizeof(int) * nfaces, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_fc, face, sizeof(int) * nfaces, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) st = cuda_fail(e, "eval_faces staging");
  if (!st) st = eval_faces_device(p, nfaces, d_el, d_fc, nq1, rule.xi, d_x, d_n, d_s, stream);
  if (!st) {
    if (x) e = cudaMemcpyAsync(x, d_x, sizeof(double) * npts * dw, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && normal) e = cudaMemcpyAsync(normal, d_n, sizeof(double) * npts * dw, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && dS) e = cudaMemcpyAsync(dS, d_s, sizeof(double) * npts, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) st = cuda_fail(e, "eval_faces D2H");
  }
  cudaStreamSynchronize(stream);
  if (d_el) cudaFree(d_el);
  if (d_fc) cudaFree(d_fc);
  if (d_x) cudaFree(d_x);
  if (d_n) cudaFree(d_n);
  if (d_s) cudaFree(d_s);
  return st;
}

// ---- differential geometry of surfaces ---------------------------------------------------------------------------------
igab_status igab_surface_forms(int dim, int dimworld, int64_t npts, const double* Jt, const double* H, double* normal,
                               double* unit_normal, double* metric, double* second_ff, double* gauss_k,
                               igab_memspace space, void* stream_) {
  if (npts < 0 || (npts > 0 && !Jt)) {
    set_error("surface_forms: null argument");
    return IGAB_INVALID_ARGUMENT;
  }
  if (dim < 1 || dim > 3 || dimworld < dim || dimworld > 3) {
    set_error("surface_forms: unsupported (dim, dimworld)");
    return IGAB_UNSUPPORTED;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  if (space == IGAB_DEVICE)
    return surface_forms_device(dim, dimworld, npts, Jt, H, normal, unit_normal, metric, second_ff, gauss_k, stream);
  if (npts == 0) return IGAB_OK;
  // host arrays: one device block [Jt | H | normal | unit | metric | b | K]
  const size_t nh = (size_t)dim * (dim + 1) / 2;
  const size_t len[7] = {(size_t)dim * dimworld, H ? nh * dimworld : 0, normal ? 3u : 0u, unit_normal ? 3u : 0u,
                         metric ? (size_t)dim * dim : 0, second_ff ? 4u : 0u, gauss_k ? 1u : 0u};
  size_t off[8] = {0};
  for (int k = 0; k < 7; ++k) off[k + 1] = off[k] + len[k] * (size_t)npts;
  double* d = nullptr;
  IGAB_CUDA_TRY(cudaMalloc(&d, off[7] * sizeof(double)));
  auto part = [&](int k) { return len[k] ? d + off[k] : nullptr; };
  cudaError_t e = cudaMemcpyAsync(d + off[0], Jt, len[0] * npts * sizeof(double), cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess && H) e = cudaMemcpyAsync(d + off[1], H, len[1] * npts * sizeof(double), cudaMemcpyHostToDevice, stream);
  igab_status st = e == cudaSuccess ? IGAB_OK : cuda_fail(e, "surface_forms H2D");
  if (!st) st = surface_forms_device(dim, dimworld, npts, part(0), part(1), part(2), part(3), part(4), part(5), part(6), stream);
  if (!st) {
    double* host[7] = {nullptr, nullptr, normal, unit_normal, metric, second_ff, gauss_k};
    for (int k = 2; k < 7 && e == cudaSuccess; ++k)
      if (host[k]) e = cudaMemcpyAsync(host[k], d + off[k], len[k] * npts * sizeof(double), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) st = cuda_fail(e, "surface_forms D2H");
  }
  cudaStreamSynchronize(stream);
  cudaFree(d);
  return st;
}

// ---- inverse of the element map ---------------------------------------------------------------------------------------
igab_status igab_geometry_local(igab_patch* p, int64_t npts, const int32_t* elem, const double* xglobal, double* xi,
                                int32_t* flag, igab_memspace space, void* stream_) {
  if (!p || npts < 0 || (npts > 0 && (!elem || !xglobal || !xi))) {
    set_error("geometry_local: null argument");
    return IGAB_INVALID_ARGUMENT;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  if (igab_status est = enter_stream(p, stream)) return est;
  const int dim = p->dev.dim, dw = p->dev.dw;
  if (dim != dw && dim > 2) {
    set_error("geometry_local: closest-point projection exists for curves and surfaces only");
    return IGAB_UNSUPPORTED;
  }
  if (space == IGAB_DEVICE) return geometry_local_device(p->dev, npts, elem, xglobal, xi, flag, stream);
  if (npts == 0) return IGAB_OK;
  for (long long k = 0; k < npts; ++k)
    if (elem[k] < 0 || elem[k] >= p->dev.nelem) {
      set_error("geometry_local: element index out of range");
      return IGAB_INVALID_ARGUMENT;
    }
  int *d_el = nullptr, *d_fl = nullptr;
  double *d_x = nullptr, *d_xi = nullptr;
  cudaError_t e = cudaMalloc(&d_el, sizeof(int) * npts);
  if (e == cudaSuccess) e = cudaMalloc(&d_fl, sizeof(int) * npts);
  if (e == cudaSuccess) e = cudaMalloc(&d_x, sizeof(double) * npts * dw);
  if (e == cudaSuccess) e = cudaMalloc(&d_xi, sizeof(double) * npts * dim);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_el, elem, sizeof(int) * npts, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_x, xglobal, sizeof(double) * npts * dw, cudaMemcpyHostToDevice, stream);
  igab_status st = e == cudaSuccess ? IGAB_OK : cuda_fail(e, "geometry_local staging");
  if (!st) st = geometry_local_device(p->dev, npts, d_el, d_x, d_xi, d_fl, stream);
  if (!st) {
    e = cudaMemcpyAsync(xi, d_xi, sizeof(double) * npts * dim, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && flag) e = cudaMemcpyAsync(flag, d_fl, sizeof(int) * npts, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) st = cuda_fail(e, "geometry_local D2H");
  }
  cudaStreamSynchronize(stream);
  if (d_el) cudaFree(d_el);
  if (d_fl) cudaFree(d_fl);
  if (d_x) cudaFree(d_x);
  if (d_xi) cudaFree(d_xi);
  return st;
}

igab_status igab_rhs_assemble(igab_patch* p, int ncomp, int64_t elem_begin, int64_t elem_end, const double* fe, double* b,
                              int accumulate, igab_memspace space, void* stream) {
  if (!p || !fe || !b) {
    set_error("rhs_assemble: null argument");
    return IGAB_INVALID_ARGUMENT;
  }
  if (elem_begin < 0 || elem_end < elem_begin || elem_end > p->dev.nelem) {
    set_error("rhs_assemble: element range outside [0, n_elem]");
    return IGAB_INVALID_ARGUMENT;
  }
  if (igab_status est = enter_stream(p, (cudaStream_t)stream)) return est;
  return rhs_assemble(p, ncomp, elem_begin, elem_end, fe, b, accumulate, space, (cudaStream_t)stream);
}
igab_status igab_csr_dirichlet(igab_patch* p, int ncomp, const uint8_t* mask, const double* g, double* values, double* b,
                               igab_memspace space, void* stream) {
  if (!p || !mask || !values) {
    set_error("csr_dirichlet: null argument");
    return IGAB_INVALID_ARGUMENT;
  }
  if (igab_status est = enter_stream(p, (cudaStream_t)stream)) return est;
  return csr_dirichlet(p, ncomp, mask, g, values, b, space, (cudaStream_t)stream);
}

igab_status igab_assemble_points(igab_patch* p, int kind, const double* D, double fconst, int64_t nlist,
                                 const int32_t* elem, const int64_t* offset, const double* xi, const double* w,
                                 double* Ke, double* fe, igab_memspace space, void* stream_) {
  if (!p || nlist < 0 || !Ke || (nlist > 0 && (!elem || !offset))) {
    set_error("assemble_points: null argument");
    return IGAB_INVALID_ARGUMENT;
  }
  if (kind != IGAB_ASM_POISSON && kind != IGAB_ASM_MASS && kind != IGAB_ASM_ELASTICITY) {
    set_error("assemble_points: unknown kind");
    return IGAB_INVALID_ARGUMENT;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  if (igab_status est = enter_stream(p, stream)) return est;
  if (nlist == 0) return IGAB_OK;
  static_assert(sizeof(long long) == sizeof(int64_t), "offset type");
  if (space == IGAB_DEVICE)
    return launch_assemble2d_points(p->dev, kind, D, fconst, nlist, elem, (const long long*)offset, xi, w, Ke, fe, stream);
  const long long npts = offset[nlist];
  for (long long k = 0; k < nlist; ++k)
    if (elem[k] < 0 || elem[k] >= p->dev.nelem || offset[k + 1] < offset[k] || offset[0] != 0) {
      set_error("assemble_points: element index out of range or offsets not ascending from 0");
      return IGAB_INVALID_ARGUMENT;
    }
  if (npts > 0 && (!xi || !w)) {
    set_error("assemble_points: null points / weights");
    return IGAB_INVALID_ARGUMENT;
  }
  const int ncomp = kind == IGAB_ASM_ELASTICITY ? p->dev.dim : 1;
  const long long n = (long long)p->dev.nb * ncomp;
  int* d_el = nullptr;
  long long* d_off = nullptr;
  double *d_xi = nullptr, *d_w = nullptr, *d_K = nullptr, *d_f = nullptr;
  cudaError_t e = cudaMalloc(&d_el, sizeof(int) * nlist);
  if (e == cudaSuccess) e = cudaMalloc(&d_off, sizeof(long long) * (nlist + 1));
  if (e == cudaSuccess) e = cudaMalloc(&d_xi, sizeof(double) * std::max(1LL, npts * p->dev.dim));
  if (e == cudaSuccess) e = cudaMalloc(&d_w, sizeof(double) * std::max(1LL, npts));
  if (e == cudaSuccess) e = cudaMalloc(&d_K, sizeof(double) * nlist * n * n);
  if (e == cudaSuccess && fe) e = cudaMalloc(&d_f, sizeof(double) * nlist * n);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_el, elem, sizeof(int) * nlist, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_off, offset, sizeof(long long) * (nlist + 1), cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess && npts) e = cudaMemcpyAsync(d_xi, xi, sizeof(double) * npts * p->dev.dim, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess && npts) e = cudaMemcpyAsync(d_w, w, sizeof(double) * npts, cudaMemcpyHostToDevice, stream);
  igab_status st = e == cudaSuccess ? IGAB_OK : cuda_fail(e, "assemble_points staging");
  if (!st) st = launch_assemble2d_points(p->dev, kind, D, fconst, nlist, d_el, d_off, d_xi, d_w, d_K, d_f, stream);
  if (!st) {
    e = cudaMemcpyAsync(Ke, d_K, sizeof(double) * nlist * n * n, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && fe) e = cudaMemcpyAsync(fe, d_f, sizeof(double) * nlist * n, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) st = cuda_fail(e, "assemble_points D2H");
  }
  cudaStreamSynchronize(stream);
  cudaFree(d_el); cudaFree(d_off); cudaFree(d_xi); cudaFree(d_w); cudaFree(d_K);
  if (d_f) cudaFree(d_f);
  return st;
}

// ---- local keys (host integers) ------------------------------------------------------------------------------------------
igab_status igab_local_keys(int dim, const int* degree, uint32_t* sub_entity, uint32_t* codim, uint32_t* index) {
  if (dim < 1 || dim > kMaxDim || !degree || !sub_entity || !codim || !index) {
    set_error("local_keys: null argument or dim outside 1..3");
    return IGAB_INVALID_ARGUMENT;
  }
  unsigned sizes[kMaxDim] = {1, 1, 1};
  unsigned nb = 1;
  for (int d = 0; d < dim; ++d) {
    if (degree[d] < 0 || degree[d] > kMaxDegree) {
      set_error("local_keys: degree outside 0..8");
      return IGAB_INVALID_ARGUMENT;
    }
    sizes[d] = (unsigned)degree[d] + 1;
    nb *= sizes[d];
  }
  for (unsigned i = 0; i < nb; ++i) {
    unsigned m[kMaxDim], r = i;
    for (int d = 0; d < dim; ++d) {
      m[d] = r % sizes[d];
      r /= sizes[d];
    }
    unsigned cd = 0, ix = 0;
    for (int d = 0; d < dim; ++d)
      if (m[d] == 0 || m[d] == sizes[d] - 1) ++cd;
    for (int d = dim - 1; d >= 0; --d)
      if (m[d] > 0 && m[d] < sizes[d] - 1) ix = (sizes[d] - 1) * ix + (m[d] - 1);
    codim[i] = cd;
    index[i] = ix;
    sub_entity[i] = 0;
  }
  if (dim == 1 && sizes[0] > 1) {
    sub_entity[nb - 1] = 1;  // 0----0----1
  } else if (dim == 2 && sizes[0] > 1 && sizes[1] > 1) {
    // vertices 0 1 / 2 3, edges: left 0, right 1, lower 2, upper 3, face 0
    for (unsigned j = 0; j < sizes[1]; ++j)
      for (unsigned i = 0; i < sizes[0]; ++i) {
        const bool lo0 = i == 0, hi0 = i == sizes[0] - 1, lo1 = j == 0, hi1 = j == sizes[1] - 1;
        unsigned s = 0;
        if (lo1) s = lo0 ? 0 : (hi0 ? 1 : 2);
        else if (hi1) s = lo0 ? 2 : 3;  // corner 2, inner dofs of the upper edge 3, corner 3
        else s = hi0 ? 1 : 0;           // left edge 0 / face 0 / right edge 1
        sub_entity[j * sizes[0] + i] = s;
      }
  }
  return IGAB_OK;
}

// ---- global CSR -----------------------------------------------------------------------------------------------------
igab_status igab_csr_pattern(igab_patch* p, int ncomp, int64_t* nrows, int64_t* nnz, int64_t* rowptr, int32_t* colidx,
                             igab_memspace space, void* stream) {
  if (!p) return IGAB_INVALID_ARGUMENT;
  if (igab_status est = enter_stream(p, (cudaStream_t)stream)) return est;
  return csr_pattern(p, ncomp, nrows, nnz, rowptr, colidx, space, (cudaStream_t)stream);
}
igab_status igab_csr_assemble(igab_patch* p, int ncomp, int64_t elem_begin, int64_t elem_end, const double* Ke,
                              const int64_t* rowptr, double* values, int accumulate, igab_memspace space,
                              void* stream) {
  if (!p || !Ke || !values) {
    set_error("csr_assemble: null argument");
    return IGAB_INVALID_ARGUMENT;
  }
  if (elem_begin < 0 || elem_end < elem_begin || elem_end > p->dev.nelem) {
    set_error("csr_assemble: element range outside [0, n_elem]");
    return IGAB_INVALID_ARGUMENT;
  }
  if (igab_status est = enter_stream(p, (cudaStream_t)stream)) return est;
  return csr_assemble(p, ncomp, elem_begin, elem_end, Ke, rowptr, values, accumulate, space, (cudaStream_t)stream);
}

// ---- the whole globalAssembler in one call --------------------------------------------------------------------------
namespace {
// element matrices of trimmed elements (their own rules) replace the tensor-rule ones in a chunk's Ke / fe
__global__ void put_trimmed_kernel(long long n2, long long m, const int32_t* __restrict__ elem, long long eb,
                                   const double* __restrict__ src, double* __restrict__ dst) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n2 * m) return;
  const long long k = idx / n2, r = idx - k * n2;
  dst[(elem[k] - eb) * n2 + r] = src[idx];
}
}  // namespace

igab_status igab_global_assemble(igab_patch* p, int kind, const double* D, double fconst, int64_t elem_begin,
                                 int64_t elem_end, const int* nq1, const double* const* xi1d,
                                 const double* const* w1d, int64_t nlist,
                                 const int32_t* elem, const int64_t* offset, const double* xi, const double* w,
                                 double* values, double* b, igab_memspace space, void* stream_) {
  if (!p || !nq1 || !xi1d || !w1d || !values || nlist < 0 || (nlist > 0 && (!elem || !offset || !xi || !w))) {
    set_error("global_assemble: null argument");
    return IGAB_INVALID_ARGUMENT;
  }
  if (kind < 0 || kind > 2) {
    set_error("global_assemble: unknown kind");
    return IGAB_INVALID_ARGUMENT;
  }
  if (elem_begin < 0 || elem_end < elem_begin || elem_end > p->dev.nelem) {
    set_error("global_assemble: element range outside [0, n_elem]");
    return IGAB_INVALID_ARGUMENT;
  }
  for (int64_t k = 0; k < nlist; ++k)
    if (elem[k] < 0 || elem[k] >= p->dev.nelem || (k > 0 && elem[k] <= elem[k - 1]) || offset[k + 1] < offset[k]) {
      set_error("global_assemble: trimmed-element list must be strictly ascending with valid offsets");
      return IGAB_INVALID_ARGUMENT;
    }
  cudaStream_t stream = (cudaStream_t)stream_;
  if (igab_status est = enter_stream(p, stream)) return est;
  const int ncomp = kind == IGAB_ASM_ELASTICITY ? p->dev.dim : 1;
  const long long n = (long long)p->dev.nb * ncomp, n2 = n * n, nrange = elem_end - elem_begin;
  int64_t nrows = 0, nnz = 0;
  igab_status st = csr_pattern(p, ncomp, &nrows, &nnz, nullptr, nullptr, IGAB_HOST, stream);
  if (st) return st;
  // Element matrices per chunk.  Every chunk re-reads and re-writes the CSR rows it shares with its neighbours (p layers
  // of DOFs on either side), so large chunks pay: measured on the C3 patch 42.7 ms with 1 GB chunks, 31.1 ms with 8 GB
  // (tools/ga_chunk_big.py).  Up to 8 GB, at most a quarter of the free device memory (this handle's scratch counted as
  // free), whole layers of elements; host outputs keep at least four chunks, so that the D2H copy of finished rows
  // overlaps the assembly of later chunks.
  const long long gb1 = std::max(1LL, (1024LL << 20) / (8 * n2));
  long long chunk = 8 * gb1;
  {
    size_t freeB = 0, totalB = 0;
    if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess) {
      const long long avail = (long long)freeB + (long long)p->ga_cap * 8;
      chunk = std::min(chunk, std::max(gb1 / 4 + 1, avail / 4 / (8 * n2)));
    } else {
      cudaGetLastError();
      chunk = gb1;
    }
    if (space == IGAB_HOST && !p->d_trim) chunk = std::min(chunk, std::max(gb1, (nrange + 3) / 4));
    long long layer = 1;
    for (int d = 0; d + 1 < p->dev.dim; ++d) layer *= p->dev.nel[d];
    if (chunk > layer) chunk -= chunk % layer;
  }
  if (const char* env = getenv("IGAB_GA_CHUNK_ELEMS")) chunk = std::max(1L, atol(env));  // test knob
  chunk = std::min(chunk, std::max(1LL, nrange));
  double *dK = nullptr, *dF = nullptr, *dV = values, *dB = b, *dKt = nullptr, *dFt = nullptr, *dXi = nullptr, *dW = nullptr;
  int32_t* dElem = nullptr;
  int64_t* dOff = nullptr;
  long long ktCap = 0;
  auto cleanup = [&]() {
    cudaStreamSynchronize(stream);
    if (dKt) cudaFree(dKt);
    if (dFt) cudaFree(dFt);
    if (dXi) cudaFree(dXi);
    if (dW) cudaFree(dW);
    if (dElem) cudaFree(dElem);
    if (dOff) cudaFree(dOff);
  };
  // one scratch block per handle, grown on demand and kept: [element matrices | element vectors | values | load vector]
  const size_t nK = (size_t)chunk * n2, nF = b ? (size_t)chunk * n : 0;
  const size_t nV = space == IGAB_HOST ? (size_t)std::max<long long>(1, nnz) : 0;
  const size_t nB = (space == IGAB_HOST && b) ? (size_t)std::max<long long>(1, nrows) : 0;
  const size_t need = nK + nF + nV + nB + 8;
  cudaError_t e = cudaSuccess;
  if (p->ga_cap < need) {
    cudaStreamSynchronize(stream);
    if (p->ga_scratch) cudaFree(p->ga_scratch);
    p->ga_scratch = nullptr;
    p->ga_cap = 0;
    e = cudaMalloc(&p->ga_scratch, sizeof(double) * need);
    if (e == cudaSuccess) p->ga_cap = need;
  }
  if (e == cudaSuccess) {
    dK = p->ga_scratch;
    dF = b ? dK + nK : nullptr;
    if (space == IGAB_HOST) {
      dV = dK + nK + nF;
      dB = b ? dV + nV : nullptr;
    }
  }
  if (e == cudaSuccess && nlist > 0) {
    const long long np = offset[nlist];
    e = cudaMalloc(&dElem, sizeof(int32_t) * nlist);
    if (e == cudaSuccess) e = cudaMalloc(&dOff, sizeof(int64_t) * (nlist + 1));
    if (e == cudaSuccess) e = cudaMalloc(&dXi, sizeof(double) * std::max(1LL, np) * p->dev.dim);
    if (e == cudaSuccess) e = cudaMalloc(&dW, sizeof(double) * std::max(1LL, np));
    if (e == cudaSuccess) e = cudaMemcpyAsync(dElem, elem, sizeof(int32_t) * nlist, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dOff, offset, sizeof(int64_t) * (nlist + 1), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dXi, xi, sizeof(double) * np * p->dev.dim, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dW, w, sizeof(double) * np, cudaMemcpyHostToDevice, stream);
  }
  if (e != cudaSuccess) {
    st = cuda_fail(e, "global_assemble: device buffers");
    cleanup();
    return st;
  }
  // overlap of the value D2H with the assembly of later chunks (host outputs, untrimmed numbering, non-empty range)
  bool overlap = space == IGAB_HOST && !p->d_trim && nrange > 0;
  cudaEvent_t evt = nullptr;
  const int dl = p->dev.dim - 1;
  long long perLayer = 1, dofStride = 1, rowsOut = 0;
  for (int d = 0; d < dl; ++d) {
    perLayer *= p->dev.nel[d];
    dofStride *= p->dev.ncp[d];
  }
  if (overlap) {
    if (!p->stage_stream) e = cudaStreamCreateWithFlags(&p->stage_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&evt, cudaEventDisableTiming);
    if (e != cudaSuccess) {
      overlap = false;
      evt = nullptr;
      e = cudaSuccess;
      cudaGetLastError();
    }
  }
  long long k0 = 0;  // first trimmed list entry not yet consumed
  while (k0 < nlist && elem[k0] < elem_begin) ++k0;
  if (nrange == 0) {  // nothing to add: the caller still gets a defined (zero) result
    e = cudaMemsetAsync(dV, 0, sizeof(double) * nnz, stream);
    if (e == cudaSuccess && b) e = cudaMemsetAsync(dB, 0, sizeof(double) * nrows, stream);
    if (e != cudaSuccess) st = cuda_fail(e, "global_assemble: memset");
  }
  for (long long eb = elem_begin; eb < elem_end && !st; eb += chunk) {
    const long long ee = std::min((long long)elem_end, eb + chunk);
    st = igab_assemble(p, kind, D, fconst, eb, ee, nq1, xi1d, w1d, dK, dF, IGAB_DEVICE, stream);
    long long k1 = k0;
    while (k1 < nlist && elem[k1] < ee) ++k1;
    if (!st && k1 > k0) {
      const long long m = k1 - k0;
      if (m > ktCap) {
        cudaStreamSynchronize(stream);
        if (dKt) cudaFree(dKt);
        if (dFt) cudaFree(dFt);
        dKt = dFt = nullptr;
        e = cudaMalloc(&dKt, sizeof(double) * m * n2);
        if (e == cudaSuccess && b) e = cudaMalloc(&dFt, sizeof(double) * m * n);
        if (e != cudaSuccess) st = cuda_fail(e, "global_assemble: trimmed staging");
        ktCap = m;
      }
      // offsets are absolute indices into xi / w, so a sub-list is the same arrays with shifted elem / offset pointers
      if (!st) st = igab_assemble_points(p, kind, D, fconst, m, dElem + k0, dOff + k0, dXi, dW, dKt, dFt, IGAB_DEVICE, stream);
      if (!st) {
        put_trimmed_kernel<<<(unsigned)((m * n2 + 255) / 256), 256, 0, stream>>>(n2, m, dElem + k0, eb, dKt, dK);
        if (b) put_trimmed_kernel<<<(unsigned)((m * n + 255) / 256), 256, 0, stream>>>(n, m, dElem + k0, eb, dFt, dF);
        count_launch(b ? 2 : 1);
        if ((e = cudaGetLastError()) != cudaSuccess) st = cuda_fail(e, "global_assemble: put_trimmed");
      }
    }
    k0 = k1;
    if (!st) st = csr_assemble(p, ncomp, eb, ee, dK, nullptr, dV, eb > elem_begin, IGAB_DEVICE, stream);
    if (!st && b) st = rhs_assemble(p, ncomp, eb, ee, dF, dB, eb > elem_begin, IGAB_DEVICE, stream);
    // host outputs: the rows below the next chunk's first row are final -- their values leave on the copy stream while
    // the next chunk is assembled (untrimmed numbering: rows of a k-slab of elements are one contiguous range)
    if (!st && overlap) {
      long long done = nrows;
      if (ee < elem_end) done = (long long)(p->hspan[dl][(int)(ee / perLayer)] - p->dev.degree[dl]) * dofStride * ncomp;
      if (done > rowsOut) {
        const long long z0 = p->csr_rowptr_host[rowsOut], z1 = p->csr_rowptr_host[done];
        e = cudaEventRecord(evt, stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(p->stage_stream, evt, 0);
        if (e == cudaSuccess && z1 > z0)
          e = cudaMemcpyAsync(values + z0, dV + z0, sizeof(double) * (z1 - z0), cudaMemcpyDeviceToHost, p->stage_stream);
        if (e != cudaSuccess) st = cuda_fail(e, "global_assemble: overlapped D2H");
        rowsOut = done;
      }
    }
  }
  if (!st && space == IGAB_HOST) {
    if (overlap) {
      e = cudaStreamSynchronize(p->stage_stream);
    } else {
      e = cudaMemcpyAsync(values, dV, sizeof(double) * nnz, cudaMemcpyDeviceToHost, stream);
    }
    if (e == cudaSuccess && b) e = cudaMemcpyAsync(b, dB, sizeof(double) * nrows, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) st = cuda_fail(e, "global_assemble: D2H");
  }
  if (evt) {
    cudaStreamSynchronize(p->stage_stream);
    cudaEventDestroy(evt);
  }
  cleanup();
  return st;
}

// ---- load vectors with a per-point source ----------------------------------------------------------------------------
namespace {
// fe[k][i*ncomp + c] = sum_q w_q detJ_q R_i(q) f_q[c]  (poisson.py:79, localB[i] += weight * integrationElement * phi[i] *
// f(posGlobal), with f already evaluated at the points).  One warp per element, lanes over the local functions, points in
// their rule order (the reference's summation order).  Tensor rules: the nq points of element k are k*nq .. with
// w = prod w1d; lists: off[k] - off[0] .. with their own weights.
struct LoadVecArgs {
  int nb, ncomp, tensor;
  long long nq;
  int nq1[kMaxDim];
  const double* w1d[kMaxDim];
  const long long* off;  // list mode: absolute offsets, off[0] is the first point of the chunk
  const double* wlist;   // list mode: weights of the chunk's points
  const double *N, *detJ, *f;
  double* fe;
};
__global__ void __launch_bounds__(128) loadvec_kernel(LoadVecArgs A, long long nel) {
  const long long k = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= nel) return;
  long long p0, p1;
  if (A.tensor) {
    p0 = k * A.nq;
    p1 = p0 + A.nq;
  } else {
    p0 = A.off[k] - A.off[0];
    p1 = A.off[k + 1] - A.off[0];
  }
  for (int i = lane; i < A.nb; i += 32) {
    double acc[3] = {0.0, 0.0, 0.0};
    for (long long pt = p0; pt < p1; ++pt) {
      double wq;
      if (A.tensor) {
        long long q = pt - p0;
        wq = A.w1d[0][q % A.nq1[0]];
        q /= A.nq1[0];
        if (A.w1d[1]) wq *= A.w1d[1][q % A.nq1[1]];
        q /= A.nq1[1];
        if (A.w1d[2]) wq *= A.w1d[2][q];
      } else {
        wq = A.wlist[pt];
      }
      const double s = wq * A.detJ[pt] * A.N[pt * A.nb + i];
      for (int c = 0; c < A.ncomp; ++c) acc[c] += s * A.f[pt * A.ncomp + c];
    }
    for (int c = 0; c < A.ncomp; ++c) A.fe[(k * A.nb + i) * A.ncomp + c] = acc[c];
  }
}
}  // namespace

igab_status igab_load_vector(igab_patch* p, int64_t elem_begin, int64_t elem_end, const int* nq1,
                             const double* const* xi1d, const double* const* w1d, const double* f, int ncomp, double* fe,
                             igab_memspace space, void* stream_) {
  if (!p || !nq1 || !xi1d || !w1d || !f || !fe || ncomp < 1 || ncomp > 3) {
    set_error("load_vector: null argument or ncomp outside 1..3");
    return IGAB_INVALID_ARGUMENT;
  }
  if (elem_begin < 0 || elem_end < elem_begin || elem_end > p->dev.nelem) {
    set_error("load_vector: element range outside [0, n_elem]");
    return IGAB_INVALID_ARGUMENT;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  if (igab_status est = enter_stream(p, stream)) return est;
  RuleDev rule;
  igab_status st = stage_rule(p, nq1, xi1d, w1d, &rule, stream);
  if (st) return st;
  const long long nel = elem_end - elem_begin;
  if (nel == 0) return IGAB_OK;
  const int nb = p->dev.nb, dim = p->dev.dim;
  long long nq = 1;
  for (int d = 0; d < dim; ++d) nq *= nq1[d];
  const long long elChunk = std::max(1LL, std::min(nel, (256LL << 20) / (8 * nq * nb)));
  // scratch: N, detJ of a chunk (+ staged f / fe for host arrays)
  const size_t nN = (size_t)elChunk * nq * nb, nJ = (size_t)elChunk * nq, nF = (size_t)elChunk * nq * ncomp,
               nE = (size_t)elChunk * nb * ncomp;
  double* scratch = nullptr;
  IGAB_CUDA_TRY(cudaMallocAsync(&scratch, sizeof(double) * (nN + nJ + (space == IGAB_HOST ? nF + nE : 0)), stream));
  double *dN = scratch, *dJ = dN + nN, *dF = dJ + nJ, *dE = dF + nF;
  cudaError_t e = cudaSuccess;
  for (long long c0 = 0; c0 < nel && !st; c0 += elChunk) {
    const long long m = std::min(elChunk, nel - c0);
    igab_eval_out o{};
    o.N = dN;
    o.detJ = dJ;
    st = run_eval_tensor_device(p, elem_begin + c0, elem_begin + c0 + m, 1, nq1, xi1d, rule, o, stream);
    if (st) break;
    LoadVecArgs A{};
    A.nb = nb;
    A.ncomp = ncomp;
    A.tensor = 1;
    A.nq = nq;
    for (int d = 0; d < kMaxDim; ++d) {
      A.nq1[d] = d < dim ? nq1[d] : 1;
      A.w1d[d] = d < dim ? rule.w[d] : nullptr;
    }
    A.N = dN;
    A.detJ = dJ;
    if (space == IGAB_HOST) {
      e = cudaMemcpyAsync(dF, f + c0 * nq * ncomp, sizeof(double) * m * nq * ncomp, cudaMemcpyHostToDevice, stream);
      if (e != cudaSuccess) { st = cuda_fail(e, "load_vector f H2D"); break; }
      A.f = dF;
      A.fe = dE;
    } else {
      A.f = f + c0 * nq * ncomp;
      A.fe = fe + c0 * nb * ncomp;
    }
    loadvec_kernel<<<(unsigned)((m + 3) / 4), 128, 0, stream>>>(A, m);
    count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) { st = cuda_fail(e, "load_vector kernel"); break; }
    if (space == IGAB_HOST) {
      e = cudaMemcpyAsync(fe + c0 * nb * ncomp, dE, sizeof(double) * m * nb * ncomp, cudaMemcpyDeviceToHost, stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(stream);  // dF / dE are reused by the next chunk
      if (e != cudaSuccess) st = cuda_fail(e, "load_vector D2H");
    }
  }
  cudaFreeAsync(scratch, stream);
  return st;
}

igab_status igab_load_vector_points(igab_patch* p, int64_t nlist, const int32_t* elem, const int64_t* offset,
                                    const double* xi, const double* w, const double* f, int ncomp, double* fe,
                                    igab_memspace space, void* stream_) {
  if (!p || nlist < 0 || ncomp < 1 || ncomp > 3 || (nlist > 0 && (!elem || !offset || !fe))) {
    set_error("load_vector_points: null argument or ncomp outside 1..3");
    return IGAB_INVALID_ARGUMENT;
  }
  if (nlist == 0) return IGAB_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (igab_status est = enter_stream(p, stream)) return est;
  const int nb = p->dev.nb, dim = p->dev.dim;
  int32_t* dEl = nullptr;   // element index of every POINT (what the point-list evaluation takes)
  long long* dOff = nullptr;
  double* scratch = nullptr;
  igab_status st = IGAB_OK;
  cudaError_t e = cudaSuccess;
  // the list's offsets are needed on the host to expand elem per point and to size the scratch
  std::vector<long long> hoff(nlist + 1);
  std::vector<int32_t> hel(nlist);
  if (space == IGAB_HOST) {
    std::copy(offset, offset + nlist + 1, hoff.begin());
    std::copy(elem, elem + nlist, hel.begin());
  } else {
    e = cudaMemcpyAsync(hoff.data(), offset, sizeof(long long) * (nlist + 1), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hel.data(), elem, sizeof(int32_t) * nlist, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return cuda_fail(e, "load_vector_points: list D2H");
  }
  const long long p0 = hoff[0], np = hoff[nlist] - p0;
  for (long long k = 0; k < nlist; ++k)
    if (hel[k] < 0 || hel[k] >= p->dev.nelem || hoff[k + 1] < hoff[k]) {
      set_error("load_vector_points: element index out of range or offsets not ascending");
      return IGAB_INVALID_ARGUMENT;
    }
  if (np > 0 && (!xi || !w || !f)) {
    set_error("load_vector_points: null points / weights / source");
    return IGAB_INVALID_ARGUMENT;
  }
  std::vector<int32_t> pel((size_t)std::max(1LL, np));
  for (long long k = 0; k < nlist; ++k)
    for (long long q = hoff[k]; q < hoff[k + 1]; ++q) pel[q - p0] = hel[k];
  const size_t nN = (size_t)np * nb, nJ = (size_t)np;
  const size_t nIn = space == IGAB_HOST ? (size_t)np * (dim + 1 + ncomp) + (size_t)nlist * nb * ncomp : 0;
  e = cudaMallocAsync(&scratch, sizeof(double) * std::max<size_t>(1, nN + nJ + nIn), stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&dEl, sizeof(int32_t) * std::max(1LL, np), stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&dOff, sizeof(long long) * (nlist + 1), stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dEl, pel.data(), sizeof(int32_t) * np, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dOff, hoff.data(), sizeof(long long) * (nlist + 1), cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) st = cuda_fail(e, "load_vector_points staging");
  double *dN = scratch, *dJ = dN + nN, *dXi = dJ + nJ, *dW = dXi + (size_t)np * dim, *dF = dW + np,
         *dE = dF + (size_t)np * ncomp;
  const double *xiD = xi ? xi + p0 * dim : nullptr, *wD = w ? w + p0 : nullptr, *fD = f ? f + p0 * ncomp : nullptr;
  if (!st && space == IGAB_HOST && np > 0) {
    e = cudaMemcpyAsync(dXi, xiD, sizeof(double) * np * dim, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dW, wD, sizeof(double) * np, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dF, fD, sizeof(double) * np * ncomp, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) st = cuda_fail(e, "load_vector_points H2D");
    xiD = dXi;
    wD = dW;
    fD = dF;
  }
  if (!st && np > 0) {
    PointSource src{};
    src.tensor = 0;
    src.elem = dEl;
    src.xi = xiD;
    EvalOutDev o{};
    o.N = dN;
    o.detJ = dJ;
    st = run_eval_points_device(p, src, np, 1, o, stream);
  }
  if (!st) {
    LoadVecArgs A{};
    A.nb = nb;
    A.ncomp = ncomp;
    A.tensor = 0;
    A.off = dOff;
    A.wlist = wD;
    A.N = dN;
    A.detJ = dJ;
    A.f = fD;
    A.fe = space == IGAB_HOST ? dE : fe;
    loadvec_kernel<<<(unsigned)((nlist + 3) / 4), 128, 0, stream>>>(A, nlist);
    count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) st = cuda_fail(e, "load_vector_points kernel");
  }
  if (!st && space == IGAB_HOST) {
    e = cudaMemcpyAsync(fe, dE, sizeof(double) * nlist * nb * ncomp, cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) st = cuda_fail(e, "load_vector_points D2H");
  }
  // pel / hoff are pageable host arrays read by the async copies above: drain before they go out of scope
  e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess && !st) st = cuda_fail(e, "load_vector_points sync");
  if (scratch) cudaFreeAsync(scratch, stream);
  if (dEl) cudaFreeAsync(dEl, stream);
  if (dOff) cudaFreeAsync(dOff, stream);
  return st;
}

igab_status igab_global_load_vector(igab_patch* p, int ncomp, int64_t elem_begin, int64_t elem_end, const int* nq1,
                                    const double* const* xi1d, const double* const* w1d, const double* f, int64_t nlist,
                                    const int32_t* elem, const int64_t* offset, const double* xi, const double* w,
                                    const double* f_list, double* b, igab_memspace space, void* stream_) {
  if (!p || !nq1 || !xi1d || !w1d || !f || !b || nlist < 0 || (nlist > 0 && (!elem || !offset || !xi || !w || !f_list))) {
    set_error("global_load_vector: null argument");
    return IGAB_INVALID_ARGUMENT;
  }
  if (elem_begin < 0 || elem_end < elem_begin || elem_end > p->dev.nelem) {
    set_error("global_load_vector: element range outside [0, n_elem]");
    return IGAB_INVALID_ARGUMENT;
  }
  for (int64_t k = 0; k < nlist; ++k)
    if (elem[k] < 0 || elem[k] >= p->dev.nelem || (k > 0 && elem[k] <= elem[k - 1]) || offset[k + 1] < offset[k]) {
      set_error("global_load_vector: trimmed-element list must be strictly ascending with valid offsets");
      return IGAB_INVALID_ARGUMENT;
    }
  cudaStream_t stream = (cudaStream_t)stream_;
  if (igab_status est = enter_stream(p, stream)) return est;
  const int nb = p->dev.nb;
  long long nq = 1;
  for (int d = 0; d < p->dev.dim; ++d) nq *= nq1[d];
  int64_t nrows = 0, nnz = 0;
  igab_status st = csr_pattern(p, ncomp, &nrows, &nnz, nullptr, nullptr, IGAB_HOST, stream);
  if (st) return st;
  const long long nrange = elem_end - elem_begin, n = (long long)nb * ncomp;
  const long long chunk = std::max(1LL, std::min(std::max(1LL, nrange), (64LL << 20) / (8 * n)));
  long long k0 = 0;
  while (k0 < nlist && elem[k0] < elem_begin) ++k0;
  const long long kFirst = k0;  // dFt / dElem hold the list entries kFirst .. kEnd of the range
  long long kEnd = k0;
  while (kEnd < nlist && elem[kEnd] < elem_end) ++kEnd;
  double *dFe = nullptr, *dB = b, *dFt = nullptr, *dFsrc = nullptr;
  int32_t* dElem = nullptr;
  cudaError_t e = cudaMallocAsync(&dFe, sizeof(double) * chunk * n, stream);
  if (e == cudaSuccess && space == IGAB_HOST) e = cudaMallocAsync(&dB, sizeof(double) * std::max<int64_t>(1, nrows), stream);
  if (e == cudaSuccess && space == IGAB_HOST) e = cudaMallocAsync(&dFsrc, sizeof(double) * chunk * nq * ncomp, stream);
  if (e == cudaSuccess && kEnd > k0) {
    // element vectors of all trimmed elements of the range from their own rules, once
    e = cudaMallocAsync(&dFt, sizeof(double) * (kEnd - k0) * n, stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&dElem, sizeof(int32_t) * (kEnd - k0), stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dElem, elem + k0, sizeof(int32_t) * (kEnd - k0), cudaMemcpyHostToDevice, stream);
  }
  if (e != cudaSuccess) st = cuda_fail(e, "global_load_vector: device buffers");
  if (!st && kEnd > k0) {
    if (space == IGAB_HOST) {
      std::vector<double> host((size_t)(kEnd - k0) * n);  // host lists in, device element vectors out: stage the result
      st = igab_load_vector_points(p, kEnd - k0, elem + k0, offset + k0, xi, w, f_list, ncomp, host.data(), IGAB_HOST, stream);
      if (!st) {
        e = cudaMemcpyAsync(dFt, host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) st = cuda_fail(e, "global_load_vector: trimmed vectors H2D");
      }
    } else {
      // device space: elem / offset are HOST arrays by contract (as in igab_global_assemble), xi / w / f_list device
      long long* dOff = nullptr;
      e = cudaMallocAsync(&dOff, sizeof(long long) * (kEnd - k0 + 1), stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(dOff, offset + k0, sizeof(long long) * (kEnd - k0 + 1), cudaMemcpyHostToDevice, stream);
      if (e != cudaSuccess) st = cuda_fail(e, "global_load_vector: offsets");
      if (!st) st = igab_load_vector_points(p, kEnd - k0, dElem, (const int64_t*)dOff, xi, w, f_list, ncomp, dFt, IGAB_DEVICE, stream);
      if (dOff) cudaFreeAsync(dOff, stream);
    }
  }
  if (!st && nrange == 0) {
    e = cudaMemsetAsync(dB, 0, sizeof(double) * nrows, stream);
    if (e != cudaSuccess) st = cuda_fail(e, "global_load_vector: memset");
  }
  for (long long eb = elem_begin; eb < elem_end && !st; eb += chunk) {
    const long long ee = std::min((long long)elem_end, eb + chunk);
    const double* fsrc = f + (eb - elem_begin) * nq * ncomp;
    if (space == IGAB_HOST) {
      e = cudaMemcpyAsync(dFsrc, fsrc, sizeof(double) * (ee - eb) * nq * ncomp, cudaMemcpyHostToDevice, stream);
      if (e != cudaSuccess) { st = cuda_fail(e, "global_load_vector: f H2D"); break; }
      fsrc = dFsrc;
    }
    st = igab_load_vector(p, eb, ee, nq1, xi1d, w1d, fsrc, ncomp, dFe, IGAB_DEVICE, stream);
    long long k1 = k0;
    while (k1 < kEnd && elem[k1] < ee) ++k1;
    if (!st && k1 > k0) {
      const long long m = k1 - k0, first = k0 - kFirst;
      put_trimmed_kernel<<<(unsigned)((m * n + 255) / 256), 256, 0, stream>>>(n, m, dElem + first, eb, dFt + first * n, dFe);
      count_launch();
      if ((e = cudaGetLastError()) != cudaSuccess) st = cuda_fail(e, "global_load_vector: put_trimmed");
    }
    k0 = k1;
    if (!st) st = rhs_assemble(p, ncomp, eb, ee, dFe, dB, eb > elem_begin, IGAB_DEVICE, stream);
    if (!st && space == IGAB_HOST) {
      e = cudaStreamSynchronize(stream);  // dFsrc is reused by the next chunk
      if (e != cudaSuccess) st = cuda_fail(e, "global_load_vector: sync");
    }
  }
  if (!st && space == IGAB_HOST) {
    e = cudaMemcpyAsync(b, dB, sizeof(double) * nrows, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) st = cuda_fail(e, "global_load_vector: D2H");
  }
  if (dFe) cudaFreeAsync(dFe, stream);
  if (space == IGAB_HOST && dB) cudaFreeAsync(dB, stream);
  if (dFsrc) cudaFreeAsync(dFsrc, stream);
  if (dFt) cudaFreeAsync(dFt, stream);
  if (dElem) cudaFreeAsync(dElem, stream);
  return st;
}

// ---- reductions -----------------------------------------------------------------------------------------------------
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
static nccl_allreduce_fn resolve_nccl_allreduce() {
  // NCCL is resolved at run time from the process (torch ships libnccl.so.2) so that the library has no link-time
  // dependency on it; used only when the caller passes a communicator.
  static const nccl_allreduce_fn fn = [] {  // initialised once, thread-safe (C++11 static)
    void* sym = dlsym(RTLD_DEFAULT, "ncclAllReduce");
    if (!sym) {
      void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (h) sym = dlsym(h, "ncclAllReduce");
    }
    return (nccl_allreduce_fn)sym;
  }();
  return fn;
}

namespace {
// sum_k detJ[k] w[k] over a point list, same fixed two-level tree as the tensor part
__global__ void __launch_bounds__(256) volume_points_partial_kernel(long long npts, const double* __restrict__ detJ,
                                                                    const double* __restrict__ w,
                                                                    double* __restrict__ partial) {
  __shared__ double sh[256];
  const long long per = (npts + gridDim.x - 1) / gridDim.x;
  const long long b0 = (long long)blockIdx.x * per, b1 = min(npts, b0 + per);
  double acc = 0.0;
  for (long long k = b0 + threadIdx.x; k < b1; k += 256) acc += detJ[k] * w[k];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
// *result += fixed-order sum of the partials
__global__ void __launch_bounds__(1024) volume_points_final_kernel(int n, const double* __restrict__ partial,
                                                                   double* __restrict__ result) {
  __shared__ double sh[1024];
  sh[threadIdx.x] = threadIdx.x < n ? partial[threadIdx.x] : 0.0;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *result += sh[0];
}
}  // namespace

igab_status igab_reduce_volume(igab_patch* p, int64_t elem_begin, int64_t elem_end, const int* nq1,
                               const double* const* xi1d, const double* const* w1d, double* result, void* nccl_comm,
                               void* stream_) {
  return igab_reduce_volume_trimmed(p, elem_begin, elem_end, nq1, xi1d, w1d, 0, nullptr, nullptr, nullptr, result,
                                    nccl_comm, stream_);
}

igab_status igab_reduce_volume_trimmed(igab_patch* p, int64_t elem_begin, int64_t elem_end, const int* nq1,
                                       const double* const* xi1d, const double* const* w1d, int64_t npts_list,
                                       const int32_t* elem, const double* xi, const double* w, double* result,
                                       void* nccl_comm, void* stream_) {
  if (!p || !nq1 || !xi1d || !w1d || !result || npts_list < 0 || (npts_list > 0 && (!elem || !xi || !w))) {
    set_error("reduce_volume: null argument");
    return IGAB_INVALID_ARGUMENT;
  }
  for (int64_t k = 0; k < npts_list; ++k)
    if (elem[k] < 0 || elem[k] >= p->dev.nelem) {
      set_error("reduce_volume: element index of a list point out of range");
      return IGAB_INVALID_ARGUMENT;
    }
  if (elem_begin < 0 || elem_end < elem_begin || elem_end > p->dev.nelem) {
    set_error("reduce_volume: element range outside [0, n_elem]");
    return IGAB_INVALID_ARGUMENT;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  if (igab_status est = enter_stream(p, stream)) return est;
  RuleDev rule;
  igab_status st = stage_rule(p, nq1, xi1d, w1d, &rule, stream);
  if (st) return st;
  const int npartial = 1024;
  if (!p->d_red) {
    IGAB_CUDA_TRY(cudaMalloc(&p->d_red, sizeof(double) * (2 * npartial + 2)));
    p->red_cap = npartial;
  }
  double* res_dev = p->d_red + 2 * npartial;
  // (a fused sum-factorised kernel like the one of the norms was tried for the tensor part: 10.0 ms against 8.3 ms for detJ by
  // the evaluation kernel + tree reduction on the C2 patch -- the block-per-element overhead, not HBM, bounds both)
  st = launch_volume(p->dev, p->sf, xi1d, elem_begin, elem_end - elem_begin, nq1, rule.xi, rule.w, p->d_red, npartial, res_dev,
                     stream, p->d_trim);
  if (st) return st;
  if (npts_list > 0) {
    // the trimmed elements' own rules (fillquadraturerule.hh:8-19): detJ at the list points, sum_k detJ w added on
    int* d_elem = nullptr;
    double *d_xi = nullptr, *d_w = nullptr, *d_det = nullptr;
    const int dim = p->dev.dim;
    cudaError_t e = cudaMallocAsync(&d_elem, sizeof(int) * npts_list, stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&d_xi, sizeof(double) * npts_list * dim, stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&d_w, sizeof(double) * npts_list, stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&d_det, sizeof(double) * npts_list, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_elem, elem, sizeof(int) * npts_list, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_xi, xi, sizeof(double) * npts_list * dim, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_w, w, sizeof(double) * npts_list, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) st = cuda_fail(e, "reduce_volume: list staging");
    if (!st) {
      PointSource src{};
      src.tensor = 0;
      src.elem = d_elem;
      src.xi = d_xi;
      EvalOutDev o{};
      o.detJ = d_det;
      st = run_eval_points_device(p, src, npts_list, 1, o, stream);
    }
    if (!st) {
      const int nblk = (int)std::min<long long>(npartial, (npts_list + 255) / 256);
      volume_points_partial_kernel<<<nblk, 256, 0, stream>>>(npts_list, d_det, d_w, p->d_red);
      volume_points_final_kernel<<<1, 1024, 0, stream>>>(nblk, p->d_red, res_dev);
      count_launch(2);
      if ((e = cudaGetLastError()) != cudaSuccess) st = cuda_fail(e, "reduce_volume: list reduction");
    }
    // host list arrays must stay valid until the copies have run
    cudaStreamSynchronize(stream);
    if (d_elem) cudaFreeAsync(d_elem, stream);
    if (d_xi) cudaFreeAsync(d_xi, stream);
    if (d_w) cudaFreeAsync(d_w, stream);
    if (d_det) cudaFreeAsync(d_det, stream);
    if (st) return st;
  }
  if (nccl_comm) {
    nccl_allreduce_fn ar = resolve_nccl_allreduce();
    if (!ar) {
      set_error("reduce_volume: NCCL communicator given but libnccl.so.2 could not be resolved");
      return IGAB_NCCL_ERROR;
    }
    // ncclDouble = 8, ncclSum = 0 (nccl.h); in place on the one-element device scalar
    const int rc = ar(res_dev, res_dev, 1, 8, 0, nccl_comm, stream);
    if (rc != 0) {
      set_error("reduce_volume: ncclAllReduce failed with code " + std::to_string(rc));
      return IGAB_NCCL_ERROR;
    }
  }
  IGAB_CUDA_TRY(cudaMemcpyAsync(result, res_dev, sizeof(double), cudaMemcpyDeviceToHost, stream));
  IGAB_CUDA_TRY(cudaStreamSynchronize(stream));
  return IGAB_OK;
}

igab_status igab_reduce_norms(igab_patch* p, int64_t elem_begin, int64_t elem_end, const int* nq1,
                              const double* const* xi1d, const double* const* w1d, const double* u, int ncomp,
                              igab_memspace u_space, double* result, void* nccl_comm, void* stream_) {
  return igab_reduce_norms_trimmed(p, elem_begin, elem_end, nq1, xi1d, w1d, u, ncomp, u_space, 0, nullptr, nullptr, nullptr,
                                   result, nccl_comm, stream_);
}

igab_status igab_reduce_norms_trimmed(igab_patch* p, int64_t elem_begin, int64_t elem_end, const int* nq1,
                                      const double* const* xi1d, const double* const* w1d, const double* u, int ncomp,
                                      igab_memspace u_space, int64_t npts_list, const int32_t* elem, const double* xi,
                                      const double* w, double* result, void* nccl_comm, void* stream_) {
  if (!p || !nq1 || !xi1d || !w1d || !u || !result || npts_list < 0 || (npts_list > 0 && (!elem || !xi || !w))) {
    set_error("reduce_norms: null argument");
    return IGAB_INVALID_ARGUMENT;
  }
  for (int64_t k = 0; k < npts_list; ++k)
    if (elem[k] < 0 || elem[k] >= p->dev.nelem) {
      set_error("reduce_norms: element index of a list point out of range");
      return IGAB_INVALID_ARGUMENT;
    }
  if (elem_begin < 0 || elem_end < elem_begin || elem_end > p->dev.nelem) {
    set_error("reduce_norms: element range outside [0, n_elem]");
    return IGAB_INVALID_ARGUMENT;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  if (igab_status est = enter_stream(p, stream)) return est;
  RuleDev rule;
  igab_status st = stage_rule(p, nq1, xi1d, w1d, &rule, stream);
  if (st) return st;
  long long ndof = p->ndof_untrimmed;
  if (p->d_trim) {  // u is numbered like the trimmed basis (prepareForTrim, nurbsbasis.hh:599-615)
    if ((st = ensure_dof_compaction(p))) return st;
    ndof = p->ndof_compact;
  }
  const int npartial = 1024;
  if (!p->d_norm) IGAB_CUDA_TRY(cudaMalloc(&p->d_norm, sizeof(double) * (2 * npartial + 2)));
  double* res_dev = p->d_norm + 2 * npartial;
  const double* u_dev = u;
  double* u_tmp = nullptr;
  if (u_space == IGAB_HOST) {
    const size_t bytes = sizeof(double) * (size_t)ndof * ncomp;
    IGAB_CUDA_TRY(cudaMallocAsync(&u_tmp, bytes, stream));
    IGAB_CUDA_TRY(cudaMemcpyAsync(u_tmp, u, bytes, cudaMemcpyHostToDevice, stream));
    u_dev = u_tmp;
  }
  int* d_le = nullptr;
  double *d_lx = nullptr, *d_lw = nullptr;
  if (npts_list > 0) {  // the trimmed elements' own points (HOST arrays)
    const int dim = p->dev.dim;
    cudaError_t e = cudaMallocAsync(&d_le, sizeof(int) * npts_list, stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&d_lx, sizeof(double) * npts_list * dim, stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&d_lw, sizeof(double) * npts_list, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_le, elem, sizeof(int) * npts_list, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_lx, xi, sizeof(double) * npts_list * dim, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_lw, w, sizeof(double) * npts_list, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) st = cuda_fail(e, "reduce_norms: list staging");
  }
  if (!st) {
    // tensor part: sum-factorised kernel where one is compiled (3-D, equal degrees), else thread per point
    bool tensor_done = false;
    const igab_status sst = launch_norms_sf(p->dev, p->sf, xi1d, elem_begin, elem_end - elem_begin, nq1, rule.xi, rule.w,
                                            u_dev, ncomp, p->d_norm, npartial, res_dev, stream, p->d_trim, p->d_dofc_map);
    if (sst == IGAB_OK) tensor_done = true;
    else if (sst != IGAB_UNSUPPORTED) st = sst;
    if (!st && !(tensor_done && npts_list == 0))
      st = launch_norms(p->dev, elem_begin, elem_end - elem_begin, nq1, rule.xi, rule.w, u_dev, ncomp, p->d_norm, npartial,
                        res_dev, stream, p->d_trim, p->d_dofc_map, npts_list, d_le, d_lx, d_lw, tensor_done);
  }
  if (npts_list > 0) cudaStreamSynchronize(stream);  // the host list arrays may go away after the call
  if (d_le) cudaFreeAsync(d_le, stream);
  if (d_lx) cudaFreeAsync(d_lx, stream);
  if (d_lw) cudaFreeAsync(d_lw, stream);
  if (u_tmp) cudaFreeAsync(u_tmp, stream);
  if (st) return st;
  if (nccl_comm) {
    nccl_allreduce_fn ar = resolve_nccl_allreduce();
    if (!ar) {
      set_error("reduce_norms: NCCL communicator given but libnccl.so.2 could not be resolved");
      return IGAB_NCCL_ERROR;
    }
    const int rc = ar(res_dev, res_dev, 2, 8, 0, nccl_comm, stream);  // ncclDouble, ncclSum
    if (rc != 0) {
      set_error("reduce_norms: ncclAllReduce failed with code " + std::to_string(rc));
      return IGAB_NCCL_ERROR;
    }
  }
  IGAB_CUDA_TRY(cudaMemcpyAsync(result, res_dev, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
  IGAB_CUDA_TRY(cudaStreamSynchronize(stream));
  return IGAB_OK;
}

}  // extern "C"

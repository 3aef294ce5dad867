// Internal declarations shared by the translation units of libigab200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/igab200.h"

namespace igab {

constexpr int kMaxDim = 3;
constexpr int kMaxDegree = 8;           // per direction
constexpr int kMaxOrder1D = kMaxDegree + 1;

// Everything a kernel needs to know about a patch; passed by value (lives in constant param space).
struct PatchDev {
  int dim, dw;
  int degree[kMaxDim];
  int nknots[kMaxDim];
  int ncp[kMaxDim];
  int nel[kMaxDim];               // elements (non-empty spans) per direction
  const double* knots[kMaxDim];   // device
  const int* spanTab[kMaxDim];    // device: span index of element e_d, bit-exact with findSpanCorrected(p,uq[e],U,e)
  const double* rden[kMaxDim];    // device [nel_d][p(p+1)/2]: reciprocals of the knot differences A2.3 divides by (per span)
  const double* cp;               // device [prod ncp][dw+1]  (x.., w)
  // 3-D patches whose weights are a tensor product w(i,j,k) = a_i b_j c_k (B-splines, conics, revolved / extruded /
  // ruled solids and every refinement of them): the per-direction factors on the device, else nullptr.  The assembly
  // kernels then carry the weights inside their univariate tables and skip the w_i w_j scaling of every stored entry.
  const double* wfac[kMaxDim];
  int nb;                         // prod (degree+1)
  long long nelem;
};

// Where the points of a batch come from.
struct PointSource {
  int tensor;                     // 1: tensor rule over elements [elem_begin, ...), 0: explicit list
  long long elem_begin;
  int nq1[kMaxDim];
  long long nq;                   // prod nq1
  const double* xi1d[kMaxDim];    // device
  const int* elem;                // device [npts]     (list mode)
  const double* xi;               // device [npts][dim] (list mode)
};

struct EvalOutDev {
  double *N, *dN, *d2N, *x, *Jt, *H, *detJ, *JinvT;
};

// Per-handle workspace of the sum-factorised tensor kernel: univariate tables of the last rule (K1 output) and the
// element counter of the dynamic scheduler.  Rebuilt only when the rule changes.
struct SfWorkspace {
  double* tab[kMaxDim] = {nullptr, nullptr, nullptr};
  size_t tabCap[kMaxDim] = {0, 0, 0};
  int nq1[kMaxDim] = {0, 0, 0};
  std::vector<double> xi[kMaxDim];
  unsigned int* counter = nullptr;  // ring of counters, one per launch in flight
  unsigned int counterSlot = 0;
  bool valid = false;
  cudaStream_t stream = nullptr;    // stream the tables were built on (rebuilt when the caller switches streams)
  // tables with second derivatives ([e][q][i][4]) of the order-2 tensor kernel, cached the same way
  double* tabO2[kMaxDim] = {nullptr, nullptr, nullptr};
  size_t tabO2Cap[kMaxDim] = {0, 0, 0};
  std::vector<double> xiO2[kMaxDim];
  bool validO2 = false;
  cudaStream_t streamO2 = nullptr;
  // the same tables multiplied by the weight factor of their control point (tensor-product weights only)
  double* wtab[kMaxDim] = {nullptr, nullptr, nullptr};
  size_t wtabCap[kMaxDim] = {0, 0, 0};
  bool wvalid = false, wuse = false;  // wuse: set per launch (off inside a graph capture)
  bool built2d = false;  // ensure_tables2d launched its table kernels in the current call
};

// orders `stream` behind the handle's previous stream when the caller switches streams (handle scratch is shared)
igab_status enter_stream(igab_patch* p, cudaStream_t stream);
void set_error(const std::string& msg);
igab_status cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);
// Programmatic dependent launch (eval_tensor_2d.cu) is for device-resident callers.  The HOST-space entry points recycle
// staging buffers behind events of a copy stream; their kernels are launched plainly (PCIe bounds them anyway), so that
// nothing rests on how a programmatic edge combines with an event wait in front of the kernel.
struct NoPdlScope {
  NoPdlScope();
  ~NoPdlScope();
};
bool pdl_allowed();

// Per (kernel, device) launch configuration: raises the dynamic shared-memory limit once and caches the SM count and the
// resident blocks per SM.  Thread-safe (handles may be driven from different host threads) and correct when a process
// uses more than one device.
struct KernelCfg {
  int sms = 148, blocksPerSm = 1;
};
igab_status kernel_config(const void* func, int threads, size_t smem, KernelCfg* out);

#define IGAB_CUDA_TRY(expr)                                           \
  do {                                                                \
    cudaError_t _e = (expr);                                          \
    if (_e != cudaSuccess) return ::igab::cuda_fail(_e, #expr);       \
  } while (0)

// ---- kernels' host launchers (each returns after enqueueing on `stream`)
igab_status launch_eval_generic(const PatchDev& P, const PointSource& src, long long npts, int order,
                                const EvalOutDev& out, cudaStream_t stream);
// general tensor-rule kernel, one warp per element (eval_warp.cu); IGAB_UNSUPPORTED when shared memory does not fit
igab_status launch_eval_warp(const PatchDev& P, const PointSource& src, long long nelem, int order,
                             const EvalOutDev& out, cudaStream_t stream);
// returns IGAB_UNSUPPORTED when no sum-factorised instantiation matches
// xi1d_host: the rule on the host (used to decide whether the cached tables are still valid)
igab_status launch_eval_tensor_sf(const PatchDev& P, SfWorkspace& ws, const PointSource& src,
                                  const double* const* xi1d_host, long long nelem, int order, const EvalOutDev& out,
                                  cudaStream_t stream);
void free_sf_workspace(SfWorkspace& ws);
igab_status ensure_tables3d(const PatchDev& Pd, SfWorkspace& ws, int Q, const double* const* xi1d_dev,
                            const double* const* xi1d_host, cudaStream_t stream);
// weighted copies of the tables (after ensure_tables3d); no-op unless Pd.wfac is set
igab_status ensure_wtables3d(const PatchDev& Pd, SfWorkspace& ws, int Q, cudaStream_t stream);
// fully fused element matrices (3-D Poisson / mass / elasticity, p = 2, 3, 4)
bool gram_fused_supported(const PatchDev& P, int kind, const int* nq1);
// assemble_sym.cu: sum-factorised Poisson / mass element matrices, upper block triangle + mirror (the default at p >= 3)
bool gram_sym_supported(int p, int q, int kind);
bool gram_sym_default(int p, int q, int kind);
igab_status launch_gram_sym(const PatchDev& Pd, SfWorkspace& ws, int kind, int q, double fconst, long long eb,
                            long long nel, const double* const* w_dev, double* Ke, double* fe, cudaStream_t stream);
igab_status launch_gram_fused(const PatchDev& P, SfWorkspace& ws, const double* const* xi1d_host,
                              const double* const* xi1d_dev, const double* const* w1d_dev, int kind, double fconst,
                              long long eb, long long nel, const int* nq1, double* Ke, double* fe, cudaStream_t stream,
                              const double* D_host);
// fused 2-D element matrices (p = 1..3, (p+1)^2 Gauss): no HBM round trip of the B-stacks
bool assemble2d_supported(const PatchDev& P, int kind, const int* nq1);
igab_status launch_assemble2d(const PatchDev& P, SfWorkspace& ws, const double* const* xi1d_host, int kind,
                              const double* D_host, double fconst, long long eb, long long nel, const int* nq1,
                              const double* const* xi1d_dev, const double* const* w1d_dev, double* Ke, double* fe,
                              cudaStream_t stream);
bool assemble2d_points_supported(const PatchDev& P, int kind);
igab_status launch_assemble2d_points(const PatchDev& P, int kind, const double* D_host, double fconst, long long nlist,
                                     const int* elem, const long long* offset, const double* xi, const double* w,
                                     double* Ke, double* fe, cudaStream_t stream);
igab_status ensure_tables2d(const PatchDev& Pd, SfWorkspace& ws, int Q, int order, const double* const* xi1d_dev,
                            const double* const* xi1d_host, cudaStream_t stream);
bool eval_tensor_sf_supported(const PatchDev& P, const int* nq1, int order, const EvalOutDev& out);
// 3-D with second derivatives (p = 2, 3)
bool eval_tensor_sf2_supported(const PatchDev& P, const int* nq1, int order, const EvalOutDev& out);
igab_status launch_eval_tensor_sf2(const PatchDev& P, SfWorkspace& ws, const PointSource& src,
                                   const double* const* xi1d_host, long long nelem, const EvalOutDev& out,
                                   cudaStream_t stream);
// 3-D, large local bases (p = 4): one block per element
bool eval_tensor_blk_supported(const PatchDev& P, const int* nq1, int order, const EvalOutDev& out);
igab_status launch_eval_tensor_blk(const PatchDev& P, SfWorkspace& ws, const PointSource& src,
                                   const double* const* xi1d_host, long long nelem, int order, const EvalOutDev& out,
                                   cudaStream_t stream);
// 2-D point lists (trimmed batches): per-point univariate rows in registers
bool eval_points_2d_supported(const PatchDev& P);
igab_status launch_eval_points_2d(const PatchDev& P, const PointSource& src, long long npts, int order,
                                  const EvalOutDev& out, cudaStream_t stream);
// 2-D family (thread per point, compile-time degree / rule / order)
bool eval_tensor_2d_supported(const PatchDev& P, const int* nq1, int order, const EvalOutDev& out);
igab_status launch_eval_tensor_2d(const PatchDev& P, SfWorkspace& ws, const PointSource& src,
                                  const double* const* xi1d_host, long long nelem, int order, const EvalOutDev& out,
                                  cudaStream_t stream);
igab_status launch_partial(const PatchDev& P, long long npts, const int* elem, const double* xi, const int* alpha,
                           double* out, cudaStream_t stream);
igab_status launch_spans(const PatchDev& P, int32_t* span, cudaStream_t stream);
igab_status launch_dofmap(const PatchDev& P, int32_t* dof, cudaStream_t stream);
igab_status launch_dof_compact(const PatchDev& P, const uint8_t* trim_dev, int32_t* dof, int32_t* scratch_used,
                               int32_t* scratch_map, long long ndof, long long* ndof_out_host, cudaStream_t stream);
// ordered exclusive scan of a 0/1 predicate (used[i] != 0, or trim[i] != EMPTY when used is null): map[i] = rank or -1,
// inv[rank] = i, *total_dev = number kept; multi-block, exact
igab_status launch_flag_scan(const int32_t* used, const uint8_t* trim, long long n, int32_t* map, int32_t* inv,
                             long long* total_dev, cudaStream_t stream);
igab_status launch_assemble(const PatchDev& P, SfWorkspace& ws, const double* const* xi1d_host, int kind, const double* D_host, const double* D_dev, double fconst, long long elem_begin,
                            long long nelem, const int* nq1, const double* const* xi1d_dev,
                            const double* const* w1d_dev, double* Ke, double* fe, cudaStream_t stream);
igab_status launch_volume(const PatchDev& P, SfWorkspace& ws, const double* const* xi1d_host, long long elem_begin, long long nelem, const int* nq1,
                          const double* const* xi1d_dev, const double* const* w1d_dev, double* partial_dev,
                          int npartial, double* result_dev, cudaStream_t stream, const uint8_t* trim_dev = nullptr);

igab_status refine_apply(int dim, int dw, const int* ncp_old, int direction, int n_new, int maxlen, const int* first,
                         const int* len, const double* coef, const double* cp_in, double* cp_out, cudaStream_t stream);
// knot refinement of all directions in HBM: Oslo operators built on the device, banded apply per direction; knots are HOST
// arrays, the nets device arrays (cp_out_dev and scratch_dev hold the refined net's size)
igab_status refine_knots_device(int dim, int dw, const int* degree, const int* nknots_old,
                                const double* const* knots_old, const int* nknots_new, const double* const* knots_new,
                                const double* cp_in_dev, double* cp_out_dev, double* scratch_dev, cudaStream_t stream);
igab_status eval_faces_device(igab_patch* p, long long nfaces, const int* elem, const int* face, const int* nqf1,
                              const double* const* xi_dev, double* x, double* normal, double* dS, cudaStream_t stream);
igab_status launch_norms(const PatchDev& P, long long elem_begin, long long nelem, const int* nq1,
                         const double* const* xi1d_dev, const double* const* w1d_dev, const double* u_dev, int ncomp,
                         double* partial_dev, int npartial, double* result_dev, cudaStream_t stream,
                         const uint8_t* trim_dev = nullptr, const int32_t* map_dev = nullptr, long long npts_list = 0,
                         const int32_t* lelem_dev = nullptr, const double* lxi_dev = nullptr,
                         const double* lw_dev = nullptr, bool tensor_done = false);
// sum-factorised tensor part (3-D, equal degrees 2..4); IGAB_UNSUPPORTED when it does not apply
igab_status launch_norms_sf(const PatchDev& P, SfWorkspace& ws, const double* const* xi1d_host, long long elem_begin,
                            long long nelem, const int* nq1, const double* const* xi1d_dev,
                            const double* const* w1d_dev, const double* u_dev, int ncomp, double* partial_dev,
                            int npartial, double* result_dev, cudaStream_t stream, const uint8_t* trim_dev,
                            const int32_t* map_dev);
// trimmed patches: builds (once) the compacted DOF numbering of prepareForTrim in the handle (d_dofc_map / d_dofc_inv)
igab_status ensure_dof_compaction(igab_patch* p);
igab_status geometry_local_device(const PatchDev& P, long long npts, const int* elem, const double* xg, double* xi,
                                  int* flag, cudaStream_t stream);
igab_status surface_forms_device(int dim, int dw, long long npts, const double* Jt, const double* H, double* normal,
                                 double* unitNormal, double* metric, double* secondFF, double* gaussK,
                                 cudaStream_t stream);
igab_status csr_pattern(igab_patch* p, int ncomp, int64_t* nrows, int64_t* nnz, int64_t* rowptr, int32_t* colidx,
                        igab_memspace space, cudaStream_t stream);
igab_status csr_assemble(igab_patch* p, int ncomp, long long eb, long long ee, const double* Ke, const int64_t* rowptr,
                         double* values, int accumulate, igab_memspace space, cudaStream_t stream);

igab_status rhs_assemble(igab_patch* p, int ncomp, long long eb, long long ee, const double* fe, double* b, int accumulate,
                         igab_memspace space, cudaStream_t stream);
igab_status csr_dirichlet(igab_patch* p, int ncomp, const uint8_t* mask, const double* gval, double* values, double* b,
                          igab_memspace space, cudaStream_t stream);

}  // namespace igab

// The host-side object behind the opaque handle.
struct igab_patch {
  igab::PatchDev dev{};
  int device = 0;
  int kernel_mode = IGAB_KERNEL_AUTO;
  std::vector<double> hknots[igab::kMaxDim];
  std::vector<int> hspan[igab::kMaxDim];
  std::vector<uint8_t> htrim;
  double* d_knots[igab::kMaxDim] = {nullptr, nullptr, nullptr};
  int* d_span[igab::kMaxDim] = {nullptr, nullptr, nullptr};
  double* d_rden[igab::kMaxDim] = {nullptr, nullptr, nullptr};
  double* d_cp = nullptr;
  double* d_wfac[igab::kMaxDim] = {nullptr, nullptr, nullptr};
  bool wfac_known = false;  // has the tensor-product check run for the current control net?
  uint8_t* d_trim = nullptr;
  long long ncp_total = 0;
  long long ndof_untrimmed = 0;
  // small device scratch for quadrature rules (xi, w per direction), reused across calls
  double* d_rule = nullptr;  // [2][kMaxDim][kMaxRule]
  igab::SfWorkspace sf;
  // last rule staged in d_rule (skip the H2D copy when the caller passes the same rule again)
  std::vector<double> rule_host;
  cudaStream_t rule_stream = nullptr;
  // persistent staging for IGAB_HOST outputs of igab_eval_tensor: two sets of device buffers, a copy stream, events
  double* stage_buf[2] = {nullptr, nullptr};
  size_t stage_cap = 0;  // bytes per set
  cudaStream_t stage_stream = nullptr;
  cudaEvent_t stage_done[2] = {nullptr, nullptr}, stage_copied[2] = {nullptr, nullptr};
  // CSR of the global matrix: first-DOF -> element tables, cached row pointer
  int* d_elem_of_first[igab::kMaxDim] = {nullptr, nullptr, nullptr};
  std::vector<long long> csr_rowptr_host;
  long long* d_csr_rowptr = nullptr;
  int csr_ncomp = 0;
  // trimmed patches: compacted DOF numbering (prepareForTrim) and the cached column indices of its pattern
  int32_t* d_dofc_map = nullptr;   // original scalar DOF -> compacted (-1: untouched)
  int32_t* d_dofc_inv = nullptr;   // compacted -> original
  long long ndof_compact = 0;
  int32_t* d_csr_colidx = nullptr;
  // igab_global_assemble: element-matrix chunk and (host-space calls) value / load-vector staging, kept between calls
  double* ga_scratch = nullptr;
  size_t ga_cap = 0;  // doubles
  // igab_exchange_interface_rows: receive buffers / staged pieces, host copy of the compaction map (trimmed patches)
  double* xch_buf = nullptr;
  size_t xch_cap = 0;  // doubles
  std::vector<int32_t> dofc_map_host;
  // igab_eval_tensor_begin in flight (host outputs still landing): collected by igab_patch_wait / the next call
  bool pending_host = false;
  cudaStream_t pending_host_stream = nullptr;
  // stream discipline: the stream of the last call, and the event that orders a new stream behind it
  cudaStream_t last_stream = nullptr;
  bool have_last_stream = false;
  cudaEvent_t stream_switch_ev = nullptr;
  // elasticity matrix D of igab_assemble, kept on the device (restaged only when it changes)
  double* d_D = nullptr;
  std::vector<double> D_host;
  cudaStream_t D_stream = nullptr;
  // igab_assemble(IGAB_HOST): two element-matrix chunks (+ element vectors) so that D2H overlaps the next chunk's kernel
  double* asm_buf[2] = {nullptr, nullptr};
  size_t asm_cap = 0;  // doubles per buffer
  // reduction scratch
  double* d_red = nullptr;
  int red_cap = 0;
  double* d_norm = nullptr;
};

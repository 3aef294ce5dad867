// refine.cu -- control-net side of knot refinement / degree elevation on the GPU.
// The reference refines direction by direction with per-hyper-surface std::function views (knotRefinement,
// dune/iga/nurbsalgorithms.hh:441-528; degreeElevate :264-439; drivers dune/iga/nurbsgrid.hh:130-165), in homogeneous
// coordinates (scale by w on entry :289-291,476, divide on exit :432-434,524-525).  Here the 1-D operator
// T (n_new x n_old), banded with <= maxlen non-zeros per row, is built once on the host
// (dune-iga_b200/patchdata.py) and applied to the whole net in one launch: one thread per NEW control point,
// out(.., i, ..) = sum_k coef[i][k] * in(.., first[i]+k, ..) on (w x, w) with the division by w fused.
#include "igab_internal.cuh"

namespace igab {

namespace {

template <int DW>
__global__ void refine_apply_kernel(long long nout, int inner, int n_old, int n_new, int maxlen,
                                    const int* __restrict__ first, const int* __restrict__ len,
                                    const double* __restrict__ coef, const double* __restrict__ cp_in,
                                    double* __restrict__ cp_out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nout) return;
  // new net index = a + inner * (i + n_new * b): a = faster directions, b = slower directions
  const int a = (int)(t % inner);
  const long long r = t / inner;
  const int i = (int)(r % n_new);
  const long long b = r / n_new;
  double acc[DW + 1];
#pragma unroll
  for (int c = 0; c <= DW; ++c) acc[c] = 0.0;
  const int f = first[i], l = len[i];
  for (int k = 0; k < l; ++k) {
    const double* p = cp_in + ((a + (long long)inner * ((f + k) + (long long)n_old * b)) * (DW + 1));
    const double w = p[DW], cw = coef[(long long)i * maxlen + k] * w;
#pragma unroll
    for (int c = 0; c < DW; ++c) acc[c] = fma(cw, p[c], acc[c]);
    acc[DW] += cw;
  }
  double* o = cp_out + t * (DW + 1);
  const double iw = 1.0 / acc[DW];
#pragma unroll
  for (int c = 0; c < DW; ++c) o[c] = acc[c] * iw;
  o[DW] = acc[DW];
}

}  // namespace

igab_status refine_apply(int dim, int dw, const int* ncp_old, int direction, int n_new, int maxlen, const int* first,
                         const int* len, const double* coef, const double* cp_in, double* cp_out, cudaStream_t stream) {
  long long inner = 1, outer = 1;
  for (int d = 0; d < direction; ++d) inner *= ncp_old[d];
  for (int d = direction + 1; d < dim; ++d) outer *= ncp_old[d];
  const long long nout = inner * n_new * outer;
  if (inner > 0x7fffffffLL) {
    set_error("refine: net too large");
    return IGAB_INVALID_ARGUMENT;
  }
  const int th = 256;
  const unsigned blocks = (unsigned)((nout + th - 1) / th);
  if (dw == 1)
    refine_apply_kernel<1><<<blocks, th, 0, stream>>>(nout, (int)inner, ncp_old[direction], n_new, maxlen, first, len, coef, cp_in, cp_out);
  else if (dw == 2)
    refine_apply_kernel<2><<<blocks, th, 0, stream>>>(nout, (int)inner, ncp_old[direction], n_new, maxlen, first, len, coef, cp_in, cp_out);
  else
    refine_apply_kernel<3><<<blocks, th, 0, stream>>>(nout, (int)inner, ncp_old[direction], n_new, maxlen, first, len, coef, cp_in, cp_out);
  count_launch();
  IGAB_CUDA_TRY(cudaGetLastError());
  return IGAB_OK;
}

}  // namespace igab

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

// Knot-insertion operator of one direction, built ON THE DEVICE: row i of T (new basis function i as a combination of the
// old ones) by the Oslo recursion for discrete B-splines (Cohen / Lyche / Riesenfeld): with mu the old span that holds
// the new knot Ubar[i],  alpha_{j,0} = [j == mu],
//     alpha_{j,k} = (Ubar[i+k] - U[j]) / (U[j+k] - U[j]) alpha_{j,k-1} + (U[j+k+1] - Ubar[i+k]) / (U[j+k+1] - U[j+1]) alpha_{j+1,k-1}
// (terms with a zero denominator dropped), non-zero for j = mu - p .. mu.  Same exact affine combinations as the
// reference's sequential insertion (knotRefinement, dune/iga/nurbsalgorithms.hh:441-528 = NURBS book A5.4) up to rounding.
// One thread per new row; first[i] = mu - p (clamped), len[i] = p + 1, coef[i][0..p].
__global__ void oslo_operator_kernel(int p, const double* __restrict__ U, int nU, const double* __restrict__ Ub, int nUb,
                                     int* __restrict__ first, int* __restrict__ len, double* __restrict__ coef) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_new = nUb - p - 1, n_old = nU - p - 1;
  if (i >= n_new) return;
  // mu: U[mu] <= Ub[i] < U[mu+1], the last non-empty span for the end knot
  int lo = p, hi = n_old;  // search in [p, n_old]: U[p] .. U[n_old] are the span boundaries of the open knot vector
  const double u = Ub[i];
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (U[mid] <= u) lo = mid; else hi = mid;
  }
  int mu = lo;
  if (u < U[p]) {  // the leading repeated knots of Ubar lie on U[0] = ... = U[p]: span of the first old function
    mu = 0;
    while (mu + 1 < nU && U[mu + 1] <= u) ++mu;
  }
  double a[kMaxDegree + 2];
  for (int k = 0; k <= p + 1; ++k) a[k] = 0.0;
  // a[r] holds alpha_{mu - k + r, k}, r = 0..k
  a[0] = 1.0;
  for (int k = 1; k <= p; ++k) {
    double nxt[kMaxDegree + 2];
    for (int r = 0; r <= k; ++r) {
      const int j = mu - k + r;
      double v = 0.0;
      // alpha_{j,k-1} lives at index r - 1 of the previous level (j = mu - (k-1) + (r-1)), alpha_{j+1,k-1} at index r
      const double t = Ub[i + k];
      if (r >= 1 && j >= 0 && j + k < nU) {
        const double d = U[j + k] - U[j];
        if (d > 0.0) v += (t - U[j]) / d * a[r - 1];
      }
      if (r <= k - 1 && j + 1 >= 0 && j + k + 1 < nU) {
        const double d = U[j + k + 1] - U[j + 1];
        if (d > 0.0) v += (U[j + k + 1] - t) / d * a[r];
      }
      nxt[r] = v;
    }
    for (int r = 0; r <= k; ++r) a[r] = nxt[r];
  }
  // clip to the old basis 0 .. n_old - 1
  int f = mu - p, l = p + 1, off = 0;
  if (f < 0) { off = -f; l -= off; f = 0; }
  if (f + l > n_old) l = n_old - f;
  first[i] = f;
  len[i] = l;
  for (int k = 0; k <= p; ++k) coef[(long long)i * (p + 1) + k] = (k < l) ? a[k + off] : 0.0;
}

}  // namespace

igab_status refine_knots_device(int dim, int dw, const int* degree, const int* nknots_old,
                                const double* const* knots_old, const int* nknots_new, const double* const* knots_new,
                                const double* cp_in_dev, double* cp_out_dev, double* scratch_dev, cudaStream_t stream) {
  // scratch_dev: a second net-sized buffer for the ping-pong; operators in stream-ordered scratch
  int ncp[kMaxDim];
  for (int d = 0; d < dim; ++d) ncp[d] = nknots_old[d] - degree[d] - 1;
  int changed = 0;
  for (int d = 0; d < dim; ++d) changed += nknots_new[d] != nknots_old[d];
  const double* src = cp_in_dev;
  int step = 0;
  for (int d = 0; d < dim; ++d) {
    if (nknots_new[d] == nknots_old[d]) continue;
    const int p = degree[d], n_new = nknots_new[d] - p - 1;
    double *dU = nullptr, *dUb = nullptr, *dCoef = nullptr;
    int *dFirst = nullptr, *dLen = nullptr;
    cudaError_t e = cudaMallocAsync(&dU, sizeof(double) * nknots_old[d], stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&dUb, sizeof(double) * nknots_new[d], stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&dCoef, sizeof(double) * n_new * (p + 1), stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&dFirst, sizeof(int) * n_new, stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&dLen, sizeof(int) * n_new, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dU, knots_old[d], sizeof(double) * nknots_old[d], cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dUb, knots_new[d], sizeof(double) * nknots_new[d], cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return cuda_fail(e, "refine_knots: operator scratch");
    oslo_operator_kernel<<<(n_new + 127) / 128, 128, 0, stream>>>(p, dU, nknots_old[d], dUb, nknots_new[d], dFirst, dLen, dCoef);
    count_launch();
    // the last changed direction writes cp_out; the others alternate between scratch and cp_out
    ++step;
    double* dst = ((changed - step) % 2 == 0) ? cp_out_dev : scratch_dev;
    igab_status st = refine_apply(dim, dw, ncp, d, n_new, p + 1, dFirst, dLen, dCoef, src, dst, stream);
    cudaFreeAsync(dU, stream);
    cudaFreeAsync(dUb, stream);
    cudaFreeAsync(dCoef, stream);
    cudaFreeAsync(dFirst, stream);
    cudaFreeAsync(dLen, stream);
    if (st) return st;
    ncp[d] = n_new;
    src = dst;
  }
  if (changed == 0) {
    long long n = 1;
    for (int d = 0; d < dim; ++d) n *= ncp[d];
    IGAB_CUDA_TRY(cudaMemcpyAsync(cp_out_dev, cp_in_dev, sizeof(double) * n * (dw + 1), cudaMemcpyDeviceToDevice, stream));
  }
  return IGAB_OK;
}

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

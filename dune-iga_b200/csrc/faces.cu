// faces.cu -- boundary / intersection evaluation: position, unit outer normal and surface integration
// element at quadrature points of element faces.  Replaces, batched over faces x points,
//   NURBSintersection::geometry().global / integrationElement, outerNormal, unitOuterNormal, integrationOuterNormal
//   (dune/iga/nurbsintersection.hh:86-160) and the sub-entity geometry with one fixed direction
//   (dune/iga/nurbsgeometry.hh:294-303,365-378).
// Face numbering is DUNE's cube numbering used there: face = 2*d + s fixes xi_d = s.
// Implementation: (1) expand (face, rule point) to full element-local coordinates, (2) evaluate x and J^T with the
// point-list evaluation kernels, (3) normals / dS from the rows of J^T exactly as the reference's switch statements.
#include "igab_internal.cuh"

namespace igab {

namespace {

__global__ void face_points_kernel(int dim, long long nfaces, int nqf, int nq0, const int* __restrict__ elem,
                                   const int* __restrict__ face, const double* __restrict__ xa,
                                   const double* __restrict__ xb, int* __restrict__ elemOut, double* __restrict__ xiOut) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nfaces * nqf) return;
  const long long f = t / nqf;
  const int q = (int)(t % nqf);
  const int fd = face[f] >> 1, side = face[f] & 1;
  elemOut[t] = elem[f];
  // free directions in increasing order take the rule's coordinates (first free direction fastest)
  int k = 0;
  for (int d = 0; d < dim; ++d) {
    double v;
    if (d == fd)
      v = side;
    else {
      v = (k == 0) ? xa[q % nq0] : xb[q / nq0];
      ++k;
    }
    xiOut[t * dim + d] = v;
  }
}

template <int DIM, int DW>
__global__ void face_normals_kernel(long long npts, int nqf, const int* __restrict__ face, const double* __restrict__ Jt,
                                    double* __restrict__ normal, double* __restrict__ dS) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npts) return;
  const int fi = face[t / nqf];
  const int fd = fi >> 1;
  double J[DIM][DW];
  for (int d = 0; d < DIM; ++d)
    for (int c = 0; c < DW; ++c) J[d][c] = Jt[(t * DIM + d) * DW + c];
  double n[DW], ds;
  if constexpr (DIM == 2 && DW == 2) {
    const double* T = J[1 - fd];  // tangent of the edge: the free direction
    if (fi == 0 || fi == 3) { n[0] = -T[1]; n[1] = T[0]; } else { n[0] = T[1]; n[1] = -T[0]; }
    ds = sqrt(T[0] * T[0] + T[1] * T[1]);
  } else if constexpr (DIM == 2 && DW == 3) {
    const double N[3] = {J[0][1] * J[1][2] - J[0][2] * J[1][1], J[0][2] * J[1][0] - J[0][0] * J[1][2],
                         J[0][0] * J[1][1] - J[0][1] * J[1][0]};
    const double* a;
    const double* b;
    // face 0: N x J1 ; 1: J1 x N ; 2: J0 x N ; 3: N x J0
    if (fi == 0) { a = N; b = J[1]; } else if (fi == 1) { a = J[1]; b = N; } else if (fi == 2) { a = J[0]; b = N; } else { a = N; b = J[0]; }
    n[0] = a[1] * b[2] - a[2] * b[1];
    n[1] = a[2] * b[0] - a[0] * b[2];
    n[2] = a[0] * b[1] - a[1] * b[0];
    const double* T = J[1 - fd];
    ds = sqrt(T[0] * T[0] + T[1] * T[1] + T[2] * T[2]);
  } else {
    // 3-D: the face's two tangents are the free rows in increasing order
    const int d0 = fd == 0 ? 1 : 0, d1 = fd == 2 ? 1 : 2;
    const double* T0 = J[d0];
    const double* T1 = J[d1];
    double c[3] = {T0[1] * T1[2] - T0[2] * T1[1], T0[2] * T1[0] - T0[0] * T1[2], T0[0] * T1[1] - T0[1] * T1[0]};
    const double sgn = (fi == 0 || fi == 3 || fi == 4) ? -1.0 : 1.0;  // cross(T1,T0) for 0,3,4 ; cross(T0,T1) else
    n[0] = sgn * c[0]; n[1] = sgn * c[1]; n[2] = sgn * c[2];
    ds = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);  // sqrtDetAAT of the 2x3 face Jacobian
  }
  double len = 0;
  for (int c = 0; c < DW; ++c) len += n[c] * n[c];
  len = sqrt(len);
  if (normal)
    for (int c = 0; c < DW; ++c) normal[t * DW + c] = n[c] / len;
  if (dS) dS[t] = ds;
}

}  // namespace

igab_status eval_faces_device(igab_patch* p, long long nfaces, const int* elem, const int* face, const int* nqf1,
                              const double* const* xi_dev, double* x, double* normal, double* dS, cudaStream_t stream) {
  const PatchDev& D = p->dev;
  if (D.dim < 2 || (D.dim == 3 && D.dw != 3)) {
    set_error("eval_faces: needs a 2-D or 3-D patch");
    return IGAB_UNSUPPORTED;
  }
  const int nq0 = nqf1[0], nq1 = D.dim == 3 ? nqf1[1] : 1, nqf = nq0 * nq1;
  const long long npts = nfaces * nqf;
  if (npts == 0) return IGAB_OK;
  int* d_el = nullptr;
  double *d_xi = nullptr, *d_Jt = nullptr;
  IGAB_CUDA_TRY(cudaMallocAsync(&d_el, sizeof(int) * npts, stream));
  IGAB_CUDA_TRY(cudaMallocAsync(&d_xi, sizeof(double) * npts * D.dim, stream));
  IGAB_CUDA_TRY(cudaMallocAsync(&d_Jt, sizeof(double) * npts * D.dim * D.dw, stream));
  const int th = 256;
  const unsigned blocks = (unsigned)((npts + th - 1) / th);
  face_points_kernel<<<blocks, th, 0, stream>>>(D.dim, nfaces, nqf, nq0, elem, face, xi_dev[0],
                                               D.dim == 3 ? xi_dev[1] : xi_dev[0], d_el, d_xi);
  count_launch();
  PointSource src{};
  src.tensor = 0;
  src.elem = d_el;
  src.xi = d_xi;
  EvalOutDev o{};
  o.x = x;
  o.Jt = d_Jt;
  igab_status st;
  if (p->kernel_mode != IGAB_KERNEL_GENERIC && eval_points_2d_supported(D))
    st = launch_eval_points_2d(D, src, npts, 1, o, stream);
  else
  {
    st = launch_eval_warp(D, src, npts, 1, o, stream);  // warp per 32 face points
    if (st == IGAB_UNSUPPORTED) st = launch_eval_generic(D, src, npts, 1, o, stream);
  }
  if (!st && (normal || dS)) {
    if (D.dim == 2 && D.dw == 2)
      face_normals_kernel<2, 2><<<blocks, th, 0, stream>>>(npts, nqf, face, d_Jt, normal, dS);
    else if (D.dim == 2)
      face_normals_kernel<2, 3><<<blocks, th, 0, stream>>>(npts, nqf, face, d_Jt, normal, dS);
    else
      face_normals_kernel<3, 3><<<blocks, th, 0, stream>>>(npts, nqf, face, d_Jt, normal, dS);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) st = cuda_fail(e, "face normals");
  }
  cudaFreeAsync(d_el, stream);
  cudaFreeAsync(d_xi, stream);
  cudaFreeAsync(d_Jt, stream);
  return st;
}

}  // namespace igab

// csr.cu -- global sparse matrix from element matrices: CSR pattern of the tensor-product NURBS basis and
// a deterministic gather assembly.  Functional spec: the scatter loop of dune/python/test/poisson.py:113-125 with the
// DOF numbering of NurbsPreBasis::createOriginalNodeIndices (dune/iga/nurbsbasis.hh:572-584).
//
// Row g = (g_0, g_1, g_2) couples with h iff |g_d - h_d| <= p_d in every direction, so a row's columns are the box
// [max(0, g_d - p_d), min(n_d - 1, g_d + p_d)]: row length and the position of a column inside the row are closed-form.
// One warp per row, lanes over the row's non-zeros; every non-zero gathers the contributions Ke[e][i][j] of the (at
// most prod (p_d + 1)) elements whose local bases contain both DOFs, in a fixed order -- no atomics.
//
// Trimmed patches (NurbsPreBasis::prepareForTrim, nurbsbasis.hh:599-615): rows and columns are the compacted DOFs (those
// touched by a non-empty element, renumbered ascending); a row keeps exactly the columns that share a non-empty element
// with it, which is the pattern the scatter loop of poisson.py would produce.  Row pointer and column indices are built
// once per handle (thread per row) and cached on the device; gather, load vector and Dirichlet rows read them.
#include <algorithm>
#include <cstdlib>

#include "igab_internal.cuh"

namespace igab {

namespace {

struct CsrGeom {
  int dim, ncomp;
  int p[kMaxDim], n[kMaxDim], nel[kMaxDim];
  const int* elemOfFirst[kMaxDim];  // first-DOF index f_d -> element e_d (or -1): inverse of span - degree
  int nb;
  // trimmed patches only (nullptr otherwise)
  const uint8_t* trim;    // element flags
  const int32_t* map;     // original scalar DOF -> compacted (or -1)
  const int32_t* inv;     // compacted scalar DOF -> original
  const int32_t* colidx;  // cached column indices of the compacted pattern
};

__device__ __forceinline__ void dof_box(const CsrGeom& G, long long g, int* gd, int* lo, int* len);

// element (e0, e1, e2) of the first-DOF triple, -1 if there is none or (trimmed patches) it is empty
__device__ __forceinline__ long long live_element(const CsrGeom& G, int f0, int f1, int f2) {
  const int e0 = G.elemOfFirst[0][f0];
  const int e1 = G.dim > 1 ? G.elemOfFirst[1][f1] : 0;
  const int e2 = G.dim > 2 ? G.elemOfFirst[2][f2] : 0;
  if (e0 < 0 || e1 < 0 || e2 < 0) return -1;
  const long long e = e0 + (long long)G.nel[0] * (e1 + (long long)(G.dim > 1 ? G.nel[1] : 1) * e2);
  if (G.trim && G.trim[e] == IGAB_TRIM_EMPTY) return -1;
  return e;
}

// do the DOFs g and h share a live element?
__device__ __forceinline__ bool coupled(const CsrGeom& G, const int* gd, const int* hd) {
  int fa[kMaxDim], fb[kMaxDim];
  for (int q = 0; q < kMaxDim; ++q) {
    fa[q] = q < G.dim ? max(max(gd[q], hd[q]) - G.p[q], 0) : 0;
    fb[q] = q < G.dim ? min(min(gd[q], hd[q]), G.n[q] - 1 - G.p[q]) : 0;
  }
  for (int f2 = fa[2]; f2 <= fb[2]; ++f2)
    for (int f1 = fa[1]; f1 <= fb[1]; ++f1)
      for (int f0 = fa[0]; f0 <= fb[0]; ++f0)
        if (live_element(G, f0, f1, f2) >= 0) return true;
  return false;
}

__device__ __forceinline__ void dof_box(const CsrGeom& G, long long g, int* gd, int* lo, int* len) {
  for (int d = 0; d < G.dim; ++d) {
    gd[d] = (int)(g % G.n[d]);
    g /= G.n[d];
    lo[d] = max(0, gd[d] - G.p[d]);
    const int hi = min(G.n[d] - 1, gd[d] + G.p[d]);
    len[d] = hi - lo[d] + 1;
  }
  for (int d = G.dim; d < kMaxDim; ++d) { gd[d] = 0; lo[d] = 0; len[d] = 1; }
}

// scalar DOF index in the matrix numbering (compacted for trimmed patches) -> box of the original DOF
__device__ __forceinline__ void row_box(const CsrGeom& G, long long r, int* gd, int* lo, int* len) {
  dof_box(G, G.inv ? (long long)G.inv[r] : r, gd, lo, len);
}

// trimmed patches: row lengths and column indices of the compacted pattern, one thread per compacted scalar DOF
__global__ void csr_rowlen_trim_kernel(CsrGeom G, long long ndofc, long long* __restrict__ rowlen) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= ndofc) return;
  int gd[kMaxDim], lo[kMaxDim], len[kMaxDim], hd[kMaxDim];
  row_box(G, r, gd, lo, len);
  long long cnt = 0;
  for (int k2 = 0; k2 < len[2]; ++k2)
    for (int k1 = 0; k1 < len[1]; ++k1)
      for (int k0 = 0; k0 < len[0]; ++k0) {
        hd[0] = lo[0] + k0; hd[1] = lo[1] + k1; hd[2] = lo[2] + k2;
        cnt += coupled(G, gd, hd) ? 1 : 0;
      }
  for (int c = 0; c < G.ncomp; ++c) rowlen[r * G.ncomp + c] = cnt * G.ncomp;
}
__global__ void csr_colidx_trim_kernel(CsrGeom G, long long ndofc, const long long* __restrict__ rowptr,
                                       int32_t* __restrict__ colidx) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= ndofc) return;
  int gd[kMaxDim], lo[kMaxDim], len[kMaxDim], hd[kMaxDim];
  row_box(G, r, gd, lo, len);
  long long cnt = 0;
  for (int k2 = 0; k2 < len[2]; ++k2)
    for (int k1 = 0; k1 < len[1]; ++k1)
      for (int k0 = 0; k0 < len[0]; ++k0) {
        hd[0] = lo[0] + k0; hd[1] = lo[1] + k1; hd[2] = lo[2] + k2;
        if (!coupled(G, gd, hd)) continue;
        const long long h = hd[0] + (long long)G.n[0] * (hd[1] + (long long)(G.dim > 1 ? G.n[1] : 1) * hd[2]);
        const long long hc = G.map[h];  // >= 0: h shares a live element with g
        for (int c = 0; c < G.ncomp; ++c)
          for (int d = 0; d < G.ncomp; ++d) colidx[rowptr[r * G.ncomp + c] + cnt * G.ncomp + d] = (int32_t)(hc * G.ncomp + d);
        ++cnt;
      }
}
__global__ void dof_inverse_kernel(long long n, const int32_t* __restrict__ map, int32_t* __restrict__ inv) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n && map[g] >= 0) inv[map[g]] = (int32_t)g;
}

// row lengths (one thread per scalar DOF); the host turns them into rowptr with a prefix sum
__global__ void csr_rowlen_kernel(CsrGeom G, long long ndof, long long* __restrict__ rowlen) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ndof) return;
  int gd[kMaxDim], lo[kMaxDim], len[kMaxDim];
  row_box(G, g, gd, lo, len);
  const long long l = (long long)len[0] * len[1] * len[2] * G.ncomp;
  for (int c = 0; c < G.ncomp; ++c) rowlen[g * G.ncomp + c] = l;
}

__global__ void csr_colidx_kernel(CsrGeom G, long long nrows, const long long* __restrict__ rowptr,
                                  int32_t* __restrict__ colidx) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  int gd[kMaxDim], lo[kMaxDim], len[kMaxDim];
  row_box(G, row / G.ncomp, gd, lo, len);
  const long long base = rowptr[row];
  const int nn = len[0] * len[1] * len[2] * G.ncomp;
  for (int k = lane; k < nn; k += 32) {
    const int d = k % G.ncomp;
    int kk = k / G.ncomp;
    const int c0 = lo[0] + kk % len[0];
    kk /= len[0];
    const int c1 = lo[1] + kk % len[1];
    const int c2 = lo[2] + kk / len[1];
    const long long h = c0 + (long long)G.n[0] * (c1 + (long long)(G.dim > 1 ? G.n[1] : 1) * c2);
    colidx[base + k] = (int32_t)(h * G.ncomp + d);
  }
}

__global__ void csr_gather_kernel(CsrGeom G, long long row_begin, long long nrows, long long elem_begin,
                                  long long elem_end, const double* __restrict__ Ke,
                                  const long long* __restrict__ rowptr, double* __restrict__ values, int accumulate) {
  const long long row = row_begin + (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  int gd[kMaxDim], lo[kMaxDim], len[kMaxDim];
  row_box(G, row / G.ncomp, gd, lo, len);
  const int c = (int)(row % G.ncomp);
  const long long base = rowptr[row];
  const int nn = (int)(rowptr[row + 1] - base);
  const long long n = (long long)G.nb * G.ncomp;
  const int o0 = G.p[0] + 1, o1 = G.dim > 1 ? G.p[1] + 1 : 1;
  for (int k = lane; k < nn; k += 32) {
    int d, hd[kMaxDim];
    if (G.colidx) {  // trimmed: the cached compacted column -> original DOF
      const int col = G.colidx[base + k];
      d = col % G.ncomp;
      long long h = G.inv[col / G.ncomp];
      hd[0] = (int)(h % G.n[0]);
      h /= G.n[0];
      hd[1] = G.dim > 1 ? (int)(h % G.n[1]) : 0;
      hd[2] = G.dim > 1 ? (int)(h / G.n[1]) : 0;
    } else {
      d = k % G.ncomp;
      int kk = k / G.ncomp;
      hd[0] = lo[0] + kk % len[0];
      kk /= len[0];
      hd[1] = lo[1] + kk % len[1];
      hd[2] = lo[2] + kk / len[1];
    }
    int fa[kMaxDim], fb[kMaxDim];  // range of first-DOF indices of the elements that contain both g and h
    for (int q = 0; q < kMaxDim; ++q) {
      if (q < G.dim) {
        fa[q] = max(max(gd[q], hd[q]) - G.p[q], 0);
        fb[q] = min(min(gd[q], hd[q]), G.n[q] - 1 - G.p[q]);
      } else {
        fa[q] = fb[q] = 0;
      }
    }
    double acc = 0.0;
    const int nel1 = G.dim > 1 ? G.nel[1] : 1;
    const double* __restrict__ Kc = Ke + (long long)c * n + d;  // component (c, d) of every (i, j) block
    for (int f2 = fa[2]; f2 <= fb[2]; ++f2) {
      const int e2 = G.dim > 2 ? G.elemOfFirst[2][f2] : 0;
      if (e2 < 0) continue;
      for (int f1 = fa[1]; f1 <= fb[1]; ++f1) {
        const int e1 = G.dim > 1 ? G.elemOfFirst[1][f1] : 0;
        if (e1 < 0) continue;
        const long long erow = (long long)G.nel[0] * (e1 + (long long)nel1 * e2);
        // local indices of g and h in an element whose first DOF is (f0, f1, f2): i = i12 + (gd0 - f0), same for j
        const int i12 = o0 * ((gd[1] - f1) + o1 * (gd[2] - f2)) + gd[0];
        const int j12 = o0 * ((hd[1] - f1) + o1 * (hd[2] - f2)) + hd[0];
        for (int f0 = fa[0]; f0 <= fb[0]; ++f0) {
          const int e0 = G.elemOfFirst[0][f0];
          if (e0 < 0) continue;
          const long long e = erow + e0;
          if (e < elem_begin || e >= elem_end) continue;
          if (G.trim && G.trim[e] == IGAB_TRIM_EMPTY) continue;
          acc += Kc[(e - elem_begin) * n * n + ((long long)(i12 - f0) * n + (j12 - f0)) * G.ncomp];
        }
      }
    }
    if (accumulate)
      values[base + k] += acc;
    else
      values[base + k] = acc;
  }
}

// The same gather organised by ROWS OF THE ELEMENT MATRICES (untrimmed numbering): one warp per matrix row (g, c).  For every
// element that contains g (fixed order) the warp reads the element matrix row of g -- n contiguous doubles, coalesced --
// and adds entry (j, d) to the row's accumulator in shared memory at the closed-form position of column h = f + j inside
// the row's box.  Every K_e entry is read once, in 1000-byte runs (the per-non-zero gather above reads it as scattered
// 8-byte words); one lane owns one column per element, so the sums have a fixed order: bitwise reproducible.
__global__ void __launch_bounds__(128) csr_gather_rows_kernel(CsrGeom G, long long row_begin, long long nrows,
                                                              long long elem_begin, long long elem_end,
                                                              const double* __restrict__ Ke,
                                                              const long long* __restrict__ rowptr,
                                                              double* __restrict__ values, int accumulate, int accStride) {
  extern __shared__ double sacc[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row = row_begin + (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= nrows) return;
  double* acc = sacc + (size_t)warp * accStride;
  int gd[kMaxDim], lo[kMaxDim], len[kMaxDim];
  dof_box(G, row / G.ncomp, gd, lo, len);
  const int c = (int)(row % G.ncomp);
  const long long base = rowptr[row];
  const int nn = (int)(rowptr[row + 1] - base);
  for (int k = lane; k < nn; k += 32) acc[k] = 0.0;
  const int n = G.nb * G.ncomp;
  const int o0 = G.p[0] + 1, o1 = G.dim > 1 ? G.p[1] + 1 : 1;
  const int nel1 = G.dim > 1 ? G.nel[1] : 1;
  int fa[kMaxDim], fb[kMaxDim];
  for (int q = 0; q < kMaxDim; ++q) {
    fa[q] = q < G.dim ? max(gd[q] - G.p[q], 0) : 0;
    fb[q] = q < G.dim ? min(gd[q], G.n[q] - 1 - G.p[q]) : 0;
  }
  __syncwarp();
  // Candidate elements in a fixed order (f0 fastest), software-pipelined: the first PF x 32 entries of the NEXT element's
  // matrix row are in flight while the current one is added into the accumulator -- a warp otherwise has only one
  // 1000-byte row outstanding and the gather runs at a quarter of the HBM bandwidth.  The position of entry jj inside the
  // row's box does not depend on the element: computed once per lane.
  constexpr int PF = 4;
  int pos[PF];
#pragma unroll
  for (int t = 0; t < PF; ++t) {
    const int jj = lane + 32 * t;
    const int j = jj / G.ncomp, d = jj - j * G.ncomp;
    const int j0 = j % o0, r = j / o0, j1 = r % o1, j2 = r / o1;
    pos[t] = ((j2 * len[1] + j1) * len[0] + j0) * G.ncomp + d;
  }
  const int nc0 = fb[0] - fa[0] + 1, nc1 = fb[1] - fa[1] + 1, nc2 = fb[2] - fa[2] + 1;
  const int total = nc0 * nc1 * nc2;
  // The candidates' element indices per direction sit in lanes 0 .. nc_d - 1 and are fetched with a shuffle; the candidate
  // counter runs as three digits (c0 fastest).  (Decoding a linear index with run-time divisions and three dependent
  // global loads per candidate made the kernel issue-bound: 67 % of the issue slots, a third of them this arithmetic.)
  const int eo0 = lane < nc0 ? G.elemOfFirst[0][fa[0] + lane] : -1;
  const int eo1 = G.dim > 1 ? (lane < nc1 ? G.elemOfFirst[1][fa[1] + lane] : -1) : 0;
  const int eo2 = G.dim > 2 ? (lane < nc2 ? G.elemOfFirst[2][fa[2] + lane] : -1) : 0;
  int c0 = 0, c1 = 0, c2 = 0;  // digits of the candidate fetched next
  auto fetch = [&](double (&v)[PF], const double*& src, int& kb) -> bool {
    const int f0 = fa[0] + c0, f1 = fa[1] + c1, f2 = fa[2] + c2;
    const int e0 = __shfl_sync(0xffffffffu, eo0, c0);
    const int e1 = G.dim > 1 ? __shfl_sync(0xffffffffu, eo1, c1) : 0;
    const int e2 = G.dim > 2 ? __shfl_sync(0xffffffffu, eo2, c2) : 0;
    if (++c0 == nc0) {
      c0 = 0;
      if (++c1 == nc1) {
        c1 = 0;
        ++c2;
      }
    }
    if (e0 < 0 || e1 < 0 || e2 < 0) return false;
    const long long e = e0 + (long long)G.nel[0] * (e1 + (long long)nel1 * e2);
    if (e < elem_begin || e >= elem_end) return false;
    const int i = (gd[0] - f0) + o0 * ((gd[1] - f1) + o1 * (gd[2] - f2));
    src = Ke + (e - elem_begin) * (long long)n * n + (long long)(i * G.ncomp + c) * n;
    // column h = f + j sits at ((h2 - lo2) len1 + (h1 - lo1)) len0 + (h0 - lo0) of the row's box
    kb = (((f2 - lo[2]) * len[1] + (f1 - lo[1])) * len[0] + (f0 - lo[0])) * G.ncomp;
#pragma unroll
    for (int t = 0; t < PF; ++t) v[t] = (lane + 32 * t < n) ? __ldg(src + lane + 32 * t) : 0.0;
    return true;
  };
  // (two candidates in flight instead of one: 94 registers instead of 80, one resident block less, 0.41 -> 0.46 ms; not kept)
  double cur[PF], nxt[PF];
  const double *srcC = nullptr, *srcN = nullptr;
  int kbC = 0, kbN = 0;
  bool okC = total > 0 && fetch(cur, srcC, kbC), okN = false;
  for (int cidx = 0; cidx < total; ++cidx) {
    okN = (cidx + 1 < total) && fetch(nxt, srcN, kbN);
    if (okC) {
#pragma unroll
      for (int t = 0; t < PF; ++t)
        if (lane + 32 * t < n) acc[kbC + pos[t]] += cur[t];
      for (int jj = lane + 32 * PF; jj < n; jj += 32) {  // rows longer than PF x 32 entries (vector-valued, p >= 5)
        const int j = jj / G.ncomp, d = jj - j * G.ncomp;
        const int j0 = j % o0, r = j / o0, j1 = r % o1, j2 = r / o1;
        acc[kbC + ((j2 * len[1] + j1) * len[0] + j0) * G.ncomp + d] += srcC[jj];
      }
    }
    __syncwarp();  // the next element's columns map to other lanes: order the read-modify-writes of acc
    okC = okN;
    srcC = srcN;
    kbC = kbN;
#pragma unroll
    for (int t = 0; t < PF; ++t) cur[t] = nxt[t];
  }
  __syncwarp();
  for (int k = lane; k < nn; k += 32) {
    if (accumulate)
      values[base + k] += acc[k];
    else
      values[base + k] = acc[k];
  }
}

// global load vector: b[g*ncomp + c] (+)= sum over the elements of [elem_begin, elem_end) that contain DOF g of
// fe[e][i*ncomp + c] (dune/python/test/poisson.py:121, b[gi] += localb[i]); one thread per row, fixed element order
__global__ void rhs_gather_kernel(CsrGeom G, long long nrows, long long elem_begin, long long elem_end,
                                  const double* __restrict__ fe, double* __restrict__ b, int accumulate) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  int gd[kMaxDim], lo[kMaxDim], len[kMaxDim];
  row_box(G, row / G.ncomp, gd, lo, len);
  const int c = (int)(row % G.ncomp);
  const long long n = (long long)G.nb * G.ncomp;
  const int o0 = G.p[0] + 1, o1 = G.dim > 1 ? G.p[1] + 1 : 1;
  int fa[kMaxDim], fb[kMaxDim];
  for (int q = 0; q < kMaxDim; ++q) {
    fa[q] = q < G.dim ? max(gd[q] - G.p[q], 0) : 0;
    fb[q] = q < G.dim ? min(gd[q], G.n[q] - 1 - G.p[q]) : 0;
  }
  double acc = 0.0;
  for (int f2 = fa[2]; f2 <= fb[2]; ++f2)
    for (int f1 = fa[1]; f1 <= fb[1]; ++f1)
      for (int f0 = fa[0]; f0 <= fb[0]; ++f0) {
        const long long e = live_element(G, f0, f1, f2);
        if (e < elem_begin || e >= elem_end) continue;
        const int i = (gd[0] - f0) + o0 * ((gd[1] - f1) + o1 * (gd[2] - f2));
        acc += fe[(e - elem_begin) * n + i * G.ncomp + c];
      }
  b[row] = accumulate ? b[row] + acc : acc;
}

// Dirichlet rows as the reference's assembler sets them (poisson.py:117-119): row gi of a marked DOF becomes the identity
// row (columns are left alone), b[gi] = g[gi].  One warp per row; the column of an entry is closed-form.
__global__ void dirichlet_rows_kernel(CsrGeom G, long long nrows, const long long* __restrict__ rowptr,
                                      const uint8_t* __restrict__ mask, const double* __restrict__ gval,
                                      double* __restrict__ values, double* __restrict__ b) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nrows || !mask[row]) return;
  int gd[kMaxDim], lo[kMaxDim], len[kMaxDim];
  row_box(G, row / G.ncomp, gd, lo, len);
  const long long base = rowptr[row];
  const int nn = (int)(rowptr[row + 1] - base);
  for (int k = lane; k < nn; k += 32) {
    long long col;
    if (G.colidx) {
      col = G.colidx[base + k];
    } else {
      const int d = k % G.ncomp;
      int kk = k / G.ncomp;
      const int c0 = lo[0] + kk % len[0];
      kk /= len[0];
      const int c1 = lo[1] + kk % len[1];
      const int c2 = lo[2] + kk / len[1];
      col = (c0 + (long long)G.n[0] * (c1 + (long long)(G.dim > 1 ? G.n[1] : 1) * c2)) * G.ncomp + d;
    }
    values[base + k] = (col == row) ? 1.0 : 0.0;
  }
  if (lane == 0 && b) b[row] = gval ? gval[row] : 0.0;
}

igab_status make_geom(igab_patch* p, int ncomp, CsrGeom* G) {
  if (ncomp < 1 || ncomp > 3) {
    set_error("csr: ncomp must be 1..3");
    return IGAB_INVALID_ARGUMENT;
  }
  const PatchDev& D = p->dev;
  if (igab_status cst = ensure_dof_compaction(p)) return cst;
  for (int d = 0; d < D.dim; ++d) {
    if (!p->d_elem_of_first[d]) {
      const int nf = D.ncp[d] - D.degree[d];  // possible first-DOF indices 0 .. n_d - p_d - 1
      std::vector<int> tab(nf, -1);
      for (int e = 0; e < D.nel[d]; ++e) tab[p->hspan[d][e] - D.degree[d]] = e;
      IGAB_CUDA_TRY(cudaMalloc(&p->d_elem_of_first[d], sizeof(int) * nf));
      IGAB_CUDA_TRY(cudaMemcpy(p->d_elem_of_first[d], tab.data(), sizeof(int) * nf, cudaMemcpyHostToDevice));
    }
  }
  G->dim = D.dim;
  G->ncomp = ncomp;
  G->nb = D.nb;
  for (int d = 0; d < kMaxDim; ++d) {
    G->p[d] = d < D.dim ? D.degree[d] : 0;
    G->n[d] = d < D.dim ? D.ncp[d] : 1;
    G->nel[d] = d < D.dim ? D.nel[d] : 1;
    G->elemOfFirst[d] = d < D.dim ? p->d_elem_of_first[d] : nullptr;
  }
  G->trim = p->d_trim;
  G->map = p->d_dofc_map;
  G->inv = p->d_dofc_inv;
  G->colidx = p->d_csr_colidx;
  return IGAB_OK;
}

}  // namespace

igab_status ensure_dof_compaction(igab_patch* p) {
  const PatchDev& D = p->dev;
  if (p->d_trim && !p->d_dofc_map) {
    // compacted numbering (prepareForTrim): the same mark / scan kernels as igab_patch_dofmap, kept in the handle
    const long long n = D.nelem * D.nb, nd = p->ndof_untrimmed;
    int32_t *dof = nullptr, *used = nullptr, *map = nullptr, *inv = nullptr;
    long long* total_dev = nullptr;
    long long total = 0;
    cudaError_t e = cudaMalloc(&dof, sizeof(int32_t) * n);
    if (e == cudaSuccess) e = cudaMalloc(&used, sizeof(int32_t) * nd);
    if (e == cudaSuccess) e = cudaMalloc(&map, sizeof(int32_t) * nd);
    if (e == cudaSuccess) e = cudaMalloc(&total_dev, sizeof(long long));
    igab_status st = e == cudaSuccess ? IGAB_OK : cuda_fail(e, "csr: compacted DOF scratch");
    if (!st) st = launch_dofmap(D, dof, nullptr);
    if (!st) st = launch_dof_compact(D, p->d_trim, dof, used, map, nd, total_dev, nullptr);
    if (!st) {
      e = cudaMemcpy(&total, total_dev, sizeof(long long), cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) e = cudaMalloc(&inv, sizeof(int32_t) * std::max(1LL, total));
      if (e == cudaSuccess) {
        dof_inverse_kernel<<<(unsigned)((nd + 255) / 256), 256>>>(nd, map, inv);
        count_launch();
        e = cudaDeviceSynchronize();
      }
      if (e != cudaSuccess) st = cuda_fail(e, "csr: compacted DOF numbering");
    }
    if (dof) cudaFree(dof);
    if (used) cudaFree(used);
    if (total_dev) cudaFree(total_dev);
    if (st) {
      if (map) cudaFree(map);
      if (inv) cudaFree(inv);
      return st;
    }
    p->d_dofc_map = map;
    p->d_dofc_inv = inv;
    p->ndof_compact = total;
  }
  return IGAB_OK;
}

igab_status csr_pattern(igab_patch* p, int ncomp, int64_t* nrows_out, int64_t* nnz_out, int64_t* rowptr, int32_t* colidx,
                        igab_memspace space, cudaStream_t stream) {
  CsrGeom G;
  igab_status st = make_geom(p, ncomp, &G);
  if (st) return st;
  const long long ndof = p->d_trim ? p->ndof_compact : p->ndof_untrimmed, nrows = ndof * ncomp;
  // row pointer: row lengths (closed form, or counted per row on trimmed patches), host prefix sum (cached in the handle)
  if ((long long)p->csr_rowptr_host.size() != nrows + 1 || p->csr_ncomp != ncomp) {
    long long* d_len = nullptr;
    IGAB_CUDA_TRY(cudaMalloc(&d_len, sizeof(long long) * std::max(1LL, nrows)));
    if (p->d_trim)
      csr_rowlen_trim_kernel<<<(unsigned)((ndof + 127) / 128), 128, 0, stream>>>(G, ndof, d_len);
    else
      csr_rowlen_kernel<<<(unsigned)((ndof + 255) / 256), 256, 0, stream>>>(G, ndof, d_len);
    count_launch();
    std::vector<long long> len(nrows);
    cudaError_t e = cudaMemcpyAsync(len.data(), d_len, sizeof(long long) * nrows, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d_len);
    if (e != cudaSuccess) return cuda_fail(e, "csr row lengths");
    p->csr_rowptr_host.assign(nrows + 1, 0);
    for (long long r = 0; r < nrows; ++r) p->csr_rowptr_host[r + 1] = p->csr_rowptr_host[r] + len[r];
    p->csr_ncomp = ncomp;
    if (p->d_csr_rowptr) cudaFree(p->d_csr_rowptr);
    p->d_csr_rowptr = nullptr;
    if (p->d_csr_colidx) cudaFree(p->d_csr_colidx);
    p->d_csr_colidx = nullptr;
    G.colidx = nullptr;
    IGAB_CUDA_TRY(cudaMalloc(&p->d_csr_rowptr, sizeof(long long) * (nrows + 1)));
    IGAB_CUDA_TRY(cudaMemcpy(p->d_csr_rowptr, p->csr_rowptr_host.data(), sizeof(long long) * (nrows + 1), cudaMemcpyHostToDevice));
    if (p->d_trim) {  // the compacted pattern has no closed form: build its column indices once and keep them
      const long long nz = p->csr_rowptr_host[nrows];
      IGAB_CUDA_TRY(cudaMalloc(&p->d_csr_colidx, sizeof(int32_t) * std::max(1LL, nz)));
      csr_colidx_trim_kernel<<<(unsigned)((ndof + 127) / 128), 128, 0, stream>>>(G, ndof, p->d_csr_rowptr, p->d_csr_colidx);
      count_launch();
      IGAB_CUDA_TRY(cudaGetLastError());
    }
  }
  const long long nnz = p->csr_rowptr_host[nrows];
  if (nrows_out) *nrows_out = nrows;
  if (nnz_out) *nnz_out = nnz;
  if (nnz > 0x7fffffffLL * 64LL) {
    set_error("csr: pattern too large");
    return IGAB_INVALID_ARGUMENT;
  }
  if (rowptr) {
    if (space == IGAB_HOST)
      std::copy(p->csr_rowptr_host.begin(), p->csr_rowptr_host.end(), rowptr);
    else
      IGAB_CUDA_TRY(cudaMemcpyAsync(rowptr, p->d_csr_rowptr, sizeof(long long) * (nrows + 1), cudaMemcpyDeviceToDevice, stream));
  }
  if (colidx) {
    int32_t* d_col = colidx;
    if (space == IGAB_HOST) IGAB_CUDA_TRY(cudaMalloc(&d_col, sizeof(int32_t) * nnz));
    cudaError_t e = cudaSuccess;
    if (p->d_trim) {
      e = cudaMemcpyAsync(d_col, p->d_csr_colidx, sizeof(int32_t) * nnz, cudaMemcpyDeviceToDevice, stream);
    } else {
      const long long threads = nrows * 32;
      csr_colidx_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(G, nrows, p->d_csr_rowptr, d_col);
      count_launch();
      e = cudaGetLastError();
    }
    if (e == cudaSuccess && space == IGAB_HOST) {
      e = cudaMemcpyAsync(colidx, d_col, sizeof(int32_t) * nnz, cudaMemcpyDeviceToHost, stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    }
    if (space == IGAB_HOST) cudaFree(d_col);
    if (e != cudaSuccess) return cuda_fail(e, "csr column indices");
  }
  return IGAB_OK;
}

igab_status csr_assemble(igab_patch* p, int ncomp, long long eb, long long ee, const double* Ke, const int64_t* rowptr,
                         double* values, int accumulate, igab_memspace space, cudaStream_t stream) {
  CsrGeom G;
  igab_status st = make_geom(p, ncomp, &G);
  if (st) return st;
  int64_t nrows = 0, nnz = 0;
  if ((st = csr_pattern(p, ncomp, &nrows, &nnz, nullptr, nullptr, IGAB_HOST, stream))) return st;
  G.colidx = p->d_csr_colidx;
  (void)rowptr;  // the handle's own copy is used; the argument documents which pattern `values` belongs to
  const long long n = (long long)p->dev.nb * ncomp, nel = ee - eb;
  const double* dK = Ke;
  double* dV = values;
  if (space == IGAB_HOST) {
    double* tK = nullptr;
    IGAB_CUDA_TRY(cudaMalloc(&tK, sizeof(double) * std::max(1LL, nel * n * n)));
    IGAB_CUDA_TRY(cudaMalloc(&dV, sizeof(double) * nnz));
    cudaError_t e = cudaMemcpyAsync(tK, Ke, sizeof(double) * nel * n * n, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess && accumulate) e = cudaMemcpyAsync(dV, values, sizeof(double) * nnz, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) {
      cudaFree(tK);
      cudaFree(dV);
      return cuda_fail(e, "csr staging");
    }
    dK = tK;
  }
  // only the rows that elements [eb, ee) can touch: the DOF layers g_last in [span(e_min) - p, span(e_max)] of the
  // slowest direction are one contiguous row range (untrimmed numbering); the other rows are zeroed when not accumulating
  long long rowBegin = 0, rowEnd = nrows;
  cudaError_t e = cudaSuccess;
  if (!p->d_trim && nel > 0) {
    const PatchDev& D = p->dev;
    const int dl = D.dim - 1;
    long long perLayer = 1, dofStride = 1;
    for (int d = 0; d < dl; ++d) {
      perLayer *= D.nel[d];
      dofStride *= D.ncp[d];
    }
    const int eLo = (int)(eb / perLayer), eHi = (int)((ee - 1) / perLayer);
    rowBegin = (long long)(p->hspan[dl][eLo] - D.degree[dl]) * dofStride * ncomp;
    rowEnd = (long long)(p->hspan[dl][eHi] + 1) * dofStride * ncomp;
    if (!accumulate && (rowBegin > 0 || rowEnd < nrows)) {
      const long long zb = p->csr_rowptr_host[rowBegin], ze = p->csr_rowptr_host[rowEnd];
      if (zb > 0) e = cudaMemsetAsync(dV, 0, sizeof(double) * zb, stream);
      if (e == cudaSuccess && ze < nnz) e = cudaMemsetAsync(dV + ze, 0, sizeof(double) * (nnz - ze), stream);
    }
  }
  if (e == cudaSuccess && rowEnd > rowBegin) {
    // untrimmed numbering: gather by rows of the element matrices (coalesced reads); trimmed: per non-zero
    long long maxRow = ncomp;
    for (int d = 0; d < p->dev.dim; ++d) maxRow *= std::min(2 * p->dev.degree[d] + 1, p->dev.ncp[d]);
    const int accStride = (int)((maxRow + 1) & ~1LL);
    const size_t smem = sizeof(double) * 4 * accStride;
    const char* env = getenv("IGAB_CSR_GATHER");
    if (!p->d_trim && smem <= 200 * 1024 && !(env && env[0] == '0')) {
      // per (kernel, device): raises the dynamic shared-memory limit once on every device a process drives
      KernelCfg kc;
      if (igab_status kst = kernel_config((const void*)csr_gather_rows_kernel, 128, 200 * 1024, &kc)) return kst;
      if (e == cudaSuccess) {
        const long long blocks = (rowEnd - rowBegin + 3) / 4;
        csr_gather_rows_kernel<<<(unsigned)blocks, 128, smem, stream>>>(G, rowBegin, rowEnd, eb, ee, dK, p->d_csr_rowptr, dV,
                                                                       accumulate, accStride);
      }
    } else {
      const long long threads = (rowEnd - rowBegin) * 32;
      csr_gather_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(G, rowBegin, rowEnd, eb, ee, dK,
                                                                              p->d_csr_rowptr, dV, accumulate);
    }
    count_launch();
    if (e == cudaSuccess) e = cudaGetLastError();
  }
  if (space == IGAB_HOST) {
    if (e == cudaSuccess) e = cudaMemcpyAsync(values, dV, sizeof(double) * nnz, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(const_cast<double*>(dK));
    cudaFree(dV);
  }
  if (e != cudaSuccess) return cuda_fail(e, "csr gather");
  return IGAB_OK;
}

igab_status rhs_assemble(igab_patch* p, int ncomp, long long eb, long long ee, const double* fe, double* b, int accumulate,
                         igab_memspace space, cudaStream_t stream) {
  CsrGeom G;
  igab_status st = make_geom(p, ncomp, &G);
  if (st) return st;
  const long long nrows = (p->d_trim ? p->ndof_compact : p->ndof_untrimmed) * ncomp, n = (long long)p->dev.nb * ncomp,
                  nel = ee - eb;
  const double* dF = fe;
  double* dB = b;
  double* tmp = nullptr;
  if (space == IGAB_HOST) {
    IGAB_CUDA_TRY(cudaMalloc(&tmp, sizeof(double) * (std::max(1LL, nel * n) + nrows)));
    dB = tmp + std::max(1LL, nel * n);
    cudaError_t e = cudaMemcpyAsync(tmp, fe, sizeof(double) * nel * n, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess && accumulate) e = cudaMemcpyAsync(dB, b, sizeof(double) * nrows, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) {
      cudaFree(tmp);
      return cuda_fail(e, "rhs staging");
    }
    dF = tmp;
  }
  rhs_gather_kernel<<<(unsigned)((nrows + 127) / 128), 128, 0, stream>>>(G, nrows, eb, ee, dF, dB, accumulate);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (space == IGAB_HOST) {
    if (e == cudaSuccess) e = cudaMemcpyAsync(b, dB, sizeof(double) * nrows, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(tmp);
  }
  if (e != cudaSuccess) return cuda_fail(e, "rhs gather");
  return IGAB_OK;
}

igab_status csr_dirichlet(igab_patch* p, int ncomp, const uint8_t* mask, const double* gval, double* values, double* b,
                          igab_memspace space, cudaStream_t stream) {
  CsrGeom G;
  igab_status st = make_geom(p, ncomp, &G);
  if (st) return st;
  int64_t nrows = 0, nnz = 0;
  if ((st = csr_pattern(p, ncomp, &nrows, &nnz, nullptr, nullptr, IGAB_HOST, stream))) return st;
  G.colidx = p->d_csr_colidx;
  const uint8_t* dM = mask;
  const double* dG = gval;
  double *dV = values, *dB = b;
  uint8_t* tM = nullptr;
  double* tmp = nullptr;
  if (space == IGAB_HOST) {
    IGAB_CUDA_TRY(cudaMalloc(&tM, (size_t)nrows));
    cudaError_t e = cudaMalloc(&tmp, sizeof(double) * (nnz + 2 * nrows));
    if (e == cudaSuccess) e = cudaMemcpyAsync(tM, mask, (size_t)nrows, cudaMemcpyHostToDevice, stream);
    dV = tmp;
    if (e == cudaSuccess) e = cudaMemcpyAsync(dV, values, sizeof(double) * nnz, cudaMemcpyHostToDevice, stream);
    if (b) {
      dB = tmp + nnz;
      if (e == cudaSuccess) e = cudaMemcpyAsync(dB, b, sizeof(double) * nrows, cudaMemcpyHostToDevice, stream);
    }
    if (gval) {
      double* t = tmp + nnz + nrows;
      if (e == cudaSuccess) e = cudaMemcpyAsync(t, gval, sizeof(double) * nrows, cudaMemcpyHostToDevice, stream);
      dG = t;
    }
    if (e != cudaSuccess) {
      cudaFree(tM);
      if (tmp) cudaFree(tmp);
      return cuda_fail(e, "dirichlet staging");
    }
    dM = tM;
  }
  const long long threads = (long long)nrows * 32;
  dirichlet_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(G, nrows, p->d_csr_rowptr, dM, dG, dV, dB);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (space == IGAB_HOST) {
    if (e == cudaSuccess) e = cudaMemcpyAsync(values, dV, sizeof(double) * nnz, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && b) e = cudaMemcpyAsync(b, dB, sizeof(double) * nrows, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(tM);
    cudaFree(tmp);
  }
  if (e != cudaSuccess) return cuda_fail(e, "dirichlet rows");
  return IGAB_OK;
}

}  // namespace igab

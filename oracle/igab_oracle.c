/* oracle/igab_oracle.c -- TEST INFRASTRUCTURE (the "port" oracle).  Plain-C CPU restatement of the algorithm of
 * dune-iga's hot path, written from the reference's behaviour, each function citing the reference file:line it
 * follows (paths relative to the reference root).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product (dune-iga_b200/csrc) never links or calls it.
 *
 * PINNING: checked in tests/test_oracle.py against (i) every literal of the reference's own unit tests that touches
 * this path (tests/golden/reference_literals.json, extracted from dune/iga/test/gridtests.cpp and
 * trimmedgridtests.cpp) and (ii) the reference's own headers compiled through oracle/ref_shim.cpp
 * (oracle/_ref/liboracle_ref.so).  Unpinned by the reference's tests, hence "parity unpinned" for them:
 * pointwise integrationElement / jacobianInverseTransposed (dune-geometry MatrixHelper, not vendored),
 * Gauss points (dune-geometry QuadratureRules), PowerPreBasis numbering (dune-functions) -- SURVEY.md 8c.
 *
 * Layout contract everywhere: first direction fastest (MultiDimensionNet, dune/iga/utils/mdnet.hh:329-345).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXP 15          /* max degree per direction */
#define MAXO (MAXP + 1)
#define MAXDIM 3
#define MAXDER 8         /* max derivative order per direction for igo_partial */

typedef struct {
  int dim, dimworld;
  int degree[MAXDIM];
  int nknots[MAXDIM];
  const double* knots[MAXDIM];
  int ncp[MAXDIM];
  const double* cp; /* [ncp_total][dimworld+1]  (x..,w), i-fastest: controlpoint.hh:15-38, nurbspatchdata.hh:17-36 */
} igo_patch;

/* ---- a1: dune/iga/bsplinealgorithms.hh:24-32 findSpanCorrected ------------------------------------------------ */
long igo_find_span(int p, double u, const double* U, int n, int offset) {
  if (u <= U[0]) return p;
  if (u >= U[n - 1]) return (long)n - p - 2;
  /* std::upper_bound(U.begin()+p-1+offset, U.end(), u): first element > u */
  long lo = p - 1 + offset, hi = n;
  while (lo < hi) {
    long mid = lo + (hi - lo) / 2;
    if (!(u < U[mid]))
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo - 1;
}

/* FloatCmp::eq, relativeWeak, eps = 8*DBL_EPSILON (dune-common default, used at bsplinealgorithms.hh:104-108) */
static int feq(double a, double b) {
  const double eps = 8.0 * 2.220446049250313e-16;
  double m = fabs(a) > fabs(b) ? fabs(a) : fabs(b);
  return fabs(a - b) <= eps * m;
}

/* ---- a2: bsplinealgorithms.hh:96-137 BsplineBasis1D::basisFunctions (A2.2 variant) ----------------------------
 * span < 0 means "search".  Clamp of u and exact unit vectors at the two patch ends: lines 103-111.           */
void igo_bspline_basis(double u, const double* U, int n, int p, int span, double* N) {
  for (int i = 0; i <= p; ++i) N[i] = 0.0;
  if (u < U[0]) u = U[0];
  if (u > U[n - 1]) u = U[n - 1];
  if (feq(u, U[n - 1])) { N[p] = 1.0; return; }
  if (feq(u, U[0])) { N[0] = 1.0; return; }
  const int sp = span >= 0 ? span : (int)igo_find_span(p, u, U, n, 0);
  /* lDiff[k] = u - U[sp - k], rDiff[k] = U[sp+1+k] - u, k = 0..p-1 (lines 115-117) */
  N[0] = 1.0;
  for (int j = 0; j < p; ++j) { /* lines 121-132: level j+1 */
    double saved = 0.0;
    for (int r = 0; r <= j; ++r) {
      const double rD = U[sp + 1 + r] - u;
      const double lD = u - U[sp - (j - r)];
      const double tmp = N[r] / (rD + lD);
      N[r] = saved + rD * tmp;
      saved = lD * tmp;
    }
    N[j + 1] = saved;
  }
}

/* ---- a3: bsplinealgorithms.hh:148-222 basisFunctionDerivatives (A2.3); no clamping, no end special case -------
 * out[(k+1)][p+1], row = derivative order.                                                                     */
void igo_bspline_ders(double u, const double* U, int n, int p, int k, int span, double* dN) {
  const int order = p + 1;
  const int sp = span >= 0 ? span : (int)igo_find_span(p, u, U, n, 0);
  double left[MAXO], right[MAXO], ndu[MAXO][MAXO], a[2][MAXO];
  ndu[0][0] = 1.0;
  for (int j = 1; j <= p; ++j) { /* lines 161-173 */
    left[j] = u - U[sp + 1 - j];
    right[j] = U[sp + j] - u;
    double saved = 0.0;
    for (int r = 0; r < j; ++r) {
      ndu[j][r] = right[r + 1] + left[j - r];
      double temp = ndu[r][j - 1] / ndu[j][r];
      ndu[r][j] = saved + right[r + 1] * temp;
      saved = left[j - r] * temp;
    }
    ndu[j][j] = saved;
  }
  for (int j = 0; j <= p; ++j) dN[j] = ndu[j][p];
  for (int r = 0; r <= p; ++r) { /* lines 183-213 */
    int s1 = 0, s2 = 1;
    a[0][0] = 1.0;
    for (int kk = 1; kk <= k; ++kk) {
      double d = 0.0;
      const int rk = r - kk, pk = p - kk;
      if (kk <= p) {
        if (r >= kk) {
          a[s2][0] = a[s1][0] / ndu[pk + 1][rk];
          d = a[s2][0] * ndu[rk][pk];
        }
        const int j1 = (rk >= -1) ? 1 : -rk;
        const int j2 = (r - 1 <= pk) ? kk - 1 : p - r;
        for (int j = j1; j <= j2; ++j) {
          a[s2][j] = (a[s1][j] - a[s1][j - 1]) / ndu[pk + 1][rk + j];
          d += a[s2][j] * ndu[rk + j][pk];
        }
        if (r <= pk) {
          a[s2][kk] = -a[s1][kk - 1] / ndu[pk + 1][r];
          d += a[s2][kk] * ndu[r][pk];
        }
      }
      dN[kk * order + r] = d;
      int t = s1; s1 = s2; s2 = t;
    }
  }
  /* lines 216-220: multiply by p!/(p-k)! */
  for (int r = p, kk = 1; kk <= k; ++kk) {
    for (int j = 0; j <= p; ++j) dN[kk * order + j] *= r;
    r *= (p - kk);
  }
  (void)order; (void)n;
}

/* ---- helpers --------------------------------------------------------------------------------------------------- */
static int nb_of(const igo_patch* P) {
  int r = 1;
  for (int d = 0; d < P->dim; ++d) r *= P->degree[d] + 1;
  return r;
}

static long binom(int n, int k) {
  if (k < 0 || k > n) return 0;
  long r = 1;
  for (int i = 0; i < k; ++i) r = r * (n - i) / (i + 1);
  return r;
}

/* unique knots: nurbspatch.hh:58-63 / nurbsbasis.hh:511-515 (unique_copy with FloatCmp::eq).
 * Returns count; uniq must hold nknots doubles.                                                                 */
int igo_unique_knots(const double* U, int n, double* uniq) {
  int m = 0;
  for (int i = 0; i < n; ++i)
    if (m == 0 || !feq(uniq[m - 1], U[i])) uniq[m++] = U[i];
  return m;
}

/* elements per direction = unique knots - 1 */
void igo_elements_per_dir(const igo_patch* P, int* nel) {
  double* tmp = (double*)malloc(sizeof(double) * 4096);
  for (int d = 0; d < P->dim; ++d) {
    double* u = P->nknots[d] > 4096 ? (double*)malloc(sizeof(double) * P->nknots[d]) : tmp;
    nel[d] = igo_unique_knots(P->knots[d], P->nknots[d], u) - 1;
    if (u != tmp) free(u);
  }
  free(tmp);
}

/* ---- a9/a11: element -> span per direction: nurbspatch.hh:248-253, nurbsbasis.hh:355-365 ----------------------
 * span_d[e_d] = findSpanCorrected(p_d, uniqueKnot_d[e_d], U_d, e_d).  out: concatenated per-direction tables.   */
void igo_span_tables(const igo_patch* P, int* nel, int** tables /* malloc'ed per dir */) {
  for (int d = 0; d < P->dim; ++d) {
    double* uq = (double*)malloc(sizeof(double) * P->nknots[d]);
    int m = igo_unique_knots(P->knots[d], P->nknots[d], uq);
    nel[d] = m - 1;
    tables[d] = (int*)malloc(sizeof(int) * (m > 1 ? m - 1 : 1));
    for (int e = 0; e < m - 1; ++e)
      tables[d][e] = (int)igo_find_span(P->degree[d], uq[e], P->knots[d], P->nknots[d], e);
    free(uq);
  }
}

/* spans of all elements, [nelem][dim], element index x-fastest (mdnet.hh:245-257, nurbsbasis.hh:683-691) */
long igo_patch_spans(const igo_patch* P, int* span) {
  int nel[MAXDIM]; int* tab[MAXDIM];
  igo_span_tables(P, nel, tab);
  long n = 1;
  for (int d = 0; d < P->dim; ++d) n *= nel[d];
  if (span)
    for (long e = 0; e < n; ++e) {
      long h = e;
      for (int d = 0; d < P->dim; ++d) { span[e * P->dim + d] = tab[d][h % nel[d]]; h /= nel[d]; }
    }
  for (int d = 0; d < P->dim; ++d) free(tab[d]);
  return n;
}

/* ---- a10: element-to-global DOF map: nurbsbasis.hh:552-587 createOriginalNodeIndices, :599-615 prepareForTrim ---
 * dof[e][i], i local index i-fastest; g_d = max(s_d - p_d, 0) + l_d; flat g = g_0 + n_0 (g_1 + n_1 g_2),
 * n_d = |U_d| - p_d - 1 (:631-633).  trimflag (nullable) per element: 1 == empty (ElementTrimFlag::empty); DOFs
 * touched only by empty elements are dropped and the rest renumbered ascending.  Returns number of DOFs.          */
long igo_dofmap(const igo_patch* P, const unsigned char* trimflag, int* dof) {
  const int dim = P->dim, nb = nb_of(P);
  int nel[MAXDIM]; int* tab[MAXDIM];
  igo_span_tables(P, nel, tab);
  long nelem = 1, ndof = 1;
  int nd[MAXDIM];
  for (int d = 0; d < dim; ++d) { nelem *= nel[d]; nd[d] = P->nknots[d] - P->degree[d] - 1; ndof *= nd[d]; }
  for (long e = 0; e < nelem; ++e) {
    int s[MAXDIM]; long h = e;
    for (int d = 0; d < dim; ++d) { s[d] = tab[d][h % nel[d]]; h /= nel[d]; }
    for (int i = 0; i < nb; ++i) {
      int l[MAXDIM], hh = i; long g[MAXDIM];
      for (int d = 0; d < dim; ++d) { l[d] = hh % (P->degree[d] + 1); hh /= (P->degree[d] + 1); }
      for (int d = 0; d < dim; ++d) { int b = s[d] - P->degree[d]; if (b < 0) b = 0; g[d] = b + l[d]; }
      long gi = g[dim - 1];
      for (int d = dim - 2; d >= 0; --d) gi = gi * nd[d] + g[d];
      dof[e * nb + i] = (int)gi;
    }
  }
  long result = ndof;
  if (trimflag) {
    char* used = (char*)calloc(ndof, 1);
    for (long e = 0; e < nelem; ++e)
      if (trimflag[e] != 1)
        for (int i = 0; i < nb; ++i) used[dof[e * nb + i]] = 1;
    int* map = (int*)malloc(sizeof(int) * ndof);
    long c = 0;
    for (long g = 0; g < ndof; ++g) map[g] = used[g] ? (int)c++ : -1;
    /* indices() of an empty element would throw in the reference (indexMap_.at); we mark such entries -1 */
    for (long k = 0; k < nelem * nb; ++k) dof[k] = map[dof[k]];
    result = c;
    free(used); free(map);
  }
  for (int d = 0; d < dim; ++d) free(tab[d]);
  return result;
}

/* ---- a4/a5: gather of the element's (p+1)^dim control points, start = span - degree: nurbsalgorithms.hh:80-89,
 * mdnet.hh:209-226 subNet (i fastest).  loc[nb][dimworld+1]                                                      */
static void gather_cp(const igo_patch* P, const int* span, double* loc) {
  const int dim = P->dim, dw1 = P->dimworld + 1, nb = nb_of(P);
  for (int i = 0; i < nb; ++i) {
    int hh = i; long g = 0, stride = 1;
    for (int d = 0; d < dim; ++d) {
      int l = hh % (P->degree[d] + 1); hh /= (P->degree[d] + 1);
      g += (long)(span[d] - P->degree[d] + l) * stride;
      stride *= P->ncp[d];
    }
    for (int c = 0; c < dw1; ++c) loc[i * dw1 + c] = P->cp[g * dw1 + c];
  }
}

/* ---- a6: nurbsalgorithms.hh:160-180 Nurbs::basisFunctions.  u in patch (knot) coordinates.  R[nb]. -------------
 * tensor product as mdnet.hh:101-117 (direction 0 copied, then scaled by direction 1, 2);
 * W = sequential dot (mdnet.hh:480-487); R = (Nhat*w)/W.                                                         */
void igo_nurbs_basis(const igo_patch* P, const double* u, const int* span_in, double* R) {
  const int dim = P->dim, nb = nb_of(P), dw1 = P->dimworld + 1;
  double N1[MAXDIM][MAXO]; int span[MAXDIM];
  for (int d = 0; d < dim; ++d) {
    igo_bspline_basis(u[d], P->knots[d], P->nknots[d], P->degree[d], span_in ? span_in[d] : -1, N1[d]);
    span[d] = span_in ? span_in[d] : (int)igo_find_span(P->degree[d], u[d], P->knots[d], P->nknots[d], 0);
  }
  double* loc = (double*)malloc(sizeof(double) * nb * dw1);
  gather_cp(P, span, loc);
  double W = 0.0;
  for (int i = 0; i < nb; ++i) {
    int hh = i; double v = 1.0;
    for (int d = 0; d < dim; ++d) { int l = hh % (P->degree[d] + 1); hh /= (P->degree[d] + 1); v = (d == 0) ? N1[0][l] : v * N1[d][l]; }
    R[i] = v;
    W += v * loc[i * dw1 + dw1 - 1];
  }
  for (int i = 0; i < nb; ++i) { R[i] *= loc[i * dw1 + dw1 - 1]; R[i] /= W; }
  free(loc);
}

/* ---- a7: nurbsalgorithms.hh:193-250 Nurbs::basisFunctionDerivatives ---------------------------------------------
 * All mixed derivatives alpha in {0..k}^dim.  out[(k+1)^dim][nb], alpha flat index first-direction-fastest.
 * A^(a) = w * prod_d N_d^(a_d) (:205-224); W^(a) = sum A^(a) (:223);
 * R^(a) = (A^(a) - sum_{0<b<=a} C(a,b) W^(b) R^(a-b)) / W^(0) (:226-248), ascending flat alpha.
 * NOTE the reference's flat inner loop double-subtracts some terms when two components of alpha are >= 2
 * (SURVEY.md Appendix C); no caller uses those.  This restatement implements the formula as written above
 * (generalised Piegl-Tiller 4.20), which coincides with the reference for every alpha with at most one
 * component >= 2 -- all alpha on the hot path (|alpha|<=2) and all alpha its tests pin ((k,0),(0,k),(k,1)).    */
void igo_nurbs_ders(const igo_patch* P, const double* u, int k, const int* span_in, double* out) {
  const int dim = P->dim, nb = nb_of(P), dw1 = P->dimworld + 1;
  double* ders1 = (double*)malloc(sizeof(double) * MAXDIM * (k + 1) * MAXO);
  int span[MAXDIM];
  for (int d = 0; d < dim; ++d) {
    span[d] = span_in ? span_in[d] : (int)igo_find_span(P->degree[d], u[d], P->knots[d], P->nknots[d], 0);
    igo_bspline_ders(u[d], P->knots[d], P->nknots[d], P->degree[d], k, span[d], ders1 + (long)d * (k + 1) * MAXO);
  }
  double* loc = (double*)malloc(sizeof(double) * nb * dw1);
  gather_cp(P, span, loc);
  int na = 1;
  for (int d = 0; d < dim; ++d) na *= (k + 1);
  double* Wd = (double*)malloc(sizeof(double) * na);
  for (int a = 0; a < na; ++a) {
    int al[MAXDIM], ha = a;
    for (int d = 0; d < dim; ++d) { al[d] = ha % (k + 1); ha /= (k + 1); }
    double W = 0.0;
    for (int i = 0; i < nb; ++i) {
      int hh = i; double v = 1.0;
      for (int d = 0; d < dim; ++d) {
        int l = hh % (P->degree[d] + 1); hh /= (P->degree[d] + 1);
        const double f = ders1[(long)d * (k + 1) * MAXO + (long)al[d] * (P->degree[d] + 1) + l];
        v = (d == 0) ? f : v * f;
      }
      W += v * loc[i * dw1 + dw1 - 1];       /* dot(netOfDerivativeNets[a], weights)  (:223) */
      out[(long)a * nb + i] = v * loc[i * dw1 + dw1 - 1]; /* R[a] *= weights (:222) */
    }
    Wd[a] = W;
  }
  for (int a = 0; a < na; ++a) {
    int al[MAXDIM], ha = a;
    for (int d = 0; d < dim; ++d) { al[d] = ha % (k + 1); ha /= (k + 1); }
    for (int b = 1; b < na; ++b) {
      int be[MAXDIM], hb = b, ok = 1; long c = 1;
      for (int d = 0; d < dim; ++d) { be[d] = hb % (k + 1); hb /= (k + 1); if (be[d] > al[d]) ok = 0; }
      if (!ok) continue;
      long amb = 0, stride = 1;
      for (int d = 0; d < dim; ++d) { c *= binom(al[d], be[d]); amb += (long)(al[d] - be[d]) * stride; stride *= (k + 1); }
      const double fac = (double)c * Wd[b];
      for (int i = 0; i < nb; ++i) out[(long)a * nb + i] -= fac * out[amb * nb + i];
    }
    for (int i = 0; i < nb; ++i) out[(long)a * nb + i] /= Wd[0];
  }
  free(Wd); free(loc); free(ders1);
}

/* ---- geometry helpers (closed forms of dune-geometry MatrixHelper; call sites nurbsgeometry.hh:200-210) -------- */
/* sqrtDetAAT: |det J| (square), ||row|| (1 x n), sqrt(sum of squared 2x2 minors) (2 x 3) */
double igo_sqrt_det_aat(int m, int n, const double* J) {
  if (m == 1) { double s = 0; for (int j = 0; j < n; ++j) s += J[j] * J[j]; return sqrt(s); }
  if (m == 2 && n == 2) return fabs(J[0] * J[3] - J[1] * J[2]);
  if (m == 2 && n == 3) {
    const double v0 = J[0] * J[4] - J[1] * J[3], v1 = J[0] * J[5] - J[3] * J[2], v2 = J[1] * J[5] - J[2] * J[4];
    return sqrt(v0 * v0 + v1 * v1 + v2 * v2);
  }
  if (m == 3 && n == 3)
    return fabs(J[0] * (J[4] * J[8] - J[5] * J[7]) - J[1] * (J[3] * J[8] - J[5] * J[6]) + J[2] * (J[3] * J[7] - J[4] * J[6]));
  return NAN;
}

/* rightInvA: JinvT = J^T (J J^T)^-1, n x m, row-major (nurbsgeometry.hh:205-210) */
void igo_right_inv(int m, int n, const double* J, double* out) {
  double G[9], Gi[9];
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += J[i * n + k] * J[j * n + k]; G[i * m + j] = s; }
  if (m == 1) Gi[0] = 1.0 / G[0];
  else if (m == 2) {
    const double det = G[0] * G[3] - G[1] * G[2];
    Gi[0] = G[3] / det; Gi[1] = -G[1] / det; Gi[2] = -G[2] / det; Gi[3] = G[0] / det;
  } else {
    const double c00 = G[4] * G[8] - G[5] * G[7], c01 = G[5] * G[6] - G[3] * G[8], c02 = G[3] * G[7] - G[4] * G[6];
    const double det = G[0] * c00 + G[1] * c01 + G[2] * c02;
    Gi[0] = c00 / det; Gi[1] = (G[2] * G[7] - G[1] * G[8]) / det; Gi[2] = (G[1] * G[5] - G[2] * G[4]) / det;
    Gi[3] = c01 / det; Gi[4] = (G[0] * G[8] - G[2] * G[6]) / det; Gi[5] = (G[2] * G[3] - G[0] * G[5]) / det;
    Gi[6] = c02 / det; Gi[7] = (G[1] * G[6] - G[0] * G[7]) / det; Gi[8] = (G[0] * G[4] - G[1] * G[3]) / det;
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j) { double s = 0; for (int k = 0; k < m; ++k) s += J[k * n + i] * Gi[k * m + j]; out[i * m + j] = s; }
}

/* second-derivative multi-indices in the reference's H row order (nurbsgeometry.hh:262-289):
 * pure ones first, then mixed (1,1,0),(1,0,1),(0,1,1).  Returns count nh = dim(dim+1)/2.                         */
int igo_second_multiindices(int dim, int (*a)[MAXDIM]) {
  int n = 0;
  for (int d = 0; d < dim; ++d) { for (int c = 0; c < MAXDIM; ++c) a[n][c] = 0; a[n][d] = 2; ++n; }
  for (int c = 0; c < dim; ++c)
    for (int d = c + 1; d < dim; ++d) {
      /* order (0,1),(0,2),(1,2) == next_permutation(greater) of (1,1,0) */
      for (int q = 0; q < MAXDIM; ++q) a[n][q] = 0;
      a[n][c] = 1; a[n][d] = 1; ++n;
    }
  return n;
}

/* ---- one point on one element: the per-point body of a8, a12-a17 --------------------------------------------------
 * xi element-local in [0,1]^dim; u_d = xi_d*scale_d + offset_d (nurbsgeometry.hh:365-378, nurbsbasis.hh:88-89).
 * N, x: from basisFunctions (A2.2 path: nurbsgeometry.hh:133-138, nurbsbasis.hh:77-82,637-643).
 * dN[i][d] = scale_d R_i^(e_d) (nurbsbasis.hh:93-95); Jt[d] = scale_d sum R_i^(e_d) P_i (nurbsgeometry.hh:189-194);
 * d2N[i][r], H[r] with scale^alpha (nurbsbasis.hh:105-107, nurbsgeometry.hh:262-289);
 * detJ = sqrtDetAAT (nurbsgeometry.hh:200-203); JinvT = rightInvA (:205-210).                                     */
static void eval_point(const igo_patch* P, const int* span, const double* loc /*[nb][dw+1]*/, const double* xi,
                       int order, double* N, double* dN, double* d2N, double* x, double* Jt, double* H, double* detJ,
                       double* JinvT, double* work /* >= 9*nb doubles */) {
  const int dim = P->dim, dw = P->dimworld, dw1 = dw + 1, nb = nb_of(P);
  double u[MAXDIM], scale[MAXDIM];
  for (int d = 0; d < dim; ++d) {
    const double* U = P->knots[d];
    scale[d] = U[span[d] + 1] - U[span[d]];
    u[d] = xi[d] * scale[d] + U[span[d]];
  }
  if (N || x) {
    double* R = work;
    igo_nurbs_basis(P, u, span, R);
    if (N) for (int i = 0; i < nb; ++i) N[i] = R[i];
    if (x)
      for (int c = 0; c < dw; ++c) { double s = 0; for (int i = 0; i < nb; ++i) s += R[i] * loc[i * dw1 + c]; x[c] = s; }
  }
  const int want1 = (dN || Jt || detJ || JinvT), want2 = (order >= 2) && (d2N || H);
  if (order >= 1 && (want1 || want2)) {
    const int k = want2 ? 2 : 1;
    double* dR = work; /* (k+1)^dim * nb */
    igo_nurbs_ders(P, u, k, span, dR);
    double J[9];
    long stride = 1;
    for (int d = 0; d < dim; ++d) {
      const double* Rd = dR + stride * nb; /* alpha = e_d: flat index (k+1)^d */
      if (dN) for (int i = 0; i < nb; ++i) dN[i * dim + d] = Rd[i] * scale[d];
      for (int c = 0; c < dw; ++c) { double s = 0; for (int i = 0; i < nb; ++i) s += Rd[i] * loc[i * dw1 + c]; J[d * dw + c] = s * scale[d]; }
      stride *= (k + 1);
    }
    if (Jt) for (int q = 0; q < dim * dw; ++q) Jt[q] = J[q];
    if (detJ) *detJ = igo_sqrt_det_aat(dim, dw, J);
    if (JinvT) igo_right_inv(dim, dw, J, JinvT);
    if (want2) {
      int a2[6][MAXDIM];
      const int nh = igo_second_multiindices(dim, a2);
      for (int r = 0; r < nh; ++r) {
        long fa = 0, st = 1; double s = 1.0;
        for (int d = 0; d < dim; ++d) { fa += a2[r][d] * st; st *= 3; for (int q = 0; q < a2[r][d]; ++q) s *= scale[d]; }
        const double* Rr = dR + fa * nb;
        if (d2N) for (int i = 0; i < nb; ++i) d2N[i * nh + r] = Rr[i] * s;
        if (H)
          for (int c = 0; c < dw; ++c) { double t = 0; for (int i = 0; i < nb; ++i) t += Rr[i] * loc[i * dw1 + c]; H[r * dw + c] = t * s; }
      }
    }
  }
}

static void element_span(const igo_patch* P, const int* nel, int* const* tab, long e, int* span) {
  long h = e;
  for (int d = 0; d < P->dim; ++d) { span[d] = tab[d][h % nel[d]]; h /= nel[d]; }
}

/* ---- a19/a20: all elements x tensor rule (nurbsgridentity.hh:120-141, nurbsleafgridview.hh:147-155) -------------
 * outputs [pt][...], pt = (e-elem_begin)*nq + q, q first-direction-fastest; any pointer may be NULL.            */
int igo_eval_tensor(const igo_patch* P, long elem_begin, long elem_end, int order, const int* nq1, const double* xi1d,
                    double* N, double* dN, double* d2N, double* x, double* Jt, double* H, double* detJ, double* JinvT,
                    int nthreads) {
  const int dim = P->dim, dw = P->dimworld, nb = nb_of(P), nh = dim * (dim + 1) / 2;
  int nel[MAXDIM]; int* tab[MAXDIM];
  igo_span_tables(P, nel, tab);
  long nq = 1; const double* xid[MAXDIM]; { long off = 0; for (int d = 0; d < dim; ++d) { xid[d] = xi1d + off; off += nq1[d]; nq *= nq1[d]; } }
  int used = 1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
  used = omp_get_max_threads();
#endif
#pragma omp parallel
  {
    double* loc = (double*)malloc(sizeof(double) * nb * (dw + 1));
    double* work = (double*)malloc(sizeof(double) * 27 * nb);
#pragma omp for schedule(static)
    for (long e = elem_begin; e < elem_end; ++e) {
      int span[MAXDIM];
      element_span(P, nel, tab, e, span);
      gather_cp(P, span, loc);
      for (long q = 0; q < nq; ++q) {
        double xi[MAXDIM]; long hq = q;
        for (int d = 0; d < dim; ++d) { xi[d] = xid[d][hq % nq1[d]]; hq /= nq1[d]; }
        const long pt = (e - elem_begin) * nq + q;
        eval_point(P, span, loc, xi, order, N ? N + pt * nb : 0, dN ? dN + pt * nb * dim : 0,
                   d2N ? d2N + pt * nb * nh : 0, x ? x + pt * dw : 0, Jt ? Jt + pt * dim * dw : 0,
                   H ? H + pt * nh * dw : 0, detJ ? detJ + pt : 0, JinvT ? JinvT + pt * dim * dw : 0, work);
      }
    }
    free(loc); free(work);
  }
  for (int d = 0; d < dim; ++d) free(tab[d]);
  return used;
}

/* arbitrary element-local points (trimmed batches, fillquadraturerule.hh:8-19 output): elem[npts], xi[npts][dim] */
int igo_eval_points(const igo_patch* P, long npts, const int* elem, const double* xi, int order, double* N, double* dN,
                    double* d2N, double* x, double* Jt, double* H, double* detJ, double* JinvT, int nthreads) {
  const int dim = P->dim, dw = P->dimworld, nb = nb_of(P), nh = dim * (dim + 1) / 2;
  int nel[MAXDIM]; int* tab[MAXDIM];
  igo_span_tables(P, nel, tab);
  int used = 1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
  used = omp_get_max_threads();
#endif
#pragma omp parallel
  {
    double* loc = (double*)malloc(sizeof(double) * nb * (dw + 1));
    double* work = (double*)malloc(sizeof(double) * 27 * nb);
#pragma omp for schedule(static)
    for (long pt = 0; pt < npts; ++pt) {
      int span[MAXDIM];
      element_span(P, nel, tab, elem[pt], span);
      gather_cp(P, span, loc);
      eval_point(P, span, loc, xi + pt * dim, order, N ? N + pt * nb : 0, dN ? dN + pt * nb * dim : 0,
                 d2N ? d2N + pt * nb * nh : 0, x ? x + pt * dw : 0, Jt ? Jt + pt * dim * dw : 0,
                 H ? H + pt * nh * dw : 0, detJ ? detJ + pt : 0, JinvT ? JinvT + pt * dim * dw : 0, work);
    }
    free(loc); free(work);
  }
  for (int d = 0; d < dim; ++d) free(tab[d]);
  return used;
}

/* partial(alpha): nurbsbasis.hh:99-108,667-677 -- derivative order handed down = sum(alpha); scaled by prod scale^alpha */
void igo_partial(const igo_patch* P, long npts, const int* elem, const double* xi, const int* alpha, double* out) {
  const int dim = P->dim, nb = nb_of(P);
  int nel[MAXDIM]; int* tab[MAXDIM];
  igo_span_tables(P, nel, tab);
  int k = 0; for (int d = 0; d < dim; ++d) k += alpha[d];
  long na = 1; for (int d = 0; d < dim; ++d) na *= (k + 1);
  double* dR = (double*)malloc(sizeof(double) * na * nb);
  for (long pt = 0; pt < npts; ++pt) {
    int span[MAXDIM]; double u[MAXDIM], s = 1.0; long fa = 0, st = 1;
    element_span(P, nel, tab, elem[pt], span);
    for (int d = 0; d < dim; ++d) {
      const double* U = P->knots[d];
      const double sc = U[span[d] + 1] - U[span[d]];
      u[d] = xi[pt * dim + d] * sc + U[span[d]];
      for (int q = 0; q < alpha[d]; ++q) s *= sc;
      fa += alpha[d] * st; st *= (k + 1);
    }
    igo_nurbs_ders(P, u, k, span, dR);
    for (int i = 0; i < nb; ++i) out[pt * nb + i] = dR[fa * nb + i] * s;
  }
  free(dR);
  for (int d = 0; d < dim; ++d) free(tab[d]);
}

/* ---- patch-coordinate geometry (nurbspatchgeometry.hh:76-82,163-185): span searched per point ------------------ */
void igo_patchgeo(const igo_patch* P, const double* u, double* x, double* Jt, double* detJ) {
  const int dim = P->dim, dw = P->dimworld, dw1 = dw + 1, nb = nb_of(P);
  int span[MAXDIM];
  for (int d = 0; d < dim; ++d) span[d] = (int)igo_find_span(P->degree[d], u[d], P->knots[d], P->nknots[d], 0);
  double* loc = (double*)malloc(sizeof(double) * nb * dw1);
  double* R = (double*)malloc(sizeof(double) * nb * 8);
  gather_cp(P, span, loc);
  if (x) {
    igo_nurbs_basis(P, u, 0, R);
    for (int c = 0; c < dw; ++c) { double s = 0; for (int i = 0; i < nb; ++i) s += R[i] * loc[i * dw1 + c]; x[c] = s; }
  }
  if (Jt || detJ) {
    double J[9];
    igo_nurbs_ders(P, u, 1, 0, R);
    long stride = 1;
    for (int d = 0; d < dim; ++d) {
      for (int c = 0; c < dw; ++c) { double s = 0; for (int i = 0; i < nb; ++i) s += R[stride * nb + i] * loc[i * dw1 + c]; J[d * dw + c] = s; }
      stride *= 2;
    }
    if (Jt) for (int q = 0; q < dim * dw; ++q) Jt[q] = J[q];
    if (detJ) *detJ = igo_sqrt_det_aat(dim, dw, J);
  }
  free(loc); free(R);
}

/* ---- a19: Gauss-Legendre on [0,1], n points (stands in for dune-geometry QuadratureRules; the C-ABI takes the rule
 * as an input, so this only feeds tests and benchmarks)                                                           */
void igo_gauss01(int n, double* x, double* w) {
  for (int i = 0; i < n; ++i) {
    double z = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 1;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1, p2 = 0;
      for (int j = 1; j <= n; ++j) { double p3 = p2; p2 = p1; p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j; }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      double dz = p1 / pp; z -= dz;
      if (fabs(dz) < 1e-16) break;
    }
    x[n - 1 - i] = 0.5 * (z + 1.0);
    w[n - 1 - i] = 1.0 / ((1.0 - z * z) * pp * pp);
  }
}

/* ---- a18: volume = sum_e sum_q detJ w  (nurbsgeometry.hh:96-113 per element, summed as in gridtests.cpp:338-341) */
double igo_volume(const igo_patch* P, const int* nq1, const double* xi1d, const double* w1d, int nthreads) {
  const int dim = P->dim;
  long nelem = igo_patch_spans(P, 0), nq = 1;
  for (int d = 0; d < dim; ++d) nq *= nq1[d];
  double* detJ = (double*)malloc(sizeof(double) * nelem * nq);
  igo_eval_tensor(P, 0, nelem, 1, nq1, xi1d, 0, 0, 0, 0, 0, 0, detJ, 0, nthreads);
  const double* wd[MAXDIM]; { long off = 0; for (int d = 0; d < dim; ++d) { wd[d] = w1d + off; off += nq1[d]; } }
  double vol = 0;
  for (long e = 0; e < nelem; ++e) {
    double ve = 0;
    for (long q = 0; q < nq; ++q) {
      double w = 1; long hq = q;
      for (int d = 0; d < dim; ++d) { w *= wd[d][hq % nq1[d]]; hq /= nq1[d]; }
      ve += detJ[e * nq + q] * w;
    }
    vol += ve;
  }
  free(detJ);
  return vol;
}

/* ---- a21: element matrices, the plain-loop generalisation of dune/python/test/poisson.py:37-81 ------------------
 * kind 0 (Poisson / scalar Laplace, poisson.py:64-77):  K_e[i][j] = sum_q w detJ grad_i . grad_j,
 *         grad_i = JinvT dN_i (poisson.py:58,64-66);  f_e[i] = sum_q w detJ N_i f  with f == fconst (poisson.py:79).
 * kind 1 (mass): K_e[i][j] = sum_q w detJ N_i N_j.
 * kind 2 (linear elasticity, B^T D B, dim == dimworld): DOF numbering flat-interleaved i*dim + c as dune-functions'
 *         PowerPreBasis<FlatInterleaved> (python/dune/iga/basis/__init__.py:53-67); strain Voigt order
 *         2-D (xx,yy,xy), 3-D (xx,yy,zz,yz,xz,xy) with engineering shear; D row-major ns x ns.
 * Ke[e][n][n] row-major, n = nb (kind 0,1) or nb*dim (kind 2); fe[e][n] (nullable).                              */
/* one quadrature point's contribution to K_e / f_e (poisson.py:52-79 generalised to B^T D B); wd_ = w detJ */
static void accum_point(int dim, int dw, int nb, int kind, const double* D, double fconst, double wd_, const double* Npt,
                        const double* dNpt, const double* JinvTpt, double* grad, double* B, double* DB, double* K,
                        double* fe_e) {
  const int ncomp = (kind == 2) ? dim : 1, n = nb * ncomp;
  const int ns = (dim == 2) ? 3 : (dim == 3 ? 6 : 1);
  /* grad_i[c] = sum_d JinvT[c][d] dN_i[d]   (JinvT is dw x dim) */
  for (int i = 0; i < nb; ++i)
    for (int c = 0; c < dw; ++c) {
      double s = 0;
      for (int d = 0; d < dim; ++d) s += JinvTpt[c * dim + d] * dNpt[i * dim + d];
      grad[i * dw + c] = s;
    }
  if (kind == 0) {
    for (int i = 0; i < nb; ++i) {
      for (int j = 0; j < nb; ++j) {
        double s = 0;
        for (int c = 0; c < dw; ++c) s += grad[i * dw + c] * grad[j * dw + c];
        K[i * n + j] += wd_ * s;
      }
      if (fe_e) fe_e[i] += wd_ * Npt[i] * fconst;
    }
  } else if (kind == 1) {
    for (int i = 0; i < nb; ++i) {
      for (int j = 0; j < nb; ++j) K[i * n + j] += wd_ * Npt[i] * Npt[j];
      if (fe_e) fe_e[i] += wd_ * Npt[i] * fconst;
    }
  } else {
    for (int q2 = 0; q2 < ns * n; ++q2) B[q2] = 0;
    for (int i = 0; i < nb; ++i) {
      const double gx = grad[i * dw + 0], gy = grad[i * dw + 1], gz = dim == 3 ? grad[i * dw + 2] : 0;
      if (dim == 2) {
        B[0 * n + i * 2 + 0] = gx; B[1 * n + i * 2 + 1] = gy;
        B[2 * n + i * 2 + 0] = gy; B[2 * n + i * 2 + 1] = gx;
      } else {
        B[0 * n + i * 3 + 0] = gx; B[1 * n + i * 3 + 1] = gy; B[2 * n + i * 3 + 2] = gz;
        B[3 * n + i * 3 + 1] = gz; B[3 * n + i * 3 + 2] = gy;
        B[4 * n + i * 3 + 0] = gz; B[4 * n + i * 3 + 2] = gx;
        B[5 * n + i * 3 + 0] = gy; B[5 * n + i * 3 + 1] = gx;
      }
    }
    for (int r = 0; r < ns; ++r)
      for (int j = 0; j < n; ++j) { double s = 0; for (int t = 0; t < ns; ++t) s += D[r * ns + t] * B[t * n + j]; DB[r * n + j] = s; }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) { double s = 0; for (int r = 0; r < ns; ++r) s += B[r * n + i] * DB[r * n + j]; K[i * n + j] += wd_ * s; }
    if (fe_e) for (int i = 0; i < nb; ++i) for (int c = 0; c < ncomp; ++c) fe_e[i * ncomp + c] += wd_ * Npt[i] * fconst;
  }
}

int igo_assemble(const igo_patch* P, int kind, const double* D, double fconst, long elem_begin, long elem_end,
                 const int* nq1, const double* xi1d, const double* w1d, double* Ke, double* fe, int nthreads) {
  const int dim = P->dim, dw = P->dimworld, nb = nb_of(P);
  const int ncomp = (kind == 2) ? dim : 1, n = nb * ncomp;
  const int ns = (dim == 2) ? 3 : (dim == 3 ? 6 : 1);
  long nq = 1; for (int d = 0; d < dim; ++d) nq *= nq1[d];
  const long ne = elem_end - elem_begin;
  int used = 1;
  /* evaluate with the restated reference path, then plain loops */
  double* N = (double*)malloc(sizeof(double) * ne * nq * nb);
  double* dN = (double*)malloc(sizeof(double) * ne * nq * nb * dim);
  double* detJ = (double*)malloc(sizeof(double) * ne * nq);
  double* JinvT = (double*)malloc(sizeof(double) * ne * nq * dim * dw);
  used = igo_eval_tensor(P, elem_begin, elem_end, 1, nq1, xi1d, N, dN, 0, 0, 0, 0, detJ, JinvT, nthreads);
  const double* wd[MAXDIM]; { long off = 0; for (int d = 0; d < dim; ++d) { wd[d] = w1d + off; off += nq1[d]; } }
#pragma omp parallel
  {
    double* grad = (double*)malloc(sizeof(double) * nb * dw);
    double* B = (double*)malloc(sizeof(double) * ns * n);
    double* DB = (double*)malloc(sizeof(double) * ns * n);
#pragma omp for schedule(static)
    for (long e = 0; e < ne; ++e) {
      double* K = Ke + e * (long)n * n;
      for (long q = 0; q < (long)n * n; ++q) K[q] = 0;
      if (fe) for (int i = 0; i < n; ++i) fe[e * n + i] = 0;
      for (long q = 0; q < nq; ++q) {
        const long pt = e * nq + q;
        double w = 1; long hq = q;
        for (int d = 0; d < dim; ++d) { w *= wd[d][hq % nq1[d]]; hq /= nq1[d]; }
        accum_point(dim, dw, nb, kind, D, fconst, w * detJ[pt], N + pt * nb, dN + pt * nb * dim, JinvT + pt * dim * dw, grad, B, DB,
                    K, fe ? fe + e * n : 0);
      }
    }
    free(grad); free(B); free(DB);
  }
  free(N); free(dN); free(detJ); free(JinvT);
  return used;
}

/* Element matrices of elements that carry their own quadrature (trimmed elements: the rule of
 * utils/fillquadraturerule.hh:8-19): list element k is patch element elem[k] with the points offset[k] .. offset[k+1]
 * of xi[npts][dim] (element-local) and w[npts].  Same per-point contribution as igo_assemble.                      */
int igo_assemble_points(const igo_patch* P, int kind, const double* D, double fconst, long nlist, const int* elem,
                        const long* offset, const double* xi, const double* w, double* Ke, double* fe) {
  const int dim = P->dim, dw = P->dimworld, nb = nb_of(P);
  const int ncomp = (kind == 2) ? dim : 1, n = nb * ncomp;
  const int ns = (dim == 2) ? 3 : (dim == 3 ? 6 : 1);
  const long npts = offset[nlist];
  int* pe = (int*)malloc(sizeof(int) * (npts > 0 ? npts : 1));
  for (long k = 0; k < nlist; ++k) for (long q = offset[k]; q < offset[k + 1]; ++q) pe[q] = elem[k];
  double* N = (double*)malloc(sizeof(double) * (npts + 1) * nb);
  double* dN = (double*)malloc(sizeof(double) * (npts + 1) * nb * dim);
  double* detJ = (double*)malloc(sizeof(double) * (npts + 1));
  double* JinvT = (double*)malloc(sizeof(double) * (npts + 1) * dim * dw);
  igo_eval_points(P, npts, pe, xi, 1, N, dN, 0, 0, 0, 0, detJ, JinvT, 0);
  double* grad = (double*)malloc(sizeof(double) * nb * dw);
  double* B = (double*)malloc(sizeof(double) * ns * n);
  double* DB = (double*)malloc(sizeof(double) * ns * n);
  for (long k = 0; k < nlist; ++k) {
    double* K = Ke + k * (long)n * n;
    for (long q = 0; q < (long)n * n; ++q) K[q] = 0;
    if (fe) for (int i = 0; i < n; ++i) fe[k * n + i] = 0;
    for (long pt = offset[k]; pt < offset[k + 1]; ++pt)
      accum_point(dim, dw, nb, kind, D, fconst, w[pt] * detJ[pt], N + pt * nb, dN + pt * nb * dim, JinvT + pt * dim * dw, grad, B,
                  DB, K, fe ? fe + k * n : 0);
  }
  free(pe); free(N); free(dN); free(detJ); free(JinvT); free(grad); free(B); free(DB);
  return 0;
}

/* ---- a21 with a non-constant source: dune/python/test/poisson.py:79  localB[i] += weight * integrationElement *
 * phi[i] * f(posGlobal), the caller's f given per quadrature point (f[pt][ncomp]; vector-valued: flat-interleaved
 * i*ncomp + c).  Tensor rule over [elem_begin, elem_end) (f indexed (e - elem_begin)*nq + q) ...                   */
int igo_load_vector(const igo_patch* P, long elem_begin, long elem_end, const int* nq1, const double* xi1d,
                    const double* w1d, const double* f, int ncomp, double* fe) {
  const int dim = P->dim, nb = nb_of(P);
  long nq = 1; for (int d = 0; d < dim; ++d) nq *= nq1[d];
  const long ne = elem_end - elem_begin;
  double* N = (double*)malloc(sizeof(double) * (ne * nq + 1) * nb);
  double* detJ = (double*)malloc(sizeof(double) * (ne * nq + 1));
  igo_eval_tensor(P, elem_begin, elem_end, 1, nq1, xi1d, N, 0, 0, 0, 0, 0, detJ, 0, 0);
  const double* wd[MAXDIM]; { long off = 0; for (int d = 0; d < dim; ++d) { wd[d] = w1d + off; off += nq1[d]; } }
  for (long e = 0; e < ne; ++e) {
    for (int i = 0; i < nb * ncomp; ++i) fe[e * nb * ncomp + i] = 0;
    for (long q = 0; q < nq; ++q) {
      const long pt = e * nq + q;
      double w = 1; long hq = q;
      for (int d = 0; d < dim; ++d) { w *= wd[d][hq % nq1[d]]; hq /= nq1[d]; }
      for (int i = 0; i < nb; ++i)
        for (int c = 0; c < ncomp; ++c) fe[(e * nb + i) * ncomp + c] += w * detJ[pt] * N[pt * nb + i] * f[pt * ncomp + c];
    }
  }
  free(N); free(detJ);
  return 0;
}
/* ... and over elements with their own rules (trimmed elements, fillquadraturerule.hh:8-19), lists as igo_assemble_points */
int igo_load_vector_points(const igo_patch* P, long nlist, const int* elem, const long* offset, const double* xi,
                           const double* w, const double* f, int ncomp, double* fe) {
  const int nb = nb_of(P);
  const long npts = offset[nlist];
  int* pe = (int*)malloc(sizeof(int) * (npts > 0 ? npts : 1));
  for (long k = 0; k < nlist; ++k) for (long q = offset[k]; q < offset[k + 1]; ++q) pe[q] = elem[k];
  double* N = (double*)malloc(sizeof(double) * (npts + 1) * nb);
  double* detJ = (double*)malloc(sizeof(double) * (npts + 1));
  igo_eval_points(P, npts, pe, xi, 1, N, 0, 0, 0, 0, 0, detJ, 0, 0);
  for (long k = 0; k < nlist; ++k) {
    for (int i = 0; i < nb * ncomp; ++i) fe[k * nb * ncomp + i] = 0;
    for (long pt = offset[k]; pt < offset[k + 1]; ++pt)
      for (int i = 0; i < nb; ++i)
        for (int c = 0; c < ncomp; ++c) fe[(k * nb + i) * ncomp + c] += w[pt] * detJ[pt] * N[pt * nb + i] * f[pt * ncomp + c];
  }
  free(pe); free(N); free(detJ);
  return 0;
}

/* ---- a11: dune/iga/nurbspatch.hh:599-631 getDirectIndex<0> / getRealIndex<0> with the element map filled at
 * :696-716 (direct indices of the trimmed and full elements in ascending order).  real_of_direct[e] = -1 where the
 * reference throws "No corresponding realIndex" (empty elements).  trimflag NULL: identity (:599-606).  Returns n_real.
 * Pinned by dune/iga/test/trimmedgridtests.cpp:335-373 (tests/test_oracle.py).                                     */
long igo_real_index_maps(long nelem, const unsigned char* trimflag, int* direct_of_real, int* real_of_direct) {
  long r = 0;
  for (long e = 0; e < nelem; ++e) {
    const int keep = !trimflag || trimflag[e] != 1; /* ElementTrimFlag::empty == 1 in this library's encoding */
    if (keep) {
      if (direct_of_real) direct_of_real[r] = (int)e;
      if (real_of_direct) real_of_direct[e] = (int)r;
      ++r;
    } else if (real_of_direct) {
      real_of_direct[e] = -1;
    }
  }
  return r;
}

/* ---- a19, Gauss-Lobatto branch of fillQuadratureRuleImpl (utils/fillquadraturerule.hh:21-44): points that several
 * sub-triangles share are merged -- exact coordinate equality, weights added in the order of appearance -- and the rule
 * comes out sorted lexicographically (x, then y): the std::map with the Comparator of :23-33.  In place on xi[n][2],
 * w[n]; returns the number of unique points.                                                                       */
long igo_merge_duplicate_points(long n, double* xi, double* w) {
  long m = 0; /* insertion into a sorted prefix = what std::map::try_emplace does, O(n^2) is fine for a checker */
  double* ox = (double*)malloc(sizeof(double) * 2 * (n > 0 ? n : 1));
  double* ow = (double*)malloc(sizeof(double) * (n > 0 ? n : 1));
  for (long k = 0; k < n; ++k) {
    const double x = xi[2 * k], y = xi[2 * k + 1];
    long pos = 0;
    while (pos < m && (ox[2 * pos] < x || (!(ox[2 * pos] > x) && ox[2 * pos + 1] < y))) ++pos;
    if (pos < m && !(x < ox[2 * pos]) && !(ox[2 * pos] < x) && !(y < ox[2 * pos + 1])) { /* equivalent key */
      ow[pos] += w[k];
      continue;
    }
    for (long j = m; j > pos; --j) { ox[2 * j] = ox[2 * j - 2]; ox[2 * j + 1] = ox[2 * j - 1]; ow[j] = ow[j - 1]; }
    ox[2 * pos] = x; ox[2 * pos + 1] = y; ow[pos] = w[k];
    ++m;
  }
  for (long k = 0; k < m; ++k) { xi[2 * k] = ox[2 * k]; xi[2 * k + 1] = ox[2 * k + 1]; w[k] = ow[k]; }
  free(ox); free(ow);
  return m;
}

/* ---- a19 (trimmed): utils/fillquadraturerule.hh:8-19 -- rule of a trimmed element from its sub-triangles -------
 * tri[ntri][3][2] in element-local [0,1]^2; simplex rule (ref_pts[nrp][2], ref_w[nrp]) on the reference triangle;
 * position = triangle.global(ip) (affine), weight = ip.w * triangle.integrationElement (= |det| = 2*area).       */
void igo_trimmed_rule(long ntri, const double* tri, int nrp, const double* ref_pts, const double* ref_w, double* xi,
                      double* w) {
  for (long t = 0; t < ntri; ++t) {
    const double* v = tri + t * 6;
    const double ax = v[2] - v[0], ay = v[3] - v[1], bx = v[4] - v[0], by = v[5] - v[1];
    const double ie = fabs(ax * by - ay * bx);
    for (int q = 0; q < nrp; ++q) {
      xi[(t * nrp + q) * 2 + 0] = v[0] + ax * ref_pts[q * 2] + bx * ref_pts[q * 2 + 1];
      xi[(t * nrp + q) * 2 + 1] = v[1] + ay * ref_pts[q * 2] + by * ref_pts[q * 2 + 1];
      w[t * nrp + q] = ref_w[q] * ie;
    }
  }
}

/* ---- a17 extras: NURBSGeometry::normal / unitNormal / metric / secondFundamentalForm / gaussianCurvature -------
 * dune/iga/nurbsgeometry.hh:216-255.  Jt [npts][dim][dw], H [npts][nh][dw] (rows uu, vv, uv); outputs nullable.
 * normal = J0 x J1, unit = normal / |normal|, metric = J J^T (MatrixHelper::AAT), b_ab = H_ab . unit,
 * K = det(b) / det(metric).  Pinned integrally by Gauss-Bonnet on the torus (gridtests.cpp:348-363).              */
void igo_surface_forms(int dim, int dw, long npts, const double* Jt, const double* H, double* normal, double* unit,
                       double* metric, double* secondff, double* gaussK) {
  for (long t = 0; t < npts; ++t) {
    const double* J = Jt + t * dim * dw;
    double g[9];
    for (int a = 0; a < dim; ++a)
      for (int b = 0; b < dim; ++b) {
        double s = 0;
        for (int c = 0; c < dw; ++c) s += J[a * dw + c] * J[b * dw + c];
        g[a * dim + b] = s;
      }
    if (metric) for (int q = 0; q < dim * dim; ++q) metric[t * dim * dim + q] = g[q];
    if (dim == 2 && dw == 3) {
      const double N[3] = {J[1] * J[5] - J[2] * J[4], J[2] * J[3] - J[0] * J[5], J[0] * J[4] - J[1] * J[3]};
      const double len = sqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
      const double n[3] = {N[0] / len, N[1] / len, N[2] / len};
      if (normal) for (int c = 0; c < 3; ++c) normal[t * 3 + c] = N[c];
      if (unit) for (int c = 0; c < 3; ++c) unit[t * 3 + c] = n[c];
      if (H && (secondff || gaussK)) {
        double b[3];
        for (int r = 0; r < 3; ++r) b[r] = H[(t * 3 + r) * 3] * n[0] + H[(t * 3 + r) * 3 + 1] * n[1] + H[(t * 3 + r) * 3 + 2] * n[2];
        if (secondff) { secondff[t * 4] = b[0]; secondff[t * 4 + 1] = secondff[t * 4 + 2] = b[2]; secondff[t * 4 + 3] = b[1]; }
        if (gaussK) gaussK[t] = (b[0] * b[1] - b[2] * b[2]) / (g[0] * g[3] - g[1] * g[2]);
      }
    }
  }
}

/* ---- NURBSGeometry::local (dune/iga/nurbsgeometry.hh:145-176): inverse of the ELEMENT map, xi in [0,1]^dim ----------
 * dim == dimworld: Newton from xi = 0.5, step dx = (Jt Jt^T)^-1 Jt (x(xi) - X) (MatrixHelper::xTRightInvA), clamp to
 *   [0,1] and stop when every coordinate sits on the boundary, else iterate while |dx|^2 > 16 eps; a singular Jacobian
 *   returns numeric_limits::max() in every coordinate (flag 1).
 * dim < dimworld (curves, surfaces): closestPointProjectionByTrustRegion, dune/iga/geometry/closestpointprojection.hh:
 *   minimise 0.5 |x(u) - X|^2 with Newton steps H du = -R limited to 1/4 of the domain, step halving while the
 *   energy grows (20 tries, then the reference throws MathError -> flag 2, xi = NaN), at most 100 iterations,
 *   stop when |R| < 1e-10 or when the clamped iterate sits on the boundary in every coordinate.
 * flag (nullable): 0 converged, 1 singular, 2 no descent direction, 3 iteration cap of the Newton branch (the
 * reference would loop forever).                                                                                */
static void geo012(const igo_patch* P, const int* span, const double* loc, const double* xi, int order, double* x,
                   double* J, double* H, double* work) {
  eval_point(P, span, loc, xi, order, 0, 0, 0, x, order >= 1 ? J : 0, order >= 2 ? H : 0, 0, 0, work);
}
static int solve_small(int n, const double* A, const double* b, double* y) {
  if (n == 1) { if (A[0] == 0.0) return 0; y[0] = b[0] / A[0]; return 1; }
  if (n == 2) {
    const double det = A[0] * A[3] - A[1] * A[2];
    if (det == 0.0) return 0;
    y[0] = (A[3] * b[0] - A[1] * b[1]) / det;
    y[1] = (A[0] * b[1] - A[2] * b[0]) / det;
    return 1;
  }
  const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  if (det == 0.0) return 0;
  const double i00 = c00, i01 = A[2] * A[7] - A[1] * A[8], i02 = A[1] * A[5] - A[2] * A[4];
  const double i10 = c01, i11 = A[0] * A[8] - A[2] * A[6], i12 = A[2] * A[3] - A[0] * A[5];
  const double i20 = c02, i21 = A[1] * A[6] - A[0] * A[7], i22 = A[0] * A[4] - A[1] * A[3];
  y[0] = (i00 * b[0] + i01 * b[1] + i02 * b[2]) / det;
  y[1] = (i10 * b[0] + i11 * b[1] + i12 * b[2]) / det;
  y[2] = (i20 * b[0] + i21 * b[1] + i22 * b[2]) / det;
  return 1;
}
static int clamp_all_boundaries(int dim, double* u) {
  int cnt = 0;
  for (int j = 0; j < dim; ++j) {
    u[j] = u[j] < 0.0 ? 0.0 : (u[j] > 1.0 ? 1.0 : u[j]);
    if (u[j] == 0.0 || u[j] == 1.0) ++cnt;
  }
  return cnt == dim;
}
int igo_geometry_local(const igo_patch* P, long npts, const int* elem, const double* xglobal, double* xi_out, int* flag) {
  const int dim = P->dim, dw = P->dimworld, nb = nb_of(P), nh = dim * (dim + 1) / 2;
  if (dim != dw && dim > 2) return 1;
  int nel[MAXDIM]; int* tab[MAXDIM];
  igo_span_tables(P, nel, tab);
  double* loc = (double*)malloc(sizeof(double) * nb * (dw + 1));
  double* work = (double*)malloc(sizeof(double) * 27 * nb);
  for (long pt = 0; pt < npts; ++pt) {
    int span[MAXDIM];
    element_span(P, nel, tab, elem[pt], span);
    gather_cp(P, span, loc);
    const double* X = xglobal + pt * dw;
    double u[MAXDIM], x[3], J[9], H[18];
    int fl = 0;
    for (int d = 0; d < dim; ++d) u[d] = 0.5;
    if (dim == dw) {
      const double tol = 16.0 * 2.220446049250313e-16;
      int it = 0;
      for (;;) {
        geo012(P, span, loc, u, 1, x, J, H, work);
        double dg[3], G[9], rhs[3], dx[3];
        for (int c = 0; c < dw; ++c) dg[c] = x[c] - X[c];
        for (int a = 0; a < dim; ++a) {
          double s = 0;
          for (int c = 0; c < dw; ++c) s += J[a * dw + c] * dg[c];
          rhs[a] = s;
          for (int b = 0; b < dim; ++b) { double q = 0; for (int c = 0; c < dw; ++c) q += J[a * dw + c] * J[b * dw + c]; G[a * dim + b] = q; }
        }
        if (!solve_small(dim, G, rhs, dx)) { fl = 1; for (int d = 0; d < dim; ++d) u[d] = 1.7976931348623157e308; break; }
        double n2 = 0;
        for (int d = 0; d < dim; ++d) { u[d] -= dx[d]; n2 += dx[d] * dx[d]; }
        if (clamp_all_boundaries(dim, u)) break;
        if (!(n2 > tol)) break;
        if (++it >= 100) { fl = 3; break; }
      }
    } else {
      const double tol = 1e-10, domainFraction = 0.25;
      double R[2], HV[4], energyVal, oldEnergy;
#define IGO_EGH()                                                                                   \
  do {                                                                                              \
    geo012(P, span, loc, u, 2, x, J, H, work);                                                      \
    double dist[3], e2 = 0;                                                                         \
    for (int c = 0; c < dw; ++c) { dist[c] = x[c] - X[c]; e2 += dist[c] * dist[c]; }                \
    for (int a = 0; a < dim; ++a) {                                                                 \
      double s = 0;                                                                                 \
      for (int c = 0; c < dw; ++c) s += J[a * dw + c] * dist[c];                                    \
      R[a] = s;                                                                                     \
      for (int b = 0; b < dim; ++b) { double q = 0; for (int c = 0; c < dw; ++c) q += J[a * dw + c] * J[b * dw + c]; HV[a * dim + b] = q; } \
    }                                                                                               \
    double dd[3];                                                                                   \
    for (int r = 0; r < nh; ++r) { double s = 0; for (int c = 0; c < dw; ++c) s += H[r * dw + c] * dist[c]; dd[r] = s; } \
    for (int a = 0; a < dim; ++a) HV[a * dim + a] += dd[a];                                         \
    if (dim == 2) { HV[1] += dd[2]; HV[2] += dd[2]; }                                               \
    oldEnergy = 0.5 * e2;                                                                           \
  } while (0)
#define IGO_ENERGY(res)                                                                             \
  do {                                                                                              \
    geo012(P, span, loc, u, 0, x, J, H, work);                                                      \
    double e2 = 0;                                                                                  \
    for (int c = 0; c < dw; ++c) e2 += (x[c] - X[c]) * (x[c] - X[c]);                               \
    res = 0.5 * e2;                                                                                 \
  } while (0)
      IGO_EGH();
      energyVal = oldEnergy;
      for (int i = 0; i < 100 && !fl; ++i) {
        double du[2], mR[2] = {-R[0], dim == 2 ? -R[1] : 0.0};
        if (!solve_small(dim, HV, mR, du)) { fl = 1; break; }
        double duNorm = 0;
        for (int d = 0; d < dim; ++d) duNorm += du[d] * du[d];
        duNorm = sqrt(duNorm);
        if (duNorm > domainFraction) for (int d = 0; d < dim; ++d) du[d] *= domainFraction / duNorm;
        double kappa = 0;
        for (int a = 0; a < dim; ++a) { double s = 0; for (int b = 0; b < dim; ++b) s += HV[a * dim + b] * du[b]; kappa += du[a] * s; }
        const int isPD = kappa > 0.0;
        double uold[2] = {u[0], dim == 2 ? u[1] : 0.0};
        IGO_ENERGY(oldEnergy);
        for (int d = 0; d < dim; ++d) u[d] += du[d];
        if (clamp_all_boundaries(dim, u)) break;
        for (int j = 0; j < 20; ++j) {
          IGO_ENERGY(energyVal);
          if (energyVal > oldEnergy && duNorm > 1e-8) {
            const double sf = ldexp(1.0, j + 1);
            for (int d = 0; d < dim; ++d) u[d] = isPD ? uold[d] + du[d] / sf : uold[d] - du[d] / sf;
          } else
            break;
          if (j == 19) fl = 2;
        }
        if (fl) break;
        IGO_EGH();
        double rn = 0;
        for (int d = 0; d < dim; ++d) rn += R[d] * R[d];
        if (sqrt(rn) < tol) break;
      }
#undef IGO_EGH
#undef IGO_ENERGY
      (void)energyVal;
      if (fl == 2 || fl == 1) for (int d = 0; d < dim; ++d) u[d] = NAN;
      else for (int d = 0; d < dim; ++d) u[d] = u[d] < 0.0 ? 0.0 : (u[d] > 1.0 ? 1.0 : u[d]);
    }
    for (int d = 0; d < dim; ++d) xi_out[pt * dim + d] = u[d];
    if (flag) flag[pt] = fl;
  }
  free(loc); free(work);
  for (int d = 0; d < dim; ++d) free(tab[d]);
  return 0;
}

/* ---- patch-level field norms (north_star: "patch-level reductions (volume, L2/energy norms)") --------------------
 * u_h = sum_i R_i u_{g(e,i)}: result[0] = sum_q w detJ |u_h|^2, result[1] = sum_q w detJ |J^-T grad_xi u_h|^2, formed per
 * point exactly as the loop of dune/python/test/poisson.py:47-79 forms its ingredients (evaluateFunction,
 * evaluateJacobian, jacobianInverseTransposed, integrationElement).  u [ndof][ncomp], DOF numbering of igo_dofmap.   */
int igo_norms(const igo_patch* P, long elem_begin, long elem_end, const int* nq1, const double* xi1d, const double* w1d,
              const double* u, int ncomp, double* result) {
  const int dim = P->dim, dw = P->dimworld, nb = nb_of(P);
  int nel[MAXDIM]; int* tab[MAXDIM];
  igo_span_tables(P, nel, tab);
  long nq = 1; const double* xid[MAXDIM]; const double* wd[MAXDIM];
  { long off = 0; for (int d = 0; d < dim; ++d) { xid[d] = xi1d + off; wd[d] = w1d + off; off += nq1[d]; nq *= nq1[d]; } }
  long nelem_all = 1; for (int d = 0; d < dim; ++d) nelem_all *= nel[d];
  int* dof = (int*)malloc(sizeof(int) * nelem_all * nb);
  igo_dofmap(P, 0, dof);
  double* loc = (double*)malloc(sizeof(double) * nb * (dw + 1));
  double* work = (double*)malloc(sizeof(double) * 27 * nb);
  double* N = (double*)malloc(sizeof(double) * nb);
  double* dN = (double*)malloc(sizeof(double) * nb * dim);
  long double l2 = 0, en = 0;
  for (long e = elem_begin; e < elem_end; ++e) {
    int span[MAXDIM];
    element_span(P, nel, tab, e, span);
    gather_cp(P, span, loc);
    for (long q = 0; q < nq; ++q) {
      double xi[MAXDIM], wq = 1.0, detJ, JinvT[9]; long hq = q;
      for (int d = 0; d < dim; ++d) { xi[d] = xid[d][hq % nq1[d]]; wq *= wd[d][hq % nq1[d]]; hq /= nq1[d]; }
      eval_point(P, span, loc, xi, 1, N, dN, 0, 0, 0, 0, &detJ, JinvT, work);
      for (int c = 0; c < ncomp; ++c) {
        double uh = 0, gxi[MAXDIM] = {0, 0, 0};
        for (int i = 0; i < nb; ++i) {
          const double ui = u[(long)dof[e * nb + i] * ncomp + c];
          uh += N[i] * ui;
          for (int d = 0; d < dim; ++d) gxi[d] += dN[i * dim + d] * ui;
        }
        double g2 = 0;
        for (int k = 0; k < dw; ++k) { double gk = 0; for (int d = 0; d < dim; ++d) gk += JinvT[k * dim + d] * gxi[d]; g2 += gk * gk; }
        l2 += (long double)(wq * detJ * uh * uh);
        en += (long double)(wq * detJ * g2);
      }
    }
  }
  result[0] = (double)l2; result[1] = (double)en;
  free(dof); free(loc); free(work); free(N); free(dN);
  for (int d = 0; d < dim; ++d) free(tab[d]);
  return 0;
}

/* ---- a9: NurbsLocalCoefficients::init (dune/iga/nurbsbasis.hh:220-271, setup1d :162-180, setup2d :182-216) ----------
 * LocalKey (subEntity, codim, index) per local shape function; sizes_d = degree_d + 1; dim 3 leaves subEntity 0.     */
void igo_local_keys(int dim, const int* degree, unsigned* sub, unsigned* codim, unsigned* index) {
  unsigned sizes[MAXDIM] = {1, 1, 1}, nb = 1;
  for (int d = 0; d < dim; ++d) { sizes[d] = degree[d] + 1; nb *= sizes[d]; }
  for (unsigned i = 0; i < nb; ++i) {
    unsigned m[MAXDIM], r = i;
    for (int d = 0; d < dim; ++d) { m[d] = r % sizes[d]; r /= sizes[d]; }
    codim[i] = 0; index[i] = 0; sub[i] = 0;
    for (int j = 0; j < dim; ++j) if (m[j] == 0 || m[j] == sizes[j] - 1) codim[i]++;
    for (int j = dim - 1; j >= 0; --j) if (m[j] > 0 && m[j] < sizes[j] - 1) index[i] = (sizes[j] - 1) * index[i] + (m[j] - 1);
  }
  if (dim == 1) {
    if (sizes[0] == 1) return;
    unsigned last = 0;
    sub[last++] = 0;
    for (unsigned i = 0; i + 2 < sizes[0]; ++i) sub[last++] = 0;
    sub[last++] = 1;
  } else if (dim == 2 && sizes[0] > 1 && sizes[1] > 1) {
    unsigned last = 0;
    sub[last++] = 0;                                                /* corner 0 */
    for (unsigned i = 0; i + 2 < sizes[0]; ++i) sub[last++] = 2;   /* lower edge */
    sub[last++] = 1;                                                /* corner 1 */
    for (unsigned e = 0; e + 2 < sizes[1]; ++e) {
      sub[last++] = 0;                                              /* left edge */
      for (unsigned i = 0; i + 2 < sizes[0]; ++i) sub[last++] = 0; /* face */
      sub[last++] = 1;                                              /* right edge */
    }
    sub[last++] = 2;                                                /* corner 2 */
    for (unsigned i = 0; i + 2 < sizes[0]; ++i) sub[last++] = 3;   /* upper edge */
    sub[last++] = 3;                                                /* corner 3 */
  }
}

int igo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

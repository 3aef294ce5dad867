/* igab200.h -- C-ABI of the B200-native NURBS evaluation / assembly path behind dune-iga's local-finite-element and
 * geometry API.  Plain pointers and sizes only; no C++ / torch types.  Every entry point returns an igab_status
 * (0 == ok), never throws and never aborts; igab_last_error_string() explains the last failure of the calling thread.
 *
 * The reference (rath3t/dune-iga v0.1.7, header-only C++) has NO plugin / FFI layer for this path: the boundary a
 * replacement sits behind is the set of C++ methods its clients call once per quadrature point.  Each entry point
 * below names the reference interface it replaces (file:line relative to the reference root) and batches it over
 * all points of all elements of a patch.  INTEGRATION.md shows the reference-side binding (a header-only adaptor
 * keeping the reference's method names, and the pybind11 / ctypes stub for dune.iga).
 *
 * Layout contract (identical to the reference's MultiDimensionNet, dune/iga/utils/mdnet.hh:329-345):
 *   - every multi-index is flattened FIRST DIRECTION FASTEST: elements, local basis functions, quadrature points,
 *     control points;
 *   - control point k is the row (x_0 .. x_{dimworld-1}, w) -- ControlPoint{p,w}, dune/iga/controlpoint.hh:15-38;
 *   - points are numbered pt = (e - elem_begin) * nq + q for tensor rules, or as given for point lists;
 *   - outputs per point, all FP64, contiguous per point:
 *       N    [pt][nb]              R_i(xi)                       NurbsLocalBasis::evaluateFunction   nurbsbasis.hh:77-82
 *       dN   [pt][nb][dim]         dR_i/dxi_d  (reference elem.) NurbsLocalBasis::evaluateJacobian   nurbsbasis.hh:87-96
 *       d2N  [pt][nb][nh]          partial(alpha), |alpha| = 2   NurbsLocalBasis::partial            nurbsbasis.hh:99-108
 *       x    [pt][dimworld]        NURBSGeometry::global                                             nurbsgeometry.hh:133-138
 *       Jt   [pt][dim][dimworld]   NURBSGeometry::jacobianTransposed                                 nurbsgeometry.hh:182-196
 *       H    [pt][nh][dimworld]    NURBSGeometry::secondDerivativeOfPosition                         nurbsgeometry.hh:257-292
 *       detJ [pt]                  NURBSGeometry::integrationElement                                 nurbsgeometry.hh:200-203
 *       JinvT[pt][dimworld][dim]   NURBSGeometry::jacobianInverseTransposed                          nurbsgeometry.hh:205-210
 *     nb = prod(degree_d + 1); nh = dim (dim+1)/2 with the reference's row order: pure second derivatives first,
 *     then mixed (0,1),(0,2),(1,2)  (nurbsgeometry.hh:262-289).
 *   - integers (spans, DOF maps) are int32 and bit-exact with the reference.
 *
 * Memory spaces: patch data is always given in HOST memory and copied to the current CUDA device by
 * igab_patch_create.  Bulk inputs/outputs carry an igab_memspace: IGAB_DEVICE pointers are used in place on the
 * given stream (asynchronous); IGAB_HOST pointers are staged through device buffers owned by the handle, and the call
 * returns after the results have landed (synchronous).  There is no CPU fallback: without a CUDA device every
 * compute entry point returns IGAB_CUDA_ERROR.
 *
 * Threading: one host thread per patch handle at a time (the reference is threadSafe=false,
 * dune/iga/gridcapabilities.hh:53-63); different handles may be used concurrently.
 */
#ifndef IGAB200_H
#define IGAB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct igab_patch igab_patch;
typedef int igab_status;

enum {
  IGAB_OK = 0,
  IGAB_INVALID_ARGUMENT = 1,
  IGAB_OUT_OF_MEMORY = 2,
  IGAB_CUDA_ERROR = 3,
  IGAB_NCCL_ERROR = 4,
  IGAB_UNSUPPORTED = 5
};

typedef enum { IGAB_HOST = 0, IGAB_DEVICE = 1 } igab_memspace;

/* igab_eval_* kernel choice (igab_patch_set_kernel): AUTO picks a tuned tensor kernel when one is compiled for the
 * patch's (dim, dimworld, degree, rule), else the general warp-per-element kernel (run-time degrees and rule), else the
 * thread-per-point kernel.  GENERIC forces the thread-per-point kernel, WARP the warp-per-element kernel, TENSOR
 * refuses anything but a tuned kernel (tests compare the paths with each other and with the oracle). */
enum { IGAB_KERNEL_AUTO = 0, IGAB_KERNEL_GENERIC = 1, IGAB_KERNEL_TENSOR = 2, IGAB_KERNEL_WARP = 3 };

/* element trim flags, values of Dune::IGA::ElementTrimFlag (dune/iga/nurbspatch.hh:633-647) */
enum { IGAB_TRIM_FULL = 0, IGAB_TRIM_EMPTY = 1, IGAB_TRIM_TRIMMED = 2 };

/* assembly kinds for igab_assemble */
enum { IGAB_ASM_POISSON = 0, IGAB_ASM_MASS = 1, IGAB_ASM_ELASTICITY = 2 };

/* Output set of an evaluation; any member may be NULL (not computed / not stored). */
typedef struct igab_eval_out {
  double* N;
  double* dN;
  double* d2N;
  double* x;
  double* Jt;
  double* H;
  double* detJ;
  double* JinvT;
} igab_eval_out;

/* ---- library ------------------------------------------------------------------------------------------------- */
const char* igab_last_error_string(void);
const char* igab_version(void);
/* number of CUDA kernels this library has launched in this process (for benchmarks' gpu_launches) */
int64_t igab_launch_count(void);
/* pinned host memory for IGAB_HOST buffers that should transfer at full PCIe rate; placed on the NUMA node of the
 * CURRENT CUDA device (one process per GPU on a multi-socket box: every rank's buffers next to its own GPU) */
igab_status igab_host_alloc(void** ptr, uint64_t bytes);
igab_status igab_host_free(void* ptr);
/* NUMA node the current CUDA device hangs off (sysfs), -1 when unknown */
int igab_device_numa_node(void);

/* ---- patch handle -------------------------------------------------------------------------------------------------
 * Replaces: NURBSPatchData (dune/iga/nurbspatchdata.hh:17-36) + the per-element state NURBSPatch / NURBSGeometry /
 * NurbsLocalFiniteElement cache: unique knots (nurbspatch.hh:58-66), element->span (nurbspatch.hh:248-253,
 * nurbsbasis.hh:355-360), scaling/offset (nurbsgeometry.hh:64-68, nurbsbasis.hh:363-365).
 * knots[d] has nknots[d] entries (open knot vector, degree[d]+1 repeated end knots, else IGAB_INVALID_ARGUMENT);
 * cp has prod(nknots[d]-degree[d]-1) rows of dimworld+1 doubles.  trimflags: NULL or one byte per element.
 * Supported: dim 1..3, dim <= dimworld <= 3, degree 1..8 per direction.                                          */
igab_status igab_patch_create(int dim, int dimworld, const int* degree, const int* nknots, const double* const* knots,
                              const double* cp, const uint8_t* trimflags, igab_patch** out);
void igab_patch_destroy(igab_patch* p);
/* n_elem total and per direction (non-empty knot spans); nb = local basis size; ndof_untrimmed = prod n_d */
igab_status igab_patch_sizes(const igab_patch* p, int64_t* n_elem, int32_t* n_elem_dir, int32_t* nb,
                             int64_t* ndof_untrimmed);
igab_status igab_patch_set_kernel(igab_patch* p, int kernel);
/* replace the control net (same sizes) without rebuilding the handle, e.g. after a geometry update */
igab_status igab_patch_update_control_points(igab_patch* p, const double* cp);

/* Replaces: NURBSPatch::spanAndDirectionFromDirectIndex<0> (nurbspatch.hh:248-253) / NurbsLocalFiniteElement::bind
 * (nurbsbasis.hh:355-365) for all elements.  span [n_elem][dim], bit-exact.                                      */
igab_status igab_patch_spans(igab_patch* p, int32_t* span, igab_memspace space, void* stream);

/* Replaces: NurbsPreBasis::createOriginalNodeIndices + prepareForTrim + indices (nurbsbasis.hh:552-615).
 * dof [n_elem][nb]; entries of IGAB_TRIM_EMPTY elements whose DOF was dropped are -1; *ndof = size of the basis.   */
igab_status igab_patch_dofmap(igab_patch* p, int32_t* dof, int64_t* ndof, igab_memspace space, void* stream);

/* Replaces NURBSPatch::getDirectIndex<0> / getRealIndex<0> (dune/iga/nurbspatch.hh:599-631): the reference numbers the
 * elements of a trimmed patch by their REAL index -- position among the non-empty (full or trimmed) elements in
 * ascending direct index (elementIndexMap, nurbspatch.hh:696-716) -- and this library by the DIRECT index (flat index of
 * the knot span).  direct_of_real [*n_real] = getDirectIndex<0>(r); real_of_direct [n_elem] = getRealIndex<0>(e), -1
 * where the reference throws ("No corresponding realIndex": IGAB_TRIM_EMPTY elements).  Identity without trim flags.
 * An ordered prefix sum over the flags on the device; either array may be NULL; int32, bit-exact.                    */
igab_status igab_patch_real_index_maps(igab_patch* p, int32_t* direct_of_real, int32_t* real_of_direct, int64_t* n_real,
                                       igab_memspace space, void* stream);

/* Replaces NurbsLocalCoefficients::init (nurbsbasis.hh:220-271): the LocalKey (subEntity, codim, index) of each of the
 * nb = prod(degree_d + 1) local shape functions, identical for every element.  codim = number of directions in which the
 * function sits on the element boundary; index = position among the functions of the same sub-entity; subEntity as
 * setup1d / setup2d number it (:162-216) -- like the reference, 0 for dim 3.  Pure integer host code, no device needed. */
igab_status igab_local_keys(int dim, const int* degree, uint32_t* sub_entity, uint32_t* codim, uint32_t* index);

/* ---- batched evaluation -------------------------------------------------------------------------------------------
 * Replaces, for every point of every element in [elem_begin, elem_end) at the tensor rule xi1d[d][0..nq1[d]):
 * the user quadrature loop calling evaluateFunction / evaluateJacobian / partial (nurbsbasis.hh:77-108) and
 * global / jacobianTransposed / integrationElement / jacobianInverseTransposed / secondDerivativeOfPosition
 * (nurbsgeometry.hh:133-138,182-210,257-292); element iteration nurbsleafgridview.hh:147-155; rule
 * nurbsgridentity.hh:120-141.  xi1d are HOST arrays in [0,1] (element-local).  deriv_order 0, 1 or 2 bounds what
 * is computed (dN/Jt/detJ/JinvT need >= 1, d2N/H need 2).                                                        */
igab_status igab_eval_tensor(igab_patch* p, int64_t elem_begin, int64_t elem_end, int deriv_order, const int* nq1,
                             const double* const* xi1d, const igab_eval_out* out, igab_memspace out_space,
                             void* stream);

/* The same with HOST outputs, without waiting: returns once the kernels and the device-to-host copies are queued, so the
 * caller's loop over the elements it already holds (the per-element calls of the reference, nurbsbasis.hh:77-108,
 * nurbsgeometry.hh:133-203) overlaps the evaluation and the transfer of the NEXT block of elements.  The output
 * arrays (pinned memory from igab_host_alloc for a truly asynchronous copy) may be read after igab_patch_wait; any
 * other call on the handle waits for a pending begin first.                                                         */
igab_status igab_eval_tensor_begin(igab_patch* p, int64_t elem_begin, int64_t elem_end, int deriv_order, const int* nq1,
                                   const double* const* xi1d, const igab_eval_out* out, void* stream);
igab_status igab_patch_wait(igab_patch* p);

/* Same at arbitrary element-local points (trimmed-element rules, utils/fillquadraturerule.hh:8-19; probes):
 * elem[npts] element direct index, xi[npts][dim]; inputs and outputs all in `space`.                             */
igab_status igab_eval_points(igab_patch* p, int64_t npts, const int32_t* elem, const double* xi, int deriv_order,
                             const igab_eval_out* out, igab_memspace space, void* stream);

/* Replaces NurbsLocalBasis::partial for an arbitrary multi-index alpha (nurbsbasis.hh:99-108,667-677):
 * out[npts][nb] = prod_d scale_d^alpha_d * d^alpha R_i.  prod(alpha_d+1) <= 64.                                    */
igab_status igab_partial(igab_patch* p, int64_t npts, const int32_t* elem, const double* xi, const int* alpha,
                         double* out, igab_memspace space, void* stream);

/* ---- element matrices ---------------------------------------------------------------------------------------------
 * Replaces the Python localAssembler loop, dune/python/test/poisson.py:37-81, generalised to B^T D B w detJ:
 *   IGAB_ASM_POISSON    Ke[i][j] = sum_q w detJ grad R_i . grad R_j ; fe[i] = sum_q w detJ R_i fconst
 *   IGAB_ASM_MASS       Ke[i][j] = sum_q w detJ R_i R_j
 *   IGAB_ASM_ELASTICITY B = strain operator (Voigt xx,yy,xy | xx,yy,zz,yz,xz,xy), D row-major ns x ns (HOST),
 *                       DOFs flat-interleaved i*dim+c (dune-functions PowerPreBasis, basis/__init__.py:53-67)
 * Ke [n_el][n][n] row-major, n = nb or nb*dim; fe [n_el][n] or NULL.  w1d HOST weights of the tensor rule.        */
igab_status igab_assemble(igab_patch* p, int kind, const double* D, double fconst, int64_t elem_begin,
                          int64_t elem_end, const int* nq1, const double* const* xi1d, const double* const* w1d,
                          double* Ke, double* fe, igab_memspace out_space, void* stream);

/* Element matrices of elements that carry their OWN quadrature: trimmed elements, whose rule is assembled from the
 * sub-triangles of the trimmed domain (utils/fillquadraturerule.hh:8-19, NURBSGridEntity::getQuadratureRule
 * nurbsgridentity.hh:120-141).  List element k is patch element elem[k] with the points offset[k] .. offset[k+1] of
 * xi [npts][2] (element-local) and w [npts]; same integrands and layouts as igab_assemble; Ke [nlist][n][n],
 * fe [nlist][n] or NULL.  2-D patches (trimming is 2-D in the reference) with equal degrees 1..3.  elem (int32),
 * offset (int64, nlist + 1 entries), xi, w, Ke, fe live in `space`; D is HOST.                                     */
igab_status igab_assemble_points(igab_patch* p, int kind, const double* D, double fconst, int64_t nlist,
                                 const int32_t* elem, const int64_t* offset, const double* xi, const double* w,
                                 double* Ke, double* fe, igab_memspace space, void* stream);

/* Element load vectors with a source given PER QUADRATURE POINT: fe[k][i*ncomp + c] = sum_q w_q detJ_q R_i(xi_q) f[pt][c]
 * -- localB[i] += pt.weight * integrationElement * phi[i] * f(posGlobal), dune/python/test/poisson.py:79, with the
 * caller's f already evaluated at the points (x from igab_eval_tensor / igab_eval_points; a C-ABI cannot take the
 * Python callable).  igab_load_vector: tensor rule over [elem_begin, elem_end), f [(elem_end - elem_begin) * nq][ncomp],
 * fe [n_el][nb * ncomp].  igab_load_vector_points: elements with their own rules (trimmed elements), same list format
 * as igab_assemble_points with f [npts][ncomp] indexed like xi / w; any dim.  f and fe (and the lists) live in `space`;
 * points are summed in rule order.  ncomp 1..3.                                                                     */
igab_status igab_load_vector(igab_patch* p, int64_t elem_begin, int64_t elem_end, const int* nq1,
                             const double* const* xi1d, const double* const* w1d, const double* f, int ncomp, double* fe,
                             igab_memspace space, void* stream);
igab_status igab_load_vector_points(igab_patch* p, int64_t nlist, const int32_t* elem, const int64_t* offset,
                                    const double* xi, const double* w, const double* f, int ncomp, double* fe,
                                    igab_memspace space, void* stream);

/* ---- refinement of the control net -----------------------------------------------------------------------------------
 * Replaces the control-net part of knotRefinement / degreeElevate (dune/iga/nurbsalgorithms.hh:264-528) for one
 * direction: new net = T applied along `direction` in homogeneous coordinates.  T (n_new x ncp_old[direction]) is given
 * row-banded: row i has len[i] <= maxlen coefficients coef[i*maxlen + k] for the old indices first[i] + k.  All arrays
 * HOST; cp_in has prod(ncp_old) rows, cp_out prod(ncp_new) rows of dimworld+1 doubles, first direction fastest.     */
igab_status igab_refine_apply(int dim, int dimworld, const int* ncp_old, int direction, int n_new, int maxlen,
                              const int32_t* first, const int32_t* len, const double* coef, const double* cp_in,
                              double* cp_out);

/* A CHAIN of such operators -- degree elevation (degreeElevate, dune/iga/nurbsalgorithms.hh:264-439) and / or knot
 * insertion of several directions, as NURBSGrid::globalDegreeElevate / globalRefine apply them one direction after the other
 * (dune/iga/nurbsgrid.hh:130-178) -- with every intermediate net in HBM: operator o (row-banded as above, HOST arrays) acts
 * along direction[o] on the net the previous operators produced.  cp_in [prod ncp_old][dimworld+1] in `in_space`, cp_out
 * [prod n_final][dimworld+1] in `out_space`; with IGAB_DEVICE on both sides and igab_patch_create_from_device only the
 * operators (n_new x maxlen doubles each) cross PCIe.  nops = 0 copies the net.  Synchronises `stream` before returning. */
igab_status igab_refine_operators(int dim, int dimworld, const int* ncp_old, int nops, const int* direction,
                                  const int* n_new, const int* maxlen, const int32_t* const* first,
                                  const int32_t* const* len, const double* const* coef, const double* cp_in,
                                  igab_memspace in_space, double* cp_out, igab_memspace out_space, void* stream);

/* Knot refinement of ALL directions without leaving HBM (the chain NURBSGrid::globalRefine -> knotRefinement per direction,
 * dune/iga/nurbsgrid.hh:130-165 -> dune/iga/nurbsalgorithms.hh:441-528): for every direction whose knot vector grows, the
 * knot-insertion operator is built on the device (Oslo recursion, one thread per new control-point row) and applied to
 * the net in homogeneous coordinates; intermediate nets stay on the device.  knots_* are HOST arrays (knots_new[d] must
 * contain knots_old[d]); cp_in [prod n_old][dimworld+1] in `in_space`, cp_out [prod n_new][dimworld+1] in `out_space` --
 * with IGAB_DEVICE on both sides and igab_patch_create_from_device the refined net never crosses PCIe.               */
igab_status igab_refine_knots(int dim, int dimworld, const int* degree, const int* nknots_old,
                              const double* const* knots_old, const int* nknots_new, const double* const* knots_new,
                              const double* cp_in, igab_memspace in_space, double* cp_out, igab_memspace out_space,
                              void* stream);
/* igab_patch_create with the control net already on the current device (copied device to device) */
igab_status igab_patch_create_from_device(int dim, int dimworld, const int* degree, const int* nknots,
                                          const double* const* knots, const double* cp_dev, const uint8_t* trimflags,
                                          igab_patch** out);

/* ---- element faces (boundary / intersection integrals) -----------------------------------------------------------------
 * Replaces, for nfaces faces x the tensor rule of dimension dim-1: NURBSintersection::geometry().global /
 * integrationElement and outerNormal / unitOuterNormal / integrationOuterNormal (dune/iga/nurbsintersection.hh:86-160;
 * sub-entity geometry nurbsgeometry.hh:294-303).  face[k] = 2*d + s (DUNE cube numbering: xi_d fixed to s) of element
 * elem[k]; the rule's coordinates fill the free directions in increasing order, first one fastest.
 * Outputs per point (nullable): x [pt][dimworld], normal [pt][dimworld] = UNIT outer normal, dS [pt] = integration
 * element of the face; integrationOuterNormal = normal * dS.  dim 2 or 3.  elem / face / outputs live in `space`. */
igab_status igab_eval_faces(igab_patch* p, int64_t nfaces, const int32_t* elem, const int32_t* face, const int* nqf1,
                            const double* const* xi1d, double* x, double* normal, double* dS, igab_memspace space,
                            void* stream);

/* ---- differential geometry of surfaces ---------------------------------------------------------------------------------
 * Replaces, batched over npts points, NURBSGeometry::normal / unitNormal (dune/iga/nurbsgeometry.hh:216-228), metric
 * (:238-243), secondFundamentalForm (:245-255) and gaussianCurvature (:230-236).  Inputs are the arrays written by
 * igab_eval_tensor / igab_eval_points: Jt [npts][dim][dimworld], H [npts][dim(dim+1)/2][dimworld] (rows uu, vv, uv;
 * needed for second_ff / gauss_k only).  Outputs (nullable): normal [npts][3] = J0 x J1, unit_normal [npts][3],
 * metric [npts][dim][dim] = J J^T, second_ff [npts][2][2] = H_ab . n, gauss_k [npts] = det(second_ff) / det(metric).
 * metric works for every (dim, dimworld); the others need dim 2, dimworld 3 as in the reference (IGAB_UNSUPPORTED
 * otherwise).  All arrays live in `space`.                                                                          */
igab_status igab_surface_forms(int dim, int dimworld, int64_t npts, const double* Jt, const double* H, double* normal,
                               double* unit_normal, double* metric, double* second_ff, double* gauss_k,
                               igab_memspace space, void* stream);

/* ---- inverse of the element map ---------------------------------------------------------------------------------------
 * Replaces NURBSGeometry::local (dune/iga/nurbsgeometry.hh:145-176), batched: for point k, the element-local
 * coordinates xi[k][dim] in [0,1]^dim of the point of element elem[k] that maps to (dim == dimworld: Newton) or lies
 * closest to (curves / surfaces in a larger space: closestPointProjectionByTrustRegion,
 * dune/iga/geometry/closestpointprojection.hh) xglobal[k][dimworld].  As in the reference the iterate is clamped to
 * the element, and a singular Jacobian returns DBL_MAX in every coordinate.  flag[k] (nullable): 0 ok, 1 singular
 * matrix, 2 no energy-decreasing direction (the reference throws Dune::MathError; xi = NaN), 3 Newton iteration cap
 * (100; the reference has none).  All arrays live in `space`.                                                       */
igab_status igab_geometry_local(igab_patch* p, int64_t npts, const int32_t* elem, const double* xglobal, double* xi,
                                int32_t* flag, igab_memspace space, void* stream);

/* ---- global sparse matrix (CSR) ---------------------------------------------------------------------------------------
 * Replaces the scatter loop of globalAssembler, dune/python/test/poisson.py:113-125 (A[gi,gj] += localA[i,j] through
 * localView.index), for the tensor-product DOF numbering of NurbsPreBasis (nurbsbasis.hh:572-584): row g couples
 * with the DOFs within +-degree in every direction.  ncomp > 1: vector-valued basis, flat-interleaved numbering
 * g*ncomp + c (dune-functions PowerPreBasis, python/dune/iga/basis/__init__.py:53-67).  Patches created with trim
 * flags use the compacted numbering of prepareForTrim (nurbsbasis.hh:599-615, the one igab_patch_dofmap returns):
 * rows / columns are the DOFs touched by a non-empty element, a row keeps exactly the columns that share a non-empty
 * element with it (the pattern the scatter loop produces), IGAB_TRIM_EMPTY elements of Ke / fe are never read, and
 * the caller puts the matrices of IGAB_TRIM_TRIMMED elements (igab_assemble_points) into their slots of Ke.
 *   igab_csr_pattern : *nrows = ndof*ncomp, *nnz; rowptr[nrows+1] (int64) and colidx[nnz] (int32) may be NULL to
 *                      query the sizes first.  Columns ascending within a row.
 *   igab_csr_assemble: values[nnz] (+)= sum over the elements [elem_begin, elem_end) of Ke[e - elem_begin][i][j];
 *                      a GATHER per non-zero in a fixed element order: no atomics, bitwise reproducible.  Rows of
 *                      DOFs that other element ranges also touch receive only this range's share -- ranks exchange
 *                      and add those interface rows (contiguous value ranges, see dune-iga_b200/partition.py).       */
igab_status igab_csr_pattern(igab_patch* p, int ncomp, int64_t* nrows, int64_t* nnz, int64_t* rowptr, int32_t* colidx,
                             igab_memspace space, void* stream);
igab_status igab_csr_assemble(igab_patch* p, int ncomp, int64_t elem_begin, int64_t elem_end, const double* Ke,
                              const int64_t* rowptr, double* values, int accumulate, igab_memspace space,
                              void* stream);

/* Global load vector, same gather: b[nrows] (+)= sum over the elements [elem_begin, elem_end) of fe[e - elem_begin][i]
 * (poisson.py:121, b[gi] += localb[i]).                                                                            */
igab_status igab_rhs_assemble(igab_patch* p, int ncomp, int64_t elem_begin, int64_t elem_end, const double* fe, double* b,
                              int accumulate, igab_memspace space, void* stream);
/* Dirichlet rows as the reference's assembler sets them (poisson.py:91-93,117-119): for every row with mask[row] != 0
 * the row of `values` becomes the identity row (columns are left alone, as in the reference) and, when b is given,
 * b[row] = g[row] (0 when g is NULL).  mask [nrows] bytes, g / b [nrows], values [nnz] -- all in `space`.          */
igab_status igab_csr_dirichlet(igab_patch* p, int ncomp, const uint8_t* mask, const double* g, double* values, double* b,
                               igab_memspace space, void* stream);

/* The whole globalAssembler of dune/python/test/poisson.py:84-127 in ONE call: element matrices (igab_assemble) in
 * device-resident chunks, the trimmed elements' matrices from their own rules (igab_assemble_points; nlist may be 0)
 * put into their slots, CSR gather and load-vector gather per chunk -- the element matrices never leave the device.
 * values [nnz] (pattern of igab_csr_pattern with ncomp = 1, or dim for IGAB_ASM_ELASTICITY) and b [nrows] (or NULL)
 * live in `space`; the rule, D and the trimmed lists (elem strictly ascending, offset absolute into xi / w) are HOST.
 * Only the elements [elem_begin, elem_end) contribute (list entries outside are ignored): a rank owning a knot-span
 * slab assembles its share of the one shared pattern, rows of DOFs on a cut are completed by the interface-row exchange
 * (dune-iga_b200/partition.py).  Dirichlet rows are applied afterwards with igab_csr_dirichlet, as the reference does
 * inside its loop.                                                                                                  */
igab_status igab_global_assemble(igab_patch* p, int kind, const double* D, double fconst, int64_t elem_begin,
                                 int64_t elem_end, const int* nq1, const double* const* xi1d,
                                 const double* const* w1d, int64_t nlist,
                                 const int32_t* elem, const int64_t* offset, const double* xi, const double* w,
                                 double* values, double* b, igab_memspace space, void* stream);

/* The load-vector half of globalAssembler (poisson.py:79,121) with a per-point source, in one call: element vectors of
 * [elem_begin, elem_end) from the tensor rule (f [(elem_end - elem_begin) * nq][ncomp]) in device-resident chunks, those
 * of the listed trimmed elements from their own rules (f_list [npts][ncomp]; elem / offset HOST arrays as in
 * igab_global_assemble), gathered into b [nrows] of igab_csr_pattern(ncomp).  f, f_list, xi, w and b live in `space`.
 * Dirichlet values g go in afterwards through igab_csr_dirichlet (poisson.py:91-93,117-119).                         */
igab_status igab_global_load_vector(igab_patch* p, int ncomp, int64_t elem_begin, int64_t elem_end, const int* nq1,
                                    const double* const* xi1d, const double* const* w1d, const double* f, int64_t nlist,
                                    const int32_t* elem, const int64_t* offset, const double* xi, const double* w,
                                    const double* f_list, double* b, igab_memspace space, void* stream);

/* ---- patch reductions ---------------------------------------------------------------------------------------------
 * Replaces the user loop summing NURBSGeometry::volume() over elements (nurbsgeometry.hh:96-113,
 * dune/iga/test/gridtests.cpp:338-341): result = sum_{e in range} sum_q detJ w.  Deterministic two-level tree sum.
 * nccl_comm: NULL, or an ncclComm_t over the ranks sharing the patch -- then *result is the all-reduced sum
 * (ncclAllReduce, ncclDouble, ncclSum) and is identical on all ranks.  *result is a HOST scalar.                */
igab_status igab_reduce_volume(igab_patch* p, int64_t elem_begin, int64_t elem_end, const int* nq1,
                               const double* const* xi1d, const double* const* w1d, double* result, void* nccl_comm,
                               void* stream);

/* The same sum on a TRIMMED patch (NURBSGeometry::volume of a trimmed element integrates its sub-triangle rule,
 * nurbsgeometry.hh:96-113 with fillquadraturerule.hh:8-19): the tensor rule over the IGAB_TRIM_FULL elements of
 * [elem_begin, elem_end) (trimmed and empty elements are skipped whenever the patch has trim flags -- also by
 * igab_reduce_volume) plus sum_k detJ(elem[k], xi[k]) w[k] over the trimmed elements' own points: elem [npts_list]
 * (int32), xi [npts_list][dim] element-local, w [npts_list], HOST arrays (quadrature.trimmedBatch).                */
igab_status igab_reduce_volume_trimmed(igab_patch* p, int64_t elem_begin, int64_t elem_end, const int* nq1,
                                       const double* const* xi1d, const double* const* w1d, int64_t npts_list,
                                       const int32_t* elem, const double* xi, const double* w, double* result,
                                       void* nccl_comm, void* stream);

/* L2 norm and energy (H1 seminorm) of a discrete field u_h = sum_i R_i u_i, fused evaluation + reduction:
 *   result[0] = sum_q w detJ |u_h|^2 (= u^T M u),  result[1] = sum_q w detJ |grad u_h|^2 (= u^T K u of the Poisson
 * operator), with R_i / dR_i as evaluateFunction / evaluateJacobian (nurbsbasis.hh:77-96), detJ = integrationElement and
 * grad = jacobianInverseTransposed * d/dxi (nurbsgeometry.hh:200-210) -- the per-point quantities of the loop
 * dune/python/test/poisson.py:47-79.  u [ndof*ncomp] in `u_space`, indexed g*ncomp + c with the DOF numbering of
 * igab_patch_dofmap, ncomp <= 3.  result[2] is HOST memory.
 * Deterministic (fixed two-level tree); nccl_comm as in igab_reduce_volume (both sums all-reduced).                   */
igab_status igab_reduce_norms(igab_patch* p, int64_t elem_begin, int64_t elem_end, const int* nq1,
                              const double* const* xi1d, const double* const* w1d, const double* u, int ncomp,
                              igab_memspace u_space, double* result, void* nccl_comm, void* stream);

/* The same on a TRIMMED patch: u [ndof*ncomp] is numbered like the trimmed basis (the compacted DOF map of
 * igab_patch_dofmap); the tensor rule runs over the IGAB_TRIM_FULL elements of the range (also in igab_reduce_norms when
 * the patch has trim flags) and the trimmed elements' own points are added: elem [npts_list], xi [npts_list][dim]
 * element-local, w [npts_list], HOST arrays (quadrature.trimmedBatch).                                             */
igab_status igab_reduce_norms_trimmed(igab_patch* p, int64_t elem_begin, int64_t elem_end, const int* nq1,
                                      const double* const* xi1d, const double* const* w1d, const double* u, int ncomp,
                                      igab_memspace u_space, int64_t npts_list, const int32_t* elem, const double* xi,
                                      const double* w, double* result, void* nccl_comm, void* stream);

/* ---- multi-GPU: element slabs, communicator, interface rows ------------------------------------------------------------
 * One process per GPU; every rank holds the whole patch handle (knots and control net replicated) and owns a contiguous
 * slab of elements: whole layers of the slowest direction, so that its elements (first direction fastest,
 * dune/iga/utils/mdnet.hh:245-257), its outputs and the rows it touches are each ONE contiguous range.  Evaluation and
 * element matrices need no collective; the reductions above take the communicator; the sharded global matrix needs the
 * exchange below.  The reference is sequential -- the place a DUNE caller hooks this in is the grid view's
 * communicate() (dune/iga/nurbsleafgridview.hh:163-165).
 *
 * igab_partition_slab: [elem_begin, elem_end) of `rank`: ceil / floor split of the n_elem_dir[dim-1] layers.  Host only. */
igab_status igab_partition_slab(int dim, const int32_t* n_elem_dir, int world, int rank, int64_t* elem_begin,
                                int64_t* elem_end);

/* Trimmed patches: elements carry unequal work (empty ones none, trimmed ones their own rules,
 * dune/iga/utils/fillquadraturerule.hh:8-19), so the slabs are balanced by the cumulative number of quadrature points
 * instead of the element count.  counts[n_elem] = points of every element in direct-index order; bounds[world + 1]:
 * rank r owns [bounds[r], bounds[r+1]), boundary r placed where the running point count first reaches r/world of the
 * total and rounded to a multiple of per_layer (the elements of one layer of the slowest direction: what
 * igab_interface_row_ranges / igab_exchange_interface_rows take as elem_bounds; 1 = no rounding).  Host only.         */
igab_status igab_partition_weighted(int64_t n_elem, const int64_t* counts, int world, int64_t per_layer,
                                    int64_t* bounds);

/* Rows of the global system (untrimmed numbering g*ncomp + c) that the elements of every rank touch: the DOF layers
 * span(l) - p .. span(l') of the slab's first / last element layer (g_d = s_d - p_d + l_d, dune/iga/nurbsbasis.hh:576).
 * elem_bounds: NULL (igab_partition_slab) or world + 1 ascending multiples of the layer size.  Rows shared by ranks r
 * and s are the intersection of their ranges; with slabs thinner than the degree that includes ranks that are not
 * neighbours.  row_begin / row_end [world].  Pure host integer code: needs no device.                                */
igab_status igab_interface_row_ranges(int dim, const int* degree, const int* nknots, const double* const* knots,
                                      int ncomp, int world, const int64_t* elem_bounds, int64_t* row_begin,
                                      int64_t* row_end);

/* NCCL communicator without the caller touching NCCL (resolved from the process with dlsym; libnccl.so.2 otherwise):
 * rank 0 fills a 128-byte id and distributes it by whatever the host code has (MPI_Bcast in a DUNE program, a file, a
 * torch.distributed broadcast); every rank then calls igab_comm_create with the CUDA device it drives current.  The
 * result is an ncclComm_t and may equally be one the caller created itself.                                         */
igab_status igab_comm_unique_id(void* id128);
igab_status igab_comm_create(int rank, int world, const void* id128, void** comm);
igab_status igab_comm_destroy(void* comm);
igab_status igab_comm_info(void* comm, int* rank, int* world);

/* Completes the rows of the sharded global matrix and load vector: every rank has gathered the element matrices of ITS
 * slab into the one shared pattern (igab_global_assemble / igab_csr_assemble / igab_rhs_assemble with its element
 * range), so rows of DOFs that elements of several ranks contain hold partial sums (poisson.py:113-125 split over
 * ranks).  For every pair of ranks with intersecting row ranges the contiguous value range of the shared rows is sent
 * both ways (ncclSend / ncclRecv in one group on `stream`), then one kernel adds the contributions in ASCENDING RANK
 * ORDER: afterwards every rank holds the complete rows it touches, bit-identical on all ranks that hold them.  values
 * [nnz] of igab_csr_pattern(ncomp) and / or b [nrows], in `space` (host arrays are staged piecewise through the
 * device); rank and size are the communicator's.  Trimmed patches use the compacted numbering.  elem_bounds as above
 * and the same on every rank.  Asynchronous on `stream` for device arrays.                                           */
igab_status igab_exchange_interface_rows(igab_patch* p, int ncomp, const int64_t* elem_bounds, double* values, double* b,
                                         igab_memspace space, void* nccl_comm, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IGAB200_H */

"""ctypes binding of dune-iga_b200/libigab200.so (the C-ABI of include/igab200.h) and a thin host-side mirror of the
reference's batched objects.  This is what a dune.iga Python caller would use in place of the per-point pybind11
calls of dune/python/test/poisson.py:47-68 -- numpy in / numpy out, or device pointers in place.

There is NO fallback: if the shared library is missing or no CUDA device is present, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("IGAB200_SO") or os.path.join(_HERE, "libigab200.so")  # the override is for instrumented builds
_lib = None

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
HOST, DEVICE = 0, 1
OK, INVALID_ARGUMENT, OUT_OF_MEMORY, CUDA_ERROR, NCCL_ERROR, UNSUPPORTED = range(6)
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_TENSOR, KERNEL_WARP = 0, 1, 2, 3
ASM_POISSON, ASM_MASS, ASM_ELASTICITY = 0, 1, 2
FIELDS = ("N", "dN", "d2N", "x", "Jt", "H", "detJ", "JinvT")
EXPORTS = ("igab_last_error_string", "igab_version", "igab_launch_count", "igab_host_alloc", "igab_host_free",
           "igab_patch_create", "igab_patch_destroy", "igab_patch_sizes", "igab_patch_set_kernel",
           "igab_patch_update_control_points", "igab_patch_spans", "igab_patch_dofmap", "igab_eval_tensor",
           "igab_eval_points", "igab_partial", "igab_assemble", "igab_reduce_volume", "igab_csr_pattern",
           "igab_csr_assemble", "igab_eval_faces", "igab_refine_apply", "igab_surface_forms", "igab_geometry_local", "igab_reduce_norms", "igab_local_keys", "igab_rhs_assemble",
           "igab_csr_dirichlet", "igab_assemble_points", "igab_global_assemble", "igab_reduce_volume_trimmed", "igab_reduce_norms_trimmed",
           "igab_partition_slab", "igab_partition_weighted", "igab_interface_row_ranges", "igab_comm_unique_id", "igab_comm_create",
           "igab_comm_destroy", "igab_comm_info", "igab_exchange_interface_rows", "igab_patch_real_index_maps",
           "igab_load_vector", "igab_load_vector_points", "igab_global_load_vector", "igab_device_numa_node", "igab_eval_tensor_begin", "igab_patch_wait", "igab_refine_knots",
           "igab_patch_create_from_device", "igab_refine_operators")


class EvalOut(C.Structure):
    _fields_ = [(f, C.c_void_p) for f in FIELDS]


class IgabError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("igab status %d: %s" % (status, msg))
        self.status = status


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError("%s not built (run __graft_entry__.build()); there is no CPU fallback" % SO_PATH)
        L = C.CDLL(SO_PATH)
        L.igab_last_error_string.restype = C.c_char_p
        L.igab_version.restype = C.c_char_p
        L.igab_launch_count.restype = C.c_int64
        L.igab_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_uint64]
        L.igab_host_free.argtypes = [C.c_void_p]
        L.igab_patch_create.argtypes = [C.c_int, C.c_int, ip, ip, C.POINTER(dp), dp, C.POINTER(C.c_ubyte),
                                        C.POINTER(C.c_void_p)]
        L.igab_patch_create_from_device.argtypes = [C.c_int, C.c_int, ip, ip, C.POINTER(dp), C.c_void_p,
                                                    C.POINTER(C.c_ubyte), C.POINTER(C.c_void_p)]
        L.igab_refine_knots.argtypes = [C.c_int, C.c_int, ip, ip, C.POINTER(dp), ip, C.POINTER(dp), C.c_void_p, C.c_int,
                                        C.c_void_p, C.c_int, C.c_void_p]
        L.igab_refine_operators.argtypes = [C.c_int, C.c_int, ip, C.c_int, ip, ip, ip, C.POINTER(C.c_void_p),
                                            C.POINTER(C.c_void_p), C.POINTER(dp), C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                            C.c_void_p]
        L.igab_patch_destroy.argtypes = [C.c_void_p]
        L.igab_patch_destroy.restype = None
        L.igab_patch_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                       C.POINTER(C.c_int64)]
        L.igab_patch_set_kernel.argtypes = [C.c_void_p, C.c_int]
        L.igab_patch_update_control_points.argtypes = [C.c_void_p, dp]
        L.igab_patch_spans.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.igab_patch_dofmap.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_void_p]
        L.igab_eval_tensor.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, ip, C.POINTER(dp),
                                       C.POINTER(EvalOut), C.c_int, C.c_void_p]
        L.igab_eval_tensor_begin.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, ip, C.POINTER(dp),
                                             C.POINTER(EvalOut), C.c_void_p]
        L.igab_patch_wait.argtypes = [C.c_void_p]
        L.igab_eval_points.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(EvalOut),
                                       C.c_int, C.c_void_p]
        L.igab_partial.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, ip, C.c_void_p, C.c_int, C.c_void_p]
        L.igab_assemble.argtypes = [C.c_void_p, C.c_int, dp, C.c_double, C.c_int64, C.c_int64, ip, C.POINTER(dp),
                                    C.POINTER(dp), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.igab_csr_pattern.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p,
                                       C.c_void_p, C.c_int, C.c_void_p]
        L.igab_reduce_norms_trimmed.argtypes = [C.c_void_p, C.c_int64, C.c_int64, ip, C.POINTER(dp), C.POINTER(dp),
                                                C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.POINTER(C.c_double), C.c_void_p, C.c_void_p]
        L.igab_reduce_volume_trimmed.argtypes = [C.c_void_p, C.c_int64, C.c_int64, ip, C.POINTER(dp), C.POINTER(dp),
                                                 C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.POINTER(C.c_double), C.c_void_p, C.c_void_p]
        L.igab_global_assemble.argtypes = [C.c_void_p, C.c_int, dp, C.c_double, C.c_int64, C.c_int64, ip, C.POINTER(dp), C.POINTER(dp),
                                           C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int, C.c_void_p]
        L.igab_csr_assemble.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_int, C.c_void_p]
        L.igab_eval_faces.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, ip, C.POINTER(dp), C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.igab_surface_forms.argtypes = [C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 7 + [C.c_int, C.c_void_p]
        L.igab_geometry_local.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_void_p]
        L.igab_reduce_norms.argtypes = [C.c_void_p, C.c_int64, C.c_int64, ip, C.POINTER(dp), C.POINTER(dp), C.c_void_p,
                                        C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p, C.c_void_p]
        L.igab_rhs_assemble.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                        C.c_void_p]
        L.igab_csr_dirichlet.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_void_p]
        L.igab_assemble_points.argtypes = [C.c_void_p, C.c_int, dp, C.c_double, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.igab_local_keys.argtypes = [C.c_int, ip, C.c_void_p, C.c_void_p, C.c_void_p]
        L.igab_refine_apply.argtypes = [C.c_int, C.c_int, ip, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, dp, dp,
                                        dp]
        L.igab_reduce_volume.argtypes = [C.c_void_p, C.c_int64, C.c_int64, ip, C.POINTER(dp), C.POINTER(dp),
                                         C.POINTER(C.c_double), C.c_void_p, C.c_void_p]
        L.igab_partition_slab.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.igab_interface_row_ranges.argtypes = [C.c_int, ip, ip, C.POINTER(dp), C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                C.c_void_p]
        L.igab_comm_unique_id.argtypes = [C.c_void_p]
        L.igab_comm_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.igab_comm_destroy.argtypes = [C.c_void_p]
        L.igab_comm_info.argtypes = [C.c_void_p, ip, ip]
        L.igab_exchange_interface_rows.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                   C.c_void_p, C.c_void_p]
        L.igab_patch_real_index_maps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int,
                                                 C.c_void_p]
        L.igab_load_vector.argtypes = [C.c_void_p, C.c_int64, C.c_int64, ip, C.POINTER(dp), C.POINTER(dp), C.c_void_p,
                                       C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.igab_load_vector_points.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.igab_global_load_vector.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, ip, C.POINTER(dp), C.POINTER(dp),
                                              C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _check(st):
    if st != 0:
        raise IgabError(st, lib().igab_last_error_string().decode())


def launch_count():
    return lib().igab_launch_count()


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a):
    """(address, memspace) of a numpy array (HOST) or of anything with data_ptr() living on a CUDA device (DEVICE)."""
    if a is None:
        return None, None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data, HOST
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous()
        return a.data_ptr(), (DEVICE if a.is_cuda else HOST)
    raise TypeError("expected numpy array or CUDA tensor")


def _comm(c):
    return c.comm if isinstance(c, Communicator) else c


def _rule_arrays(rule1d):
    arrs = [_f64(x) for x in rule1d]
    ptrs = (dp * len(arrs))(*[a.ctypes.data_as(dp) for a in arrs])
    return arrs, ptrs


def banded(T, tol=0.0):
    """Row-banded form (first, len, coef[n_new, maxlen]) of a dense 1-D refinement operator T."""
    T = np.asarray(T, float)
    nz = np.abs(T) > tol
    first = np.argmax(nz, axis=1).astype(np.int32)
    last = (T.shape[1] - 1 - np.argmax(nz[:, ::-1], axis=1)).astype(np.int32)
    ln = (last - first + 1).astype(np.int32)
    maxlen = int(ln.max())
    coef = np.zeros((T.shape[0], maxlen))
    for i in range(T.shape[0]):
        coef[i, :ln[i]] = T[i, first[i]:first[i] + ln[i]]
    return first, ln, coef


def refineApply(pd, direction, T, newKnots, newDegree=None):
    """GPU version of patchdata._apply_along: apply the 1-D operator T to the control net of `pd` along `direction`
    (knotRefinement / degreeElevate, nurbsalgorithms.hh:264-528) and return the new NurbsPatchData."""
    from .patchdata import NurbsPatchData
    first, ln, coef = banded(T)
    ncp = _i32(pd.ncp)
    cp_in = _f64(pd.cp)
    ncp_new = list(pd.ncp)
    ncp_new[direction] = T.shape[0]
    cp_out = np.empty((int(np.prod(ncp_new)), pd.dimworld + 1))
    _check(lib().igab_refine_apply(pd.patchDim, pd.dimworld, ncp.ctypes.data_as(ip), direction, T.shape[0],
                                   coef.shape[1], first.ctypes.data, ln.ctypes.data, coef.ctypes.data_as(dp),
                                   cp_in.ctypes.data_as(dp), cp_out.ctypes.data_as(dp)))
    knots = [k.copy() for k in pd.knotSpans]
    knots[direction] = np.asarray(newKnots, float)
    deg = list(pd.degree)
    if newDegree is not None:
        deg[direction] = newDegree
    return NurbsPatchData(knots, cp_out, deg, pd.dimworld)


def refineOperators(pd, ops, cp_dev=None, device_out=False, stream=None):
    """A chain of 1-D refinement operators applied in HBM (igab_refine_operators): ops = [(direction, T), ...] with T the
    dense [n_new, n_old] operator of that step (patchdata.knotInsertionOperator / degreeElevationOperator), acting on the
    net the previous steps produced.  cp_dev: the coarse net as a CUDA tensor (else pd.cp is copied in).  Returns the new
    net [prod n_new][dimworld+1] as numpy array, or as CUDA tensor with device_out=True."""
    dim = pd.patchDim
    ncp = _i32(pd.ncp)
    cur = list(pd.ncp)
    band = []
    for d, T in ops:
        T = np.asarray(T, float)
        assert T.shape[1] == cur[d], "operator does not fit the net it is applied to"
        band.append(banded(T))
        cur[d] = T.shape[0]
    nops = len(ops)
    dirs, nnew = _i32([d for d, _ in ops]), _i32([b[2].shape[0] for b in band])
    mlen = _i32([b[2].shape[1] for b in band])
    firsts = (C.c_void_p * max(nops, 1))(*[b[0].ctypes.data for b in band])
    lens = (C.c_void_p * max(nops, 1))(*[b[1].ctypes.data for b in band])
    coefs = (dp * max(nops, 1))(*[b[2].ctypes.data_as(dp) for b in band])
    src = cp_dev if cp_dev is not None else _f64(pd.cp)
    sa, ss = _ptr(src)
    nout = int(np.prod(cur))
    if device_out:
        import torch
        out = torch.empty((nout, pd.dimworld + 1), dtype=torch.float64, device="cuda")
    else:
        out = np.empty((nout, pd.dimworld + 1))
    oa, os_ = _ptr(out)
    _check(lib().igab_refine_operators(dim, pd.dimworld, ncp.ctypes.data_as(ip), nops, dirs.ctypes.data_as(ip),
                                       nnew.ctypes.data_as(ip), mlen.ctypes.data_as(ip), firsts, lens, coefs, sa, ss, oa,
                                       os_, stream))
    return out


def degreeElevateAll(pd, factors, cp_dev=None, device_out=False, stream=None):
    """degreeElevate of every direction with factors[d] > 0 (NURBSGrid::globalDegreeElevate, nurbsgrid.hh:166-178 ->
    degreeElevate, nurbsalgorithms.hh:264-439) without the net leaving HBM: the 1-D elevation operators are built on the
    host (they depend on one knot vector each), the chain runs on the device.  Returns (degree, knotSpans, net)."""
    from .patchdata import degreeElevationOperator
    ops, knots, deg = [], [k.copy() for k in pd.knotSpans], list(pd.degree)
    for d, t in enumerate(factors):
        if t > 0:
            knots[d], E = degreeElevationOperator(pd.knotSpans[d], pd.degree[d], t)
            deg[d] += t
            ops.append((d, E))
    return deg, knots, refineOperators(pd, ops, cp_dev=cp_dev, device_out=device_out, stream=stream)


def refineKnots(pd, newKnotVectors, cp_dev=None, device_out=False, stream=None):
    """Knot refinement of all directions on the device (igab_refine_knots): pd = coarse NurbsPatchData, newKnotVectors =
    the refined knot vector per direction (each containing the old one).  cp_dev: the coarse net as a CUDA tensor (else
    pd.cp is copied in).  Returns the refined net [prod n_new][dimworld+1] as numpy array, or as CUDA tensor with
    device_out=True (feed it to Patch.fromDeviceNet: the refined net never leaves HBM)."""
    dim = pd.patchDim
    deg = _i32(pd.degree)
    ko = [_f64(k) for k in pd.knotSpans]
    kn = [_f64(k) for k in newKnotVectors]
    kop = (dp * dim)(*[k.ctypes.data_as(dp) for k in ko])
    knp = (dp * dim)(*[k.ctypes.data_as(dp) for k in kn])
    nko, nkn = _i32([len(k) for k in ko]), _i32([len(k) for k in kn])
    nout = int(np.prod([len(kn[d]) - pd.degree[d] - 1 for d in range(dim)]))
    src = cp_dev if cp_dev is not None else _f64(pd.cp)
    sa, ss = _ptr(src)
    if device_out:
        import torch
        out = torch.empty((nout, pd.dimworld + 1), dtype=torch.float64, device="cuda")
    else:
        out = np.empty((nout, pd.dimworld + 1))
    oa, os_ = _ptr(out)
    _check(lib().igab_refine_knots(dim, pd.dimworld, deg.ctypes.data_as(ip), nko.ctypes.data_as(ip), kop,
                                   nkn.ctypes.data_as(ip), knp, sa, ss, oa, os_, stream))
    return out


def pinned_empty(shape, dtype=np.float64):
    """numpy array over page-locked memory from igab_host_alloc; release it with pinned_free(arr) (views of the array
    must not outlive that call -- nothing is freed behind the caller's back)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    _check(lib().igab_host_alloc(C.byref(p), max(n, 1)))
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[arr.ctypes.data] = p.value
    return arr


_PINNED = {}


def pinned_free(arr):
    p = _PINNED.pop(arr.ctypes.data, None)
    if p is not None:
        _check(lib().igab_host_free(C.c_void_p(p)))


class Patch:
    """Device-resident patch handle.  Mirrors what NURBSPatch + NURBSGeometry + NurbsPreBasis hold for the hot path
    (dune/iga/nurbspatch.hh:53-69, nurbsgeometry.hh:58-85, nurbsbasis.hh:503-518)."""

    def __init__(self, degree, knotSpans, cp, dimworld, trimflags=None):
        self._init(degree, knotSpans, cp, dimworld, trimflags)

    def _init(self, degree, knotSpans, cp, dimworld, trimflags=None):
        self.dim = len(degree)
        self.dimworld = int(dimworld)
        self.degree = [int(p) for p in degree]
        kn = [_f64(k) for k in knotSpans]
        on_device = hasattr(cp, "data_ptr")
        if not on_device:
            cp = _f64(cp).reshape(-1, self.dimworld + 1)
        ncp = [len(kn[d]) - self.degree[d] - 1 for d in range(self.dim)]
        if int(cp.shape[0]) != int(np.prod(ncp)):
            raise ValueError("control net has %d points, knots/degree imply %s" % (cp.shape[0], ncp))
        kptr = (dp * self.dim)(*[k.ctypes.data_as(dp) for k in kn])
        tf = None
        if trimflags is not None:
            tf = np.ascontiguousarray(trimflags, dtype=np.uint8)
        h = C.c_void_p()
        tfp = tf.ctypes.data_as(C.POINTER(C.c_ubyte)) if tf is not None else None
        if on_device:
            assert cp.is_cuda and cp.is_contiguous()
            _check(lib().igab_patch_create_from_device(self.dim, self.dimworld, _i32(self.degree).ctypes.data_as(ip),
                                                       _i32([len(k) for k in kn]).ctypes.data_as(ip), kptr,
                                                       cp.data_ptr(), tfp, C.byref(h)))
        else:
            _check(lib().igab_patch_create(self.dim, self.dimworld, _i32(self.degree).ctypes.data_as(ip),
                                           _i32([len(k) for k in kn]).ctypes.data_as(ip), kptr, cp.ctypes.data_as(dp),
                                           tfp, C.byref(h)))
        self.h = h
        ne, nb, nd = C.c_int64(), C.c_int32(), C.c_int64()
        ned = (C.c_int32 * 3)()
        _check(lib().igab_patch_sizes(self.h, C.byref(ne), ned, C.byref(nb), C.byref(nd)))
        self.nelem, self.nb, self.ndof_untrimmed = ne.value, nb.value, nd.value
        self.nel_dir = [ned[d] for d in range(self.dim)]
        self.nh = self.dim * (self.dim + 1) // 2

    @classmethod
    def fromPatchData(cls, pd, trimflags=None):
        return cls(pd.degree, pd.knotSpans, pd.cp, pd.dimworld, trimflags)

    @classmethod
    def fromDeviceNet(cls, degree, knotSpans, cp_dev, dimworld, trimflags=None):
        """Handle from a control net that already lives on the device (CUDA tensor [n][dimworld+1]), e.g. the output of
        refineKnots(..., device_out=True): igab_patch_create_from_device."""
        self = cls.__new__(cls)
        self._init(degree, knotSpans, cp_dev, dimworld, trimflags)
        return self

    def close(self):
        if getattr(self, "h", None):
            lib().igab_patch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setKernel(self, mode):
        _check(lib().igab_patch_set_kernel(self.h, mode))

    def updateControlPoints(self, cp):
        cp = _f64(cp)
        _check(lib().igab_patch_update_control_points(self.h, cp.ctypes.data_as(dp)))

    # ---- integers
    def spans(self):
        out = np.empty((self.nelem, self.dim), dtype=np.int32)
        _check(lib().igab_patch_spans(self.h, out.ctypes.data, HOST, None))
        return out

    def dofmap(self):
        out = np.empty((self.nelem, self.nb), dtype=np.int32)
        n = C.c_int64()
        _check(lib().igab_patch_dofmap(self.h, out.ctypes.data, C.byref(n), HOST, None))
        return out, n.value

    def realIndexMaps(self):
        """(direct_of_real [n_real], real_of_direct [n_elem], -1 for empty elements): getDirectIndex<0> / getRealIndex<0>
        of the reference's trimmed patch (nurbspatch.hh:599-631)."""
        dor = np.empty(self.nelem, dtype=np.int32)
        rod = np.empty(self.nelem, dtype=np.int32)
        n = C.c_int64()
        _check(lib().igab_patch_real_index_maps(self.h, dor.ctypes.data, rod.ctypes.data, C.byref(n), HOST, None))
        return dor[:n.value].copy(), rod

    # ---- evaluation
    def shapes(self, npts):
        return dict(N=(npts, self.nb), dN=(npts, self.nb, self.dim), d2N=(npts, self.nb, self.nh),
                    x=(npts, self.dimworld), Jt=(npts, self.dim, self.dimworld), H=(npts, self.nh, self.dimworld),
                    detJ=(npts,), JinvT=(npts, self.dimworld, self.dim))

    def _outset(self, npts, order, want, out):
        sh = self.shapes(npts)
        res, eo, space = {}, EvalOut(), None
        for f in FIELDS:
            need = f in want and not (f in ("d2N", "H") and order < 2) and not (
                f in ("dN", "Jt", "detJ", "JinvT") and order < 1)
            if not need:
                continue
            arr = out[f] if (out is not None and f in out) else np.empty(sh[f])
            addr, sp = _ptr(arr)
            if space is not None and sp != space:
                raise ValueError("all outputs must live in the same memory space")
            space = sp
            setattr(eo, f, addr)
            res[f] = arr
        return res, eo, (HOST if space is None else space)

    def evalTensor(self, xi1d, order=1, elem_begin=0, elem_end=None, want=FIELDS, out=None, stream=None):
        """All points of all elements [elem_begin, elem_end) at the tensor rule xi1d (list of 1-D arrays in [0,1])."""
        if elem_end is None:
            elem_end = self.nelem
        arrs, ptrs = _rule_arrays(xi1d)
        nq1 = _i32([len(a) for a in arrs])
        npts = (elem_end - elem_begin) * int(np.prod(nq1))
        res, eo, space = self._outset(npts, order, want, out)
        _check(lib().igab_eval_tensor(self.h, elem_begin, elem_end, order, nq1.ctypes.data_as(ip), ptrs,
                                      C.byref(eo), space, stream))
        return res

    def evalTensorBegin(self, xi1d, out, order=1, elem_begin=0, elem_end=None, want=FIELDS, stream=None):
        """igab_eval_tensor_begin: queue the evaluation of [elem_begin, elem_end) into the HOST arrays `out` (pinned_empty)
        and return at once; wait() before reading them."""
        if elem_end is None:
            elem_end = self.nelem
        arrs, ptrs = _rule_arrays(xi1d)
        nq1 = _i32([len(a) for a in arrs])
        npts = (elem_end - elem_begin) * int(np.prod(nq1))
        res, eo, space = self._outset(npts, order, want, out)
        assert space == HOST
        _check(lib().igab_eval_tensor_begin(self.h, elem_begin, elem_end, order, nq1.ctypes.data_as(ip), ptrs,
                                            C.byref(eo), stream))
        return res

    def wait(self):
        _check(lib().igab_patch_wait(self.h))

    def evalPoints(self, elem, xi, order=1, want=FIELDS, out=None, stream=None):
        ea, es = _ptr(elem if not isinstance(elem, (list, tuple)) else _i32(elem))
        if isinstance(elem, (list, tuple, np.ndarray)):
            elem = _i32(elem)
            xi = _f64(xi).reshape(len(elem), self.dim)
            ea, es = _ptr(elem)
        xa, xs = _ptr(xi)
        npts = len(elem) if es == HOST else elem.numel()
        res, eo, space = self._outset(npts, order, want, out)
        if npts and space != es:
            raise ValueError("inputs and outputs must live in the same memory space")
        _check(lib().igab_eval_points(self.h, npts, ea, xa, order, C.byref(eo), es, stream))
        return res

    def partial(self, elem, xi, alpha):
        elem = _i32(elem)
        xi = _f64(xi).reshape(len(elem), self.dim)
        out = np.empty((len(elem), self.nb))
        _check(lib().igab_partial(self.h, len(elem), elem.ctypes.data, xi.ctypes.data,
                                  _i32(alpha).ctypes.data_as(ip), out.ctypes.data, HOST, None))
        return out

    # ---- element matrices / reductions
    def assemble(self, kind, xi1d, w1d, D=None, fconst=1.0, elem_begin=0, elem_end=None, want_fe=True, Ke=None,
                 fe=None, stream=None):
        if elem_end is None:
            elem_end = self.nelem
        n = self.nb * (self.dim if kind == ASM_ELASTICITY else 1)
        ne = elem_end - elem_begin
        if Ke is None:
            Ke = np.empty((ne, n, n))
            fe = np.empty((ne, n)) if want_fe else None
        ka, ks = _ptr(Ke)
        fa, _ = _ptr(fe)
        xa, xp = _rule_arrays(xi1d)
        wa, wp = _rule_arrays(w1d)
        nq1 = _i32([len(a) for a in xa])
        Dm = _f64(D) if D is not None else None
        _check(lib().igab_assemble(self.h, kind, Dm.ctypes.data_as(dp) if Dm is not None else None, fconst,
                                   elem_begin, elem_end, nq1.ctypes.data_as(ip), xp, wp, ka, fa, ks, stream))
        return Ke, fe

    def assemblePoints(self, kind, elem, offset, xi, w, D=None, fconst=1.0, want_fe=True, stream=None):
        """Element matrices of elements with their own quadrature (trimmed elements): list element k is patch element
        elem[k] with points offset[k]:offset[k+1] of xi [npts][2], w [npts] (quadrature.trimmedBatch produces them)."""
        elem = _i32(elem)
        offset = np.ascontiguousarray(offset, dtype=np.int64)
        xi, w = _f64(xi), _f64(w)
        n = self.nb * (self.dim if kind == ASM_ELASTICITY else 1)
        Ke = np.empty((len(elem), n, n))
        fe = np.empty((len(elem), n)) if want_fe else None
        Dm = _f64(D) if D is not None else None
        _check(lib().igab_assemble_points(self.h, kind, Dm.ctypes.data_as(dp) if Dm is not None else None, fconst,
                                          len(elem), elem.ctypes.data, offset.ctypes.data, xi.ctypes.data, w.ctypes.data,
                                          Ke.ctypes.data, fe.ctypes.data if fe is not None else None, HOST, stream))
        return Ke, fe

    def loadVector(self, xi1d, w1d, f, ncomp=1, elem_begin=0, elem_end=None, fe=None, stream=None):
        """Element load vectors with the source given per quadrature point: f [(elem_end-elem_begin)*nq][ncomp]
        (poisson.py:79 with f(posGlobal) evaluated by the caller)."""
        if elem_end is None:
            elem_end = self.nelem
        if isinstance(f, np.ndarray):
            f = _f64(f)
        fa, fs = _ptr(f)
        if fe is None:
            shape = (elem_end - elem_begin, self.nb * ncomp)
            if fs == HOST:
                fe = np.empty(shape)
            else:
                import torch
                fe = torch.empty(shape, dtype=torch.float64, device=f.device)
        xa, xp = _rule_arrays(xi1d)
        wa, wp = _rule_arrays(w1d)
        nq1 = _i32([len(a) for a in xa])
        _check(lib().igab_load_vector(self.h, elem_begin, elem_end, nq1.ctypes.data_as(ip), xp, wp, fa, ncomp,
                                      _ptr(fe)[0], fs, stream))
        return fe

    def loadVectorPoints(self, elem, offset, xi, w, f, ncomp=1):
        """Element load vectors of elements with their own rules (trimmed elements): lists as in assemblePoints,
        f [npts][ncomp] indexed like xi / w."""
        elem = _i32(elem)
        offset = np.ascontiguousarray(offset, dtype=np.int64)
        xi, w, f = _f64(xi), _f64(w), _f64(f)
        fe = np.empty((len(elem), self.nb * ncomp))
        _check(lib().igab_load_vector_points(self.h, len(elem), elem.ctypes.data, offset.ctypes.data, xi.ctypes.data,
                                             w.ctypes.data, f.ctypes.data, ncomp, fe.ctypes.data, HOST, None))
        return fe

    def globalLoadVector(self, xi1d, w1d, f, ncomp=1, trimmed=None, f_list=None, elem_begin=0, elem_end=None, stream=None):
        """b [nrows] of the global system from a per-point source (numpy in / out, or CUDA tensors in place of f, f_list
        and the trimmed xi / w): the load-vector half of globalAssembler, element vectors stay on the device."""
        if elem_end is None:
            elem_end = self.nelem
        if isinstance(f, np.ndarray):
            f = _f64(f)
        fa, fs = _ptr(f)
        nr = C.c_int64()
        _check(lib().igab_csr_pattern(self.h, ncomp, C.byref(nr), None, None, None, HOST, None))
        if fs == HOST:
            b = np.empty(nr.value)
        else:
            import torch
            b = torch.empty(nr.value, dtype=torch.float64, device=f.device)
        xa, xp = _rule_arrays(xi1d)
        wa, wp = _rule_arrays(w1d)
        nq1 = _i32([len(a) for a in xa])
        nl, ea, oa, xia, wqa, fla = 0, None, None, None, None, None
        if trimmed is not None and len(trimmed[0]):
            te, to = _i32(trimmed[0]), np.ascontiguousarray(trimmed[1], dtype=np.int64)
            tx, tw = trimmed[2], trimmed[3]
            if fs == HOST:
                tx, tw, f_list = _f64(tx), _f64(tw), _f64(f_list)
            nl, ea, oa = len(te), te.ctypes.data, to.ctypes.data
            xia, wqa, fla = _ptr(tx)[0], _ptr(tw)[0], _ptr(f_list)[0]
        _check(lib().igab_global_load_vector(self.h, ncomp, elem_begin, elem_end, nq1.ctypes.data_as(ip), xp, wp, fa, nl,
                                             ea, oa, xia, wqa, fla, _ptr(b)[0], fs, stream))
        return b

    # ---- faces
    def evalFaces(self, elem, face, xi1d):
        """Position, UNIT outer normal and integration element at the face rule xi1d (dim-1 arrays) of faces
        face[k] = 2*d+s of elements elem[k]  (nurbsintersection.hh:86-160)."""
        elem, face = _i32(elem), _i32(face)
        arrs, ptrs = _rule_arrays(xi1d)
        nqf1 = _i32([len(a) for a in arrs])
        npts = len(elem) * int(np.prod(nqf1))
        x = np.empty((npts, self.dimworld))
        nrm = np.empty((npts, self.dimworld))
        dS = np.empty(npts)
        _check(lib().igab_eval_faces(self.h, len(elem), elem.ctypes.data, face.ctypes.data, nqf1.ctypes.data_as(ip),
                                     ptrs, x.ctypes.data, nrm.ctypes.data, dS.ctypes.data, HOST, None))
        return dict(x=x, normal=nrm, dS=dS)

    def rhsAssemble(self, fe, ncomp=1, elem_begin=0, elem_end=None, b=None, accumulate=False, stream=None):
        """b[nrows] (+)= gather of the element load vectors fe of elements [elem_begin, elem_end) (poisson.py:121)."""
        if elem_end is None:
            elem_end = self.nelem
        fa, fs = _ptr(fe)
        if b is None:
            nr = C.c_int64()  # rows of the global system: compacted DOFs on a trimmed patch
            _check(lib().igab_csr_pattern(self.h, ncomp, C.byref(nr), None, None, None, HOST, None))
            nrows = nr.value
            if fs == HOST:
                b = np.zeros(nrows)
            else:
                import torch
                b = torch.zeros(nrows, dtype=torch.float64, device=fe.device)
        ba, bs = _ptr(b)
        assert bs == fs
        _check(lib().igab_rhs_assemble(self.h, ncomp, elem_begin, elem_end, fa, ba, int(accumulate), fs, stream))
        return b

    def csrDirichlet(self, mask, values, b=None, g=None, ncomp=1, stream=None):
        """Identity rows for the marked DOFs and b = g there, in place (poisson.py:91-93,117-119)."""
        if isinstance(mask, np.ndarray):
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
        ma, ms = _ptr(mask)
        va, vs = _ptr(values)
        assert ms == vs
        _check(lib().igab_csr_dirichlet(self.h, ncomp, ma, _ptr(g)[0] if g is not None else None, va,
                                        _ptr(b)[0] if b is not None else None, vs, stream))
        return values, b

    # ---- inverse map
    def local(self, elem, xglobal):
        """Element-local coordinates of the points of elements elem[k] that map to / lie closest to xglobal[k]
        (NURBSGeometry::local, nurbsgeometry.hh:145-176).  Returns (xi [npts][dim], flag [npts])."""
        elem = _i32(elem)
        X = _f64(xglobal).reshape(len(elem), self.dimworld)
        xi = np.empty((len(elem), self.dim))
        flag = np.empty(len(elem), dtype=np.int32)
        _check(lib().igab_geometry_local(self.h, len(elem), elem.ctypes.data, X.ctypes.data, xi.ctypes.data,
                                         flag.ctypes.data, HOST, None))
        return xi, flag

    # ---- global sparse matrix
    def csrPattern(self, ncomp=1):
        """(rowptr int64 [nrows+1], colidx int32 [nnz]) of the global matrix; replaces the lil_matrix of poisson.py:87."""
        nr, nnz = C.c_int64(), C.c_int64()
        _check(lib().igab_csr_pattern(self.h, ncomp, C.byref(nr), C.byref(nnz), None, None, HOST, None))
        rowptr = np.empty(nr.value + 1, dtype=np.int64)
        colidx = np.empty(nnz.value, dtype=np.int32)
        _check(lib().igab_csr_pattern(self.h, ncomp, C.byref(nr), C.byref(nnz), rowptr.ctypes.data, colidx.ctypes.data,
                                      HOST, None))
        return rowptr, colidx

    def csrAssemble(self, Ke, ncomp=1, elem_begin=0, elem_end=None, values=None, accumulate=False, stream=None):
        """values[nnz] (+)= gather of the element matrices Ke of elements [elem_begin, elem_end) (poisson.py:113-125)."""
        if elem_end is None:
            elem_end = self.nelem
        ka, ks = _ptr(Ke)
        if values is None:
            nr, nnz = C.c_int64(), C.c_int64()
            _check(lib().igab_csr_pattern(self.h, ncomp, C.byref(nr), C.byref(nnz), None, None, HOST, None))
            if ks == HOST:
                values = np.zeros(nnz.value)
            else:
                import torch
                values = torch.zeros(nnz.value, dtype=torch.float64, device=Ke.device)
        va, vs = _ptr(values)
        if vs != ks:
            raise ValueError("Ke and values must live in the same memory space")
        _check(lib().igab_csr_assemble(self.h, ncomp, elem_begin, elem_end, ka, None, va, 1 if accumulate else 0, ks,
                                       stream))
        return values

    def globalAssemble(self, kind, xi1d, w1d, D=None, fconst=1.0, trimmed=None, want_b=True, device=False, stream=None,
                       out=None, elem_begin=0, elem_end=None):
        """globalAssembler of poisson.py:84-127 in one call: (values [nnz] of csrPattern(ncomp), b [nrows]).  The element
        matrices stay on the device.  trimmed = (elem, offset, xi, w) of the trimmed elements' own rules
        (quadrature.trimmedBatch; elem ascending) or None.  device=True returns CUDA tensors."""
        if elem_end is None:
            elem_end = self.nelem
        xa, xp = _rule_arrays(xi1d)
        wa, wp = _rule_arrays(w1d)
        nq1 = _i32([len(a) for a in xa])
        ncomp = self.dim if kind == ASM_ELASTICITY else 1
        nr, nnz = C.c_int64(), C.c_int64()
        _check(lib().igab_csr_pattern(self.h, ncomp, C.byref(nr), C.byref(nnz), None, None, HOST, None))
        if out is not None:  # caller's buffers (e.g. pinned_empty arrays): (values [nnz], b [nrows] or None)
            values, b = out
        elif device:
            import torch
            values = torch.empty(nnz.value, dtype=torch.float64, device="cuda")
            b = torch.empty(nr.value, dtype=torch.float64, device="cuda") if want_b else None
        else:
            values = np.empty(nnz.value)
            b = np.empty(nr.value) if want_b else None
        Dm = _f64(D) if D is not None else None
        nl, ea, oa, xia, wqa = 0, None, None, None, None
        if trimmed is not None and len(trimmed[0]):
            te, to = _i32(trimmed[0]), np.ascontiguousarray(trimmed[1], dtype=np.int64)
            tx, tw = _f64(trimmed[2]), _f64(trimmed[3])
            nl, ea, oa, xia, wqa = len(te), te.ctypes.data, to.ctypes.data, tx.ctypes.data, tw.ctypes.data
        va, vs = _ptr(values)
        _check(lib().igab_global_assemble(self.h, kind, Dm.ctypes.data_as(dp) if Dm is not None else None, fconst,
                                          elem_begin, elem_end, nq1.ctypes.data_as(ip), xp, wp, nl, ea, oa, xia, wqa, va,
                                          _ptr(b)[0] if b is not None else None, vs, stream))
        return values, b

    def exchangeInterfaceRows(self, comm, values=None, b=None, ncomp=1, elem_bounds=None, stream=None):
        """In place: complete the rows of the sharded global matrix / load vector that several ranks touch
        (igab_exchange_interface_rows; comm = Communicator or a raw ncclComm_t)."""
        va, vs = _ptr(values)
        ba, bs = _ptr(b)
        space = vs if vs is not None else bs
        if vs is not None and bs is not None and vs != bs:
            raise ValueError("values and b must live in the same memory space")
        eb = np.ascontiguousarray(elem_bounds, dtype=np.int64) if elem_bounds is not None else None
        c = comm.comm if isinstance(comm, Communicator) else comm
        _check(lib().igab_exchange_interface_rows(self.h, ncomp, eb.ctypes.data if eb is not None else None, va, ba, space,
                                                  c, stream))
        return values, b

    def volume(self, xi1d, w1d, elem_begin=0, elem_end=None, nccl_comm=None, stream=None, trimmed=None):
        """sum of detJ w over the tensor rule of the (full) elements [elem_begin, elem_end); trimmed = (elem [npts],
        xi [npts][dim], w [npts]) adds the trimmed elements' own points (quadrature.trimmedBatch)."""
        if elem_end is None:
            elem_end = self.nelem
        xa, xp = _rule_arrays(xi1d)
        wa, wp = _rule_arrays(w1d)
        nq1 = _i32([len(a) for a in xa])
        r = C.c_double()
        nccl_comm = _comm(nccl_comm)
        if trimmed is not None:
            te, tx, tw = _i32(trimmed[0]), _f64(trimmed[1]), _f64(trimmed[2])
            _check(lib().igab_reduce_volume_trimmed(self.h, elem_begin, elem_end, nq1.ctypes.data_as(ip), xp, wp, len(te),
                                                    te.ctypes.data, tx.ctypes.data, tw.ctypes.data, C.byref(r), nccl_comm,
                                                    stream))
            return r.value
        _check(lib().igab_reduce_volume(self.h, elem_begin, elem_end, nq1.ctypes.data_as(ip), xp, wp, C.byref(r),
                                        nccl_comm, stream))
        return r.value

    def norms(self, xi1d, w1d, u, ncomp=1, elem_begin=0, elem_end=None, nccl_comm=None, stream=None, trimmed=None):
        """(sum w detJ |u_h|^2, sum w detJ |grad u_h|^2) of the discrete field with coefficients u [ndof*ncomp] (numpy
        or CUDA tensor) over the elements [elem_begin, elem_end): squared L2 norm and energy, one fused pass."""
        if elem_end is None:
            elem_end = self.nelem
        if isinstance(u, np.ndarray):
            u = _f64(u)
        ua, us = _ptr(u)
        xa, xp = _rule_arrays(xi1d)
        wa, wp = _rule_arrays(w1d)
        nq1 = _i32([len(a) for a in xa])
        r = (C.c_double * 2)()
        nccl_comm = _comm(nccl_comm)
        if trimmed is not None:  # (elem [npts], xi [npts][dim], w [npts]): the trimmed elements' own points
            te, tx, tw = _i32(trimmed[0]), _f64(trimmed[1]), _f64(trimmed[2])
            _check(lib().igab_reduce_norms_trimmed(self.h, elem_begin, elem_end, nq1.ctypes.data_as(ip), xp, wp, ua, ncomp,
                                                   us, len(te), te.ctypes.data, tx.ctypes.data, tw.ctypes.data, r,
                                                   nccl_comm, stream))
            return np.array([r[0], r[1]])
        _check(lib().igab_reduce_norms(self.h, elem_begin, elem_end, nq1.ctypes.data_as(ip), xp, wp, ua, ncomp, us, r,
                                       nccl_comm, stream))
        return np.array([r[0], r[1]])


class Communicator:
    """ncclComm_t created through the C-ABI (igab_comm_unique_id / igab_comm_create).  `bcast(bytes_or_None) -> bytes`
    distributes rank 0's 128-byte id with whatever the host program has (MPI in a DUNE program); fromTorch uses an
    initialised torch.distributed process group for that and nothing else."""

    def __init__(self, rank, world, bcast):
        uid = (C.c_char * 128)()
        if rank == 0:
            _check(lib().igab_comm_unique_id(uid))
        raw = bcast(bytes(uid.raw) if rank == 0 else None)
        uid = (C.c_char * 128).from_buffer_copy(raw)
        self.comm = C.c_void_p()
        _check(lib().igab_comm_create(rank, world, uid, C.byref(self.comm)))
        self.rank, self.world = rank, world

    @classmethod
    def fromTorch(cls, dist, device=None):
        import torch

        def bcast(raw):
            t = torch.zeros(128, dtype=torch.uint8, device=device)
            if raw is not None:
                t = torch.tensor(list(raw), dtype=torch.uint8, device=device)
            if dist.is_initialized() and dist.get_world_size() > 1:
                dist.broadcast(t, 0)
            return bytes(t.cpu().tolist())
        multi = dist.is_initialized()
        return cls(dist.get_rank() if multi else 0, dist.get_world_size() if multi else 1, bcast)

    def info(self):
        r, w = C.c_int(), C.c_int()
        _check(lib().igab_comm_info(self.comm, C.byref(r), C.byref(w)))
        return r.value, w.value

    def close(self):
        if getattr(self, "comm", None):
            lib().igab_comm_destroy(self.comm)
            self.comm = None


def partitionSlab(nel_dir, world, rank):
    """[elem_begin, elem_end) of `rank` (igab_partition_slab): whole layers of the slowest direction."""
    n = _i32(nel_dir)
    eb, ee = C.c_int64(), C.c_int64()
    _check(lib().igab_partition_slab(len(n), n.ctypes.data, world, rank, C.byref(eb), C.byref(ee)))
    return eb.value, ee.value



def partitionWeighted(counts, world, per_layer=1):
    """world + 1 element bounds with (nearly) equal cumulative point counts (igab_partition_weighted): trimmed patches,
    counts[e] = quadrature points of element e; per_layer > 1 rounds the bounds to whole layers."""
    c = np.ascontiguousarray(counts, dtype=np.int64)
    out = np.empty(world + 1, dtype=np.int64)
    _check(lib().igab_partition_weighted(C.c_int64(c.size), c.ctypes.data_as(C.c_void_p), world, C.c_int64(max(1, int(per_layer or 1))),
                                         out.ctypes.data_as(C.c_void_p)))
    return [int(v) for v in out]

def interfaceRowRanges(degree, knotSpans, world, ncomp=1, elem_bounds=None):
    """(row_begin, row_end) int64 [world]: the rows every rank's slab touches (igab_interface_row_ranges; host only)."""
    deg = _i32(degree)
    kn = [_f64(k) for k in knotSpans]
    kptr = (dp * len(kn))(*[k.ctypes.data_as(dp) for k in kn])
    nk = _i32([len(k) for k in kn])
    rb, re = np.zeros(world, dtype=np.int64), np.zeros(world, dtype=np.int64)
    eb = np.ascontiguousarray(elem_bounds, dtype=np.int64) if elem_bounds is not None else None
    _check(lib().igab_interface_row_ranges(len(deg), deg.ctypes.data_as(ip), nk.ctypes.data_as(ip), kptr, ncomp, world,
                                           eb.ctypes.data if eb is not None else None, rb.ctypes.data, re.ctypes.data))
    return rb, re


def localKeys(degree):
    """(subEntity, codim, index) uint32 arrays of the local shape functions: NurbsLocalCoefficients::init,
    nurbsbasis.hh:220-271.  Host integers; works without a device."""
    deg = _i32(degree)
    nb = int(np.prod(deg + 1))
    sub, cod, idx = (np.empty(nb, dtype=np.uint32) for _ in range(3))
    _check(lib().igab_local_keys(len(deg), deg.ctypes.data_as(ip), sub.ctypes.data, cod.ctypes.data, idx.ctypes.data))
    return sub, cod, idx


def surfaceForms(Jt, H=None, want=("normal", "unitNormal", "metric", "secondFF", "gaussK"), stream=None):
    """normal / unitNormal / metric / secondFundamentalForm / gaussianCurvature (nurbsgeometry.hh:216-255) for every
    point of the Jt [npts][dim][dimworld] (and H [npts][nh][dimworld]) arrays an evaluation call returned.  numpy
    arrays are staged through the device; CUDA tensors are processed in place (outputs are torch tensors)."""
    if isinstance(Jt, np.ndarray):
        Jt = _f64(Jt)
        H = _f64(H) if H is not None else None
    ja, space = _ptr(Jt)
    npts, dim, dw = tuple(Jt.shape)
    ha = _ptr(H)[0] if H is not None else None
    shapes = dict(normal=(npts, 3), unitNormal=(npts, 3), metric=(npts, dim, dim), secondFF=(npts, 2, 2),
                  gaussK=(npts,))
    out, ptrs = {}, {}
    for k, shp in shapes.items():
        if k not in want:
            ptrs[k] = None
            continue
        if space == HOST:
            out[k] = np.empty(shp)
        else:
            import torch
            out[k] = torch.empty(shp, dtype=torch.float64, device=Jt.device)
        ptrs[k] = _ptr(out[k])[0]
    _check(lib().igab_surface_forms(dim, dw, npts, ja, ha, ptrs["normal"], ptrs["unitNormal"], ptrs["metric"],
                                    ptrs["secondFF"], ptrs["gaussK"], space, stream))
    return out

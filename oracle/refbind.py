"""ctypes binding of oracle/_ref/liboracle_ref.so -- TEST INFRASTRUCTURE.

The shared object is the reference's own headers (dune/iga/bsplinealgorithms.hh, nurbsalgorithms.hh,
nurbspatchgeometry.hh) compiled through oracle/ref_shim.cpp.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "liboracle_ref.so")
_lib = None

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_SO)
        L.oref_patch_create.restype = C.c_void_p
        L.oref_patch_create.argtypes = [C.c_int, C.c_int, ip, ip, dp, ip, dp]
        L.oref_patch_destroy.argtypes = [C.c_void_p]
        L.oref_patch_info.argtypes = [C.c_void_p, ip, ip, ip, ip, ip, C.POINTER(C.c_long)]
        L.oref_patch_get.argtypes = [C.c_void_p, dp, dp]
        L.oref_patch_spans.argtypes = [C.c_void_p, ip]
        for f in ("oref_patch_knot_refine",):
            getattr(L, f).restype = C.c_void_p
        L.oref_patch_knot_refine.argtypes = [C.c_void_p, C.c_int, dp, C.c_int]
        L.oref_patch_refine_uniform.restype = C.c_void_p
        L.oref_patch_refine_uniform.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oref_patch_degree_elevate.restype = C.c_void_p
        L.oref_patch_degree_elevate.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oref_make_circular_arc.restype = C.c_void_p
        L.oref_make_circular_arc.argtypes = [C.c_double, C.c_double, C.c_double]
        L.oref_make_surface_of_revolution.restype = C.c_void_p
        L.oref_make_surface_of_revolution.argtypes = [C.c_void_p, dp, dp, C.c_double]
        L.oref_find_span.restype = C.c_long
        L.oref_find_span.argtypes = [C.c_int, C.c_double, dp, C.c_int, C.c_int]
        L.oref_bspline_basis.argtypes = [C.c_double, dp, C.c_int, C.c_int, C.c_int, dp]
        L.oref_bspline_ders.argtypes = [C.c_double, dp, C.c_int, C.c_int, C.c_int, C.c_int, dp]
        L.oref_nurbs_basis.argtypes = [C.c_void_p, dp, ip, dp]
        L.oref_nurbs_ders.argtypes = [C.c_void_p, dp, C.c_int, ip, dp]
        L.oref_patchgeo_global.argtypes = [C.c_void_p, dp, dp]
        L.oref_patchgeo_jt.argtypes = [C.c_void_p, dp, dp]
        L.oref_patchgeo_detj.restype = C.c_double
        L.oref_patchgeo_detj.argtypes = [C.c_void_p, dp]
        L.oref_patchgeo_012.argtypes = [C.c_void_p, dp, dp, dp, dp]
        L.oref_patchgeo_local.restype = C.c_int
        L.oref_patchgeo_local.argtypes = [C.c_void_p, dp, dp]
        L.oref_patchgeo_volume.restype = C.c_double
        L.oref_patchgeo_volume.argtypes = [C.c_void_p, C.c_int]
        L.oref_eval_tensor.restype = C.c_int
        L.oref_eval_tensor.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_int, ip, dp] + [dp] * 7 + [C.c_int]
        L.oref_eval_points.restype = C.c_int
        L.oref_eval_points.argtypes = [C.c_void_p, C.c_long, ip, dp, C.c_int] + [dp] * 7 + [C.c_int]
        L.oref_partial.argtypes = [C.c_void_p, C.c_long, ip, dp, ip, dp]
        L.oref_max_threads.restype = C.c_int
        if hasattr(L, "oref_as_shipped_points"):
            L.oref_as_shipped_points.restype = C.c_int
            L.oref_as_shipped_points.argtypes = [C.c_void_p, C.c_long, dp, dp, C.c_int]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(ip) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class RefPatch:
    """A NURBSPatchData held by the reference code.  knots: list of 1-D arrays; cp: [..., dimworld+1] flat i-fastest."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("reference patch creation failed")
        self.h = C.c_void_p(handle)
        L = lib()
        dim, dw = C.c_int(), C.c_int()
        deg = (C.c_int * 3)()
        nk = (C.c_int * 3)()
        ncp = (C.c_int * 3)()
        nel = C.c_long()
        L.oref_patch_info(self.h, C.byref(dim), C.byref(dw), deg, nk, ncp, C.byref(nel))
        self.dim, self.dimworld = dim.value, dw.value
        self.degree = [deg[i] for i in range(self.dim)]
        self.nknots = [nk[i] for i in range(self.dim)]
        self.ncp = [ncp[i] for i in range(self.dim)]
        self.nelem = nel.value
        self.nb = int(np.prod([p + 1 for p in self.degree]))
        kn = np.empty(sum(self.nknots))
        cp = np.empty((int(np.prod(self.ncp)), self.dimworld + 1))
        L.oref_patch_get(self.h, _d(kn), _d(cp))
        self.knots = list(np.split(kn, np.cumsum(self.nknots)[:-1]))
        self.cp = cp

    @classmethod
    def create(cls, degree, knots, cp, dimworld):
        dim = len(degree)
        knots = [_f64(k) for k in knots]
        cp = _f64(cp).reshape(-1, dimworld + 1)
        ncp = [len(knots[d]) - degree[d] - 1 for d in range(dim)]
        assert cp.shape[0] == int(np.prod(ncp)), (cp.shape, ncp)
        kcat = _f64(np.concatenate(knots))
        h = lib().oref_patch_create(dim, dimworld, _i(_i32(degree)), _i(_i32([len(k) for k in knots])), _d(kcat),
                                    _i(_i32(ncp)), _d(cp))
        return cls(h)

    def __del__(self):
        try:
            lib().oref_patch_destroy(self.h)
        except Exception:
            pass

    # --- reference refinement (nurbsalgorithms.hh:49-66,264-528)
    def knot_refine(self, direction, new_knots):
        nk = _f64(new_knots)
        return RefPatch(lib().oref_patch_knot_refine(self.h, direction, _d(nk), len(nk)))

    def refine_uniform(self, direction, level):
        return RefPatch(lib().oref_patch_refine_uniform(self.h, direction, level))

    def degree_elevate(self, direction, t):
        return RefPatch(lib().oref_patch_degree_elevate(self.h, direction, t))

    # --- pointwise
    def spans(self):
        out = np.empty((self.nelem, self.dim), dtype=np.int32)
        lib().oref_patch_spans(self.h, _i(out))
        return out

    def nurbs_basis(self, u, span=None):
        out = np.empty(self.nb)
        u = _f64(u)
        s = _i32(span) if span is not None else None
        lib().oref_nurbs_basis(self.h, _d(u), _i(s), _d(out))
        return out

    def nurbs_ders(self, u, k, span=None):
        out = np.empty(((k + 1) ** self.dim, self.nb))
        u = _f64(u)
        s = _i32(span) if span is not None else None
        lib().oref_nurbs_ders(self.h, _d(u), k, _i(s), _d(out))
        return out

    def geo_global(self, u):
        out = np.empty(self.dimworld)
        lib().oref_patchgeo_global(self.h, _d(_f64(u)), _d(out))
        return out

    def geo_jt(self, u):
        out = np.empty((self.dim, self.dimworld))
        lib().oref_patchgeo_jt(self.h, _d(_f64(u)), _d(out))
        return out

    def geo_detj(self, u):
        return lib().oref_patchgeo_detj(self.h, _d(_f64(u)))

    def geo_012(self, u):
        nh = self.dim * (self.dim + 1) // 2
        pos, J, H = np.empty(self.dimworld), np.empty((self.dim, self.dimworld)), np.empty((nh, self.dimworld))
        lib().oref_patchgeo_012(self.h, _d(_f64(u)), _d(pos), _d(J), _d(H))
        return pos, J, H

    def volume(self, scale_order=1):
        return lib().oref_patchgeo_volume(self.h, scale_order)

    def geo_local(self, xglobal):
        """NURBSPatchGeometry::local (nurbspatchgeometry.hh:126-158); patch coordinates.  None when the reference threw."""
        out = np.empty(self.dim)
        rc = lib().oref_patchgeo_local(self.h, _d(_f64(xglobal)), _d(out))
        return out if rc == 0 else None

    def _alloc(self, npts, order, want):
        nh = self.dim * (self.dim + 1) // 2
        shapes = dict(N=(npts, self.nb), dN=(npts, self.nb, self.dim), d2N=(npts, self.nb, nh),
                      x=(npts, self.dimworld), Jt=(npts, self.dim, self.dimworld), H=(npts, nh, self.dimworld),
                      detJ=(npts,))
        out = {}
        for k, s in shapes.items():
            need = k in want and not (k in ("d2N", "H") and order < 2) and not (k in ("dN", "Jt", "detJ") and order < 1)
            out[k] = np.empty(s) if need else None
        return out

    def eval_tensor(self, xi1d, order=1, elem_begin=0, elem_end=None, want=("N", "dN", "d2N", "x", "Jt", "H", "detJ"),
                    nthreads=0, out=None):
        """Element-batched evaluation at the tensor rule xi1d (list of 1-D arrays in [0,1]); reference kernel path."""
        if elem_end is None:
            elem_end = self.nelem
        nq1 = _i32([len(x) for x in xi1d])
        nq = int(np.prod(nq1))
        npts = (elem_end - elem_begin) * nq
        if out is None:  # a caller timing the evaluation passes its buffers in (no allocation / first touch per call)
            out = self._alloc(npts, order, want)
        xcat = _f64(np.concatenate([_f64(x) for x in xi1d]))
        used = lib().oref_eval_tensor(self.h, elem_begin, elem_end, order, _i(nq1), _d(xcat), _d(out["N"]),
                                      _d(out["dN"]), _d(out["d2N"]), _d(out["x"]), _d(out["Jt"]), _d(out["H"]),
                                      _d(out["detJ"]), nthreads)
        out["threads"] = used
        return out

    def as_shipped_points(self, u, nthreads=0):
        """The per-point path as shipped (whole-patch extractWeights / extractControlCoordinates per call):
        evaluateFunction + evaluateJacobian + patch geometry global / jacobianTransposed / integrationElement at the
        patch coordinates u [npts][dim].  Returns (threads used, checksum).  Timing baseline only."""
        u = _f64(u).reshape(-1, self.dim)
        acc = np.zeros(1)
        used = lib().oref_as_shipped_points(self.h, len(u), _d(u), _d(acc), nthreads)
        return used, float(acc[0])

    def eval_points(self, elem, xi, order=1, want=("N", "dN", "d2N", "x", "Jt", "H", "detJ"), nthreads=0):
        elem = _i32(elem)
        xi = _f64(xi).reshape(len(elem), self.dim)
        out = self._alloc(len(elem), order, want)
        used = lib().oref_eval_points(self.h, len(elem), _i(elem), _d(xi), order, _d(out["N"]), _d(out["dN"]),
                                      _d(out["d2N"]), _d(out["x"]), _d(out["Jt"]), _d(out["H"]), _d(out["detJ"]),
                                      nthreads)
        out["threads"] = used
        return out

    def partial(self, elem, xi, alpha):
        elem = _i32(elem)
        xi = _f64(xi).reshape(len(elem), self.dim)
        out = np.empty((len(elem), self.nb))
        lib().oref_partial(self.h, len(elem), _i(elem), _d(xi), _i(_i32(alpha)), _d(out))
        return out


def make_circular_arc(radius=1.0, start=0.0, end=360.0):
    return RefPatch(lib().oref_make_circular_arc(radius, start, end))


def make_surface_of_revolution(generatrix, point, axis, angle=360.0):
    return RefPatch(lib().oref_make_surface_of_revolution(generatrix.h, _d(_f64(point)), _d(_f64(axis)), angle))


def find_span(p, u, U, offset=0):
    U = _f64(U)
    return lib().oref_find_span(p, u, _d(U), len(U), offset)


def bspline_basis(u, U, p, span=-1):
    U = _f64(U)
    out = np.empty(p + 1)
    lib().oref_bspline_basis(u, _d(U), len(U), p, span, _d(out))
    return out


def bspline_ders(u, U, p, k, span=-1):
    U = _f64(U)
    out = np.empty((k + 1, p + 1))
    lib().oref_bspline_ders(u, _d(U), len(U), p, k, span, _d(out))
    return out


def max_threads():
    return lib().oref_max_threads()

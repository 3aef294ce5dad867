"""ctypes binding of oracle/liboracle_port.so (igab_oracle.c, the plain-C restatement) -- TEST INFRASTRUCTURE.

Same method names as refbind.RefPatch so tests can run either checker.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_port.so")
_lib = None
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
MAXDIM = 3


class _CPatch(C.Structure):
    _fields_ = [("dim", C.c_int), ("dimworld", C.c_int), ("degree", C.c_int * MAXDIM), ("nknots", C.c_int * MAXDIM),
                ("knots", dp * MAXDIM), ("ncp", C.c_int * MAXDIM), ("cp", dp)]


def build():
    src = os.path.join(_HERE, "igab_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        PP = C.POINTER(_CPatch)
        L.igo_find_span.restype = C.c_long
        L.igo_find_span.argtypes = [C.c_int, C.c_double, dp, C.c_int, C.c_int]
        L.igo_bspline_basis.argtypes = [C.c_double, dp, C.c_int, C.c_int, C.c_int, dp]
        L.igo_bspline_ders.argtypes = [C.c_double, dp, C.c_int, C.c_int, C.c_int, C.c_int, dp]
        L.igo_patch_spans.restype = C.c_long
        L.igo_patch_spans.argtypes = [PP, ip]
        L.igo_dofmap.restype = C.c_long
        L.igo_dofmap.argtypes = [PP, C.POINTER(C.c_ubyte), ip]
        L.igo_nurbs_basis.argtypes = [PP, dp, ip, dp]
        L.igo_nurbs_ders.argtypes = [PP, dp, C.c_int, ip, dp]
        L.igo_eval_tensor.restype = C.c_int
        L.igo_eval_tensor.argtypes = [PP, C.c_long, C.c_long, C.c_int, ip, dp] + [dp] * 8 + [C.c_int]
        L.igo_eval_points.restype = C.c_int
        L.igo_eval_points.argtypes = [PP, C.c_long, ip, dp, C.c_int] + [dp] * 8 + [C.c_int]
        L.igo_partial.argtypes = [PP, C.c_long, ip, dp, ip, dp]
        L.igo_patchgeo.argtypes = [PP, dp, dp, dp, dp]
        L.igo_gauss01.argtypes = [C.c_int, dp, dp]
        L.igo_volume.restype = C.c_double
        L.igo_volume.argtypes = [PP, ip, dp, dp, C.c_int]
        L.igo_assemble.restype = C.c_int
        L.igo_assemble.argtypes = [PP, C.c_int, dp, C.c_double, C.c_long, C.c_long, ip, dp, dp, dp, dp, C.c_int]
        L.igo_trimmed_rule.argtypes = [C.c_long, dp, C.c_int, dp, dp, dp, dp]
        L.igo_surface_forms.argtypes = [C.c_int, C.c_int, C.c_long] + [dp] * 7
        L.igo_geometry_local.restype = C.c_int
        L.igo_geometry_local.argtypes = [PP, C.c_long, ip, dp, dp, ip]
        L.igo_norms.restype = C.c_int
        L.igo_norms.argtypes = [PP, C.c_long, C.c_long, ip, dp, dp, dp, C.c_int, dp]
        L.igo_assemble_points.restype = C.c_int
        L.igo_assemble_points.argtypes = [PP, C.c_int, dp, C.c_double, C.c_long, ip, C.POINTER(C.c_long), dp, dp, dp, dp]
        L.igo_max_threads.restype = C.c_int
        L.igo_load_vector.argtypes = [PP, C.c_long, C.c_long, ip, dp, dp, dp, C.c_int, dp]
        L.igo_load_vector_points.argtypes = [PP, C.c_long, ip, C.POINTER(C.c_long), dp, dp, dp, C.c_int, dp]
        L.igo_real_index_maps.restype = C.c_long
        L.igo_real_index_maps.argtypes = [C.c_long, C.POINTER(C.c_ubyte), ip, ip]
        L.igo_merge_duplicate_points.restype = C.c_long
        L.igo_merge_duplicate_points.argtypes = [C.c_long, dp, dp]
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


class PortPatch:
    def __init__(self, degree, knots, cp, dimworld):
        self.dim = len(degree)
        self.dimworld = dimworld
        self.degree = [int(p) for p in degree]
        self.knots = [_f64(k) for k in knots]
        self.nknots = [len(k) for k in self.knots]
        self.ncp = [self.nknots[d] - self.degree[d] - 1 for d in range(self.dim)]
        self.cp = _f64(cp).reshape(-1, dimworld + 1)
        assert self.cp.shape[0] == int(np.prod(self.ncp))
        self.nb = int(np.prod([p + 1 for p in self.degree]))
        c = _CPatch()
        c.dim, c.dimworld = self.dim, dimworld
        for d in range(self.dim):
            c.degree[d], c.nknots[d], c.ncp[d] = self.degree[d], self.nknots[d], self.ncp[d]
            c.knots[d] = _d(self.knots[d])
        c.cp = _d(self.cp)
        self.c = c
        self.nelem = lib().igo_patch_spans(C.byref(c), None)

    @classmethod
    def create(cls, degree, knots, cp, dimworld):
        return cls(degree, knots, cp, dimworld)

    def spans(self):
        out = np.empty((self.nelem, self.dim), dtype=np.int32)
        lib().igo_patch_spans(C.byref(self.c), _i(out))
        return out

    def dofmap(self, trimflag=None):
        out = np.empty((self.nelem, self.nb), dtype=np.int32)
        tf = None
        if trimflag is not None:
            tf = np.ascontiguousarray(trimflag, dtype=np.uint8)
        n = lib().igo_dofmap(C.byref(self.c), tf.ctypes.data_as(C.POINTER(C.c_ubyte)) if tf is not None else None,
                             _i(out))
        return out, n

    def nurbs_basis(self, u, span=None):
        out = np.empty(self.nb)
        s = _i32(span) if span is not None else None
        lib().igo_nurbs_basis(C.byref(self.c), _d(_f64(u)), _i(s), _d(out))
        return out

    def nurbs_ders(self, u, k, span=None):
        out = np.empty(((k + 1) ** self.dim, self.nb))
        s = _i32(span) if span is not None else None
        lib().igo_nurbs_ders(C.byref(self.c), _d(_f64(u)), k, _i(s), _d(out))
        return out

    def geo_global(self, u):
        out = np.empty(self.dimworld)
        lib().igo_patchgeo(C.byref(self.c), _d(_f64(u)), _d(out), None, None)
        return out

    def geo_jt(self, u):
        out = np.empty((self.dim, self.dimworld))
        lib().igo_patchgeo(C.byref(self.c), _d(_f64(u)), None, _d(out), None)
        return out

    def geo_detj(self, u):
        out = C.c_double()
        lib().igo_patchgeo(C.byref(self.c), _d(_f64(u)), None, None, C.cast(C.byref(out), dp))
        return out.value

    def _alloc(self, npts, order, want):
        nh = self.dim * (self.dim + 1) // 2
        shapes = dict(N=(npts, self.nb), dN=(npts, self.nb, self.dim), d2N=(npts, self.nb, nh),
                      x=(npts, self.dimworld), Jt=(npts, self.dim, self.dimworld), H=(npts, nh, self.dimworld),
                      detJ=(npts,), JinvT=(npts, self.dimworld, self.dim))
        out = {}
        for k, s in shapes.items():
            need = k in want and not (k in ("d2N", "H") and order < 2) and not (
                k in ("dN", "Jt", "detJ", "JinvT") and order < 1)
            out[k] = np.empty(s) if need else None
        return out

    def eval_tensor(self, xi1d, order=1, elem_begin=0, elem_end=None,
                    want=("N", "dN", "d2N", "x", "Jt", "H", "detJ", "JinvT"), nthreads=0, out=None):
        if elem_end is None:
            elem_end = self.nelem
        nq1 = _i32([len(x) for x in xi1d])
        npts = (elem_end - elem_begin) * int(np.prod(nq1))
        if out is None:
            out = self._alloc(npts, order, want)
        xcat = _f64(np.concatenate([_f64(x) for x in xi1d]))
        used = lib().igo_eval_tensor(C.byref(self.c), elem_begin, elem_end, order, _i(nq1), _d(xcat), _d(out["N"]),
                                     _d(out["dN"]), _d(out["d2N"]), _d(out["x"]), _d(out["Jt"]), _d(out["H"]),
                                     _d(out["detJ"]), _d(out["JinvT"]), nthreads)
        out["threads"] = used
        return out

    def eval_points(self, elem, xi, order=1, want=("N", "dN", "d2N", "x", "Jt", "H", "detJ", "JinvT"), nthreads=0):
        elem = _i32(elem)
        xi = _f64(xi).reshape(len(elem), self.dim)
        out = self._alloc(len(elem), order, want)
        used = lib().igo_eval_points(C.byref(self.c), len(elem), _i(elem), _d(xi), order, _d(out["N"]), _d(out["dN"]),
                                     _d(out["d2N"]), _d(out["x"]), _d(out["Jt"]), _d(out["H"]), _d(out["detJ"]),
                                     _d(out["JinvT"]), nthreads)
        out["threads"] = used
        return out

    def partial(self, elem, xi, alpha):
        elem = _i32(elem)
        xi = _f64(xi).reshape(len(elem), self.dim)
        out = np.empty((len(elem), self.nb))
        lib().igo_partial(C.byref(self.c), len(elem), _i(elem), _d(xi), _i(_i32(alpha)), _d(out))
        return out

    def norms(self, xi1d, w1d, u, ncomp=1, elem_begin=0, elem_end=None):
        """(sum w detJ |u_h|^2, sum w detJ |grad u_h|^2) over the elements [elem_begin, elem_end)."""
        if elem_end is None:
            elem_end = self.nelem
        nq1 = _i32([len(x) for x in xi1d])
        out = np.empty(2)
        lib().igo_norms(C.byref(self.c), elem_begin, elem_end, _i(nq1), _d(_f64(np.concatenate(xi1d))),
                        _d(_f64(np.concatenate(w1d))), _d(_f64(u)), ncomp, _d(out))
        return out

    def assemble_points(self, kind, elem, offset, xi, w, D=None, fconst=1.0, want_fe=True):
        elem = _i32(elem)
        offset = np.ascontiguousarray(offset, dtype=np.int64)
        n = self.nb * (self.dim if kind == 2 else 1)
        Ke = np.empty((len(elem), n, n))
        fe = np.empty((len(elem), n)) if want_fe else None
        Dm = _f64(D) if D is not None else None
        lib().igo_assemble_points(C.byref(self.c), kind, _d(Dm), fconst, len(elem), _i(elem),
                                  offset.ctypes.data_as(C.POINTER(C.c_long)), _d(_f64(xi)), _d(_f64(w)), _d(Ke), _d(fe))
        return Ke, fe

    def load_vector(self, xi1d, w1d, f, ncomp=1, elem_begin=0, elem_end=None):
        """Element load vectors with a per-point source f [(ee-eb)*nq][ncomp] (poisson.py:79)."""
        if elem_end is None:
            elem_end = self.nelem
        fe = np.empty((elem_end - elem_begin, self.nb * ncomp))
        nq1 = _i32([len(x) for x in xi1d])
        lib().igo_load_vector(C.byref(self.c), elem_begin, elem_end, _i(nq1), _d(_f64(np.concatenate(xi1d))),
                              _d(_f64(np.concatenate(w1d))), _d(_f64(f)), ncomp, _d(fe))
        return fe

    def load_vector_points(self, elem, offset, xi, w, f, ncomp=1):
        elem = _i32(elem)
        offset = np.ascontiguousarray(offset, dtype=np.int64)
        fe = np.empty((len(elem), self.nb * ncomp))
        lib().igo_load_vector_points(C.byref(self.c), len(elem), _i(elem), offset.ctypes.data_as(C.POINTER(C.c_long)),
                                     _d(_f64(xi)), _d(_f64(w)), _d(_f64(f)), ncomp, _d(fe))
        return fe

    def local(self, elem, xglobal):
        """NURBSGeometry::local per point (nurbsgeometry.hh:145-176): element-local xi and a flag per point."""
        elem = _i32(elem)
        X = _f64(xglobal).reshape(len(elem), self.dimworld)
        xi = np.empty((len(elem), self.dim))
        flag = np.empty(len(elem), np.int32)
        rc = lib().igo_geometry_local(C.byref(self.c), len(elem), _i(elem), _d(X), _d(xi), _i(flag))
        if rc:
            raise ValueError("local: unsupported (dim, dimworld)")
        return xi, flag

    def volume(self, xi1d, w1d, nthreads=0):
        nq1 = _i32([len(x) for x in xi1d])
        return lib().igo_volume(C.byref(self.c), _i(nq1), _d(_f64(np.concatenate(xi1d))),
                                _d(_f64(np.concatenate(w1d))), nthreads)

    def assemble(self, kind, xi1d, w1d, D=None, fconst=1.0, elem_begin=0, elem_end=None, want_fe=True, nthreads=0):
        """kind: 0 poisson, 1 mass, 2 elasticity.  Returns Ke[ne][n][n], fe[ne][n]."""
        if elem_end is None:
            elem_end = self.nelem
        n = self.nb * (self.dim if kind == 2 else 1)
        ne = elem_end - elem_begin
        Ke = np.empty((ne, n, n))
        fe = np.empty((ne, n)) if want_fe else None
        nq1 = _i32([len(x) for x in xi1d])
        Dm = _f64(D) if D is not None else None
        lib().igo_assemble(C.byref(self.c), kind, _d(Dm), fconst, elem_begin, elem_end, _i(nq1),
                           _d(_f64(np.concatenate(xi1d))), _d(_f64(np.concatenate(w1d))), _d(Ke), _d(fe), nthreads)
        return Ke, fe


def find_span(p, u, U, offset=0):
    U = _f64(U)
    return lib().igo_find_span(p, u, _d(U), len(U), offset)


def bspline_basis(u, U, p, span=-1):
    U = _f64(U)
    out = np.empty(p + 1)
    lib().igo_bspline_basis(u, _d(U), len(U), p, span, _d(out))
    return out


def bspline_ders(u, U, p, k, span=-1):
    U = _f64(U)
    out = np.empty((k + 1, p + 1))
    lib().igo_bspline_ders(u, _d(U), len(U), p, k, span, _d(out))
    return out


def gauss01(n):
    x, w = np.empty(n), np.empty(n)
    lib().igo_gauss01(n, _d(x), _d(w))
    return x, w


def trimmed_rule(tri, ref_pts, ref_w):
    tri = _f64(tri).reshape(-1, 3, 2)
    ref_pts, ref_w = _f64(ref_pts), _f64(ref_w)
    n = tri.shape[0] * len(ref_w)
    xi, w = np.empty((n, 2)), np.empty(n)
    lib().igo_trimmed_rule(tri.shape[0], _d(tri), len(ref_w), _d(ref_pts), _d(ref_w), _d(xi), _d(w))
    return xi, w


def surface_forms(Jt, H=None, want=("normal", "unitNormal", "metric", "secondFF", "gaussK")):
    """nurbsgeometry.hh:216-255 from Jt [npts][dim][dw] and H [npts][nh][dw]."""
    Jt = _f64(Jt)
    npts, dim, dw = Jt.shape
    H = _f64(H) if H is not None else None
    shapes = dict(normal=(npts, 3), unitNormal=(npts, 3), metric=(npts, dim, dim), secondFF=(npts, 2, 2), gaussK=(npts,))
    out = {k: (np.empty(shapes[k]) if k in want else None) for k in shapes}
    lib().igo_surface_forms(dim, dw, npts, _d(Jt), _d(H), _d(out["normal"]), _d(out["unitNormal"]), _d(out["metric"]),
                            _d(out["secondFF"]), _d(out["gaussK"]))
    return {k: v for k, v in out.items() if v is not None}


def real_index_maps(nelem, trimflag=None):
    """(direct_of_real, real_of_direct) of NURBSPatch::getDirectIndex<0> / getRealIndex<0> (nurbspatch.hh:599-631)."""
    dor = np.empty(nelem, dtype=np.int32)
    rod = np.empty(nelem, dtype=np.int32)
    tf = np.ascontiguousarray(trimflag, dtype=np.uint8) if trimflag is not None else None
    n = lib().igo_real_index_maps(nelem, tf.ctypes.data_as(C.POINTER(C.c_ubyte)) if tf is not None else None, _i(dor),
                                  _i(rod))
    return dor[:n].copy(), rod


def merge_duplicate_points(xi, w):
    """Gauss-Lobatto branch of fillQuadratureRuleImpl (fillquadraturerule.hh:21-44)."""
    xi = _f64(xi).reshape(-1, 2).copy()
    w = _f64(w).copy()
    m = lib().igo_merge_duplicate_points(len(w), _d(xi), _d(w))
    return xi[:m].copy(), w[:m].copy()


def local_keys(degree):
    deg = _i32(degree)
    nb = int(np.prod(deg + 1))
    sub, cod, idx = (np.empty(nb, dtype=np.uint32) for _ in range(3))
    up = C.POINTER(C.c_uint)
    lib().igo_local_keys.argtypes = [C.c_int, ip, up, up, up]
    lib().igo_local_keys(len(deg), _i(deg), sub.ctypes.data_as(up), cod.ctypes.data_as(up), idx.ctypes.data_as(up))
    return sub, cod, idx


def max_threads():
    return lib().igo_max_threads()

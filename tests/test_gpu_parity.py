"""GPU parity tests (-m gpu): the CUDA path, called through the C-ABI, against the oracle on the same seeded inputs,
against the reference's own literals, and -- at sizes the oracle cannot reach -- through size-independent properties.

Bar (BASELINE.json north_star): bit-exact for span indices and element-to-global DOF maps; <= 1e-12 relative (see
parity_util.relerr) for basis values, derivatives, Jacobians, integration elements and element matrices.
"""
import json
import os

import numpy as np
import pytest

import igab200  # noqa: F401
import patches_for_tests as PFT
import portbind as PT
from dune_iga_b200 import capi
from dune_iga_b200 import patchdata as PD
from parity_util import TOL, oracle_for, port_for, relerr

pytestmark = pytest.mark.gpu

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_literals.json")))
NAMES = ["curve_p3", "plate_hole_p2", "cylinder_p3", "cylinder_p4", "torus_p2", "square_p3", "mixed_deg_3d"]
EVAL_FIELDS = ("N", "dN", "d2N", "x", "Jt", "H", "detJ")


def gpu_patch(spec, **kw):
    return capi.Patch(spec["degree"], spec["knots"], spec["cp"], spec["dimworld"], **kw)


def gauss(spec, extra=0):
    return [PD.gaussLegendre01(p + 1 + extra)[0] for p in spec["degree"]]


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("kernel", [capi.KERNEL_GENERIC, capi.KERNEL_AUTO, capi.KERNEL_WARP], ids=["generic", "auto", "warp"])
def test_eval_tensor_matches_oracle(name, kernel):
    spec = PFT.small_patch(name)
    O, kind = oracle_for(spec)
    P = gpu_patch(spec)
    P.setKernel(kernel)
    xi1d = gauss(spec)
    ref = O.eval_tensor(xi1d, order=2, want=EVAL_FIELDS)
    got = P.evalTensor(xi1d, order=2, want=EVAL_FIELDS)
    for f in EVAL_FIELDS:
        assert relerr(got[f], ref[f]) <= TOL, (f, kind, relerr(got[f], ref[f]))
    # partition of unity and derivative sums (gridtests.cpp:787,804)
    assert np.abs(got["N"].sum(1) - 1.0).max() < 1e-13
    assert np.abs(got["dN"].sum(1)).max() < 1e-10 * max(1.0, np.abs(got["dN"]).max())


@pytest.mark.parametrize("name", NAMES)
def test_eval_tensor_order1_subrange_and_jinvt(name):
    spec = PFT.small_patch(name)
    Pp = port_for(spec)
    P = gpu_patch(spec)
    xi1d = gauss(spec, extra=1)  # a rule that is NOT p+1 points: exercises the general path
    eb, ee = Pp.nelem // 3, Pp.nelem - 1
    ref = Pp.eval_tensor(xi1d, order=1, elem_begin=eb, elem_end=ee)
    got = P.evalTensor(xi1d, order=1, elem_begin=eb, elem_end=ee)
    assert "d2N" not in got and "H" not in got
    for f in ("N", "dN", "x", "Jt", "detJ", "JinvT"):
        assert relerr(got[f], ref[f]) <= TOL, (f, relerr(got[f], ref[f]))
    # empty range is legal
    assert P.evalTensor(xi1d, order=1, elem_begin=2, elem_end=2)["N"].shape[0] == 0


@pytest.mark.parametrize("name", NAMES)
def test_integers_bit_exact(name):
    spec = PFT.small_patch(name)
    Pp = port_for(spec)
    P = gpu_patch(spec)
    assert P.nelem == Pp.nelem and P.nb == Pp.nb
    assert np.array_equal(P.spans(), Pp.spans())
    dof, n = P.dofmap()
    dref, nref = Pp.dofmap()
    assert n == nref and np.array_equal(dof, dref)


def test_dofmap_trim_compaction_bit_exact():
    spec = PFT.small_patch("square_p3", level=3)
    Pp = port_for(spec)
    rng = np.random.default_rng(5)
    flags = rng.choice([0, 1, 2], size=Pp.nelem, p=[0.5, 0.3, 0.2]).astype(np.uint8)
    flags[:9] = 1  # make sure a whole corner drops out
    P = gpu_patch(spec, trimflags=flags)
    dof, n = P.dofmap()
    dref, nref = Pp.dofmap(flags)
    assert n == nref and n < P.ndof_untrimmed
    keep = flags != 1
    assert np.array_equal(dof[keep], dref[keep])
    assert np.array_equal(dof, dref)


@pytest.mark.parametrize("name", NAMES)
def test_eval_points_corners_and_random(name):
    """Arbitrary element-local points incl. element corners and corners +- 1e-8, the 13-point pattern of
    checkJacobianAndPartialConsistency (gridtests.cpp:514-576)."""
    spec = PFT.small_patch(name)
    O, kind = oracle_for(spec)
    P = gpu_patch(spec)
    rng = np.random.default_rng(42)
    n = 96
    elem = rng.integers(0, P.nelem, n).astype(np.int32)
    elem[:4] = [0, P.nelem - 1, 0, P.nelem - 1]
    xi = rng.random((n, P.dim))
    xi[0:2] = 0.0
    xi[2:4] = 1.0
    xi[4:12] = 1e-8
    xi[12:20] = 1 - 1e-8
    xi[20:28] = (rng.random((8, P.dim)) > 0.5).astype(float)  # random corners
    ref = O.eval_points(elem, xi, order=2, want=EVAL_FIELDS)
    got = P.evalPoints(elem, xi, order=2, want=EVAL_FIELDS)
    for f in EVAL_FIELDS:
        assert relerr(got[f], ref[f]) <= TOL, (f, kind, relerr(got[f], ref[f]))
    # evaluateJacobian column d == partial(e_d) to 1e-14 (the reference's own consistency test)
    for d in range(P.dim):
        al = [0] * P.dim
        al[d] = 1
        part = P.partial(elem, xi, al)
        assert np.abs(part - got["dN"][:, :, d]).max() <= 1e-14 * max(1.0, np.abs(part).max())
    # second derivatives through partial as well
    al = [0] * P.dim
    al[0] = 2
    assert relerr(P.partial(elem, xi, al), got["d2N"][:, :, 0]) <= 1e-13
    assert np.all(np.isfinite(got["d2N"])) and np.all(np.isfinite(got["H"]))  # gridtests.cpp:1085-1124


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("mode", [capi.KERNEL_GENERIC, capi.KERNEL_WARP], ids=["generic", "warp"])
def test_eval_points_kernels_incl_points_outside_the_patch(name, mode):
    """Point lists through the thread-per-point kernel and the warp-per-32-points kernel: several chunks with a tail,
    points grouped by element and scattered, element corners, and points slightly OUTSIDE the patch, where the
    reference's value path clamps (bsplinealgorithms.hh:103) and its derivative path does not -- both kernels must
    reproduce that.  Orders 0, 1, 2, output subsets."""
    spec = PFT.small_patch(name)
    O, kind = oracle_for(spec)
    P = gpu_patch(spec)
    P.setKernel(mode)
    rng = np.random.default_rng(7)
    n = 32 * 3 + 11
    elem = np.sort(rng.integers(0, P.nelem, n)).astype(np.int32)
    elem[-8:] = rng.integers(0, P.nelem, 8)
    xi = rng.random((n, P.dim))
    elem[:6] = [0, 0, 0, P.nelem - 1, P.nelem - 1, P.nelem - 1]
    xi[0], xi[1], xi[2] = 0.0, -0.02, -1e-9
    xi[3], xi[4], xi[5] = 1.0, 1.03, 1 + 1e-9
    ref = O.eval_points(elem, xi, order=2, want=EVAL_FIELDS)
    got = P.evalPoints(elem, xi, order=2, want=EVAL_FIELDS)
    for f in EVAL_FIELDS:
        assert relerr(got[f], ref[f]) <= TOL, (f, kind, relerr(got[f], ref[f]))
    got1 = P.evalPoints(elem, xi, order=1, want=("N", "dN", "x", "Jt", "detJ", "JinvT"))
    for f in ("N", "dN", "x", "Jt", "detJ"):
        assert relerr(got1[f], ref[f]) <= TOL, (f, "order 1")
    got0 = P.evalPoints(elem, xi, order=0)
    assert set(got0) == {"N", "x"} and relerr(got0["N"], ref["N"]) <= TOL and relerr(got0["x"], ref["x"]) <= TOL
    sub = P.evalPoints(elem, xi, order=2, want=("d2N", "detJ"))
    assert np.array_equal(sub["d2N"], got["d2N"]) and np.array_equal(sub["detJ"], got["detJ"])


def test_reference_literals_through_cuda():
    """The reference's known-answer vectors (Appendix B) replayed through the CUDA kernels."""
    n = G["nurbs2d"]
    w = np.ones((3, 7))
    for i in range(3):
        for j in range(3):
            w[j, i] = n["weights_3x3_block"][i][j]
    cp = np.zeros((21, 3))
    cp[:, 2] = w.reshape(-1)
    P = capi.Patch(n["degree"], n["knots"], cp, 2)
    eps = 64 * np.finfo(float).eps
    # element (0,0) spans [0,0.5] x [0,2]
    for v in n["values"]:
        xi = [v["u"][0] / 0.5, v["u"][1] / 2.0]
        R = P.evalPoints([0], [xi], order=0, want=("N",))["N"][0]
        assert np.abs(R - np.array(v["R"])).max() <= eps
        assert abs(R.sum() - 1.0) <= eps
    xi = [[0.0, 0.1 / 2.0]]
    for key, val in n["derivatives_at_0_0.1"].items():
        a, b = map(int, key.split(","))
        got = P.partial([0], xi, [a, b])[0] / (0.5 ** a * 2.0 ** b)  # undo the element scaling: literals are d/du
        val = np.array(val)
        assert np.abs(got - val).max() <= 1e-13 * max(1.0, np.abs(val).max()), key
    # 1-D B-spline literals: weights 1 => NURBS == B-spline; p=3 knots {0,0,0,0,.5,1,1,2,2,2,2}, u=1.45 in span [1,2]
    c = G["bspline1d_ders"]
    P1 = capi.Patch([3], [c["knots"]], np.ones((7, 2)), 1)
    e = 2  # elements: [0,.5],[.5,1],[1,2]
    assert P1.nelem == 3
    for k in range(4):
        got = P1.partial([e], [[0.45]], [k])[0]
        assert np.abs(got - np.array(c["dN"][k])).max() <= 1e-13 * max(1.0, np.abs(c["dN"][k]).max())
    # geometry literals (trimmedgridtests.cpp:38-186); single-element patches => xi == u
    g = G["geometry_curve"]
    Pc = capi.Patch(g["degree"], g["knots"], g["cp"], g["dimworld"])
    for u, x in g["global"]:
        assert np.abs(Pc.evalPoints([0], [u], order=0, want=("x",))["x"][0] - x).max() <= eps * 8
    for u, J in g["jacobianTransposed"]:
        assert np.abs(Pc.evalPoints([0], [u], order=1, want=("Jt",))["Jt"][0] - J).max() <= eps * 64
    s = G["geometry_surface"]
    cps = np.array([[s["cp_table_ij"][i][j] for i in range(3)] for j in range(3)], float).reshape(9, 4)
    Ps = capi.Patch(s["degree"], s["knots"], cps, 3)
    for u, x in s["global"]:
        assert np.abs(Ps.evalPoints([0], [u], order=0, want=("x",))["x"][0] - x).max() <= eps * 4
    for u, J in s["jacobianTransposed"]:
        assert np.abs(Ps.evalPoints([0], [u], order=1, want=("Jt",))["Jt"][0] - J).max() <= eps * 4


def test_torus_invariants_through_cuda():
    """gridtests.cpp:338-363 on the GPU: area = 4 pi^2 r R (1e-4), Gauss-Bonnet (1e-5), curvature bounds."""
    t = G["torus"]
    spec = PFT.small_patch("torus_p2", level=2)
    P = gpu_patch(spec)
    x, w = PD.gaussLegendre01(3)
    area = P.volume([x, x], [w, w])
    assert abs(area - 4 * np.pi ** 2 * t["r"] * t["R"]) < t["area_tol"]
    o = P.evalTensor([x, x], order=2, want=("Jt", "H", "detJ"))
    J, H, det = o["Jt"], o["H"], o["detJ"]
    nrm = np.cross(J[:, 0], J[:, 1])
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    b00, b11, b01 = (H[:, 0] * nrm).sum(1), (H[:, 1] * nrm).sum(1), (H[:, 2] * nrm).sum(1)
    g00, g11, g01 = (J[:, 0] ** 2).sum(1), (J[:, 1] ** 2).sum(1), (J[:, 0] * J[:, 1]).sum(1)
    K = (b00 * b11 - b01 ** 2) / (g00 * g11 - g01 ** 2)
    ww = np.tile(np.outer(w, w).reshape(-1), P.nelem)
    assert abs((K * det * ww).sum()) < t["gauss_bonnet_tol"]
    r, R = t["r"], t["R"]
    assert K.min() >= -1 / (r * (R - r)) - 1e-8 and K.max() <= 1 / (r * (R + r)) + 1e-8


@pytest.mark.parametrize("name,kind", [("plate_hole_p2", capi.ASM_POISSON), ("plate_hole_p2", capi.ASM_MASS),
                                       ("plate_hole_p2", capi.ASM_ELASTICITY), ("cylinder_p3", capi.ASM_POISSON),
                                       ("cylinder_p3", capi.ASM_ELASTICITY), ("cylinder_p4", capi.ASM_POISSON),
                                       ("torus_p2", capi.ASM_POISSON), ("square_p3", capi.ASM_MASS)])
def test_element_matrices_match_oracle(name, kind):
    spec = PFT.small_patch(name)
    Pp = port_for(spec)
    P = gpu_patch(spec)
    rule = [PD.gaussLegendre01(p + 1) for p in spec["degree"]]
    xi1d, w1d = [r[0] for r in rule], [r[1] for r in rule]
    D = None
    if kind == capi.ASM_ELASTICITY:
        E, nu = 1.0, 0.3
        lam, mu = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
        if P.dim == 2:
            D = np.array([[lam + 2 * mu, lam, 0], [lam, lam + 2 * mu, 0], [0, 0, mu]])
        else:
            D = np.zeros((6, 6))
            D[:3, :3] = lam
            D[np.arange(3), np.arange(3)] = lam + 2 * mu
            D[np.arange(3, 6), np.arange(3, 6)] = mu
    ne = min(Pp.nelem, 24)
    Kref, fref = Pp.assemble(kind, xi1d, w1d, D=D, fconst=1.5, elem_end=ne)
    K, f = P.assemble(kind, xi1d, w1d, D=D, fconst=1.5, elem_end=ne)
    assert relerr(K.reshape(ne, -1), Kref.reshape(ne, -1)) <= TOL, relerr(K.reshape(ne, -1), Kref.reshape(ne, -1))
    assert relerr(f, fref) <= TOL
    assert np.abs(K - K.transpose(0, 2, 1)).max() <= 1e-12 * np.abs(K).max()


@pytest.mark.parametrize("name", ["plate_hole_p2", "cylinder_p3", "torus_p2"])
def test_volume_matches_oracle_and_is_reproducible(name):
    spec = PFT.small_patch(name)
    Pp = port_for(spec)
    P = gpu_patch(spec)
    rule = [PD.gaussLegendre01(p + 1) for p in spec["degree"]]
    xi1d, w1d = [r[0] for r in rule], [r[1] for r in rule]
    v = P.volume(xi1d, w1d)
    assert abs(v - Pp.volume(xi1d, w1d)) <= 1e-12 * abs(v)
    assert P.volume(xi1d, w1d) == v  # deterministic tree: bitwise reproducible
    half = P.nelem // 2
    assert abs(P.volume(xi1d, w1d, 0, half) + P.volume(xi1d, w1d, half, P.nelem) - v) <= 1e-13 * abs(v)


def test_full_size_properties_c1():
    """BASELINE config 1 at full size (256 x 256 elements, p=2, 3x3 Gauss): properties that need no oracle run --
    partition of unity, sum of gradients = 0, detJ > 0, area = 16 - pi/4, J consistent with finite differences of x,
    plus the oracle on a random sample of elements."""
    Pd = PD.plateWithHole(7, 8)
    assert Pd.elementsPerDirection == [256, 256]
    P = capi.Patch.fromPatchData(Pd)
    x, w = PD.gaussLegendre01(3)
    o = P.evalTensor([x, x], order=1, want=("N", "dN", "x", "Jt", "detJ"))
    assert o["N"].shape == (256 * 256 * 9, 9)
    assert np.abs(o["N"].sum(1) - 1).max() < 1e-13
    assert np.abs(o["dN"].sum(1)).max() < 1e-12
    assert o["detJ"].min() > 0
    area = P.volume([x, x], [w, w])
    assert abs(area - (16 - np.pi / 4)) < 1e-9
    Pp = PT.PortPatch(Pd.degree, Pd.knotSpans, Pd.cp, 2)
    rng = np.random.default_rng(3)
    for e in rng.integers(0, P.nelem, 40):
        ref = Pp.eval_tensor([x, x], order=1, elem_begin=int(e), elem_end=int(e) + 1)
        sl = slice(int(e) * 9, int(e) * 9 + 9)
        for f in ("N", "dN", "x", "Jt", "detJ"):
            assert relerr(o[f][sl], ref[f]) <= TOL, f


def test_device_pointers_in_place():
    """IGAB_DEVICE outputs: results land in caller-owned device memory on the caller's stream."""
    import torch
    spec = PFT.small_patch("cylinder_p3")
    P = gpu_patch(spec)
    xi1d = gauss(spec)
    npts = P.nelem * 64
    sh = P.shapes(npts)
    out = {f: torch.empty(sh[f], dtype=torch.float64, device="cuda") for f in ("N", "dN", "x", "Jt", "detJ")}
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        P.evalTensor(xi1d, order=1, want=tuple(out), out=out, stream=s.cuda_stream)
    s.synchronize()
    host = P.evalTensor(xi1d, order=1, want=tuple(out))
    for f in out:
        assert np.array_equal(out[f].cpu().numpy(), host[f]), f


def test_error_behaviour():
    spec = PFT.small_patch("plate_hole_p2")
    P = gpu_patch(spec)
    x = PD.gaussLegendre01(3)[0]
    with pytest.raises(capi.IgabError) as ei:
        P.evalTensor([x, x], elem_begin=0, elem_end=P.nelem + 1)
    assert ei.value.status == 1
    with pytest.raises(capi.IgabError):
        P.evalPoints([P.nelem], [[0.5, 0.5]])
    with pytest.raises(capi.IgabError):
        P.evalTensor([x, x], order=3)
    with pytest.raises(capi.IgabError):
        P.partial([0], [[0.5, 0.5]], [70, 0])


def test_degree_elevation_invariance_through_cuda():
    """testCurveHigherOrderDerivatives / testSurfaceHigherOrderDerivatives (gridtests.cpp:1177-1200,1285-1337):
    position, Jacobian and Hessian are invariant under degree elevation (1e-10 / 1e-8 there)."""
    c = G["degree_elevate_curve"]
    base = PD.NurbsPatchData(c["knots"], c["cp"], c["degree"], c["dimworld"])
    u = np.linspace(0, 1, 21)

    def sample(pd):
        P = capi.Patch.fromPatchData(pd)
        # two elements [0,.5],[.5,1]: map patch coordinate u to (element, xi)
        el = np.minimum((u * 2).astype(np.int32), 1)
        xi = (u * 2 - el).reshape(-1, 1)
        o = P.evalPoints(el, xi, order=2, want=("x", "Jt", "H"))
        # undo the element scaling (span length 0.5) to compare in patch coordinates
        return o["x"], o["Jt"] / 0.5, o["H"] / 0.25

    x0, J0, H0 = sample(base)
    for t in range(1, 5):
        x1, J1, H1 = sample(base.degreeElevate(0, t))
        assert np.abs(x1 - x0).max() < 1e-10 and np.abs(J1 - J0).max() < 1e-10 and np.abs(H1 - H0).max() < 1e-8
    # surface: torus, elevate both directions
    s0 = PD.torusSurface(0)
    s1 = s0.degreeElevate(0, 1).degreeElevate(1, 2)
    rng = np.random.default_rng(3)
    el = rng.integers(0, 16, 9).astype(np.int32)
    xi = rng.random((9, 2))
    a = capi.Patch.fromPatchData(s0).evalPoints(el, xi, order=2, want=("x", "Jt", "H"))
    b = capi.Patch.fromPatchData(s1).evalPoints(el, xi, order=2, want=("x", "Jt", "H"))
    for f, tol in (("x", 1e-10), ("Jt", 1e-10), ("H", 1e-8)):
        assert np.abs(a[f] - b[f]).max() < tol, f


def test_flat_plate_second_derivatives():
    """testPlate (gridtests.cpp:1085-1124): on a flat plate in R^3 every N, dN, d2N is finite at the element corners and
    the out-of-plane components of the Hessian vanish."""
    pd = PD.NurbsPatchData([[0, 0, 0, 1, 1, 1], [0, 0, 0, 1, 1, 1]],
                           [(x, y, 0.0, 1.0) for y in (0, .5, 1) for x in (0, .5, 1)], [2, 2], 3).globalRefine(1)
    P = capi.Patch.fromPatchData(pd)
    corners = np.array([[a, b] for a in (0.0, 1.0) for b in (0.0, 1.0)])
    for e in range(P.nelem):
        o = P.evalPoints([e] * 4, corners, order=2)
        for f in ("N", "dN", "d2N"):
            assert np.all(np.isfinite(o[f]))
        assert np.abs(o["H"][:, :, 2]).max() == 0.0
        assert np.abs(o["Jt"][:, :, 2]).max() == 0.0


@pytest.mark.parametrize("name,kind", [("plate_hole_p2", capi.ASM_POISSON), ("torus_p2", capi.ASM_MASS),
                                       ("cylinder_p3", capi.ASM_POISSON), ("plate_hole_p2", capi.ASM_ELASTICITY),
                                       ("mixed_deg_3d", capi.ASM_MASS)])
def test_global_csr_matches_scatter_of_oracle(name, kind):
    """Global matrix = scatter of the element matrices through the DOF map (poisson.py:113-125): CSR gather on the GPU
    against a dense numpy scatter of the oracle's element matrices."""
    spec = PFT.small_patch(name, level=1)
    Pp = port_for(spec)
    P = gpu_patch(spec)
    rule = [PD.gaussLegendre01(p + 1) for p in spec["degree"]]
    xi1d, w1d = [r[0] for r in rule], [r[1] for r in rule]
    ncomp = P.dim if kind == capi.ASM_ELASTICITY else 1
    D = np.array([[1.35, 0.58, 0], [0.58, 1.35, 0], [0, 0, 0.385]]) if kind == capi.ASM_ELASTICITY else None
    Kref, _ = Pp.assemble(kind, xi1d, w1d, D=D)
    dof, ndof = Pp.dofmap()
    n = ndof * ncomp
    dense = np.zeros((n, n))
    gd = (dof[:, :, None] * ncomp + np.arange(ncomp)[None, None, :]).reshape(P.nelem, -1)
    for e in range(P.nelem):
        np.add.at(dense, (gd[e][:, None], gd[e][None, :]), Kref[e])
    K, _ = P.assemble(kind, xi1d, w1d, D=D)
    rowptr, colidx = P.csrPattern(ncomp)
    assert len(rowptr) == n + 1 and rowptr[-1] == len(colidx)
    assert all(np.all(np.diff(colidx[rowptr[r]:rowptr[r + 1]]) > 0) for r in range(0, n, max(1, n // 50)))
    vals = P.csrAssemble(K, ncomp)
    got = np.zeros((n, n))
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    got[rows, colidx] = vals
    assert np.abs(got - dense).max() <= 1e-12 * np.abs(dense).max()
    # everything outside the pattern is structurally zero in the scatter as well
    mask = np.zeros((n, n), bool)
    mask[rows, colidx] = True
    assert np.abs(dense[~mask]).max() == 0.0
    # two element ranges accumulate to the same matrix (the multi-rank path), bitwise reproducible
    half = P.nelem // 2
    v2 = P.csrAssemble(K[:half], ncomp, 0, half)
    v2 = P.csrAssemble(K[half:], ncomp, half, P.nelem, values=v2, accumulate=True)
    assert np.abs(v2 - vals).max() <= 1e-13 * np.abs(vals).max()
    assert np.array_equal(P.csrAssemble(K, ncomp), vals)


def _faces_reference(O, dim, dw, elem, face, xif):
    """numpy restatement of NURBSintersection::outerNormal / unitOuterNormal and the face integration element
    (dune/iga/nurbsintersection.hh:86-160) on top of the ORACLE's x and J^T at the face points."""
    nqf = xif.shape[0]
    el = np.repeat(elem, nqf)
    fc = np.repeat(face, nqf)
    xi = np.zeros((len(el), dim))
    for k in range(len(el)):
        fd, side = fc[k] >> 1, fc[k] & 1
        free = [d for d in range(dim) if d != fd]
        xi[k, fd] = side
        for m, d in enumerate(free):
            xi[k, d] = xif[k % nqf, m]
    o = O.eval_points(el, xi, order=1, want=("x", "Jt"))
    J = o["Jt"]
    n = np.zeros((len(el), dw))
    ds = np.zeros(len(el))
    for k in range(len(el)):
        fi, fd = fc[k], fc[k] >> 1
        if dim == 2 and dw == 2:
            T = J[k, 1 - fd]
            n[k] = (-T[1], T[0]) if fi in (0, 3) else (T[1], -T[0])
            ds[k] = np.linalg.norm(T)
        elif dim == 2:
            N = np.cross(J[k, 0], J[k, 1])
            n[k] = {0: np.cross(N, J[k, 1]), 1: np.cross(J[k, 1], N), 2: np.cross(J[k, 0], N), 3: np.cross(N, J[k, 0])}[fi]
            ds[k] = np.linalg.norm(J[k, 1 - fd])
        else:
            free = [d for d in range(3) if d != fd]
            T0, T1 = J[k, free[0]], J[k, free[1]]
            n[k] = np.cross(T1, T0) if fi in (0, 3, 4) else np.cross(T0, T1)
            ds[k] = np.linalg.norm(np.cross(T0, T1))
    n /= np.linalg.norm(n, axis=1)[:, None]
    return o["x"], n, ds


@pytest.mark.parametrize("name", ["plate_hole_p2", "torus_p2", "cylinder_p3", "mixed_deg_3d"])
def test_faces_match_reference_formulas(name):
    spec = PFT.small_patch(name)
    O, _ = oracle_for(spec)
    P = gpu_patch(spec)
    dim, dw = P.dim, P.dimworld
    rng = np.random.default_rng(11)
    nf = 40
    elem = rng.integers(0, P.nelem, nf).astype(np.int32)
    face = rng.integers(0, 2 * dim, nf).astype(np.int32)
    x1 = PD.gaussLegendre01(3)[0]
    if dim == 2:
        xif = x1.reshape(-1, 1)
        rule = [x1]
    else:
        xif = np.array([[x1[q % 3], x1[q // 3]] for q in range(9)])
        rule = [x1, x1]
    got = P.evalFaces(elem, face, rule)
    xr, nr, dsr = _faces_reference(O, dim, dw, elem, face, xif)
    assert relerr(got["x"], xr) <= TOL
    assert np.abs(got["normal"] - nr).max() <= 1e-12
    assert relerr(got["dS"].reshape(-1, 1), dsr.reshape(-1, 1)) <= TOL


def test_cylinder_boundary_normals_and_divergence_theorem():
    """Thick quarter cylinder: radial normals on the curved faces, +-e_z on the end caps, area of the outer face, and
    sum over the boundary of x.n dS = 3 * volume."""
    pd = PD.thickCylinder(3, (2, 2, 2))  # 4 x 4 x 4 elements
    P = capi.Patch.fromPatchData(pd)
    n = 4
    x1, w1 = PD.gaussLegendre01(4)
    ww = np.outer(w1, w1).reshape(-1)  # first free direction fastest
    total = 0.0
    # the reference's cross-product formulas (nurbsintersection.hh:139-152) give the OUTER normal for a positively
    # oriented parametrisation; (angle, radius, axis) is left-handed, so every normal comes out flipped: orient = -1
    orient = np.sign(np.linalg.det(P.evalPoints([0], [[.5, .5, .5]], order=1, want=("Jt",))["Jt"][0]))
    assert orient == -1.0
    for d in range(3):
        for s in (0, 1):
            idx = [np.arange(n)] * 3
            idx[d] = np.array([0 if s == 0 else n - 1])
            e0, e1, e2 = np.meshgrid(idx[0], idx[1], idx[2], indexing="ij")
            elem = (e0 + n * (e1 + n * e2)).reshape(-1).astype(np.int32)
            face = np.full(len(elem), 2 * d + s, np.int32)
            o = P.evalFaces(elem, face, [x1, x1])
            w = np.tile(ww, len(elem))
            total += orient * ((o["x"] * o["normal"]).sum(1) * o["dS"] * w).sum()
            r = np.hypot(o["x"][:, 0], o["x"][:, 1])
            if d == 1:  # radial direction: inner (s=0) and outer (s=1) curved faces
                radial = np.stack([o["x"][:, 0] / r, o["x"][:, 1] / r, 0 * r], 1)
                assert np.abs(orient * o["normal"] - (radial if s == 1 else -radial)).max() < 1e-12
                assert abs((o["dS"] * w).sum() - np.pi / 2 * (2.0 if s == 1 else 1.0) * 3.0) < 1e-10
            if d == 2:
                assert np.abs(orient * o["normal"] - np.array([0, 0, 1.0 if s == 1 else -1.0])).max() < 1e-12
    vol = np.pi / 4 * 3.0 * 3.0
    assert abs(total - 3 * vol) < 1e-9


def test_refinement_on_gpu_matches_host_and_reference():
    """igab_refine_apply (control-net side of knotRefinement / degreeElevate, nurbsalgorithms.hh:264-528) against the
    host operators and, when available, the reference's own refinement code."""
    import refbind as RF
    cw = 1.0 / np.sqrt(2.0)
    cp0 = []
    for z in (0.0, 3.0):
        for r in (1.0, 2.0):
            cp0 += [(r, 0, z, 1.0), (r, r, z, cw), (0, r, z, 1.0)]
    P0 = PD.NurbsPatchData([[0, 0, 0, 1, 1, 1], [0, 0, 1, 1], [0, 0, 1, 1]], cp0, [2, 1, 1], 3)  # coarse (2,1,1)
    cur_gpu, cur_host = P0, P0
    for d, t in ((0, 1), (1, 2), (2, 2)):  # elevate to (3,3,3)
        U, E = PD.degreeElevationOperator(cur_host.knotSpans[d], cur_host.degree[d], t)
        cur_gpu = capi.refineApply(cur_gpu, d, E, U, cur_gpu.degree[d] + t)
        cur_host = cur_host.degreeElevate(d, t)
    for d in range(3):  # refine 2^3 per direction
        nk = PD.generateRefinedKnots(cur_host.knotSpans, d, 3)
        U, T = PD.knotInsertionOperator(cur_host.knotSpans[d], cur_host.degree[d], nk)
        cur_gpu = capi.refineApply(cur_gpu, d, T, U)
        cur_host = cur_host.knotRefinement(nk, d)
    assert cur_gpu.ncp == cur_host.ncp == [11, 11, 11]
    assert np.abs(cur_gpu.cp - cur_host.cp).max() <= 1e-14
    if RF.available():
        R = RF.RefPatch.create([2, 1, 1], [[0, 0, 0, 1, 1, 1], [0, 0, 1, 1], [0, 0, 1, 1]], cp0, 3)
        R = R.degree_elevate(0, 1).degree_elevate(1, 2).degree_elevate(2, 2)
        for d in range(3):
            R = R.refine_uniform(d, 3)
        assert np.abs(cur_gpu.cp - R.cp).max() <= 1e-12


@pytest.mark.parametrize("name", ["cylinder_p3", "cylinder_p4", "plate_hole_p2", "torus_p2"])
@pytest.mark.parametrize("want", [("N",), ("dN",), ("detJ",), ("x", "JinvT"), ("Jt", "N"), ("dN", "detJ", "JinvT")])
def test_output_subsets(name, want):
    """Any subset of the output set may be requested (nullable members of igab_eval_out)."""
    spec = PFT.small_patch(name)
    Pp = port_for(spec)
    P = gpu_patch(spec)
    xi1d = gauss(spec)
    ref = Pp.eval_tensor(xi1d, order=1, want=want)
    got = P.evalTensor(xi1d, order=1, want=want)
    assert set(got) == set(want)
    for f in want:
        assert relerr(got[f], ref[f]) <= TOL, f


@pytest.mark.parametrize("name", ["cylinder_p3", "cylinder_p4", "square_p3"])
def test_misaligned_device_outputs(name):
    """Device output pointers that are only 8-byte aligned (views at an odd double offset): the 16-byte paths
    (TMA bulk stores, vector stores) must fall back, results unchanged."""
    import torch
    spec = PFT.small_patch(name)
    P = gpu_patch(spec)
    xi1d = gauss(spec)
    nq = int(np.prod([len(x) for x in xi1d]))
    npts = P.nelem * nq
    sh = P.shapes(npts)
    fields = ("N", "dN", "x", "Jt", "detJ", "JinvT")
    host = P.evalTensor(xi1d, order=1, want=fields)
    out = {}
    for f in fields:
        buf = torch.empty(int(np.prod(sh[f])) + 1, dtype=torch.float64, device="cuda")
        out[f] = buf[1:].view(sh[f])
        assert out[f].data_ptr() % 16 == 8
    P.evalTensor(xi1d, order=1, want=fields, out=out, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for f in fields:
        assert np.array_equal(out[f].cpu().numpy(), host[f]), f


def test_surface_forms_match_oracle_and_gauss_bonnet():
    """normal / unitNormal / metric / secondFundamentalForm / gaussianCurvature (nurbsgeometry.hh:216-255) on the device
    from the Jt / H arrays of the evaluation kernel: oracle parity, Gauss-Bonnet (gridtests.cpp:348-363), host and
    device-resident arrays give the same bits."""
    import torch
    t = G["torus"]
    spec = PFT.small_patch("torus_p2", level=2)
    P = gpu_patch(spec)
    x, w = PD.gaussLegendre01(3)
    o = P.evalTensor([x, x], order=2, want=("Jt", "H", "detJ"))
    got = capi.surfaceForms(o["Jt"], o["H"])
    ref = PT.surface_forms(o["Jt"], o["H"])
    for k in ("normal", "unitNormal", "metric", "secondFF", "gaussK"):
        assert relerr(got[k], ref[k]) <= TOL, (k, relerr(got[k], ref[k]))
    ww = np.tile(np.outer(w, w).reshape(-1), P.nelem)
    assert abs((got["gaussK"] * o["detJ"] * ww).sum()) < t["gauss_bonnet_tol"]
    dev = capi.surfaceForms(torch.from_numpy(o["Jt"]).cuda(), torch.from_numpy(o["H"]).cuda())
    torch.cuda.synchronize()
    for k in got:
        assert np.array_equal(dev[k].cpu().numpy(), got[k]), k
    # metric alone works for every (dim, dimworld); the surface quantities do not
    spec3 = PFT.small_patch("cylinder_p3")
    P3 = gpu_patch(spec3)
    g3 = gauss(spec3)
    J3 = P3.evalTensor(g3, order=1, want=("Jt",))["Jt"]
    m3 = capi.surfaceForms(J3, want=("metric",))["metric"]
    assert relerr(m3, np.einsum("pac,pbc->pab", J3, J3)) <= TOL
    with pytest.raises(capi.IgabError) as ei:
        capi.surfaceForms(J3, want=("normal",))
    assert ei.value.status == capi.UNSUPPORTED
    with pytest.raises(capi.IgabError):
        capi.surfaceForms(o["Jt"], None, want=("gaussK",))


def _literal_patches():
    c = G["geometry_curve"]
    s = G["geometry_surface"]
    cps = np.array([[s["cp_table_ij"][i][j] for i in range(3)] for j in range(3)], float).reshape(9, 4)
    return (c, np.array(c["cp"], float)), (s, cps)


def test_geometry_local_literals_through_cuda():
    """trimmedgridtests.cpp:90-95,157-162: geometry.local(x) == u for the curve and the surface."""
    for spec, cp in _literal_patches():
        P = capi.Patch(spec["degree"], spec["knots"], cp, spec["dimworld"])
        for x, u in spec["local"]:
            xi, flag = P.local([0], [x])
            assert flag[0] == 0 and np.abs(xi[0] - np.array(u, float)).max() <= 64 * np.finfo(float).eps, (x, u, xi)


@pytest.mark.parametrize("name", ["square_p3", "cylinder_p3", "curve_p3", "torus_p2", "mixed_deg_3d"])
def test_geometry_local_matches_oracle(name):
    """Batched NURBSGeometry::local on the device vs the port oracle: images of random element-local points (round trip,
    gridtests.cpp:1085-1100: corners to 1e-14) and, for curves / surfaces, off-manifold targets through the trust-region
    projection (closestpointprojection.hh).  Tolerance on xi: 1e-12 for on-manifold targets; the projection stops at
    |grad E| < 1e-10; off-manifold targets whose projection ends inside the element are compared at 1e-12 as well."""
    spec = PFT.small_patch(name)
    Pp = port_for(spec)
    P = gpu_patch(spec)
    rng = np.random.default_rng(17)
    dim, dw = Pp.dim, Pp.dimworld
    corners = np.array([[(k >> d) & 1 for d in range(dim)] for k in range(2 ** dim)], float)
    per = len(corners) + 4
    elem = np.repeat(np.arange(Pp.nelem, dtype=np.int32), per)
    xi = np.concatenate([np.concatenate([corners, rng.uniform(0.05, 0.95, (4, dim))]) for _ in range(Pp.nelem)])
    X = Pp.eval_points(elem, xi, order=0, want=("x",))["x"]
    got, flag = P.local(elem, X)
    ref, rflag = Pp.local(elem, X)
    assert np.array_equal(flag, rflag) and (flag == 0).all()
    assert np.abs(got - ref).max() <= 1e-12
    if dim == dw:  # Newton branch: exact round trip; the projection may stop on the element boundary it overshot to
        assert np.abs(got - xi).max() <= 1e-12
    if dim < dw:
        Xo = X + rng.normal(0, 0.05, X.shape)
        got, flag = P.local(elem, Xo)
        ref, rflag = Pp.local(elem, Xo)
        same = flag == rflag
        assert same.mean() >= 0.97
        ok = same & (flag == 0)
        # Iterates the reference clamps onto an element edge never satisfy |grad E| < 1e-10 and wander for the full 100
        # iterations; where they end is decided by last-bit comparisons (`energy > oldEnergy`,
        # closestpointprojection.hh:100-101; perturbing the target by 1e-15 moves the ORACLE's own answer by up to
        # 1e-3 there).  Projections that end inside the element are stable and must agree to 1e-12.
        inner = ok & ((ref > 0) & (ref < 1)).all(1) & ((got > 0) & (got < 1)).all(1)
        assert inner.sum() >= 0.5 * ok.sum()
        assert np.abs(got[inner] - ref[inner]).max() <= 1e-12
        edge = ok & ~inner
        xg = Pp.eval_points(elem[edge], got[edge], order=0, want=("x",))["x"]
        xr = Pp.eval_points(elem[edge], ref[edge], order=0, want=("x",))["x"]
        dg, dr = np.linalg.norm(xg - Xo[edge], axis=1), np.linalg.norm(xr - Xo[edge], axis=1)
        assert np.abs(dg - dr).max() <= 1e-4  # same distance to the target even where the parameter wandered
        assert np.isnan(got[flag == 2]).all()


def test_geometry_local_errors_and_singular_jacobian():
    spec = PFT.small_patch("square_p3")
    P = gpu_patch(spec)
    with pytest.raises(capi.IgabError) as ei:
        P.local([P.nelem], [[0.0, 0.0]])
    assert ei.value.status == capi.INVALID_ARGUMENT
    xi, flag = P.local(np.zeros(0, np.int32), np.zeros((0, 2)))
    assert xi.shape == (0, 2)
    # a degenerate element (all control points equal): singular Jacobian -> numeric_limits::max (nurbsgeometry.hh:163-164)
    k = [0, 0, 1, 1.0]
    Pd = capi.Patch([1, 1], [k, k], np.array([[1, 1, 1.0]] * 4), 2)
    xi, flag = Pd.local([0], [[2.0, 2.0]])
    assert flag[0] == 1 and (xi[0] == np.finfo(float).max).all()


@pytest.mark.parametrize("name,ncomp", [("curve_p3", 1), ("plate_hole_p2", 2), ("cylinder_p3", 1), ("cylinder_p4", 3),
                                        ("torus_p2", 1), ("mixed_deg_3d", 2)])
def test_field_norms_match_oracle(name, ncomp):
    """igab_reduce_norms (fused evaluation + deterministic reduction) vs the port oracle's per-point loop: squared L2
    norm and energy of a random field, whole patch and a sub-range; device-resident coefficients give the same bits;
    two calls give the same bits."""
    import torch
    spec = PFT.small_patch(name)
    Pp = port_for(spec)
    P = gpu_patch(spec)
    rule = [PD.gaussLegendre01(p + 1) for p in spec["degree"]]
    xi1d, w1d = [r[0] for r in rule], [r[1] for r in rule]
    rng = np.random.default_rng(23)
    u = rng.normal(size=(Pp.cp.shape[0], ncomp))
    ref = Pp.norms(xi1d, w1d, u, ncomp=ncomp)
    got = P.norms(xi1d, w1d, u, ncomp=ncomp)
    assert np.abs(got - ref).max() <= TOL * np.abs(ref).max(), (got, ref)
    assert np.array_equal(got, P.norms(xi1d, w1d, u, ncomp=ncomp))
    assert np.array_equal(got, P.norms(xi1d, w1d, torch.from_numpy(u).cuda(), ncomp=ncomp))
    eb, ee = Pp.nelem // 3, Pp.nelem - 1
    ref = Pp.norms(xi1d, w1d, u, ncomp=ncomp, elem_begin=eb, elem_end=ee)
    got = P.norms(xi1d, w1d, u, ncomp=ncomp, elem_begin=eb, elem_end=ee)
    assert np.abs(got - ref).max() <= TOL * np.abs(ref).max()
    assert (P.norms(xi1d, w1d, u, ncomp=ncomp, elem_begin=2, elem_end=2) == 0).all()
    # constant field: L2^2 = volume, energy = 0
    one = P.norms(xi1d, w1d, np.ones(Pp.cp.shape[0]))
    assert abs(one[0] - P.volume(xi1d, w1d)) <= 1e-13 * one[0] and abs(one[1]) <= 1e-18 * max(1.0, one[0])


def test_field_norms_are_the_quadratic_forms_of_the_assembled_matrices():
    """energy = u^T K u and L2^2 = u^T M u with K, M gathered into the global CSR by the CUDA path itself."""
    spec = PFT.small_patch("cylinder_p3")
    P = gpu_patch(spec)
    rule = [PD.gaussLegendre01(p + 1) for p in spec["degree"]]
    xi1d, w1d = [r[0] for r in rule], [r[1] for r in rule]
    rowptr, colidx = P.csrPattern()
    u = np.random.default_rng(3).normal(size=len(rowptr) - 1)
    got = P.norms(xi1d, w1d, u)
    for kind, val in ((capi.ASM_MASS, got[0]), (capi.ASM_POISSON, got[1])):
        Ke, _ = P.assemble(kind, xi1d, w1d, want_fe=False)
        vals = P.csrAssemble(Ke)
        Au = np.add.reduceat(vals * u[colidx], rowptr[:-1])
        assert abs(u @ Au - val) <= 1e-11 * abs(val)


def test_field_norms_errors():
    spec = PFT.small_patch("plate_hole_p2")
    rule = [PD.gaussLegendre01(3) for _ in range(2)]
    xi1d, w1d = [r[0] for r in rule], [r[1] for r in rule]
    P = gpu_patch(spec)
    n = P.dofmap()[1]
    with pytest.raises(capi.IgabError) as ei:
        P.norms(xi1d, w1d, np.zeros((n, 4)), ncomp=4)
    assert ei.value.status == capi.INVALID_ARGUMENT
    with pytest.raises(capi.IgabError) as ei:
        P.norms(xi1d, w1d, np.zeros(n), elem_begin=0, elem_end=P.nelem + 1)
    assert ei.value.status == capi.INVALID_ARGUMENT
    # a patch with trim flags takes u in the trimmed basis' numbering: the empty element is skipped, the norm of the
    # constant field 1 is the area of the remaining elements
    flags = np.zeros(P.nelem, np.uint8)
    flags[0] = 1
    Pt = gpu_patch(spec, trimflags=flags)
    nt = Pt.dofmap()[1]
    det = Pt.evalTensor(xi1d, order=1, want=("detJ",))["detJ"].reshape(P.nelem, -1)
    area = (det[1:] * np.outer(w1d[1], w1d[0]).reshape(-1)).sum()
    r = Pt.norms(xi1d, w1d, np.ones(nt))
    assert abs(r[0] - area) <= 1e-12 * area and abs(r[1]) <= 1e-20 + 1e-12 * area


@pytest.mark.parametrize("name", ["cylinder_p3", "cylinder_p4"])
@pytest.mark.parametrize("material", ["isotropic", "anisotropic", "nonsymmetric"])
def test_elasticity_fused_kernel_materials(name, material):
    """3-D elasticity element matrices of the fused nine-Gram-block kernel for the three coefficient cases: isotropic
    (the specialised 21-coefficient combination), fully populated symmetric D (general 81-coefficient combination) and a
    non-symmetric D (mirror block needs the transposed coefficients)."""
    spec = PFT.small_patch(name)
    Pp = port_for(spec)
    P = gpu_patch(spec)
    rule = [PD.gaussLegendre01(p + 1) for p in spec["degree"]]
    xi1d, w1d = [r[0] for r in rule], [r[1] for r in rule]
    rng = np.random.default_rng(31)
    lam, mu = 0.3 / (1.3 * 0.4), 1 / 2.6
    D = np.zeros((6, 6))
    D[:3, :3] = lam
    D[np.arange(3), np.arange(3)] = lam + 2 * mu
    D[np.arange(3, 6), np.arange(3, 6)] = mu
    if material != "isotropic":
        A = rng.normal(size=(6, 6))
        D = D + 0.1 * (A @ A.T if material == "anisotropic" else A)
    ne = min(Pp.nelem, 6)
    Kref, _ = Pp.assemble(capi.ASM_ELASTICITY, xi1d, w1d, D=D, elem_end=ne, want_fe=False)
    K, _ = P.assemble(capi.ASM_ELASTICITY, xi1d, w1d, D=D, elem_end=ne, want_fe=False)
    assert relerr(K.reshape(ne, -1), Kref.reshape(ne, -1)) <= TOL, relerr(K.reshape(ne, -1), Kref.reshape(ne, -1))
    sym = np.abs(K - K.transpose(0, 2, 1)).max() / np.abs(K).max()
    assert (sym > 1e-6) if material == "nonsymmetric" else (sym <= 1e-12)


@pytest.mark.parametrize("p,extra", [(3, 0), (3, 1), (2, 0), (2, 1)])
def test_3d_second_derivative_tensor_kernel(p, extra):
    """eval_sf3d_o2_kernel (3-D tensor rules with second derivatives, p = 2, 3; Q = p+1, p+2) against the oracle: every
    field, an element sub-range, and output subsets (the kernel skips phases whose outputs are absent)."""
    pd = PD.thickCylinder(p, (1, 1, 2))
    spec = dict(degree=pd.degree, knots=pd.knotSpans, cp=pd.cp, dimworld=3)
    Pp = port_for(spec)
    P = gpu_patch(spec)
    xi1d = [PD.gaussLegendre01(p + 1 + extra)[0]] * 3
    ref = Pp.eval_tensor(xi1d, order=2, want=EVAL_FIELDS + ("JinvT",))
    got = P.evalTensor(xi1d, order=2, want=EVAL_FIELDS + ("JinvT",))
    for f in EVAL_FIELDS + ("JinvT",):
        assert relerr(got[f], ref[f]) <= TOL, (f, relerr(got[f], ref[f]))
    P.setKernel(capi.KERNEL_GENERIC)
    gen = P.evalTensor(xi1d, order=2, want=EVAL_FIELDS)
    P.setKernel(capi.KERNEL_AUTO)
    for f in EVAL_FIELDS:
        assert relerr(got[f], gen[f]) <= TOL
    eb, ee = 3, Pp.nelem - 2
    nq = len(xi1d[0]) ** 3
    sub = P.evalTensor(xi1d, order=2, elem_begin=eb, elem_end=ee, want=("d2N", "H"))
    assert set(sub) == {"d2N", "H"}
    assert np.array_equal(sub["d2N"], got["d2N"][eb * nq:ee * nq]) and np.array_equal(sub["H"], got["H"][eb * nq:ee * nq])
    only = P.evalTensor(xi1d, order=2, want=("N", "d2N"))
    assert np.array_equal(only["N"], got["N"]) and np.array_equal(only["d2N"], got["d2N"])
    # second derivatives of a partition of unity vanish
    assert np.abs(got["d2N"].sum(1)).max() < 1e-9 * max(1.0, np.abs(got["d2N"]).max())


def test_element_centres_after_refinement_through_cuda():
    """trimmedgridtests.cpp:197-217: geometry().center() of the four elements of the once-refined bilinear unit square,
    compared with == as the reference does; the refinement itself runs on the device (igab_refine_apply)."""
    c = G["element_centres"]
    pd = PD.NurbsPatchData(c["knots"], c["cp"], c["degree"], c["dimworld"]).globalRefine(c["refine_level"])
    P = capi.Patch.fromPatchData(pd)
    assert P.nelem == 4
    x = P.evalPoints(np.arange(4, dtype=np.int32), np.full((4, 2), 0.5), order=0, want=("x",))["x"]
    assert np.array_equal(x, np.array(c["centres"]))
    xt = P.evalTensor([np.array([0.5])] * 2, order=0, want=("x",))["x"]
    assert np.array_equal(xt, np.array(c["centres"]))


def test_poisson_global_assembler_on_the_device():
    """The whole globalAssembler of dune/python/test/poisson.py:84-127 on the device: element matrices and load vectors,
    CSR gather, load-vector gather, Dirichlet rows (first floor(sqrt(N)) DOFs, identity rows, columns untouched) -- against
    the same loop written with the oracle's element matrices and a dense scatter; then both systems are solved."""
    import math
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    spec = PFT.small_patch("plate_hole_p2")
    Pp = port_for(spec)
    P = gpu_patch(spec)
    rule = [PD.gaussLegendre01(p + 1) for p in spec["degree"]]
    xi1d, w1d = [r[0] for r in rule], [r[1] for r in rule]
    Ke, fe = P.assemble(capi.ASM_POISSON, xi1d, w1d, fconst=1.0)
    rowptr, colidx = P.csrPattern()
    vals = P.csrAssemble(Ke)
    b = P.rhsAssemble(fe)
    N = len(rowptr) - 1
    mask = np.zeros(N, dtype=np.uint8)
    mask[:math.floor(math.sqrt(N))] = 1
    vals, b = P.csrDirichlet(mask, vals, b=b)
    A = sp.csr_matrix((vals, colidx, rowptr), shape=(N, N))
    # the reference loop with the oracle's element matrices
    Kref, fref = Pp.assemble(capi.ASM_POISSON, xi1d, w1d, fconst=1.0)
    dof, _ = Pp.dofmap()
    Ad, bd = np.zeros((N, N)), np.zeros(N)
    for e in range(Pp.nelem):
        for i in range(Pp.nb):
            gi = dof[e, i]
            if mask[gi]:
                Ad[gi, gi] = 1.0
                bd[gi] = 0.0
            else:
                bd[gi] += fref[e, i]
                Ad[gi, dof[e]] += Kref[e, i]
    assert np.abs(A.toarray() - Ad).max() <= 1e-12 * np.abs(Ad).max()
    assert np.abs(b - bd).max() <= 1e-12 * np.abs(bd).max()
    x = spla.spsolve(A.tocsc(), b)
    xd = np.linalg.solve(Ad, bd)
    assert np.abs(x - xd).max() <= 1e-9 * np.abs(xd).max()
    assert np.abs(x[:math.floor(math.sqrt(N))]).max() == 0.0
    # device-resident arrays, two element ranges accumulated: same bits as the single call
    import torch
    h = P.nelem // 2
    fd = torch.from_numpy(fe).cuda()
    b2 = P.rhsAssemble(fd[:h].contiguous(), elem_end=h)
    b2 = P.rhsAssemble(fd[h:].contiguous(), elem_begin=h, b=b2, accumulate=True)
    assert np.abs(b2.cpu().numpy() - P.rhsAssemble(fe)).max() <= 1e-14 * np.abs(bd).max()


def _random_patch(rng, dim, dw, degree):
    """Open knot vectors with random interior knots of random multiplicity (<= p: C^0 lines included), random positive
    weights, control points = a perturbed lattice (so that detJ stays away from 0)."""
    knots, ncp = [], []
    for p in degree:
        inner = []
        for u in np.sort(rng.uniform(0.1, 0.9, rng.integers(1, 4))):
            inner += [float(u)] * int(rng.integers(1, p + 1))
        U = np.array([0.0] * (p + 1) + inner + [1.0] * (p + 1))
        knots.append(U)
        ncp.append(len(U) - p - 1)
    grids = np.meshgrid(*[np.linspace(0, 1 + 0.3 * d, n) for d, n in enumerate(ncp)], indexing="ij")
    pts = np.stack([g.transpose(*range(dim - 1, -1, -1)).reshape(-1) for g in grids], axis=1)  # first direction fastest
    cp = np.zeros((pts.shape[0], dw + 1))
    cp[:, :dim] = pts
    cp[:, :dw] += rng.normal(0, 0.02, (pts.shape[0], dw))
    if dw > dim:
        cp[:, dim] += 0.3 * np.sin(2 * pts[:, 0]) * (pts[:, -1] + 0.5)
    cp[:, dw] = rng.uniform(0.5, 2.0, pts.shape[0])
    return dict(degree=list(degree), knots=knots, cp=cp, dimworld=dw)


@pytest.mark.parametrize("case", [(1, 2, (3,)), (1, 3, (2,)), (2, 2, (2, 2)), (2, 2, (3, 3)), (2, 2, (1, 1)), (2, 2, (3, 2)),
                                  (2, 3, (2, 2)), (2, 3, (3, 3)), (3, 3, (2, 2, 2)), (3, 3, (3, 3, 3)), (3, 3, (4, 4, 4)),
                                  (3, 3, (1, 2, 3)), (3, 3, (1, 1, 1))],
                         ids=lambda c: "d%dw%d_p%s" % (c[0], c[1], "".join(map(str, c[2]))))
def test_random_patches_with_repeated_knots(case):
    """Seeded random patches (repeated interior knots up to multiplicity p, random weights) through the tuned kernels
    (auto), the thread-per-point kernel and the warp-per-element kernel against the oracle, all fields up to second order; spans and DOF maps bit-exact."""
    dim, dw, degree = case
    rng = np.random.default_rng(1000 + 17 * dim + sum(degree) + 100 * dw)
    spec = _random_patch(rng, dim, dw, degree)
    O, kind = oracle_for(spec)
    Pp = port_for(spec)
    P = gpu_patch(spec)
    assert np.array_equal(P.spans(), Pp.spans())
    assert np.array_equal(P.dofmap()[0], Pp.dofmap()[0])
    for extra in (0, 1):
        xi1d = [PD.gaussLegendre01(p + 1 + extra)[0] for p in degree]
        ref = O.eval_tensor(xi1d, order=2, want=EVAL_FIELDS)
        assert ref["detJ"].min() > 1e-3
        for mode in (capi.KERNEL_AUTO, capi.KERNEL_GENERIC, capi.KERNEL_WARP):
            P.setKernel(mode)
            got = P.evalTensor(xi1d, order=2, want=EVAL_FIELDS)
            for f in EVAL_FIELDS:
                assert relerr(got[f], ref[f]) <= TOL, (f, mode, extra, relerr(got[f], ref[f]))
            got1 = P.evalTensor(xi1d, order=1, want=("N", "dN", "x", "Jt", "detJ", "JinvT"))
            for f in ("N", "dN", "x", "Jt", "detJ"):
                assert relerr(got1[f], ref[f]) <= TOL, (f, mode, extra, "order1")


@pytest.mark.parametrize("case", [(1, 1, (5,), (7,)), (1, 3, (8,), (33,)), (2, 2, (4, 2), (3, 6)), (2, 3, (5, 5), (7, 5)),
                                  (3, 3, (2, 1, 1), (3, 2, 2)), (3, 3, (5, 2, 3), (2, 7, 3)), (3, 3, (4, 4, 4), (6, 6, 6)),
                                  (3, 3, (6, 6, 6), (2, 2, 2)), (3, 3, (3, 3, 3), (1, 1, 1))],
                         ids=lambda c: "d%dw%d_p%s_q%s" % (c[0], c[1], "".join(map(str, c[2])), "x".join(map(str, c[3]))))
def test_warp_per_element_kernel_general_degrees_and_rules(case):
    """The general warp-per-element kernel (eval_warp.cu) on what no tuned kernel covers: degrees up to 8, mixed degrees,
    rules with a different number of points per direction (fewer and more than 32 points per element, chunk tails),
    1-D / 2-D / 3-D, against the oracle and the thread-per-point kernel; then sub-ranges and output subsets.  AUTO must
    route these configurations to it (same bits as WARP)."""
    dim, dw, degree, nq1 = case
    rng = np.random.default_rng(7000 + 31 * dim + sum(degree) + 100 * dw + sum(nq1))
    spec = _random_patch(rng, dim, dw, degree)
    Pp = port_for(spec)
    P = gpu_patch(spec)
    xi1d = [PD.gaussLegendre01(q)[0] for q in nq1]
    fields1 = ("N", "dN", "x", "Jt", "detJ", "JinvT")
    ref = Pp.eval_tensor(xi1d, order=2, want=EVAL_FIELDS)
    P.setKernel(capi.KERNEL_WARP)
    got = P.evalTensor(xi1d, order=2, want=EVAL_FIELDS)
    for f in EVAL_FIELDS:
        assert relerr(got[f], ref[f]) <= TOL, (f, relerr(got[f], ref[f]))
    assert np.abs(got["N"].sum(1) - 1.0).max() < 1e-13
    P.setKernel(capi.KERNEL_AUTO)
    auto = P.evalTensor(xi1d, order=2, want=EVAL_FIELDS)
    if not (dim == 3 and len(set(degree)) == 1 and len(set(nq1)) == 1 and degree[0] in (2, 3)):  # tuned order-2 kernel
        for f in EVAL_FIELDS:
            assert np.array_equal(auto[f], got[f]), f
    # order 1 and 0, a sub-range that does not start at 0, JinvT
    eb, ee = Pp.nelem // 3, Pp.nelem
    P.setKernel(capi.KERNEL_WARP)
    ref1 = Pp.eval_tensor(xi1d, order=1, elem_begin=eb, elem_end=ee)
    got1 = P.evalTensor(xi1d, order=1, elem_begin=eb, elem_end=ee)
    for f in fields1:
        assert relerr(got1[f], ref1[f]) <= TOL, (f, "order1", relerr(got1[f], ref1[f]))
    got0 = P.evalTensor(xi1d, order=0, elem_begin=eb, elem_end=ee)
    assert set(got0) == {"N", "x"}
    assert np.array_equal(got0["N"], got1["N"]) and np.array_equal(got0["x"], got1["x"])
    # output subsets give the same bits as the full set
    for want in (("dN",), ("detJ",), ("x", "JinvT"), ("N", "Jt")):
        sub = P.evalTensor(xi1d, order=1, elem_begin=eb, elem_end=ee, want=want)
        assert set(sub) == set(want)
        for f in want:
            assert np.array_equal(sub[f], got1[f]), (want, f)
    sub2 = P.evalTensor(xi1d, order=2, want=("d2N", "H"))
    assert np.array_equal(sub2["d2N"], got["d2N"]) and np.array_equal(sub2["H"], got["H"])
    assert P.evalTensor(xi1d, order=1, elem_begin=1, elem_end=1)["N"].shape[0] == 0


@pytest.mark.parametrize("p", [1, 2, 3, 4])
@pytest.mark.parametrize("dw,kind", [(2, capi.ASM_POISSON), (2, capi.ASM_MASS), (2, capi.ASM_ELASTICITY),
                                     (3, capi.ASM_POISSON), (3, capi.ASM_MASS)])
def test_fused_2d_element_matrices(p, dw, kind):
    """assemble2d_kernel (2-D patches, p = 1..3 and p = 4 for the scalar problems, fused evaluation + contraction) against the oracle and against the
    general B-stack path, on a random patch with repeated knots; odd element ranges exercise the partial last group of
    a warp; a non-symmetric D exercises the mirrored coefficients of the elasticity combination."""
    rng = np.random.default_rng(50 + 7 * p + dw + 3 * kind)
    spec = _random_patch(rng, 2, dw, (p, p))
    Pp = port_for(spec)
    P = gpu_patch(spec)
    D = None
    if kind == capi.ASM_ELASTICITY:
        D = np.array([[1.35, 0.58, 0.02], [0.58, 1.35, -0.03], [0.05, 0.01, 0.385]])
    # reduced (p points) and over-integrated (p+2) rules are fused as well; poisson.py:44 itself uses 2 points
    for q in (max(p, 2), p + 2):
        xq, wq = PD.gaussLegendre01(q)
        Kref, fref = Pp.assemble(kind, [xq] * 2, [wq] * 2, D=D, fconst=0.7)
        K, f = P.assemble(kind, [xq] * 2, [wq] * 2, D=D, fconst=0.7)
        assert relerr(K.reshape(len(K), -1), Kref.reshape(len(K), -1)) <= TOL, (q, relerr(K.reshape(len(K), -1), Kref.reshape(len(K), -1)))
        assert relerr(f, fref) <= TOL
    rule = [PD.gaussLegendre01(p + 1)] * 2
    xi1d, w1d = [r[0] for r in rule], [r[1] for r in rule]
    for eb, ee in ((0, Pp.nelem), (1, Pp.nelem - 1), (2, 3), (3, 3)):
        Kref, fref = Pp.assemble(kind, xi1d, w1d, D=D, fconst=0.7, elem_begin=eb, elem_end=ee)
        K, f = P.assemble(kind, xi1d, w1d, D=D, fconst=0.7, elem_begin=eb, elem_end=ee)
        assert K.shape == Kref.shape
        if ee > eb:
            assert relerr(K.reshape(ee - eb, -1), Kref.reshape(ee - eb, -1)) <= TOL
            assert relerr(f, fref) <= TOL
    os.environ["IGAB_NO_FUSED_ASSEMBLY"] = "1"
    try:
        Kg, fg = P.assemble(kind, xi1d, w1d, D=D, fconst=0.7)
    finally:
        del os.environ["IGAB_NO_FUSED_ASSEMBLY"]
    K, f = P.assemble(kind, xi1d, w1d, D=D, fconst=0.7)
    assert relerr(K.reshape(len(K), -1), Kg.reshape(len(K), -1)) <= TOL and relerr(f, fg) <= TOL
    assert capi.launch_count() > 0


@pytest.mark.parametrize("kind", [capi.ASM_POISSON, capi.ASM_MASS, capi.ASM_ELASTICITY])
def test_fused_3d_element_matrices_p2(kind):
    """The fused tensor-core kernels at p = 2 (27 functions padded to one 32 x 32 block, 3^3 Gauss) against the oracle and
    the general B-stack path, on a random patch with repeated knots and on the cylinder; non-symmetric D for elasticity."""
    rng = np.random.default_rng(77 + kind)
    pdc = PD.thickCylinder(2, (1, 1, 1))
    specs = [_random_patch(rng, 3, 3, (2, 2, 2)), dict(degree=pdc.degree, knots=pdc.knotSpans, cp=pdc.cp, dimworld=3)]
    x, w = PD.gaussLegendre01(3)
    D = None
    if kind == capi.ASM_ELASTICITY:
        lam, mu = 0.3 / (1.3 * 0.4), 1 / 2.6
        D = np.zeros((6, 6))
        D[:3, :3] = lam
        D[np.arange(3), np.arange(3)] = lam + 2 * mu
        D[np.arange(3, 6), np.arange(3, 6)] = mu
        D = D + 0.05 * rng.normal(size=(6, 6))
    for spec in specs:
        Pp = port_for(spec)
        P = gpu_patch(spec)
        ne = min(Pp.nelem, 7)
        Kref, fref = Pp.assemble(kind, [x] * 3, [w] * 3, D=D, fconst=1.3, elem_begin=1, elem_end=ne)
        K, f = P.assemble(kind, [x] * 3, [w] * 3, D=D, fconst=1.3, elem_begin=1, elem_end=ne)
        assert relerr(K.reshape(ne - 1, -1), Kref.reshape(ne - 1, -1)) <= TOL, relerr(K.reshape(ne - 1, -1), Kref.reshape(ne - 1, -1))
        assert relerr(f, fref) <= TOL
        os.environ["IGAB_NO_FUSED_ASSEMBLY"] = "1"
        try:
            Kg, _ = P.assemble(kind, [x] * 3, [w] * 3, D=D, fconst=1.3, elem_begin=1, elem_end=ne)
        finally:
            del os.environ["IGAB_NO_FUSED_ASSEMBLY"]
        assert relerr(K.reshape(ne - 1, -1), Kg.reshape(ne - 1, -1)) <= TOL


@pytest.mark.parametrize("p,q", [(4, 7), (4, 6), (3, 5), (2, 4)])
def test_reference_default_rules_in_3d_take_the_tuned_kernels(p, q):
    """The reference's default rule has order dim * max(p) (nurbsgridentity.hh:123-124): q = floor(3p / 2) + 1 points per
    direction in 3-D, i.e. 4 / 5 / 7 for p = 2 / 3 / 4.  Those rules must be served by the tuned kernels (KERNEL_TENSOR
    refuses to fall back) and agree with the oracle."""
    pd = PD.thickCylinder(p, (1, 1, 1))
    spec = dict(degree=pd.degree, knots=pd.knotSpans, cp=pd.cp, dimworld=3)
    Pp = port_for(spec)
    P = gpu_patch(spec)
    P.setKernel(capi.KERNEL_TENSOR)
    xi1d = [PD.gaussLegendre01(q)[0]] * 3
    want = ("N", "dN", "x", "Jt", "detJ", "JinvT")
    ref = Pp.eval_tensor(xi1d, order=1, want=want)
    got = P.evalTensor(xi1d, order=1, want=want)
    for f in want:
        assert relerr(got[f], ref[f]) <= TOL, (f, relerr(got[f], ref[f]))


@pytest.mark.parametrize("p", [2, 3, 4])
@pytest.mark.parametrize("kind", [capi.ASM_POISSON, capi.ASM_MASS, capi.ASM_ELASTICITY])
def test_fused_3d_element_matrices_over_integrated(p, kind):
    """Fused 3-D tensor-core assembly with the over-integrated rule (p+2 points per direction) against the oracle."""
    pd = PD.thickCylinder(p, (1, 1, 0))
    spec = dict(degree=pd.degree, knots=pd.knotSpans, cp=pd.cp, dimworld=3)
    Pp = port_for(spec)
    P = gpu_patch(spec)
    x, w = PD.gaussLegendre01(p + 2)
    D = None
    if kind == capi.ASM_ELASTICITY:
        lam, mu = 0.3 / (1.3 * 0.4), 1 / 2.6
        D = np.zeros((6, 6))
        D[:3, :3] = lam
        D[np.arange(3), np.arange(3)] = lam + 2 * mu
        D[np.arange(3, 6), np.arange(3, 6)] = mu
    ne = min(Pp.nelem, 3)
    Kref, fref = Pp.assemble(kind, [x] * 3, [w] * 3, D=D, fconst=0.9, elem_end=ne)
    l0 = capi.launch_count()
    K, f = P.assemble(kind, [x] * 3, [w] * 3, D=D, fconst=0.9, elem_end=ne)
    assert capi.launch_count() - l0 <= 9  # tables + weight check + weighted tables + one fused kernel, not the evaluate / stack / contract sequence
    assert relerr(K.reshape(ne, -1), Kref.reshape(ne, -1)) <= TOL, relerr(K.reshape(ne, -1), Kref.reshape(ne, -1))
    assert relerr(f, fref) <= TOL


@pytest.mark.parametrize("p,extra", [(2, 0), (3, 0), (3, 1), (4, 0), (4, 1), (4, 2)])
@pytest.mark.parametrize("kind", [capi.ASM_POISSON, capi.ASM_MASS, capi.ASM_ELASTICITY])
@pytest.mark.parametrize("sf", ["0", "1"], ids=["dense", "sumfact"])
def test_fused_3d_element_matrices_both_contractions(p, extra, kind, sf, monkeypatch):
    """Both fused 3-D assembly kernels -- the dense Gram contraction and the sum-factorised one (pairs (i_d, j_d)
    contracted direction by direction) -- forced through IGAB_GRAM_SF / IGAB_ELAS_SF at every degree, with a random
    NON-symmetric material for elasticity and a patch whose weights vary in every direction, against the oracle."""
    monkeypatch.setenv("IGAB_GRAM_SF", sf)
    monkeypatch.setenv("IGAB_ELAS_SF", sf)
    rng = np.random.default_rng(40 + p)
    pd = PD.thickCylinder(p, (1, 1, 1))
    cp = pd.cp.copy()
    cp[:, 3] *= rng.uniform(0.7, 1.3, cp.shape[0])          # weights without tensor-product structure
    spec = dict(degree=pd.degree, knots=pd.knotSpans, cp=cp, dimworld=3)
    Pp = port_for(spec)
    P = gpu_patch(spec)
    x, w = PD.gaussLegendre01(p + 1 + extra)
    D = rng.normal(size=(6, 6)) + 4 * np.eye(6) if kind == capi.ASM_ELASTICITY else None
    eb, ee = 3, 8
    Kref, fref = Pp.assemble(kind, [x] * 3, [w] * 3, D=D, fconst=1.3, elem_begin=eb, elem_end=ee)
    l0 = capi.launch_count()
    K, f = P.assemble(kind, [x] * 3, [w] * 3, D=D, fconst=1.3, elem_begin=eb, elem_end=ee)
    assert capi.launch_count() - l0 <= 9   # tables (+ weight check, weighted tables) + ONE fused kernel (p = 4 with 7 points = the reference's default rule)
    assert relerr(K.reshape(ee - eb, -1), Kref.reshape(ee - eb, -1)) <= TOL, relerr(K.reshape(ee - eb, -1), Kref.reshape(ee - eb, -1))
    assert relerr(f, fref) <= TOL


@pytest.mark.parametrize("kind,q", [(capi.ASM_POISSON, 6), (capi.ASM_MASS, 6), (capi.ASM_MASS, 7)])
def test_fused_3d_element_matrices_p5(kind, q):
    """p = 5 (216 functions): the sum-factorised kernel alone (one block per SM, general stage-2 tile loop) against the
    oracle, one fused launch."""
    rng = np.random.default_rng(55)
    pd = PD.thickCylinder(5, (1, 1, 0))
    cp = pd.cp.copy()
    cp[:, 3] *= rng.uniform(0.8, 1.2, cp.shape[0])
    spec = dict(degree=pd.degree, knots=pd.knotSpans, cp=cp, dimworld=3)
    Pp = port_for(spec)
    P = gpu_patch(spec)
    x, w = PD.gaussLegendre01(q)
    Kref, fref = Pp.assemble(kind, [x] * 3, [w] * 3, fconst=0.4, elem_begin=1, elem_end=4)
    l0 = capi.launch_count()
    K, f = P.assemble(kind, [x] * 3, [w] * 3, fconst=0.4, elem_begin=1, elem_end=4)
    assert capi.launch_count() - l0 <= 9
    assert relerr(K.reshape(3, -1), Kref.reshape(3, -1)) <= TOL, relerr(K.reshape(3, -1), Kref.reshape(3, -1))
    assert relerr(f, fref) <= TOL


def test_device_path_is_cuda_graph_capturable():
    """The device-resident entry points enqueue on the caller's stream without host synchronisation (once the rule is
    staged), so a quadrature loop can be captured in a CUDA graph and replayed -- here: evaluation + fused element
    matrices of the C1-type patch, replayed after the control net changed."""
    import torch
    spec = PFT.small_patch("plate_hole_p2", level=3)
    P = gpu_patch(spec)
    x, w = PD.gaussLegendre01(3)
    fields = ("N", "dN", "x", "Jt", "detJ")
    sh = P.shapes(P.nelem * 9)
    out = {f: torch.zeros(sh[f], dtype=torch.float64, device="cuda") for f in fields}
    Ke = torch.zeros((P.nelem, 9, 9), dtype=torch.float64, device="cuda")
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        s = side.cuda_stream
        P.assemble(capi.ASM_POISSON, [x, x], [w, w], Ke=Ke, fe=None, stream=s)  # warm-up: stages the rule (with weights) ...
        P.evalTensor([x, x], order=1, want=fields, out=out, stream=s)           # ... and builds the tables
    side.synchronize()
    ref_x, ref_K = out["x"].clone(), Ke.clone()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        s = torch.cuda.current_stream().cuda_stream
        P.evalTensor([x, x], order=1, want=fields, out=out, stream=s)
        P.assemble(capi.ASM_POISSON, [x, x], [w, w], Ke=Ke, fe=None, stream=s)
    out["x"].zero_()
    Ke.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out["x"], ref_x) and torch.equal(Ke, ref_K)
    # the graph holds pointers into the handle: new control points, same graph
    cp2 = np.array(spec["cp"], float).copy()
    cp2[:, :2] *= 2.0
    P.updateControlPoints(cp2)
    g.replay()
    torch.cuda.synchronize()
    assert torch.allclose(out["x"], 2.0 * ref_x, rtol=1e-14, atol=0) and torch.allclose(Ke, ref_K, rtol=1e-12, atol=1e-15)
    # a rule that is not resident cannot be staged from inside a capture (the copy node would read host memory at replay)
    x4, w4 = PD.gaussLegendre01(4)
    g2 = torch.cuda.CUDAGraph()
    with pytest.raises(capi.IgabError) as ei:
        with torch.cuda.graph(g2, stream=side):
            P.evalTensor([x4, x4], order=1, want=fields, out=None if False else {f: torch.zeros(P.shapes(P.nelem * 16)[f], dtype=torch.float64, device="cuda") for f in fields},
                         stream=torch.cuda.current_stream().cuda_stream)
    assert ei.value.status == capi.INVALID_ARGUMENT


@pytest.mark.parametrize("p", [1, 2, 3])
@pytest.mark.parametrize("dw,kind", [(2, capi.ASM_POISSON), (2, capi.ASM_MASS), (2, capi.ASM_ELASTICITY),
                                     (3, capi.ASM_POISSON), (3, capi.ASM_MASS)])
def test_element_matrices_from_point_lists(p, dw, kind):
    """igab_assemble_points: element matrices of elements with their own quadrature (trimmed-element batches: ragged
    point counts per element from sub-triangle rules, an element without points, more than 32 points) against the
    oracle; and the tensor rule written as a point list reproduces igab_assemble."""
    from dune_iga_b200 import quadrature as QD
    rng = np.random.default_rng(90 + 5 * p + dw + 2 * kind)
    spec = _random_patch(rng, 2, dw, (p, p))
    Pp = port_for(spec)
    P = gpu_patch(spec)
    D = np.array([[1.35, 0.58, 0.02], [0.58, 1.35, -0.03], [0.05, 0.01, 0.385]]) if kind == capi.ASM_ELASTICITY else None
    # ragged batch: element -> 1..8 sub-triangles inside [0,1]^2, 6-point rule on each (6 .. 48 points per element)
    chosen = sorted(rng.choice(Pp.nelem, size=min(5, Pp.nelem), replace=False).tolist())
    batch = {e: np.stack([np.array([[0, 0], [1, 0], [0, 1.0]]) * rng.uniform(0.3, 1.0) for _ in range(int(rng.integers(1, 9)))])
             for e in chosen}
    batch[chosen[0]] = np.stack([np.array([[0, 0], [1, 0], [0, 1.0]]) * s for s in np.linspace(0.3, 1.0, 8)])  # > 32 points
    elem_pt, xi, w, off = QD.trimmedBatch(batch, order=4)
    elems = np.array(chosen, np.int32)
    # add an element without points and one with the plain tensor rule
    x1, w1 = PD.gaussLegendre01(p + 1)
    Xt = np.array([[x1[q % (p + 1)], x1[q // (p + 1)]] for q in range((p + 1) ** 2)])
    Wt = np.array([w1[q % (p + 1)] * w1[q // (p + 1)] for q in range((p + 1) ** 2)])
    elems = np.concatenate([elems, [0, 1]]).astype(np.int32)
    off = np.concatenate([off, [off[-1], off[-1] + len(Wt)]]).astype(np.int64)
    xi = np.concatenate([xi, Xt])
    w = np.concatenate([w, Wt])
    assert np.diff(off).max() > 32 and np.diff(off).min() == 0
    Kref, fref = Pp.assemble_points(kind, elems, off, xi, w, D=D, fconst=0.8)
    K, f = P.assemblePoints(kind, elems, off, xi, w, D=D, fconst=0.8)
    assert relerr(K.reshape(len(K), -1), Kref.reshape(len(K), -1)) <= TOL, relerr(K.reshape(len(K), -1), Kref.reshape(len(K), -1))
    assert relerr(f, fref) <= TOL
    assert np.abs(K[-2]).max() == 0.0 and np.abs(f[-2]).max() == 0.0          # the element without points
    Kt, ft = P.assemble(kind, [x1, x1], [w1, w1], D=D, fconst=0.8, elem_begin=1, elem_end=2)
    assert relerr(K[-1], Kt[0]) <= TOL and relerr(f[-1], ft[0]) <= TOL           # tensor rule as a list
    with pytest.raises(capi.IgabError):
        P.assemblePoints(kind, np.array([Pp.nelem], np.int32), np.array([0, 1], np.int64), xi[:1], w[:1], D=D)


def test_trimmed_patch_poisson_end_to_end():
    """A trimmed patch end to end on the device path: trim flags (empty / trimmed / full elements) -> compacted DOF map
    (prepareForTrim, nurbsbasis.hh:599-615), element matrices of the full elements from the tensor rule and of the trimmed
    elements from their own sub-triangle rules (fillquadraturerule.hh:8-19), COO scatter through the compacted DOF map
    (poisson.py:113-125) -- against the same loop with the oracle.  The area integrated by the load vector equals the
    untrimmed area of the full elements plus the triangulated area of the trimmed ones."""
    import scipy.sparse as sp
    from dune_iga_b200 import quadrature as QD
    rng = np.random.default_rng(123)
    pd = PD.squarePlate(2, 2)                      # 4 x 4 elements, p = 2, unit weights, the unit square
    spec = dict(degree=pd.degree, knots=pd.knotSpans, cp=pd.cp, dimworld=2)
    nel = 16
    flags = np.zeros(nel, np.uint8)
    flags[[0, 1, 4]] = 1                           # empty corner
    trimmed = [2, 5, 8]
    flags[trimmed] = 2
    P = gpu_patch(spec, trimflags=flags)
    Pp = port_for(spec)
    dof, ndof = P.dofmap()
    dref, nref = Pp.dofmap(flags)
    assert ndof == nref and np.array_equal(dof, dref) and ndof < 36
    # each trimmed element keeps the upper-right triangle of its parameter square plus a small one
    tris = {e: np.array([[[1, 0], [1, 1], [0, 1.0]], [[0.6, 0.1], [0.9, 0.1], [0.9, 0.4]]]) for e in trimmed}
    elem_pt, xi, w, off = QD.trimmedBatch(tris, order=4)
    x, wg = PD.gaussLegendre01(3)
    Kfull, ffull = P.assemble(capi.ASM_POISSON, [x, x], [wg, wg], fconst=1.0)
    Ktrim, ftrim = P.assemblePoints(capi.ASM_POISSON, np.array(trimmed, np.int32), off, xi, w, fconst=1.0)
    Kref_full, fref_full = Pp.assemble(capi.ASM_POISSON, [x, x], [wg, wg], fconst=1.0)
    Kref_trim, fref_trim = Pp.assemble_points(capi.ASM_POISSON, np.array(trimmed, np.int32), off, xi, w, fconst=1.0)

    def scatter(Kf, ff, Kt, ft):
        rows, cols, vals = [], [], []
        b = np.zeros(ndof)
        for e in range(nel):
            if flags[e] == 1:
                continue
            Ke, fe = (Kt[trimmed.index(e)], ft[trimmed.index(e)]) if flags[e] == 2 else (Kf[e], ff[e])
            rows.append(np.repeat(dof[e], 9)); cols.append(np.tile(dof[e], 9)); vals.append(Ke.reshape(-1))
            np.add.at(b, dof[e], fe)
        A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(ndof, ndof)).tocsr()
        return A, b

    A, b = scatter(Kfull, ffull, Ktrim, ftrim)
    Ar, br = scatter(Kref_full, fref_full, Kref_trim, fref_trim)
    assert abs(A - Ar).max() <= 1e-12 * abs(Ar).max() and np.abs(b - br).max() <= 1e-12 * np.abs(br).max()
    assert (dof[flags != 1] >= 0).all() and np.abs(A @ np.ones(ndof)).max() <= 1e-12 * abs(A).max()   # constants in the kernel
    # sum_i f_i = integral of 1 over the kept domain = sum of w detJ over the full elements' rule and the trimmed rules
    det = P.evalTensor([x, x], order=1, want=("detJ",))["detJ"].reshape(nel, 9)
    area = (det[flags == 0] * np.outer(wg, wg).reshape(-1)).sum() + \
        (P.evalPoints(elem_pt, xi, order=1, want=("detJ",))["detJ"] * w).sum()
    assert abs(b.sum() - area) <= 1e-13
    # NURBSGeometry::volume summed over the trimmed grid (full elements: tensor rule, trimmed: their sub-triangle rules)
    assert abs(P.volume([x, x], [wg, wg], trimmed=(elem_pt, xi, w)) - area) <= 1e-14
    assert abs(P.volume([x, x], [wg, wg]) - (det[flags == 0] * np.outer(wg, wg).reshape(-1)).sum()) <= 1e-14
    # the same system through the device-side global assembler (compacted CSR pattern, gather, load vector, Dirichlet
    # rows): element matrices of the trimmed elements replace the tensor-rule ones, empty elements are never read
    Kall, fall = Kfull.copy(), ffull.copy()
    Kall[trimmed], fall[trimmed] = Ktrim, ftrim
    Kall[flags == 1], fall[flags == 1] = np.nan, np.nan
    rowptr, colidx = P.csrPattern()
    A.sort_indices()
    assert len(rowptr) == ndof + 1 and rowptr[-1] == A.nnz == len(colidx)
    assert np.array_equal(rowptr, A.indptr) and np.array_equal(colidx, A.indices)        # exactly the scatter's pattern
    vals = P.csrAssemble(Kall)
    Ad = sp.csr_matrix((vals, colidx, rowptr), shape=(ndof, ndof))
    assert abs(Ad - Ar).max() <= 1e-12 * abs(Ar).max()
    bd = P.rhsAssemble(fall)
    assert bd.shape == (ndof,) and np.abs(bd - br).max() <= 1e-12 * np.abs(br).max()
    # ... and the whole thing in ONE call, trimmed rules included
    vg, bg = P.globalAssemble(capi.ASM_POISSON, [x, x], [wg, wg], fconst=1.0,
                              trimmed=(np.array(trimmed, np.int32), off, xi, w))
    assert np.abs(vg - vals).max() <= 1e-14 * np.abs(vals).max() and np.abs(bg - bd).max() <= 1e-14 * np.abs(bd).max()
    with pytest.raises(capi.IgabError):  # the trimmed list must be ascending
        P.globalAssemble(capi.ASM_POISSON, [x, x], [wg, wg], trimmed=(np.array(trimmed[::-1], np.int32), off, xi, w))
    # L2 norm and energy of a field on the trimmed patch (u numbered like the trimmed basis) = the quadratic forms of
    # the trimmed mass / stiffness matrices assembled above by a different kernel chain
    uf = rng.normal(size=ndof)
    vm, _ = P.globalAssemble(capi.ASM_MASS, [x, x], [wg, wg], trimmed=(np.array(trimmed, np.int32), off, xi, w), want_b=False)
    Md = sp.csr_matrix((vm, colidx, rowptr), shape=(ndof, ndof))
    nrm = P.norms([x, x], [wg, wg], uf, trimmed=(elem_pt, xi, w))
    assert abs(nrm[0] - uf @ (Md @ uf)) <= 1e-12 * abs(nrm[0]) and abs(nrm[1] - uf @ (Ad @ uf)) <= 1e-12 * abs(nrm[1])
    # without the lists only the full elements count
    nfull = P.norms([x, x], [wg, wg], uf)
    vmf, _ = P.globalAssemble(capi.ASM_MASS, [x, x], [wg, wg], want_b=False)   # trimmed elements with the tensor rule ...
    assert nfull[0] < nrm[0] and nfull[1] < nrm[1] + 1e-12
    # two element ranges accumulated give the same matrix
    v2 = P.csrAssemble(Kall[:7], elem_end=7)
    v2 = P.csrAssemble(Kall[7:], elem_begin=7, values=v2, accumulate=True)
    assert np.abs(v2 - vals).max() <= 1e-14 * np.abs(vals).max()
    # Dirichlet rows on the compacted numbering
    mask = np.zeros(ndof, np.uint8)
    mask[[0, ndof - 1]] = 1
    gD = np.full(ndof, 2.5)
    vD, bD = P.csrDirichlet(mask, vals.copy(), b=bd.copy(), g=gD)
    AD = sp.csr_matrix((vD, colidx, rowptr), shape=(ndof, ndof)).toarray()
    for r in (0, ndof - 1):
        e = np.zeros(ndof); e[r] = 1.0
        assert np.array_equal(AD[r], e) and bD[r] == 2.5
    keep = np.setdiff1d(np.arange(ndof), [0, ndof - 1])
    assert np.array_equal(AD[keep], Ad.toarray()[keep]) and np.array_equal(bD[keep], bd[keep])


@pytest.mark.parametrize("name,kind", [("plate_hole_p2", capi.ASM_POISSON), ("cylinder_p3", capi.ASM_MASS),
                                       ("plate_hole_p2", capi.ASM_ELASTICITY), ("cylinder_p4", capi.ASM_POISSON)])
def test_global_assemble_one_call_matches_the_stepwise_path(name, kind, monkeypatch):
    """igab_global_assemble (the whole globalAssembler in one call, element matrices kept on the device in chunks)
    against element matrices -> csrAssemble -> rhsAssemble, host and device outputs."""
    spec = PFT.small_patch(name)
    P = gpu_patch(spec)
    dim = len(spec["degree"])
    rule = [PD.gaussLegendre01(p + 1) for p in spec["degree"]]
    xi, w = [r[0] for r in rule], [r[1] for r in rule]
    ncomp = dim if kind == capi.ASM_ELASTICITY else 1
    ns = 3 if dim == 2 else 6
    D = np.eye(ns) + 0.1 if kind == capi.ASM_ELASTICITY else None
    Ke, fe = P.assemble(kind, xi, w, D=D, fconst=0.7)
    vref = P.csrAssemble(Ke, ncomp)
    bref = P.rhsAssemble(fe, ncomp)
    v, b = P.globalAssemble(kind, xi, w, D=D, fconst=0.7)
    assert np.abs(v - vref).max() <= 1e-14 * np.abs(vref).max() and np.abs(b - bref).max() <= 1e-14 * np.abs(bref).max()
    vd, bd = P.globalAssemble(kind, xi, w, D=D, fconst=0.7, device=True)
    assert np.array_equal(vd.cpu().numpy(), v) and np.array_equal(bd.cpu().numpy(), b)
    v2, b2 = P.globalAssemble(kind, xi, w, D=D, fconst=0.7, want_b=False)
    assert b2 is None and np.array_equal(v2, v)
    # many small chunks (the value D2H of finished rows overlaps the assembly of later chunks): same bits
    monkeypatch.setenv("IGAB_GA_CHUNK_ELEMS", "3")
    vc, bc = P.globalAssemble(kind, xi, w, D=D, fconst=0.7)
    vcd, bcd = P.globalAssemble(kind, xi, w, D=D, fconst=0.7, device=True)
    monkeypatch.delenv("IGAB_GA_CHUNK_ELEMS")
    assert np.array_equal(vc, vcd.cpu().numpy()) and np.array_equal(bc, bcd.cpu().numpy())
    assert np.abs(vc - v).max() <= 1e-14 * np.abs(v).max() and np.abs(bc - b).max() <= 1e-14 * np.abs(b).max()
    # the two gather kernels (by rows of the element matrices / per non-zero) add in the same order: identical bits
    monkeypatch.setenv("IGAB_CSR_GATHER", "0")
    assert np.array_equal(P.csrAssemble(Ke, ncomp), vref)
    monkeypatch.delenv("IGAB_CSR_GATHER")
    # element ranges (a rank's slab): the shares add up to the whole system; an empty range gives zeros
    h = P.nelem // 3
    va, ba = P.globalAssemble(kind, xi, w, D=D, fconst=0.7, elem_end=h)
    vb, bb = P.globalAssemble(kind, xi, w, D=D, fconst=0.7, elem_begin=h)
    assert np.abs(va + vb - v).max() <= 1e-14 * np.abs(v).max() and np.abs(ba + bb - b).max() <= 1e-14 * np.abs(b).max()
    vz, bz = P.globalAssemble(kind, xi, w, D=D, fconst=0.7, elem_begin=h, elem_end=h)
    assert not vz.any() and not bz.any()


def test_trimmed_csr_pattern_random_flags_3d_and_vector_valued():
    """Compacted CSR of patches with random empty elements (2-D with two components, 3-D scalar; repeated knots):
    pattern and values equal the COO scatter through the compacted DOF map, bit for bit in the pattern."""
    import scipy.sparse as sp
    for dim, dw, degree, ncomp, seed in ((2, 2, (2, 3), 2, 5), (3, 3, (2, 1, 2), 1, 6)):
        rng = np.random.default_rng(seed)
        spec = _random_patch(rng, dim, dw, degree)
        Pp = port_for(spec)
        nel = Pp.nelem
        flags = rng.choice([0, 1, 2], nel, p=[0.5, 0.35, 0.15]).astype(np.uint8)
        P = gpu_patch(spec, trimflags=flags)
        dof, ndof = P.dofmap()
        n = P.nb * ncomp
        Ke = rng.normal(size=(nel, n, n))
        fe = rng.normal(size=(nel, n))
        rows, cols, vals = [], [], []
        b = np.zeros(ndof * ncomp)
        for e in range(nel):
            if flags[e] == 1:
                continue
            gd = (dof[e][:, None] * ncomp + np.arange(ncomp)[None, :]).reshape(-1)   # flat-interleaved power basis
            rows.append(np.repeat(gd, n)); cols.append(np.tile(gd, n)); vals.append(Ke[e].reshape(-1))
            np.add.at(b, gd, fe[e])
        A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(ndof * ncomp,) * 2).tocsr()
        A.sort_indices()
        Ke[flags == 1] = np.nan
        fe[flags == 1] = np.nan
        rowptr, colidx = P.csrPattern(ncomp)
        assert np.array_equal(rowptr, A.indptr) and np.array_equal(colidx, A.indices)
        v = P.csrAssemble(Ke, ncomp)
        assert np.abs(v - A.data).max() <= 1e-13 * np.abs(A.data).max()
        assert np.abs(P.rhsAssemble(fe, ncomp) - b).max() <= 1e-13 * np.abs(b).max()

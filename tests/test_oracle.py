"""CPU tests (-m "not gpu"): pin the oracles.

 * oracle/_ref/liboracle_ref.so  = the reference's own headers compiled through oracle/ref_shim.cpp
 * oracle/liboracle_port.so      = oracle/igab_oracle.c, the plain-C restatement
are both replayed against every literal the reference's unit tests hold for this path
(tests/golden/reference_literals.json, see make_reference_literals.py) and against each other on seeded patches.
Tolerance for literals: the reference's own, Dune::FloatCmp::eq default = 8 eps relative, widened to 64 eps because the
literals are printed with 16 digits and the reference's last bits are compiler-dependent.
"""
import json
import os

import numpy as np
import pytest

import portbind as PT
import refbind as RF
import patches_for_tests as PFT

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_literals.json")))
EPS = 64 * np.finfo(float).eps

needs_ref = pytest.mark.skipif(not RF.available(), reason="oracle/_ref not built")
IMPLS = [pytest.param("port", id="port"), pytest.param("ref", id="ref", marks=needs_ref)]


def mod(name):
    return PT if name == "port" else RF


def make(name, degree, knots, cp, dw):
    return PT.PortPatch(degree, knots, cp, dw) if name == "port" else RF.RefPatch.create(degree, knots, cp, dw)


def close(a, b, tol=EPS):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)).max()))


@pytest.mark.parametrize("impl", IMPLS)
def test_bspline1d_literals(impl):
    M = mod(impl)
    for c in G["bspline1d_values"]:
        assert close(M.bspline_basis(c["u"], c["knots"], c["p"]), c["N"]), c["cite"]
    c = G["bspline1d_ders"]
    assert close(M.bspline_basis(c["u"], c["knots"], c["p"]), c["N"])
    assert close(M.bspline_ders(c["u"], c["knots"], c["p"], 3), c["dN"])
    c = G["bspline1d_N0_sweep"]
    got = [M.bspline_basis(u, c["knots"], c["p"])[0] for u in c["u"]]
    assert close(got, c["N0"])


@pytest.mark.parametrize("impl", IMPLS)
def test_find_span_semantics(impl):
    # bsplinealgorithms.hh:24-32: clamps at both ends, upper_bound semantics with repeated interior knots
    M = mod(impl)
    U = [0, 0, 0, 0.5, 0.5, 2, 2, 3, 3, 3]
    assert M.find_span(2, -1.0, U) == 2
    assert M.find_span(2, 0.0, U) == 2
    assert M.find_span(2, 0.3, U) == 2
    assert M.find_span(2, 0.5, U) == 4
    assert M.find_span(2, 1.0, U) == 4
    assert M.find_span(2, 2.0, U) == 6
    assert M.find_span(2, 3.0, U) == len(U) - 2 - 2
    assert M.find_span(2, 9.0, U) == len(U) - 2 - 2
    # element path: value = unique knot e, offset = e
    uq = [0, 0.5, 2, 3]
    assert [M.find_span(2, uq[e], U, e) for e in range(3)] == [2, 4, 6]


def nurbs2d_patch(impl):
    n = G["nurbs2d"]
    w = np.ones((3, 7))  # [j][i], proper 7x3 net; only the 3x3 block of the first span is read
    for i in range(3):
        for j in range(3):
            w[j, i] = n["weights_3x3_block"][i][j]
    cp = np.zeros((21, 3))
    cp[:, 2] = w.reshape(-1)
    return make(impl, n["degree"], n["knots"], cp, 2)


@pytest.mark.parametrize("impl", IMPLS)
def test_nurbs2d_literals(impl):
    P = nurbs2d_patch(impl)
    n = G["nurbs2d"]
    for v in n["values"]:
        R = P.nurbs_basis(v["u"])
        assert close(R, v["R"])
        assert abs(R.sum() - 1.0) < EPS  # partition of unity, gridtests.cpp:787,804
    d = P.nurbs_ders([0.0, 0.1], 5)
    for key, val in n["derivatives_at_0_0.1"].items():
        a, b = map(int, key.split(","))
        assert close(d[a + 6 * b], val), key


@pytest.mark.parametrize("impl", IMPLS)
def test_geometry_literals(impl):
    c = G["geometry_curve"]
    P = make(impl, c["degree"], c["knots"], c["cp"], c["dimworld"])
    for u, x in c["global"]:
        assert close(P.geo_global(u), x)
    for u, J in c["jacobianTransposed"]:
        assert close(P.geo_jt(u), J)
    s = G["geometry_surface"]
    cp = np.array([[s["cp_table_ij"][i][j] for i in range(3)] for j in range(3)], float).reshape(9, 4)
    P = make(impl, s["degree"], s["knots"], cp, s["dimworld"])
    for u, x in s["global"]:
        assert close(P.geo_global(u), x)
    for u, J in s["jacobianTransposed"]:
        assert close(P.geo_jt(u), J)
    for k, corner in enumerate(s["corners"]):
        assert close(P.geo_global([(k >> 0) & 1, (k >> 1) & 1]), corner)


@needs_ref
def test_degree_elevate_literal():
    c = G["degree_elevate_curve"]
    P = RF.RefPatch.create(c["degree"], c["knots"], c["cp"], c["dimworld"]).degree_elevate(0, c["elevate_by"])
    assert P.degree == [4]
    assert np.allclose(P.cp[:, :3], c["expected_cp"], rtol=0, atol=1e-10 * 6)
    assert np.allclose(P.cp[:, 3], c["expected_w"], rtol=0, atol=1e-10 * 12)


@needs_ref
@pytest.mark.parametrize("name", ["plate_hole_p2", "cylinder_p3", "torus_p2", "square_p3", "mixed_deg_3d"])
def test_port_equals_reference_on_seeded_patches(name):
    """The restatement against the reference itself, pointwise, all outputs, order 2."""
    spec = PFT.small_patch(name)
    Pp = PT.PortPatch(spec["degree"], spec["knots"], spec["cp"], spec["dimworld"])
    Pr = RF.RefPatch.create(spec["degree"], spec["knots"], spec["cp"], spec["dimworld"])
    assert Pp.nelem == Pr.nelem
    assert np.array_equal(Pp.spans(), Pr.spans())  # integers: bit-exact
    xi1d = [PT.gauss01(p + 1)[0] for p in spec["degree"]]
    a = Pp.eval_tensor(xi1d, order=2)
    b = Pr.eval_tensor(xi1d, order=2)
    for k in ("N", "dN", "d2N", "x", "Jt", "H", "detJ"):
        scale = max(1.0, np.abs(b[k]).max())
        assert np.abs(a[k] - b[k]).max() <= 1e-13 * scale, k
    # corners and corner +- 1e-8, as checkJacobianAndPartialConsistency does (gridtests.cpp:514-576)
    rng = np.random.default_rng(42)
    elem = rng.integers(0, Pp.nelem, 64).astype(np.int32)
    xi = rng.random((64, Pp.dim))
    xi[:8] = 0.0
    xi[8:16] = 1.0
    xi[16:24] = 1e-8
    xi[24:32] = 1 - 1e-8
    a = Pp.eval_points(elem, xi, order=2)
    b = Pr.eval_points(elem, xi, order=2)
    for k in ("N", "dN", "d2N", "x", "Jt", "H", "detJ"):
        scale = max(1.0, np.abs(b[k]).max())
        assert np.abs(a[k] - b[k]).max() <= 1e-13 * scale, k
    # evaluateJacobian column d == partial(e_d)  (gridtests.cpp:514-576, 1e-14)
    for d in range(Pp.dim):
        al = [0] * Pp.dim
        al[d] = 1
        for Q, o in ((Pp, a), (Pr, b)):
            part = Q.partial(elem, xi, al)
            assert np.abs(part - o["dN"][:, :, d]).max() <= 1e-14 * max(1.0, np.abs(part).max())


@needs_ref
def test_torus_invariants_reference():
    """gridtests.cpp:338-363: torus area 4 pi^2 r R (1e-4) and Gauss-Bonnet (1e-5) through the oracle."""
    t = G["torus"]
    spec = PFT.small_patch("torus_p2", level=2)
    Pp = PT.PortPatch(spec["degree"], spec["knots"], spec["cp"], 3)
    x, w = PT.gauss01(3)  # order 2*max(p)=4 -> 3 points (nurbsgridentity.hh:123-133)
    area = Pp.volume([x, x], [w, w])
    assert abs(area - 4 * np.pi ** 2 * t["r"] * t["R"]) < t["area_tol"]
    o = Pp.eval_tensor([x, x], order=2)
    J, H, det = o["Jt"], o["H"], o["detJ"]
    nrm = np.cross(J[:, 0], J[:, 1])
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    b00, b11, b01 = (H[:, 0] * nrm).sum(1), (H[:, 1] * nrm).sum(1), (H[:, 2] * nrm).sum(1)
    g00, g11, g01 = (J[:, 0] ** 2).sum(1), (J[:, 1] ** 2).sum(1), (J[:, 0] * J[:, 1]).sum(1)
    K = (b00 * b11 - b01 ** 2) / (g00 * g11 - g01 ** 2)
    ww = np.tile(np.outer(w, w).reshape(-1), Pp.nelem)
    assert abs((K * det * ww).sum()) < t["gauss_bonnet_tol"]
    r, R = t["r"], t["R"]
    assert K.min() >= -1 / (r * (R - r)) - 1e-8 and K.max() <= 1 / (r * (R + r)) + 1e-8


def test_dofmap_port():
    """nurbsbasis.hh:552-587: simple interior knots -> g_d = e_d + l_d; repeated knots shift; trim compaction."""
    kn = [np.array([0, 0, 0, .25, .5, .75, 1, 1, 1.]), np.array([0, 0, 0, .5, .5, 1, 1, 1.])]
    ncp = [6, 5]
    cp = np.ones((30, 3))
    P = PT.PortPatch([2, 2], kn, cp, 2)
    dof, n = P.dofmap()
    assert n == 30 and P.nelem == 8
    assert list(dof[0]) == [0, 1, 2, 6, 7, 8, 12, 13, 14]
    # element (1,1): span0 = 3, span1 = 4 (double knot) -> start (1,2)
    assert list(dof[1 + 4 * 1]) == [13, 14, 15, 19, 20, 21, 25, 26, 27]
    flags = np.zeros(8, np.uint8)
    flags[[0, 1]] = 1  # two empty elements
    dof2, n2 = P.dofmap(flags)
    used = sorted(set(dof[2:].reshape(-1)))
    assert n2 == len(used)
    remap = {g: k for k, g in enumerate(used)}
    assert all(dof2[e, i] == remap[dof[e, i]] for e in range(2, 8) for i in range(9))


def test_assembly_port_properties():
    """K_e symmetric, rows sum to zero (constants in the kernel of the Laplacian), sum of mass = volume."""
    spec = PFT.small_patch("plate_hole_p2")
    P = PT.PortPatch(spec["degree"], spec["knots"], spec["cp"], 2)
    x, w = PT.gauss01(3)
    Ke, fe = P.assemble(0, [x, x], [w, w])
    assert np.abs(Ke - Ke.transpose(0, 2, 1)).max() < 1e-13
    assert np.abs(Ke.sum(2)).max() < 1e-12
    Me, _ = P.assemble(1, [x, x], [w, w])
    assert abs(Me.sum() - P.volume([x, x], [w, w])) < 1e-12
    assert abs(fe.sum() - P.volume([x, x], [w, w])) < 1e-12


@pytest.mark.parametrize("impl", IMPLS)
def test_geometry_local_literals(impl):
    """trimmedgridtests.cpp:90-95,157-162: geometry.local(x) of the curve and the surface.  The port restates the
    element-level NURBSGeometry::local, which coincides with the patch-level one on these one-element patches."""
    c = G["geometry_curve"]
    s = G["geometry_surface"]
    cp = np.array([[s["cp_table_ij"][i][j] for i in range(3)] for j in range(3)], float).reshape(9, 4)
    for spec, cpn in ((c, c["cp"]), (s, cp)):
        P = make(impl, spec["degree"], spec["knots"], cpn, spec["dimworld"])
        for x, u in spec["local"]:
            got = P.local([0], [x])[0][0] if impl == "port" else P.geo_local(x)
            assert close(got, u), (x, u, got)


@needs_ref
def test_geometry_local_port_matches_reference():
    """NURBSPatchGeometry::local compiled from the reference vs the port, seeded targets on one-element patches:
    Newton branch (2-D in 2-D, 3-D in 3-D) and trust-region projection (curve in 2-D, surface in 3-D)."""
    rng = np.random.default_rng(11)
    c = G["geometry_curve"]
    s = G["geometry_surface"]
    cps = np.array([[s["cp_table_ij"][i][j] for i in range(3)] for j in range(3)], float).reshape(9, 4)
    k2 = [0, 0, 0, 1, 1, 1.0]
    quad = np.array([[x + 0.2 * y * y, y + 0.1 * x * y, 1 + 0.3 * x * (1 - y)] for y in (0, .5, 1) for x in (0, .6, 1)])
    hexa = np.array([[x + 0.1 * z, y + 0.2 * x * z, z + 0.1 * y * y, 1 + 0.2 * x * y * z]
                     for z in (0, .5, 1) for y in (0, .5, 1) for x in (0, .4, 1)])
    cases = [(c["degree"], c["knots"], np.array(c["cp"], float), 2, (-5, 5)),
             (s["degree"], s["knots"], cps, 3, (-0.5, 2.5)),
             ([2, 2], [k2, k2], quad, 2, (-0.2, 1.3)),
             ([2, 2, 2], [k2, k2, k2], hexa, 3, (-0.1, 1.2))]
    for degree, knots, cp, dw, (lo, hi) in cases:
        Pp = PT.PortPatch(degree, knots, cp, dw)
        Pr = RF.RefPatch.create(degree, knots, cp, dw)
        X = rng.uniform(lo, hi, (120, dw))
        if len(degree) == dw:
            # Newton branch: targets inside the image only -- the reference's patch-level loop has no iteration cap and
            # does not terminate for points it cannot reach (nurbspatchgeometry.hh:136-155)
            X = np.array([Pr.geo_global(u) for u in rng.uniform(0, 1, (120, dw))])
        xi, flag = Pp.local(np.zeros(len(X), np.int32), X)
        ref = [Pr.geo_local(x) for x in X]
        threw = np.array([r is None for r in ref])
        # failures (MathError in the reference) coincide up to comparisons that sit on a rounding boundary
        assert (threw != (flag != 0)).sum() <= 2
        ok = ~threw & (flag == 0)
        assert ok.sum() >= 40
        d = np.abs(xi[ok] - np.array([r for r, o in zip(ref, ok) if o]))
        assert d.max() <= 1e-9, d.max()
        assert np.median(d) <= 1e-14


def test_geometry_local_round_trip_port():
    """gridtests.cpp:1085-1100: local(global(corner)) == corner to 1e-14, here for corners and interior points of every
    element of a perturbed square and of the cylinder solid (the plate with a hole has a singular corner, where Newton
    converges only linearly)."""
    rng = np.random.default_rng(5)
    for name in ("square_p3", "cylinder_p3"):
        spec = PFT.small_patch(name)
        P = PT.PortPatch(spec["degree"], spec["knots"], spec["cp"], spec["dimworld"])
        dim = P.dim
        corners = np.array([[(k >> d) & 1 for d in range(dim)] for k in range(2 ** dim)], float)
        elem = np.repeat(np.arange(P.nelem, dtype=np.int32), len(corners) + 3)
        xi = np.concatenate([np.concatenate([corners, rng.uniform(0.05, 0.95, (3, dim))]) for _ in range(P.nelem)])
        X = P.eval_points(elem, xi, order=0, want=("x",))["x"]
        got, flag = P.local(elem, X)
        assert (flag == 0).all()
        assert np.abs(got - xi).max() < 1e-12


def test_surface_forms_port_on_torus_and_sphere_like_patch():
    """nurbsgeometry.hh:216-255 through the port: Gauss-Bonnet on the torus (gridtests.cpp:348-363), unit normals
    orthogonal to both tangents, metric = J J^T, and K = 1/R^2 on a surface of revolution that is a sphere."""
    t = G["torus"]
    spec = PFT.small_patch("torus_p2", level=2)
    Pp = PT.PortPatch(spec["degree"], spec["knots"], spec["cp"], 3)
    x, w = PT.gauss01(3)
    o = Pp.eval_tensor([x, x], order=2)
    f = PT.surface_forms(o["Jt"], o["H"])
    ww = np.tile(np.outer(w, w).reshape(-1), Pp.nelem)
    assert abs((f["gaussK"] * o["detJ"] * ww).sum()) < t["gauss_bonnet_tol"]
    assert np.abs((f["unitNormal"][:, None, :] * o["Jt"]).sum(2)).max() < 1e-13
    assert np.abs(np.linalg.norm(f["unitNormal"], axis=1) - 1).max() < 1e-14
    assert np.abs(f["metric"] - np.einsum("pac,pbc->pab", o["Jt"], o["Jt"])).max() < 1e-13
    assert np.abs(np.sqrt(np.linalg.det(f["metric"])) - o["detJ"]).max() < 1e-12
    assert np.abs(f["secondFF"][:, 0, 1] - f["secondFF"][:, 1, 0]).max() == 0.0
    r, R = t["r"], t["R"]
    assert f["gaussK"].min() >= -1 / (r * (R - r)) - 1e-8 and f["gaussK"].max() <= 1 / (r * (R + r)) + 1e-8


def test_field_norms_port_properties():
    """igo_norms: u = 1 -> (volume, 0); u = x-coordinate of the control net on a polynomial patch -> (int x^2, volume);
    and the quadratic forms u^T M u / u^T K u of the element matrices of the same rule."""
    spec = PFT.small_patch("square_p3")  # unit weights: the field with coefficients P_i reproduces x
    Pp = PT.PortPatch(spec["degree"], spec["knots"], spec["cp"], spec["dimworld"])
    rule = [PT.gauss01(p + 1) for p in spec["degree"]]
    xi1d, w1d = [r[0] for r in rule], [r[1] for r in rule]
    vol = Pp.volume(xi1d, w1d)
    ndof = Pp.cp.shape[0]
    n1 = Pp.norms(xi1d, w1d, np.ones(ndof))
    assert abs(n1[0] - vol) < 1e-13 * vol and abs(n1[1]) < 1e-20
    nx = Pp.norms(xi1d, w1d, Pp.cp[:, 0].copy())
    o = Pp.eval_tensor(xi1d, order=1, want=("x", "detJ"))
    ww = np.tile(np.outer(w1d[1], w1d[0]).reshape(-1), Pp.nelem)
    assert abs(nx[0] - (o["x"][:, 0] ** 2 * o["detJ"] * ww).sum()) < 1e-13
    assert abs(nx[1] - vol) < 1e-12 * vol
    rng = np.random.default_rng(2)
    u = rng.normal(size=ndof)
    dof, _ = Pp.dofmap()
    Ke, _ = Pp.assemble(0, xi1d, w1d, want_fe=False)
    Me, _ = Pp.assemble(1, xi1d, w1d, want_fe=False)
    ue = u[dof]
    nu = Pp.norms(xi1d, w1d, u)
    assert abs(nu[0] - np.einsum("ei,eij,ej->", ue, Me, ue)) < 1e-12 * nu[0]
    assert abs(nu[1] - np.einsum("ei,eij,ej->", ue, Ke, ue)) < 1e-12 * nu[1]
    # two components, sub-range: additivity
    u2 = rng.normal(size=(ndof, 2))
    h = Pp.nelem // 2
    full = Pp.norms(xi1d, w1d, u2, ncomp=2)
    parts = Pp.norms(xi1d, w1d, u2, ncomp=2, elem_end=h) + Pp.norms(xi1d, w1d, u2, ncomp=2, elem_begin=h)
    assert np.abs(full - parts).max() < 1e-12 * full.max()


@pytest.mark.parametrize("impl", IMPLS)
def test_element_centres_after_refinement(impl):
    """trimmedgridtests.cpp:197-217: centres of the four elements of the once-refined unit square, compared with ==."""
    from dune_iga_b200 import patchdata as PD
    c = G["element_centres"]
    pd = PD.NurbsPatchData(c["knots"], c["cp"], c["degree"], c["dimworld"]).globalRefine(c["refine_level"])
    P = make(impl, pd.degree, pd.knotSpans, pd.cp, 2)
    spans = [(0.0, 0.5), (0.5, 1.0)]
    for e, centre in enumerate(c["centres"]):
        u = [0.5 * sum(spans[e % 2]), 0.5 * sum(spans[e // 2])]
        assert np.array_equal(P.geo_global(u), np.array(centre)), (e, P.geo_global(u))


def test_assemble_points_port_equals_tensor_assembly():
    """igo_assemble_points with the tensor rule written out as a point list reproduces igo_assemble (same per-point
    contribution), element by element, for Poisson / mass / plane elasticity."""
    spec = PFT.small_patch("plate_hole_p2")
    Pp = PT.PortPatch(spec["degree"], spec["knots"], spec["cp"], 2)
    x, w = PT.gauss01(3)
    X = np.array([[x[q % 3], x[q // 3]] for q in range(9)])
    W = np.array([w[q % 3] * w[q // 3] for q in range(9)])
    elem = np.array([0, 3, Pp.nelem - 1], np.int32)
    off = np.array([0, 9, 18, 27], np.int64)
    D = np.array([[1.35, 0.58, 0], [0.58, 1.35, 0], [0, 0, 0.385]])
    for kind in (0, 1, 2):
        Kp, fp = Pp.assemble_points(kind, elem, off, np.tile(X, (3, 1)), np.tile(W, 3), D=D if kind == 2 else None, fconst=1.2)
        for k, e in enumerate(elem):
            Kt, ft = Pp.assemble(kind, [x, x], [w, w], D=D if kind == 2 else None, fconst=1.2, elem_begin=int(e), elem_end=int(e) + 1)
            assert np.abs(Kp[k] - Kt[0]).max() <= 1e-14 * np.abs(Kt).max()
            assert np.abs(fp[k] - ft[0]).max() <= 1e-14 * np.abs(ft).max()


def test_real_index_maps_port_against_the_reference_tests():
    """getDirectIndex<0> / getRealIndex<0> (nurbspatch.hh:599-631).  The flag patterns are the ones the reference's own
    test produces (trimmedgridtests.cpp:335-373): one trimmed element; after one refinement 1 full + 3 trimmed (1:1
    maps); element_trim_xb after one refinement: element 0 empty, 1..3 trimmed -> real(i) = i - 1, direct(i - 1) = i."""
    EMPTY, TRIMMED, FULL = 1, 2, 0
    dor, rod = PT.real_index_maps(1, [TRIMMED])
    assert dor.tolist() == [0] and rod.tolist() == [0]
    dor, rod = PT.real_index_maps(4, [FULL, TRIMMED, TRIMMED, TRIMMED])
    assert dor.tolist() == [0, 1, 2, 3] and rod.tolist() == [0, 1, 2, 3]
    dor, rod = PT.real_index_maps(4, [EMPTY, TRIMMED, TRIMMED, TRIMMED])
    assert [rod[i] for i in range(1, 4)] == [0, 1, 2] and [dor[i - 1] for i in range(1, 4)] == [1, 2, 3]
    assert rod[0] == -1  # the reference throws "No corresponding realIndex" (:616-622)
    dor, rod = PT.real_index_maps(7)  # no trimming: identity (:599-606)
    assert dor.tolist() == list(range(7)) == rod.tolist()
    rng = np.random.default_rng(0)
    fl = rng.choice([0, 1, 2], size=1000).astype(np.uint8)
    dor, rod = PT.real_index_maps(1000, fl)
    keep = np.flatnonzero(fl != EMPTY)
    assert np.array_equal(dor, keep) and np.array_equal(rod[keep], np.arange(len(keep))) and (rod[fl == EMPTY] == -1).all()


def test_load_vector_port_properties():
    """Per-point source (poisson.py:79): a constant source reproduces the load vector of igo_assemble; the tensor-rule
    and the point-list variants agree when the list holds the tensor points; the vector-valued numbering is
    flat-interleaved."""
    from dune_iga_b200 import patchdata as PD
    import patches_for_tests as PFT
    for name in ("plate_hole_p2", "cylinder_p3"):
        spec = PFT.small_patch(name)
        O = PT.PortPatch(spec["degree"], spec["knots"], spec["cp"], spec["dimworld"])
        dim = len(spec["degree"])
        x, w = PD.gaussLegendre01(max(spec["degree"]) + 1)
        nq = len(x) ** dim
        _, fe0 = O.assemble(0, [x] * dim, [w] * dim, fconst=2.5)
        fe1 = O.load_vector([x] * dim, [w] * dim, np.full((O.nelem * nq, 1), 2.5))
        assert np.abs(fe0 - fe1).max() <= 1e-15 * np.abs(fe0).max()
        rng = np.random.default_rng(5)
        f = rng.normal(size=(O.nelem * nq, 3))
        fe3 = O.load_vector([x] * dim, [w] * dim, f, ncomp=3)
        for c in range(3):
            assert np.array_equal(fe3[:, c::3], O.load_vector([x] * dim, [w] * dim, f[:, c:c + 1].copy()))
        # the same points as explicit lists
        grids = np.meshgrid(*([x] * dim), indexing="ij")
        xi = np.stack([g.T.reshape(-1) for g in grids], axis=1) if dim > 1 else x[:, None]
        ww = np.ones(nq)
        for d in range(dim):
            ww = ww * np.tile(np.repeat(w, len(x) ** d), len(x) ** (dim - 1 - d))
        xi = np.stack([np.tile(np.repeat(x, len(x) ** d), len(x) ** (dim - 1 - d)) for d in range(dim)], axis=1)
        elem = np.arange(O.nelem, dtype=np.int32)
        off = np.arange(O.nelem + 1) * nq
        fel = O.load_vector_points(elem, off, np.tile(xi, (O.nelem, 1)), np.tile(ww, O.nelem), f, ncomp=3)
        assert np.abs(fel - fe3).max() <= 1e-14 * np.abs(fe3).max()

"""GPU tests at BASELINE.json's full sizes (-m gpu): the oracle cannot reach these sizes in seconds, so the whole
output is checked on the device through size-independent properties (partition of unity, gradient sums, exact radius of
the cylinder map, detJ == |det J|, volume in closed form, scaling / linearity of the geometry, run-to-run bitwise
reproducibility), and a random sample of elements is compared with the oracle."""
import numpy as np
import pytest

import igab200  # noqa: F401
import portbind as PT
from dune_iga_b200 import capi, patchdata as PD, quadrature as Q
from parity_util import oracle_for,  TOL, relerr

pytestmark = pytest.mark.gpu


def dev_out(P, npts, fields):
    import torch
    sh = P.shapes(npts)
    return {f: torch.empty(sh[f], dtype=torch.float64, device="cuda") for f in fields}


def test_c2_full_patch_properties():
    import torch
    pd = PD.thickCylinder(3, (7, 7, 7))  # 128^3 elements, 131^3 control points
    P = capi.Patch.fromPatchData(pd)
    P.setKernel(capi.KERNEL_TENSOR)
    assert P.nelem == 128 ** 3
    x1, w1 = Q.gaussLegendre01(4)
    xi, w = [x1] * 3, [w1] * 3
    fields = ("N", "dN", "x", "Jt", "detJ")
    layers = 4
    nel = 128 * 128 * layers
    npts = nel * 64
    out = dev_out(P, npts, fields)
    s = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(0)
    # the reference's own headers (oracle/_ref) when they are there, else the pinned port
    Pp, _ = oracle_for(dict(degree=pd.degree, knots=pd.knotSpans, cp=pd.cp, dimworld=3))
    for k0 in (0, 62, 124):  # first, middle, last slab of elements
        eb = 128 * 128 * k0
        P.evalTensor(xi, order=1, elem_begin=eb, elem_end=eb + nel, want=fields, out=out, stream=s)
        torch.cuda.synchronize()
        assert float((out["N"].sum(1) - 1).abs().max()) < 2e-14
        assert float(out["dN"].sum(1).abs().max()) < 1e-11
        assert float(out["detJ"].min()) > 0
        r = torch.hypot(out["x"][:, 0], out["x"][:, 1])
        assert float(r.min()) >= 1 - 1e-12 and float(r.max()) <= 2 + 1e-12
        # radius is affine in the radial parameter and z in the axial one: check against the rule's coordinates
        det = torch.linalg.det(out["Jt"])
        assert float((det.abs() - out["detJ"]).abs().max() / out["detJ"].max()) < 1e-13
        # bitwise reproducible
        ref = {f: out[f].clone() for f in fields}
        P.evalTensor(xi, order=1, elem_begin=eb, elem_end=eb + nel, want=fields, out=out, stream=s)
        torch.cuda.synchronize()
        assert all(torch.equal(ref[f], out[f]) for f in fields)
        # oracle on a random sample of elements of this slab
        for e in rng.integers(eb, eb + nel, 6):
            o = Pp.eval_tensor(xi, order=1, elem_begin=int(e), elem_end=int(e) + 1, want=fields)
            sl = slice((int(e) - eb) * 64, (int(e) - eb + 1) * 64)
            for f in fields:
                assert relerr(out[f][sl].cpu().numpy(), o[f]) <= TOL, (f, int(e))
    # closed-form volume of the quarter cylinder shell: pi/4 (r_o^2 - r_i^2) L
    vol = P.volume(xi, w)
    assert abs(vol - np.pi / 4 * 3.0 * 3.0) < 1e-11
    # scaling the control net by 2 doubles x and J, multiplies detJ by 8, leaves N untouched
    P.evalTensor(xi, order=1, elem_begin=0, elem_end=nel, want=fields, out=out, stream=s)
    torch.cuda.synchronize()
    base = {f: out[f].clone() for f in fields}
    cp2 = pd.cp.copy()
    cp2[:, :3] *= 2.0
    P.updateControlPoints(cp2)
    P.evalTensor(xi, order=1, elem_begin=0, elem_end=nel, want=fields, out=out, stream=s)
    torch.cuda.synchronize()
    assert torch.equal(out["N"], base["N"]) and torch.equal(out["dN"], base["dN"])
    assert torch.equal(out["x"], 2 * base["x"]) and torch.equal(out["Jt"], 2 * base["Jt"])
    assert float((out["detJ"] - 8 * base["detJ"]).abs().max() / base["detJ"].max()) < 1e-14


def test_c4_full_torus_shell():
    """1024 x 1024 elements, 4 x 4 Gauss, second derivatives: torus area and Gauss-Bonnet on the device
    (gridtests.cpp:338-363 at config-C4 size)."""
    import torch
    pd = PD.torusSurface(8)
    P = capi.Patch.fromPatchData(pd)
    assert P.nelem == 1024 * 1024
    x1, w1 = Q.gaussLegendre01(4)
    fields = ("N", "d2N", "Jt", "H", "detJ")
    out = dev_out(P, P.nelem * 16, fields)
    P.evalTensor([x1, x1], order=2, want=fields, out=out, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ww = torch.from_numpy(np.outer(w1, w1).reshape(-1)).cuda().repeat(P.nelem)
    area = float((out["detJ"] * ww).sum())
    assert abs(area - 4 * np.pi ** 2 * 1.0 * 2.0) < 1e-9
    J, H = out["Jt"], out["H"]
    nrm = torch.linalg.cross(J[:, 0], J[:, 1])
    nrm = nrm / nrm.norm(dim=1, keepdim=True)
    b00, b11, b01 = (H[:, 0] * nrm).sum(1), (H[:, 1] * nrm).sum(1), (H[:, 2] * nrm).sum(1)
    g00, g11, g01 = (J[:, 0] ** 2).sum(1), (J[:, 1] ** 2).sum(1), (J[:, 0] * J[:, 1]).sum(1)
    K = (b00 * b11 - b01 ** 2) / (g00 * g11 - g01 ** 2)
    assert abs(float((K * out["detJ"] * ww).sum())) < 1e-8
    assert float(K.min()) >= -1.0 - 1e-8 and float(K.max()) <= 1 / 3.0 + 1e-8
    assert float((out["N"].sum(1) - 1).abs().max()) < 2e-14
    assert float(out["d2N"].sum(1).abs().max()) < 1e-6 * float(out["d2N"].abs().max())


def test_c5_trimmed_batches():
    """Config C5: tensor rule on full elements + per-element point lists on trimmed elements (fillquadraturerule
    format).  Square plate [0,1]^2, p=3, hole of radius 0.25 at the centre triangulated per element."""
    # unit weights: the map is the identity on [0,1]^2, so trimmed areas are known in closed form
    pd = PD.NurbsPatchData([[0, 0, 1, 1], [0, 0, 1, 1]], [(0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1)], [1, 1], 2)
    pd = pd.degreeElevate(0, 2).degreeElevate(1, 2).globalRefine(5)  # p = 3, 32 x 32 elements
    P = capi.Patch.fromPatchData(pd)
    Pp = PT.PortPatch(pd.degree, pd.knotSpans, pd.cp, 2)
    n = 32
    rng = np.random.default_rng(42)
    # synthetic trimmed elements: random sub-triangulations covering a known fraction of each chosen element
    chosen = np.sort(rng.choice(n * n, 60, replace=False))
    tris, frac = {}, {}
    for e in chosen:
        c = rng.uniform(0.3, 0.7, 2)
        # fan of 4 triangles around an interior point covers the whole element; drop one => known area
        corners = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], float)
        fan = [np.array([c, corners[k], corners[(k + 1) % 4]]) for k in range(4)]
        keep = fan[:3]
        tris[int(e)] = np.array(keep)
        frac[int(e)] = sum(0.5 * abs((t[1] - t[0])[0] * (t[2] - t[0])[1] - (t[1] - t[0])[1] * (t[2] - t[0])[0]) for t in keep)
    elem, xi, w, off = Q.trimmedBatch(tris, order=4)
    got = P.evalPoints(elem, xi, order=1, want=("N", "dN", "x", "Jt", "detJ"))
    Pr, _ = oracle_for(dict(degree=pd.degree, knots=pd.knotSpans, cp=pd.cp, dimworld=2))
    ref = Pr.eval_points(elem, xi, order=1, want=("N", "dN", "x", "Jt", "detJ"))
    for f in ("N", "dN", "x", "Jt", "detJ"):
        assert relerr(got[f], ref[f]) <= TOL, f
    # the plate is the identity map scaled by 1/32 per element: area of the trimmed part = frac / n^2
    for k, e in enumerate(sorted(tris)):
        a = (got["detJ"][off[k]:off[k + 1]] * w[off[k]:off[k + 1]]).sum()
        assert abs(a - frac[e] / n ** 2) < 1e-13 * frac[e] / n ** 2
    # trim flags: empty elements drop out of the DOF map, trimmed ones stay
    flags = np.zeros(n * n, np.uint8)
    flags[chosen] = 2
    flags[:n] = 1
    Pt = capi.Patch(pd.degree, pd.knotSpans, pd.cp, 2, trimflags=flags)
    dof, nd = Pt.dofmap()
    dref, nref = Pp.dofmap(flags)
    assert nd == nref and np.array_equal(dof, dref)


def test_general_assembly_path_beyond_65535_elements():
    """Regression: the general B-stack contraction indexes elements with gridDim.z; 2-D patch with 512 x 256 = 131,072
    elements and a rule that is not (p+1)^2 (so the fused 2-D kernel does not apply).  Properties: rows of the Poisson
    matrices sum to zero (constants are in the kernel), sum of all mass-matrix entries = area; oracle on sampled elements."""
    import portbind as PT
    pd = PD.plateWithHole(8, 8)
    P = capi.Patch.fromPatchData(pd)
    nel = P.nelem
    assert nel > 65535
    x, w = PD.gaussLegendre01(4)
    Ke, _ = P.assemble(capi.ASM_POISSON, [x, x], [w, w], want_fe=False)
    assert Ke.shape == (nel, 9, 9)
    assert np.abs(Ke.sum(2)).max() <= 1e-11 * np.abs(Ke).max()
    Me, _ = P.assemble(capi.ASM_MASS, [x, x], [w, w], want_fe=False)
    assert abs(Me.sum() - P.volume([x, x], [w, w])) <= 1e-11 * Me.sum()
    Pp = PT.PortPatch(pd.degree, pd.knotSpans, pd.cp, pd.dimworld)
    for e in (0, 65535, 65536, nel - 1):
        Kref, _ = Pp.assemble(capi.ASM_POISSON, [x, x], [w, w], elem_begin=e, elem_end=e + 1, want_fe=False)
        assert np.abs(Ke[e] - Kref[0]).max() <= 1e-12 * np.abs(Kref).max()


def test_c3_full_size_element_matrices(monkeypatch):
    """Config C3 at full size (p = 4, 64^3 elements, 5^3 Gauss): element matrices of whole k-slabs checked on the device
    through properties that do not need the oracle -- symmetry, constants in the kernel of the stiffness matrix (row sums
    0), positive diagonal, sum of all mass-matrix entries == patch volume in closed form, rigid translations in the
    kernel of the elasticity matrix, bitwise reproducibility -- then the sum-factorised contraction against the dense
    Gram kernel on a whole chunk, and a random sample of elements against the oracle."""
    import torch
    p, lev = 4, 6
    pd = PD.thickCylinder(p, (lev, lev, lev))
    P = capi.Patch.fromPatchData(pd)
    assert P.nelem == 64 ** 3 and P.nb == 125
    x1, w1 = Q.gaussLegendre01(5)
    xi, w = [x1] * 3, [w1] * 3
    nel = 64 * 64 * 2  # two k-layers per call: 1.02 GB of element matrices
    K = torch.empty((nel, 125, 125), dtype=torch.float64, device="cuda")
    F = torch.empty((nel, 125), dtype=torch.float64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    Pp = PT.PortPatch(pd.degree, pd.knotSpans, pd.cp, 3)
    rng = np.random.default_rng(3)
    vol = 0.0
    for k0 in range(0, 64, 2):
        eb = 64 * 64 * k0
        P.assemble(capi.ASM_MASS, xi, w, elem_begin=eb, elem_end=eb + nel, Ke=K, fe=None, stream=s)
        vol += float(K.sum())
        if k0 in (0, 30, 62):
            assert float((K - K.transpose(1, 2)).abs().max()) <= 1e-13 * float(K.abs().max())
            assert float(K.min()) > 0                     # products of positive rational functions
    assert abs(vol - np.pi / 4 * (2.0 ** 2 - 1.0 ** 2) * 3.0) <= 1e-11 * vol   # sum_ij M_ij = int 1 = volume of the solid
    for k0 in (0, 30, 62):
        eb = 64 * 64 * k0
        P.assemble(capi.ASM_POISSON, xi, w, fconst=1.0, elem_begin=eb, elem_end=eb + nel, Ke=K, fe=F, stream=s)
        torch.cuda.synchronize()
        amax = float(K.abs().max())
        assert float((K - K.transpose(1, 2)).abs().max()) <= 1e-12 * amax
        assert float(K.sum(2).abs().max()) <= 1e-11 * amax          # grad of the constant function is 0
        assert float(torch.diagonal(K, dim1=1, dim2=2).min()) > 0
        ref = K.clone()
        P.assemble(capi.ASM_POISSON, xi, w, fconst=1.0, elem_begin=eb, elem_end=eb + nel, Ke=K, fe=F, stream=s)
        torch.cuda.synchronize()
        assert torch.equal(ref, K)                                   # bitwise reproducible
        # the dense-Gram kernel on the same chunk (a different algorithm, same numbers)
        monkeypatch.setenv("IGAB_GRAM_SF", "0")
        P.assemble(capi.ASM_POISSON, xi, w, fconst=1.0, elem_begin=eb, elem_end=eb + nel, Ke=K, fe=None, stream=s)
        monkeypatch.delenv("IGAB_GRAM_SF")
        torch.cuda.synchronize()
        assert float((K - ref).abs().max()) <= 1e-12 * amax
        for e in rng.integers(eb, eb + nel, 3):
            Ko, fo = Pp.assemble(capi.ASM_POISSON, xi, w, fconst=1.0, elem_begin=int(e), elem_end=int(e) + 1)
            assert relerr(ref[int(e) - eb].cpu().numpy(), Ko[0]) <= TOL
            assert relerr(F[int(e) - eb].cpu().numpy(), fo[0]) <= TOL
    # elasticity (n = 375): a chunk of one quarter k-layer, symmetric material
    lam, mu = 0.3 / (1.3 * 0.4), 1 / 2.6
    D = np.zeros((6, 6))
    D[:3, :3] = lam
    D[np.arange(3), np.arange(3)] = lam + 2 * mu
    D[np.arange(3, 6), np.arange(3, 6)] = mu
    ne = 1024
    KE = torch.empty((ne, 375, 375), dtype=torch.float64, device="cuda")
    eb = 64 * 64 * 31 + 517
    P.assemble(capi.ASM_ELASTICITY, xi, w, D=D, elem_begin=eb, elem_end=eb + ne, Ke=KE, fe=None, stream=s)
    torch.cuda.synchronize()
    amax = float(KE.abs().max())
    assert float((KE - KE.transpose(1, 2)).abs().max()) <= 1e-12 * amax
    for c in range(3):                                               # rigid translations produce no strain
        t = torch.zeros(375, dtype=torch.float64, device="cuda")
        t[c::3] = 1.0
        assert float((KE @ t).abs().max()) <= 1e-11 * amax
    ref = KE.clone()
    monkeypatch.setenv("IGAB_ELAS_SF", "0")
    P.assemble(capi.ASM_ELASTICITY, xi, w, D=D, elem_begin=eb, elem_end=eb + ne, Ke=KE, fe=None, stream=s)
    monkeypatch.delenv("IGAB_ELAS_SF")
    torch.cuda.synchronize()
    assert float((KE - ref).abs().max()) <= 1e-12 * amax


@pytest.mark.parametrize("case", [(2, 2, (4, 5), (6, 5), (7, 6)), (3, 3, (3, 2, 4), (4, 3, 5), (5, 4, 3)),
                                  (2, 3, (5, 4), (7, 7), (7, 7))],
                         ids=lambda c: "d%dw%d_p%s" % (c[0], c[1], "".join(map(str, c[2]))))
def test_warp_per_element_kernel_many_elements(case):
    """The general warp-per-element kernel with far more elements than resident warps (grid-stride loop, chunk tails,
    mixed degrees, second derivatives): every field against the thread-per-point kernel on the device, a random sample
    of elements against the oracle."""
    import torch
    dim, dw, degree, nq1, lev = case
    if dim == 2 and dw == 2:
        pd = PD.plateWithHole(0, 0)
        for d, p in enumerate(degree):
            pd = pd.degreeElevate(d, p - 2)
    elif dim == 2:
        pd = PD.torusSurface(0)
        for d, p in enumerate(degree):
            pd = pd.degreeElevate(d, p - 2)
    else:
        pd = PD.thickCylinder(2)
        for d, p in enumerate(degree):
            pd = pd.degreeElevate(d, p - 2)
    for d in range(dim):
        pd = pd.globalRefineInDirection(d, lev[d])
    P = capi.Patch.fromPatchData(pd)
    assert P.nelem > 148 * 5 * 4
    xi = [Q.gaussLegendre01(q)[0] for q in nq1]
    fields = ("N", "dN", "d2N", "x", "Jt", "H", "detJ")
    npts = P.nelem * int(np.prod(nq1))
    s = torch.cuda.current_stream().cuda_stream
    outs = {}
    for mode in (capi.KERNEL_WARP, capi.KERNEL_GENERIC):
        P.setKernel(mode)
        outs[mode] = dev_out(P, npts, fields)
        P.evalTensor(xi, order=2, want=fields, out=outs[mode], stream=s)
    torch.cuda.synchronize()
    for f in fields:
        a, b = outs[capi.KERNEL_WARP][f], outs[capi.KERNEL_GENERIC][f]
        assert float((a - b).abs().max()) <= 1e-12 * max(1.0, float(b.abs().max())), f
    assert float((outs[capi.KERNEL_WARP]["N"].sum(1) - 1).abs().max()) < 1e-13
    Pp, _ = oracle_for(dict(degree=pd.degree, knots=pd.knotSpans, cp=pd.cp, dimworld=dw))
    nq = int(np.prod(nq1))
    for e in np.random.default_rng(5).integers(0, P.nelem, 5):
        o = Pp.eval_tensor(xi, order=2, elem_begin=int(e), elem_end=int(e) + 1, want=fields)
        for f in fields:
            # second derivatives of a degree-5 basis on 2^-7 spans are second differences over tiny knot intervals: two
            # correct FP64 algorithms (the oracle's sums of R'' P and the kernel's homogeneous quotient rule) agree to
            # ~ eps (p / h)^2 h^2 / |H|, a few 1e-12 here -- 1e-11 for the order-2 fields, 1e-12 for everything else
            tol = 1e-11 if f in ("d2N", "H") else TOL
            assert relerr(outs[capi.KERNEL_WARP][f][int(e) * nq:(int(e) + 1) * nq].cpu().numpy(), o[f]) <= tol, (f, int(e))

"""GPU parity tests (-m gpu) of the entry points added in round 2: real <-> direct element index maps, the multi-block
ordered scan behind them and behind the trim compaction of the DOF map, load vectors with a per-point source, and the
handle's stream discipline.  Everything goes through the C-ABI and is compared with the oracle."""
import ctypes as C

import numpy as np
import pytest

import igab200  # noqa: F401
import patches_for_tests as PFT
import portbind as PT
from dune_iga_b200 import capi
from dune_iga_b200 import patchdata as PD
from dune_iga_b200 import quadrature as QD
from parity_util import TOL, port_for, relerr

pytestmark = pytest.mark.gpu


def gpu_patch(spec, **kw):
    return capi.Patch(spec["degree"], spec["knots"], spec["cp"], spec["dimworld"], **kw)


def test_real_index_maps_reference_golden_and_random_flags():
    """igab_patch_real_index_maps == getDirectIndex<0> / getRealIndex<0> (nurbspatch.hh:599-631): the patterns of the
    reference's own test (trimmedgridtests.cpp:335-373), random flags over more elements than one scan tile (4,096), and
    the identity on an untrimmed patch -- int32, bit-exact against the oracle."""
    pd = PD.squarePlate(2, 1)  # 2 x 2 elements
    spec = dict(degree=pd.degree, knots=pd.knotSpans, cp=pd.cp, dimworld=2)
    EMPTY, TRIMMED, FULL = 1, 2, 0
    for flags, real, direct in (([FULL, TRIMMED, TRIMMED, TRIMMED], [0, 1, 2, 3], [0, 1, 2, 3]),
                                ([EMPTY, TRIMMED, TRIMMED, TRIMMED], [-1, 0, 1, 2], [1, 2, 3])):
        P = gpu_patch(spec, trimflags=np.array(flags, np.uint8))
        dor, rod = P.realIndexMaps()
        assert rod.tolist() == real and dor.tolist() == direct
    P = gpu_patch(spec)
    dor, rod = P.realIndexMaps()
    assert dor.tolist() == [0, 1, 2, 3] == rod.tolist()
    rng = np.random.default_rng(17)
    pd = PD.squarePlate(2, 7)  # 128 x 128 = 16,384 elements: several scan tiles
    spec = dict(degree=pd.degree, knots=pd.knotSpans, cp=pd.cp, dimworld=2)
    for probs in ([0.5, 0.3, 0.2], [0.0, 1.0, 0.0], [0.02, 0.97, 0.01]):
        flags = rng.choice([0, 1, 2], size=128 * 128, p=probs).astype(np.uint8)
        P = gpu_patch(spec, trimflags=flags)
        dor, rod = P.realIndexMaps()
        odor, orod = PT.real_index_maps(len(flags), flags)
        assert dor.dtype == np.int32 and np.array_equal(dor, odor) and np.array_equal(rod, orod)
        # the DOF compaction runs through the same scan (18,496 DOFs: several tiles)
        dof, n = P.dofmap()
        dref, nref = port_for(spec).dofmap(flags)
        assert n == nref and np.array_equal(dof, dref)


@pytest.mark.parametrize("name,ncomp", [("plate_hole_p2", 1), ("cylinder_p3", 3), ("torus_p2", 2), ("curve_p3", 1),
                                        ("mixed_deg_3d", 1)])
def test_load_vector_with_per_point_source_matches_oracle(name, ncomp):
    """fe[i] = sum_q w detJ R_i f(x_q) (poisson.py:79) with f given per point: tensor rule, sub-ranges, device arrays."""
    import torch
    spec = PFT.small_patch(name)
    P, O = gpu_patch(spec), port_for(spec)
    dim = len(spec["degree"])
    rule = [PD.gaussLegendre01(p + 1) for p in spec["degree"]]
    xi, w = [r[0] for r in rule], [r[1] for r in rule]
    nq = int(np.prod([len(a) for a in xi]))
    x = P.evalTensor(xi, order=0, want=("x",))["x"]
    # a smooth source evaluated by the caller at the quadrature points, as f(posGlobal) in the reference's loop
    f = np.stack([np.sin(1.0 + c + x @ np.arange(1, x.shape[1] + 1)) for c in range(ncomp)], axis=1)
    fe = P.loadVector(xi, w, f, ncomp)
    ref = O.load_vector(xi, w, f, ncomp)
    assert relerr(fe, ref) <= TOL
    eb, ee = P.nelem // 3, P.nelem - 1
    sub = P.loadVector(xi, w, f[eb * nq:ee * nq], ncomp, elem_begin=eb, elem_end=ee)
    assert np.array_equal(sub, fe[eb:ee])
    fd = P.loadVector(xi, w, torch.from_numpy(f).cuda(), ncomp)
    assert np.array_equal(fd.cpu().numpy(), fe)
    # a constant source reproduces igab_assemble's load vector (Poisson kind; scalar)
    if ncomp == 1 and dim == spec["dimworld"]:
        _, f0 = P.assemble(capi.ASM_POISSON, xi, w, fconst=1.75)
        f1 = P.loadVector(xi, w, np.full((P.nelem * nq, 1), 1.75))
        assert relerr(f1, f0) <= TOL
    # global load vector = scatter of the element vectors through the DOF map (poisson.py:121)
    dof, ndof = P.dofmap()
    b = P.globalLoadVector(xi, w, f, ncomp)
    bref = np.zeros(ndof * ncomp)
    for c in range(ncomp):
        np.add.at(bref, dof.reshape(-1) * ncomp + c, ref[:, c::ncomp].reshape(-1))
    assert np.abs(b - bref).max() <= TOL * np.abs(bref).max()
    assert np.array_equal(P.globalLoadVector(xi, w, torch.from_numpy(f).cuda(), ncomp).cpu().numpy(), b)


def test_load_vector_on_a_trimmed_patch_with_own_rules():
    """Trimmed elements integrate the source with their own sub-triangle rules (fillquadraturerule.hh:8-19), empty
    elements contribute nothing, numbering compacted (prepareForTrim)."""
    pd = PD.squarePlate(2, 2)
    spec = dict(degree=pd.degree, knots=pd.knotSpans, cp=pd.cp, dimworld=2)
    flags = np.zeros(16, np.uint8)
    flags[[0, 1, 4]] = 1
    trimmed = [2, 5, 8]
    flags[trimmed] = 2
    P, O = gpu_patch(spec, trimflags=flags), port_for(spec)
    tris = {e: np.array([[[1, 0], [1, 1], [0, 1.0]], [[0.6, 0.1], [0.9, 0.1], [0.9, 0.4]]]) for e in trimmed}
    for qt in ("GaussLegendre", "GaussLobatto"):
        elem_pt, xl, wl, off = QD.trimmedBatch(tris, order=3 if qt == "GaussLobatto" else 4, qt=qt)
        x, wg = PD.gaussLegendre01(3)
        xq = P.evalTensor([x, x], order=0, want=("x",))["x"]
        xlq = P.evalPoints(elem_pt, xl, order=0, want=("x",))["x"]
        src = lambda X: (1.0 + X[:, 0] + 2.0 * X[:, 1] ** 2)[:, None]  # noqa: E731
        f, fl = src(xq), src(xlq)
        fel = P.loadVectorPoints(np.array(trimmed, np.int32), off, xl, wl, fl)
        assert relerr(fel, O.load_vector_points(np.array(trimmed, np.int32), off, xl, wl, fl)) <= TOL
        b = P.globalLoadVector([x, x], [wg, wg], f, trimmed=(np.array(trimmed, np.int32), off, xl, wl), f_list=fl)
        dof, ndof = P.dofmap()
        fe = O.load_vector([x, x], [wg, wg], f)
        fe[trimmed] = O.load_vector_points(np.array(trimmed, np.int32), off, xl, wl, fl)
        bref = np.zeros(ndof)
        for e in range(16):
            if flags[e] != 1:
                np.add.at(bref, dof[e], fe[e])
        assert b.shape == (ndof,) and np.abs(b - bref).max() <= TOL * np.abs(bref).max()
        # integral of the source over the kept domain = sum of the load vector (partition of unity)
        det = P.evalTensor([x, x], order=1, want=("detJ",))["detJ"].reshape(16, 9)
        detl = P.evalPoints(elem_pt, xl, order=1, want=("detJ",))["detJ"]
        tot = (det * np.outer(wg, wg).reshape(-1) * f.reshape(16, 9))[flags == 0].sum() + (detl * wl * fl[:, 0]).sum()
        assert abs(b.sum() - tot) <= 1e-13 * abs(tot)


def test_elasticity_device_path_does_not_synchronise():
    """igab_assemble(ELASTICITY, IGAB_DEVICE) keeps D in the handle: no allocation, no stream synchronisation -- the call
    can be captured in a CUDA graph and replayed (streams and graphs)."""
    import torch
    spec = PFT.small_patch("cylinder_p3")
    P, O = gpu_patch(spec), port_for(spec)
    x, w = PD.gaussLegendre01(4)
    E, nu = 1.0, 0.3
    lam, mu = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
    D = np.zeros((6, 6))
    D[:3, :3] = lam
    D[np.arange(3), np.arange(3)] += 2 * mu
    D[np.arange(3, 6), np.arange(3, 6)] = mu
    n = P.nb * 3
    Ke = torch.empty((P.nelem, n, n), dtype=torch.float64, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        P.assemble(capi.ASM_ELASTICITY, [x] * 3, [w] * 3, D=D, Ke=Ke, fe=None, stream=s.cuda_stream)  # stages rule + D
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        Ke.zero_()
        with torch.cuda.graph(g, stream=s):
            P.assemble(capi.ASM_ELASTICITY, [x] * 3, [w] * 3, D=D, Ke=Ke, fe=None, stream=s.cuda_stream)
        g.replay()
        s.synchronize()
    ref, _ = O.assemble(capi.ASM_ELASTICITY, [x] * 3, [w] * 3, D=D, want_fe=False)
    assert relerr(Ke.cpu().numpy().reshape(P.nelem, -1), ref.reshape(P.nelem, -1)) <= TOL
    # a different material is picked up (restaged) by the next plain call
    Ke2 = torch.empty_like(Ke)
    P.assemble(capi.ASM_ELASTICITY, [x] * 3, [w] * 3, D=2.0 * D, Ke=Ke2, fe=None)
    torch.cuda.synchronize()
    assert relerr(Ke2.cpu().numpy().reshape(P.nelem, -1), 2.0 * ref.reshape(P.nelem, -1)) <= TOL


def test_assemble_host_outputs_in_chunks_match_device_outputs(monkeypatch):
    """igab_assemble(IGAB_HOST) streams chunks through two device buffers with the D2H copy of chunk k overlapping the
    kernel of chunk k+1: identical bits to the device-resident result, also with a ragged last chunk."""
    import torch
    spec = PFT.small_patch("cylinder_p3", level=2)
    P = gpu_patch(spec)
    x, w = PD.gaussLegendre01(4)
    Kd = torch.empty((P.nelem, 64, 64), dtype=torch.float64, device="cuda")
    fd = torch.empty((P.nelem, 64), dtype=torch.float64, device="cuda")
    P.assemble(capi.ASM_POISSON, [x] * 3, [w] * 3, Ke=Kd, fe=fd)
    torch.cuda.synchronize()
    monkeypatch.setenv("IGAB_ASM_CHUNK_ELEMS", "7")
    Kh, fh = P.assemble(capi.ASM_POISSON, [x] * 3, [w] * 3)
    assert np.array_equal(Kh, Kd.cpu().numpy()) and np.array_equal(fh, fd.cpu().numpy())
    Kp = capi.pinned_empty((P.nelem, 64, 64))
    P.assemble(capi.ASM_POISSON, [x] * 3, [w] * 3, Ke=Kp, fe=None)
    assert np.array_equal(Kp, Kh)
    capi.pinned_free(Kp)


def test_switching_streams_on_one_handle_is_ordered():
    """ADVICE (round 1): handle-owned scratch (staged rule, univariate tables, counters) is re-staged when the caller
    switches streams; the library orders the new stream behind the work still in flight on the old one, so alternating
    rules and streams without host synchronisation gives the same results as a serial run."""
    import torch
    spec = PFT.small_patch("cylinder_p3", level=3)
    P = gpu_patch(spec)
    rules = [PD.gaussLegendre01(4)[0], PD.gaussLegendre01(5)[0]]
    serial = [P.evalTensor([r] * 3, order=1, want=("N", "dN", "detJ")) for r in rules]
    s = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = []
    for it in range(6):
        k = it & 1
        npts = P.nelem * len(rules[k]) ** 3
        sh = P.shapes(npts)
        o = {f: torch.empty(sh[f], dtype=torch.float64, device="cuda") for f in ("N", "dN", "detJ")}
        P.evalTensor([rules[k]] * 3, order=1, want=("N", "dN", "detJ"), out=o, stream=s[k].cuda_stream)
        outs.append((k, o))
    torch.cuda.synchronize()
    for k, o in outs:
        for f in ("N", "dN", "detJ"):
            assert np.array_equal(o[f].cpu().numpy(), serial[k][f]), (k, f)


@pytest.mark.parametrize("p", [3, 4])
@pytest.mark.parametrize("kind", [capi.ASM_POISSON, capi.ASM_MASS, capi.ASM_ELASTICITY])
def test_warp_specialised_assembly_kernel_matches_oracle_and_the_barrier_separated_kernel(p, kind, monkeypatch):
    """gram_ws_kernel (DMMA warps + helper warps, named barriers, tile buffer) against the oracle and, bit for bit, against
    gram_sf_kernel: more elements than resident blocks (grid-stride + pipeline drain between elements), non-symmetric
    material, load vector."""
    import torch
    pd = PD.thickCylinder(p, (3, 3, 2))  # 8 x 8 x 4 = 256 elements > 148 resident blocks
    spec = dict(degree=pd.degree, knots=pd.knotSpans, cp=PD.perturbed(pd, seed=3, amp=0.05).cp, dimworld=3)
    P, O = gpu_patch(spec), port_for(spec)
    x, w = PD.gaussLegendre01(p + 1)
    rng = np.random.default_rng(9)
    D = rng.normal(size=(6, 6)) + 4 * np.eye(6) if kind == capi.ASM_ELASTICITY else None
    n = P.nb * (3 if kind == capi.ASM_ELASTICITY else 1)
    out = {}
    monkeypatch.setenv("IGAB_ELAS_SF", "1")
    monkeypatch.setenv("IGAB_GRAM_SYM", "0")   # bit for bit against the kernel that computes every block
    for mode in ("0", "1"):
        monkeypatch.setenv("IGAB_GRAM_WS", mode)
        Ke = torch.zeros((P.nelem, n, n), dtype=torch.float64, device="cuda")
        fe = torch.zeros((P.nelem, n), dtype=torch.float64, device="cuda")
        l0 = capi.launch_count()
        P.assemble(kind, [x] * 3, [w] * 3, D=D, fconst=0.7, Ke=Ke, fe=fe)
        torch.cuda.synchronize()
        assert capi.launch_count() > l0
        out[mode] = (Ke.cpu().numpy(), fe.cpu().numpy())
    assert np.array_equal(out["0"][0], out["1"][0]) and np.array_equal(out["0"][1], out["1"][1])
    sample = np.array([0, 1, 147, 148, 149, 200, 255])
    for e in sample:
        Kr, fr = O.assemble(kind, [x] * 3, [w] * 3, D=D, fconst=0.7, elem_begin=int(e), elem_end=int(e) + 1)
        assert relerr(out["1"][0][e].reshape(1, -1), Kr.reshape(1, -1)) <= TOL
        assert relerr(out["1"][1][e].reshape(1, -1), fr.reshape(1, -1)) <= TOL


def test_device_resident_degree_elevation_chain_matches_host_and_reference():
    """igab_refine_operators: degree elevation of several directions (and a mixed elevation + knot-insertion chain) applied
    in HBM, against the host operators (patchdata.degreeElevate), the reference's own degreeElevate
    (nurbsalgorithms.hh:264-439 through oracle/_ref) and -- the curve must not change -- the geometry map itself."""
    import torch
    import refbind as RF
    cases = [(PFT.small_patch("torus_p2", level=0)["patch"], (1, 2)), (PD.perturbed(PD.thickCylinder(2, (1, 0, 1)), seed=5), (1, 0, 2)),
             (PFT.small_patch("mixed_deg_3d", level=0)["patch"], (2, 1, 1)), (PFT.small_patch("curve_p3", level=0)["patch"], (1,))]
    for pd, fac in cases:
        host = pd
        for d, t in enumerate(fac):
            host = host.degreeElevate(d, t)
        deg, knots, got = capi.degreeElevateAll(pd, fac)
        assert list(deg) == list(host.degree) and all(np.array_equal(a, b) for a, b in zip(knots, host.knotSpans))
        scale = np.abs(host.cp).max()
        assert got.shape == host.cp.shape and np.abs(got - host.cp).max() <= 1e-12 * scale
        if RF.available():
            R = RF.RefPatch.create(pd.degree, pd.knotSpans, pd.cp, pd.dimworld)
            for d, t in enumerate(fac):
                if t:
                    R = R.degree_elevate(d, t)
            assert np.abs(got - np.asarray(R.cp).reshape(got.shape)).max() <= 1e-11 * scale
        # device in / device out, handle from the device net; the map x(xi) is the one of the original patch
        gdev = capi.degreeElevateAll(pd, fac, cp_dev=torch.from_numpy(np.ascontiguousarray(pd.cp)).cuda(), device_out=True)[2]
        assert np.array_equal(gdev.cpu().numpy(), got)
        Pe = capi.Patch.fromDeviceNet(deg, knots, gdev, pd.dimworld)
        P0 = capi.Patch.fromPatchData(pd)
        xi = [PD.gaussLegendre01(3)[0]] * pd.patchDim
        a, b = Pe.evalTensor(xi, order=1, want=("x", "Jt")), P0.evalTensor(xi, order=1, want=("x", "Jt"))
        assert np.abs(a["x"] - b["x"]).max() <= 1e-12 * np.abs(b["x"]).max()
        assert np.abs(a["Jt"] - b["Jt"]).max() <= 1e-10 * np.abs(b["Jt"]).max()
    # a mixed chain: elevate direction 0, then insert knots in direction 1 of the elevated net
    pd = cases[0][0]
    U0, E0 = PD.degreeElevationOperator(pd.knotSpans[0], pd.degree[0], 1)
    nk = PD.generateRefinedKnots(pd.knotSpans, 1, 2)
    U1, T1 = PD.knotInsertionOperator(pd.knotSpans[1], pd.degree[1], nk)
    got = capi.refineOperators(pd, [(0, E0), (1, T1)])
    host = pd.degreeElevate(0, 1).knotRefinement(nk, 1)
    assert np.abs(got - host.cp).max() <= 1e-12 * np.abs(host.cp).max()
    # no operator: a copy; an operator that does not fit: refused
    assert np.array_equal(capi.refineOperators(pd, []), pd.cp)
    first, ln, coef = capi.banded(E0)
    first = first.copy(); first[-1] += 1
    ncp = np.asarray(pd.ncp, np.int32)
    out = np.empty((E0.shape[0] * pd.ncp[1], pd.dimworld + 1))
    one = lambda a: np.asarray([a], np.int32)
    st = capi.lib().igab_refine_operators(pd.patchDim, pd.dimworld, ncp.ctypes.data_as(capi.ip), 1, one(0).ctypes.data_as(capi.ip),
                                          one(E0.shape[0]).ctypes.data_as(capi.ip), one(coef.shape[1]).ctypes.data_as(capi.ip),
                                          (C.c_void_p * 1)(first.ctypes.data), (C.c_void_p * 1)(ln.ctypes.data),
                                          (capi.dp * 1)(coef.ctypes.data_as(capi.dp)), np.ascontiguousarray(pd.cp).ctypes.data,
                                          capi.HOST, out.ctypes.data, capi.HOST, None)
    assert st == capi.INVALID_ARGUMENT


def test_device_resident_refinement_chain_matches_host_and_reference():
    """igab_refine_knots: operators built on the device (Oslo), all directions applied in HBM, the handle created from the
    device net (igab_patch_create_from_device) -- against the host operators (patchdata.knotRefinement) and the
    reference's own knotRefinement (nurbsalgorithms.hh:441-528 through oracle/_ref) on 2-D / 3-D patches with repeated
    interior knots, non-uniform weights and uneven refinement per direction; then the whole chain at 128^3 elements."""
    import torch
    import refbind as RF
    cases = [PFT.small_patch("torus_p2", level=0)["patch"], PFT.small_patch("mixed_deg_3d", level=0)["patch"],
             PD.perturbed(PD.thickCylinder(3, (1, 0, 2)), seed=2), PFT.small_patch("curve_p3", level=0)["patch"]]
    for pd in cases:
        levels = [2, 1, 3][:pd.patchDim]
        fine = pd
        for d, lev in enumerate(levels):
            fine = fine.globalRefineInDirection(d, lev)
        got = capi.refineKnots(pd, fine.knotSpans)
        assert got.shape == fine.cp.shape
        scale = np.abs(fine.cp).max()
        assert np.abs(got - fine.cp).max() <= 1e-13 * scale
        if RF.available():
            R = RF.RefPatch.create(pd.degree, pd.knotSpans, pd.cp, pd.dimworld)
            for d, lev in enumerate(levels):
                R = R.refine_uniform(d, lev)
            assert np.abs(got - np.asarray(R.cp).reshape(got.shape)).max() <= 1e-13 * scale
        # device in / device out, then a handle straight from the device net: evaluation equals the host-built handle
        gdev = capi.refineKnots(pd, fine.knotSpans, cp_dev=torch.from_numpy(np.ascontiguousarray(pd.cp)).cuda(), device_out=True)
        assert np.array_equal(gdev.cpu().numpy(), got)
        Pd_ = capi.Patch.fromDeviceNet(fine.degree, fine.knotSpans, gdev, fine.dimworld)
        Ph = capi.Patch(fine.degree, fine.knotSpans, got, fine.dimworld)
        xi = [PD.gaussLegendre01(p + 1)[0] for p in fine.degree]
        a, b = Pd_.evalTensor(xi, order=1), Ph.evalTensor(xi, order=1)
        assert all(np.array_equal(a[f], b[f]) for f in a)
    # errors: not a refinement / unsorted
    pd = cases[0]
    bad = [k.copy() for k in pd.knotSpans]
    bad[0] = np.concatenate([bad[0][:3], [0.3, 0.2], bad[0][3:]])
    with pytest.raises(capi.IgabError):
        capi.refineKnots(pd, bad)
    # the C2 chain: coarse solid (elevated to p=3) -> 128^3 elements, entirely on the device; the map is unchanged
    coarse = PD.thickCylinder(3, (0, 0, 0))
    knots = []
    for d in range(3):
        knots.append(np.sort(np.concatenate([coarse.knotSpans[d], PD.generateRefinedKnots(coarse.knotSpans, d, 7)])))
    net = capi.refineKnots(coarse, knots, device_out=True)
    assert net.shape == (131 ** 3, 4)
    P = capi.Patch.fromDeviceNet(coarse.degree, knots, net, 3)
    assert P.nelem == 128 ** 3
    x, w = PD.gaussLegendre01(4)
    vol = P.volume([x] * 3, [w] * 3)
    assert abs(vol - np.pi / 4 * 3.0 * 3.0) <= 1e-12 * vol
    if RF.available():  # the reference's own chain to 32^3 on the CPU, the device chain to the same level: same net
        R = RF.RefPatch.create(coarse.degree, coarse.knotSpans, coarse.cp, 3)
        k5 = []
        for d in range(3):
            R = R.refine_uniform(d, 5)
            k5.append(np.sort(np.concatenate([coarse.knotSpans[d], PD.generateRefinedKnots(coarse.knotSpans, d, 5)])))
        n5 = capi.refineKnots(coarse, k5)
        assert np.abs(n5 - np.asarray(R.cp).reshape(n5.shape)).max() <= 1e-13 * np.abs(n5).max()


@pytest.mark.parametrize("p", [3, 4])
@pytest.mark.parametrize("kind", [capi.ASM_POISSON, capi.ASM_MASS, capi.ASM_ELASTICITY])
def test_tensor_product_weights_folded_into_the_tables(p, kind, monkeypatch):
    """Patches whose weights are a tensor product (here: the quarter cylinder, weights (1, 1/sqrt 2, 1) x 1 x 1 and its
    refinements) take the assembly path with the weights inside the univariate tables (no w_i w_j scaling per stored
    entry); IGAB_NO_WTAB=1 at handle creation keeps the general path.  Same element matrices to rounding, both kernels
    (barrier-separated / warp-specialised), and the oracle on a sample; a handle whose weights stop being a tensor
    product after igab_patch_update_control_points falls back by itself."""
    import torch
    pd = PD.thickCylinder(p, (2, 2, 2))
    spec = dict(degree=pd.degree, knots=pd.knotSpans, cp=pd.cp, dimworld=3)
    x, w = PD.gaussLegendre01(p + 1)
    D = np.eye(6) * 2.0 + 0.3 if kind == capi.ASM_ELASTICITY else None
    monkeypatch.setenv("IGAB_ELAS_SF", "1")
    monkeypatch.setenv("IGAB_GRAM_SYM", "0")   # bitwise comparisons between the kernels that compute every block
    res = {}
    for wtab in ("on", "off"):
        if wtab == "off":
            monkeypatch.setenv("IGAB_NO_WTAB", "1")
        else:
            monkeypatch.delenv("IGAB_NO_WTAB", raising=False)
        P = gpu_patch(spec)
        for ws in ("0", "1"):
            monkeypatch.setenv("IGAB_GRAM_WS", ws)
            res[(wtab, ws)] = P.assemble(kind, [x] * 3, [w] * 3, D=D, fconst=1.3)
    monkeypatch.delenv("IGAB_NO_WTAB", raising=False)
    ref = res[("off", "0")]
    assert np.array_equal(res[("off", "1")][0], ref[0]) and np.array_equal(res[("on", "1")][0], res[("on", "0")][0])
    n = ref[0].shape[1]
    assert relerr(res[("on", "0")][0].reshape(-1, n * n), ref[0].reshape(-1, n * n)) <= 1e-13
    assert relerr(res[("on", "0")][1], ref[1]) <= 1e-13
    O = port_for(spec)
    for e in (0, 17, 63):
        Kr, fr = O.assemble(kind, [x] * 3, [w] * 3, D=D, fconst=1.3, elem_begin=e, elem_end=e + 1)
        assert relerr(res[("on", "0")][0][e].reshape(1, -1), Kr.reshape(1, -1)) <= TOL
        assert relerr(res[("on", "0")][1][e].reshape(1, -1), fr.reshape(1, -1)) <= TOL
    # weights perturbed through igab_patch_update_control_points: not a tensor product any more
    P = gpu_patch(spec)
    cp2 = PD.perturbed(pd, seed=4, amp=0.0).cp
    assert np.abs(cp2[:, 3] - pd.cp[:, 3]).max() > 0.1
    P.updateControlPoints(cp2)
    K2, _ = P.assemble(kind, [x] * 3, [w] * 3, D=D, elem_end=2)
    O2 = PT.PortPatch(pd.degree, pd.knotSpans, cp2, 3)
    Kr2, _ = O2.assemble(kind, [x] * 3, [w] * 3, D=D, elem_end=2)
    assert relerr(K2.reshape(2, -1), Kr2.reshape(2, -1)) <= TOL
    P.updateControlPoints(pd.cp)  # ... and back
    K3, _ = P.assemble(kind, [x] * 3, [w] * 3, D=D, elem_end=2)
    assert np.array_equal(K3, res[("on", "1")][0][:2]) or relerr(K3.reshape(2, -1), ref[0][:2].reshape(2, -1)) <= 1e-13


@pytest.mark.parametrize("p,q", [(3, 4), (3, 5), (4, 5), (4, 6), (4, 7), (5, 6), (5, 7)])
@pytest.mark.parametrize("kind", [capi.ASM_POISSON, capi.ASM_MASS])
@pytest.mark.parametrize("wtab", [True, False], ids=["tensor_product_weights", "general_weights"])
def test_symmetric_assembly_kernel_matches_oracle_and_the_full_kernel(p, q, kind, wtab, monkeypatch):
    """gram_sym_kernel (assemble_sym.cu: blocks j2 >= i2 contracted in paired passes, strictly upper blocks mirrored from a
    shared-memory buffer) forced with IGAB_GRAM_SYM=1 wherever it is compiled, against the oracle
    (localAssembler, dune/python/test/poisson.py:37-81) and against gram_sf_kernel, which computes every block: more
    elements than resident blocks, load vector, both weight paths (weights folded into the tables / scaled epilogue),
    exact symmetry of the mirrored blocks."""
    import torch
    if kind == capi.ASM_POISSON and (p, q) == (5, 7):
        pytest.skip("p = 5 with 7 points: the Poisson operands do not fit 227 KB of shared memory (mass only)")
    lev = (3, 3, 2) if p < 5 else (2, 2, 2)
    pd = PD.thickCylinder(p, lev)
    rng = np.random.default_rng(p * 10 + q)
    cp = pd.cp.copy()
    cp[:, :3] += rng.uniform(-0.01, 0.01, (cp.shape[0], 3))               # weights stay a tensor product
    if not wtab:
        cp[:, 3] *= rng.uniform(0.7, 1.3, cp.shape[0])
    spec = dict(degree=pd.degree, knots=pd.knotSpans, cp=cp, dimworld=3)
    P, O = gpu_patch(spec), port_for(spec)
    x, w = PD.gaussLegendre01(q)
    n = P.nb
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("IGAB_GRAM_SYM", mode)
        Ke = torch.full((P.nelem, n, n), float("nan"), dtype=torch.float64, device="cuda")
        fe = torch.full((P.nelem, n), float("nan"), dtype=torch.float64, device="cuda")
        l0 = capi.launch_count()
        P.assemble(kind, [x] * 3, [w] * 3, fconst=0.7, Ke=Ke, fe=fe)
        torch.cuda.synchronize()
        assert capi.launch_count() > l0
        out[mode] = (Ke.cpu().numpy(), fe.cpu().numpy())
    K0, K1 = out["0"][0], out["1"][0]
    assert np.isfinite(K1).all() and np.isfinite(out["1"][1]).all()       # every entry written
    scale = np.abs(K0).reshape(P.nelem, -1).max(axis=1)[:, None, None]
    assert (np.abs(K1 - K0) / scale).max() <= 1e-13
    assert np.array_equal(out["0"][1], out["1"][1])                      # the load vector code is the same
    nb1 = p + 1
    B = K1.reshape(P.nelem, nb1, nb1 * nb1, nb1, nb1 * nb1)               # [e][i2][i01][j2][j01]
    for i2 in range(nb1):
        for j2 in range(i2 + 1, nb1):
            assert np.array_equal(B[:, j2, :, i2, :], np.swapaxes(B[:, i2, :, j2, :], 1, 2)), (i2, j2)
    sample = [0, 1, 63, P.nelem // 2, P.nelem - 1]
    for e in sample:
        Kr, fr = O.assemble(kind, [x] * 3, [w] * 3, fconst=0.7, elem_begin=int(e), elem_end=int(e) + 1)
        assert relerr(K1[e].reshape(1, -1), Kr.reshape(1, -1)) <= TOL
        assert relerr(out["1"][1][e].reshape(1, -1), fr.reshape(1, -1)) <= TOL


def test_programmatic_dependent_launches_keep_the_stream_order():
    """The 2-D evaluation kernels are launched with the programmatic-stream-serialization attribute (the next launch is
    resident while the current one drains and waits only before its first store).  What must still hold on one stream:
    launches queued back to back into the SAME outputs leave the last launch's values; a control-net update between two
    launches is seen by the second one; a rule switch (tables rebuilt: plain launch) right behind a launch is correct;
    point lists produced on the stream just before the call are read after they are complete.  Compared with the same
    calls made one by one with a synchronisation after each (IGAB_NO_PDL=1 for the reference values)."""
    import os
    import torch
    spec = PFT.small_patch("plate_hole_p2", level=4)
    pd = spec["patch"]
    cpA = np.ascontiguousarray(pd.cp)
    cpB = PD.perturbed(pd, seed=11, amp=0.02).cp
    fields = ("N", "dN", "x", "Jt", "detJ")
    x3, x4 = PD.gaussLegendre01(3)[0], PD.gaussLegendre01(4)[0]

    ref = {}
    os.environ["IGAB_NO_PDL"] = "1"
    try:
        for key, cp, xi in (("A3", cpA, x3), ("B3", cpB, x3), ("B4", cpB, x4)):
            ref[key] = capi.Patch(pd.degree, pd.knotSpans, cp, pd.dimworld).evalTensor([xi, xi], order=1, want=fields)
    finally:
        del os.environ["IGAB_NO_PDL"]
    P = capi.Patch(pd.degree, pd.knotSpans, cpA, pd.dimworld)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        sh3, sh4 = P.shapes(P.nelem * 9), P.shapes(P.nelem * 16)
        o3 = {f: torch.full(sh3[f], float("nan"), dtype=torch.float64, device="cuda") for f in fields}
        o4 = {f: torch.full(sh4[f], float("nan"), dtype=torch.float64, device="cuda") for f in fields}
        P.evalTensor([x3, x3], order=1, want=fields, out=o3, stream=s.cuda_stream)   # builds the tables (plain launch)
        for _ in range(8):                                                            # resident rule: dependent launches
            P.evalTensor([x3, x3], order=1, want=fields, out=o3, stream=s.cuda_stream)
        P.updateControlPoints(cpB)                                                    # copy between two launches
        for _ in range(4):
            P.evalTensor([x3, x3], order=1, want=fields, out=o3, stream=s.cuda_stream)
        P.evalTensor([x4, x4], order=1, want=fields, out=o4, stream=s.cuda_stream)   # rule switch right behind a launch
        s.synchronize()
        for f in fields:
            assert np.array_equal(o3[f].cpu().numpy(), ref["B3"][f]), f
            assert np.array_equal(o4[f].cpu().numpy(), ref["B4"][f]), f
        # point lists written on the stream immediately before the call
        rng = np.random.default_rng(3)
        npts = 40000
        elem = torch.from_numpy(np.sort(rng.integers(0, P.nelem, npts)).astype(np.int32)).cuda()
        xi_h = rng.random((npts, 2))
        shl = P.shapes(npts)
        ol = {f: torch.empty(shl[f], dtype=torch.float64, device="cuda") for f in fields}
        xi_d = torch.zeros((npts, 2), dtype=torch.float64, device="cuda")
        src = torch.from_numpy(xi_h).cuda()
        s.synchronize()
        P.evalTensor([x3, x3], order=1, want=fields, out=o3, stream=s.cuda_stream)   # a launch that triggers early ...
        xi_d.copy_(src)                                                               # ... a producer kernel / copy ...
        xi_d.mul_(1.0)                                                                # (a kernel writing the list)
        P.evalPoints(elem, xi_d, order=1, want=fields, out=ol, stream=s.cuda_stream)  # ... and the consumer
        s.synchronize()
    got = {f: ol[f].cpu().numpy() for f in fields}
    want = capi.Patch(pd.degree, pd.knotSpans, cpB, pd.dimworld).evalPoints(elem.cpu().numpy(), xi_h, order=1, want=fields)
    for f in fields:
        assert np.array_equal(got[f], want[f]), f


@pytest.mark.parametrize("p", [3, 4])
@pytest.mark.parametrize("kind", [capi.ASM_POISSON, capi.ASM_MASS, capi.ASM_ELASTICITY])
def test_warp_specialised_kernel_bulk_store_epilogue(p, kind, monkeypatch):
    """gram_ws_kernel with tensor-product weights: the rows of K_e leave the tile buffer through async bulk stores issued by
    the DMMA warps (row start shifted to the 16-byte phase of its destination; rows of 125 / 375 doubles alternate in
    phase, rows of 64 / 192 do not) against the helpers' copy loop (IGAB_WS_BULK=0): same bits for K_e and the load vector,
    also when K_e starts at an address that is 8 mod 16 (nothing written outside the view), and the oracle on a sample --
    with an isotropic D (zero coefficients skipped in the point matrices) and a full D (nothing skipped)."""
    import torch
    pd = PD.thickCylinder(p, (2, 1, 2))
    spec = dict(degree=pd.degree, knots=pd.knotSpans, cp=pd.cp, dimworld=3)
    x, w = PD.gaussLegendre01(p + 1)
    monkeypatch.setenv("IGAB_ELAS_SF", "1")
    monkeypatch.setenv("IGAB_GRAM_WS", "1")
    monkeypatch.setenv("IGAB_GRAM_SYM", "0")
    P, O = gpu_patch(spec), port_for(spec)
    n = P.nb * (3 if kind == capi.ASM_ELASTICITY else 1)
    rng = np.random.default_rng(7)
    A = rng.uniform(-1, 1, (6, 6))
    lam, mu = 0.6, 0.4
    Diso = np.zeros((6, 6)); Diso[:3, :3] = lam; Diso[np.arange(3), np.arange(3)] = lam + 2 * mu; Diso[np.arange(3, 6), np.arange(3, 6)] = mu
    for D in ([Diso, A @ A.T + np.eye(6)] if kind == capi.ASM_ELASTICITY else [None]):
        out = {}
        for mode in ("0", "1"):
            monkeypatch.setenv("IGAB_WS_BULK", mode)
            for off in (0, 1):
                buf = torch.full((P.nelem * n * n + 2,), -7.0, dtype=torch.float64, device="cuda")
                Ke = buf[off:off + P.nelem * n * n].view(P.nelem, n, n)
                assert Ke.data_ptr() % 16 == 8 * off
                fe = torch.full((P.nelem, n), float("nan"), dtype=torch.float64, device="cuda")
                P.assemble(kind, [x] * 3, [w] * 3, D=D, fconst=0.9, Ke=Ke, fe=fe)
                torch.cuda.synchronize()
                assert float(buf[-1]) == -7.0 and float(buf[-2 + off]) == -7.0 and (off == 0 or float(buf[0]) == -7.0)
                out[(mode, off)] = (Ke.cpu().numpy(), fe.cpu().numpy())
        ref = out[("0", 0)]
        assert np.isfinite(ref[0]).all() and np.isfinite(ref[1]).all()
        for key, (K, f) in out.items():
            assert np.array_equal(K, ref[0]) and np.array_equal(f, ref[1]), key
        for e in (0, 1, P.nelem - 1):
            Kr, fr = O.assemble(kind, [x] * 3, [w] * 3, D=D, fconst=0.9, elem_begin=e, elem_end=e + 1)
            assert relerr(ref[0][e].reshape(1, -1), Kr.reshape(1, -1)) <= TOL
            assert relerr(ref[1][e].reshape(1, -1), fr.reshape(1, -1)) <= TOL


@pytest.mark.parametrize("name", ["plate_hole_p2", "cylinder_p3", "square_p3"])
def test_output_pointers_that_are_only_8_byte_aligned(name):
    """The async bulk stores (TMA) of the tensor-rule kernels need 16-byte aligned global rows; a caller's double* is only
    guaranteed 8-byte alignment.  Every field is written once into freshly allocated tensors and once into views that
    start one double into a buffer (address = 8 mod 16): same bits, nothing written outside the views."""
    import torch
    spec = PFT.small_patch(name, level=2)
    P = gpu_patch(spec)
    q = spec["degree"][0] + 1
    xi = [PD.gaussLegendre01(q)[0]] * P.dim
    want = ("N", "dN", "x", "Jt", "detJ", "JinvT")
    npts = P.nelem * q ** P.dim
    sh = P.shapes(npts)
    ref = {f: torch.empty(sh[f], dtype=torch.float64, device="cuda") for f in want}
    P.evalTensor(xi, order=1, want=want, out=ref)
    bufs, views = {}, {}
    for f in want:
        n = int(np.prod(sh[f]))
        bufs[f] = torch.full((n + 2,), -7.0, dtype=torch.float64, device="cuda")
        assert bufs[f].data_ptr() % 16 == 0
        views[f] = bufs[f][1:n + 1].view(sh[f])
        assert views[f].data_ptr() % 16 == 8
    P.evalTensor(xi, order=1, want=want, out=views)
    torch.cuda.synchronize()
    for f in want:
        assert torch.equal(views[f], ref[f]), f
        assert float(bufs[f][0]) == -7.0 and float(bufs[f][-1]) == -7.0, f

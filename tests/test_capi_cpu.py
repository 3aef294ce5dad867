"""CPU-side checks of the C-ABI (-m "not gpu"): the in-tree library loads, exports every symbol include/igab200.h
declares, validates arguments before touching CUDA, and fails loudly (no CPU fallback) without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import igab200  # noqa: F401
from dune_iga_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "igab200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(igab_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), n
    assert set(names) == set(capi.EXPORTS)
    assert b"sm_100a" in L.igab_version()


def test_only_cuda_code_in_the_library():
    # the product must not link the oracle
    import subprocess
    out = subprocess.run(["ldd", capi.SO_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "libcudart" in out


def test_argument_validation_before_cuda():
    L = capi.lib()
    h = C.c_void_p()
    deg = (C.c_int * 1)(2)
    nk = (C.c_int * 1)(5)
    kn = np.array([0, 0, 1, 1, 1.0])
    kp = (capi.dp * 1)(kn.ctypes.data_as(capi.dp))
    cp = np.ones((2, 3))
    # closed/short knot vector: the reference asserts open knot vectors (bsplinealgorithms.hh:98-99)
    st = L.igab_patch_create(1, 2, deg, nk, kp, cp.ctypes.data_as(capi.dp), None, C.byref(h))
    assert st == 1 and b"knot" in L.igab_last_error_string()
    st = L.igab_patch_create(4, 4, deg, nk, kp, cp.ctypes.data_as(capi.dp), None, C.byref(h))
    assert st == 1
    assert L.igab_patch_sizes(None, None, None, None, None) == 1


def test_no_cpu_fallback_without_device():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("device present")
    except ImportError:
        pass
    with pytest.raises(capi.IgabError) as ei:
        capi.Patch([1], [[0, 0, 1, 1]], [[0, 0, 1], [1, 0, 1]], 2)
    assert ei.value.status in (3, 2)


def test_local_keys_match_oracle_and_reference_diagram():
    """igab_local_keys (host integers, no device) == the oracle's restatement of NurbsLocalCoefficients::init
    (nurbsbasis.hh:220-271) for every degree combination up to 4, and the p=(2,2) table read off the reference's edge /
    vertex diagram (:182-216)."""
    import itertools
    import portbind as PT
    for dim in (1, 2, 3):
        for deg in itertools.product(range(0, 5), repeat=dim):
            got = capi.localKeys(deg)
            ref = PT.local_keys(deg)
            for a, b in zip(got, ref):
                assert np.array_equal(a, b), (deg, a, b)
    sub, cod, idx = capi.localKeys([2, 2])
    assert list(sub) == [0, 2, 1, 0, 0, 1, 2, 3, 3]
    assert list(cod) == [2, 1, 2, 1, 0, 1, 2, 1, 2]
    assert list(idx) == [0] * 9
    sub, cod, idx = capi.localKeys([3, 2])  # two inner dofs on the horizontal edges
    assert list(idx) == [0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0]
    assert list(capi.localKeys([3])[0]) == [0, 0, 0, 1]
    # (subEntity, codim, index) identifies a shape function uniquely in 1-D and 2-D
    for deg in ([4], [3, 4], [2, 2]):
        keys = set(zip(*[k.tolist() for k in capi.localKeys(deg)]))
        assert len(keys) == int(np.prod(np.array(deg) + 1))
    with pytest.raises(capi.IgabError):
        capi.localKeys([9])


def test_slab_partition_and_interface_row_planner_host_only():
    """igab_partition_slab / igab_interface_row_ranges are pure host integer code (no device): the planner's row range of
    a slab is exactly the set of rows the slab's elements touch through the oracle's DOF map (nurbsbasis.hh:576), also
    with repeated interior knots, vector-valued numbering, custom layer bounds, thin slabs and empty ranks."""
    import portbind as PT
    from dune_iga_b200 import patchdata as PD
    rng = np.random.default_rng(3)
    cases = [([2, 3], [[0, 0, 0, .2, .2, .5, .7, 1, 1, 1], [0, 0, 0, 0, .25, .5, .5, .75, 1, 1, 1, 1]]),
             ([3, 2, 3], [[0, 0, 0, 0, .5, 1, 1, 1, 1], [0, 0, 0, .3, 1, 1, 1],
                          [0, 0, 0, 0, .1, .2, .3, .3, .6, .8, .9, 1, 1, 1, 1]])]
    for deg, kn in cases:
        dim = len(deg)
        ncp = [len(kn[d]) - deg[d] - 1 for d in range(dim)]
        cp = np.concatenate([rng.random((int(np.prod(ncp)), dim)), np.ones((int(np.prod(ncp)), 1))], axis=1)
        O = PT.PortPatch(deg, kn, cp, dim)
        dof, ndof = O.dofmap()
        nel = [len(np.unique(k)) - 1 for k in kn]
        per_layer = int(np.prod(nel[:-1]))
        for world in (1, 2, 3, nel[-1], nel[-1] + 2):
            bounds = [capi.partitionSlab(nel, world, r) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == O.nelem
            assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
            for ncomp in (1, 3):
                rb, re = capi.interfaceRowRanges(deg, kn, world, ncomp)
                for r, (eb, ee) in enumerate(bounds):
                    if ee == eb:
                        assert rb[r] == re[r]
                        continue
                    t = np.unique(dof[eb:ee])
                    assert rb[r] == t.min() * ncomp and re[r] == (t.max() + 1) * ncomp
                    assert len(t) * ncomp == re[r] - rb[r]  # contiguous: whole DOF layers
        # custom bounds on layer boundaries; anything else is refused
        eb = [0, per_layer, per_layer, O.nelem]
        rb, re = capi.interfaceRowRanges(deg, kn, 3, 1, eb)
        assert rb[1] == re[1] and rb[0] == 0 and re[2] == ndof
        with pytest.raises(capi.IgabError):
            capi.interfaceRowRanges(deg, kn, 2, 1, [0, 1, O.nelem])
    # thin slabs: 4 element layers of degree 3 over 4 ranks -> ranks 0 and 3 share DOF layers without being neighbours
    pd = PD.thickCylinder(3, (1, 1, 2))
    rb, re = capi.interfaceRowRanges(pd.degree, pd.knotSpans, 4)
    assert min(re[0], re[3]) > max(rb[0], rb[3])


def test_partition_weighted_balances_cumulative_point_counts():
    """igab_partition_weighted (trimmed patches are partitioned by cumulative point count): every boundary sits
    where the running count first reaches r / world of the total, bounds are monotone, rounding to whole layers, empty
    input, and the error path -- host integer code, no device needed."""
    rng = np.random.default_rng(3)
    counts = rng.integers(0, 241, 4096).astype(np.int64)
    counts[rng.random(4096) < 0.3] = 0            # empty elements
    cum = np.concatenate([[0], np.cumsum(counts)])
    for world in (1, 2, 3, 8):
        b = capi.partitionWeighted(counts, world)
        assert b[0] == 0 and b[-1] == 4096 and all(x <= y for x, y in zip(b, b[1:]))
        for r in range(1, world):
            assert cum[b[r]] * world >= cum[-1] * r and (b[r] == 0 or cum[b[r] - 1] * world < cum[-1] * r)
        work = [cum[b[r + 1]] - cum[b[r]] for r in range(world)]
        assert max(work) - min(work) <= 2 * 240
        bl = capi.partitionWeighted(counts, world, per_layer=64)
        assert all(x % 64 == 0 for x in bl) and bl[0] == 0 and bl[-1] == 4096
        assert all(abs(x - y) <= 32 for x, y in zip(bl, b))
    assert capi.partitionWeighted(np.zeros(0, np.int64), 4) == [0, 0, 0, 0, 0]
    assert capi.partitionWeighted(np.zeros(10, np.int64), 2) == [0, 0, 10]
    with pytest.raises(Exception):
        capi.partitionWeighted(np.array([1, -1], np.int64), 2)


def test_register_budget_of_the_role_split_kernel():
    """gram_ws_kernel moves registers from its helper warps to its DMMA warps (setmaxnreg); ptxas reports only the launch-time
    count, so the SASS is checked: nothing a helper warp can reach names a register outside the helpers' budget, nothing at
    all one outside the DMMA warps' (tools/ws_regs_check.py)."""
    import shutil
    import subprocess
    import sys
    obj = os.path.join(ROOT, "dune-iga_b200", "csrc", "_build", "assemble_fused.o")
    if not (shutil.which("cuobjdump") and os.path.exists(obj)):
        pytest.skip("needs cuobjdump and the built object file")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ws_regs_check.py"), obj], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.count("ok") >= 6, r.stdout + r.stderr

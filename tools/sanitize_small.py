"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck)."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import igab200  # noqa
import patches_for_tests as PFT
from dune_iga_b200 import capi, patchdata as PD, quadrature as Q

for name in ["curve_p3", "plate_hole_p2", "cylinder_p3", "cylinder_p4", "torus_p2", "square_p3", "mixed_deg_3d"]:
    spec = PFT.small_patch(name)
    P = capi.Patch(spec["degree"], spec["knots"], spec["cp"], spec["dimworld"])
    xi = [Q.gaussLegendre01(p + 1)[0] for p in spec["degree"]]
    w = [Q.gaussLegendre01(p + 1)[1] for p in spec["degree"]]
    for mode in (capi.KERNEL_GENERIC, capi.KERNEL_AUTO):
        P.setKernel(mode)
        o = P.evalTensor(xi, order=2)
        o = P.evalTensor(xi, order=1, elem_begin=1, elem_end=P.nelem - 1)
    rng = np.random.default_rng(1)
    el = rng.integers(0, P.nelem, 37).astype(np.int32)
    pts = rng.random((37, P.dim))
    P.evalPoints(el, pts, order=2)
    P.partial(el, pts, [2] + [1] * (P.dim - 1))
    P.spans(); P.dofmap()
    if P.dim >= 2:
        P.volume(xi, w)
        P.assemble(capi.ASM_POISSON, xi, w, elem_end=min(P.nelem, 5))
        P.assemble(capi.ASM_MASS, xi, w, elem_end=min(P.nelem, 5))
    print(name, "ok", flush=True)
flags = np.zeros(16, np.uint8); flags[:3] = 1
sp = PFT.small_patch("square_p3")
capi.Patch(sp["degree"], sp["knots"], sp["cp"], 2, trimflags=flags).dofmap()
print("done")

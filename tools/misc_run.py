"""Short run for ncu: the second-derivative 3-D kernel (p=3, 64x64x8 elements), the fused 2-D assembly kernel (p=2 Poisson,
1024^2 elements), the C1 evaluation kernel (p=2, 256^2) and the fused norms kernel."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igab200  # noqa: E402,F401
from dune_iga_b200 import capi, patchdata as PD  # noqa: E402

s = torch.cuda.current_stream().cuda_stream
P3 = capi.Patch.fromPatchData(PD.thickCylinder(3, (6, 6, 6)))
x4 = PD.gaussLegendre01(4)
nel = 64 * 64 * 8
f3 = ("N", "dN", "d2N", "x", "Jt", "H", "detJ")
sh = P3.shapes(nel * 64)
o3 = {f: torch.empty(sh[f], dtype=torch.float64, device="cuda") for f in f3}
P2 = capi.Patch.fromPatchData(PD.plateWithHole(9, 10))
x3 = PD.gaussLegendre01(3)
K2 = torch.empty((P2.nelem, 9, 9), dtype=torch.float64, device="cuda")
P1 = capi.Patch.fromPatchData(PD.plateWithHole(7, 8))
f1 = ("N", "dN", "x", "Jt", "detJ")
sh1 = P1.shapes(P1.nelem * 9)
o1 = {f: torch.empty(sh1[f], dtype=torch.float64, device="cuda") for f in f1}
u = torch.randn(int(np.prod([n + 3 for n in PD.thickCylinder(3, (6, 6, 6)).elementsPerDirection])), dtype=torch.float64, device="cuda")
for _ in range(3):
    P3.evalTensor([x4[0]] * 3, order=2, elem_end=nel, want=f3, out=o3, stream=s)
    P2.assemble(capi.ASM_POISSON, [x3[0]] * 2, [x3[1]] * 2, Ke=K2, fe=None, stream=s)
    P1.evalTensor([x3[0]] * 2, order=1, want=f1, out=o1, stream=s)
    r = P3.norms([x4[0]] * 3, [x4[1]] * 3, u)
torch.cuda.synchronize()
print("ok", float(o3["H"].abs().max()), float(K2.abs().max()), float(o1["N"].max()), r)

"""Short run of the general warp-per-element evaluation kernels for ncu: 3-D p=3 with 6^3 Gauss points (no tuned kernel),
2-D p=4 with 5x5 points, and a 3-D point list grouped by element."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igab200  # noqa: E402,F401
from dune_iga_b200 import capi, patchdata as PD  # noqa: E402

F1 = ("N", "dN", "x", "Jt", "detJ")
s = torch.cuda.current_stream().cuda_stream


def tensor(pd, nq1, nel):
    P = capi.Patch.fromPatchData(pd)
    xi = [PD.gaussLegendre01(n)[0] for n in nq1]
    sh = P.shapes(nel * int(np.prod(nq1)))
    out = {f: torch.empty(sh[f], dtype=torch.float64, device="cuda") for f in F1}
    for _ in range(3):
        P.evalTensor(xi, order=1, elem_begin=0, elem_end=nel, want=F1, out=out, stream=s)
    torch.cuda.synchronize()
    return float(out["detJ"].sum())


a = tensor(PD.thickCylinder(3, (6, 6, 6)), [6, 6, 6], 64 * 64 * 8)
plate4 = PD.plateWithHole(0, 0).degreeElevate(0, 2).degreeElevate(1, 2)
plate4 = plate4.globalRefineInDirection(0, 8).globalRefineInDirection(1, 9)
b = tensor(plate4, [5, 5], plate4.numElements)
pd = PD.thickCylinder(3, (5, 5, 5))
P = capi.Patch.fromPatchData(pd)
rng = np.random.default_rng(1)
npts = 1 << 20
elem = np.repeat(rng.choice(P.nelem, npts // 40 + 1, replace=False), 40)[:npts].astype(np.int32)
d_elem, d_xi = torch.from_numpy(elem).cuda(), torch.from_numpy(rng.random((npts, 3))).cuda()
sh = P.shapes(npts)
out = {f: torch.empty(sh[f], dtype=torch.float64, device="cuda") for f in F1}
for _ in range(3):
    P.evalPoints(d_elem, d_xi, order=1, want=F1, out=out, stream=s)
torch.cuda.synchronize()
print("ok", a, b, float(out["detJ"].sum()))

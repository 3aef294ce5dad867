"""Throughput of the patch reductions (volume, L2/energy norms) on the C2 patch (p=3, 128^3 elements, 4^3 Gauss)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igab200  # noqa: E402,F401
from dune_iga_b200 import capi, patchdata as PD  # noqa: E402

lev = int(sys.argv[1]) if len(sys.argv) > 1 else 7
pd = PD.thickCylinder(3, (lev, lev, lev))
P = capi.Patch.fromPatchData(pd)
x, w = PD.gaussLegendre01(4)
ndof = P.dofmap()[1] if lev <= 5 else int(np.prod([n + 3 for n in pd.elementsPerDirection]))
u = torch.randn(ndof, dtype=torch.float64, device="cuda")
npts = P.nelem * 64
for name, f in (("volume", lambda: P.volume([x] * 3, [w] * 3)), ("norms", lambda: P.norms([x] * 3, [w] * 3, u))):
    f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = f()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%s: %.1f ms, %.3g points/s, result %s" % (name, dt * 1e3, npts / dt, r))

"""3-D element-matrix throughput at p=2 (27 functions, 3^3 Gauss) and p=1."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igab200  # noqa: E402,F401
from dune_iga_b200 import capi, patchdata as PD  # noqa: E402
from tools.bench_configs import timed  # noqa: E402

lam, mu = 0.3 / (1.3 * 0.4), 1 / 2.6
D = np.zeros((6, 6)); D[:3, :3] = lam; D[np.arange(3), np.arange(3)] = lam + 2 * mu; D[np.arange(3, 6), np.arange(3, 6)] = mu
for p, lev in ((2, 6), (3, 5)):
    pd = PD.thickCylinder(p, (lev, lev, lev))
    P = capi.Patch.fromPatchData(pd)
    x, w = PD.gaussLegendre01(p + 1)
    for kind, kname, nc in ((capi.ASM_POISSON, "Poisson", 1), (capi.ASM_MASS, "mass", 1), (capi.ASM_ELASTICITY, "elasticity", 3)):
        n = P.nb * nc
        nel = min(P.nelem, (2 << 30) // (n * n * 8))
        Ke = torch.empty((nel, n, n), dtype=torch.float64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        ms = timed(lambda: P.assemble(kind, [x] * 3, [w] * 3, D=D if kind == capi.ASM_ELASTICITY else None, elem_end=nel, Ke=Ke, fe=None, stream=st), reps=3, warm=1)
        print("3-D p=%d %-10s n=%3d  %d elements %8.3f ms  %.3g el/s  output %.0f GB/s" % (p, kname, n, nel, ms, nel / ms * 1e3, nel * n * n * 8 / ms / 1e6), flush=True)

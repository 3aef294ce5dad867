"""2-D element-matrix throughput (general B-stack path vs fused small-n kernel): plate with hole p=2, 3x3 Gauss; square p=3, 4x4 Gauss."""
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
D2 = np.array([[lam + 2 * mu, lam, 0], [lam, lam + 2 * mu, 0], [0, 0, mu]])
for name, pd, q in (("plate p=2 1024^2", PD.plateWithHole(9, 10), 3), ("square p=3 512^2", PD.squarePlate(3, 9), 4)):
    P = capi.Patch.fromPatchData(pd)
    x, w = PD.gaussLegendre01(q)
    for kind, kname, nc in ((capi.ASM_POISSON, "Poisson", 1), (capi.ASM_MASS, "mass", 1), (capi.ASM_ELASTICITY, "elasticity", 2)):
        n = P.nb * nc
        nel = P.nelem
        Ke = torch.empty((nel, n, n), dtype=torch.float64, device="cuda")
        fe = torch.empty((nel, n), dtype=torch.float64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        ms = timed(lambda: P.assemble(kind, [x, x], [w, w], D=D2 if kind == capi.ASM_ELASTICITY else None, Ke=Ke, fe=fe, stream=st), reps=3, warm=1)
        gb = nel * (n * n + n) * 8 / 1e9
        print("%-18s %-10s n=%2d  %8.3f ms  %.3g el/s  output %.0f GB/s" % (name, kname, n, ms, nel / ms * 1e3, gb / ms * 1e3), flush=True)

"""Elasticity element matrices at p = 4 / 5 points (gram_ws_kernel, the default there) with the rows of K_e leaving by async
bulk stores issued by the DMMA warps (default) against the helpers' copy loop (IGAB_WS_BULK=0), and at p = 3 / 4 points
(gram_sf_kernel, the default there): element matrices/s from CUDA events, bitwise comparison of the two epilogues.
IGAB200_SO=<other build> runs the same sweep on another build of the library."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igab200  # noqa: E402,F401
from dune_iga_b200 import capi, patchdata as PD  # noqa: E402


def timed(f, reps=7):
    f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


lam, mu = 0.3 / (1.3 * 0.4), 1 / 2.6
D = np.zeros((6, 6))
D[:3, :3] = lam
D[np.arange(3), np.arange(3)] = lam + 2 * mu
D[np.arange(3, 6), np.arange(3, 6)] = mu
for p, q, lev, nel in ((4, 5, 5, 148 * 48), (3, 4, 5, 148 * 96)):
    pd = PD.thickCylinder(p, (lev, lev, lev))
    P = capi.Patch.fromPatchData(pd)
    x, w = PD.gaussLegendre01(q)
    n = P.nb * 3
    res, out = {}, {}
    for mode in ("0", "1"):
        os.environ["IGAB_WS_BULK"] = mode
        Ke = torch.zeros((nel, n, n), dtype=torch.float64, device="cuda")
        s = torch.cuda.current_stream().cuda_stream
        ms = timed(lambda: P.assemble(capi.ASM_ELASTICITY, [x] * 3, [w] * 3, D=D, elem_end=nel, Ke=Ke, fe=None, stream=s))
        res[mode] = nel / ms * 1e3
        out[mode] = Ke
    same = bool(torch.equal(out["0"], out["1"]))
    print(json.dumps({"p": p, "q": q, "elements": nel, "el_per_s by IGAB_WS_BULK": res, "bitwise_equal": same}), flush=True)
    del out

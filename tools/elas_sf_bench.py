"""3-D elasticity element matrices: the nine-Gram-block kernel K5c (IGAB_ELAS_SF=0) against the sum-factorised kernels
(IGAB_ELAS_SF=1: gram_sf_kernel / gram_ws_kernel) per degree and rule, element matrices/s from CUDA events."""
import json
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igab200  # noqa
from dune_iga_b200 import capi, patchdata as PD
lam, mu = 0.3 / (1.3 * 0.4), 1 / 2.6
D = np.zeros((6, 6)); D[:3, :3] = lam; D[np.arange(3), np.arange(3)] = lam + 2 * mu; D[np.arange(3, 6), np.arange(3, 6)] = mu
for p, q, nel in ((2, 3, 148 * 64), (2, 4, 148 * 64), (3, 4, 148 * 32), (3, 5, 148 * 32), (4, 5, 148 * 16)):
    pd = PD.thickCylinder(p, (5, 5, 5))
    P = capi.Patch.fromPatchData(pd)
    x, w = PD.gaussLegendre01(q)
    n = P.nb * 3
    Ke = torch.empty((nel, n, n), dtype=torch.float64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    res = {}
    for mode in ("0", "1"):
        os.environ["IGAB_ELAS_SF"] = mode
        f = lambda: P.assemble(capi.ASM_ELASTICITY, [x] * 3, [w] * 3, D=D, elem_end=nel, Ke=Ke, fe=None, stream=s)
        f(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[mode] = nel / float(np.median(ts)) * 1e3
    print(json.dumps({"p": p, "q": q, "elements": nel, "nine_gram_blocks_el_per_s": round(res["0"]), "sum_factorised_el_per_s": round(res["1"]),
                      "ratio": round(res["1"] / res["0"], 3)}), flush=True)
    del Ke

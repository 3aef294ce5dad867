"""Barrier-separated (IGAB_GRAM_WS=0) against warp-specialised (IGAB_GRAM_WS=1) sum-factorised assembly kernel:
bitwise comparison on whole chunks and element matrices/s from CUDA events (context for profiles/README.md)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igab200  # noqa: E402,F401
from dune_iga_b200 import capi, patchdata as PD  # noqa: E402


def timed(f, reps=5):
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
for p, q, lev in ((4, 5, 5), (3, 4, 5)):
    pd = PD.thickCylinder(p, (lev, lev, lev))
    P = capi.Patch.fromPatchData(pd)
    x, w = PD.gaussLegendre01(q)
    for kind, name, nel in ((capi.ASM_POISSON, "poisson", 148 * 32), (capi.ASM_MASS, "mass", 148 * 32),
                            (capi.ASM_ELASTICITY, "elasticity", 148 * 16)):
        n = P.nb * (3 if kind == capi.ASM_ELASTICITY else 1)
        res = {}
        out = {}
        for mode in ("0", "1"):
            os.environ["IGAB_GRAM_WS"] = mode
            os.environ["IGAB_ELAS_SF"] = "1"
            Ke = torch.zeros((nel, n, n), dtype=torch.float64, device="cuda")
            fe = torch.zeros((nel, n), dtype=torch.float64, device="cuda")
            s = torch.cuda.current_stream().cuda_stream
            ms = timed(lambda: P.assemble(kind, [x] * 3, [w] * 3, D=D if kind == capi.ASM_ELASTICITY else None,
                                          elem_end=nel, Ke=Ke, fe=(fe if os.environ.get("WS_FE") else None), stream=s))
            res[mode] = nel / ms * 1e3
            out[mode] = (Ke, fe)
        same = bool(torch.equal(out["0"][0], out["1"][0]) and torch.equal(out["0"][1], out["1"][1]))
        err = float((out["0"][0] - out["1"][0]).abs().max() / out["0"][0].abs().max())
        print(json.dumps({"p": p, "q": q, "kind": name, "elements": nel, "barrier_separated_el_per_s": res["0"],
                          "warp_specialised_el_per_s": res["1"], "speedup": res["1"] / res["0"], "bitwise_equal": same,
                          "max_rel_diff": err}), flush=True)
        del out

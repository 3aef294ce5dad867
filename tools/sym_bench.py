"""gram_sf_kernel (IGAB_GRAM_SYM=0: all blocks) against gram_sym_kernel (upper block triangle + mirror): max relative
difference, symmetry defect and element matrices/s from CUDA events.  usage: python tools/sym_bench.py [fe]"""
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


want_fe = len(sys.argv) > 1 and sys.argv[1] == "fe"
cases = ((4, 5, 5, 148 * 64), (3, 4, 5, 148 * 128), (4, 6, 4, 148 * 16), (4, 7, 4, 148 * 16), (3, 5, 5, 148 * 64), (5, 6, 4, 148 * 8))
if os.environ.get("SYM_CASES"):
    cases = cases[:int(os.environ["SYM_CASES"])]
modes = os.environ.get("SYM_MODES", "0").split(",")
for p, q, lev, nel in cases:
    pd = PD.thickCylinder(p, (lev, lev, lev))
    P = capi.Patch.fromPatchData(pd)
    nel = min(nel, P.nelem)
    x, w = PD.gaussLegendre01(q)
    for kind, name in ((capi.ASM_POISSON, "poisson"), (capi.ASM_MASS, "mass")):
        n = P.nb
        res, out = {}, {}
        for mode in ["off"] + modes:
            os.environ["IGAB_GRAM_SYM"] = "0" if mode == "off" else "1"
            os.environ["IGAB_SYM_MODE"] = mode
            Ke = torch.zeros((nel, n, n), dtype=torch.float64, device="cuda")
            fe = torch.zeros((nel, n), dtype=torch.float64, device="cuda")
            s = torch.cuda.current_stream().cuda_stream
            ms = timed(lambda: P.assemble(kind, [x] * 3, [w] * 3, elem_end=nel, Ke=Ke, fe=(fe if want_fe else None), stream=s))
            res[mode] = nel / ms * 1e3
            out[mode] = (Ke, fe)
        scale = float(out["off"][0].abs().max())
        line = {"p": p, "q": q, "kind": name, "elements": nel, "all_blocks_el_per_s": round(res["off"])}
        for m in modes:
            err = float((out["off"][0] - out[m][0]).abs().max() / scale)
            ferr = float((out["off"][1] - out[m][1]).abs().max() / max(float(out["off"][1].abs().max()), 1e-300)) if want_fe else 0.0
            line["mode" + m] = {"el_per_s": round(res[m]), "speedup": round(res[m] / res["off"], 3), "max_rel_diff": err, "fe_rel_diff": ferr}
        print(json.dumps(line), flush=True)
        del out

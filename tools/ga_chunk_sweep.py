"""igab_global_assemble (device outputs) on the C3 patch for different chunk sizes (IGAB_GA_CHUNK_ELEMS): does keeping a
chunk's element matrices inside the 126 MB L2 pay for the extra launches?"""
import os
import sys
import time

import torch

sys.path.insert(0, ".")
import igab200  # noqa
from dune_iga_b200 import capi, patchdata as PD

pd = PD.thickCylinder(4, (6, 6, 6))
P = capi.Patch.fromPatchData(pd)
x, w = PD.gaussLegendre01(5)
xi, ww = [x] * 3, [w] * 3
rp, _ = P.csrPattern()
vals = torch.empty(int(rp[-1]), dtype=torch.float64, device="cuda")
b = torch.empty(len(rp) - 1, dtype=torch.float64, device="cuda")
s = torch.cuda.current_stream()


def t(f, n=3):
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ref = None
for chunk in (0, 296, 592, 888, 1184, 2368, 4096, 8192):
    if chunk:
        os.environ["IGAB_GA_CHUNK_ELEMS"] = str(chunk)
    else:
        os.environ.pop("IGAB_GA_CHUNK_ELEMS", None)
    ms = t(lambda: P.globalAssemble(capi.ASM_POISSON, xi, ww, out=(vals, b), stream=s.cuda_stream))
    if ref is None:
        ref = vals.clone()
    print("chunk %5s elements (%6.1f MB of K_e): %7.2f ms  %.3e element matrices/s  same bits %s" % (
        chunk or "dflt", (chunk or 8589) * 125 * 125 * 8 / 1e6, ms, P.nelem / ms * 1e3, bool(torch.equal(ref, vals))), flush=True)

"""3-D tensor-rule evaluation kernels beside the headline configuration: points/s and fraction of the measured copy
bandwidth (CUDA events) for p = 2 (3^3 and 4^3 Gauss), p = 3 (4^3, 5^3) and p = 4 (5^3), outputs N+dN+x+Jt+detJ."""
import json
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igab200  # noqa
from dune_iga_b200 import capi, patchdata as PD
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
F1 = ("N", "dN", "x", "Jt", "detJ")
stream = torch.cuda.current_stream()
cases = [(2, 3, 6), (2, 4, 6), (3, 4, 6), (3, 5, 6), (4, 5, 5)]
if os.environ.get("CASES"):
    cases = [cases[int(k)] for k in os.environ["CASES"].split(",")]
for p, q, lev in cases:
    pd = PD.thickCylinder(p, (lev, lev, lev))
    P = capi.Patch.fromPatchData(pd)
    P.setKernel(capi.KERNEL_TENSOR)
    xi = [PD.gaussLegendre01(q)[0]] * 3
    nel = min(P.nelem, int(12e9 / (8 * (P.nb * 4 + 13) * q ** 3)))
    npts = nel * q ** 3
    sh = P.shapes(npts)
    o = {f: torch.empty(sh[f], dtype=torch.float64, device="cuda") for f in F1}
    f = lambda: P.evalTensor(xi, order=1, want=F1, out=o, elem_end=nel, stream=stream.cuda_stream)
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); f(); e1.record(stream); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    bpp = 8 * (P.nb * 4 + 3 + 9 + 1) + 8 * P.nb * 4 / q ** 3
    print("3-D p=%d, %d^3 Gauss, %d elements: %8.3f ms  %.3g points/s  %.0f GB/s  %.3f of copy" % (
        p, q, nel, ms, npts / ms * 1e3, bpp * npts / ms / 1e6, bpp * npts / ms / 1e6 / PEAK), flush=True)
    del o
    P.close()

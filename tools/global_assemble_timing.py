import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import igab200
from dune_iga_b200 import capi, patchdata as PD
pd = PD.thickCylinder(4, (6, 6, 4))
P = capi.Patch.fromPatchData(pd)
x, w = PD.gaussLegendre01(5)
xi, ww = [x] * 3, [w] * 3
def t(f, n=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
rp, _ = P.csrPattern()
v, b = capi.pinned_empty(int(rp[-1])), capi.pinned_empty(len(rp) - 1)
print("global host pinned  %.1f ms" % t(lambda: P.globalAssemble(capi.ASM_POISSON, xi, ww, out=(v, b))))
print("global device       %.1f ms" % t(lambda: P.globalAssemble(capi.ASM_POISSON, xi, ww, device=True)))
nel = 8192
K = torch.empty((nel, 125, 125), dtype=torch.float64, device="cuda")
F = torch.empty((nel, 125), dtype=torch.float64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
print("assemble 8192 el    %.2f ms" % t(lambda: P.assemble(capi.ASM_POISSON, xi, ww, elem_end=nel, Ke=K, fe=F, stream=s)))
vals = torch.zeros(int(rp[-1]), dtype=torch.float64, device="cuda")
print("csrAssemble 8192 el %.2f ms" % t(lambda: P.csrAssemble(K, 1, 0, nel, values=vals, accumulate=True)))
import os
os.environ["IGAB_CSR_GATHER"] = "0"
print("csrAssemble (per-nnz) %.2f ms" % t(lambda: P.csrAssemble(K, 1, 0, nel, values=vals, accumulate=True)))
bb = torch.zeros(len(rp) - 1, dtype=torch.float64, device="cuda")
print("rhsAssemble 8192 el %.2f ms" % t(lambda: P.rhsAssemble(F, 1, 0, nel, b=bb, accumulate=True)))

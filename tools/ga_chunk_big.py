import os, sys, torch
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
def t(f, n=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for chunk in (4096, 8192, 16384, 32768, 65536, 131072):
    os.environ["IGAB_GA_CHUNK_ELEMS"] = str(chunk)
    ms = t(lambda: P.globalAssemble(capi.ASM_POISSON, xi, ww, out=(vals, b), device=True))
    print("chunk %6d elements (%.1f GB of K_e): %.1f ms  %.3g el/s" % (chunk, chunk * 125e3 / 1e9, ms, P.nelem / ms * 1e3), flush=True)

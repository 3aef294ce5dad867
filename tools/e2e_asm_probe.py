import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import igab200
from dune_iga_b200 import capi, patchdata as PD
sub = PD.thickCylinder(4, (6, 6, 4))
Ps = capi.Patch.fromPatchData(sub)
x, w = PD.gaussLegendre01(5)
xi, ww = [x] * 3, [w] * 3
rp, _ = Ps.csrPattern()
v, bb = capi.pinned_empty(int(rp[-1])), capi.pinned_empty(len(rp) - 1)
Ps.globalAssemble(capi.ASM_POISSON, xi, ww, out=(v, bb))
for upd in (False, True, True, False):
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        if upd:
            Ps.updateControlPoints(sub.cp)
        t1 = time.perf_counter()
        Ps.globalAssemble(capi.ASM_POISSON, xi, ww, out=(v, bb))
        t2 = time.perf_counter()
        ts.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3))
    print("update" if upd else "no update", ["%.1f + %.1f ms" % t for t in ts], flush=True)
# what bench does before: big allocations
ring = [torch.empty((16384, 125, 125), dtype=torch.float64, device="cuda") for _ in range(2)]
del ring
torch.cuda.synchronize()
free, total = torch.cuda.mem_get_info()
print("free GB", free / 1e9, "torch reserved GB", torch.cuda.memory_reserved() / 1e9)
ts = []
for _ in range(4):
    t1 = time.perf_counter()
    Ps.globalAssemble(capi.ASM_POISSON, xi, ww, out=(v, bb))
    ts.append((time.perf_counter() - t1) * 1e3)
print("after torch allocations", ["%.1f" % t for t in ts])

"""C1 (p=2, 256x256 elements, 3x3 Gauss) launch anatomy: host time of one igab_eval_tensor call through ctypes, device
time per call when calls are queued back to back, and device time of a single isolated call (CUDA events)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igab200  # noqa: E402,F401
from dune_iga_b200 import capi, patchdata as PD  # noqa: E402

pd = PD.plateWithHole(7, 8)
P = capi.Patch.fromPatchData(pd)
xi = [PD.gaussLegendre01(3)[0]] * 2
fields = ("N", "dN", "x", "Jt", "detJ")
npts = P.nelem * 9
sh = P.shapes(npts)
ring = [{f: torch.empty(sh[f], dtype=torch.float64, device="cuda") for f in fields} for _ in range(4)]  # 4 x 175 MB > L2
s = torch.cuda.current_stream().cuda_stream
for k in range(8):
    P.evalTensor(xi, order=1, want=fields, out=ring[k % 4], stream=s)
torch.cuda.synchronize()
n = 200
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(n):
    P.evalTensor(xi, order=1, want=fields, out=ring[k % 4], stream=s)
t1 = time.perf_counter()
e1.record()
torch.cuda.synchronize()
print("host us/call %.1f   device us/call (queued) %.1f" % ((t1 - t0) / n * 1e6, e0.elapsed_time(e1) / n * 1e3))
ts = []
for k in range(20):
    torch.cuda.synchronize()
    e0.record()
    P.evalTensor(xi, order=1, want=fields, out=ring[k % 4], stream=s)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print("isolated call, events around the python call: median %.1f us" % np.median(ts))
bpp = 8 * (9 + 18 + 2 + 4 + 1) + 8 * 9 * 3 / 9
print("algorithmic bytes/launch %.1f MB" % (bpp * npts / 1e6))

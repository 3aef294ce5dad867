"""Short run for ncu / timing: the 2-D evaluation kernels of configs C1, C4, C5 (tensor part and trimmed point lists),
one launch each after a warm-up; prints points/s and the fraction of MEASURED_PEAKS.json's copy bandwidth from CUDA events."""
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
reps = int(os.environ.get("REPS", "5"))
stream = torch.cuda.current_stream()


def timed(f, inner=1):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(inner):
            f()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    return float(np.median(ts))


def tensor_cfg(name, pd, nq1, order, fields, inner=1):
    P = capi.Patch.fromPatchData(pd)
    P.setKernel(capi.KERNEL_TENSOR)
    xi = [PD.gaussLegendre01(n)[0] for n in nq1]
    nq = int(np.prod(nq1))
    npts = P.nelem * nq
    sh = P.shapes(npts)
    o = {f: torch.empty(sh[f], dtype=torch.float64, device="cuda") for f in fields}
    ms = timed(lambda: P.evalTensor(xi, order=order, want=fields, out=o, stream=stream.cuda_stream), inner)
    nb, dim, dw, nh = P.nb, P.dim, P.dimworld, P.nh
    per = dict(N=nb, dN=nb * dim, d2N=nb * nh, x=dw, Jt=dim * dw, H=nh * dw, detJ=1, JinvT=dim * dw)
    bpp = 8 * sum(per[f] for f in fields) + 8 * nb * (dw + 1) / nq
    print("%-28s %8.4f ms  %.3g points/s  %.0f GB/s  %.3f of copy" % (name, ms, npts / ms * 1e3, bpp * npts / ms / 1e6, bpp * npts / ms / 1e6 / PEAK), flush=True)
    P.close()


which = os.environ.get("CFGS", "C1,C4,C5,C5L").split(",")
if "C1" in which:
    tensor_cfg("C1 p=2 256^2 3x3", PD.plateWithHole(7, 8), [3, 3], 1, F1, inner=int(os.environ.get("C1_INNER", "30")))
if "C1big" in which:
    tensor_cfg("C1 refined 1024^2", PD.plateWithHole(9, 10), [3, 3], 1, F1)
if "C4" in which:
    tensor_cfg("C4 shell p=2 1024^2 4x4 o2", PD.torusSurface(8), [4, 4], 2, ("N", "dN", "d2N", "x", "Jt", "H", "detJ"))
if "C5" in which:
    tensor_cfg("C5 p=3 1024^2 4x4", PD.squarePlate(3, 10), [4, 4], 1, F1)
if "C5L" in which:
    pd = PD.squarePlate(3, 9)
    P = capi.Patch.fromPatchData(pd)
    rng = np.random.default_rng(42)
    nel_t = P.nelem // 20
    elems = np.sort(rng.choice(P.nelem, nel_t, replace=False)).astype(np.int32)
    per = rng.integers(3, 41, nel_t) * 6
    elem = np.repeat(elems, per)
    d_elem, d_xi = torch.from_numpy(elem).cuda(), torch.from_numpy(rng.random((len(elem), 2))).cuda()
    sh = P.shapes(len(elem))
    o = {f: torch.empty(sh[f], dtype=torch.float64, device="cuda") for f in F1}
    ms = timed(lambda: P.evalPoints(d_elem, d_xi, order=1, want=F1, out=o, stream=stream.cuda_stream), inner=10)
    bpp = 8 * (16 * 3 + 2 + 4 + 1) + 8 * 3
    print("%-28s %8.4f ms  %.3g points/s  %.0f GB/s  %.3f of copy" % ("C5 trimmed lists", ms, len(elem) / ms * 1e3, bpp * len(elem) / ms / 1e6, bpp * len(elem) / ms / 1e6 / PEAK), flush=True)

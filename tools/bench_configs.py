"""Per-config device-resident timings for the five BASELINE.json configs (context for DESIGN.md; bench.py is the
contract benchmark).  Prints one JSON line per config: points/s (or elements/s), algorithmic GB/s, fraction of the
measured HBM copy bandwidth (or GFLOP/s for assembly)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igab200  # noqa: E402,F401
from dune_iga_b200 import capi, patchdata as PD  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(f, reps=5, warm=2):
    """Median device time of one call in ms.  Short calls are queued back to back between the two events (so that the
    host time of the Python/ctypes call, ~25 us, is hidden behind the previous launch as it is in a real element loop);
    the number of queued calls is chosen so that a sample lasts >= 2 ms."""
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    f()
    e1.record()
    torch.cuda.synchronize()
    inner = max(1, min(200, int(2.0 / max(e0.elapsed_time(e1), 1e-3))))
    ts = []
    for _ in range(reps):
        e0.record()
        for _ in range(inner):
            f()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    return float(np.median(ts))


def eval_config(name, pd, nq1, order, fields, el_chunk=None, kernel=capi.KERNEL_AUTO):
    P = capi.Patch.fromPatchData(pd)
    P.setKernel(kernel)
    xi = [PD.gaussLegendre01(n)[0] for n in nq1]
    nq = int(np.prod(nq1))
    nel = P.nelem if el_chunk is None else min(P.nelem, el_chunk)
    npts = nel * nq
    sh = P.shapes(npts)
    out = {f: torch.empty(sh[f], dtype=torch.float64, device="cuda") for f in fields}
    s = torch.cuda.current_stream().cuda_stream
    ms = timed(lambda: P.evalTensor(xi, order=order, elem_begin=0, elem_end=nel, want=fields, out=out, stream=s))
    nb, dim, dw, nh = P.nb, P.dim, P.dimworld, P.nh
    per = dict(N=nb, dN=nb * dim, d2N=nb * nh, x=dw, Jt=dim * dw, H=nh * dw, detJ=1, JinvT=dim * dw)
    bpp = 8 * sum(per[f] for f in fields) + 8 * nb * (dw + 1) / nq
    gbs = bpp * npts / ms / 1e6
    print(json.dumps({"config": name, "elements": nel, "points": npts, "ms": ms, "points_per_s": npts / ms * 1e3,
                      "bytes_per_point": bpp, "GBs": gbs, "frac_of_copy_peak": gbs / PEAK, "fields": fields}), flush=True)


def main():
    which = sys.argv[1:] or ["C1", "C2", "C3e", "C4", "C5", "C5t", "C3a"]
    if "C1" in which:
        eval_config("C1 2-D plate-with-hole p=2 256x256, 3x3 Gauss, order 1", PD.plateWithHole(7, 8), [3, 3], 1,
                    ("N", "dN", "x", "Jt", "detJ"))
        eval_config("C1x16 same patch refined to 1024x1024 (amortises launch latency)", PD.plateWithHole(9, 10), [3, 3],
                    1, ("N", "dN", "x", "Jt", "detJ"))
    if "C2" in which:
        eval_config("C2 3-D cylinder p=3 128^3, 4^3 Gauss, order 1 (8 k-layers per launch)",
                    PD.thickCylinder(3, (7, 7, 7)), [4, 4, 4], 1, ("N", "dN", "x", "Jt", "detJ"), el_chunk=128 * 128 * 8)
    if "C3e" in which:
        eval_config("C3 evaluation part: 3-D cylinder p=4 64^3, 5^3 Gauss, order 1 (2 k-layers per launch)",
                    PD.thickCylinder(4, (6, 6, 6)), [5, 5, 5], 1, ("N", "dN", "x", "Jt", "detJ"), el_chunk=64 * 64 * 2)
    if "X" in which:  # beyond the five named configs
        eval_config("X1 3-D cylinder p=2 128^3 (16 k-layers), 3^3 Gauss, order 1", PD.thickCylinder(2, (7, 7, 7)), [3, 3, 3], 1,
                    ("N", "dN", "x", "Jt", "detJ"), el_chunk=128 * 128 * 16)
        eval_config("X2 3-D cylinder p=3 64^3 (8 k-layers), 4^3 Gauss, order 2", PD.thickCylinder(3, (6, 6, 6)), [4, 4, 4], 2,
                    ("N", "dN", "d2N", "x", "Jt", "H", "detJ"), el_chunk=64 * 64 * 8)
        eval_config("X3 3-D cylinder p=3 64^3 (8 k-layers), 5^3 Gauss, order 1 (general kernel: rule != p+1)", PD.thickCylinder(3, (6, 6, 6)),
                    [5, 5, 5], 1, ("N", "dN", "x", "Jt", "detJ"), el_chunk=64 * 64 * 8)
        eval_config("X4 2-D plate p=2 1024^2, 3x3 Gauss, detJ only (volume path)", PD.plateWithHole(9, 10), [3, 3], 1, ("detJ",))
    if "W" in which:  # the general kernels: thread per point (generic) against warp per element (warp)
        F1 = ("N", "dN", "x", "Jt", "detJ")
        F2 = ("N", "dN", "d2N", "x", "Jt", "H", "detJ")
        cyl3, cyl4 = PD.thickCylinder(3, (6, 6, 6)), PD.thickCylinder(4, (5, 5, 5))
        mixed = PD.thickCylinder(2).degreeElevate(0, 1)  # coarse solid at (3,2,2): mixed degrees
        for d in range(3):
            mixed = mixed.globalRefineInDirection(d, 6)
        plate4 = PD.plateWithHole(0, 0).degreeElevate(0, 2).degreeElevate(1, 2)
        plate4 = plate4.globalRefineInDirection(0, 8).globalRefineInDirection(1, 9)
        for kn, km in (("generic", capi.KERNEL_GENERIC), ("warp", capi.KERNEL_WARP), ("auto", capi.KERNEL_AUTO)):
            eval_config("W1[%s] 3-D p=3 64^3 (8 k-layers), 4^3 Gauss, order 1" % kn, cyl3, [4, 4, 4], 1, F1,
                        el_chunk=64 * 64 * 8, kernel=km)
            eval_config("W2[%s] 3-D p=3 64^3 (8 k-layers), 6^3 Gauss, order 1 (no tuned kernel)" % kn, cyl3, [6, 6, 6], 1,
                        F1, el_chunk=64 * 64 * 8, kernel=km)
            eval_config("W3[%s] 3-D p=(3,2,2), 4x3x3 Gauss, order 1 (mixed degrees)" % kn, mixed, [4, 3, 3], 1, F1,
                        el_chunk=64 * 64 * 16, kernel=km)
            eval_config("W4[%s] 3-D p=4 32^3 (8 k-layers), 5^3 Gauss, order 2 (no tuned kernel)" % kn, cyl4, [5, 5, 5], 2,
                        F2, el_chunk=32 * 32 * 8, kernel=km)
            eval_config("W5[%s] 2-D plate p=2 elevated to p=4 512^2, 5x5 Gauss, order 1" % kn,
                        plate4, [5, 5], 1, F1, kernel=km)
    if "L" in which:  # explicit point lists in 3-D: thread per point against warp per 32 points
        pd = PD.thickCylinder(3, (5, 5, 5))
        rng = np.random.default_rng(1)
        for label, grouped in (("grouped by element (40 points each)", True), ("scattered", False)):
            P = capi.Patch.fromPatchData(pd)
            npts = 1 << 20
            elem = (np.repeat(rng.choice(P.nelem, npts // 40 + 1, replace=False), 40)[:npts] if grouped
                    else rng.integers(0, P.nelem, npts)).astype(np.int32)
            d_elem, d_xi = torch.from_numpy(elem).cuda(), torch.from_numpy(rng.random((npts, 3))).cuda()
            fields = ("N", "dN", "x", "Jt", "detJ")
            sh = P.shapes(npts)
            out = {f: torch.empty(sh[f], dtype=torch.float64, device="cuda") for f in fields}
            s = torch.cuda.current_stream().cuda_stream
            for kn, km in (("generic", capi.KERNEL_GENERIC), ("warp", capi.KERNEL_WARP)):
                P.setKernel(km)
                ms = timed(lambda: P.evalPoints(d_elem, d_xi, order=1, want=fields, out=out, stream=s))
                bpp = 8 * (64 * 4 + 3 + 9 + 1) + 8 * 4
                print(json.dumps({"config": "L[%s] 3-D p=3 point list, %s" % (kn, label), "points": npts, "ms": ms,
                                  "points_per_s": npts / ms * 1e3, "GBs": bpp * npts / ms / 1e6,
                                  "frac_of_copy_peak": bpp * npts / ms / 1e6 / PEAK}), flush=True)
    if "C4" in which:
        eval_config("C4 torus shell p=2 in R^3 1024x1024, 4x4 Gauss, order 2", PD.torusSurface(8), [4, 4], 2,
                    ("N", "dN", "d2N", "x", "Jt", "H", "detJ"))
    if "C5" in which:
        eval_config("C5 tensor part: 2-D square p=3 1024x1024, 4x4 Gauss, order 1", PD.squarePlate(3, 10), [4, 4], 1,
                    ("N", "dN", "x", "Jt", "detJ"))
    if "C5t" in which:
        pd = PD.squarePlate(3, 9)
        P = capi.Patch.fromPatchData(pd)
        rng = np.random.default_rng(42)
        nel_t = P.nelem // 20  # 5 % trimmed elements
        elems = np.sort(rng.choice(P.nelem, nel_t, replace=False)).astype(np.int32)
        per = rng.integers(3, 41, nel_t) * 6  # 3..40 sub-triangles x 6-point rule
        elem = np.repeat(elems, per)
        xi = rng.random((len(elem), 2))
        d_elem = torch.from_numpy(elem).cuda()
        d_xi = torch.from_numpy(xi).cuda()
        fields = ("N", "dN", "x", "Jt", "detJ")
        sh = P.shapes(len(elem))
        out = {f: torch.empty(sh[f], dtype=torch.float64, device="cuda") for f in fields}
        s = torch.cuda.current_stream().cuda_stream
        ms = timed(lambda: P.evalPoints(d_elem, d_xi, order=1, want=fields, out=out, stream=s))
        bpp = 8 * (16 * 3 + 2 + 4 + 1) + 8 * 3
        print(json.dumps({"config": "C5 trimmed batches: 5 % of 512x512 elements, 18..240 points each (point lists)",
                          "points": len(elem), "ms": ms, "points_per_s": len(elem) / ms * 1e3,
                          "GBs": bpp * len(elem) / ms / 1e6, "frac_of_copy_peak": bpp * len(elem) / ms / 1e6 / PEAK}),
              flush=True)
    if "C3a" in which:
        for kind, kname, s_rows in ((capi.ASM_POISSON, "Poisson", 3), (capi.ASM_ELASTICITY, "elasticity", 6)):
            pd = PD.thickCylinder(4, (5, 5, 5))
            P = capi.Patch.fromPatchData(pd)
            rule = [PD.gaussLegendre01(5)] * 3
            xi, w = [r[0] for r in rule], [r[1] for r in rule]
            n = P.nb * (3 if kind == capi.ASM_ELASTICITY else 1)
            nel = 4736 if kind == capi.ASM_POISSON else 2368  # multiples of 2 x 148 resident blocks
            Ke = torch.empty((nel, n, n), dtype=torch.float64, device="cuda")
            D = None
            if kind == capi.ASM_ELASTICITY:
                lam, mu = 0.3 / (1.3 * 0.4), 1 / 2.6
                D = np.zeros((6, 6))
                D[:3, :3] = lam
                D[np.arange(3), np.arange(3)] = lam + 2 * mu
                D[np.arange(3, 6), np.arange(3, 6)] = mu
            st = torch.cuda.current_stream().cuda_stream
            ms = timed(lambda: P.assemble(kind, xi, w, D=D, elem_end=nel, Ke=Ke, fe=None, stream=st), reps=3, warm=1)
            flops = 125 * (2 * s_rows * n * n + 2 * s_rows * s_rows * n)
            print(json.dumps({"config": "C3 %s element matrices p=4 (n=%d), 5^3 Gauss" % (kname, n), "elements": nel,
                              "ms": ms, "elements_per_s": nel / ms * 1e3, "algorithmic_GFLOPs": flops * nel / ms / 1e6}),
                  flush=True)


if __name__ == "__main__":
    main()

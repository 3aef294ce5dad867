"""Phase accounting of gram_sym_kernel (clock64 deltas of lane 0 of every warp, summed over the launch):
  python tools/sym_phase.py --build   (here: instrumented library -> tools/_phase/libigab200.so)
  python tools/sym_phase.py           (on the GPU box)"""
import ctypes as C
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tools", "_phase", "libigab200.so")
if "--build" in sys.argv:
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = sorted(glob.glob(os.path.join(ROOT, "dune-iga_b200", "csrc", "*.cu")))
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode",
                           "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
                           "-DIGAB_PHASE_TIMING", "-shared", "-o", SO, "-t", "8"] + src + ["-lcudart", "-ldl"], env=env)
    print("built", SO)
    sys.exit(0)
os.environ["IGAB200_SO"] = SO
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import igab200  # noqa: E402,F401
from dune_iga_b200 import capi, patchdata as PD  # noqa: E402

L = capi.lib()
NAMES = ["prepare (tables, control points, geometry, constants)", "operands (T1 zero, B2, B3, M4, fragments, T2 zero)",
         "stage 1", "stage 2", "stage 3 + stores", "mirror copy-out", "barrier waits inside the passes", "load vector"]
for p, q in ((4, 5), (3, 4)):
    pd = PD.thickCylinder(p, (5, 5, 5))
    P = capi.Patch.fromPatchData(pd)
    x, w = PD.gaussLegendre01(q)
    nel = 148 * 32
    for kind, name in ((capi.ASM_POISSON, "poisson"), (capi.ASM_MASS, "mass")):
        K = torch.empty((nel, P.nb, P.nb), dtype=torch.float64, device="cuda")
        f = lambda: P.assemble(kind, [x] * 3, [w] * 3, elem_end=nel, Ke=K, fe=None)
        buf = (C.c_ulonglong * 16)()
        f(); torch.cuda.synchronize()
        L.igab_debug_phase_ticks_sym(buf, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        L.igab_debug_phase_ticks_sym(buf, 1)
        t = np.array(list(buf)[:8], dtype=float) / (8.0 * nel)   # cycles per element per warp
        print("p=%d %s: %.3g element matrices/s (instrumented); cycles per element and warp: total %.0f" % (p, name, nel / e0.elapsed_time(e1) * 1e3, t.sum()))
        for nm, v in zip(NAMES, t):
            print("   %-55s %8.0f  %5.1f %%" % (nm, v, 100 * v / t.sum()))

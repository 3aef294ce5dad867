"""Phase accounting of the fused assembly kernels (clock64 per warp, summed): how much warp time goes to the element
preparation, the chunk builds, the DMMA loop, barrier waits and the epilogue.
  python tools/phase_timing.py --build     (here, no GPU: instrumented library -> tools/_phase/libigab200.so)
  python tools/phase_timing.py             (on the GPU box)"""
import ctypes as C
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tools", "_phase", "libigab200.so")

if "--build" in sys.argv:
    src = sorted(glob.glob(os.path.join(ROOT, "dune-iga_b200", "csrc", "*.cu")))
    cmd = ["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-DIGAB_PHASE_TIMING", "-shared", "-o", SO] + src + ["-lcudart", "-ldl"]
    subprocess.check_call(cmd)
    print("built", SO)
    sys.exit(0)

os.environ["IGAB200_SO"] = SO
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import igab200  # noqa: E402,F401
from dune_iga_b200 import capi, patchdata as PD  # noqa: E402

L = capi.lib()
NAMES = ["prepare (tables, control points, geometry)", "build gradient rows of a chunk", "DMMA loop", "barrier wait", "epilogue (stores / combination)"]


def phases(label, f):
    buf = (C.c_ulonglong * 8)()
    f()
    torch.cuda.synchronize()
    L.igab_debug_phase_ticks(buf, 1)
    f()
    torch.cuda.synchronize()
    L.igab_debug_phase_ticks(buf, 1)
    t = np.array(list(buf)[:5], dtype=float)
    print(label)
    for n, v in zip(NAMES, t / t.sum()):
        print("   %-45s %5.1f %%" % (n, 100 * v))


pd = PD.thickCylinder(4, (5, 5, 5))
P = capi.Patch.fromPatchData(pd)
x, w = PD.gaussLegendre01(5)
nel = 2368
K1 = torch.empty((nel, 125, 125), dtype=torch.float64, device="cuda")
K3 = torch.empty((nel, 375, 375), dtype=torch.float64, device="cuda")
lam, mu = 0.3 / (1.3 * 0.4), 1 / 2.6
D = np.zeros((6, 6)); D[:3, :3] = lam; D[np.arange(3), np.arange(3)] = lam + 2 * mu; D[np.arange(3, 6), np.arange(3, 6)] = mu
phases("gram_fused_kernel<4,5,3> (Poisson p=4)", lambda: P.assemble(capi.ASM_POISSON, [x] * 3, [w] * 3, elem_end=nel, Ke=K1, fe=None))
phases("elasticity_fused_kernel<4,5> (p=4)", lambda: P.assemble(capi.ASM_ELASTICITY, [x] * 3, [w] * 3, D=D, elem_end=nel, Ke=K3, fe=None))

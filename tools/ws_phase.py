"""Phase accounting of gram_ws_kernel (clock64 deltas of one lane per 4 warps, summed over the launch):
  python tools/ws_phase.py --build   (here: instrumented library -> tools/_phase/libigab200.so)
  python tools/ws_phase.py           (on the GPU box)"""
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
                           "-DIGAB_PHASE_TIMING", "-shared", "-o", SO] + src + ["-lcudart", "-ldl"], env=env)
    print("built", SO)
    sys.exit(0)
os.environ["IGAB200_SO"] = SO
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import igab200  # noqa: E402,F401
from dune_iga_b200 import capi, patchdata as PD  # noqa: E402

L = capi.lib()
NAMES = ["mma: wait T1 full", "mma: stage 2", "mma: wait T2 / tile buffer free", "mma: stage 3", "hlp: wait T1 free",
         "hlp: M4 + stage 1", "hlp: wait tile buffer full", "hlp: epilogue", "mma: prepare", "hlp: prepare",
         "mma: operands", "hlp: operands", "hlp: (loop end)"]
pd = PD.thickCylinder(4, (5, 5, 5))
P = capi.Patch.fromPatchData(pd)
x, w = PD.gaussLegendre01(5)
lam, mu = 0.3 / (1.3 * 0.4), 1 / 2.6
D = np.zeros((6, 6)); D[:3, :3] = lam; D[np.arange(3), np.arange(3)] = lam + 2 * mu; D[np.arange(3, 6), np.arange(3, 6)] = mu
for kind, name, n, nel in ((capi.ASM_POISSON, "poisson", 125, 4736), (capi.ASM_ELASTICITY, "elasticity", 375, 2368)):
    K = torch.empty((nel, n, n), dtype=torch.float64, device="cuda")
    f = lambda: P.assemble(kind, [x] * 3, [w] * 3, D=D if kind == 2 else None, elem_end=nel, Ke=K, fe=None)
    buf = (C.c_ulonglong * 16)()
    f(); torch.cuda.synchronize()
    L.igab_debug_phase_ticks(buf, 1)
    f(); torch.cuda.synchronize()
    L.igab_debug_phase_ticks(buf, 1)
    t = np.array(list(buf)[:13], dtype=float)
    # one sampling lane per 4 warps: 1 for the DMMA group, 3 for the helpers
    per_el = t / nel
    print(name, "(cycles per element per sampled warp group)")
    for i, nm in enumerate(NAMES):
        div = 2.0
        print("   %-36s %9.0f" % (nm, per_el[i] / div))

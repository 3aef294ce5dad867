"""Timing experiments on gram_sym_kernel (library built with -DIGAB_SYM_EXPERIMENT into tools/_exp by --build): element matrices/s with
single pieces of the kernel switched off (results are wrong on purpose) -- an upper bound of what optimising a piece can buy."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tools", "_exp", "libigab200.so")
if "--build" in sys.argv:  # here (no GPU): python tools/sym_skip.py --build ; then on the GPU box: python tools/sym_skip.py
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = sorted(glob.glob(os.path.join(ROOT, "dune-iga_b200", "csrc", "*.cu")))
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode",
                           "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
                           "-DIGAB_SYM_EXPERIMENT", "-shared", "-o", SO, "-t", "8"] + src + ["-lcudart", "-ldl"], env=env)
    print("built", SO)
    sys.exit(0)
os.environ["IGAB200_SO"] = SO
os.environ["IGAB_GRAM_SYM"] = "1"
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import igab200  # noqa: E402,F401
from dune_iga_b200 import capi, patchdata as PD  # noqa: E402

NAMES = {0: "all on", 1: "no stage 1", 2: "no B2 build", 4: "no mirror", 8: "no K_e stores", 16: "no stage 2", 32: "no stage-3 DMMA",
         12: "no mirror, no K_e stores", 64: "no element_prepare (tables, control points, geometry, constants)", 64 + 63: "all off incl. element_prepare", 1 + 16 + 32: "no stage 1/2/3 math", 63: "everything off (prepare + operands + barriers left)"}
for p, q, kind, name in ((4, 5, capi.ASM_POISSON, "poisson"), (4, 5, capi.ASM_MASS, "mass")):
    pd = PD.thickCylinder(p, (5, 5, 5))
    P = capi.Patch.fromPatchData(pd)
    x, w = PD.gaussLegendre01(q)
    nel = 148 * 64
    K = torch.empty((nel, P.nb, P.nb), dtype=torch.float64, device="cuda")
    for skip, what in NAMES.items():
        os.environ["IGAB_SYM_SKIP"] = str(skip)
        ts = []
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); P.assemble(kind, [x] * 3, [w] * 3, elem_end=nel, Ke=K, fe=None); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts[1:]))
        print("p=%d %-8s %-52s %8.3f ms  %.3g el/s" % (p, name, what, ms, nel / ms * 1e3), flush=True)

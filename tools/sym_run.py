"""Short run for ncu: gram_sym_kernel on C3 (p=4, 5^3 Gauss), Poisson and mass, 2368 elements each."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["IGAB_GRAM_SYM"] = "1"
import igab200  # noqa
from dune_iga_b200 import capi, patchdata as PD
pd = PD.thickCylinder(4, (5, 5, 5))
P = capi.Patch.fromPatchData(pd)
x, w = PD.gaussLegendre01(5)
nel = 2368
K1 = torch.empty((nel, 125, 125), dtype=torch.float64, device="cuda")
for _ in range(3):
    P.assemble(capi.ASM_POISSON, [x] * 3, [w] * 3, elem_end=nel, Ke=K1, fe=None)
    P.assemble(capi.ASM_MASS, [x] * 3, [w] * 3, elem_end=nel, Ke=K1, fe=None)
torch.cuda.synchronize()
print("ok", float(K1.abs().max()))

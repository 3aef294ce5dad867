"""Short assembly run for ncu: C3 Poisson and elasticity element matrices, p=4, 5^3 Gauss, 2368 elements each."""
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igab200  # noqa
from dune_iga_b200 import capi, patchdata as PD
pd = PD.thickCylinder(4, (5, 5, 5))
P = capi.Patch.fromPatchData(pd)
x, w = PD.gaussLegendre01(5)
nel = 2368
K1 = torch.empty((nel, 125, 125), dtype=torch.float64, device="cuda")
K3 = torch.empty((nel, 375, 375), dtype=torch.float64, device="cuda")
lam, mu = 0.3 / (1.3 * 0.4), 1 / 2.6
D = np.zeros((6, 6)); D[:3, :3] = lam; D[np.arange(3), np.arange(3)] = lam + 2 * mu; D[np.arange(3, 6), np.arange(3, 6)] = mu
for _ in range(3):
    P.assemble(capi.ASM_POISSON, [x] * 3, [w] * 3, elem_end=nel, Ke=K1, fe=None)
    P.assemble(capi.ASM_ELASTICITY, [x] * 3, [w] * 3, D=D, elem_end=nel, Ke=K3, fe=None)
torch.cuda.synchronize()
print("ok", float(K1.abs().max()), float(K3.abs().max()))

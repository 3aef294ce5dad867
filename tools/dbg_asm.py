import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import igab200
import patches_for_tests as PFT
import portbind as PT
from dune_iga_b200 import capi, patchdata as PD
for name in ("cylinder_p3", "cylinder_p4"):
    spec = PFT.small_patch(name)
    Pp = PT.PortPatch(spec["degree"], spec["knots"], spec["cp"], 3)
    P = capi.Patch(spec["degree"], spec["knots"], spec["cp"], 3)
    rule = [PD.gaussLegendre01(p + 1) for p in spec["degree"]]
    xi, w = [r[0] for r in rule], [r[1] for r in rule]
    for kind in (0, 1):
        Kr, fr = Pp.assemble(kind, xi, w, elem_end=8)
        K, f = P.assemble(kind, xi, w, elem_end=8)
        err = np.abs(K - Kr).reshape(8, -1).max(1) / np.abs(Kr).max()
        print(name, kind, "err per element", err, "fe err", np.abs(f - fr).max())
        if err.max() > 1e-10:
            e = int(np.argmax(err)); d = np.abs(K[e] - Kr[e]); i, j = np.unravel_index(np.argmax(d), d.shape)
            print(" worst elem", e, "entry", i, j, K[e, i, j], Kr[e, i, j], "nonzero rows got", (np.abs(K[e]).sum(1) > 0).sum())
            bad = np.argwhere(d > 1e-10 * np.abs(Kr).max()); print(" bad entries", len(bad), bad[:6].tolist(), bad[-3:].tolist())

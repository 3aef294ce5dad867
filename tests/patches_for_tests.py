"""Small seeded patches shared by the CPU and GPU parity tests (built with the product's host generators)."""
import numpy as np

import igab200  # noqa: F401  (registers dune_iga_b200)
from dune_iga_b200 import patchdata as PD


def small_patch(name, level=None):
    if name == "plate_hole_p2":
        P = PD.plateWithHole(2 if level is None else level, 2 if level is None else level)
    elif name == "cylinder_p3":
        lv = 1 if level is None else level
        P = PD.thickCylinder(3, (lv, lv, lv))
    elif name == "cylinder_p4":
        lv = 1 if level is None else level
        P = PD.thickCylinder(4, (lv, lv, lv))
    elif name == "torus_p2":
        P = PD.torusSurface(1 if level is None else level)
    elif name == "square_p3":
        P = PD.squarePlate(3, 2 if level is None else level)
    elif name == "mixed_deg_3d":
        # different degree per direction, non-uniform knots with an interior double knot, random weights
        P = PD.thickCylinder(2, (1, 0, 1))
        P = P.degreeElevate(2, 1)  # degrees (2,2,3)
        P = P.knotRefinement([0.3, 0.3], 1)
        P = PD.perturbed(P, seed=7, amp=0.05)
    elif name == "curve_p3":
        P = PD.NurbsPatchData([[0, 0, 0, 0, 1, 1, 1, 1]], [[-4, -4, 1], [-3, 2.8, 2.5], [2, -4, 1], [4, 4, 1]], [3], 2)
        P = P.globalRefine(2 if level is None else level)
    else:
        raise KeyError(name)
    return dict(degree=P.degree, knots=P.knotSpans, cp=P.cp, dimworld=P.dimworld, patch=P)

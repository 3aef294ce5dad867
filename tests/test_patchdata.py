"""Host-side patch generators (dune-iga_b200/patchdata.py) against the reference's own refinement code
(dune/iga/nurbsalgorithms.hh:49-66,264-528,579-705 through oracle/_ref) and its degree-elevation literals."""
import json
import os

import numpy as np
import pytest

import refbind as RF
import igab200  # noqa: F401
from dune_iga_b200 import patchdata as PD

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_literals.json")))
needs_ref = pytest.mark.skipif(not RF.available(), reason="oracle/_ref not built")


def same_patch(P, R, tol=1e-12):
    assert P.degree == R.degree
    for a, b in zip(P.knotSpans, R.knots):
        assert np.abs(a - b).max() <= 1e-15
    assert P.cp.shape == R.cp.shape
    assert np.abs(P.cp - R.cp).max() <= tol * max(1.0, np.abs(R.cp).max())


def test_degree_elevate_literal():
    c = G["degree_elevate_curve"]
    P = PD.NurbsPatchData(c["knots"], c["cp"], c["degree"], c["dimworld"]).degreeElevate(0, c["elevate_by"])
    assert P.degree == [4]
    assert np.allclose(P.cp[:, :3], c["expected_cp"], rtol=0, atol=6e-10)
    assert np.allclose(P.cp[:, 3], c["expected_w"], rtol=0, atol=12e-10)


@needs_ref
def test_circular_arc_and_revolution_match_reference():
    same_patch(PD.makeCircularArc(1.0), RF.make_circular_arc(1.0))
    same_patch(PD.makeCircularArc(2.5, 0.0, 180.0), RF.make_circular_arc(2.5, 0.0, 180.0))
    arc, rarc = PD.makeCircularArc(1.0), RF.make_circular_arc(1.0)
    same_patch(PD.makeSurfaceOfRevolution(arc, (2, 0, 0), (0, 1, 0), 360.0),
               RF.make_surface_of_revolution(rarc, (2, 0, 0), (0, 1, 0), 360.0))


@needs_ref
def test_refinement_matches_reference():
    # 2-D rational: plate with hole, refine both directions (generateRefinedKnots + knotRefinement)
    P0 = PD.plateWithHole()
    R = RF.RefPatch.create(P0.degree, P0.knotSpans, P0.cp, 2).refine_uniform(0, 3).refine_uniform(1, 2)
    same_patch(P0.globalRefineInDirection(0, 3).globalRefineInDirection(1, 2), R)
    # 3-D: elevate (2,1,1) -> (3,3,3) then refine, i.e. config C2's construction at small size
    Pc = PD.thickCylinder(2, (0, 0, 0))
    Pc = PD.NurbsPatchData(Pc.knotSpans, Pc.cp, Pc.degree, 3)
    coarse = PD.NurbsPatchData([[0, 0, 0, 1, 1, 1], [0, 0, 1, 1], [0, 0, 1, 1]],
                               PD.thickCylinder.__globals__["np"].array(_coarse_cyl()), [2, 1, 1], 3)
    R = RF.RefPatch.create(coarse.degree, coarse.knotSpans, coarse.cp, 3)
    R = R.degree_elevate(0, 1).degree_elevate(1, 2).degree_elevate(2, 2)
    for d in range(3):
        R = R.refine_uniform(d, 2)
    same_patch(PD.thickCylinder(3, (2, 2, 2)), R, tol=1e-12)
    # arbitrary sorted knots incl. one already present
    R2 = RF.RefPatch.create(P0.degree, P0.knotSpans, P0.cp, 2).knot_refine(0, [0.1, 0.5, 0.77])
    same_patch(P0.knotRefinement([0.1, 0.5, 0.77], 0), R2)


def _coarse_cyl(r_i=1.0, r_o=2.0, L=3.0):
    cw = 1.0 / np.sqrt(2.0)
    cp = []
    for z in (0.0, L):
        for r in (r_i, r_o):
            cp += [(r, 0, z, 1.0), (r, r, z, cw), (0, r, z, 1.0)]
    return cp


def test_refinement_keeps_geometry():
    """Refinement / elevation must not move the map (the reference checks this at 1e-10, gridtests.cpp:1177-1200)."""
    import portbind as PT
    P0 = PD.thickCylinder(2, (0, 0, 0))
    P1 = PD.thickCylinder(3, (1, 2, 1))
    a = PT.PortPatch(P0.degree, P0.knotSpans, P0.cp, 3)
    b = PT.PortPatch(P1.degree, P1.knotSpans, P1.cp, 3)
    rng = np.random.default_rng(1)
    for u in rng.random((20, 3)):
        xa, xb = a.geo_global(u), b.geo_global(u)
        assert np.abs(xa - xb).max() < 1e-13
        assert abs(np.hypot(xa[0], xa[1]) - (1.0 + u[1])) < 1e-13  # radius exact: r_i + u_1 (r_o - r_i)
        assert abs(a.geo_detj(u) - b.geo_detj(u)) < 1e-12


def test_validation_errors():
    with pytest.raises(ValueError):
        PD.NurbsPatchData([[0, 0, 1, 1, 1]], [[0, 0, 1], [1, 0, 1]], [2], 2)  # not open
    with pytest.raises(ValueError):
        PD.plateWithHole().knotRefinement([0.5, 0.2], 0)  # unsorted

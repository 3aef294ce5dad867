#!/usr/bin/env python
"""Extract the known-answer literals of the reference's own unit tests into tests/golden/reference_literals.json.

Run in the authoring container only (needs /root/reference).  The JSON it writes is committed; tests never read
/root/reference at run time.  Sources:
  dune/iga/test/gridtests.cpp:677-988        testBsplineBasisFunctions  (1-D B-spline, 2-D NURBS values + derivatives)
  dune/iga/test/gridtests.cpp:1129-1221      testCurveHigherOrderDerivatives (degreeElevate control points)
  dune/iga/test/trimmedgridtests.cpp:38-186  testPatchGeometryCurve / testPatchGeometrySurface
Only numbers and the data they belong to are taken; inputs (knots, control points) are restated by hand below
with the line they come from.
"""
import json
import os
import re

REF = "/root/reference/dune/iga/test"
src = open(os.path.join(REF, "gridtests.cpp")).read()
lines = src.splitlines()
seg = "\n".join(lines[676:988])  # testBsplineBasisFunctions

out = {"_source": "rath3t/dune-iga v0.1.7 unit tests; tolerance there: Dune::FloatCmp::eq default (8 eps relative)"}

# --- 1-D B-spline values  (gridtests.cpp:678-717)
out["bspline1d_values"] = [
    {"cite": "gridtests.cpp:678-686", "knots": [0, 0, 0, 0.5, 0.5, 2, 2, 3, 3, 3], "p": 2, "u": 0.3,
     "N": [0.16, 0.48, 0.36]},
    {"cite": "gridtests.cpp:688-692", "knots": [0, 0, 0, 0.5, 0.5, 2, 2, 3, 3, 3], "p": 2, "u": 0.5,
     "N": [1.0, 0.0, 0.0]},
    {"cite": "gridtests.cpp:694-698", "knots": [0, 0, 0, 0.5, 0.5, 2, 2, 3, 3, 3], "p": 2, "u": 0.0,
     "N": [1.0, 0.0, 0.0]},
    {"cite": "gridtests.cpp:700-704", "knots": [0, 0, 0, 0.5, 0.5, 2, 2, 3, 3, 3], "p": 2, "u": 3.0,
     "N": [0.0, 0.0, 1.0]},
    {"cite": "gridtests.cpp:706-712", "knots": [0, 0, 0, 0.5, 1, 1, 1], "p": 2, "u": 0.1, "N": [0.64, 0.34, 0.02]},
    {"cite": "gridtests.cpp:714-717", "knots": [0, 0, 0, 0.5, 1, 1, 1], "p": 2, "u": 0.01,
     "N": [0.9604, 0.0394, 0.0002]},
]

# --- 1-D p=3 values + derivatives 0..3 at u=1.45 (gridtests.cpp:719-751)
ders = [[None] * 4 for _ in range(4)]
for m in re.finditer(r"eq\(dN\[(\d)\]\[(\d)\],\s*([-0-9.eE]+)\)", seg):
    ders[int(m.group(1))][int(m.group(2))] = float(m.group(3))
vals = [float(m.group(2)) for m in re.finditer(r"eq\(N2\[(\d)\],\s*([-0-9.eE]+)\)", seg)]
assert all(v is not None for r in ders for v in r) and len(vals) == 4
out["bspline1d_ders"] = {"cite": "gridtests.cpp:719-751", "knots": [0, 0, 0, 0, 0.5, 1, 1, 2, 2, 2, 2], "p": 3,
                         "u": 1.45, "N": vals, "dN": ders}

# --- N_0(u) at 10 equispaced u in [0,2] (gridtests.cpp:753-766)
m = re.search(r"NAtEvalPoints\s*=\s*\{([^}]*)\}", seg)
out["bspline1d_N0_sweep"] = {"cite": "gridtests.cpp:753-766", "knots": [0, 0, 0, 0, 0.5, 1, 1, 2, 2, 2, 2], "p": 3,
                             "u": [i / 9.0 * 2.0 for i in range(10)],
                             "N0": [float(x) for x in m.group(1).replace("\n", " ").split(",")]}

# --- 2-D NURBS, p=(2,2) (gridtests.cpp:768-986).  The reference builds the weight net with dimsize {3,7} from a 3x7
# table, value (i,j) = 1 + 7 i + j; only the 3x3 block of the first span is ever read (netOfSpan start (0,0)).
nurbs = {"cite": "gridtests.cpp:768-986",
         "knots": [[0, 0, 0, 0.5, 0.5, 2, 2, 3, 3, 3], [0, 0, 0, 2, 2, 2]], "degree": [2, 2],
         "weights_3x3_block": [[1 + 7 * i + j for j in range(3)] for i in range(3)],
         "weights_note": "w(i,j) for i,j<3 (i = first direction); the rest of a proper 7x3 net is never touched"}
v1 = [None] * 9
first_block = seg[seg.index("xieta{0.2, 0.25}"):seg.index("xieta   = {0, 0.1}")]
for m in re.finditer(r"eq\(N_Nurbs\[(\d)\],\s*([-0-9.eE]+)\)", first_block):
    v1[int(m.group(1))] = float(m.group(2))
second_block = seg[seg.index("xieta   = {0, 0.1}"):seg.index("xieta         = {0, 0.1}")]
v2 = [None] * 9
for m in re.finditer(r"eq\(N_Nurbs\[(\d)\],\s*([-0-9.eE]+)\)", second_block):
    v2[int(m.group(1))] = float(m.group(2))
nurbs["values"] = [{"u": [0.2, 0.25], "R": v1}, {"u": [0.0, 0.1], "R": v2}]
dd = {}
for m in re.finditer(r"eq\(dN_Nurbs\.get\(\{(\d), (\d)\}\)\.get\(\{(\d), (\d)\}\),\s*([-0-9.eE]+)\)", seg):
    a = "%s,%s" % (m.group(1), m.group(2))
    dd.setdefault(a, [None] * 9)[int(m.group(3)) + 3 * int(m.group(4))] = float(m.group(5))
assert all(v is not None for r in dd.values() for v in r), dd
nurbs["derivatives_at_0_0.1"] = dd  # key "a,b" = derivative multi-index; value i-fastest over the 3x3 local net
out["nurbs2d"] = nurbs

# --- curve / surface geometry (trimmedgridtests.cpp:38-186)
out["geometry_curve"] = {
    "cite": "trimmedgridtests.cpp:38-110", "degree": [3], "knots": [[0, 0, 0, 0, 1, 1, 1, 1]], "dimworld": 2,
    "cp": [[-4, -4, 1], [-3, 2.8, 2.5], [2, -4, 1], [4, 4, 1]],
    "global": [[[0.0], [-4, -4]], [[0.5], [-1.32, 0.72]], [[1.0], [4, 4]],
               [[0.25], [-2.7607655502392343, 0.4688995215311005]],
               [[0.4], [-1.9854368932038828, 0.7669902912621357]]],
    "jacobianTransposed": [[[0.0], [[7.5, 51]]], [[0.5], [[7.4496, -0.9216]]]],
    # geometry.local(x) == u (trimmedgridtests.cpp:90-95)
    "local": [[[-4, -4], [0.0]], [[-1.32, 0.72], [0.5]]]}
out["geometry_surface"] = {
    "cite": "trimmedgridtests.cpp:112-186", "degree": [2, 2], "knots": [[0, 0, 0, 1, 1, 1], [0, 0, 0, 1, 1, 1]],
    "dimworld": 3,
    # controlPoints[i][j] sits at multi-index (i,j); stored here i-fastest (flat = i + 3 j)
    "cp_table_ij": [[[0, 0, 1, 1], [1, 0, 1, 1], [2, 0, 2, 1]],
                    [[0, 1, 0, 1], [1, 1, 0, 1], [2, 1, 0, 1]],
                    [[0, 2, 1, 1], [1, 2, 2, 1], [2, 2, 2, 1]]],
    "global": [[[0.5, 0.5], [1.0, 1.0, 0.75]], [[0, 0], [0, 0, 1]], [[0, 1], [2, 0, 2]]],
    "jacobianTransposed": [[[0.5, 0.5], [[0.0, 2.0, 0.5], [2.0, 0.0, 0.5]]]],
    "corners": [[0, 0, 1], [0, 2, 1], [2, 0, 2], [2, 2, 2]],
    # geometry.local(x) == u (trimmedgridtests.cpp:157-162)
    "local": [[[1.0, 1.0, 0.75], [0.5, 0.5]], [[2, 0, 2], [0, 1]]]}

# --- element centres after one global refinement of the bilinear unit square element.ibra (trimmedgridtests.cpp:197-217;
# fixture auxiliaryfiles/element.ibra: degree (1,1), knots {0,1}x{0,1}, poles (0,0),(0,1),(1,0),(1,1), unit weights).
# The reference compares with operator== (exact).  Element index = position, first direction fastest.
out["element_centres"] = {
    "cite": "trimmedgridtests.cpp:197-217", "degree": [1, 1], "knots": [[0, 0, 1, 1], [0, 0, 1, 1]], "dimworld": 2,
    "cp": [[0, 0, 1], [1, 0, 1], [0, 1, 1], [1, 1, 1]], "refine_level": 1,
    "centres": [[0.25, 0.25], [0.75, 0.25], [0.25, 0.75], [0.75, 0.75]]}

# --- degreeElevate by 2 of a p=2 rational curve (gridtests.cpp:1136-1166,1203-1221), tolerance 1e-10
out["degree_elevate_curve"] = {
    "cite": "gridtests.cpp:1136-1166,1203-1221", "tol": 1e-10, "degree": [2], "knots": [[0, 0, 0, 0.5, 1, 1, 1]],
    "dimworld": 3,
    "cp": [[0.086956521739130, -0.434782608695652, 0, 11.5], [0.2, 5.4, 0.2, 5],
           [1.857142857142857, 0.142857142857143, 0, 7], [3.714285714285714, 0.285714285714286, 2.0, 3.5]],
    "elevate_by": 2,
    "expected_cp": [[0.086956521739130, -0.434782608695652, 0], [0.121212121212121, 1.333333333333333, 0.060606060606061],
                    [0.32, 3.12, 0.12], [0.727272727272727, 3.727272727272727, 0.136363636363636],
                    [1.538461538461539, 1.153846153846154, 0.038461538461538], [1.92, 0.506666666666667, 0.2],
                    [2.476190476190476, 0.190476190476190, 0.666666666666667],
                    [3.714285714285714, 0.285714285714286, 2.0]],
    "expected_w": [11.5, 8.25, 6.25, 5.5, 6.5, 6.25, 5.25, 3.5]}

# --- integral invariants (gridtests.cpp:338-363): torus r=1 (arc), R=2 -> area 4 pi^2 r R (1e-4), Gauss-Bonnet 0 (1e-5)
out["torus"] = {"cite": "gridtests.cpp:307-311,338-363", "r": 1.0, "R": 2.0, "area_tol": 1e-4, "gauss_bonnet_tol": 1e-5,
                "refine_levels": 2}

dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_literals.json")
json.dump(out, open(dst, "w"), indent=1)
print("wrote", dst, {k: (len(v) if hasattr(v, "__len__") else v) for k, v in out.items()})

"""Host-side patch data and patch generators (numpy), mirroring the reference's value type and constructors.

Mirrors (names and argument meaning; the algorithms are written from scratch as linear operators on the control net):
  Dune::IGA::NURBSPatchData          dune/iga/nurbspatchdata.hh:17-36   -> NurbsPatchData
  Dune::IGA::ControlPoint{p, w}      dune/iga/controlpoint.hh:15-38     -> rows (x.., w) of NurbsPatchData.cp
  generateRefinedKnots               dune/iga/nurbsalgorithms.hh:49-66  -> generateRefinedKnots
  knotRefinement                     dune/iga/nurbsalgorithms.hh:441-528 -> knotRefinement
  degreeElevate                      dune/iga/nurbsalgorithms.hh:264-439 -> degreeElevate
  makeCircularArc                    dune/iga/nurbsalgorithms.hh:579-633 -> makeCircularArc
  makeSurfaceOfRevolution            dune/iga/nurbsalgorithms.hh:636-705 -> makeSurfaceOfRevolution
  Python registrations               dune/python/iga/nurbspatchdata.hh:67-98, python/dune/iga/_nurbsalgorithms.py

This is the step *before* the hot path: it runs once per refinement on the host and only has to
produce the same patch as the reference to rounding (checked against the reference in tests/test_patchdata.py);
the evaluation kernels take whatever control net they are given.

Layout contract: control net flat, FIRST DIRECTION FASTEST (MultiDimensionNet, dune/iga/utils/mdnet.hh:329-345),
one row (x_0..x_{dimworld-1}, w) per control point, coordinates NOT pre-multiplied by w.
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np

FLOATCMP_EPS = 8 * np.finfo(float).eps  # Dune::FloatCmp default (relativeWeak)


def _feq(a, b):
    return abs(a - b) <= FLOATCMP_EPS * max(abs(a), abs(b))


def uniqueKnots(U):
    """unique_copy with FloatCmp::eq, as dune/iga/nurbspatch.hh:58-63."""
    out = []
    for v in np.asarray(U, float):
        if not out or not _feq(out[-1], v):
            out.append(float(v))
    return np.array(out)


@dataclass
class NurbsPatchData:
    knotSpans: List[np.ndarray]
    cp: np.ndarray  # [prod(ncp), dimworld+1]
    degree: List[int]
    dimworld: int = field(default=None)

    def __post_init__(self):
        self.knotSpans = [np.ascontiguousarray(k, dtype=np.float64) for k in self.knotSpans]
        self.degree = [int(p) for p in self.degree]
        self.cp = np.ascontiguousarray(self.cp, dtype=np.float64)
        if self.dimworld is None:
            self.dimworld = self.cp.shape[-1] - 1
        self.cp = self.cp.reshape(-1, self.dimworld + 1)
        for d in range(self.patchDim):
            U, p = self.knotSpans[d], self.degree[d]
            # open knot vector asserted by the reference (bsplinealgorithms.hh:98-99, nurbsgrid.hh:112-122)
            if len(U) < 2 * (p + 1) or np.any(np.diff(U) < 0):
                raise ValueError("knot vector %d is not a sorted open knot vector of degree %d" % (d, p))
            if not (np.all(U[:p + 1] == U[0]) and np.all(U[-p - 1:] == U[-1])):
                raise ValueError("knot vector %d must repeat its end knots degree+1 times" % d)
        if self.cp.shape[0] != int(np.prod(self.ncp)):
            raise ValueError("control net has %d points, knots/degree imply %s" % (self.cp.shape[0], self.ncp))

    # ---- sizes
    @property
    def patchDim(self):
        return len(self.degree)

    @property
    def ncp(self):
        return [len(self.knotSpans[d]) - self.degree[d] - 1 for d in range(self.patchDim)]

    @property
    def elementsPerDirection(self):
        return [len(uniqueKnots(U)) - 1 for U in self.knotSpans]

    @property
    def numElements(self):
        return int(np.prod(self.elementsPerDirection))

    @property
    def nb(self):
        return int(np.prod([p + 1 for p in self.degree]))

    def net(self):
        """Control net as an array [n_{dim-1}, ..., n_0, dimworld+1] (C order == first direction fastest)."""
        return self.cp.reshape(list(reversed(self.ncp)) + [self.dimworld + 1])

    def copy(self):
        return NurbsPatchData([k.copy() for k in self.knotSpans], self.cp.copy(), list(self.degree), self.dimworld)

    # ---- reference-named operations
    def knotRefinement(self, newKnots, refinementDirection):
        return knotRefinement(self, newKnots, refinementDirection)

    def degreeElevate(self, refinementDirection, elevationFactor):
        return degreeElevate(self, refinementDirection, elevationFactor)

    def globalRefine(self, level):
        """NURBSGrid::globalRefine: every direction (dune/iga/nurbsgrid.hh:133-141)."""
        out = self
        for d in range(self.patchDim):
            out = out.globalRefineInDirection(d, level)
        return out

    def globalRefineInDirection(self, direction, level):
        """dune/iga/nurbsgrid.hh:143-149."""
        if level == 0:
            return self
        return knotRefinement(self, generateRefinedKnots(self.knotSpans, direction, level), direction)


# ------------------------------------------------------------------------------------------------ B-spline basics
def findSpan(p, u, U):
    """bsplinealgorithms.hh:24-32 semantics (host helper for the generators only)."""
    U = np.asarray(U)
    if u <= U[0]:
        return p
    if u >= U[-1]:
        return len(U) - p - 2
    return int(np.searchsorted(U, u, side="right") - 1)


def _all_basis(U, p, u):
    """Row of all n = len(U)-p-1 B-spline values at u (Cox-de Boor), right end included in the last span."""
    n = len(U) - p - 1
    s = findSpan(p, u, U)
    N = np.zeros(p + 1)
    N[0] = 1.0
    left, right = np.zeros(p + 1), np.zeros(p + 1)
    for j in range(1, p + 1):
        left[j] = u - U[s + 1 - j]
        right[j] = U[s + j] - u
        saved = 0.0
        for r in range(j):
            t = N[r] / (right[r + 1] + left[j - r])
            N[r] = saved + right[r + 1] * t
            saved = left[j - r] * t
        N[j] = saved
    row = np.zeros(n)
    row[s - p:s + 1] = N
    return row


def generateRefinedKnots(knotSpans, direction, refinementLevel):
    """2^level - 1 equispaced new knots inside every non-empty span (nurbsalgorithms.hh:49-66)."""
    U = np.asarray(knotSpans[direction], float)
    uq = U[np.concatenate([[True], np.diff(U) != 0])]  # std::ranges::unique (exact ==)
    m = 2 ** refinementLevel
    out = []
    for i in range(len(uq) - 1):
        inc = (uq[i + 1] - uq[i]) / m
        out.extend(uq[i] + inc * j for j in range(1, m))
    return np.array(out)


def knotInsertionOperator(U, p, newKnots):
    """(U_new, T) with T [n_new, n_old]: homogeneous control points of the refined curve = T @ old ones.
    Built by repeated single-knot (Boehm) insertion; exact affine combinations."""
    U = np.asarray(U, float).copy()
    n = len(U) - p - 1
    T = np.eye(n)
    for ub in np.sort(np.asarray(newKnots, float)):
        k = int(np.searchsorted(U, ub, side="right") - 1)  # U[k] <= ub < U[k+1]
        n = len(U) - p - 1
        A = np.zeros((n + 1, n))
        for i in range(n + 1):
            if i <= k - p:
                A[i, i] = 1.0
            elif i <= k:
                a = (ub - U[i]) / (U[i + p] - U[i])
                A[i, i] = a
                A[i, i - 1] = 1.0 - a
            else:
                A[i, i - 1] = 1.0
        T = A @ T
        U = np.insert(U, k + 1, ub)
    return U, T


def degreeElevationOperator(U, p, t):
    """(U_new, E): raise degree p -> p+t keeping the curve; every distinct knot gains multiplicity t.
    E solves the collocation system  B_new(g) E = B_old(g)  at the Greville abscissae g of the new basis
    (non-singular by Schoenberg-Whitney)."""
    U = np.asarray(U, float)
    uq, mult = np.unique(U, return_counts=True)
    Un = np.repeat(uq, mult + t)
    q = p + t
    nn = len(Un) - q - 1
    g = np.array([Un[i + 1:i + q + 1].mean() for i in range(nn)])
    Bn = np.array([_all_basis(Un, q, x) for x in g])
    Bo = np.array([_all_basis(U, p, x) for x in g])
    E = np.linalg.solve(Bn, Bo)
    E[np.abs(E) < 1e-15] = 0.0
    return Un, E


def _apply_along(patch, direction, M, newU, newDegree=None):
    """Apply the 1-D operator M to the homogeneous control net along `direction`."""
    dim, dw = patch.patchDim, patch.dimworld
    net = patch.net().copy()
    net[..., :dw] *= net[..., dw:dw + 1]  # (w x, w)
    axis = dim - 1 - direction
    out = np.moveaxis(np.tensordot(M, np.moveaxis(net, axis, 0), axes=(1, 0)), 0, axis)
    out = np.ascontiguousarray(out)
    out[..., :dw] /= out[..., dw:dw + 1]
    knots = [k.copy() for k in patch.knotSpans]
    knots[direction] = newU
    deg = list(patch.degree)
    if newDegree is not None:
        deg[direction] = newDegree
    return NurbsPatchData(knots, out.reshape(-1, dw + 1), deg, dw)


def knotRefinement(oldData, newKnots, refinementDirection):
    newKnots = np.asarray(newKnots, float)
    if np.any(np.diff(newKnots) < 0):
        raise ValueError("You passed a non-sorted vector of new knots.")  # assert at nurbsalgorithms.hh:444-445
    if len(newKnots) == 0:
        return oldData.copy()
    U, T = knotInsertionOperator(oldData.knotSpans[refinementDirection], oldData.degree[refinementDirection], newKnots)
    return _apply_along(oldData, refinementDirection, T, U)


def degreeElevate(oldData, refinementDirection, elevationFactor):
    if elevationFactor == 0:
        return oldData.copy()
    p = oldData.degree[refinementDirection]
    U, E = degreeElevationOperator(oldData.knotSpans[refinementDirection], p, elevationFactor)
    return _apply_along(oldData, refinementDirection, E, U, p + elevationFactor)


# ------------------------------------------------------------------------------------------------ constructors
def makeCircularArc(radius=1.0, startAngle=0.0, endAngle=360.0, origin=(0, 0, 0), X=(1, 0, 0), Y=(0, 1, 0)):
    """Quadratic rational arc in R^3 made of <= 90 degree pieces (NURBS book A7.1; nurbsalgorithms.hh:579-633).
    Angles in degrees for the sweep; like the reference, the *start direction* uses startAngle as given."""
    origin, X, Y = (np.asarray(v, float) for v in (origin, X, Y))
    if endAngle < startAngle:
        endAngle += 360.0
    theta = endAngle - startAngle
    narcs = int(np.ceil(theta / 90.0))
    dth = theta / narcs * np.pi / 180.0
    w1 = np.cos(dth / 2.0)
    P = np.zeros((2 * narcs + 1, 4))
    P[:, 3] = 1.0
    P0 = origin + radius * np.cos(startAngle) * X + radius * np.sin(startAngle) * Y
    T0 = -np.sin(startAngle) * X + np.cos(startAngle) * Y
    P[0, :3] = P0
    ang = startAngle
    idx = 0
    for i in range(1, narcs + 1):
        ang += dth
        P2 = origin + radius * np.cos(ang) * X + radius * np.sin(ang) * Y
        T2 = -np.sin(ang) * X + np.cos(ang) * Y
        P[idx + 2, :3] = P2
        P[idx + 1, :3] = _intersect_lines(P0, T0, P2, T2)
        P[idx + 1, 3] = w1
        idx += 2
        P0, T0 = P2, T2
    U = np.zeros(2 * (narcs + 2))
    for i in range(1, narcs):
        U[2 * i + 1] = U[2 * i + 2] = i / narcs
    U[-3:] = 1.0
    return NurbsPatchData([U], P, [2], 3)


def _intersect_lines(p1, d1, p2, d2):
    """Closest point on line 1 to line 2 (they intersect for tangent lines of one circle)."""
    b = p1 - p2
    a11, a12, a22 = d1 @ d1, d1 @ d2, d2 @ d2
    D = a11 * a22 - a12 * a12
    if abs(D) < 1e-8:
        raise ValueError("The two lines are almost parallel.")
    s = (a12 * (d2 @ b) - a22 * (d1 @ b)) / D
    return p1 + s * d1


def makeSurfaceOfRevolution(generatrix, point, revolutionaxis, revolutionAngle=360.0):
    """NURBS book A8.1 (nurbsalgorithms.hh:636-705): direction 0 = angular (p=2), direction 1 = the generatrix."""
    if generatrix.patchDim != 1 or generatrix.dimworld != 3:
        raise ValueError("The passed nurbs patchdata has to be a curve in 3D.")
    point = np.asarray(point, float)
    axis = np.asarray(revolutionaxis, float)
    axis = axis / np.linalg.norm(axis)
    narcs = int(np.ceil(revolutionAngle / 90.0))
    U = np.zeros(2 * (narcs + 2))
    for i in range(1, narcs):
        U[2 * i + 1] = U[2 * i + 2] = i / narcs
    U[-3:] = 1.0
    dth = revolutionAngle / narcs * np.pi / 180.0
    wm = np.cos(dth / 2.0)
    ang = dth * np.arange(1, narcs + 1)
    cosines, sines = np.cos(ang), np.sin(ang)
    gcp = generatrix.cp
    m = gcp.shape[0]
    n0 = 2 * narcs + 1
    S = np.zeros((m, n0, 4))  # [j][i]
    for j in range(m):
        Pj, wj = gcp[j, :3], gcp[j, 3]
        Om = point + ((Pj - point) @ axis) * axis  # projection onto the axis
        Xv = Pj - Om
        r = np.linalg.norm(Xv)
        Xv = Xv / r
        Yv = np.cross(axis, Xv)
        S[j, 0, :3], S[j, 0, 3] = Pj, wj
        P0, T0 = Pj, Yv
        idx = 0
        for i in range(narcs):
            P2 = Om + r * cosines[i] * Xv + r * sines[i] * Yv
            T2 = -sines[i] * Xv + cosines[i] * Yv
            S[j, idx + 2, :3], S[j, idx + 2, 3] = P2, wj
            S[j, idx + 1, :3] = _intersect_lines(P0, T0, P2, T2)
            S[j, idx + 1, 3] = wm * wj
            idx += 2
            P0, T0 = P2, T2
    return NurbsPatchData([U, generatrix.knotSpans[0].copy()], S.reshape(-1, 4), [2, generatrix.degree[0]], 3)


# ------------------------------------------------------------------------------------------------ benchmark patches
def gaussLegendre01(n):
    """n-point Gauss-Legendre rule on [0,1] (the C-ABI takes the rule as input; dune-geometry's is not vendored)."""
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def plateWithHole(level_xi=0, level_eta=0):
    """C1: classical quarter plate with circular hole, p=(2,2), R=1, L=4, 2x1 coarse elements."""
    s = (1.0 + 1.0 / np.sqrt(2.0)) / 2.0
    rows = [
        [(-1, 0, 1), (-1, np.sqrt(2) - 1, s), (1 - np.sqrt(2), 1, s), (0, 1, 1)],
        [(-2.5, 0, 1), (-2.5, 0.75, 1), (-0.75, 2.5, 1), (0, 2.5, 1)],
        [(-4, 0, 1), (-4, 4, 1), (-4, 4, 1), (0, 4, 1)],
    ]
    cp = np.array([c for row in rows for c in row], float)
    P = NurbsPatchData([[0, 0, 0, .5, 1, 1, 1], [0, 0, 0, 1, 1, 1]], cp, [2, 2], 2)
    return P.globalRefineInDirection(0, level_xi).globalRefineInDirection(1, level_eta)


def thickCylinder(p=3, levels=(0, 0, 0), r_i=1.0, r_o=2.0, L=3.0):
    """C2/C3: thick-walled quarter cylinder: coarse degrees (2,1,1), elevated to (p,p,p), then refined 2^level per
    direction.  Direction 0 angular, 1 radial, 2 axial."""
    cw = 1.0 / np.sqrt(2.0)
    cp = []
    for z in (0.0, L):
        for r in (r_i, r_o):
            cp += [(r, 0, z, 1.0), (r, r, z, cw), (0, r, z, 1.0)]
    P = NurbsPatchData([[0, 0, 0, 1, 1, 1], [0, 0, 1, 1], [0, 0, 1, 1]], np.array(cp, float), [2, 1, 1], 3)
    for d, base in enumerate((2, 1, 1)):
        if p > base:
            P = P.degreeElevate(d, p - base)
    for d in range(3):
        P = P.globalRefineInDirection(d, levels[d])
    return P


def torusSurface(level=0, r=1.0, R=2.0):
    """C4: torus surface p=(2,2) in R^3, as gridtests.cpp:307-311 / readgrid.py:101-107 build it:
    makeSurfaceOfRevolution(makeCircularArc(r) shifted, {R,0,0}, {0,1,0}, 360).  4x4 coarse elements with double
    interior knots."""
    arc = makeCircularArc(r)
    # the reference test moves only control point 0 (gridtests.cpp:309) for gridCheck; the area test uses the plain arc
    surf = makeSurfaceOfRevolution(arc, (R, 0, 0), (0, 1, 0), 360.0)
    return surf.globalRefine(level)


def squarePlate(p=3, level=0, L=1.0):
    """C5 tensor part: flat square p=(p,p) with mildly non-uniform weights."""
    P = NurbsPatchData([[0, 0, 1, 1], [0, 0, 1, 1]],
                       np.array([(0, 0, 1.0), (L, 0, 0.8), (0, L, 1.2), (L, L, 1.0)], float), [1, 1], 2)
    for d in range(2):
        P = P.degreeElevate(d, p - 1)
    return P.globalRefine(level)


def perturbed(P, seed=42, amp=0.1, wlo=0.5, whi=1.5):
    """Robustness variant used by the survey probe: control points + U(-amp,amp)*h, weights U(wlo,whi)."""
    rng = np.random.default_rng(seed)
    Q = P.copy()
    ext = Q.cp[:, :Q.dimworld].max(0) - Q.cp[:, :Q.dimworld].min(0)
    h = ext.max() / max(Q.ncp)
    Q.cp[:, :Q.dimworld] += rng.uniform(-amp, amp, (Q.cp.shape[0], Q.dimworld)) * h
    Q.cp[:, Q.dimworld] = rng.uniform(wlo, whi, Q.cp.shape[0])
    return Q

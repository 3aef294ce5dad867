"""Quadrature rules handed to the batched entry points (the C-ABI takes rules as INPUT: dune-geometry's tabulated
rules are not vendored with the reference).

Mirrors:  NURBSGridEntity<0>::fillQuadratureRule / getQuadratureRule   dune/iga/nurbsgridentity.hh:120-141
          fillQuadratureRuleImpl (trimmed elements)                   dune/iga/utils/fillquadraturerule.hh:8-45
"""
import numpy as np


def gaussLegendre01(n):
    """n-point Gauss-Legendre rule on [0,1]."""
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def pointsForOrder(order):
    """Number of Gauss points per direction dune-geometry uses for polynomial order `order`: floor(order/2) + 1."""
    return order // 2 + 1


def defaultRule(degree):
    """Untrimmed element: QuadratureRules<double,dim>::rule(cube, dim * max(p))  (nurbsgridentity.hh:123-124,131-132).
    Returns (xi1d, w1d) lists per direction."""
    dim = len(degree)
    n = pointsForOrder(dim * max(degree))
    x, w = gaussLegendre01(n)
    return [x.copy() for _ in range(dim)], [w.copy() for _ in range(dim)]


# symmetric simplex rules on the reference triangle {(0,0),(1,0),(0,1)}, weights sum to 1/2
_TRI = {
    1: (np.array([[1 / 3, 1 / 3]]), np.array([0.5])),
    2: (np.array([[1 / 6, 1 / 6], [2 / 3, 1 / 6], [1 / 6, 2 / 3]]), np.array([1 / 6, 1 / 6, 1 / 6])),
    4: (np.array([[0.445948490915965, 0.445948490915965], [0.445948490915965, 0.108103018168070],
                  [0.108103018168070, 0.445948490915965], [0.091576213509771, 0.091576213509771],
                  [0.091576213509771, 0.816847572980459], [0.816847572980459, 0.091576213509771]]),
        np.array([0.223381589678011, 0.223381589678011, 0.223381589678011, 0.109951743655322, 0.109951743655322,
                  0.109951743655322]) * 0.5),
}


def triangleRule(order):
    """A symmetric rule exact for polynomials of degree <= order (order <= 4): 1, 3 or 6 points."""
    for o in sorted(_TRI):
        if order <= o:
            return _TRI[o]
    raise ValueError("triangle rules up to order 4 are tabulated")


# CLOSED rules on the reference triangle (nodes on vertices / edges, so neighbouring sub-triangles share points): what
# QuadratureType::GaussLobatto stands for in fillQuadratureRuleImpl.  dune-geometry's own Lobatto points are not
# vendored with the reference (parity of the point positions unpinned); the MERGE semantics below are the
# reference's.  Weights sum to 1/2.
_TRI_CLOSED = {
    1: (np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]), np.array([1 / 6, 1 / 6, 1 / 6])),
    2: (np.array([[0.5, 0.0], [0.5, 0.5], [0.0, 0.5]]), np.array([1 / 6, 1 / 6, 1 / 6])),
    3: (np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [0.5, 0.0], [0.5, 0.5], [0.0, 0.5], [1 / 3, 1 / 3]]),
        np.array([1 / 40, 1 / 40, 1 / 40, 1 / 15, 1 / 15, 1 / 15, 9 / 40])),
}


def closedTriangleRule(order):
    """A closed rule exact for polynomials of degree <= order (order <= 3): vertex / edge-midpoint / centroid nodes."""
    for o in sorted(_TRI_CLOSED):
        if order <= o:
            return _TRI_CLOSED[o]
    raise ValueError("closed triangle rules up to order 3 are tabulated")


def mergeDuplicatePoints(xi, w):
    """Gauss-Lobatto branch of fillQuadratureRuleImpl (fillquadraturerule.hh:21-44): points with EXACTLY equal
    coordinates are merged, their weights added in order of appearance, and the rule is returned sorted
    lexicographically (first coordinate, then second) -- the std::map with the Comparator of :23-33."""
    xi = np.asarray(xi, float).reshape(-1, 2)
    w = np.asarray(w, float)
    if len(w) == 0:
        return xi.copy(), w.copy()
    order = np.lexsort((xi[:, 1], xi[:, 0]))          # stable: equal keys keep their order of appearance
    xs, ws = xi[order], w[order]
    new = np.ones(len(ws), bool)
    new[1:] = np.any(xs[1:] != xs[:-1], axis=1)
    starts = np.flatnonzero(new)
    out_w = np.empty(len(starts))
    ends = np.append(starts[1:], len(ws))
    for k, (a, b) in enumerate(zip(starts, ends)):   # sequential sums: the reference adds one weight at a time
        acc = ws[a]
        for j in range(a + 1, b):
            acc += ws[j]
        out_w[k] = acc
    return xs[starts].copy(), out_w


def fillQuadratureRule(triangles, order=4, qt="GaussLegendre"):
    """Rule of a TRIMMED element from its sub-triangles (element-local [0,1]^2), fillquadraturerule.hh:13-19:
    position = triangle.global(ip.position()), weight = ip.weight() * triangle.integrationElement(ip.position()).
    qt = "GaussLobatto": closed sub-triangle rules, duplicate points merged and the rule sorted (:21-44).
    triangles: [ntri, 3, 2].  Returns (xi [npts, 2], w [npts])."""
    tri = np.asarray(triangles, float).reshape(-1, 3, 2)
    if qt == "GaussLobatto":
        rp, rw = closedTriangleRule(order)
        a = tri[:, 1] - tri[:, 0]
        b = tri[:, 2] - tri[:, 0]
        ie = np.abs(a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0])
        # vertex / midpoint nodes are evaluated so that a point shared by two triangles gets bit-identical coordinates
        # from both: barycentric combination with the vertices in a canonical (sorted) order
        pts = np.empty((len(tri), len(rw), 2))
        for t in range(len(tri)):
            for q, (l1, l2) in enumerate(rp):
                lam = np.array([1.0 - l1 - l2, l1, l2])
                idx = np.lexsort((tri[t][:, 1], tri[t][:, 0]))
                acc = np.zeros(2)
                for k in idx:
                    if lam[k] != 0.0:
                        acc = acc + lam[k] * tri[t][k]
                pts[t, q] = acc
        return mergeDuplicatePoints(pts.reshape(-1, 2), (rw[None, :] * ie[:, None]).reshape(-1))
    rp, rw = triangleRule(order)
    a = tri[:, 1] - tri[:, 0]
    b = tri[:, 2] - tri[:, 0]
    ie = np.abs(a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0])
    xi = tri[:, None, 0, :] + rp[None, :, 0, None] * a[:, None, :] + rp[None, :, 1, None] * b[:, None, :]
    w = rw[None, :] * ie[:, None]
    return xi.reshape(-1, 2), w.reshape(-1)


def trimmedBatch(elem_triangles, order=4, qt="GaussLegendre"):
    """CSR-style batch for igab_eval_points from {element index: triangles}: (elem [npts], xi [npts,2], w [npts],
    offsets [nel+1]) -- the per-element point lists config C5 feeds to the evaluation kernel."""
    elems, xis, ws, off = [], [], [], [0]
    for e in sorted(elem_triangles):
        xi, w = fillQuadratureRule(elem_triangles[e], order, qt)
        elems.append(np.full(len(w), e, dtype=np.int32))
        xis.append(xi)
        ws.append(w)
        off.append(off[-1] + len(w))
    if not elems:
        return np.zeros(0, np.int32), np.zeros((0, 2)), np.zeros(0), np.array(off)
    return np.concatenate(elems), np.concatenate(xis), np.concatenate(ws), np.array(off)

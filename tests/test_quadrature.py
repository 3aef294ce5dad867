"""Host-side rules: Gauss points, default order (nurbsgridentity.hh:120-133), trimmed-element rule assembly
(utils/fillquadraturerule.hh:8-19) against the oracle restatement and the reference's own property
'sum of weights == triangulated area' (trimmedgridtests.cpp:395-413)."""
import numpy as np

import igab200  # noqa: F401
import portbind as PT
from dune_iga_b200 import quadrature as Q


def test_gauss_matches_oracle_and_integrates():
    for n in range(1, 9):
        x, w = Q.gaussLegendre01(n)
        xo, wo = PT.gauss01(n)
        assert np.abs(x - xo).max() < 1e-15 and np.abs(w - wo).max() < 1e-15
        for k in range(2 * n):
            assert abs((w * x ** k).sum() - 1.0 / (k + 1)) < 1e-14


def test_default_rule_orders():
    # order = dim * max(p): 3, 5, 7 points per direction for C1 (p=2, 2-D), C2 (p=3, 3-D), C3 (p=4, 3-D)
    assert len(Q.defaultRule([2, 2])[0][0]) == 3
    assert len(Q.defaultRule([3, 3, 3])[0][0]) == 5
    assert len(Q.defaultRule([4, 4, 4])[0][0]) == 7


def test_trimmed_rule_matches_oracle_and_area():
    rng = np.random.default_rng(42)
    tri = rng.random((17, 3, 2))
    for order in (1, 2, 4):
        rp, rw = Q.triangleRule(order)
        xi, w = Q.fillQuadratureRule(tri, order)
        xo, wo = PT.trimmed_rule(tri, rp, rw)
        assert np.abs(xi - xo).max() < 1e-15 and np.abs(w - wo).max() < 1e-15
        a, b = tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
        area = 0.5 * np.abs(a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]).sum()
        assert abs(w.sum() - area) < 1e-14  # trimmedgridtests.cpp:395-413
        # exactness: integrate x^a y^b over one triangle against the closed form a! b! / (a+b+2)! on the reference
    rp, rw = Q.triangleRule(4)
    from math import factorial as f
    for a in range(5):
        for b in range(5 - a):
            assert abs((rw * rp[:, 0] ** a * rp[:, 1] ** b).sum() - f(a) * f(b) / f(a + b + 2)) < 1e-14
    el, xi, w, off = Q.trimmedBatch({7: tri[:3], 2: tri[3:5]}, 2)
    assert list(el[:6]) == [2] * 6 and off.tolist() == [0, 6, 15] and len(w) == 15


def test_gauss_lobatto_merge_matches_oracle_and_reference_semantics():
    """Gauss-Lobatto branch of fillQuadratureRuleImpl (fillquadraturerule.hh:21-44): duplicates merged by exact
    coordinate equality, weights added, output sorted by (x, y) -- host code against the oracle's std::map restatement."""
    rng = np.random.default_rng(11)
    xi = rng.integers(0, 4, size=(200, 2)) / 4.0          # many exact duplicates
    w = rng.random(200)
    mx, mw = Q.mergeDuplicatePoints(xi, w)
    ox, ow = PT.merge_duplicate_points(xi, w)
    assert np.array_equal(mx, ox) and np.array_equal(mw, ow)   # same order of additions: identical bits
    assert len(mw) == len({tuple(p) for p in xi}) and abs(mw.sum() - w.sum()) < 1e-13
    assert all(tuple(mx[k]) < tuple(mx[k + 1]) for k in range(len(mx) - 1))
    # a fan of sub-triangles around a common vertex: closed rules share vertices and edge midpoints
    c = np.array([0.5, 0.5])
    ring = np.array([[0, 0], [1, 0], [1, 1], [0, 1.0]])
    tri = np.array([[c, ring[k], ring[(k + 1) % 4]] for k in range(4)])
    for order, npts in ((1, 5), (2, 8), (3, 17)):
        xi, w = Q.fillQuadratureRule(tri, order, qt="GaussLobatto")
        assert len(w) == npts and abs(w.sum() - 1.0) < 1e-15
        for a in range(order + 1):
            for b in range(order + 1 - a):
                assert abs((w * xi[:, 0] ** a * xi[:, 1] ** b).sum() - 1.0 / ((a + 1) * (b + 1))) < 1e-14
    # Gauss-Legendre rules are left alone (no merge, no sort): fillquadraturerule.hh:21-22
    x0, w0 = Q.fillQuadratureRule(tri, 2)
    assert len(w0) == 12
    el, xi, w, off = Q.trimmedBatch({3: tri}, 3, qt="GaussLobatto")
    assert off.tolist() == [0, 17] and set(el) == {3}

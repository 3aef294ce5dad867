"""Shared helpers of the parity tests: tolerance definition and oracle selection."""
import numpy as np

import portbind as PT
import refbind as RF

# north_star: "within 1e-12 relative error in FP64 for basis values, derivatives, Jacobians and element matrices".
# Relative to the largest entry of the same quantity AT THE SAME POINT (individual second derivatives cancel heavily,
# SURVEY.md 7 "hard parts"), floored at 1e-3 of the field's global maximum so exact zeros do not divide by zero.
TOL = 1e-12


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    n = a.shape[0]
    d = np.abs(a - b).reshape(n, -1).max(1)
    s = np.abs(b).reshape(n, -1).max(1)
    floor = max(1e-3 * np.abs(b).max(), 1e-300)
    return float((d / np.maximum(s, floor)).max())


def oracle_for(spec):
    """Prefer the reference itself (oracle/_ref); fall back to the pinned C restatement."""
    if RF.available():
        return RF.RefPatch.create(spec["degree"], spec["knots"], spec["cp"], spec["dimworld"]), "reference"
    return PT.PortPatch(spec["degree"], spec["knots"], spec["cp"], spec["dimworld"]), "port"


def port_for(spec):
    return PT.PortPatch(spec["degree"], spec["knots"], spec["cp"], spec["dimworld"])

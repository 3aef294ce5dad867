"""Element partitioning across ranks: contiguous knot-span blocks.

Elements are independent on the hot path (no state between elements in dune/iga/nurbsgeometry.hh / nurbsbasis.hh), so
each rank owns a contiguous range of the x-fastest element index (dune/iga/utils/mdnet.hh:245-257): whole layers of the
slowest direction.  Knots and control net are replicated; evaluation and element matrices need no collective.  Only
patch-level scalars (volume, norms) are all-reduced, and rows of the assembled global matrix that belong to DOFs touched
from both sides of a cut (p layers of n_0 x n_1 basis functions, from g_d = s_d - p_d + l_d, nurbsbasis.hh:576) are
exchanged with the neighbour.
"""
import numpy as np


def slab_range(nel_dir, world, rank):
    """[elem_begin, elem_end) of `rank`: ceil/floor split of the layers of the slowest direction."""
    nel_dir = [int(n) for n in nel_dir]
    layers = nel_dir[-1]
    per_layer = int(np.prod(nel_dir[:-1])) if len(nel_dir) > 1 else 1
    base, rem = divmod(layers, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo * per_layer, hi * per_layer


def weighted_ranges(counts, world):
    """Trimmed patches: contiguous element ranges with (nearly) equal cumulative point counts.
    counts[e] = quadrature points of element e (0 for empty elements).  Returns world+1 boundaries."""
    counts = np.asarray(counts, dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(counts)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(cum, total * r / world, side="left")))
    bounds.append(len(counts))
    return [min(max(b, bounds[i - 1] if i else 0), len(counts)) for i, b in enumerate(bounds)]


def interface_dofs(dofmap, ranges):
    """For each cut between consecutive ranges: the sorted global DOFs touched by elements on both sides."""
    out = []
    for r in range(len(ranges) - 2):
        a = np.unique(dofmap[ranges[r]:ranges[r + 1]])
        b = np.unique(dofmap[ranges[r + 1]:ranges[r + 2]])
        both = np.intersect1d(a[a >= 0], b[b >= 0])
        out.append(both)
    return out

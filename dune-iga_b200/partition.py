"""Element partitioning across ranks: contiguous knot-span blocks.

Elements are independent on the hot path (no state between elements in dune/iga/nurbsgeometry.hh / nurbsbasis.hh), so
each rank owns a contiguous range of the x-fastest element index (dune/iga/utils/mdnet.hh:245-257): whole layers of the
slowest direction.  Knots and control net are replicated; evaluation and element matrices need no collective.  Only
patch-level scalars (volume, norms) are all-reduced, and rows of the assembled global matrix that belong to DOFs touched
by elements of several ranks (from g_d = s_d - p_d + l_d, nurbsbasis.hh:576) are exchanged.

The partition, the row ranges and the exchange itself live behind the C-ABI (igab_partition_slab,
igab_interface_row_ranges, igab_exchange_interface_rows in csrc/comm.cu: ncclSend / ncclRecv + an ordered add kernel);
this module is the thin Python face of those calls plus the point-count balancing of trimmed patches.
"""
import numpy as np

from . import capi


def slab_range(nel_dir, world, rank):
    """[elem_begin, elem_end) of `rank`: ceil/floor split of the layers of the slowest direction."""
    return capi.partitionSlab(nel_dir, world, rank)


def slab_bounds(nel_dir, world):
    """world + 1 element bounds of the slab partition."""
    return [slab_range(nel_dir, world, r)[0] for r in range(world)] + [int(np.prod([int(n) for n in nel_dir]))]


def weighted_ranges(counts, world, per_layer=None):
    """Trimmed patches: contiguous element ranges with (nearly) equal cumulative point counts.
    counts[e] = quadrature points of element e (0 for empty elements).  Returns world+1 boundaries; with per_layer the
    boundaries are rounded to whole layers of the slowest direction (what the interface-row exchange needs)."""
    return capi.partitionWeighted(counts, world, per_layer or 1)


def interface_dofs(dofmap, ranges):
    """For each cut between consecutive ranges: the sorted global DOFs touched by elements on both sides."""
    out = []
    for r in range(len(ranges) - 2):
        a = np.unique(dofmap[ranges[r]:ranges[r + 1]])
        b = np.unique(dofmap[ranges[r + 1]:ranges[r + 2]])
        both = np.intersect1d(a[a >= 0], b[b >= 0])
        out.append(both)
    return out


def touched_row_ranges(pd, world, ncomp=1, elem_bounds=None):
    """(row_begin, row_end) [world]: the contiguous range of global rows (untrimmed numbering g*ncomp + c) that the
    elements of every rank's slab touch; two ranks share the intersection of their ranges."""
    return capi.interfaceRowRanges(pd.degree, pd.knotSpans, world, ncomp, elem_bounds)


def interface_row_ranges(pd, world, ncomp=1):
    """Per cut c (between rank c and c+1): the rows [row_begin, row_end) both sides touch -- the intersection of the two
    neighbours' ranges (p layers of n_0 x n_1 functions for simple knots)."""
    rb, re = touched_row_ranges(pd, world, ncomp)
    return [(int(max(rb[c], rb[c + 1])), int(max(min(re[c], re[c + 1]), max(rb[c], rb[c + 1])))) for c in range(world - 1)]


def exchange_interface_rows(P, comm, values=None, b=None, ncomp=1, elem_bounds=None, stream=None):
    """In place on every rank: add the other ranks' partial sums on the shared rows (igab_exchange_interface_rows)."""
    return P.exchangeInterfaceRows(comm, values=values, b=b, ncomp=ncomp, elem_bounds=elem_bounds, stream=stream)

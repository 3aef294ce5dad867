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


# ---- interface rows of the assembled global matrix (north star: "NCCL ... for the interface rows") ---------------------
def element_first_dofs(U, p):
    """First global DOF index (per direction) of every element: span - p with span = findSpanCorrected(p, uq[e], U, e)
    (dune/iga/nurbsbasis.hh:355-360,576)."""
    from .patchdata import uniqueKnots
    U = np.asarray(U, float)
    uq = uniqueKnots(U)
    spans = np.searchsorted(U, uq[:-1], side="right") - 1
    return spans - p


def interface_row_ranges(pd, world, ncomp=1):
    """For the slab partition of `pd` over `world` ranks: per cut c (between rank c and c+1) the contiguous global row
    range [row_begin, row_end) of the DOF layers that elements on BOTH sides touch (p layers of n_0 x n_1 functions
    for simple knots).  Rows are numbered g*ncomp + comp with g first-direction-fastest."""
    nel = pd.elementsPerDirection
    last = pd.patchDim - 1
    first = element_first_dofs(pd.knotSpans[last], pd.degree[last])
    per_layer = int(np.prod(pd.ncp[:last])) if last > 0 else 1
    cuts = []
    for r in range(world - 1):
        eb, ee = slab_range(nel, world, r)
        per_el_layer = int(np.prod(nel[:last])) if last > 0 else 1
        l_end = ee // per_el_layer  # first element layer of rank r+1
        if l_end <= 0 or l_end >= nel[last]:
            cuts.append((0, 0))
            continue
        lo = int(first[l_end])                           # lowest DOF layer the right side touches
        hi = int(first[l_end - 1]) + pd.degree[last]     # highest DOF layer the left side touches
        cuts.append((lo * per_layer * ncomp, (hi + 1) * per_layer * ncomp))
    return cuts


def exchange_interface_rows(values, rowptr, cuts, rank, world, dist):
    """In place: add the neighbour's partial sums on the interface rows (both sides end up with the complete rows).
    `values` is a torch tensor (CUDA with the nccl backend, CPU with gloo); one send/recv pair per cut."""
    import torch
    ops, bufs = [], []
    for c, (r0, r1) in enumerate(cuts):
        if r1 <= r0 or rank not in (c, c + 1):
            continue
        peer = c + 1 if rank == c else c
        a, b = int(rowptr[r0]), int(rowptr[r1])
        recv = torch.empty(b - a, dtype=values.dtype, device=values.device)
        send = values[a:b].clone()
        ops += [dist.P2POp(dist.isend, send, peer), dist.P2POp(dist.irecv, recv, peer)]
        bufs.append((a, b, recv, send))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for a, b, recv, _ in bufs:
        values[a:b] += recv
    return values

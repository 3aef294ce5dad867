"""N>1 host logic on CPU: world_size-2 gloo processes partition a patch into knot-span slabs, each evaluates its slab
(with the oracle standing in for the device on this GPU-less box), and the patch-level reduction is all-reduced."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import igab200  # noqa: F401
    import patches_for_tests as PFT
    import portbind as PT
    from dune_iga_b200 import partition, patchdata as PD
    spec = PFT.small_patch("cylinder_p3", level=2)
    P = PT.PortPatch(spec["degree"], spec["knots"], spec["cp"], 3)
    nel_dir = spec["patch"].elementsPerDirection
    eb, ee = partition.slab_range(nel_dir, world, rank)
    x, w = PD.gaussLegendre01(4)
    o = P.eval_tensor([x, x, x], order=1, elem_begin=eb, elem_end=ee, want=("detJ",))
    ww = np.einsum("k,j,i->kji", w, w, w).reshape(-1)
    part = torch.tensor([(o["detJ"].reshape(-1, 64) * ww).sum()], dtype=torch.float64)
    dist.all_reduce(part)
    total = P.volume([x, x, x], [w, w, w])
    dof, n = P.dofmap()
    cuts = partition.interface_dofs(dof, [partition.slab_range(nel_dir, world, r)[0] for r in range(world)] + [P.nelem])
    q.put((rank, eb, ee, float(part.item()), total, [len(c) for c in cuts], spec["patch"].ncp))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_slab_partition_and_reduction():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, b0, e0, v0, t0, c0, ncp), (r1, b1, e1, v1, t1, c1, _) = res
    assert b0 == 0 and e0 == b1 and e1 == 64  # 4x4x4 elements, two slabs of 2 layers
    assert v0 == v1  # all-reduced value identical on both ranks
    assert abs(v0 - t0) <= 1e-12 * abs(t0)
    # interface: p = 3 layers of n0 x n1 basis functions straddle the cut
    assert c0 == [3 * ncp[0] * ncp[1]]


def test_partition_helpers():
    import igab200  # noqa: F401
    from dune_iga_b200 import partition
    for world in (1, 2, 3, 4, 8):
        rs = [partition.slab_range([5, 7, 13], world, r) for r in range(world)]
        assert rs[0][0] == 0 and rs[-1][1] == 5 * 7 * 13
        assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
        assert all((b - a) % 35 == 0 for a, b in rs)
        sizes = [(b - a) // 35 for a, b in rs]
        assert max(sizes) - min(sizes) <= 1
    counts = np.array([16] * 50 + [0] * 20 + [120] * 10 + [16] * 20)
    b = partition.weighted_ranges(counts, 4)
    assert b[0] == 0 and b[-1] == 100 and all(b[i] <= b[i + 1] for i in range(4))
    loads = [counts[b[i]:b[i + 1]].sum() for i in range(4)]
    assert max(loads) - min(loads) <= 2 * 120


def _gloo_exchange(vals, rowptr, rb, re, rank, world):
    """The algorithm of igab_exchange_interface_rows (csrc/comm.cu) with gloo standing in for ncclSend / ncclRecv: every
    pair of ranks with intersecting row ranges swaps the value range of the shared rows, contributions are added in
    ascending rank order."""
    peers = []
    for s in range(world):
        r0, r1 = max(rb[s], rb[rank]), min(re[s], re[rank])
        if s != rank and r1 > r0:
            peers.append((s, int(rowptr[r0]), int(rowptr[r1])))
    own = vals.clone()
    ops, recv = [], {}
    for s, a, z in peers:
        recv[s] = torch.empty(z - a, dtype=vals.dtype)
        ops += [dist.P2POp(dist.isend, own[a:z].clone(), s), dist.P2POp(dist.irecv, recv[s], s)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    acc = torch.zeros_like(vals)
    seen = torch.zeros(len(vals), dtype=torch.bool)
    for s in sorted([q[0] for q in peers] + [rank]):
        if s == rank:
            a, z = int(rowptr[rb[rank]]), int(rowptr[re[rank]])
            acc[a:z] = torch.where(seen[a:z], acc[a:z] + own[a:z], own[a:z])
        else:
            a, z = next((q[1], q[2]) for q in peers if q[0] == s)
            acc[a:z] = torch.where(seen[a:z], acc[a:z] + recv[s], recv[s])
        seen[a:z] = True
    a, z = int(rowptr[rb[rank]]), int(rowptr[re[rank]])
    vals[a:z] = acc[a:z]
    return vals


def _csr_worker(rank, world, port, q, name="plate_hole_p2", level=2):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import igab200  # noqa: F401
    import patches_for_tests as PFT
    import portbind as PT
    from dune_iga_b200 import partition, patchdata as PD
    spec = PFT.small_patch(name, level=level)
    pd = spec["patch"]
    P = PT.PortPatch(pd.degree, pd.knotSpans, pd.cp, spec["dimworld"])
    x, w = PD.gaussLegendre01(max(pd.degree) + 1)
    Ke, _ = P.assemble(0, [x] * pd.patchDim, [w] * pd.patchDim)
    dof, n = P.dofmap()
    # dense stand-in for the device CSR gather on this GPU-less box: full band pattern, row-major
    rowptr = np.arange(n + 1) * n
    eb, ee = partition.slab_range(pd.elementsPerDirection, world, rank)
    mine = np.zeros((n, n))
    for e in range(eb, ee):
        np.add.at(mine, (dof[e][:, None], dof[e][None, :]), Ke[e])
    full = np.zeros((n, n))
    for e in range(P.nelem):
        np.add.at(full, (dof[e][:, None], dof[e][None, :]), Ke[e])
    cuts = partition.interface_row_ranges(pd, world)
    rb, re = partition.touched_row_ranges(pd, world)
    vals = torch.from_numpy(mine.reshape(-1).copy())
    _gloo_exchange(vals, rowptr, rb, re, rank, world)
    got = vals.numpy().reshape(n, n)
    # rows of the DOF layers this rank touches are now complete
    touched = np.unique(dof[eb:ee])
    assert touched.min() == rb[rank] and touched.max() == re[rank] - 1  # the planner's range is exactly what is touched
    err = np.abs(got[touched] - full[touched]).max() / np.abs(full).max()
    # every rank holding a shared row holds the same bits (ascending-rank summation)
    allv = [None] * world
    dist.all_gather_object(allv, got)
    same = True
    for s in range(world):
        r0, r1 = max(rb[s], rb[rank]), min(re[s], re[rank])
        if r1 > r0:
            same = same and np.array_equal(allv[s][r0:r1], got[r0:r1])
    q.put((rank, cuts, float(err), pd.ncp, pd.degree, bool(same), [int(v) for v in rb], [int(v) for v in re]))
    dist.barrier()
    dist.destroy_process_group()


def test_interface_rows_exchange_two_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_csr_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, cuts, err, ncp, deg, same, rb, re in res:
        assert err < 1e-14 and same
        (r0, r1), = cuts
        assert (r1 - r0) == deg[1] * ncp[0]  # p layers of n_0 basis functions


def test_interface_rows_exchange_thin_slabs_three_and_four_ranks():
    """Slabs thinner than the degree: a DOF layer is shared by more than two ranks and the outer ranks must receive each
    other's partial sums too (4 element layers, p = 3 in the slowest direction, 3 and 4 ranks)."""
    for world in (3, 4):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = 33500 + (os.getpid() % 2000) + world
        procs = [ctx.Process(target=_csr_worker, args=(r, world, port, q, "cylinder_p3", 2)) for r in range(world)]
        for p in procs:
            p.start()
        res = sorted(q.get(timeout=300) for _ in range(world))
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        for rank, cuts, err, ncp, deg, same, rb, re in res:
            assert err < 1e-14 and same, (world, rank, err, same)
        # rank 0 and rank 2 are not neighbours but share DOF layers
        rb, re = res[0][6], res[0][7]
        assert min(re[0], re[2]) > max(rb[0], rb[2])

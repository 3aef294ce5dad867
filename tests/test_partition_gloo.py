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


def _csr_worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import igab200  # noqa: F401
    import patches_for_tests as PFT
    import portbind as PT
    from dune_iga_b200 import partition, patchdata as PD
    spec = PFT.small_patch("plate_hole_p2", level=2)
    pd = spec["patch"]
    P = PT.PortPatch(pd.degree, pd.knotSpans, pd.cp, 2)
    x, w = PD.gaussLegendre01(3)
    Ke, _ = P.assemble(0, [x, x], [w, w])
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
    vals = torch.from_numpy(mine.reshape(-1).copy())
    partition.exchange_interface_rows(vals, rowptr, cuts, rank, world, dist)
    got = vals.numpy().reshape(n, n)
    # rows of the DOF layers this rank touches are now complete
    touched = np.unique(dof[eb:ee])
    err = np.abs(got[touched] - full[touched]).max() / np.abs(full).max()
    q.put((rank, cuts, float(err), pd.ncp, pd.degree))
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
    for rank, cuts, err, ncp, deg in res:
        assert err < 1e-14
        (r0, r1), = cuts
        assert (r1 - r0) == deg[1] * ncp[0]  # p layers of n_0 basis functions

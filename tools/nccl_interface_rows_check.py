"""Multi-GPU check (torchrun, >= 2 GPUs): every rank assembles the element matrices of its knot-span slab on its GPU,
gathers them into the global CSR pattern, completes the shared rows with igab_exchange_interface_rows (ncclSend /
ncclRecv + ordered add behind the C-ABI, communicator from igab_comm_create) and compares the rows it touches with a
single-GPU assembly of the whole patch.  IGAB_CHECK_LAYERS_LOG2 sets the number of element layers of the slowest
direction (default 4 -> 16 layers; 3 -> 8 layers = slabs thinner than the degree at 4+ ranks)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igab200  # noqa: E402,F401
from dune_iga_b200 import capi, partition, patchdata as PD  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = capi.Communicator.fromTorch(dist, device="cuda")
    assert comm.info() == (rank, world)
    worst = 0.0
    for lev in ([int(os.environ["IGAB_CHECK_LAYERS_LOG2"])] if "IGAB_CHECK_LAYERS_LOG2" in os.environ else [4, 2]):
        pd = PD.thickCylinder(3, (3, 3, lev))  # 8 x 8 x 2^lev elements, p = 3
        P = capi.Patch.fromPatchData(pd)
        x, w = PD.gaussLegendre01(4)
        xi, ww = [x] * 3, [w] * 3
        n = P.nb
        eb, ee = partition.slab_range(pd.elementsPerDirection, world, rank)
        Ke = torch.empty((max(ee - eb, 1), n, n), dtype=torch.float64, device="cuda")
        fe = torch.empty((max(ee - eb, 1), n), dtype=torch.float64, device="cuda")
        P.assemble(capi.ASM_POISSON, xi, ww, elem_begin=eb, elem_end=ee, Ke=Ke, fe=fe)
        rowptr, colidx = P.csrPattern(1)
        vals = P.csrAssemble(Ke, 1, eb, ee)
        rhs = P.rhsAssemble(fe, 1, eb, ee)
        P.exchangeInterfaceRows(comm, values=vals, b=rhs)
        # the same share from the one-call global assembler restricted to this rank's slab (identical bits), then exchanged
        vals1, rhs1 = P.globalAssemble(capi.ASM_POISSON, xi, ww, device=True, elem_begin=eb, elem_end=ee)
        P.exchangeInterfaceRows(comm, values=vals1, b=rhs1)
        assert torch.equal(vals1, vals), "igab_global_assemble over a slab differs from assemble + csr_assemble"
        assert torch.equal(rhs1, rhs)
        # host arrays take the staged path of the same call
        vh, bh = P.globalAssemble(capi.ASM_POISSON, xi, ww, elem_begin=eb, elem_end=ee)
        P.exchangeInterfaceRows(comm, values=vh, b=bh)
        assert np.array_equal(vh, vals.cpu().numpy()) and np.array_equal(bh, rhs.cpu().numpy())
        # single-GPU reference of the whole matrix
        full, bfull = P.globalAssemble(capi.ASM_POISSON, xi, ww, device=True)
        dof, _ = P.dofmap()
        touched = np.unique(dof[eb:ee])
        mask = np.zeros(len(rowptr) - 1, bool)
        mask[touched] = True
        sel = torch.from_numpy(np.repeat(mask, np.diff(rowptr))).cuda()
        err = float((vals[sel] - full[sel]).abs().max() / full.abs().max())
        mt = torch.from_numpy(mask).cuda()
        errb = float((rhs[mt] - bfull[mt]).abs().max() / bfull.abs().max())
        # bit-identical on every rank that holds a shared row
        rb, re = partition.touched_row_ranges(pd, world)
        mine = vals.cpu().numpy()
        allv = [None] * world
        dist.all_gather_object(allv, mine)
        same = True
        for s in range(world):
            r0, r1 = max(rb[s], rb[rank]), min(re[s], re[rank])
            if r1 > r0:
                same = same and np.array_equal(allv[s][rowptr[r0]:rowptr[r1]], mine[rowptr[r0]:rowptr[r1]])
        res = [None] * world
        dist.all_gather_object(res, (rank, eb, ee, err, errb, bool(same), int(sel.sum())))
        if rank == 0:
            print("layers", 2 ** lev, res)
            assert all(r[3] < 1e-14 and r[4] < 1e-14 and r[5] for r in res), res
        worst = max(worst, err, errb)
        P.close()
    # ---- vector-valued (plane elasticity, ncomp = 2) on a TRIMMED 2-D patch: compacted numbering (prepareForTrim), rows of
    #      DOFs only empty elements touch drop out of every range; 16 x 16 elements p = 2, a hole of empty elements across
    #      the cuts, a few trimmed elements treated with the tensor rule (the exchange does not care where K_e comes from)
    pd2 = PD.squarePlate(2, 4)
    nel = 256
    flags = np.zeros(nel, np.uint8)
    ex, ey = np.meshgrid(np.arange(16), np.arange(16), indexing="xy")
    hole = ((ex - 7.5) ** 2 + (ey - 7.5) ** 2 < 9.0).reshape(-1)
    flags[hole] = 1
    flags[[3, 40, 200]] = 2
    P2 = capi.Patch.fromPatchData(pd2, trimflags=flags)
    x3, w3 = PD.gaussLegendre01(3)
    D3 = np.array([[1.1, 0.3, 0.0], [0.3, 1.1, 0.0], [0.0, 0.0, 0.4]])
    eb, ee = partition.slab_range(pd2.elementsPerDirection, world, rank)
    vals, rhs = P2.globalAssemble(capi.ASM_ELASTICITY, [x3] * 2, [w3] * 2, D=D3, device=True, elem_begin=eb, elem_end=ee)
    P2.exchangeInterfaceRows(comm, values=vals, b=rhs, ncomp=2)
    full, bfull = P2.globalAssemble(capi.ASM_ELASTICITY, [x3] * 2, [w3] * 2, D=D3, device=True)
    rowptr2, _ = P2.csrPattern(2)
    dof2, ndof2 = P2.dofmap()
    mine = dof2[eb:ee][flags[eb:ee] != 1]
    touched = np.unique(mine[mine >= 0])
    mask = np.zeros(ndof2 * 2, bool)
    mask[touched * 2] = True
    mask[touched * 2 + 1] = True
    sel = torch.from_numpy(np.repeat(mask, np.diff(rowptr2))).cuda()
    err2 = float((vals[sel] - full[sel]).abs().max() / full.abs().max()) if int(sel.sum()) else 0.0
    mt = torch.from_numpy(mask).cuda()
    errb2 = float((rhs[mt] - bfull[mt]).abs().max() / bfull.abs().max()) if int(mt.sum()) else 0.0
    res = [None] * world
    dist.all_gather_object(res, (rank, eb, ee, err2, errb2, int(mask.sum()), int(ndof2)))
    if rank == 0:
        print("trimmed, ncomp 2:", res)
        assert all(r[3] < 1e-14 and r[4] < 1e-14 for r in res), res
        assert ndof2 < 18 * 18
    worst = max(worst, err2, errb2)
    P2.close()
    if rank == 0:
        print("interface rows check ok; worst %.2e over %d ranks" % (worst, world))
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

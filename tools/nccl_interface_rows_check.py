"""Multi-GPU check (torchrun, >= 2 GPUs): every rank assembles the element matrices of its knot-span slab on its GPU,
gathers them into the global CSR pattern (igab_csr_assemble), exchanges the interface rows with its neighbours over
NCCL (send/recv of contiguous value ranges) and compares the rows it touches with a single-GPU assembly of the patch."""
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
    pd = PD.thickCylinder(3, (3, 3, 4))  # 8 x 8 x 16 elements, p = 3
    P = capi.Patch.fromPatchData(pd)
    x, w = PD.gaussLegendre01(4)
    xi, ww = [x] * 3, [w] * 3
    n = P.nb
    eb, ee = partition.slab_range(pd.elementsPerDirection, world, rank)
    Ke = torch.empty((ee - eb, n, n), dtype=torch.float64, device="cuda")
    P.assemble(capi.ASM_POISSON, xi, ww, elem_begin=eb, elem_end=ee, Ke=Ke, fe=None)
    rowptr, colidx = P.csrPattern(1)
    vals = P.csrAssemble(Ke, 1, eb, ee)
    cuts = partition.interface_row_ranges(pd, world)
    partition.exchange_interface_rows(vals, rowptr, cuts, rank, world, dist)
    # the same share from the one-call global assembler restricted to this rank's slab (identical bits), then exchanged
    vals1, _ = P.globalAssemble(capi.ASM_POISSON, xi, ww, want_b=False, device=True, elem_begin=eb, elem_end=ee)
    partition.exchange_interface_rows(vals1, rowptr, cuts, rank, world, dist)
    assert torch.equal(vals1, vals), "igab_global_assemble over a slab differs from assemble + csr_assemble"
    # single-GPU reference of the whole matrix
    Kall = torch.empty((P.nelem, n, n), dtype=torch.float64, device="cuda")
    P.assemble(capi.ASM_POISSON, xi, ww, Ke=Kall, fe=None)
    full = P.csrAssemble(Kall, 1)
    dof, _ = P.dofmap()
    touched = np.unique(dof[eb:ee])
    mask = np.zeros(len(rowptr) - 1, bool)
    mask[touched] = True
    sel = torch.from_numpy(np.repeat(mask, np.diff(rowptr))).cuda()
    err = float((vals[sel] - full[sel]).abs().max() / full.abs().max())
    sym = float(full.abs().max())
    res = [None] * world
    dist.all_gather_object(res, (rank, eb, ee, cuts, err, int(sel.sum())))
    if rank == 0:
        print(res)
        assert all(r[4] < 1e-14 for r in res), res
        print("interface rows check ok; |A|max = %.3e, nnz = %d" % (sym, len(colidx)))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

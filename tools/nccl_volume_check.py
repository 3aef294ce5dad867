"""Multi-GPU check of the patch reduction fused with its NCCL all-reduce (run under torchrun on >= 2 GPUs):
each rank reduces the volume of its element slab on its own GPU, igab_reduce_volume all-reduces the scalar through the
communicator the C-ABI created (igab_comm_unique_id on rank 0, the 128-byte id broadcast through torch.distributed,
igab_comm_create on every rank)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import igab200  # noqa: E402,F401
from dune_iga_b200 import capi, partition, patchdata as PD  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = capi.Communicator.fromTorch(dist, device="cuda")  # igab_comm_unique_id / igab_comm_create
    assert comm.info() == (rank, world)

    pd = PD.thickCylinder(3, (4, 4, 5))
    P = capi.Patch.fromPatchData(pd)
    eb, ee = partition.slab_range(pd.elementsPerDirection, world, rank)
    x, w = PD.gaussLegendre01(4)
    mine = P.volume([x] * 3, [w] * 3, eb, ee)
    total = P.volume([x] * 3, [w] * 3, eb, ee, nccl_comm=comm)
    whole = P.volume([x] * 3, [w] * 3)
    exact = np.pi / 4 * (2.0 ** 2 - 1.0 ** 2) * 3.0
    allv = [None] * world
    dist.all_gather_object(allv, (mine, total))
    if rank == 0:
        print("ranks:", allv, "whole:", whole, "exact:", exact)
        assert all(abs(v[1] - whole) <= 1e-12 * whole for v in allv)
        assert len({v[1] for v in allv}) == 1
        assert abs(sum(v[0] for v in allv) - whole) <= 1e-12 * whole
        assert abs(whole - exact) < 1e-9
        print("nccl volume check ok")
    # field norms: slab partial sums all-reduced inside igab_reduce_norms vs the whole-patch call
    u = np.random.default_rng(7).normal(size=P.dofmap()[1])
    mine_n = P.norms([x] * 3, [w] * 3, u, elem_begin=eb, elem_end=ee)
    total_n = P.norms([x] * 3, [w] * 3, u, elem_begin=eb, elem_end=ee, nccl_comm=comm)
    whole_n = P.norms([x] * 3, [w] * 3, u)
    alln = [None] * world
    dist.all_gather_object(alln, (mine_n.tolist(), total_n.tolist()))
    if rank == 0:
        print("norms ranks:", alln, "whole:", whole_n)
        assert all(np.abs(np.array(v[1]) - whole_n).max() <= 1e-12 * whole_n.max() for v in alln)
        assert len({tuple(v[1]) for v in alln}) == 1
        assert np.abs(sum(np.array(v[0]) for v in alln) - whole_n).max() <= 1e-12 * whole_n.max()
        print("nccl norms check ok")
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

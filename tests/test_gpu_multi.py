"""Multi-GPU checks (-m gpu; skipped on a single-GPU box): the NCCL all-reduce fused into the patch reductions and the
interface-row exchange of the sharded global matrix, each run as two ranks under torchrun on one node."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs on this box")


def _torchrun(script, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", script)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600)


def test_single_rank_communicator_through_the_c_abi():
    """The product's own communicator path on a one-GPU box: igab_comm_unique_id / igab_comm_create with one rank, the
    reductions all-reduce through it (ncclAllReduce on the call's stream) and the interface-row exchange finds no peer."""
    import sys
    sys.path.insert(0, ROOT)
    import numpy as np
    import igab200  # noqa: F401
    from dune_iga_b200 import capi, patchdata as PD
    comm = capi.Communicator(0, 1, lambda raw: raw)
    assert comm.info() == (0, 1)
    pd = PD.thickCylinder(3, (3, 3, 3))
    P = capi.Patch.fromPatchData(pd)
    x, w = PD.gaussLegendre01(4)
    v0 = P.volume([x] * 3, [w] * 3)
    v1 = P.volume([x] * 3, [w] * 3, nccl_comm=comm)
    assert v0 == v1 and abs(v0 - np.pi / 4 * 3.0 * 3.0) < 1e-9
    u = np.random.default_rng(1).normal(size=P.dofmap()[1])
    assert np.array_equal(P.norms([x] * 3, [w] * 3, u), P.norms([x] * 3, [w] * 3, u, nccl_comm=comm))
    vals, b = P.globalAssemble(capi.ASM_POISSON, [x] * 3, [w] * 3)
    keep = vals.copy()
    P.exchangeInterfaceRows(comm, values=vals, b=b)
    assert np.array_equal(keep, vals)
    comm.close()


@needs2
def test_reductions_all_reduce_over_nccl():
    r = _torchrun("nccl_volume_check.py", 29611)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "nccl volume check ok" in r.stdout and "nccl norms check ok" in r.stdout


@needs2
def test_interface_rows_exchange_over_nccl():
    r = _torchrun("nccl_interface_rows_check.py", 29612)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "interface rows check ok" in r.stdout


@needs2
def test_two_devices_in_one_process_large_shared_memory_gather():
    """ADVICE (round 1): per-kernel attributes (dynamic shared memory above 48 KB) are per DEVICE.  Two handles on two
    devices in one process, p=4 elasticity (the CSR gather needs 70 KB per block, the assembly kernel 213 KB): both must
    launch, and give the same bits."""
    import sys
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import igab200  # noqa: F401
    from dune_iga_b200 import capi, patchdata as PD
    pd = PD.thickCylinder(4, (1, 1, 1))
    x, w = PD.gaussLegendre01(5)
    D = np.eye(6) + 0.1
    res = []
    for dev in (1, 0):  # start with the device that is NOT the process default
        torch.cuda.set_device(dev)
        P = capi.Patch.fromPatchData(pd)
        vals, b = P.globalAssemble(capi.ASM_ELASTICITY, [x] * 3, [w] * 3, D=D)
        res.append((vals, b))
        P.close()
    torch.cuda.set_device(0)
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert np.isfinite(res[0][0]).all() and np.abs(res[0][0]).max() > 0

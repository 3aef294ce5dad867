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

"""Two-GPU run of the slab-decomposed solver (NCCL send/recv halo exchange + max-allreduce):
must be BIT-identical to the single-GPU run.  Needs >= 2 GPUs (gpurun --gpus 2); on a
one-GPU box the test reports itself as skipped."""
import os
import subprocess
import sys

import numpy as np
import pytest

from julia_b200 import lib as L
from util import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("bc", ["periodic", "outflow"])
def test_two_slabs_bit_identical(built, tmp_path, bc):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path / "res.npz")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300),
           os.path.join(ROOT, "scripts", "multi_gpu_check.py"), "--bc", bc, "--out", out]
    subprocess.check_call(cmd, cwd=ROOT)
    r = np.load(out)
    assert float(r["max_abs_diff"]) == 0.0
    assert float(r["dt_diff"]) == 0.0

"""Multi-GPU runs of the slab-decomposed solver (halo exchange per stage + max-allreduce per step):
must be BIT-identical to the single-GPU run at 2, 4 and 8 ranks -- Orszag-Tang and blast wave 512x1024,
1024-row slabs, and the outflow shock-cloud problem (julia_b200/slabcheck.py).  Needs that many GPUs
(gpurun --gpus N); on a smaller box the case reports itself as skipped."""
import json
import os
import subprocess
import sys

import pytest

from util import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slabs_bit_identical(built, tmp_path, world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    out = str(tmp_path / "res.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300),
           os.path.join(ROOT, "scripts", "multi_gpu_check.py"), "--out", out]
    subprocess.check_call(cmd, cwd=ROOT)
    with open(out) as f:
        r = json.load(f)
    assert r["ranks"] == world
    for case, v in r["cases"].items():
        assert v["max_abs_diff"] == 0.0 and v["max_dt_diff"] == 0.0, (case, v)
    assert r["slab_bit_identical"] is True

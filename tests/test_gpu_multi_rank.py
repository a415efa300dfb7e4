"""One process per GPU (torchrun, NCCL): the multi-GPU owner-computes step against the oracle, rank by rank
(tests/owner_rank_check.py). Needs two visible GPUs; skipped on a single-GPU box, where tests/test_gpu_owner_partition.py
covers the same comparison with the ranks' handles created one after the other on one device."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("order,cells", [(2, 5), (1, 8)])
def test_two_ranks_against_the_oracle(order, cells):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + order), os.path.join(ROOT, "tests", "owner_rank_check.py"), "--cells", str(cells), "--order", str(order)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("-> OK") == world

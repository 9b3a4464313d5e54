"""GPU test of the slab-sharded path (needs >= 2 GPUs; skipped otherwise): sharded == single-GPU, bit for bit."""
import os
import subprocess
import sys

import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim", [2, 3])
def test_sharded_step_is_bit_identical(dim):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + dim), os.path.join(REPO, "tools", "sharded_check.py"), str(dim), "12"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "BIT-IDENTICAL" in r.stdout

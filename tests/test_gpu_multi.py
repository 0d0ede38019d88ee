"""GPU suite, part 3: the sharded path on 2 GPUs (skipped on a 1-GPU box): tools/dist_check.py under torchrun."""
import os
import subprocess
import sys

import pytest

import common
from pangene_b200 import engine

pytestmark = pytest.mark.gpu


def test_two_ranks_match_the_oracle():
    if engine.lib().pgx_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29541",
                        os.path.join(common.ROOT, "tools", "dist_check.py")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert r.returncode == 0 and b"DIST_CHECK PASS" in r.stdout, r.stdout.decode()[-3000:]

"""GPU suite, part 3: the sharded path (genomes partitioned over ranks, collectives inside libpgx): tools/dist_check.py under torchrun against
the single-process oracle -- replicated graph identical on every rank and equal to the oracle's, per-rank W and BED lines equal to the oracle's.
  * with >= 2 GPUs: one rank per GPU over NCCL (the production transport);
  * ALWAYS, also on a single-GPU box: 2 and 3 ranks sharing the device over the loopback transport (shared-memory staged collectives,
    pgx_comm.cuh) -- every partition / merge / fallback decision of a multi-rank run, minus NVLink."""
import os
import subprocess
import sys

import pytest

import common
from pangene_b200 import engine

pytestmark = pytest.mark.gpu


def run_dist_check(n_ranks, port, env_extra):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % n_ranks, "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(common.ROOT, "tools", "dist_check.py")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900, env=env)
    assert r.returncode == 0 and b"DIST_CHECK PASS" in r.stdout, r.stdout.decode()[-3000:]


def test_two_ranks_match_the_oracle():
    if engine.lib().pgx_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    run_dist_check(2, 29541, {})


@pytest.mark.parametrize("n_ranks", [2, 3])
def test_ranks_sharing_one_gpu_over_the_loopback_transport(n_ranks):
    run_dist_check(n_ranks, 29543 + n_ranks, {"PGX_LOOPBACK": "1"})

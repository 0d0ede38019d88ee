"""CPU suite, part 3: the launcher-side plumbing of the sharded path with world_size 2 over gloo (no GPU):
partitioning, the unique-id broadcast, re-assembly of per-rank output in genome order, and -- with the oracle
standing in for the device -- that the cross-genome reductions the sharded path relies on are shard-invariant."""
import ctypes as C
import os
import subprocess
import sys
import textwrap

import pytest

import common
from pangene_b200 import dist as pgdist, hostio


def test_partition_covers_everything_once():
    for n in [1, 2, 7, 100, 1000]:
        for w in [1, 2, 3, 4, 8]:
            parts = pgdist.partition(n, w)
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(e - b for b, e in parts) - min(e - b for b, e in parts) <= 1


WORKER = textwrap.dedent("""
    import os, sys, hashlib
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
    import torch, torch.distributed as dist
    import common
    from pangene_b200 import dist as pgdist, hostio
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    uid = pgdist.broadcast_unique_id(lambda: bytes(range(128)), rank)
    assert uid == bytes(range(128))
    G = 5
    b, e = pgdist.partition(G, world)[rank]
    opt = hostio.default_opt()
    d = hostio.synth(opt, hostio.synth_params(2, n_genome=e - b, genome_begin=b, n_gene=300, iso_per_gene=1, hits_per_genome=700, n_genome_total=G))
    # per-rank partial of a cross-genome reduction the device path allreduces (hit.c:231-238): max score per protein
    mx = [0] * d.contents.n_prot
    for j in range(d.contents.n_genome):
        g = d.contents.genome[j]
        for i in range(g.n_hit):
            h = g.hit[i]
            mx[h.pid] = max(mx[h.pid], h.score_ori)
    t = torch.tensor(mx, dtype=torch.int32)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    labels = [d.contents.genome[j].label.decode() for j in range(d.contents.n_genome)]
    parts = pgdist.gather_in_rank_order(labels, world)
    if rank == 0:
        whole = hostio.synth(opt, hostio.synth_params(2, n_genome=G, n_gene=300, iso_per_gene=1, hits_per_genome=700, n_genome_total=G))
        ref = [0] * whole.contents.n_prot
        for j in range(G):
            g = whole.contents.genome[j]
            for i in range(g.n_hit):
                ref[g.hit[i].pid] = max(ref[g.hit[i].pid], g.hit[i].score_ori)
        assert ref == t.tolist()
        assert sum(parts, []) == [whole.contents.genome[j].label.decode() for j in range(G)]
        print("GLOO_OK")
    dist.destroy_process_group()
""")


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % {"root": common.ROOT})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=240)
    assert r.returncode == 0 and b"GLOO_OK" in r.stdout, r.stdout.decode()[-2000:]

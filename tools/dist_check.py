#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun (one rank per GPU):
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
With PGX_LOOPBACK=1 the ranks share whatever GPUs are visible (two ranks on one GPU work) and libpgx's collectives go through shared memory.
Every rank generates ITS genome shard, the sharded GPU path runs with NCCL collectives inside libpgx, and the
result is compared with the single-process oracle on the whole input: replicated graph identical on every rank
and equal to the oracle's, per-rank BED/W lines equal to the oracle's lines for those genomes."""
import hashlib
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist

import common
from oracle import pgoracle
from pangene_b200 import dist as pgdist, engine, hostio

CASES = [dict(config=2, n_genome=7, n_gene=1800, iso_per_gene=4, hits_per_genome=16000, args=[]),
         dict(config=2, n_genome=9, n_gene=2500, iso_per_gene=1, hits_per_genome=6000, args=["-S"]),
         dict(config=5, n_genome=64, n_gene=1200, iso_per_gene=1, hits_per_genome=1000, args=["-p0", "-a2"]),
         dict(config=2, n_genome=3, n_gene=900, iso_per_gene=4, hits_per_genome=8000, args=[])]   # fewer genomes than ranks at 4/8 GPUs: empty shards
# PGX_HOST_BUILD=1 (tests/test_emulated_path.py): the library behind PGX_LIB is the host build of libpgx on the SIMT shim -- no device to select,
# loopback transport, and inputs small enough for a build that is ~10^4 times slower
CASES_HOST_BUILD = [dict(config=2, n_genome=4, n_gene=150, iso_per_gene=4, hits_per_genome=1100, args=[]),
                    dict(config=5, n_genome=7, n_gene=160, iso_per_gene=1, hits_per_genome=150, args=["-p0", "-a2"]),
                    dict(config=2, n_genome=1, n_gene=120, iso_per_gene=1, hits_per_genome=300, args=["-S"])]   # one genome: the second rank's shard is empty


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    loopback = os.environ.get("PGX_LOOPBACK") == "1"       # ranks share the GPUs there are; collectives through shared memory (pgx_comm.cuh)
    host_build = os.environ.get("PGX_HOST_BUILD") == "1"
    if host_build:
        loopback, local = True, 0
    else:
        if loopback:
            local = local % torch.cuda.device_count()
        torch.cuda.set_device(local)
    if loopback:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = engine.lib()
    L.pgx_set_device(local)
    engine.set_verbose(1)
    ok = True
    only = os.environ.get("PGX_HOST_BUILD_CASES")                  # (host build: a comma-separated subset of its cases)
    for ci, case in enumerate(CASES_HOST_BUILD if host_build else CASES):
        if host_build and only and str(ci) not in only.split(","):
            continue
        G = case["n_genome"]
        b, e = pgdist.partition(G, world)[rank]
        if loopback:
            pgdist.init_comm_loopback(L, rank, world, G, b, tag=str(ci))
        else:
            pgdist.init_comm(L, rank, world, G, b)
        opt = common.opt_from_args(case["args"])
        p = hostio.synth_params(case["config"], n_genome=e - b, genome_begin=b, n_gene=case["n_gene"], iso_per_gene=case["iso_per_gene"], hits_per_genome=case["hits_per_genome"], n_genome_total=G)
        d = hostio.synth(opt, p)
        g = engine.hot_path(opt, d, lazy=True)
        sl = common.gfa_bytes(g, walk=False)
        with tempfile.NamedTemporaryFile() as t:
            hostio.write_gfa(g, t.name, False, True)
            wl = open(t.name, "rb").read()
        bed = common.bed_bytes(d)
        if host_build:   # the GFA route of the sharded run (lazy mode 2: graph + walk tokens, no pg_hit_t back) must print the same S/L and W lines
            d2 = hostio.synth(opt, p)
            g2, tok, off, n_tok = engine.graph_and_tokens(opt, d2)
            with tempfile.NamedTemporaryFile() as t:
                hostio.write_gfa_tokens(g2, tok, off, t.name, True, False)
                sl2 = open(t.name, "rb").read()
                hostio.write_gfa_tokens(g2, tok, off, t.name, False, True)
                wl2 = open(t.name, "rb").read()
            engine.release_tokens(d2, tok)
            if sl2 != sl or wl2 != wl:
                print("case %d rank %d: the token route differs from the pg_hit_t route (%d tokens)" % (ci, rank, n_tok), flush=True)
                ok = False
            engine.destroy(g2); hostio.destroy(d2)
        parts = pgdist.gather_in_rank_order((hashlib.md5(sl).hexdigest(), wl, bed), world)
        if rank == 0:
            po = hostio.synth_params(case["config"], n_genome=G, n_gene=case["n_gene"], iso_per_gene=case["iso_per_gene"], hits_per_genome=case["hits_per_genome"], n_genome_total=G)
            do = hostio.synth(opt, po)
            go = pgoracle.hot_path(opt, do)
            same_graph = len(set(x[0] for x in parts)) == 1
            sl_ok = common.gfa_bytes(go, walk=False) == sl
            with tempfile.NamedTemporaryFile() as t:
                hostio.write_gfa(go, t.name, False, True)
                wo = open(t.name, "rb").read()
            w_ok = sorted(wo.splitlines()) == sorted(b"".join(x[1] for x in parts).splitlines())
            bed_ok = sorted(common.bed_bytes(do).splitlines()) == sorted(b"".join(x[2] for x in parts).splitlines())
            print("case %d (%d genomes over %d ranks): replicated graph identical=%s  S/L==oracle=%s  W==oracle=%s  BED==oracle=%s  (%d S, %d L)"
                  % (ci, G, world, same_graph, sl_ok, w_ok, bed_ok, g.contents.n_seg, g.contents.n_arc), flush=True)
            ok = ok and same_graph and sl_ok and w_ok and bed_ok
        engine.destroy(g); hostio.destroy(d)
        L.pgx_comm_destroy()
        dist.barrier()
    if host_build:   # (every rank checked its own token route)
        oks = pgdist.gather_in_rank_order(ok, world)
        ok = all(oks)
    if rank == 0:
        print("DIST_CHECK", "PASS" if ok else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""BASELINE.json configs[4] at FULL size under torchrun (one rank per GPU): 1000 bacterial genomes x 5k hits, `-p0 -a2`, genomes sharded over the
ranks, NCCL collectives inside libpgx; the GFA assembled from the replicated S/L lines and the rank-gathered W lines goes through gfa2matrix
(pgio_gfa2matrix, the restatement of pangene.js:1168-1246) and must give the SAME presence/absence matrix -- and the same sorted GFA lines -- as the
single-process oracle on the whole input.
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 tools/dist_c5.py [n_genomes=1000]"""
import hashlib
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist

import common
from oracle import pgoracle
from pangene_b200 import dist as pgdist, engine, hostio


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = engine.lib()
    L.pgx_set_device(local)
    engine.set_verbose(1)
    b, e = pgdist.partition(G, world)[rank]
    if world > 1:
        pgdist.init_comm(L, rank, world, G, b)
    opt = common.opt_from_args(["-p0", "-a2"])
    d = hostio.synth(opt, hostio.synth_params(5, n_genome=e - b, genome_begin=b, n_genome_total=G))
    n_local = hostio.total_hits(d)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    g = engine.hot_path(opt, d, lazy=True)
    dt = time.perf_counter() - t0
    st = engine.stats()
    sl = common.gfa_bytes(g, walk=False)
    with tempfile.NamedTemporaryFile() as t:
        hostio.write_gfa(g, t.name, False, True)
        wl = open(t.name, "rb").read()
    parts = pgdist.gather_in_rank_order((hashlib.md5(sl).hexdigest(), wl, n_local, dt, st.redo), world) if world > 1 else [(hashlib.md5(sl).hexdigest(), wl, n_local, dt, st.redo)]
    ok = True
    if rank == 0:
        same_graph = len(set(x[0] for x in parts)) == 1
        gpu_gfa = sl + b"".join(x[1] for x in parts)
        t1 = time.perf_counter()
        do = hostio.synth(opt, hostio.synth_params(5, n_genome=G, n_genome_total=G))
        go = pgoracle.hot_path(opt, do)
        t_oracle = time.perf_counter() - t1
        with tempfile.TemporaryDirectory() as td:
            f1, f2, m1, m2 = [os.path.join(td, x) for x in ("cpu.gfa", "gpu.gfa", "cpu.tsv", "gpu.tsv")]
            hostio.write_gfa(go, f1)
            open(f2, "wb").write(gpu_gfa)
            cpu_gfa = open(f1, "rb").read()
            n1 = hostio.lib().pgio_gfa2matrix(os.fsencode(f1), os.fsencode(m1), 0)
            n2 = hostio.lib().pgio_gfa2matrix(os.fsencode(f2), os.fsencode(m2), 0)
            mat_ok = n1 == n2 and open(m1, "rb").read() == open(m2, "rb").read()
            mat_md5 = hashlib.md5(open(m2, "rb").read()).hexdigest()
        gfa_ok = sorted(cpu_gfa.splitlines()) == sorted(gpu_gfa.splitlines())
        n_hits = sum(x[2] for x in parts)
        print("config 5: %d genomes, %d hits over %d rank(s): replicated graph identical=%s  GFA lines==oracle=%s  matrix==oracle=%s (%d genes x %d assemblies, md5 %s)  "
              "S=%d L=%d  hot path %.1f ms end to end (max over ranks), careful redos %d, oracle %.1f s on one host core"
              % (G, n_hits, world, same_graph, gfa_ok, mat_ok, n2, G, mat_md5, g.contents.n_seg, g.contents.n_arc, max(x[3] for x in parts) * 1e3, max(x[4] for x in parts), t_oracle), flush=True)
        ok = same_graph and gfa_ok and mat_ok
        print("DIST_C5", "PASS" if ok else "FAIL", flush=True)
    engine.destroy(g); hostio.destroy(d)
    if world > 1:
        dist.barrier()
        L.pgx_comm_destroy()
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

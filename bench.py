#!/usr/bin/env python
"""bench.py -- PAF hits/s from parsed hits to GFA arcs (BASELINE.json metric) on N B200s of one node.

A "step" is one pass of the whole hot path (per-genome load-time block of read.c:243-260 + pg_post_process +
pg_graph_gen, 15 branch rounds) over the synthetic configuration BASELINE.json quotes the metric on:
configs[2] = 100 genomes x 200k hits (80k proteins / 20k genes, 4 isoforms, 24 contigs), generated in memory
by libpgio (pgio_synth, seed 3).  N > 1 shards the SAME 100 genomes across ranks ("strong" scaling); one process
per GPU, NCCL collectives inside libpgx.

  value        hits/s with the inputs already resident in HBM (pgx_upload done): K timed steps of pgx_run,
               device time from CUDA events on the library's stream, max over ranks.
  e2e          the same metric through the reference-facing C ABI with HOST buffers: pg_post_process() +
               pg_graph_init() + pg_graph_gen() on a malloc'd pg_data_t, H2D of the raw hits and D2H of the
               finished hits + graph inside the timed region (wall clock around the calls, which return
               synchronously).
               Median of five timed calls after two untimed ones (`ms_samples` lists the five); `detail.e2e_gfa_route`
               (single-GPU lines) is the same C ABI on its GFA route: graph + walk tokens out, no pg_hit_t rebuilt.
  roofline     the first (load-time) overlap sweep -- the one launch that streams ~all hits: algorithmic bytes
               (52 + 8 n_exon per unfiltered hit, 4 per filtered hit) / CUDA-event time, against
               the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline the compiled, unmodified reference (oracle/_ref, 1 thread -- pangene is single-threaded) on a
               bounded sample of the same workload, timed on this box's host cores in the same run.

  parity       correctness gate on the graph the timed calls produced: S/L lines (replicated on every rank) and W lines (gathered in
               rank order) are hashed and compared with tests/golden/c3_golden.json, which was printed by the UNMODIFIED reference
               (stock binary and in-process objects, tests/golden/make_golden_c3.py).  "ok" or the process exits non-zero.

  --impl reference   times the reference's own CPU implementation (oracle/_ref, unmodified, single thread) on the SAME configuration:
                     ONE full pass over all 100 genomes (~95 s on one host core; K/W are not repeated, the line says steps 1).
"""
import argparse
import ctypes as C
import faulthandler
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "PAF hits/s to GFA (parsed hits -> GFA arcs; block of read.c:243-260 + pg_post_process + pg_graph_gen)"
WORKLOAD = "BASELINE configs[2]: synthetic 100 genomes x 200k hits, 80k proteins / 20k genes (4 isoforms), 24 contigs, default options (T=15)"
CPU_SAMPLE_GENOMES = 5


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(sm)}


def cpu_reference(hostio, n_genomes, steps=1, warmup=0):
    """hits/s of the reference's own CPU implementation (1 thread) on the first n_genomes genomes of the workload (100 = all of it)"""
    import refharness
    opt = hostio.default_opt()
    p = hostio.synth_params(3, n_genome=n_genomes, n_genome_total=100)
    kind = "reference" if refharness.available() else "port"
    times, hits, digest = [], 0, None
    for it in range(warmup + steps):
        d = hostio.synth(opt, p)
        hits = hostio.total_hits(d)
        t0 = time.perf_counter()
        if kind == "reference":
            rh = refharness.RefHot()
            g, _ = rh.hot_path(opt, d)
            dt = time.perf_counter() - t0
            if n_genomes == 100:
                digest = hostio.gfa_digest(g)
            rh.destroy(g)
        else:
            from oracle import pgoracle
            g = pgoracle.hot_path(opt, d)
            dt = time.perf_counter() - t0
            if n_genomes == 100:
                digest = hostio.gfa_digest(g)
            pgoracle.destroy(g)
        hostio.destroy(d)
        if it >= warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    sample = ("none: all 100 genomes (%d hits), %.1f s per pass" % (hits, dt)) if n_genomes == 100 else \
        ("%d of the 100 genomes of the same workload (%d hits), %.1f s per pass" % (n_genomes, hits, dt))
    return {"value": hits / dt, "unit": "hits/s", "cores": 1, "kind": kind, "sample": sample + ", single thread (the reference has no threads)"}, dt, hits, digest


def load_golden():
    p = os.path.join(ROOT, "tests", "golden", "c3_golden.json")
    return json.load(open(p)) if os.path.exists(p) else None


def parity_verdict(sl_md5s, w_bytes, gold):
    """sl_md5s: one S/L digest per rank (the graph is replicated); w_bytes: W lines of all ranks in rank (= genome) order"""
    w_md5 = hashlib.md5(w_bytes).hexdigest()
    out = {"golden": "tests/golden/c3_golden.json (unmodified reference: stock binary + in-process objects)", "sl_md5": sl_md5s[0], "w_md5": w_md5,
           "n_W": w_bytes.count(b"\n"), "ranks_agree": len(set(sl_md5s)) == 1}
    if gold is None:
        out["status"] = "unchecked (golden file missing)"
    else:
        ok = out["ranks_agree"] and sl_md5s[0] == gold["sl_md5"] and w_md5 == gold["w_md5"] and out["n_W"] == gold["n_W"]
        out["status"] = "ok" if ok else "FAIL"
    return out


def workload_config(n_total, G_total, world):
    return {"workload": WORKLOAD if G_total == 100 else "DEBUG %d genomes" % G_total, "hits": n_total, "genomes": G_total,
            "parallelism": "genomes sharded over %d rank(s)" % world,
            "l2": "inputs (%.1f GB of hit columns + exons per step) exceed the 126 MB L2; no explicit flush" % ((n_total or 0) * 140 / 1e9)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genomes", type=int, default=100, help="override the genome count (debug; invalidates the number)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-sample", type=int, default=0, help="--impl reference on this many genomes instead of the full configuration (debug)")
    ap.add_argument("--no-parity", action="store_true", help="skip the golden comparison (debug)")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    from pangene_b200 import hostio

    faulthandler.enable()
    if a.impl == "reference":
        if rank != 0:
            return 0
        full = a.genomes == 100 and not a.ref_sample
        cb, dt, hits, digest = cpu_reference(hostio, 100 if full else (a.ref_sample or a.genomes))
        par = parity_verdict([digest[0]], digest[1], load_golden()) if digest else {"status": "unchecked (sample run)"}
        cfg = workload_config(hits if full else None, 100, a.gpus)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "hits/s", "n_gpus": a.gpus, "steps": 1, "warmup": 0,
                          "steps_requested": a.steps, "warmup_requested": a.warmup,
                          "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
                          "config": cfg, "cpu_baseline": cb, "parity": par,
                          "note": "one full pass of the whole configuration (K/W are not repeated: a pass takes ~1.5 min on the one core the reference can use)",
                          "e2e": {"value": cb["value"], "unit": "hits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return 0

    import torch
    from pangene_b200 import engine
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libpgx has no CPU fallback")
    torch.cuda.set_device(local_rank)
    L = engine.lib()
    L.pgx_set_device(local_rank)
    engine.set_verbose(1)
    dist = None
    G_total = a.genomes
    g_begin, g_end = G_total * rank // world, G_total * (rank + 1) // world
    if world > 1:
        import torch.distributed as dist
        from pangene_b200 import dist as pgdist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        pgdist.init_comm(L, rank, world, G_total, g_begin)

    opt = hostio.default_opt()
    p = hostio.synth_params(3, n_genome=g_end - g_begin, genome_begin=g_begin, n_genome_total=G_total)
    d0 = hostio.synth(opt, p)               # this rank's shard, raw parsed hits
    n_local = hostio.total_hits(d0)
    n_exon = sum(d0.contents.genome[j].n_exon for j in range(d0.contents.n_genome))
    n_total = n_local
    if world > 1:
        t = torch.tensor([n_local], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        n_total = int(t.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: device-resident ---------------------------------------------------------------------
    sess = L.pgx_upload(d0)
    L.pgx_set_sweep_stat(1)                       # the roofline numerator: counted once, in the first warm-up pass
    for w in range(a.warmup):
        L.pgx_run(sess, C.byref(opt), 3)
        L.pgx_set_sweep_stat(0)
    L.pgx_set_sweep_stat(0)
    barrier()
    st0 = engine.stats()
    dev_ms, sweep_ms, st_ms = [], [], ([], [])
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        for _ in range(a.steps):
            L.pgx_run(sess, C.byref(opt), 3)
            s = engine.stats()
            dev_ms.append(s.run_ms[2]); sweep_ms.append(s.sweep_ms); st_ms[0].append(s.run_ms[0]); st_ms[1].append(s.run_ms[1])
        barrier()
        wall = time.perf_counter() - t0
    st1 = engine.stats()
    ms_step = max_over_ranks(sum(dev_ms) / len(dev_ms))
    stage1_ms, stage2_ms = max_over_ranks(sum(st_ms[0]) / len(st_ms[0])), max_over_ranks(sum(st_ms[1]) / len(st_ms[1]))
    wall_step = max_over_ranks(wall / a.steps * 1e3)
    launches = (st1.launches - st0.launches) // a.steps
    fast_rounds, redos = (st1.fast_rounds - st0.fast_rounds) // a.steps, st1.redo - st0.redo
    sweep_bytes, sweep_t = st1.sweep_bytes, sum(sweep_ms) / len(sweep_ms)
    n_active = st1.n_active
    L.pgx_session_free(sess)

    # ---- e2e: host pg_data_t in, host pg_data_t + pg_graph_t out --------------------------------------
    e2e_t, h2d, d2h = [], 0, 0
    digest = None
    L.pgx_set_lazy(1)
    n_e2e = 2 + 5
    for it in range(n_e2e):
        d = hostio.lib().pgio_data_clone(d0)
        barrier()
        s0 = engine.stats()
        t0 = time.perf_counter()
        L.pg_post_process(C.byref(opt), d)
        g = L.pg_graph_init(d)
        L.pg_graph_gen(C.byref(opt), g)
        dt = time.perf_counter() - t0
        s1 = engine.stats()
        barrier()
        if it >= 2:
            e2e_t.append(max_over_ranks(dt))
        h2d, d2h = s1.h2d_bytes - s0.h2d_bytes, s1.d2h_bytes - s0.d2h_bytes
        if it == n_e2e - 1 and not a.no_parity:
            digest = hostio.gfa_digest(g)       # the graph and hits the LAST timed call returned (outside the timed region)
        L.pg_graph_destroy(g)
        hostio.lib().pgio_data_clone_destroy(d)
    e2e_s = statistics.median(e2e_t)           # median of 5 timed calls after 2 untimed ones (host page-fault/NUMA noise is +-30 %)

    # ---- the GFA-only route of the same C ABI (what the CLI runs for `pangene *.paf > out.gfa`): lazy mode 2 returns the graph, the hits stay
    # on the device and only the walk tokens of the W lines (12 B per printed hit) come back; checked against the same golden
    # (single-GPU lines only: the sharded run's contract is value / e2e / parity above)
    gfa_t, gfa_d2h, digest_tok = [], 0, None
    n_gfa = (1 + 3) if world == 1 else 0
    for it in range(n_gfa):
        d = hostio.lib().pgio_data_clone(d0)
        barrier()
        s0 = engine.stats()
        t0 = time.perf_counter()
        g, tok, off, n_tok = engine.graph_and_tokens(opt, d)
        dt = time.perf_counter() - t0
        gfa_d2h = engine.stats().d2h_bytes - s0.d2h_bytes
        barrier()
        if it >= 1:
            gfa_t.append(max_over_ranks(dt))
        if it == n_gfa - 1 and not a.no_parity:
            digest_tok = hostio.gfa_digest_tokens(g, tok, off)
        engine.release_tokens(d, tok)
        L.pg_graph_destroy(g)
        hostio.lib().pgio_data_clone_destroy(d)
    gfa_s = statistics.median(gfa_t) if gfa_t else None

    # ---- parity gate: every rank's replicated S/L lines + the W lines of all ranks in genome order, against the reference's golden
    parity = {"status": "unchecked (--no-parity)"}
    if digest is not None:
        parts = [digest]
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, digest)
        if rank == 0:
            parity = parity_verdict([x[0] for x in parts], b"".join(x[1] for x in parts), load_golden() if G_total == 100 else None)
            if G_total != 100:
                parity["status"] = "unchecked (debug genome count)"
    parity_tok = {"status": "unchecked (--no-parity)"}
    if digest_tok is not None:
        parts = [digest_tok]
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, digest_tok)
        if rank == 0:
            parity_tok = parity_verdict([x[0] for x in parts], b"".join(x[1] for x in parts), load_golden() if G_total == 100 else None)
            if G_total != 100:
                parity_tok["status"] = "unchecked (debug genome count)"
            if parity_tok["status"] == "FAIL":
                parity["status"] = "FAIL"          # one gate: either route failing fails the run

    # ---- teardown, the same on every rank: flush what we know first, then barrier -> libpgx communicator -> torch process group
    def teardown():
        if world > 1:
            barrier()
            L.pgx_comm_destroy()
            dist.barrier()
            dist.destroy_process_group()

    if rank != 0:
        teardown()
        return 0
    peak, peak_src = peaks()
    achieved = sweep_bytes / (sweep_t * 1e-3) / 1e9 if sweep_t > 0 else 0.0
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    if os.path.exists(tp) and world == 1 and a.genomes == 100:   # (a constant from the committed ncu capture of this kernel, not measured in this run: the label says which build)
        tj = json.load(open(tp))
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), "%s; %s" % (tj.get("source"), tj.get("build"))
    gfa_route = None
    if gfa_s is not None:
        gfa_route = {"value": n_total / gfa_s, "unit": "hits/s", "ms_per_step": gfa_s * 1e3, "d2h_bytes_per_step": int(gfa_d2h), "parity": parity_tok["status"],
                     "api": "pgx_set_lazy(2): pg_post_process + pg_graph_init + pg_graph_gen + pgx_walk_tokens -- host pg_data_t in, graph + walk tokens out (everything a GFA needs; no pg_hit_t record is rebuilt)"}
    out = {
        "metric": METRIC, "value": n_total / (ms_step * 1e-3), "unit": "hits/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
        "config": workload_config(n_total, G_total, world),
        "detail": {"active_hits_after_load_filters_rank0": int(n_active), "wall_ms_per_step_incl_host": wall_step, "stage_ms": [stage1_ms, stage2_ms],
                   "ms_per_step_rank0": {"min": min(dev_ms), "median": statistics.median(dev_ms), "max": max(dev_ms)}, "fast_rounds_per_step": int(fast_rounds), "careful_redos": int(redos)},
        "parity": parity,
        "e2e": {"value": n_total / e2e_s, "unit": "hits/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3,
                "ms_samples": [round(x * 1e3, 1) for x in e2e_t],       # the five timed calls behind the median (this leg is host-memory bound and varies from box to box: DESIGN.md 6)
                "api": "pg_post_process + pg_graph_init + pg_graph_gen on malloc'd pg_data_t (lazy materialisation, as the CLI does); bytes are rank 0's, counted by the library: 48 B/hit + exons up, 24 B/hit + graph down"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_sweep_tiled<SW_LOAD, 1> (first overlap sweep, pg_shadow cal_dom_sc=1; at N > 1: this rank's shard)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": sweep_bytes, "ms_per_launch": sweep_t},
        "clocks": clk.summary(),
    }
    if gfa_route is not None:
        out["detail"]["e2e_gfa_route"] = gfa_route
    out["detail"]["library"] = os.path.relpath(getattr(engine, "LIB_PATH", "?"), ROOT)   # pangene_b200/csrc/libpgx.so unless PGX_LIB pointed the binding at another build
    if world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"] = cpu_reference(hostio, CPU_SAMPLE_GENOMES)[0]
    print(json.dumps(out), flush=True)        # ONE line, flushed before the teardown below can go wrong
    teardown()
    return 0 if parity["status"] in ("ok",) or parity["status"].startswith("unchecked") else 1


if __name__ == "__main__":
    sys.exit(main())

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
  roofline     the first (load-time) overlap sweep -- the one launch that streams ~all hits: algorithmic bytes
               (52 + 8 n_exon per unfiltered hit, 4 per filtered hit) / CUDA-event time, against
               the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline the compiled, unmodified reference (oracle/_ref, 1 thread -- pangene is single-threaded) on a
               bounded sample of the same workload, timed on this box's host cores in the same run.

  --impl reference   times only that CPU baseline and prints it in the same JSON shape.
"""
import argparse
import ctypes as C
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


def cpu_reference_sample(hostio, steps=1, warmup=0):
    """hits/s of the reference's own CPU implementation (1 thread) on the first CPU_SAMPLE_GENOMES genomes of the workload"""
    import refharness
    opt = hostio.default_opt()
    p = hostio.synth_params(3, n_genome=CPU_SAMPLE_GENOMES, n_genome_total=100)
    kind = "reference" if refharness.available() else "port"
    times, hits = [], 0
    for it in range(warmup + steps):
        d = hostio.synth(opt, p)
        hits = hostio.total_hits(d)
        t0 = time.perf_counter()
        if kind == "reference":
            rh = refharness.RefHot()
            g, _ = rh.hot_path(opt, d)
            dt = time.perf_counter() - t0
            rh.destroy(g)
        else:
            from oracle import pgoracle
            g = pgoracle.hot_path(opt, d)
            dt = time.perf_counter() - t0
            pgoracle.destroy(g)
        hostio.destroy(d)
        if it >= warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    return {"value": hits / dt, "unit": "hits/s", "cores": 1, "kind": kind,
            "sample": "%d of the 100 genomes of the same workload (%d hits), %.1f s per pass, single thread (the reference has no threads)" % (CPU_SAMPLE_GENOMES, hits, dt)}, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genomes", type=int, default=100, help="override the genome count (debug; invalidates the number)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    from pangene_b200 import hostio

    if a.impl == "reference":
        if rank != 0:
            return
        cb, dt = cpu_reference_sample(hostio, steps=max(1, a.steps), warmup=min(a.warmup, 1))
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "hits/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
                          "config": {"workload": WORKLOAD, "sample": cb["sample"]}, "cpu_baseline": cb,
                          "e2e": {"value": cb["value"], "unit": "hits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return

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
    dev_ms, sweep_ms = [], []
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        for _ in range(a.steps):
            L.pgx_run(sess, C.byref(opt), 3)
            s = engine.stats()
            dev_ms.append(s.run_ms[2]); sweep_ms.append(s.sweep_ms)
        barrier()
        wall = time.perf_counter() - t0
    st1 = engine.stats()
    ms_step = max_over_ranks(sum(dev_ms) / len(dev_ms))
    wall_step = max_over_ranks(wall / a.steps * 1e3)
    launches = (st1.launches - st0.launches) // a.steps
    sweep_bytes, sweep_t = st1.sweep_bytes, sum(sweep_ms) / len(sweep_ms)
    n_active = st1.n_active
    L.pgx_session_free(sess)

    # ---- e2e: host pg_data_t in, host pg_data_t + pg_graph_t out --------------------------------------
    e2e_t, h2d, d2h = [], 0, 0
    L.pgx_set_lazy(1)
    for it in range(2 + 5):
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
        L.pg_graph_destroy(g)
        hostio.lib().pgio_data_clone_destroy(d)
    e2e_s = statistics.median(e2e_t)           # median of 5 timed calls after 2 untimed ones (host page-fault/NUMA noise is +-30 %)

    if rank != 0:
        return
    peak, peak_src = peaks()
    achieved = sweep_bytes / (sweep_t * 1e-3) / 1e9 if sweep_t > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    if os.path.exists(tp) and world == 1 and a.genomes == 100:
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    out = {
        "metric": METRIC, "value": n_total / (ms_step * 1e-3), "unit": "hits/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
        "config": {"workload": WORKLOAD if a.genomes == 100 else "DEBUG %d genomes" % a.genomes, "hits": n_total, "genomes": G_total, "parallelism": "genomes sharded over %d rank(s)" % world,
                   "l2": "inputs (%.1f GB of hit columns + exons per step) exceed the 126 MB L2; no explicit flush" % ((n_total * 140) / 1e9),
                   "active_hits_after_load_filters_rank0": int(n_active), "wall_ms_per_step_incl_host": wall_step},
        "e2e": {"value": n_total / e2e_s, "unit": "hits/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3,
                "api": "pg_post_process + pg_graph_init + pg_graph_gen on malloc'd pg_data_t (lazy materialisation, as the CLI does); bytes are rank 0's, counted by the library: 48 B/hit + exons up, 24 B/hit + graph down"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_sweep_tiled<SW_LOAD> (first overlap sweep, pg_shadow cal_dom_sc=1)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": sweep_bytes, "ms_per_launch": sweep_t},
        "clocks": clk.summary(),
    }
    if world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"], _ = cpu_reference_sample(hostio)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

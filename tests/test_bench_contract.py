"""bench.py's driver-facing contract, as far as it can be checked without a GPU: the reference arm (`--impl reference`, here on a one-genome
sample of the workload) prints ONE JSON line with the agreed keys, rank != 0 prints nothing and exits 0, the parity verdict compares what it
should, and the GPU arm refuses to run without a device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_bench(args, env_extra=None):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)


def test_reference_arm_line():
    r = run_bench(["--impl", "reference", "--ref-sample", "1", "--steps", "3", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    b = json.loads(lines[0])
    assert b["impl"] == "reference" and b["unit"] == "hits/s" and b["higher_is_better"] is True and b["vs_baseline"] is None and b["data"] == "synthetic"
    assert b["metric"].startswith("PAF hits/s to GFA") and b["scaling"] == "strong" and b["n_gpus"] == 1 and b["gpu_launches"] == 0
    assert b["value"] > 0 and abs(b["value"] - b["cpu_baseline"]["value"]) < 1e-6 and b["e2e"]["value"] == b["value"]
    assert b["e2e"]["h2d_bytes_per_step"] == 0 and b["e2e"]["d2h_bytes_per_step"] == 0
    assert b["cpu_baseline"]["kind"] == "reference" and b["cpu_baseline"]["cores"] == 1 and "1 of the 100 genomes" in b["cpu_baseline"]["sample"]
    assert b["config"]["workload"].startswith("BASELINE configs[2]") and "model" not in b["config"]
    assert b["parity"]["status"].startswith("unchecked")            # a sample run has no golden; the full run compares with c3_golden.json


def test_reference_arm_other_ranks_do_nothing():
    r = run_bench(["--impl", "reference", "--gpus", "2"], {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_fallback():
    import torch
    if torch.cuda.is_available():
        return
    r = run_bench(["--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_parity_verdict():
    import bench
    gold = {"sl_md5": "a" * 32, "w_md5": __import__("hashlib").md5(b"W\t1\nW\t2\n").hexdigest(), "n_W": 2}
    ok = bench.parity_verdict(["a" * 32, "a" * 32], b"W\t1\nW\t2\n", gold)
    assert ok["status"] == "ok" and ok["ranks_agree"] and ok["n_W"] == 2
    assert bench.parity_verdict(["a" * 32, "b" * 32], b"W\t1\nW\t2\n", gold)["status"] == "FAIL"       # a rank disagrees about the replicated graph
    assert bench.parity_verdict(["b" * 32], b"W\t1\nW\t2\n", gold)["status"] == "FAIL"
    assert bench.parity_verdict(["a" * 32], b"W\t1\n", gold)["status"] == "FAIL"
    assert bench.parity_verdict(["a" * 32], b"W\t1\nW\t2\n", None)["status"].startswith("unchecked")
    assert bench.load_golden() is not None                         # the golden of the headline configuration is committed


def test_pick_variant(tmp_path):
    log = tmp_path / "ab.log"
    log.write_text("== PGX_SWEEP_VARIANT=0\nper pass (stage 1 + stage 2 ms): 90.1+12.0 8.25+11.66 8.26+11.70  redo: 0\nhits 1 active 2 seg 3 arc 4 launches 5 run_ms 8.2 11.6 sweep_ms 1.100 sweep_GB 0.0\n"
                   "== PGX_SWEEP_VARIANT=1\nper pass (stage 1 + stage 2 ms): 99.1+12.0 8.05+11.66 8.06+11.70  redo: 0\nhits 1 active 2 seg 3 arc 4 launches 5 run_ms 8.2 11.6 sweep_ms 1.000 sweep_GB 0.0\n")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "pick_variant.py"), str(log)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "1" and "variant 0: median 19.935" in r.stderr


def test_gpu_arm_line_with_stub_engine():
    """the Python side of the GPU arm end to end (tests/bench_stub_driver.py: stub engine / hostio / torch.cuda): ONE JSON line with every key of
    the contract, exit code 0 on parity ok"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_stub_driver.py"), "--steps", "4", "--warmup", "3"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    b = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks", "parity", "detail"):
        assert k in b, k
    assert b["steps"] == 4 and b["warmup"] == 3 and b["n_gpus"] == 1 and b["unit"] == "hits/s" and b["dtype"] == "i32" and b["vs_baseline"] is None
    assert b["gpu_launches"] == 415 and abs(b["ms_per_step"] - 20.0) < 0.1 and abs(b["value"] - 19977625 / (b["ms_per_step"] * 1e-3)) < 1
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(b["roofline"]) and b["roofline"]["bound"] == "hbm"
    assert abs(b["roofline"]["frac"] - b["roofline"]["achieved"] / b["roofline"]["peak"]) < 1e-12 and abs(b["roofline"]["achieved"] - 1876.0) < 1e-6
    assert b["roofline"]["traffic"] == json.load(open(os.path.join(ROOT, "profiles", "sweep_traffic.json")))["dram_bytes_per_launch"]
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(b["e2e"]) and len(b["e2e"]["ms_samples"]) == 5
    assert b["e2e"]["h2d_bytes_per_step"] == 1904701000 and b["e2e"]["d2h_bytes_per_step"] == 483129304
    assert b["parity"]["status"] == "ok" and b["detail"]["e2e_gfa_route"]["parity"] == "ok"
    assert "library" in b["detail"]
    assert "workload" in b["config"] and "model" not in b["config"] and "l2" in b["config"]
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(b["clocks"])


def test_gpu_arm_two_ranks_with_stub_engine():
    """the sharded launch of the same arm (torchrun-style environment, two processes, gloo standing in for nccl): rank 0 prints the one line --
    n_gpus 2, the W lines of both ranks put together in rank order for the parity verdict, no single-GPU-only legs -- rank 1 prints nothing, both
    tear down and exit 0"""
    import socket
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        port = so.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "bench_stub_driver.py"), "--gpus", "2", "--steps", "3", "--warmup", "3"],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (so_, se_) in zip(procs, outs):
        assert p.returncode == 0, se_[-3000:]
    lines = [l for l in outs[0][0].splitlines() if l.startswith("{")]
    assert len(lines) == 1 and not [l for l in outs[1][0].splitlines() if l.startswith("{")]
    b = json.loads(lines[0])
    assert b["n_gpus"] == 2 and b["scaling"] == "strong" and b["parity"]["status"] == "ok" and b["parity"]["ranks_agree"] is True
    assert b["value"] > 0 and b["config"]["parallelism"].endswith("2 rank(s)") and "e2e_gfa_route" not in b["detail"] and "cpu_baseline" not in b
    assert b["config"]["hits"] == 2 * 19977625                      # (the stub gives every rank the same count; the line reports the all-reduced total)

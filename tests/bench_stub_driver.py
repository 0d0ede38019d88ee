"""TEST INFRASTRUCTURE for tests/test_bench_contract.py: runs bench.main() -- the GPU arm -- against stub `engine` / `hostio` modules and a stub
torch.cuda, so that the Python side of the benchmark of record (argument handling, timing loops, parity plumbing, the JSON line) is executed on a
box without a GPU.  Nothing here measures anything; the numbers in the line are the stubs'."""
import hashlib
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import pangene_b200  # noqa: E402

gold = json.load(open(os.path.join(ROOT, "tests", "golden", "c3_golden.json")))
W_BYTES = b"".join(b"W\tstub\n" for _ in range(gold["n_W"]))
gold_stub = dict(gold, w_md5=hashlib.md5(W_BYTES).hexdigest())


class Stats:
    def __init__(self, k):
        self.run_ms = [8.0, 12.0, 20.0 + 0.01 * (k % 3)]
        self.sweep_ms, self.sweep_bytes = 1.0, 1.876e9
        self.launches, self.fast_rounds, self.redo = 415 * k, 17 * k, 0
        self.n_active, self.h2d_bytes, self.d2h_bytes = 3278512, 1904701000 * k, 483129304 * k


class FakeLib:
    def __init__(self):
        self.runs = 0

    def __getattr__(self, name):
        def f(*a):
            if name in ("pgx_run", "pg_graph_gen"):
                self.runs += 1
            return 1
        return f


LIB = FakeLib()
engine = types.ModuleType("pangene_b200.engine")
engine.lib = lambda: LIB
engine.set_verbose = lambda v: None
engine.stats = lambda: Stats(LIB.runs)
engine.graph_and_tokens = lambda opt, d: (LIB.pg_graph_gen() and "g", "tok", "off", 7)
engine.release_tokens = lambda d, tok: None


class Genome:
    n_exon = 1190000


class Data:
    n_genome = 100
    genome = [Genome() for _ in range(100)]


class DataPtr:
    contents = Data()


hostio = types.ModuleType("pangene_b200.hostio")
hostio.default_opt = lambda: (__import__("ctypes").c_int * 64)()
hostio.synth_params = lambda *a, **k: None
hostio.synth = lambda opt, p: DataPtr()
hostio.total_hits = lambda d: 19977625
hostio.lib = lambda: LIB
RANK, WORLD = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
n_w_lo, n_w_hi = gold["n_W"] * RANK // WORLD, gold["n_W"] * (RANK + 1) // WORLD          # this rank's share of the W lines
MY_W = W_BYTES[len(b"W\tstub\n") * n_w_lo:len(b"W\tstub\n") * n_w_hi]
hostio.gfa_digest = lambda g: (gold["sl_md5"], MY_W)
hostio.gfa_digest_tokens = lambda g, tok, off: (gold["sl_md5"], MY_W)
hostio.destroy = lambda d: None

sys.modules["pangene_b200.engine"] = engine
sys.modules["pangene_b200.hostio"] = hostio
pangene_b200.engine, pangene_b200.hostio = engine, hostio
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a: None
if WORLD > 1:   # the sharded launch on a CPU box: gloo instead of nccl, host tensors instead of device tensors
    import torch.distributed as dist
    _init, _tensor = dist.init_process_group, torch.tensor
    dist.init_process_group = lambda backend, **kw: _init("gloo")
    torch.tensor = lambda data, **kw: _tensor(data, **{k: v for k, v in kw.items() if k != "device"})

import bench  # noqa: E402
bench.load_golden = lambda: gold_stub
sys.argv = ["bench.py", "--no-cpu-baseline"] + sys.argv[1:]
sys.exit(bench.main())

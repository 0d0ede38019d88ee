#!/usr/bin/env python
"""TEST INFRASTRUCTURE: randomized campaign of the CUDA path (through the C ABI) against the oracle restatement on a GPU box --
the same case generator as tools/fuzz_oracle.py (which pins the oracle against the compiled reference in the build container).
usage: tools/fuzz_gpu.py [first_case | case,case,...] [n_cases] [time_budget_s] [local|cmtie]      (modes: see tools/fuzz_oracle.py)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import common  # noqa: E402
from fuzz_oracle import data_for, random_case  # noqa: E402
from oracle import pgoracle  # noqa: E402
from pangene_b200 import engine, hostio  # noqa: E402


def main():
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    if len(sys.argv) > 1 and "," in sys.argv[1]:
        cases = [int(x) for x in sys.argv[1].split(",")]
    else:
        first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
        cases = list(range(first, first + n))
    budget = float(sys.argv[3]) if len(sys.argv) > 3 else 1e9
    local = len(sys.argv) > 4 and sys.argv[4] == "local"
    cmtie = len(sys.argv) > 4 and sys.argv[4] == "cmtie"
    engine.set_verbose(1)
    t0, bad, done = time.time(), 0, 0
    for k in cases:
        if time.time() - t0 > budget:
            break
        case = random_case(k, local, cmtie)
        opt, d1 = data_for(case)
        _, d2 = data_for(case)
        g1 = pgoracle.hot_path(opt, d1)
        g2 = engine.hot_path(opt, d2, lazy=bool(k & 1))
        st = engine.stats()
        a, b = common.gfa_bytes(g1), common.gfa_bytes(g2)
        ok_gfa = sorted(a.splitlines()) == sorted(b.splitlines())
        ok_graph = common.graph_tuple(g1) == common.graph_tuple(g2)
        ok_bed = sorted(common.bed_bytes(d1).splitlines()) == sorted(common.bed_bytes(d2).splitlines())
        ok_tie = list(st.tie)[:3] == pgoracle.tie_report()[:3]
        ok = ok_gfa and ok_graph and ok_bed   # (the tie counters are reported, not compared: the two detectors count pairs from different sides)
        bad += not ok
        done += 1
        print("case %d %s hits=%d S=%d L=%d tie=%s gfa=%s graph=%s bed=%s tie_same=%s %s" % (k, "ok" if ok else "MISMATCH", hostio.total_hits(d1), g1.contents.n_seg, g1.contents.n_arc, list(st.tie), ok_gfa, ok_graph, ok_bed, ok_tie,
              {x: case[x] for x in ("config", "n_genome", "n_gene", "iso_per_gene", "hits_per_genome", "n_ctg", "seed", "args", "local_count") if x in case}), flush=True)
        if not ok_gfa:
            print(common.first_diff(a, b), flush=True)
        pgoracle.destroy(g1); engine.destroy(g2); hostio.destroy(d1); hostio.destroy(d2)
    print("GPU_FUZZ %s (%d cases in %.0f s, %d mismatches)" % ("PASS" if bad == 0 else "FAIL", done, time.time() - t0, bad))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()

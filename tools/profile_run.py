#!/usr/bin/env python
"""One device-resident pass of the hot path over the bench workload, for ncu (see profiles/README.md).
usage: python tools/profile_run.py [n_genomes=100] [passes=1]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pangene_b200 import engine, hostio  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
opt = hostio.default_opt()
d = hostio.synth(opt, hostio.synth_params(3, n_genome=n, n_genome_total=n))
L = engine.lib()
engine.set_verbose(1)
s = L.pgx_upload(d)
L.pgx_set_sweep_stat(1 if os.environ.get("PGX_SWEEP_STAT") else 0)
per_pass = []
for _ in range(passes):
    L.pgx_run(s, C.byref(opt), 3)
    t = engine.stats()
    per_pass.append("%.2f+%.2f" % (t.run_ms[0], t.run_ms[1]))
st = engine.stats()
if passes > 1:
    print("per pass (stage 1 + stage 2 ms):", " ".join(per_pass), " redo:", st.redo)
print("hits %d active %d seg %d arc %d launches %d run_ms %.2f %.2f sweep_ms %.3f sweep_GB %.3f" % (st.n_hit, st.n_active, st.n_seg, st.n_arc, st.launches, st.run_ms[0], st.run_ms[1], st.sweep_ms, st.sweep_bytes / 1e9))
L.pgx_session_free(s)

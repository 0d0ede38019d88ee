#!/usr/bin/env python
"""device-resident passes with per-stage timing (PGX_TIMING=1 prints the host-side pieces)"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pangene_b200 import engine, hostio
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
opt = hostio.default_opt()
d = hostio.synth(opt, hostio.synth_params(3, n_genome=n, n_genome_total=n))
L = engine.lib(); engine.set_verbose(1)
s = L.pgx_upload(d)
for it in range(4):
    t0 = time.perf_counter(); L.pgx_run(s, C.byref(opt), 1); t1 = time.perf_counter(); L.pgx_run(s, C.byref(opt), 2); t2 = time.perf_counter()
    st = engine.stats()
    print("pass %d: stage1 wall %.2f ms (dev %.2f)  stage2 wall %.2f ms (dev %.2f)  launches %d" % (it, (t1 - t0) * 1e3, st.run_ms[0], (t2 - t1) * 1e3, st.run_ms[1], st.launches), flush=True)
L.pgx_session_free(s)

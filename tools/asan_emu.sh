#!/bin/bash
# TEST INFRASTRUCTURE: a CPU-side "memcheck" of the device code.  tests/emu/build.py::build_full(sanitize=True) compiles all of libpgx (C ABI, engine,
# every kernel) for the host on the SIMT shim under AddressSanitizer + UBSan; this script drives the whole hot path through it -- the reference's
# C4 example against its goldens, an isoform-rich synthetic set against the oracle on the sync-free rounds and again on the careful rounds, and
# the GFA route (walk tokens) -- so that an out-of-bounds index into a column, a table or a shared-memory array in ANY kernel is a sanitizer report.
# usage: tools/asan_emu.sh        (needs gcc with libasan; ~4 min)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SO=$(python -c "import sys; sys.path.insert(0, '$ROOT/tests'); from emu import build; print(build.build_full(True))")
cd $ROOT/tests
run() {
  LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1 PGX_LIB=$SO "$@" python - <<'PY'
import os, sys, tempfile
sys.path.insert(0, "..")
import common
from oracle import pgoracle
from pangene_b200 import engine, hostio
engine.set_verbose(1)
opt = hostio.default_opt()
mode = os.environ.get("PGX_ROUNDS", "fast")
with tempfile.TemporaryDirectory() as td:
    files = common.c4_files(td)
    d = hostio.read_pafs(opt, files)
    g = engine.hot_path(opt, d)
    assert common.gfa_bytes(g) == open(os.path.join(common.GOLDEN, "C4.gfa"), "rb").read(), "C4 GFA"
    assert common.bed_bytes(d) == open(os.path.join(common.GOLDEN, "C4.bed_flag"), "rb").read(), "C4 BED"
    print("[%s] C4: GFA and BED byte-identical to the goldens" % mode)
p = hostio.synth_params(2, n_genome=3, n_gene=200, iso_per_gene=4, hits_per_genome=1500, n_genome_total=3)
d1, d2, d3 = hostio.synth(opt, p), hostio.synth(opt, p), hostio.synth(opt, p)
g1, g2 = pgoracle.hot_path(opt, d1), engine.hot_path(opt, d2, lazy=True)
a, b = common.gfa_bytes(g1), common.gfa_bytes(g2)
assert sorted(a.splitlines()) == sorted(b.splitlines()) and sorted(common.bed_bytes(d1).splitlines()) == sorted(common.bed_bytes(d2).splitlines())
st = engine.stats()
print("[%s] synthetic: %d hits, %d segments, %d arcs, %d launches, %d fast rounds, %d redo: GFA and BED lines equal to the oracle's" % (mode, st.n_hit, st.n_seg, st.n_arc, st.launches, st.fast_rounds, st.redo))
g, tok, off, n = engine.graph_and_tokens(opt, d3)
with tempfile.NamedTemporaryFile(suffix=".gfa") as t:
    hostio.write_gfa_tokens(g, tok, off, t.name, True, True)
    c = open(t.name, "rb").read()
engine.release_tokens(d3, tok)
assert sorted(c.splitlines()) == sorted(a.splitlines()), "token route"
print("[%s] GFA route: %d walk tokens, GFA lines equal to the oracle's" % (mode, n))
PY
}
run env
run env PGX_ROUNDS=careful
echo "ASAN_EMU PASS"

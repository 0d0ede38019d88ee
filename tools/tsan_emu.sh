#!/bin/bash
# TEST INFRASTRUCTURE: a CPU-side "racecheck" of ALL the device code: libpgx's own source compiled for the host on the SIMT shim under ThreadSanitizer
# (tests/emu/build.py::build_full("thread")); CUDA threads are OS threads there and __syncthreads() / the warp collectives are real barriers, so two
# accesses of a CTA to one shared- or global-memory word that no barrier orders are a reported race.  Expected reports (benign by construction, also on
# the device): threads storing the SAME value to one byte (k_chain_mark), and the optimistic plain read of a hash slot that the compare-and-swap
# right after it validates (arc_slot / ovf_group_add).      usage: tools/tsan_emu.sh      (~2 min)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SO=$(python -c "import sys; sys.path.insert(0, '$ROOT/tests'); from emu import build; print(build.build_full('thread'))" 2>/dev/null | tail -n 1)
cd $ROOT/tests
LD_PRELOAD=$(gcc -print-file-name=libtsan.so) TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=4 exitcode=0" PGX_LIB=$SO python - 2> /tmp/tsan_emu.err <<'PY'
import sys
sys.path.insert(0, "..")
import common
from oracle import pgoracle
from pangene_b200 import engine, hostio
engine.set_verbose(1)
opt = hostio.default_opt()
p = hostio.synth_params(2, n_genome=2, n_gene=120, iso_per_gene=4, hits_per_genome=900, n_genome_total=2)
d1, d2 = hostio.synth(opt, p), hostio.synth(opt, p)
g1, g2 = pgoracle.hot_path(opt, d1), engine.hot_path(opt, d2, lazy=True)
a, b = common.gfa_bytes(g1), common.gfa_bytes(g2)
assert sorted(a.splitlines()) == sorted(b.splitlines()) and sorted(common.bed_bytes(d1).splitlines()) == sorted(common.bed_bytes(d2).splitlines())
st = engine.stats()
print("whole hot path under ThreadSanitizer: %d hits, %d launches, GFA and BED lines equal to the oracle's" % (st.n_hit, st.launches))
PY
echo "ThreadSanitizer reports: $(grep -c 'WARNING: ThreadSanitizer' /tmp/tsan_emu.err || true)"
grep -A3 "WARNING: ThreadSanitizer" /tmp/tsan_emu.err | grep "#0" | sed 's/(libpgx.*//' | sort | uniq -c
UNEXPECTED=$(grep -A3 "WARNING: ThreadSanitizer" /tmp/tsan_emu.err | grep "#0" | grep -v "k_chain_mark\|arc_slot\|ovf_group_add\|atomic" | wc -l)
if [ "$UNEXPECTED" != "0" ]; then echo "TSAN_EMU FAIL: unexpected report(s)"; exit 1; fi
echo "TSAN_EMU PASS"

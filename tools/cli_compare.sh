#!/bin/bash
# Whole-CLI comparison on the same PAF files (BASELINE.md §3 "second line"): the stock reference binary (oracle/_ref/pangene,
# 1 thread) vs our `pangene` (parallel reader + libpgx + writers).  usage: tools/cli_compare.sh [n_genomes=16]
set -e -o pipefail
N=${1:-16}
D=$(mktemp -d)
python - "$N" "$D" <<'PY'
import sys
sys.path.insert(0, ".")
from pangene_b200 import hostio
n, d = int(sys.argv[1]), sys.argv[2]
opt = hostio.default_opt()
hostio.synth_write_pafs(opt, hostio.synth_params(3, n_genome=n, n_genome_total=n), d)
PY
du -sh $D | cut -f1
TIMEFORMAT="%R s wall (%U user)"
echo -n "reference CLI (stock pangene, 1 thread): "; time oracle/_ref/pangene $D/*.paf > $D/ref.gfa 2> $D/ref.log
echo -n "pangene-b200 CLI (first call: CUDA context + pinned ring): "; time pangene_b200/csrc/pangene -v1 $D/*.paf > $D/ours.gfa
echo -n "pangene-b200 CLI (again): "; time pangene_b200/csrc/pangene -v3 $D/*.paf 2> $D/ours.log > $D/ours2.gfa
grep -E "parse|Real time" $D/ours.log || true
if cmp -s <(LC_ALL=C sort $D/ref.gfa) <(LC_ALL=C sort $D/ours.gfa); then echo "GFA identical (sorted): $(wc -l < $D/ref.gfa) lines"; else echo "GFA DIFFERS"; fi
rm -rf $D

#!/bin/bash
# TEST INFRASTRUCTURE: the host-side code (libpgio reader/writers/generator) and the oracle under AddressSanitizer + UBSan + LeakSanitizer:
#   pipeline   generator -> PAF text -> parallel + sequential reader -> oracle -> GFA/BED writers -> gfa2matrix
#   tie_cases  campaign cases in which the oracle's sort-chain replays are active (DESIGN.md "Tie order")
#   odd_lines  hand-made pathological PAF lines (empty, truncated, 16-byte boundaries of the SSE2 split, NUL, CRLF, 20-digit numbers, ...)
# usage: tools/asan_host.sh     (needs gcc/g++ with libasan; scratch files under /tmp/asan)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p /tmp/asan && cd /tmp/asan
SAN="-O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer"
INC="-I$ROOT/include -I$ROOT/oracle -I$ROOT/pangene_b200/csrc"
gcc -std=gnu99 $SAN -ffp-contract=off $INC -c $ROOT/oracle/pg_oracle.c -o pg_oracle.o
for t in pipeline tie_cases odd_lines; do
	g++ -std=c++17 $SAN $INC $ROOT/tools/asan/$t.cpp $ROOT/pangene_b200/csrc/pgio.cpp $ROOT/pangene_b200/csrc/pgio_synth.cpp pg_oracle.o -o $t -lz -lpthread -lm
	echo "== $t"; ASAN_OPTIONS=detect_leaks=1 ./$t
done
echo "ASAN_HOST PASS"

#!/usr/bin/env python
"""TEST INFRASTRUCTURE: randomized campaign on REAL alignments -- the reference's own test/C4 example (tests/golden/C4_inputs.txt,
33 genomes), perturbed: random subsets and orders of the genomes, random line drops / duplications / score edits inside the files,
random option mixes.  Every case goes through the stock reference CLI (oracle/_ref/pangene) and through our reader + the oracle +
our writers; GFA (as sorted lines) and `--bed=flag` (as sorted lines) must be identical.   usage: tools/fuzz_c4.py [first] [n]"""
import os
import random
import re
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
import refharness  # noqa: E402
from oracle import pgoracle  # noqa: E402
from pangene_b200 import hostio  # noqa: E402

OPTS = [[], ["-S"], ["-J"], ["-E"], ["-F"], ["--ori-sc"], ["-T0"], ["-T3"], ["-p0"], ["-p0.2"], ["-a2"], ["-f0.2"], ["-f0.9"], ["-c2"], ["-g4"], ["-r1"], ["-b0.01"], ["-y0.3"], ["-B0.9"]]


def perturb(src, dst, rnd):
    lines = open(src, "rb").read().splitlines(keepends=True)
    mode = rnd.choice(["asis", "asis", "drop", "dup", "score", "shuffle"])
    out = []
    for l in lines:
        r = rnd.random()
        if mode == "drop" and r < 0.15:
            continue
        if mode == "score" and r < 0.2:
            l = re.sub(rb"ms:i:(\d+)", lambda m: b"ms:i:%d" % max(1, int(m.group(1)) + rnd.randint(-40, 40)), l)
        out.append(l)
        if mode == "dup" and r < 0.05:
            out.append(l)
    if mode == "shuffle":
        rnd.shuffle(out)
    open(dst, "wb").write(b"".join(out))


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    bad = 0
    base = tempfile.mkdtemp()
    src_files = common.c4_files(base)
    for k in range(first, first + n):
        rnd = random.Random(424200 + k)
        files = rnd.sample(src_files, rnd.randint(2, len(src_files)))
        args = sorted(set(sum(rnd.sample(OPTS, rnd.randint(0, 3)), [])))
        args = [a for a in args if not (a.startswith("-T") and sum(x.startswith("-T") for x in args) > 1 and a != "-T0")]
        args = [a for a in args if not (a.startswith("-p") and sum(x.startswith("-p") for x in args) > 1 and a != "-p0")]
        args = [a for a in args if not (a.startswith("-f") and sum(x.startswith("-f") for x in args) > 1 and a != "-f0.2")]
        d = tempfile.mkdtemp(dir=base)
        mine = []
        for f in files:
            g = os.path.join(d, os.path.basename(f))
            perturb(f, g, rnd)
            mine.append(g)
        ref_gfa = refharness.ref_cli(mine, args)
        ref_bed = refharness.ref_cli(mine, args + ["--bed=flag"])
        opt = common.opt_from_args(args)
        dd = hostio.read_pafs(opt, mine)
        g = pgoracle.hot_path(opt, dd)
        a, b = common.gfa_bytes(g), common.bed_bytes(dd)
        ok_gfa = sorted(a.splitlines()) == sorted(ref_gfa.splitlines())
        ok_bed = sorted(b.splitlines()) == sorted(ref_bed.splitlines())
        tie = pgoracle.tie_report()
        status = "ok" if ok_gfa and ok_bed else "MISMATCH"
        bad += status != "ok"
        print("case %d %s genomes=%d hits=%d tie=%s gfa=%s bed=%s args=%s" % (k, status, len(mine), hostio.total_hits(dd), tie, ok_gfa, ok_bed, args), flush=True)
        if not ok_gfa:
            print(common.first_diff(ref_gfa, a), flush=True)
        elif not ok_bed:
            print(common.first_diff(ref_bed, b), flush=True)
        pgoracle.destroy(g); hostio.destroy(dd)
        shutil.rmtree(d)
    shutil.rmtree(base)
    print("C4_FUZZ %s (%d cases, %d mismatches)" % ("PASS" if bad == 0 else "FAIL", n, bad))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()

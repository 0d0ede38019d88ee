#!/usr/bin/env python
"""TEST INFRASTRUCTURE: randomized campaign of the oracle restatement against the compiled, unmodified reference
(oracle/_ref/libpgref.so) -- many more shapes, seeds and option mixes than the CPU test suite has time for.
Needs /root/reference (this container only).   usage: tools/fuzz_oracle.py [first_case] [n_cases] [local|cmtie]
"local" runs every case with local_dist = 1 and local_count = case % 4 (pangene -D/-C): the locality test of branch.c:31-46 is
then decided by survivor ordinals, which makes the reference's tie order among survivors with one (contig, cs) observable.
Prints one line per case; a mismatch that comes with a non-zero tie detector (DESIGN.md "tie order") is reported as TIE."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
import refharness  # noqa: E402
from oracle import pgoracle  # noqa: E402
from pangene_b200 import hostio  # noqa: E402

FLAGS = ["-S", "-J", "-E", "-F", "--ori-sc"]


def random_case(k, local=False, cmtie=False):
    rnd = random.Random(777000 + k)
    args = [f for f in FLAGS if rnd.random() < 0.2]
    if rnd.random() < 0.5: args.append("-T%d" % rnd.choice([0, 1, 2, 3, 7, 15]))
    if rnd.random() < 0.3: args.append("-p%s" % rnd.choice(["0", "0.01", "0.2", "0.5"]))
    if rnd.random() < 0.3: args.append("-a%d" % rnd.choice([1, 2, 3]))
    if rnd.random() < 0.3: args.append("-f%s" % rnd.choice(["0.1", "0.2", "0.5", "0.9"]))
    if rnd.random() < 0.2: args.append("-c%d" % rnd.choice([1, 2, 4, 20]))
    if rnd.random() < 0.2: args.append("-g%d" % rnd.choice([2, 4, 6, 30]))
    if rnd.random() < 0.2: args.append("-r%d" % rnd.choice([0, 1, 3]))
    if rnd.random() < 0.2: args.append("-b%s" % rnd.choice(["0.01", "0.1", "0.3"]))
    if rnd.random() < 0.2: args.append("-y%s" % rnd.choice(["0.1", "0.3", "0.8"]))
    if rnd.random() < 0.2: args.append("-B%s" % rnd.choice(["0.5", "0.9", "0.99"]))
    if rnd.random() < 0.25:
        case = dict(config=5, n_genome=rnd.randint(3, 80), n_gene=rnd.randint(100, 1200), iso_per_gene=1, args=args)
        case["hits_per_genome"] = int(case["n_gene"] * rnd.uniform(0.5, 1.1))
    else:
        case = dict(config=2, n_genome=rnd.randint(1, 9), n_gene=rnd.randint(50, 2000), iso_per_gene=rnd.choice([1, 4]), args=args)
        case["hits_per_genome"] = int(case["n_gene"] * case["iso_per_gene"] * rnd.uniform(1.2, 3.5))
    case["seed"] = rnd.getrandbits(40)
    case["n_ctg"] = rnd.choice([1, 2, 5, 24])
    if local:
        case["local_dist"], case["local_count"] = 1, k % 4
    if cmtie:
        case["cmtie"] = True
        case["n_gene"] = min(case["n_gene"], 600)       # (the tie injection walks every hit in Python)
        case["hits_per_genome"] = min(case["hits_per_genome"], 2500)
    return case


def force_cm_ties(d, seed):
    """"cmtie" mode: give hits that follow a non-overlapping hit on their contig the SAME mid-point (pg_hit_t::cm) as that hit.  Both usually
    survive, so the arc between them (graph.c:106-122) and their place in the W line (format.c:203-208) depend on how the reference's unstable
    sort chain orders two equal (contig, cm) keys -- the case DESIGN.md "tie order" calls tie[0]."""
    rnd = random.Random(seed)
    dd = d.contents
    n_forced = 0
    for j in range(dd.n_genome):
        g = dd.genome[j]
        by_ctg = {}
        for i in range(g.n_hit):
            by_ctg.setdefault(g.hit[i].cid, []).append(i)
        for idx in by_ctg.values():
            idx.sort(key=lambda i: (g.hit[i].cs, i))
            k = 1
            while k < len(idx):
                a, b = g.hit[idx[k - 1]], g.hit[idx[k]]
                if a.ce <= b.cs and rnd.random() < 0.08:
                    b.cm = a.cm
                    n_forced += 1
                    k += 2
                else:
                    k += 1
    return n_forced


def data_for(case):
    opt = common.opt_from_args(case["args"])
    if "local_dist" in case:
        opt.local_dist, opt.local_count = case["local_dist"], case["local_count"]
    p = hostio.synth_params(case["config"], n_genome=case["n_genome"], n_gene=case["n_gene"], iso_per_gene=case["iso_per_gene"],
                            hits_per_genome=case["hits_per_genome"], n_genome_total=case["n_genome"], seed=case["seed"], n_ctg=case["n_ctg"])
    d = hostio.synth(opt, p)
    if case.get("cmtie"):
        force_cm_ties(d, case["seed"])
    return opt, d


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    local = len(sys.argv) > 3 and sys.argv[3] == "local"
    cmtie = len(sys.argv) > 3 and sys.argv[3] == "cmtie"
    rh = refharness.RefHot()
    bad = 0
    for k in range(first, first + n):
        case = random_case(k, local, cmtie)
        opt, d1 = data_for(case)
        _, d2 = data_for(case)
        g1, _ = rh.hot_path(opt, d1)
        g2 = pgoracle.hot_path(opt, d2)
        tie = pgoracle.tie_report()
        a, b = common.gfa_bytes(g1), common.gfa_bytes(g2)
        ok_gfa = sorted(a.splitlines()) == sorted(b.splitlines())
        ok_graph = common.graph_tuple(g1) == common.graph_tuple(g2)
        ok_bed = sorted(common.bed_bytes(d1).splitlines()) == sorted(common.bed_bytes(d2).splitlines())
        ok = ok_gfa and ok_graph and ok_bed
        status = "ok" if ok else ("TIE" if ((tie[0] and not cmtie) or tie[2]) else "MISMATCH")
        bad += status == "MISMATCH"
        print("case %d %s hits=%d S=%d L=%d tie=%s gfa=%s graph=%s bed=%s %s" % (k, status, hostio.total_hits(d1), g1.contents.n_seg, g1.contents.n_arc, tie, ok_gfa, ok_graph, ok_bed,
              {x: case[x] for x in ("config", "n_genome", "n_gene", "iso_per_gene", "hits_per_genome", "n_ctg", "seed", "args", "local_count") if x in case}), flush=True)
        if status == "MISMATCH":
            print(common.first_diff(a, b), flush=True)
        rh.destroy(g1); pgoracle.destroy(g2); hostio.destroy(d1); hostio.destroy(d2)
    print("FUZZ %s (%d cases, %d mismatches)" % ("PASS" if bad == 0 else "FAIL", n, bad))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()

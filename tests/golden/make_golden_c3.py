#!/usr/bin/env python
"""Golden for the HEADLINE configuration (BASELINE.json configs[2]/[3]: synthetic 100 genomes x 200k hits, 80k proteins /
20k genes, 24 contigs, seed 3, default options) -> tests/golden/c3_golden.json.   Run in the build container (needs
oracle/_ref, i.e. /root/reference at build time); takes ~10 min of one core and ~6 GB of disk under --tmp.

Two independent routes through the UNMODIFIED reference, which must agree:
  (1) the stock binary: pgio_synth_write_paf() writes the 100 PAF files, `oracle/_ref/pangene *.paf` prints the GFA;
  (2) the reference's objects in-process (oracle/_ref/libpgref.so driven by oracle/ref_hot.c) on the in-memory pg_data_t that
      bench.py uploads, printed by the reference's own writers (format.c).
What is committed (a checksum of checksums, cheap to verify at bench time at any rank count):
  sl_md5    md5 of the S and L lines, in the reference's order (segments by gene id, arcs by key)
  w_md5     md5 of the W lines, in genome order (rank r of a sharded run holds a contiguous slice of them)
  gfa_sorted_md5, n_S, n_L, n_W, n_hit
"""
import argparse, hashlib, json, os, subprocess, sys, tempfile, time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.path.join(ROOT, "oracle", "_ref", "pangene")


def digest(gfa):
    lines = gfa.splitlines(keepends=True)
    sl = b"".join(l for l in lines if l[:1] in (b"S", b"L"))
    w = b"".join(l for l in lines if l[:1] == b"W")
    assert len(sl) + len(w) == len(gfa), "unexpected record type in the GFA"
    return dict(sl_md5=hashlib.md5(sl).hexdigest(), w_md5=hashlib.md5(w).hexdigest(),
                gfa_sorted_md5=hashlib.md5(b"\n".join(sorted(gfa.splitlines()))).hexdigest(),
                n_S=sum(l[:1] == b"S" for l in lines), n_L=sum(l[:1] == b"L" for l in lines), n_W=sum(l[:1] == b"W" for l in lines))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=100)
    ap.add_argument("--tmp", default="/tmp")
    ap.add_argument("--route", default="both", choices=["both", "cli", "inproc"])
    ap.add_argument("--out", default=os.path.join(HERE, "c3_golden.json"))
    a = ap.parse_args()
    from pangene_b200 import hostio
    import refharness
    opt = hostio.default_opt()
    p = hostio.synth_params(3, n_genome=a.genomes, n_genome_total=a.genomes)
    res = {}
    if a.route in ("both", "cli"):
        with tempfile.TemporaryDirectory(dir=a.tmp) as td:
            t0 = time.time()
            files = hostio.synth_write_pafs(opt, p, td)
            t1 = time.time()
            gfa = subprocess.run([REF] + files, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
            print("stock binary: PAF written in %.0f s, pangene ran %.0f s, %d bytes of GFA" % (t1 - t0, time.time() - t1, len(gfa)), flush=True)
        res["cli"] = digest(gfa)
        del gfa
    if a.route in ("both", "inproc"):
        d = hostio.synth(opt, p)
        n_hit = hostio.total_hits(d)
        rh = refharness.RefHot()
        t0 = time.time()
        g, t = rh.hot_path(opt, d)
        print("in-process reference: hot path %.0f s (block %.1f, post %.1f, graph %.1f)" % (time.time() - t0, t[0], t[1], t[2]), flush=True)
        with tempfile.NamedTemporaryFile(dir=a.tmp) as tf:
            rh.write(g, d, 0, tf.name)
            res["inproc"] = digest(open(tf.name, "rb").read())
        res["inproc"]["n_hit"] = n_hit
        res["inproc"]["ref_hot_path_s"] = t
        rh.destroy(g)
    if a.route == "both":
        for k in ("sl_md5", "w_md5", "gfa_sorted_md5", "n_S", "n_L", "n_W"):
            assert res["cli"][k] == res["inproc"][k], "the two reference routes disagree on " + k
    gold = dict(res.get("inproc") or res["cli"])
    gold.update(workload="pgio_synth config 3 (seed 3): %d genomes x 200k hits, 80k proteins / 20k genes, 24 contigs, default options" % a.genomes,
                genomes=a.genomes, routes=sorted(res.keys()), routes_agree=(a.route == "both"))
    json.dump(gold, open(a.out, "w"), indent=1, sort_keys=True)
    print(json.dumps(gold, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Regenerate tests/golden/ from the reference (run in the build container, where /root/reference exists).

  C4_inputs.txt   the 33 PAFs of the reference's README example (test/C4/*.paf.gz, README.md:13), decompressed
                  and concatenated; a line ">name" starts a file.  Data fixture, not source.
  C4.gfa, C4.bed_raw, C4.bed_flag, C4.bed_walk, C4_w.gfa
                  stdout of the UNMODIFIED reference binary (oracle/_ref/pangene) on those files with
                  no option / --bed=raw / --bed=flag / --bed / -w.  md5s match SURVEY.md §4.
  synth_golden.json
                  for a few small synthetic inputs (pgio_synth parameters + reference options): md5 of the
                  LC_ALL=C-sorted reference GFA and of the sorted --bed=flag output, plus S/L/W counts.
"""
import glob, gzip, hashlib, json, os, subprocess, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "pangene")

SYNTH_CASES = [
    dict(name="euk_iso4", config=2, n_genome=5, n_gene=1500, iso_per_gene=4, hits_per_genome=15000, args=[]),
    dict(name="euk_iso1", config=2, n_genome=8, n_gene=3000, iso_per_gene=1, hits_per_genome=7500, args=[]),
    dict(name="euk_iso4_S_F", config=2, n_genome=5, n_gene=1500, iso_per_gene=4, hits_per_genome=15000, args=["-S", "-F"]),
    dict(name="euk_iso4_J_T3", config=2, n_genome=5, n_gene=1500, iso_per_gene=4, hits_per_genome=15000, args=["-J", "-T3"]),
    dict(name="bact", config=5, n_genome=40, n_gene=1200, iso_per_gene=1, hits_per_genome=1000, args=["-p0", "-a2"]),
]


def run(args, files):
    return subprocess.run([REF] + args + files, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout


def main():
    files = sorted(glob.glob("/root/reference/test/C4/*.paf.gz"))
    with open(os.path.join(HERE, "C4_inputs.txt"), "wb") as out:
        for f in files:
            out.write(b">" + os.path.basename(f)[:-3].encode() + b"\n")
            out.write(gzip.open(f).read())
    for name, args in [("C4.gfa", []), ("C4.bed_raw", ["--bed=raw"]), ("C4.bed_flag", ["--bed=flag"]), ("C4.bed_walk", ["--bed"]), ("C4_w.gfa", ["-w"])]:
        data = run(args, files)
        open(os.path.join(HERE, name), "wb").write(data)
        print(name, hashlib.md5(data).hexdigest())
    from pangene_b200 import hostio
    res = {}
    for c in SYNTH_CASES:
        opt = hostio.default_opt()
        p = hostio.synth_params(c["config"], n_genome=c["n_genome"], n_gene=c["n_gene"], iso_per_gene=c["iso_per_gene"],
                                hits_per_genome=c["hits_per_genome"], n_genome_total=c["n_genome"])
        with tempfile.TemporaryDirectory() as td:
            fs = hostio.synth_write_pafs(opt, p, td)
            gfa = run(c["args"], fs)
            bed = run(c["args"] + ["--bed=flag"], fs)
        ls = gfa.splitlines()
        res[c["name"]] = dict(c, gfa_sorted_md5=hashlib.md5(b"\n".join(sorted(ls))).hexdigest(),
                              bed_flag_sorted_md5=hashlib.md5(b"\n".join(sorted(bed.splitlines()))).hexdigest(),
                              n_S=sum(l[:1] == b"S" for l in ls), n_L=sum(l[:1] == b"L" for l in ls), n_W=sum(l[:1] == b"W" for l in ls))
        print(c["name"], res[c["name"]]["gfa_sorted_md5"], res[c["name"]]["n_S"], res[c["name"]]["n_L"], res[c["name"]]["n_W"])
    json.dump(res, open(os.path.join(HERE, "synth_golden.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

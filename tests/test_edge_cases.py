"""Edge cases of the path: empty and ragged inputs, duplicates/collisions, negative scores, name-set options.
Hand-built miniprot-dialect PAFs go through (a) the stock reference CLI, (b) our reader + the oracle, and -- with a GPU --
(c) our reader + libpgx.  All three must print the same GFA/BED."""
import os
import random

import pytest

import common
import refharness
from oracle import pgoracle
from pangene_b200 import abi, hostio

needs_ref = pytest.mark.skipif(not refharness.available(), reason="oracle/_ref not built")


def line(q, qlen, qs, qe, strand, ctg, clen, cs, cigar, ms, mlen=None, fs=0, st=0):
    span = 0
    num = ""
    for ch in cigar:
        if ch.isdigit():
            num += ch
        else:
            n = int(num); num = ""
            span += n * 3 if ch in "MXD=" else n if ch in "NUVFG" else 0
    blen = (qe - qs) * 3
    mlen = blen if mlen is None else mlen
    return "\t".join(map(str, [q, qlen, qs, qe, strand, ctg, clen, cs, cs + span, mlen, blen, 0, "AS:i:%d" % ms, "ms:i:%d" % ms, "np:i:%d" % (qe - qs),
                               "fs:i:%d" % fs, "st:i:%d" % st, "cg:Z:" + cigar])) + "\n"


def build_cases(tmp):
    """returns {name: (files, cli_args)}"""
    cases = {}
    d = os.path.join(str(tmp), "ragged"); os.makedirs(d)
    a = os.path.join(d, "A.paf"); b = os.path.join(d, "EMPTY.paf"); c = os.path.join(d, "C.paf"); e = os.path.join(d, "ALLBAD.paf")
    with open(a, "w") as f:
        f.write(line("G1:P1", 100, 0, 100, "+", "s1#0#c1", 100000, 1000, "50M200N50M", 500))
        f.write(line("G1:P1", 100, 0, 100, "+", "s1#0#c1", 100000, 1000, "50M200N50M", 500))        # exact duplicate hit: same key, ranks decide
        f.write(line("G1:P1b", 90, 0, 90, "+", "s1#0#c1", 100000, 1003, "40M209N50M", 430))
        f.write(line("G2:P2", 80, 0, 80, "-", "s1#0#c1", 100000, 1100, "80M", 300))                  # single exon inside G1's intron/exon
        f.write(line("G3:P3", 60, 0, 60, "+", "s1#0#c1", 100000, 5000, "60M", 280))
        f.write(line("G3:P3", 60, 5, 55, "+", "s1#0#c2", 5000, 100, "50M", 120, mlen=100))           # second contig with ONE hit, low identity
        f.write(line("G4:P4", 70, 0, 70, "+", "s1#0#c1", 100000, 5100, "30M100N40M", -50))           # negative score
        f.write(line("G5:P5", 50, 0, 50, "+", "s1#0#c1", 100000, 9000, "50M", 0))                    # zero score
        f.write(line("G6:P6", 200, 0, 80, "+", "s1#0#c1", 100000, 9500, "80M", 300))                 # dropped by -l (40 % aligned)
        f.write("\n")                                                                                # empty line
        f.write("G7:P7\t10\n")                                                                       # truncated line
    open(b, "w").close()
    with open(c, "w") as f:
        f.write(line("G1:P1", 100, 0, 100, "-", "s2#1#x", 50000, 2000, "50M150N50M", 480))
        f.write(line("G3:P3", 60, 0, 60, "-", "s2#1#x", 50000, 2100, "60M", 300))                    # G3 inside G1 on the other strand
        f.write(line("G2:P2", 80, 0, 80, "+", "nolabel", 40000, 300, "80M", 310))                    # contig name without sample#hap
        f.write(line("G8:P8", 40, 0, 40, "+", "s2#1#x", 50000, 2000, "40M", 200))                    # same start as G1's hit
    with open(e, "w") as f:
        f.write(line("G1:P1", 100, 0, 20, "+", "s3#0#c1", 1000, 10, "20M", 50))                      # every line dropped by the reader
    cases["ragged"] = ([a, b, c, e], [])
    cases["ragged_S_p0"] = ([a, b, c, e], ["-S", "-p0"])
    cases["ragged_T0_J"] = ([a, b, c, e], ["-T0", "-J"])
    cases["single_file"] = ([c], ["-p0"])
    cases["empty_only"] = ([b], [])
    # reader corner cases (tools/asan/odd_lines.cpp runs the same lines under ASan/UBSan): fields ending on the 16-byte boundaries of
    # the SSE2 tab split, a NUL inside a line (the reference stops the line there), CRLF, a 20-digit score, frameshift CIGAR
    # operators, very long names, a last line without newline
    o = os.path.join(d, "ODD.paf")
    good = line("G1:P1", 100, 0, 100, "+", "s1#0#c1", 100000, 1000, "50M200N50M", 500)
    with open(o, "wb") as f:
        f.write(good.encode())
        f.write(b"0123456789abcdef\n0123456789abcde\tx\n0123456789abcdef\t\n\t\t\t\t\t\t\t\t\t\t\t\t\t\t\t\t\t\t\n")
        f.write(good[:40].encode() + b"\0" + good[41:].encode())
        f.write(good.rstrip("\n").encode() + b"\r\n")
        f.write(line("G3:P3", 50, 0, 50, "+", "ctg", 1000, 10, "50M", 0).replace("ms:i:0", "ms:i:99999999999999999999").encode())
        f.write(line("G4:" + "x" * 5000, 50, 0, 50, "+", "c" * 3000, 1000, 10, "50M", 77).encode())
        f.write(line("G5:P5", 60, 0, 60, "+", "s1#0#c1", 100000, 5000, "10M1F20M2G28M", 280, fs=2).encode())
        f.write(line("G6:P6", 80, 0, 80, "-", "s1#0#c1", 100000, 7000, "20M100U20M100V40M", 300).rstrip("\n").encode())
    cases["odd_lines"] = ([a, o], [])
    return cases


def run_ours(hot_path, destroy, files, args):
    opt = common.opt_from_args(args)
    d = hostio.read_pafs(opt, files)
    g = hot_path(opt, d)
    out = (common.gfa_bytes(g), common.bed_bytes(d))
    destroy(g); hostio.destroy(d)
    return out


@needs_ref
@pytest.mark.parametrize("name", ["ragged", "ragged_S_p0", "ragged_T0_J", "single_file", "empty_only", "odd_lines"])
def test_oracle_matches_stock_cli_on_edge_cases(tmp_path, name):
    files, args = build_cases(tmp_path)[name]
    gfa, bed = run_ours(pgoracle.hot_path, pgoracle.destroy, files, args)
    assert gfa == refharness.ref_cli(files, args)
    assert sorted(bed.splitlines()) == sorted(refharness.ref_cli(files, args + ["--bed=flag"]).splitlines())


@needs_ref
def test_name_sets_included_preferred_excluded(tmp_path):
    """-I / -P / -X are reader-side (read.c:142-148) but change vertex election (vertex.c:55,69) -- our reader + oracle == stock CLI"""
    case = dict(config=2, n_genome=4, n_gene=400, iso_per_gene=4, hits_per_genome=3000)
    opt0 = hostio.default_opt()
    p = hostio.synth_params(2, n_genome=4, n_gene=400, iso_per_gene=4, hits_per_genome=3000, n_genome_total=4)
    files = hostio.synth_write_pafs(opt0, p, str(tmp_path / "paf"))
    incl, pref, excl = "G000003,G000010", "G000005,G000006,G000007", "G000001,G000002:P000002.2"
    opt = hostio.default_opt()
    L = hostio.lib()
    opt.incl, opt.preferred, opt.excl = L.pgio_name_set(incl.encode()), L.pgio_name_set(pref.encode()), L.pgio_name_set(excl.encode())
    d = hostio.read_pafs(opt, files)
    g = pgoracle.hot_path(opt, d)
    assert common.gfa_bytes(g) == refharness.ref_cli(files, ["-I", incl, "-P", pref, "-X", excl])
    pgoracle.destroy(g); hostio.destroy(d)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ragged", "ragged_S_p0", "ragged_T0_J", "single_file", "empty_only"])
def test_gpu_matches_oracle_on_edge_cases(tmp_path, name):
    from pangene_b200 import engine
    engine.set_verbose(1)
    files, args = build_cases(tmp_path)[name]
    a = run_ours(pgoracle.hot_path, pgoracle.destroy, files, args)
    b = run_ours(lambda o, d: engine.hot_path(o, d, lazy=(name != "ragged")), engine.destroy, files, args)
    assert a[0] == b[0]
    assert sorted(a[1].splitlines()) == sorted(b[1].splitlines())


@pytest.mark.gpu
def test_gpu_name_sets(tmp_path):
    from pangene_b200 import engine
    engine.set_verbose(1)
    opt0 = hostio.default_opt()
    p = hostio.synth_params(2, n_genome=4, n_gene=400, iso_per_gene=4, hits_per_genome=3000, n_genome_total=4)
    files = hostio.synth_write_pafs(opt0, p, str(tmp_path / "paf"))
    outs = []
    for hp, ds in [(pgoracle.hot_path, pgoracle.destroy), (engine.hot_path, engine.destroy)]:
        opt = hostio.default_opt()
        L = hostio.lib()
        opt.incl, opt.preferred = L.pgio_name_set(b"G000003,G000010"), L.pgio_name_set(b"G000005,G000006,G000007")
        d = hostio.read_pafs(opt, files)
        g = hp(opt, d)
        outs.append(common.gfa_bytes(g))
        ds(g); hostio.destroy(d)
    assert outs[0] == outs[1]


@pytest.mark.gpu
def test_gpu_rerun_same_session_is_idempotent():
    """pgx_run twice on one upload gives the same graph (bench.py relies on it)"""
    import ctypes as C
    from pangene_b200 import engine
    engine.set_verbose(1)
    opt = hostio.default_opt()
    p = hostio.synth_params(2, n_genome=3, n_gene=1000, iso_per_gene=4, hits_per_genome=9000, n_genome_total=3)
    d = hostio.synth(opt, p)
    L = engine.lib()
    s = L.pgx_upload(d)
    res = []
    for _ in range(2):
        L.pgx_run(s, C.byref(opt), 3)
        d2 = hostio.lib().pgio_data_clone(d)
        g = L.pg_graph_init(d2)
        L.pgx_download(s, d2, g)
        res.append((common.gfa_bytes(g), sorted(common.bed_bytes(d2).splitlines())))
        L.pg_graph_destroy(g); hostio.lib().pgio_data_clone_destroy(d2)
    L.pgx_session_free(s)
    assert res[0] == res[1]
    d1 = hostio.synth(opt, p)
    g1 = pgoracle.hot_path(opt, d1)
    assert res[0][0] == common.gfa_bytes(g1)

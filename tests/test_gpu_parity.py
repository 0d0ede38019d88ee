"""GPU suite, part 2: the CUDA hot path, called through the C ABI (libpgx.so), against
   (a) the committed goldens printed by the unmodified reference binary,
   (b) the oracle restatement (libpgoracle.so) on the same inputs,
   (c) the compiled reference itself in-process (oracle/_ref/libpgref.so), which travels with the snapshot.
Bit-exact: GFA lines as sorted sets (S/L/W "modulo record order"), BED lines as sorted sets (every per-hit flag),
and the seg/arc structs element by element."""
import os
import random

import pytest

import common
import refharness
from oracle import pgoracle
from pangene_b200 import abi, engine, hostio

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _quiet():
    engine.set_verbose(1)


def check_against(opt, make, ref_kind="oracle", lazy=False):
    d_ref, d_gpu = make(), make()
    if ref_kind == "reference":
        rh = refharness.RefHot()
        g_ref, _ = rh.hot_path(opt, d_ref)
    else:
        g_ref = pgoracle.hot_path(opt, d_ref)
    g_gpu = engine.hot_path(opt, d_gpu, lazy=lazy)
    a, b = common.gfa_bytes(g_ref), common.gfa_bytes(g_gpu)
    assert sorted(a.splitlines()) == sorted(b.splitlines()), common.first_diff(a, b)
    ba, bb = common.bed_bytes(d_ref), common.bed_bytes(d_gpu)
    assert sorted(ba.splitlines()) == sorted(bb.splitlines()), common.first_diff(ba, bb)
    assert common.graph_tuple(g_ref) == common.graph_tuple(g_gpu)
    q1, q2 = g_ref.contents, g_gpu.contents
    assert [q1.g2s[i] for i in range(d_ref.contents.n_gene)] == [q2.g2s[i] for i in range(d_gpu.contents.n_gene)]
    assert [q1.idx[i] for i in range(2 * q1.n_seg)] == [q2.idx[i] for i in range(2 * q2.n_seg)]
    for i in range(d_ref.contents.n_prot):
        p1, p2 = d_ref.contents.prot[i], d_gpu.contents.prot[i]
        assert (p1.n, p1.avg_score_adj, p1.max_score_ori, p1.rep) == (p2.n, p2.avg_score_adj, p2.max_score_ori, p2.rep)
    st = engine.stats()
    engine.destroy(g_gpu)
    (refharness.RefHot().destroy if ref_kind == "reference" else pgoracle.destroy)(g_ref)
    hostio.destroy(d_ref); hostio.destroy(d_gpu)
    return st, a


def test_c4_golden(tmp_path):
    files = common.c4_files(tmp_path)
    opt = hostio.default_opt()
    d = hostio.read_pafs(opt, files)
    g = engine.hot_path(opt, d)
    assert common.gfa_bytes(g) == open(os.path.join(common.GOLDEN, "C4.gfa"), "rb").read()
    assert common.bed_bytes(d) == open(os.path.join(common.GOLDEN, "C4.bed_flag"), "rb").read()
    assert common.bed_bytes(d, True) == open(os.path.join(common.GOLDEN, "C4.bed_walk"), "rb").read()
    engine.destroy(g); hostio.destroy(d)


def test_c4_bed_raw_after_post_process(tmp_path):
    opt = hostio.default_opt()
    d = hostio.read_pafs(opt, common.c4_files(tmp_path))
    engine.hot_path(opt, d, stop_after_post=True)
    assert common.bed_bytes(d) == open(os.path.join(common.GOLDEN, "C4.bed_raw"), "rb").read()
    hostio.destroy(d)


@pytest.mark.parametrize("name", sorted(common.synth_golden().keys()))
def test_synthetic_goldens(name):
    case = common.synth_golden()[name]
    opt, d = common.synth_case_data(case)
    g = engine.hot_path(opt, d, lazy=True)
    assert common.sorted_md5(common.gfa_bytes(g)) == case["gfa_sorted_md5"]
    assert common.sorted_md5(common.bed_bytes(d)) == case["bed_flag_sorted_md5"]
    engine.destroy(g); hostio.destroy(d)


OPTION_SETS = [[], ["-S"], ["-J"], ["-E"], ["-F", "-T2"], ["-T0"], ["--ori-sc"], ["-p0", "-a2"], ["-f0.2", "-c4", "-g6"], ["-b0.1", "-y0.3", "-B0.9", "-r1"]]


@pytest.mark.parametrize("k", range(len(OPTION_SETS)))
def test_randomized_vs_oracle(k):
    rnd = random.Random(5000 + k)
    case = dict(config=2, n_genome=rnd.randint(2, 9), n_gene=rnd.randint(300, 2500), iso_per_gene=rnd.choice([1, 4]), hits_per_genome=0, args=OPTION_SETS[k])
    case["hits_per_genome"] = int(case["n_gene"] * case["iso_per_gene"] * rnd.uniform(1.5, 3.0))
    opt = common.opt_from_args(case["args"])
    check_against(opt, lambda: common.synth_case_data(case)[1], "oracle", lazy=bool(k & 1))


@pytest.mark.parametrize("k", [177, 179, 225, 463, 502])
def test_first_hit_tie_cases(k):
    """Genomes whose first (contig, cs) is shared by several hits: the front of the reference's array -- the one hit pg_shadow never
    un-shadows (overlap.c:108) -- moves with the reference's unstable sort chain; libpgx replays the chain on the host for those
    contigs (Session::prepare_first_replay).  Campaign cases of tools/fuzz_oracle.py on which a stable-order build differs."""
    import sys
    sys.path.insert(0, os.path.join(common.ROOT, "tools"))
    import fuzz_oracle
    case = fuzz_oracle.random_case(k)
    opt = common.opt_from_args(case["args"])
    kind = "reference" if refharness.available() else "oracle"
    check_against(opt, lambda: fuzz_oracle.data_for(case)[1], kind, lazy=bool(k & 1))


@pytest.mark.parametrize("k", [1006, 2573, 176])
def test_tied_survivor_ordinal_cases(k):
    """local_dist = 1, local_count = k % 4: the locality test of branch.c:31-46 is decided by survivor ordinals, and survivors that
    share one (contig, cs) are numbered in the order the reference's sort chain left them in (branch.c:18-25).  On 1006 and 2573 a
    stable-order build differs from the reference; on 176 the stable order IS the reference's (a regression guard for the replay:
    a tied hit that does not own its gene's table entry must not be renumbered).  libpgx: k_cs_tie_mark, Session::fix_tied_ordinals."""
    import sys
    sys.path.insert(0, os.path.join(common.ROOT, "tools"))
    import fuzz_oracle
    case = fuzz_oracle.random_case(k, local=True)
    kind = "reference" if refharness.available() else "oracle"
    check_against(fuzz_oracle.data_for(case)[0], lambda: fuzz_oracle.data_for(case)[1], kind, lazy=bool(k & 1))


@pytest.mark.parametrize("k", [16, 97, 110, 115])
def test_cm_tied_survivor_cases(k):
    """Survivors that share one (contig, cm) (tools/fuzz_oracle.py "cmtie" mode forces them): the reference walks them in the order its
    unstable sort chain left them in, and the arc between two of them -- with everything the branch rounds derive from it -- depends on that
    order (graph.c:103-122, branch.c:116-136).  On case 16 a stable-order build differs from the reference.  libpgx: the fast rounds raise the
    fallback on such a pair, the careful path replays the chain (Session::fix_cm_tied_survivors)."""
    import sys
    sys.path.insert(0, os.path.join(common.ROOT, "tools"))
    import fuzz_oracle
    case = fuzz_oracle.random_case(k, cmtie=True)
    kind = "reference" if refharness.available() else "oracle"
    st, _ = check_against(fuzz_oracle.data_for(case)[0], lambda: fuzz_oracle.data_for(case)[1], kind, lazy=bool(k & 1))
    assert st.tie[0] > 0 and st.redo > 0


@pytest.mark.parametrize("k", [0, 9, 17, 25])
def test_perturbed_real_alignments(k, tmp_path):
    """Cases of tools/fuzz_c4.py: the reference's test/C4 alignments with genomes subsampled and reordered, lines dropped,
    duplicated (exact duplicate hits: equal keys, equal coordinates), re-scored or shuffled, under random option mixes."""
    import random
    import sys
    sys.path.insert(0, os.path.join(common.ROOT, "tools"))
    import fuzz_c4
    rnd = random.Random(424200 + k)
    src = common.c4_files(tmp_path)
    files = rnd.sample(src, rnd.randint(2, len(src)))
    args = sorted(set(sum(rnd.sample(fuzz_c4.OPTS, rnd.randint(0, 3)), [])))
    args = [a for a in args if not (a[:2] in ("-T", "-p", "-f") and sum(x[:2] == a[:2] for x in args) > 1 and a not in ("-T0", "-p0", "-f0.2"))]
    mine = []
    os.makedirs(str(tmp_path / "p"))
    for f in files:
        g = str(tmp_path / "p" / os.path.basename(f))
        fuzz_c4.perturb(f, g, rnd)
        mine.append(g)
    opt = common.opt_from_args(args)
    check_against(opt, lambda: hostio.read_pafs(opt, mine), "oracle", lazy=bool(k & 1))


@pytest.mark.skipif(not refharness.available(), reason="oracle/_ref did not travel")
@pytest.mark.parametrize("cfg", ["euk", "bact"])
def test_vs_compiled_reference(cfg):
    if cfg == "euk":
        case = dict(config=2, n_genome=6, n_gene=4000, iso_per_gene=4, hits_per_genome=40000, args=[])
    else:
        case = dict(config=5, n_genome=150, n_gene=2000, iso_per_gene=1, hits_per_genome=1700, args=["-p0", "-a2"])
    opt = common.opt_from_args(case["args"])
    check_against(opt, lambda: common.synth_case_data(case)[1], "reference")


def test_config2_bit_exact_and_matrix(tmp_path):
    """BASELINE.json configs[1]: 10 genomes x 50k hits, 20k proteins, 1 GPU, bit-exact GFA (vs the oracle)"""
    case = dict(config=2, n_genome=10, n_gene=20000, iso_per_gene=1, hits_per_genome=50000, args=[])
    opt = common.opt_from_args([])
    st, gfa = check_against(opt, lambda: common.synth_case_data(case)[1], "oracle", lazy=True)
    assert st.n_hit > 450000 and st.n_seg > 16000
    assert gfa.count(b"\nS\t") + gfa.startswith(b"S\t") >= 0.8 * 20000   # generator non-degeneracy


def test_full_size_properties():
    """configs[2] shape at reduced genome count (the oracle needs minutes at 100 genomes): size-independent properties
    of the GFA: arc mirror symmetry, ng <= nc, sortedness, walks consistent with segments."""
    opt = hostio.default_opt()
    p = hostio.synth_params(3, n_genome=12, n_genome_total=12)
    d = hostio.synth(opt, p)
    g = engine.hot_path(opt, d, lazy=True)
    segs, arcs = common.graph_tuple(g)
    amap = {a[0]: a for a in arcs}
    assert [a[0] for a in arcs] == sorted(amap)                      # sorted, unique keys
    for x, ng, nc, ad, s1, s2 in arcs:
        v, w = x >> 32, x & 0xffffffff
        m = amap[((w ^ 1) << 32) | (v ^ 1)]                          # graph.c:117-119: every arc has its mirror
        assert (m[1], m[2], m[3], m[4], m[5]) == (ng, nc, ad, s2, s1)
        assert 1 <= ng <= nc and ng <= 12
    assert [s[0] for s in segs] == sorted(s[0] for s in segs)
    assert all(s[3] <= s[4] for s in segs)
    # S-line nc == number of unfiltered hits of that gene over all genomes (graph.c:125-126 after the last round)
    cnt = {}
    dd = d.contents
    for j in range(dd.n_genome):
        gg = dd.genome[j]
        for i in range(gg.n_hit):
            h = gg.hit[i]
            if not (h.bits & 2):
                gid = dd.prot[h.pid].gid
                cnt[gid] = cnt.get(gid, 0) + 1
    assert {s[0]: s[4] for s in segs if s[4]} == cnt
    engine.destroy(g); hostio.destroy(d)


def test_headline_config_matches_the_reference_golden():
    """BASELINE.json configs[2] at FULL size (100 genomes x 200k hits, the configuration bench.py quotes its numbers on): the S/L
    lines and the W lines of the GPU path, hashed, against tests/golden/c3_golden.json -- printed by the unmodified reference (stock
    binary on the PAF files AND its objects in-process; tests/golden/make_golden_c3.py, ~6 min of one core).  Same gate as bench.py."""
    import hashlib
    import json
    gold = json.load(open(os.path.join(common.GOLDEN, "c3_golden.json")))
    opt = hostio.default_opt()
    d = hostio.synth(opt, hostio.synth_params(3, n_genome=100, n_genome_total=100))
    assert hostio.total_hits(d) == gold["n_hit"]
    g = engine.hot_path(opt, d, lazy=True)
    sl, w = hostio.gfa_digest(g)
    assert (g.contents.n_seg, g.contents.n_arc, w.count(b"\n")) == (gold["n_S"], gold["n_L"], gold["n_W"])
    assert sl == gold["sl_md5"]
    assert hashlib.md5(w).hexdigest() == gold["w_md5"]
    engine.destroy(g); hostio.destroy(d)


def test_tie_detector_reports_zero_on_goldens():
    case = common.synth_golden()["euk_iso4"]
    opt, d = common.synth_case_data(case)
    g = engine.hot_path(opt, d, lazy=True)
    st = engine.stats()
    assert st.tie[0] == 0 and st.tie[2] == 0
    engine.destroy(g); hostio.destroy(d)


def test_cli_binary_prints_the_golden_gfa(tmp_path):
    """`pangene *.paf > out.gfa` (our main(), parallel reader, libpgx, our writers) == the stock binary's output"""
    import subprocess
    files = common.c4_files(tmp_path)
    exe = os.path.join(common.ROOT, "pangene_b200", "csrc", "pangene")
    out = subprocess.run([exe, "-v1"] + files, check=True, stdout=subprocess.PIPE).stdout
    assert out == open(os.path.join(common.GOLDEN, "C4.gfa"), "rb").read()
    out = subprocess.run([exe, "-v1", "--bed=raw"] + files, check=True, stdout=subprocess.PIPE).stdout
    assert out == open(os.path.join(common.GOLDEN, "C4.bed_raw"), "rb").read()
    out = subprocess.run([exe, "-v1", "-w"] + files, check=True, stdout=subprocess.PIPE).stdout
    assert out == open(os.path.join(common.GOLDEN, "C4_w.gfa"), "rb").read()


DROPIN = os.path.join(common.ROOT, "oracle", "_ref", "pangene_pgx")


@pytest.mark.skipif(not (refharness.available() and os.path.exists(DROPIN)), reason="oracle/_ref/pangene_pgx did not travel")
def test_reference_host_linked_against_libpgx(tmp_path):
    """THE drop-in: oracle/_ref/pangene_pgx is the reference's OWN main(), reader, dictionary, option parser and writers (main.c sys.c dict.c
    option.c format.c read.c hit.c, compiled from the reference tree by oracle/Makefile) linked against libpgx.so instead of graph.o vertex.o
    branch.o overlap.o, with the two edits of INTEGRATION.md (read.c:243-260 compiled out, option.c:4 dropped).  It must print what the stock
    binary prints: on the reference's C4 example (goldens) and on synthetic PAF files (stock binary run next to it)."""
    import subprocess
    files = common.c4_files(tmp_path)
    run = lambda exe, args, fs: subprocess.run([exe, "-v1"] + args + fs, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    assert run(DROPIN, [], files) == open(os.path.join(common.GOLDEN, "C4.gfa"), "rb").read()
    assert run(DROPIN, ["-w"], files) == open(os.path.join(common.GOLDEN, "C4_w.gfa"), "rb").read()
    for flag, gold in (("--bed=raw", "C4.bed_raw"), ("--bed=flag", "C4.bed_flag"), ("--bed", "C4.bed_walk")):   # BED line order follows the reference's unstable sort chain: compare as sets
        assert sorted(run(DROPIN, [flag], files).splitlines()) == sorted(open(os.path.join(common.GOLDEN, gold), "rb").read().splitlines()), flag
    opt = hostio.default_opt()
    p = hostio.synth_params(2, n_genome=6, n_gene=1500, iso_per_gene=4, hits_per_genome=12000, n_genome_total=6)
    fs = hostio.synth_write_pafs(opt, p, str(tmp_path / "paf"))
    for args in ([], ["-S", "-T3"], ["-p0", "-a2"]):
        a, b = run(refharness.REF_BIN, args, fs), run(DROPIN, args, fs)
        assert sorted(a.splitlines()) == sorted(b.splitlines()), (args, common.first_diff(a, b))


def test_config5_bacterial_with_matrix(tmp_path):
    """BASELINE.json configs[4] shape (bacterial, -p0 -a2) at 300 of the 1000 genomes: bit-exact GFA vs the oracle and the
    same presence/absence matrix (gfa2matrix, pangene.js:1168-1246) from both GFAs"""
    opt = common.opt_from_args(["-p0", "-a2"])
    p = hostio.synth_params(5, n_genome=300, n_genome_total=300)
    d1, d2 = hostio.synth(opt, p), hostio.synth(opt, p)
    g1 = pgoracle.hot_path(opt, d1)
    g2 = engine.hot_path(opt, d2, lazy=True)
    f1, f2 = str(tmp_path / "cpu.gfa"), str(tmp_path / "gpu.gfa")
    hostio.write_gfa(g1, f1); hostio.write_gfa(g2, f2)
    a, b = open(f1, "rb").read(), open(f2, "rb").read()
    assert sorted(a.splitlines()) == sorted(b.splitlines()), common.first_diff(a, b)
    m1, m2 = str(tmp_path / "cpu.tsv"), str(tmp_path / "gpu.tsv")
    n1 = hostio.lib().pgio_gfa2matrix(os.fsencode(f1), os.fsencode(m1), 0)
    n2 = hostio.lib().pgio_gfa2matrix(os.fsencode(f2), os.fsencode(m2), 0)
    assert n1 == n2 > 2000 and open(m1, "rb").read() == open(m2, "rb").read()
    engine.destroy(g2); pgoracle.destroy(g1); hostio.destroy(d1); hostio.destroy(d2)


def test_walk_tokens_route(tmp_path):
    """the GFA and the gfa2matrix table from the DEVICE-RESIDENT final state (lazy mode 2: graph + 12-byte walk tokens,
    no pg_hit_t record comes back).  C4: the golden GFA byte for byte; synthetic sets incl. the bacterial -p0 -a2 shape: the oracle's GFA,
    the tokens the oracle's hits give (tests/common.tokens_from_hits) and the table pgio_gfa2matrix computes from the oracle's GFA text.
    pgx_fetch_hits() afterwards still materialises every per-hit flag."""
    opt = hostio.default_opt()
    d = hostio.read_pafs(opt, common.c4_files(tmp_path))
    g, tok, off, n = engine.graph_and_tokens(opt, d)
    assert common.gfa_bytes_tokens(g, tok, off) == open(os.path.join(common.GOLDEN, "C4.gfa"), "rb").read()
    assert common.gfa_bytes_tokens(g, tok, off, walk=False) == open(os.path.join(common.GOLDEN, "C4_w.gfa"), "rb").read()
    engine.lib().pgx_fetch_hits(d)
    assert common.bed_bytes(d) == open(os.path.join(common.GOLDEN, "C4.bed_flag"), "rb").read()
    engine.release_tokens(d, tok); engine.destroy(g); hostio.destroy(d)
    cases = [([], hostio.synth_params(2, n_genome=7, n_gene=3000, iso_per_gene=4, hits_per_genome=30000)),
             (["-S", "-T3"], hostio.synth_params(2, n_genome=5, n_gene=2000, iso_per_gene=1, hits_per_genome=9000)),
             (["-p0", "-a2"], hostio.synth_params(5, n_genome=200, n_genome_total=200))]
    for args, p in cases:
        o = common.opt_from_args(args)
        d1, d2 = hostio.synth(o, p), hostio.synth(o, p)
        g1 = pgoracle.hot_path(o, d1)
        g2, tok, off, n = engine.graph_and_tokens(o, d2)
        t1, off1, n1 = common.tokens_from_hits(d1)
        assert n == n1 and list(off) == list(off1)
        want, got = common.gfa_bytes(g1), common.gfa_bytes_tokens(g2, tok, off)
        assert sorted(got.splitlines()) == sorted(want.splitlines()), common.first_diff(want, got)
        assert common.gfa_bytes_tokens(g1, t1, off1) == want
        for cn in (False, True):
            assert common.matrix_bytes_tokens(g2, tok, off, cn) == common.matrix_bytes_gfa(want, cn)
        engine.lib().pgx_fetch_hits(d2)
        assert sorted(common.bed_bytes(d2).splitlines()) == sorted(common.bed_bytes(d1).splitlines())
        engine.release_tokens(d2, tok); engine.destroy(g2); pgoracle.destroy(g1); hostio.destroy(d1); hostio.destroy(d2)


def test_recycled_hit_arrays_keep_results_intact():
    """the download hands out recycled (already mapped) arrays instead of fresh ones: back-to-back runs over data sets of different
    shapes -- arrays of one run become destinations of the next -- must each still equal the oracle"""
    opt = hostio.default_opt()
    for seed, (ng, hp) in enumerate([(6, 30000), (9, 12000), (4, 50000), (9, 12000)]):
        p = hostio.synth_params(2, n_genome=ng, n_gene=2500, iso_per_gene=4, hits_per_genome=hp, seed=11 + seed)
        check_against(opt, lambda: hostio.synth(opt, p), "oracle", lazy=bool(seed & 1))


def test_packed_reader_route(tmp_path):
    """SURVEY §8f row 1 end to end: PAF text -> device-ready records (pgio_read_pafs_packed) -> pgx_post_process_packed -> graph + walk tokens ->
    GFA, without a single pg_hit_t record.  C4: the golden GFA byte for byte; synthetic PAF files: the oracle's GFA on the pg_hit_t route."""
    import ctypes as C
    opt = hostio.default_opt()
    d, pk = hostio.read_pafs_packed(opt, common.c4_files(tmp_path))
    g, tok, off, n = engine.graph_and_tokens(opt, d, packed=pk)
    assert common.gfa_bytes_tokens(g, tok, off) == open(os.path.join(common.GOLDEN, "C4.gfa"), "rb").read()
    engine.release_tokens(d, tok); engine.destroy(g); hostio.lib().pgio_packed_destroy(C.byref(pk)); hostio.destroy(d)
    for args, p in [([], hostio.synth_params(2, n_genome=6, n_gene=1500, iso_per_gene=4, hits_per_genome=12000)),
                    (["-p0", "-a2"], hostio.synth_params(5, n_genome=60, n_gene=800, iso_per_gene=1, hits_per_genome=700))]:
        o = common.opt_from_args(args)
        files = hostio.synth_write_pafs(o, p, str(tmp_path / ("paf%d" % p.shape)))
        d1 = hostio.read_pafs(o, files)
        g1 = pgoracle.hot_path(o, d1)
        want = common.gfa_bytes(g1)
        d2, pk = hostio.read_pafs_packed(o, files)
        g2, tok, off, n = engine.graph_and_tokens(o, d2, packed=pk)
        got = common.gfa_bytes_tokens(g2, tok, off)
        assert sorted(got.splitlines()) == sorted(want.splitlines()), common.first_diff(want, got)
        engine.release_tokens(d2, tok); engine.destroy(g2); hostio.lib().pgio_packed_destroy(C.byref(pk)); hostio.destroy(d2)
        pgoracle.destroy(g1); hostio.destroy(d1)


_VARIANT_SCRIPT = r"""
import hashlib, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import common
from pangene_b200 import engine, hostio
engine.set_verbose(1)
opt = hostio.default_opt()
out = []
for cfg, kw in ((3, dict(n_genome=5, n_gene=1500, iso_per_gene=4, hits_per_genome=30000, n_genome_total=5)), (2, dict(n_genome=6, n_gene=1200, iso_per_gene=1, hits_per_genome=9000, n_genome_total=6))):
    d = hostio.synth(opt, hostio.synth_params(cfg, **kw))
    g = engine.hot_path(opt, d)
    out.append(hashlib.md5(b"\n".join(sorted(common.gfa_bytes(g).splitlines()))).hexdigest())
    out.append(hashlib.md5(b"\n".join(sorted(common.bed_bytes(d).splitlines()))).hexdigest())
print("VARIANT_RESULT", " ".join(out))
"""


def test_sweep_kernel_variants_agree_with_the_oracle():
    """k_sweep_tiled<MODE, VAR> for VAR 0 ... 3 (PGX_SWEEP_VARIANT; one process each: the choice is read once) and the plain thread-per-hit
    kernel (PGX_SWEEP=global) give the oracle's GFA lines and per-hit flag lines on an isoform-rich and a single-isoform set."""
    import hashlib
    import subprocess
    import sys
    opt = hostio.default_opt()
    want = []
    for cfg, kw in ((3, dict(n_genome=5, n_gene=1500, iso_per_gene=4, hits_per_genome=30000, n_genome_total=5)), (2, dict(n_genome=6, n_gene=1200, iso_per_gene=1, hits_per_genome=9000, n_genome_total=6))):
        d = hostio.synth(opt, hostio.synth_params(cfg, **kw))
        g = pgoracle.hot_path(opt, d)
        want.append(hashlib.md5(b"\n".join(sorted(common.gfa_bytes(g).splitlines()))).hexdigest())
        want.append(hashlib.md5(b"\n".join(sorted(common.bed_bytes(d).splitlines()))).hexdigest())
        pgoracle.destroy(g); hostio.destroy(d)
    script = _VARIANT_SCRIPT % dict(root=common.ROOT, tests=os.path.join(common.ROOT, "tests"))
    for env_extra in ({"PGX_SWEEP_VARIANT": "0"}, {"PGX_SWEEP_VARIANT": "1"}, {"PGX_SWEEP_VARIANT": "2"}, {"PGX_SWEEP_VARIANT": "3"}, {"PGX_SWEEP": "global"}):
        env = dict(os.environ); env.pop("PGX_SWEEP", None); env.pop("PGX_SWEEP_VARIANT", None); env.update(env_extra)
        r = subprocess.run([sys.executable, "-c", script], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        assert r.returncode == 0, (env_extra, r.stderr[-2000:])
        got = [l for l in r.stdout.splitlines() if l.startswith("VARIANT_RESULT")][-1].split()[1:]
        assert got == want, (env_extra, got, want)

"""CPU suite, part 2: host I/O (reader, generator, writers, gfa2matrix) and the C-ABI surface of libpgx."""
import ctypes as C
import os
import re
import subprocess

import pytest

import common
import refharness
from pangene_b200 import abi, hostio

ROOT = common.ROOT


def test_reader_plus_generator_round_trip(tmp_path):
    """PAF text of the synthetic records parses back to exactly the in-memory pg_data_t (every byte of every hit)"""
    opt = hostio.default_opt()
    for cfg, kw in [(2, dict(n_genome=3, n_gene=800, iso_per_gene=4, hits_per_genome=8000)), (5, dict(n_genome=12, n_gene=500, iso_per_gene=1, hits_per_genome=400))]:
        p = hostio.synth_params(cfg, **kw)
        d = hostio.synth(opt, p)
        files = hostio.synth_write_pafs(opt, p, str(tmp_path / ("c%d" % cfg)))
        d2 = hostio.read_pafs(opt, files)
        assert (d.contents.n_genome, d.contents.n_gene, d.contents.n_prot) == (d2.contents.n_genome, d2.contents.n_gene, d2.contents.n_prot)
        for j in range(d.contents.n_genome):
            a, b = d.contents.genome[j], d2.contents.genome[j]
            assert (a.n_hit, a.n_exon, a.n_ctg) == (b.n_hit, b.n_exon, b.n_ctg)
            assert C.string_at(a.hit, a.n_hit * 88) == C.string_at(b.hit, b.n_hit * 88)
            assert C.string_at(a.exon, a.n_exon * 8) == C.string_at(b.exon, b.n_exon * 8)
            assert [a.ctg[c].name for c in range(a.n_ctg)] == [b.ctg[c].name for c in range(b.n_ctg)]
            assert a.label == b.label
        for i in range(d.contents.n_prot):
            assert (d.contents.prot[i].name, d.contents.prot[i].len, d.contents.prot[i].gid) == (d2.contents.prot[i].name, d2.contents.prot[i].len, d2.contents.prot[i].gid)
        hostio.destroy(d); hostio.destroy(d2)


def test_generator_is_deterministic_and_shardable():
    opt = hostio.default_opt()
    whole = hostio.synth(opt, hostio.synth_params(2, n_genome=4, n_gene=500, iso_per_gene=1, hits_per_genome=1200, n_genome_total=4))
    part = hostio.synth(opt, hostio.synth_params(2, n_genome=2, genome_begin=2, n_gene=500, iso_per_gene=1, hits_per_genome=1200, n_genome_total=4))
    for j in range(2):
        a, b = whole.contents.genome[2 + j], part.contents.genome[j]
        assert a.n_hit == b.n_hit and C.string_at(a.hit, a.n_hit * 88) == C.string_at(b.hit, b.n_hit * 88)
    hostio.destroy(whole); hostio.destroy(part)


@pytest.mark.skipif(not refharness.available(), reason="oracle/_ref not built")
def test_reader_and_writers_against_reference_cli(tmp_path):
    """our reader + the reference's own hot path + our writers == the stock CLI, byte for byte"""
    files = common.c4_files(tmp_path)
    opt = hostio.default_opt()
    d = hostio.read_pafs(opt, files)
    rh = refharness.RefHot()
    g, _ = rh.hot_path(opt, d)
    assert common.gfa_bytes(g) == refharness.ref_cli(files)
    assert common.bed_bytes(d) == refharness.ref_cli(files, ["--bed=flag"])
    assert common.bed_bytes(d, True) == refharness.ref_cli(files, ["--bed"])
    rh.destroy(g); hostio.destroy(d)


def test_gfa2matrix(tmp_path):
    gfa = os.path.join(common.GOLDEN, "C4.gfa")
    out = str(tmp_path / "m.tsv")
    n = hostio.lib().pgio_gfa2matrix(os.fsencode(gfa), os.fsencode(out), 0)
    rows = [l.rstrip("\n").split("\t") for l in open(out)]
    assert n == 7 and rows[0][0] == "Gene" and len(rows) == 8
    n_asm = len(rows[0]) - 1
    assert all(len(r) == n_asm + 1 and set(r[1:]) <= {"0", "1"} for r in rows[1:])
    # presence must agree with a direct count of the W lines
    walks = [l.split("\t") for l in open(gfa) if l.startswith("W")]
    asm = []
    for w in walks:
        a = w[1] + "#" + w[2]
        if a not in asm: asm.append(a)
    assert rows[0][1:] == asm
    for r in rows[1:]:
        for k, a in enumerate(asm):
            present = any((w[1] + "#" + w[2]) == a and re.search(r"[><]%s(?=[><\s]|$)" % re.escape(r[0]), w[6]) for w in walks)
            assert (r[1 + k] == "1") == bool(present)


def test_token_writers_equal_the_writers_on_materialised_hits(tmp_path):
    """SURVEY §8f rows 2-3, host half: W lines and the gfa2matrix table formatted from walk tokens are byte-identical to the GFA written from
    the pg_hit_t records / the table computed from that GFA's text (tokens built here from the oracle's hits; the device builds the same
    tokens in tests/test_gpu_parity.py)."""
    from oracle import pgoracle
    opt = hostio.default_opt()
    cases = [hostio.read_pafs(opt, common.c4_files(tmp_path)),
             hostio.synth(opt, hostio.synth_params(2, n_genome=5, n_gene=1500, iso_per_gene=4, hits_per_genome=15000)),
             hostio.synth(opt, hostio.synth_params(5, n_genome=40, n_gene=600, iso_per_gene=1, hits_per_genome=500))]
    for k, d in enumerate(cases):
        o = opt if k < 2 else common.opt_from_args(["-p0", "-a2"])
        g = pgoracle.hot_path(o, d)
        tok, off, n = common.tokens_from_hits(d)
        assert n > 0
        want = common.gfa_bytes(g)
        assert common.gfa_bytes_tokens(g, tok, off) == want
        for cn in (False, True):
            assert common.matrix_bytes_tokens(g, tok, off, cn) == common.matrix_bytes_gfa(want, cn)
        pgoracle.destroy(g); hostio.destroy(d)


def test_packed_reader_emits_what_pg_post_process_would_pack(tmp_path):
    """pgio_read_pafs_packed() writes the device-ready 48-byte records itself.  Field by field they must be what
    libpgx packs out of the pg_hit_t records of pgio_read_pafs() (pgx_engine.cuh upload_hits), and every other table must be the same."""
    import numpy as np
    opt = hostio.default_opt()
    sets = [common.c4_files(tmp_path),
            hostio.synth_write_pafs(opt, hostio.synth_params(2, n_genome=5, n_gene=900, iso_per_gene=4, hits_per_genome=7000), str(tmp_path / "e")),
            hostio.synth_write_pafs(opt, hostio.synth_params(5, n_genome=9, n_gene=400, iso_per_gene=1, hits_per_genome=300), str(tmp_path / "b"))]
    HIT = np.dtype([(n, "<i4") for n in ("pid", "qs", "qe", "cid", "mlen", "blen", "lof", "rank", "score_ori", "score_adj", "score_dom", "n_exon", "off_exon", "pid_dom", "pid_dom0")]
                   + [("bits", "<u4"), ("cs", "<i8"), ("cm", "<i8"), ("ce", "<i8")])
    PK = np.dtype([(n, "<i4") for n in ("pid", "cid", "rank", "score_ori", "score_adj", "n_exon", "off_exon", "rev", "cs", "cm", "ce", "lof")])
    for files in sets:
        for threads in (1, 0):
            d = hostio.read_pafs(opt, files, threads=0)
            d2, pk = hostio.read_pafs_packed(opt, files, threads=threads)
            a, b = d.contents, d2.contents
            assert (a.n_genome, a.n_gene, a.n_prot) == (b.n_genome, b.n_gene, b.n_prot) and pk.n_genome == a.n_genome
            assert hostio.lib().pgio_packed_total_hits(C.byref(pk)) == hostio.total_hits(d) > 0
            assert [(a.prot[i].name, a.prot[i].gid, a.prot[i].len) for i in range(a.n_prot)] == [(b.prot[i].name, b.prot[i].gid, b.prot[i].len) for i in range(b.n_prot)]
            assert [(a.gene[i].name, a.gene[i].bits) for i in range(a.n_gene)] == [(b.gene[i].name, b.gene[i].bits) for i in range(b.n_gene)]
            for j in range(a.n_genome):
                ga, gb = a.genome[j], b.genome[j]
                assert gb.n_hit == 0 and not gb.hit and pk.n_hit[j] == ga.n_hit and ga.n_exon == gb.n_exon and ga.label == gb.label
                assert [(ga.ctg[c].name, ga.ctg[c].len) for c in range(ga.n_ctg)] == [(gb.ctg[c].name, gb.ctg[c].len) for c in range(gb.n_ctg)]
                assert C.string_at(ga.exon, ga.n_exon * 8) == C.string_at(gb.exon, gb.n_exon * 8)
                h = np.frombuffer(C.string_at(ga.hit, ga.n_hit * 88), dtype=HIT)
                p = np.frombuffer(C.string_at(pk.hit[j], ga.n_hit * 48), dtype=PK)
                for f in ("pid", "cid", "rank", "score_ori", "score_adj", "n_exon", "off_exon", "lof", "cs", "cm", "ce"):
                    assert (h[f] == p[f]).all(), f
                assert ((h["bits"] & 1) == p["rev"]).all()
            hostio.lib().pgio_packed_destroy(C.byref(pk))
            hostio.destroy(d); hostio.destroy(d2)


def test_libpgx_exports_every_declared_symbol():
    """the C-ABI library loads without a GPU and exports every function include/pgx.h declares (no compute call)"""
    hdr = open(os.path.join(ROOT, "include", "pgx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(pgx?_[a-z0-9_]+)\s*\(", hdr))
    assert {"pg_post_process", "pg_graph_init", "pg_graph_gen", "pg_graph_destroy", "pgx_upload", "pgx_run", "pgx_download", "pgx_comm_init"} <= names
    L = C.CDLL(os.path.join(ROOT, "pangene_b200", "csrc", "libpgx.so"))
    for n in sorted(names):
        assert hasattr(L, n), n
    assert C.c_int.in_dll(L, "pg_verbose").value == 3
    # built for sm_100a and nothing else
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "pangene_b200", "csrc", "libpgx.so")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT).stdout.decode()
    if "sm_" in out:
        assert set(re.findall(r"sm_\d+a?", out)) == {"sm_100a"}, out


def test_host_side_radix_replay_of_libpgx_matches_oracle():
    import random
    from oracle import pgoracle
    from pangene_b200 import engine
    rnd = random.Random(7)
    for n, bits in [(50, 4), (3000, 5), (3000, 64), (70000, 17)]:
        A = (C.c_uint64 * (2 * n))()
        for i in range(n):
            A[2 * i] = rnd.getrandbits(bits) << rnd.choice([0, 8, 31]) & (2 ** 64 - 1); A[2 * i + 1] = i
        B = (C.c_uint64 * (2 * n))(*A)
        engine.lib().pgx_test_flag_sort_replay(C.byref(A), n)
        pgoracle.lib().pgo_radix_replay128(C.byref(B), n)
        assert list(A) == list(B)


def test_representatives_shortcut_equals_the_replayed_sort():
    """hit.c:205-217: the representative of a gene is its protein with the largest z; libpgx replays the reference's unstable sort only when
    two proteins of one gene share that maximum -- both routes must agree wherever the shortcut applies, and ties must take the replay"""
    import numpy as np
    from pangene_b200 import engine
    L = engine.lib()
    rng = np.random.default_rng(4)
    for n_gene, iso, ties in [(500, 4, False), (3000, 3, False), (20000, 4, False), (800, 4, True)]:
        n_prot = n_gene * iso
        gid = np.repeat(np.arange(n_gene, dtype=np.int32), iso)
        perm = rng.permutation(n_prot); gid = gid[perm]
        score = rng.integers(0, 1 << 22, n_prot).astype(np.uint64); cnt = rng.integers(1, 100, n_prot).astype(np.uint64)
        z = (score << np.uint64(32)) | cnt
        z[rng.random(n_prot) < 0.3] = 0                       # proteins without a counted hit (never the unique maximum of a gene with hits)
        if not ties:                                           # make every gene's maximum unique
            for g in range(n_gene):
                idx = np.nonzero(gid == g)[0]
                z[idx[0]] = (np.uint64(1 << 23) + np.uint64(g)) << np.uint64(32) | np.uint64(1)
        else:
            idx = np.nonzero(gid == 7)[0]; z[idx] = z[idx[0]] | np.uint64(1 << 60)     # one gene whose isoforms all share the maximum
        a = np.zeros(n_gene, np.int32); b = np.zeros(n_gene, np.int32)
        ra = L.pgx_test_choose_representatives(z.ctypes.data, gid.ctypes.data, n_prot, n_gene, 0, a.ctypes.data)
        rb = L.pgx_test_choose_representatives(z.ctypes.data, gid.ctypes.data, n_prot, n_gene, 1, b.ctypes.data)
        assert rb == 1 and ra == (1 if ties else 0)
        assert np.array_equal(a, b)


def _same_data(d, d2):
    assert (d.contents.n_genome, d.contents.n_gene, d.contents.n_prot) == (d2.contents.n_genome, d2.contents.n_gene, d2.contents.n_prot)
    for j in range(d.contents.n_genome):
        a, b = d.contents.genome[j], d2.contents.genome[j]
        assert (a.n_hit, a.n_exon, a.n_ctg, a.label) == (b.n_hit, b.n_exon, b.n_ctg, b.label)
        assert C.string_at(a.hit, a.n_hit * 88) == C.string_at(b.hit, b.n_hit * 88)
        assert C.string_at(a.exon, a.n_exon * 8) == C.string_at(b.exon, b.n_exon * 8)
        assert [(a.ctg[c].name, a.ctg[c].len) for c in range(a.n_ctg)] == [(b.ctg[c].name, b.ctg[c].len) for c in range(b.n_ctg)]
    for i in range(d.contents.n_prot):
        x, y = d.contents.prot[i], d2.contents.prot[i]
        assert (x.name, x.len, x.gid) == (y.name, y.len, y.gid)
    for i in range(d.contents.n_gene):
        x, y = d.contents.gene[i], d2.contents.gene[i]
        assert (x.name, x.bits) == (y.name, y.bits)


def test_parallel_reader_reproduces_sequential_ids(tmp_path):
    """per-file parsing with local ids + ordered merge == file-by-file parsing: ids, names, lengths, ranks, every hit byte"""
    opt = hostio.default_opt()
    files = common.c4_files(tmp_path)
    for thr in (0, 3):
        a, b = hostio.read_pafs(opt, files, threads=1), hostio.read_pafs(opt, files, threads=thr)
        _same_data(a, b)
        hostio.destroy(a); hostio.destroy(b)
    # proteins first seen in LATER files, files listed in a scrambled order, one missing file in the middle
    p = hostio.synth_params(2, n_genome=6, n_gene=500, iso_per_gene=4, hits_per_genome=4000, n_genome_total=6)
    fs = hostio.synth_write_pafs(opt, p, str(tmp_path / "syn"))
    order = [fs[3], fs[5], str(tmp_path / "does_not_exist.paf"), fs[0], fs[4], fs[1], fs[2]]
    a, b = hostio.read_pafs(opt, order, threads=1), hostio.read_pafs(opt, order, threads=4)
    assert a.contents.n_genome == 6
    _same_data(a, b)
    hostio.destroy(a); hostio.destroy(b)


def test_reader_line_ending_and_container_variants(tmp_path):
    """the same records as plain text, gzip, CRLF text and text without a final newline parse to the same pg_data_t; odd trailing
    fields (a short tag-like last field, non-numeric columns) are harmless"""
    import gzip
    opt = hostio.default_opt()
    p = hostio.synth_params(2, n_genome=1, n_gene=300, iso_per_gene=4, hits_per_genome=2500, n_genome_total=1)
    (plain,) = hostio.synth_write_pafs(opt, p, str(tmp_path / "v"))
    text = open(plain, "rb").read()
    name = os.path.basename(plain)
    variants = {}
    for tag, data in [("gz", None), ("crlf", text.replace(b"\n", b"\r\n")), ("nonl", text.rstrip(b"\n")), ("tail", text.rstrip(b"\n") + b"\tzz:\tab\tms:")]:
        d = tmp_path / tag
        d.mkdir()
        if tag == "gz":
            fn = str(d / (name + ".gz"))
            with gzip.open(fn, "wb") as f:
                f.write(text)
        else:
            fn = str(d / name)
            open(fn, "wb").write(data)
        variants[tag] = fn
    a = hostio.read_pafs(opt, [plain])
    for tag, fn in variants.items():
        b = hostio.read_pafs(opt, [fn])
        _same_data(a, b)
        hostio.destroy(b)
    hostio.destroy(a)

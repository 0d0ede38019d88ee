"""The sync-free round kernels (pangene_b200/csrc/pgx_round.cuh), checked on the CPU: tests/emu/ compiles them for the host (SIMT shim) and
this test runs the launch sequence of Session::gen_arc_fast -- table insert, (key, genome) marks, singleton fold, overflow groups, the
rank-to-rank merge, vertex scan, counting sort -- on random survivor lists against a plain-Python restatement of graph.c:97-176 (per-genome
collapse with the rounded mean, global collapse, arcs sorted by key, idx of graph.c:202-211).  Also: the one-pass survivor compaction, the
two-launch segment filter (graph.c:237-263) and the dirty-tile marks of the incremental sweep against brute force."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu import build as emu_build  # noqa: E402

ARC = np.dtype([("x", "<u8"), ("n_genome", "<i4"), ("tot_cnt", "<i4"), ("avg_dist", "<i4"), ("s1", "<i4"), ("s2", "<i4"), ("bits", "<u4")])
SEG = np.dtype([("gid", "<i4"), ("n_dom", "<i4"), ("n_sub", "<i4"), ("n_genome", "<i4"), ("tot_cnt", "<i4"), ("del", "<u4"), ("l0", "<i4"), ("l1", "<i4")])


class EmuArcIn(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("n_hit", "n_sv", "n_seg0", "n_gene", "G", "use_ori", "grid", "H", "H2", "arc_cap", "ranks")] + \
               [(k, C.c_void_p) for k in ("sv", "bits", "ctg", "cm", "gidp", "sori", "sdom", "pdom0", "g2s", "prot_gid", "ctg2g", "arc", "idx", "seg_ng", "seg_tot", "nl_off", "out_counts")]


@pytest.fixture(scope="module")
def emu():
    L = C.CDLL(emu_build.build_round())
    L.round_emu_arc_build.argtypes = [C.POINTER(EmuArcIn)]
    L.round_emu_compact.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.round_emu_seg_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]
    L.round_emu_dirty.argtypes = [C.c_void_p] * 4 + [C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    return L


def ref_arcs(sv, ctg, cm, sid, rev, score, genome_of_ctg, n_seg0):
    """graph.c:97-176 on a survivor list in (contig, cm) order"""
    per_genome = {}
    seg_cnt = {}
    prev = None
    for h in sv:
        g = genome_of_ctg[ctg[h]]
        w = (sid[h] << 1) | rev[h]
        seg_cnt.setdefault(g, {}).setdefault(sid[h], 0)
        seg_cnt[g][sid[h]] += 1
        if prev is not None and ctg[prev] == ctg[h]:
            v = (sid[prev] << 1) | rev[prev]
            d = int(cm[h]) - int(cm[prev])
            per_genome.setdefault(g, {}).setdefault((v << 32) | w, []).append((d, score[prev], score[h]))
            per_genome[g].setdefault(((w ^ 1) << 32) | (v ^ 1), []).append((d, score[h], score[prev]))
        prev = h
    glob = {}
    for g, recs in per_genome.items():
        for x, lst in recs.items():
            n = len(lst)
            dist = int(float(sum(r[0] for r in lst)) / n + .499)
            s1 = max([0] + [r[1] for r in lst]); s2 = max([0] + [r[2] for r in lst])
            glob.setdefault(x, []).append((n, dist, s1, s2))
    arcs = []
    for x in sorted(glob):
        lst = glob[x]
        ng, tot = len(lst), sum(r[0] for r in lst)
        arcs.append((x, ng, tot, int(float(sum(r[1] * r[0] for r in lst)) / tot + .499), int(float(sum(r[2] for r in lst)) / ng + .499), int(float(sum(r[3] for r in lst)) / ng + .499)))
    seg_ng, seg_tot = np.zeros(n_seg0, np.int32), np.zeros(n_seg0, np.int32)
    for g, cnt in seg_cnt.items():
        for s, c in cnt.items():
            seg_ng[s] += 1; seg_tot[s] += c
    return arcs, seg_ng, seg_tot


def make_case(rng, G, n_seg0, n_per_genome, dup_frac, cm_ties=False):
    """survivors of G genomes: genes in a mostly conserved order (so that arcs repeat across genomes), some tandem repeats of one gene pair
    inside a genome (groups of several records per (key, genome): the overflow path) and some v -> v^1 self-mirror adjacencies"""
    n_gene = n_seg0 + 7
    g2s = np.full(n_gene, -1, np.int32)
    genes = rng.permutation(n_gene)[:n_seg0]
    g2s[np.sort(genes)] = np.arange(n_seg0)
    vtx_genes = np.sort(genes)
    ctg, cm, gene, rev, sori, sdom, pdom0, ctg2g = [], [], [], [], [], [], [], []
    n_prot = n_gene
    for g in range(G):
        n_ctg = rng.integers(1, 4)
        for c in range(n_ctg):
            ctg2g.append(g)
            cid = len(ctg2g) - 1
            order = list(vtx_genes[rng.integers(0, n_seg0, size=max(2, n_per_genome // n_ctg))]) if rng.random() < 0.3 else list(vtx_genes[(np.arange(max(2, n_per_genome // n_ctg)) + rng.integers(0, 5)) % n_seg0])
            pos = 1000
            k = 0
            while k < len(order):
                gid = order[k]
                r = int(rng.integers(0, 2))
                reps = 1
                if rng.random() < dup_frac: reps = int(rng.integers(2, 5))
                for q in range(reps):
                    for gg, rr in ((gid, r), (order[(k + 1) % len(order)], r ^ (q & 1))) if reps > 1 else ((gid, r),):
                        pos += int(rng.integers(1, 3000)) if (not cm_ties or rng.random() < 0.9) else 0      # cm_ties: some equal (contig, cm)
                        ctg.append(cid); cm.append(pos); gene.append(gg); rev.append(rr)
                        sori.append(int(rng.integers(0, 900))); sdom.append(int(rng.integers(0, 900))); pdom0.append(int(rng.integers(-1, n_prot)))
                if rng.random() < 0.05:    # the same gene again on the other strand: v -> v^1
                    pos += 10; ctg.append(cid); cm.append(pos); gene.append(gid); rev.append(r ^ 1); sori.append(5); sdom.append(700); pdom0.append(-1)
                k += 1
    n = len(ctg)
    a = lambda x, t: np.array(x, dtype=t)
    cols = dict(ctg=a(ctg, np.int32), cm=a(cm, np.int32), gidp=a(gene, np.int32) << 1, sori=a(sori, np.int32), sdom=a(sdom, np.int32), pdom0=a(pdom0, np.int32),
                bits=a(rev, np.uint32), g2s=g2s, prot_gid=np.arange(n_prot, dtype=np.int32), ctg2g=a(ctg2g, np.int32))
    # survivors: a random 85 % of the hits, in (contig, cm) order (the generator emits them that way)
    keep = rng.random(n) < 0.85
    sv = np.nonzero(keep)[0].astype(np.uint32)
    return cols, sv, n_gene


@pytest.mark.parametrize("seed,G,n_seg0,n_per,dup,ranks,use_ori", [(1, 1, 12, 40, 0.0, 1, 0), (2, 5, 40, 120, 0.1, 1, 0), (3, 7, 60, 200, 0.3, 1, 1), (4, 6, 50, 150, 0.15, 2, 0), (5, 9, 30, 90, 0.2, 3, 0), (6, 3, 25, 60, 0.1, 4, 1)])
def test_arc_build_matches_graph_c(emu, seed, G, n_seg0, n_per, dup, ranks, use_ori):
    rng = np.random.default_rng(seed)
    cols, sv, n_gene = make_case(rng, G, n_seg0, n_per, dup)
    n = len(cols["ctg"])
    sid = cols["g2s"][cols["gidp"] >> 1]
    rev = (cols["bits"] & 1).astype(np.int64)
    # graph.c:82-85
    dom_vtx = np.where(cols["pdom0"] >= 0, cols["g2s"][cols["prot_gid"][np.maximum(cols["pdom0"], 0)]] >= 0, True)
    score = np.where((use_ori != 0) | (cols["sori"] > cols["sdom"]) | (cols["pdom0"] < 0) | dom_vtx, cols["sori"], cols["sdom"])
    want, want_ng, want_tot = ref_arcs(list(sv), cols["ctg"], cols["cm"], sid.astype(np.int64), rev, score.astype(np.int64), cols["ctg2g"], n_seg0)
    arc_cap = 4 * len(sv) + 16
    arc = np.zeros(arc_cap, dtype=ARC); idx = np.zeros(2 * n_seg0 + 2, np.uint64); seg_ng = np.zeros(n_seg0, np.int32); seg_tot = np.zeros(n_seg0, np.int32)
    nl_off = np.zeros(2 * n_seg0 + 8, np.uint32); out = np.zeros(4, np.uint32)
    H = 1
    while H < 4 * max(len(want), 8): H *= 2
    e = EmuArcIn(n_hit=n, n_sv=len(sv), n_seg0=n_seg0, n_gene=n_gene, G=G, use_ori=use_ori, grid=3, H=H, H2=8192, arc_cap=arc_cap, ranks=ranks)
    keepalive = [sv, arc, idx, seg_ng, seg_tot, nl_off, out]
    for k in ("ctg", "cm", "gidp", "sori", "sdom", "pdom0", "g2s", "prot_gid", "ctg2g", "bits"):
        setattr(e, k, cols[k].ctypes.data)
    e.sv = sv.ctypes.data; e.arc = arc.ctypes.data; e.idx = idx.ctypes.data; e.seg_ng = seg_ng.ctypes.data; e.seg_tot = seg_tot.ctypes.data
    e.nl_off = nl_off.ctypes.data; e.out_counts = out.ctypes.data
    assert emu.round_emu_arc_build(C.byref(e)) == 0
    assert out[2] == 0, "fallback raised"
    got = [(int(a["x"]), int(a["n_genome"]), int(a["tot_cnt"]), int(a["avg_dist"]), int(a["s1"]), int(a["s2"])) for a in arc[:out[0]]]
    assert got == want
    if dup > 0 and ranks == 1: assert out[3] > 0          # the overflow path was exercised
    assert np.array_equal(seg_ng, want_ng) and np.array_equal(seg_tot, want_tot)
    # idx (graph.c:202-211) and the pair-table offsets (n (n + 1) / 2 for every vertex with >= 2 out-arcs)
    want_idx = np.zeros(2 * n_seg0, np.uint64); deg = np.zeros(2 * n_seg0, np.int64)
    for p, a in enumerate(want):
        v = a[0] >> 32
        if deg[v] == 0: want_idx[v] = p << 32
        deg[v] += 1
    want_idx |= deg.astype(np.uint64)
    assert np.array_equal(idx[:2 * n_seg0], want_idx)
    pairs = np.where(deg >= 2, deg * (deg + 1) // 2, 0)
    assert np.array_equal(nl_off[:2 * n_seg0 + 1], np.concatenate([[0], np.cumsum(pairs)]).astype(np.uint32)) and out[1] == pairs.sum()
    del keepalive


def test_arc_build_raises_fallback_on_cm_tied_survivors(emu):
    """two adjacent survivors with one (contig, cm): which of them the reference walks first depends on its sort chain -- the fast path must
    hand such a run over to the careful path (Session::fix_cm_tied_survivors), never decide the order itself"""
    rng = np.random.default_rng(12)
    cols, sv, n_gene = make_case(rng, 4, 40, 150, 0.0, cm_ties=True)
    assert np.any((np.diff(cols["cm"][sv].astype(np.int64)) == 0) & (np.diff(cols["ctg"][sv]) == 0))
    arc_cap = 4 * len(sv) + 16
    arc = np.zeros(arc_cap, dtype=ARC); idx = np.zeros(82, np.uint64); seg_ng = np.zeros(40, np.int32); seg_tot = np.zeros(40, np.int32); nl_off = np.zeros(88, np.uint32); out = np.zeros(4, np.uint32)
    e = EmuArcIn(n_hit=len(cols["ctg"]), n_sv=len(sv), n_seg0=40, n_gene=n_gene, G=4, use_ori=0, grid=2, H=4096, H2=256, arc_cap=arc_cap, ranks=1)
    for k in ("ctg", "cm", "gidp", "sori", "sdom", "pdom0", "g2s", "prot_gid", "ctg2g", "bits"):
        setattr(e, k, cols[k].ctypes.data)
    e.sv = sv.ctypes.data; e.arc = arc.ctypes.data; e.idx = idx.ctypes.data; e.seg_ng = seg_ng.ctypes.data; e.seg_tot = seg_tot.ctypes.data
    e.nl_off = nl_off.ctypes.data; e.out_counts = out.ctypes.data
    emu.round_emu_arc_build(C.byref(e))
    assert out[2] == 1


def test_arc_build_raises_fallback_when_the_table_is_too_small(emu):
    rng = np.random.default_rng(11)
    cols, sv, n_gene = make_case(rng, 4, 40, 150, 0.1)
    arc_cap = 4 * len(sv) + 16
    arc = np.zeros(arc_cap, dtype=ARC); idx = np.zeros(82, np.uint64); seg_ng = np.zeros(40, np.int32); seg_tot = np.zeros(40, np.int32); nl_off = np.zeros(88, np.uint32); out = np.zeros(4, np.uint32)
    e = EmuArcIn(n_hit=len(cols["ctg"]), n_sv=len(sv), n_seg0=40, n_gene=n_gene, G=4, use_ori=0, grid=2, H=64, H2=256, arc_cap=arc_cap, ranks=1)
    for k in ("ctg", "cm", "gidp", "sori", "sdom", "pdom0", "g2s", "prot_gid", "ctg2g", "bits"):
        setattr(e, k, cols[k].ctypes.data)
    e.sv = sv.ctypes.data; e.arc = arc.ctypes.data; e.idx = idx.ctypes.data; e.seg_ng = seg_ng.ctypes.data; e.seg_tot = seg_tot.ctypes.data
    e.nl_off = nl_off.ctypes.data; e.out_counts = out.ctypes.data
    emu.round_emu_arc_build(C.byref(e))
    assert out[2] == 1


def test_arc_build_raises_fallback_when_the_overflow_table_is_too_small(emu):
    """many (key, genome) groups of several records (tandem repeats) against an overflow table of 16 entries: k_arc_fold's group insert must
    raise the fallback (the run is then redone on the careful path), and the table must still be handed back empty for the next round"""
    rng = np.random.default_rng(13)
    cols, sv, n_gene = make_case(rng, 4, 40, 200, 0.6)
    arc_cap = 4 * len(sv) + 16
    arc = np.zeros(arc_cap, dtype=ARC); idx = np.zeros(82, np.uint64); seg_ng = np.zeros(40, np.int32); seg_tot = np.zeros(40, np.int32); nl_off = np.zeros(88, np.uint32); out = np.zeros(4, np.uint32)
    e = EmuArcIn(n_hit=len(cols["ctg"]), n_sv=len(sv), n_seg0=40, n_gene=n_gene, G=4, use_ori=0, grid=2, H=4096, H2=16, arc_cap=arc_cap, ranks=1)
    for k in ("ctg", "cm", "gidp", "sori", "sdom", "pdom0", "g2s", "prot_gid", "ctg2g", "bits"):
        setattr(e, k, cols[k].ctypes.data)
    e.sv = sv.ctypes.data; e.arc = arc.ctypes.data; e.idx = idx.ctypes.data; e.seg_ng = seg_ng.ctypes.data; e.seg_tot = seg_tot.ctypes.data
    e.nl_off = nl_off.ctypes.data; e.out_counts = out.ctypes.data
    rc = emu.round_emu_arc_build(C.byref(e))
    assert rc == 0, "the overflow table was not handed back clean (rc %d)" % rc
    assert out[3] > 32 and out[2] == 1


@pytest.mark.parametrize("n,with_perm", [(0, False), (1, False), (2047, True), (2048, False), (2049, True), (9000, True)])
def test_survivor_compaction(emu, n, with_perm):
    rng = np.random.default_rng(n)
    bits = rng.integers(0, 1 << 11, size=n + 1, dtype=np.uint32)
    perm = rng.permutation(n).astype(np.uint32) if with_perm else None
    pos = np.full(n + 1 + 4, 0xdeadbeef, np.uint32); lst = np.full(n + 4, 0xdeadbeef, np.uint32); cnt = np.zeros(1, np.uint32)
    assert emu.round_emu_compact(bits.ctypes.data, perm.ctypes.data if with_perm else None, n, pos.ctypes.data, lst.ctypes.data, cnt.ctypes.data) == 0
    order = perm if with_perm else np.arange(n, dtype=np.uint32)
    alive = (bits[order] & (0x002 | 0x080)) == 0 if n else np.zeros(0, bool)
    assert cnt[0] == alive.sum()
    assert np.array_equal(lst[:cnt[0]], order[alive])
    assert np.array_equal(pos[:n + 1], np.concatenate([[0], np.cumsum(alive)]).astype(np.uint32))
    assert np.all(pos[n + 1:] == 0xdeadbeef) and np.all(lst[cnt[0]:] == 0xdeadbeef)


@pytest.mark.parametrize("n_seg", [1, 700, 1024, 1025, 3333])
def test_segment_filter(emu, n_seg):
    rng = np.random.default_rng(n_seg)
    n_gene = n_seg + 50
    seg = np.zeros(n_seg, dtype=SEG)
    seg["gid"] = np.sort(rng.permutation(n_gene)[:n_seg]); seg["tot_cnt"] = rng.integers(0, 60, n_seg); seg["l0"] = rng.integers(0, 5, n_seg); seg["l1"] = rng.integers(0, 5, n_seg)
    seg["n_dom"] = rng.integers(0, 9, n_seg)
    idx = (rng.integers(0, 1 << 20, 2 * n_seg).astype(np.uint64) << np.uint64(32)) | rng.integers(0, 12, 2 * n_seg).astype(np.uint64)
    deg = (idx & np.uint64(0xffffffff)).astype(np.int64)
    keep = ~((seg["tot_cnt"] > 40) | (deg[0::2] > 8) | (deg[1::2] > 8) | (np.maximum(seg["l0"], seg["l1"]) > 3))
    want = seg[keep].copy()
    g2s = np.full(n_gene, 12345, np.int32); n = np.array([n_seg], np.int32)
    work = np.zeros(n_seg + 300, dtype=SEG)
    assert emu.round_emu_seg_filter(seg.ctypes.data, work.ctypes.data, n.ctypes.data, n_seg + 300, idx.ctypes.data, 40, 8, 3, g2s.ctypes.data, n_gene) == 0   # capacity > count: segments were deleted in earlier rounds
    assert n[0] == keep.sum() and np.array_equal(work[:n[0]], want)
    want_g2s = np.full(n_gene, -1, np.int32); want_g2s[want["gid"]] = np.arange(len(want))
    assert np.array_equal(g2s, want_g2s)


def test_dirty_tiles_cover_every_hit_that_sees_a_changed_hit(emu):
    rng = np.random.default_rng(5)
    n, tile = 3000, 64
    ctg = np.sort(rng.integers(0, 6, n)).astype(np.int32)
    cs = np.zeros(n, np.int32); ce = np.zeros(n, np.int32)
    for c in range(6):
        m = ctg == c
        s = np.sort(rng.integers(0, 200000, m.sum()))
        cs[m] = s; ce[m] = s + np.where(rng.random(m.sum()) < 0.02, rng.integers(1000, 60000, m.sum()), rng.integers(50, 900, m.sum()))
    pmax = np.zeros(n, np.int32)
    for i in range(n):
        pmax[i] = ce[i] if i == 0 or ctg[i - 1] != ctg[i] else max(pmax[i - 1], ce[i])
    changed = rng.choice(n, 25, replace=False).astype(np.int32)
    tiles = np.zeros(n // tile // 32 + 2, np.uint32)
    assert emu.round_emu_dirty(ctg.ctypes.data, cs.ctypes.data, ce.ctypes.data, pmax.ctypes.data, n, tile, changed.ctypes.data, len(changed), tiles.ctypes.data) == 0
    marked = lambda t: bool(tiles[t >> 5] >> np.uint32(t & 31) & np.uint32(1))
    n_marked = sum(marked(t) for t in range((n + tile - 1) // tile))
    for h in changed:
        assert marked(h // tile)
        for i in range(n):      # every hit that has h as a window partner (overlap.c:67-71,114-120: same contig, intervals overlap) sits in a marked tile
            if i != h and ctg[i] == ctg[h] and cs[i] < ce[h] and ce[i] > cs[h]:
                assert marked(i // tile), (h, i)
    assert n_marked < (n + tile - 1) // tile          # and not simply everything

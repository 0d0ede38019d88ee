"""The production sweep kernel's control flow, checked on the CPU: tests/emu/ compiles the sweep kernels of
pangene_b200/csrc/pgx_kernels.cuh for the host (SIMT shim: one OS thread per CUDA thread) and this test runs every variant on the
same hit columns.  The plain thread-per-hit kernel `k_sweep` is the restatement of overlap.c:58-178 that the GPU parity tests
pin against the oracle; the tiled production kernel `k_sweep_tiled` must reproduce its outputs bit for bit in all three modes."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import common  # noqa: E402
from pangene_b200 import abi, hostio  # noqa: E402
from emu import build as emu_build  # noqa: E402

HIT = np.dtype([(n, "<i4") for n in ("pid", "qs", "qe", "cid", "mlen", "blen", "lof", "rank", "score_ori", "score_adj", "score_dom", "n_exon", "off_exon", "pid_dom", "pid_dom0")]
               + [("bits", "<u4"), ("cs", "<i8"), ("cm", "<i8"), ("ce", "<i8")])
PAD = 64   # the tiled kernels stage 16-byte-granular column windows: columns carry slack past n


class EmuCols(C.Structure):
    _fields_ = [("n", C.c_int32)] + [(k, C.c_void_p) for k in ("ctg", "cs", "ce", "pid", "gidp", "rank", "cds", "sori", "pmax", "key", "bits", "eoff", "ex")]


@pytest.fixture(scope="module")
def emu():
    L = C.CDLL(emu_build.build())
    L.sweep_emu.argtypes = [C.c_int, C.c_int, C.POINTER(EmuCols), C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.sweep_emu.restype = C.c_int
    return L


def mix32(k):
    k = k.astype(np.uint32)
    with np.errstate(over="ignore"):
        k = k + ~(k << np.uint32(15)); k ^= k >> np.uint32(10); k = k + (k << np.uint32(3)); k ^= k >> np.uint32(6)
        k = k + ~(k << np.uint32(11)); k ^= k >> np.uint32(16)
    return k


def padded(a, dtype, fill=0):
    out = np.full(len(a) + PAD, fill, dtype=dtype)
    out[:len(a)] = a
    return out


def columns(d, rng, flt_frac, wbr_frac, stress=False, score_ties=False):
    """hit columns as stage_load builds them (k_gather / k_gather_exons), all genomes, (contig, cs) order.  stress: some hits are
    stretched over hundreds of followers (windows that leave the staged range, windows longer than the range-maximum table
    covers) and some end exactly where a follower starts (the strict/non-strict window bounds).  score_ties: adjusted scores are rounded
    down to multiples of 64, so that the strongest hits of many overlap clusters share the upper key word (the tie path of VAR 1 / 2)"""
    dd = d.contents
    gid = np.array([dd.prot[i].gid for i in range(dd.n_prot)], dtype=np.int64)
    pref = np.array([(dd.gene[i].bits >> 30) & 1 for i in range(dd.n_gene)], dtype=np.int64)
    cols = {k: [] for k in ("ctg", "cs", "ce", "pid", "rank", "sori", "sadj", "rev", "nex", "first")}
    exons = []
    coff = 0
    for j in range(dd.n_genome):
        g = dd.genome[j]
        h = np.frombuffer((C.c_char * (g.n_hit * 88)).from_address(C.addressof(g.hit.contents)), dtype=HIT)
        ex = np.frombuffer((C.c_char * (g.n_exon * 8)).from_address(C.addressof(g.exon.contents)), dtype="<i4").reshape(-1, 2)
        order = np.lexsort((h["cs"], h["cid"]))
        h = h[order]
        cols["ctg"].append(h["cid"] + coff); cols["cs"].append(h["cs"]); cols["ce"].append(h["ce"]); cols["pid"].append(h["pid"])
        cols["rank"].append(h["rank"]); cols["sori"].append(h["score_ori"]); cols["sadj"].append(h["score_adj"]); cols["rev"].append(h["bits"] & 1)
        cols["nex"].append(h["n_exon"])
        first = np.zeros(len(h), dtype=np.uint32); first[:1] = 0x80000000
        cols["first"].append(first)
        for o, n in zip(h["off_exon"], h["n_exon"]):
            exons.append(ex[o:o + n])
        coff += g.n_ctg
    c = {k: np.concatenate(v) for k, v in cols.items()}
    n = len(c["cs"])
    ex = np.concatenate(exons).astype(np.int32)
    eoff = np.zeros(n + 1, dtype=np.uint32); eoff[1:] = np.cumsum(c["nex"])
    cds = np.add.reduceat(ex[:, 1] - ex[:, 0], eoff[:-1].astype(np.int64)).astype(np.int32)
    if stress:
        ce, cs, ctg = c["ce"].copy(), c["cs"], c["ctg"]
        for i in np.nonzero(rng.random(n) < 0.05)[0]:
            k = min(n - 1, i + int(rng.integers(1, 4)))
            if ctg[k] == ctg[i] and cs[k] > cs[i]: ce[i] = cs[k]
        for i in np.nonzero(rng.random(n) < 0.006)[0]:
            k = min(n - 1, i + int(rng.integers(40, 400)))
            if ctg[k] == ctg[i]: ce[i] = max(ce[i], cs[k] + 1)
        c["ce"] = ce
    pmax = np.empty(n, dtype=np.int64)
    run, prev = -1, -1
    for i in range(n):
        if c["ctg"][i] != prev: run, prev = -1, c["ctg"][i]
        run = max(run, c["ce"][i]); pmax[i] = run
    gidp = gid[c["pid"]] << 1 | pref[gid[c["pid"]]]
    if score_ties:
        c["sadj"] = c["sadj"] // 64 * 64
    key = (c["sadj"].astype(np.int64).astype(np.uint64) << np.uint64(33)) | ((gidp & 1).astype(np.uint64) << np.uint64(32)) | mix32(c["pid"]).astype(np.uint64)
    bits = c["rev"].astype(np.uint32) | c["first"]
    bits |= np.where(rng.random(n) < flt_frac, np.uint32(abi.PG_HB_FLT | abi.PG_HB_PSEUDO), np.uint32(0)).astype(np.uint32)
    bits |= np.where(rng.random(n) < wbr_frac, rng.integers(1, 3, n).astype(np.uint32) << np.uint32(9), np.uint32(0)).astype(np.uint32)
    bits |= np.where(rng.random(n) < 0.3, np.uint32(abi.PG_HB_SHADOW), np.uint32(0)).astype(np.uint32)   # stale shadow bits must be reset
    out = dict(n=n, ctg=padded(c["ctg"], np.int32, 0x7fffffff), cs=padded(c["cs"], np.int32), ce=padded(c["ce"], np.int32), pid=padded(c["pid"], np.int32),
               gidp=padded(gidp, np.int32), rank=padded(c["rank"], np.int32), cds=padded(cds, np.int32), sori=padded(c["sori"], np.int32), pmax=padded(pmax, np.int32),
               key=padded(key, np.uint64), bits=padded(bits, np.uint32), eoff=padded(eoff, np.uint32, int(eoff[-1])), ex=np.concatenate([ex.reshape(-1), np.zeros(2 * PAD, np.int32)]))
    return out


def run(L, variant, mode, cols, ratio=0.5, check_strand=0):
    e = EmuCols(n=cols["n"], **{k: cols[k].ctypes.data for k in ("ctg", "cs", "ce", "pid", "gidp", "rank", "cds", "sori", "pmax", "key", "bits", "eoff", "ex")})
    n = cols["n"]
    bits_out = np.full(n, 0xdeadbeef, np.uint32); pdom = np.full(n, -7, np.int32); sdom = np.full(n, -7, np.int32); tie = np.zeros(1, np.uint32)
    assert L.sweep_emu(variant, mode, C.byref(e), ratio, check_strand, bits_out.ctypes.data, pdom.ctypes.data, sdom.ctypes.data, tie.ctypes.data) == 0
    return bits_out, pdom, sdom, int(tie[0])


CASES = [
    # name, synth kwargs, filtered fraction, weak_br fraction, check_strand
    ("euk4iso", dict(config=3, n_genome=2, n_gene=300, iso_per_gene=4, hits_per_genome=2600), 0.25, 0.0, 0),
    ("euk_weakbr", dict(config=3, n_genome=1, n_gene=250, iso_per_gene=4, hits_per_genome=2200), 0.05, 0.04, 0),
    ("euk_strand", dict(config=2, n_genome=2, n_gene=400, iso_per_gene=1, hits_per_genome=1500), 0.1, 0.0, 1),
    ("bacterial", dict(config=5, n_genome=3, n_gene=600, iso_per_gene=1, hits_per_genome=700), 0.0, 0.02, 0),
    ("euk_stress", dict(config=3, n_genome=1, n_gene=300, iso_per_gene=4, hits_per_genome=1200, n_ctg=2, stress=True), 0.2, 0.0, 0),
    ("euk_score_ties", dict(config=3, n_genome=1, n_gene=300, iso_per_gene=4, hits_per_genome=2400, score_ties=True), 0.15, 0.0, 0),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tiled_sweeps_equal_plain_sweep(emu, case):
    name, kw, flt, wbr, strand = case
    opt = hostio.default_opt()
    p = hostio.synth_params(kw["config"], n_genome=kw["n_genome"], n_gene=kw["n_gene"], iso_per_gene=kw["iso_per_gene"], hits_per_genome=kw["hits_per_genome"],
                            n_genome_total=kw["n_genome"], **({"n_ctg": kw["n_ctg"]} if "n_ctg" in kw else {}))
    d = hostio.synth(opt, p)
    cols = columns(d, np.random.default_rng(11), flt, wbr, stress=kw.get("stress", False), score_ties=kw.get("score_ties", False))
    if kw.get("score_ties"):   # the case must really put equal upper key words side by side
        hi = (cols["key"][:cols["n"]] >> np.uint64(32)).astype(np.int64)
        near = (hi[1:] == hi[:-1]) & (cols["ctg"][1:cols["n"]] == cols["ctg"][:cols["n"] - 1]) & (cols["cs"][1:cols["n"]] < cols["ce"][:cols["n"] - 1])
        assert int(near.sum()) > cols["n"] // 20
    assert cols["n"] > 1000
    for mode in (0, 1, 2):
        ref = run(emu, 0, mode, cols, check_strand=strand)
        n_dom = int((ref[1] >= 0).sum()) if mode != 2 else int((ref[0] & abi.PG_HB_FLT_OV != 0).sum())
        if mode != 2 or kw["iso_per_gene"] > 1:
            assert n_dom > cols["n"] // 50, "the case does not exercise the sweep"
        # k_sweep_tiled<MODE, VAR>: VAR 1 (the default; emu variant 2) on every case, all four on the two isoform-rich ones (tests/emu/build.py; each
        # variant costs a full emulated sweep per mode)
        for variant in ((1, 2, 3, 4) if name in ("euk4iso", "euk_score_ties") else (2,)):
            got = run(emu, variant, mode, cols, check_strand=strand)
            for k, what in enumerate(("bits", "pid_dom", "score_dom")):
                if k == 0 and mode == 0 and variant > 1:
                    continue   # SW_LOAD leaves the flag words as they are; VAR 1 / 2 do not copy them out (the isoform sweep that follows writes them all)
                bad = np.nonzero(ref[k] != got[k])[0]
                assert len(bad) == 0, "%s: variant %d mode %d: %s differs at %s (plain %s, tiled %s)" % (name, variant, mode, what, bad[:5], ref[k][bad[:5]], got[k][bad[:5]])
    hostio.destroy(d)


def test_related_threshold_shortcut_is_exact(emu):
    """related_threshold() answers most pairs without a division (a margin argument on ceil(ratio * shorter)); the result must be the smallest
    integer overlap ov with NOT ((double)ov / shorter < ratio) -- the reference's own predicate (overlap.c:134-136) -- for every input"""
    rng = np.random.default_rng(9)
    shorter = np.concatenate([np.arange(1, 200001, dtype=np.int32), rng.integers(1, 5_000_000, 400000).astype(np.int32)])
    out = np.zeros(len(shorter), np.int32)
    emu.sweep_emu_thresholds.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    for r in [0.5, 0.2, 1.0 / 3, 0.9, 0.05, 0.999999, 0.6180339887, 1.0]:
        ratio = np.full(len(shorter), r, np.float64)
        emu.sweep_emu_thresholds(shorter.ctypes.data, ratio.ctypes.data, len(shorter), out.ctypes.data)
        need = out.astype(np.int64)
        s = shorter.astype(np.float64)
        assert np.all(~(need.astype(np.float64) / s < r))                                       # need itself puts the pair in relation
        assert np.all((need == 1) | ((need - 1).astype(np.float64) / s < r))                    # need - 1 does not
    ratio = rng.random(len(shorter))
    emu.sweep_emu_thresholds(shorter.ctypes.data, ratio.ctypes.data, len(shorter), out.ctypes.data)
    need = out.astype(np.int64); s = shorter.astype(np.float64)
    assert np.all(~(need.astype(np.float64) / s < ratio)) and np.all((need == 1) | ((need - 1).astype(np.float64) / s < ratio))

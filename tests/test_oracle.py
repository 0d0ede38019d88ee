"""CPU suite, part 1: the oracle restatement (oracle/pg_oracle.c) is pinned to the reference.

 * against the committed goldens printed by the UNMODIFIED reference binary (tests/golden);
 * against the compiled reference run in-process (oracle/_ref/libpgref.so) on randomized synthetic inputs and
   option combinations -- only where oracle/_ref exists (the build container and the GPU box both have it);
 * the exact radix-permutation replay against the reference's own radix_sort_pg128x.
"""
import ctypes as C
import hashlib
import os
import random

import pytest

import common
import refharness
from oracle import pgoracle
from pangene_b200 import abi, hostio

C4_MD5 = {"C4.gfa": "518822a3debc3e12a4f908cfc123c9e5", "C4.bed_raw": "e56f1f73b9960dfd842f0178d0f58da5",
          "C4.bed_flag": "7257ac152f3589c55def695cd69e7ba1", "C4.bed_walk": "ea05e71a76e3162430854e17700aef7a",
          "C4_w.gfa": "a1d1594057b80a03a273d89d6d7c33e2"}


def test_golden_files_are_the_surveyed_ones():
    for name, md5 in C4_MD5.items():
        assert hashlib.md5(open(os.path.join(common.GOLDEN, name), "rb").read()).hexdigest() == md5


def test_oracle_c4_gfa_byte_identical(tmp_path):
    opt = hostio.default_opt()
    d = hostio.read_pafs(opt, common.c4_files(tmp_path))
    g = pgoracle.hot_path(opt, d)
    assert common.gfa_bytes(g) == open(os.path.join(common.GOLDEN, "C4.gfa"), "rb").read()
    assert common.gfa_bytes(g, walk=False) == open(os.path.join(common.GOLDEN, "C4_w.gfa"), "rb").read()
    # C4 has <= 12 hits per contig: the reference's sort is stable there, so even BED line ORDER must agree
    assert common.bed_bytes(d) == open(os.path.join(common.GOLDEN, "C4.bed_flag"), "rb").read()
    assert common.bed_bytes(d, is_walk=True) == open(os.path.join(common.GOLDEN, "C4.bed_walk"), "rb").read()
    assert pgoracle.tie_report() == [0, 0, 0, 0]
    pgoracle.destroy(g); hostio.destroy(d)


def test_oracle_c4_bed_raw(tmp_path):
    opt = hostio.default_opt()
    d = hostio.read_pafs(opt, common.c4_files(tmp_path))
    pgoracle.hot_path(opt, d, stop_after_post=True)
    assert common.bed_bytes(d) == open(os.path.join(common.GOLDEN, "C4.bed_raw"), "rb").read()
    hostio.destroy(d)


@pytest.mark.parametrize("name", sorted(common.synth_golden().keys()))
def test_oracle_synthetic_goldens(name):
    case = common.synth_golden()[name]
    opt, d = common.synth_case_data(case)
    g = pgoracle.hot_path(opt, d)
    gfa = common.gfa_bytes(g)
    assert common.sorted_md5(gfa) == case["gfa_sorted_md5"]
    assert common.sorted_md5(common.bed_bytes(d)) == case["bed_flag_sorted_md5"]
    ls = gfa.splitlines()
    assert (sum(l[:1] == b"S" for l in ls), sum(l[:1] == b"L" for l in ls), sum(l[:1] == b"W" for l in ls)) == (case["n_S"], case["n_L"], case["n_W"])
    pgoracle.destroy(g); hostio.destroy(d)


needs_ref = pytest.mark.skipif(not refharness.available(), reason="oracle/_ref not built (reference sources absent)")

OPTION_SETS = [[], ["-S"], ["-J"], ["-E"], ["-F", "-T2"], ["-T0"], ["--ori-sc"], ["-p0", "-a2"], ["-f0.2", "-c4", "-g6"], ["-b0.1", "-y0.3", "-B0.9", "-r1"]]


@needs_ref
@pytest.mark.parametrize("k", range(len(OPTION_SETS)))
def test_oracle_vs_compiled_reference_randomized(k):
    rnd = random.Random(1000 + k)
    args = OPTION_SETS[k]
    case = dict(config=2, n_genome=rnd.randint(2, 7), n_gene=rnd.randint(300, 1500), iso_per_gene=rnd.choice([1, 4]), hits_per_genome=0, args=args)
    case["hits_per_genome"] = int(case["n_gene"] * case["iso_per_gene"] * rnd.uniform(1.5, 3.0))
    opt, d1 = common.synth_case_data(case)
    _, d2 = common.synth_case_data(case)
    rh = refharness.RefHot()
    g1, _ = rh.hot_path(opt, d1)
    g2 = pgoracle.hot_path(opt, d2)
    a, b = common.gfa_bytes(g1), common.gfa_bytes(g2)
    assert sorted(a.splitlines()) == sorted(b.splitlines()), common.first_diff(a, b)
    assert common.graph_tuple(g1) == common.graph_tuple(g2)
    ba, bb = common.bed_bytes(d1), common.bed_bytes(d2)
    assert sorted(ba.splitlines()) == sorted(bb.splitlines()), common.first_diff(ba, bb)
    rh.destroy(g1); pgoracle.destroy(g2); hostio.destroy(d1); hostio.destroy(d2)


@needs_ref
def test_oracle_bacterial_vs_compiled_reference():
    case = dict(config=5, n_genome=60, n_gene=900, iso_per_gene=1, hits_per_genome=800, args=["-p0", "-a2"])
    opt, d1 = common.synth_case_data(case)
    _, d2 = common.synth_case_data(case)
    rh = refharness.RefHot()
    g1, _ = rh.hot_path(opt, d1)
    g2 = pgoracle.hot_path(opt, d2)
    assert common.gfa_bytes(g1) == common.gfa_bytes(g2)
    rh.destroy(g1); pgoracle.destroy(g2); hostio.destroy(d1); hostio.destroy(d2)


@needs_ref
@pytest.mark.parametrize("n,key_bits", [(10, 8), (64, 64), (65, 3), (1000, 6), (5000, 64), (20000, 20), (20000, 1)])
def test_radix_permutation_replay_matches_reference_sort(n, key_bits):
    """equal keys are common (key_bits small): the PERMUTATION, not just sortedness, must match ksort.h:52-87"""
    ref = C.CDLL(refharness.REF_LIB)
    rnd = random.Random(n * 131 + key_bits)
    arr = (abi.pg_opt_t * 0)()  # noqa: F841 (keep abi imported)
    A = (C.c_uint64 * (2 * n))()
    for i in range(n):
        A[2 * i] = rnd.getrandbits(key_bits) << rnd.choice([0, 0, 8, 29, 32]) & (2 ** 64 - 1)
        A[2 * i + 1] = i
    B = (C.c_uint64 * (2 * n))(*A)
    ref.radix_sort_pg128x(C.byref(A), C.byref(A, 16 * n))
    pgoracle.lib().pgo_radix_replay128(C.byref(B), n)
    assert list(A) == list(B)


@needs_ref
@pytest.mark.parametrize("k,local", [(177, False), (179, False), (463, False), (502, False), (1006, True), (2573, True), (176, True)])
def test_oracle_replays_the_reference_sort_chain_where_ties_are_observable(k, local):
    """Campaign cases (tools/fuzz_oracle.py).  177-502: the first (contig, cs) of a genome is shared by several hits -- the hit at
    the front of the reference's array is never un-shadowed (overlap.c:108), and which hit that is changes from sort to sort of the
    reference's unstable radix-sort chain; a stable-order oracle differs from the reference on all four.  1006, 2573 ("local"
    mode: ordinals decide the locality test): survivors sharing one (contig, cs) take the reference's ordinals (branch.c:18-25);
    176: the same situation where the stable order is the reference's."""
    import sys
    sys.path.insert(0, os.path.join(common.ROOT, "tools"))
    import fuzz_oracle
    case = fuzz_oracle.random_case(k, local)
    opt, d1 = fuzz_oracle.data_for(case)
    _, d2 = fuzz_oracle.data_for(case)
    rh = refharness.RefHot()
    g1, _ = rh.hot_path(opt, d1)
    g2 = pgoracle.hot_path(opt, d2)
    a, b = common.gfa_bytes(g1), common.gfa_bytes(g2)
    assert sorted(a.splitlines()) == sorted(b.splitlines()), common.first_diff(a, b)
    assert common.graph_tuple(g1) == common.graph_tuple(g2)
    ba, bb = common.bed_bytes(d1), common.bed_bytes(d2)
    assert sorted(ba.splitlines()) == sorted(bb.splitlines()), common.first_diff(ba, bb)
    rh.destroy(g1); pgoracle.destroy(g2); hostio.destroy(d1); hostio.destroy(d2)


@needs_ref
@pytest.mark.parametrize("k", [176, 463])
def test_sort_chain_replay_matches_reference_pg_hit_sort(k):
    """pgo_test_chain_order (the replay behind the two tie fixes) against the compiled reference's own pg_hit_sort() applied
    again and again to the same genome: identical array order after every cs sort and every cm sort of the chain."""
    import sys
    sys.path.insert(0, os.path.join(common.ROOT, "tools"))
    import fuzz_oracle
    case = fuzz_oracle.random_case(k)
    _, d = fuzz_oracle.data_for(case)
    _, d2 = fuzz_oracle.data_for(case)
    ref = C.CDLL(refharness.REF_LIB)
    ref.pg_hit_sort.argtypes = [C.POINTER(abi.pg_genome_t), C.c_int32]
    L = pgoracle.lib()
    L.pgo_test_chain_order.argtypes = [C.POINTER(abi.pg_genome_t), C.c_int32, C.c_int32, C.c_void_p]
    for j in range(min(d.contents.n_genome, 6)):
        g, g2 = d.contents.genome[j], d2.contents.genome[j]
        n = g.n_hit
        for i in range(n):
            g.hit[i].lof = i                     # tag every record with its file index
        gp = C.pointer(g)
        ref.pg_hit_sort(gp, 0)                   # read.c:247
        for pairs in range(5):
            out = (C.c_int32 * n)()
            L.pgo_test_chain_order(C.byref(g2), pairs, 0, out)
            assert list(out) == [g.hit[i].lof for i in range(n)], (j, pairs, "cs")
            ref.pg_hit_sort(gp, 1)               # graph.c:103 / branch.c:116
            L.pgo_test_chain_order(C.byref(g2), pairs, 1, out)
            assert list(out) == [g.hit[i].lof for i in range(n)], (j, pairs, "cm")
            ref.pg_hit_sort(gp, 0)               # graph.c:123 / branch.c:140
    hostio.destroy(d); hostio.destroy(d2)


@needs_ref
@pytest.mark.parametrize("k", [16, 97, 110])
def test_oracle_replays_cm_tied_survivors_like_the_reference(k):
    """tools/fuzz_oracle.py "cmtie" cases: survivors forced to share one (contig, cm); the order the reference walks them in comes out of its
    unstable sort chain (graph.c:103, branch.c:116) and decides arcs -- the oracle replays the chain (survivors_in_ref_cm_order)"""
    import sys
    sys.path.insert(0, os.path.join(common.ROOT, "tools"))
    import fuzz_oracle
    case = fuzz_oracle.random_case(k, cmtie=True)
    opt, d1 = fuzz_oracle.data_for(case)
    _, d2 = fuzz_oracle.data_for(case)
    rh = refharness.RefHot()
    g1, _ = rh.hot_path(opt, d1)
    g2 = pgoracle.hot_path(opt, d2)
    assert pgoracle.tie_report()[0] > 0
    assert common.graph_tuple(g1) == common.graph_tuple(g2)
    assert sorted(common.gfa_bytes(g1).splitlines()) == sorted(common.gfa_bytes(g2).splitlines())
    rh.destroy(g1); pgoracle.destroy(g2); hostio.destroy(d1); hostio.destroy(d2)

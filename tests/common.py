"""Shared helpers for the parity tests (TEST INFRASTRUCTURE)."""
import ctypes as C
import hashlib
import json
import os
import tempfile

from pangene_b200 import abi, hostio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def c4_files(tmpdir):
    """Unpack tests/golden/C4_inputs.txt (the reference's test/C4 example) into one PAF per genome."""
    files, fp = [], None
    for line in open(os.path.join(GOLDEN, "C4_inputs.txt"), "rb"):
        if line.startswith(b">"):
            if fp:
                fp.close()
            path = os.path.join(str(tmpdir), line[1:].strip().decode())
            files.append(path)
            fp = open(path, "wb")
        else:
            fp.write(line)
    if fp:
        fp.close()
    return files


def synth_golden():
    return json.load(open(os.path.join(GOLDEN, "synth_golden.json")))


def opt_from_args(args):
    """pg_opt_t for a list of reference CLI flags (subset used by the tests)."""
    o = hostio.default_opt()
    for a in args:
        if a == "-S": o.flag |= abi.PG_F_CHECK_STRAND
        elif a == "-F": o.flag |= abi.PG_F_FRAG_MODE
        elif a == "-J": o.flag |= abi.PG_F_NO_JOINT_PSEUDO
        elif a == "-E": o.flag |= abi.PG_F_DROP_SGL_EXON
        elif a == "--ori-sc": o.flag |= abi.PG_F_ORI_FOR_BRANCH
        elif a.startswith("-T"): o.n_branch_flt = int(a[2:])
        elif a.startswith("-p"): o.min_vertex_ratio = float(a[2:])
        elif a.startswith("-a"): o.min_arc_cnt = int(a[2:])
        elif a.startswith("-f"): o.min_ov_ratio = float(a[2:])
        elif a.startswith("-c"): o.max_avg_occ = int(a[2:])
        elif a.startswith("-g"): o.max_degree = int(a[2:])
        elif a.startswith("-r"): o.max_dist_loci = int(a[2:])
        elif a.startswith("-b"): o.branch_diff = float(a[2:])
        elif a.startswith("-y"): o.branch_diff_dist = float(a[2:])
        elif a.startswith("-B"): o.branch_diff_cut = float(a[2:])
        else: raise ValueError(a)
    return o


def synth_case_data(case, opt=None):
    opt = opt or opt_from_args(case["args"])
    p = hostio.synth_params(case["config"], n_genome=case["n_genome"], n_gene=case["n_gene"], iso_per_gene=case["iso_per_gene"],
                            hits_per_genome=case["hits_per_genome"], n_genome_total=case["n_genome"])
    return opt, hostio.synth(opt, p)


def gfa_bytes(graph, walk=True):
    with tempfile.NamedTemporaryFile(suffix=".gfa") as t:
        hostio.write_gfa(graph, t.name, True, walk)
        return open(t.name, "rb").read()


def bed_bytes(d, is_walk=False):
    with tempfile.NamedTemporaryFile(suffix=".bed") as t:
        hostio.write_bed(d, t.name, is_walk)
        return open(t.name, "rb").read()


def sorted_md5(data):
    return hashlib.md5(b"\n".join(sorted(data.splitlines()))).hexdigest()


def graph_tuple(g):
    """(segments, arcs) as plain tuples for struct-level comparison."""
    q = g.contents
    segs = [(s.gid, s.n_dom, s.n_sub, s.n_genome, s.tot_cnt) for s in (q.seg[i] for i in range(q.n_seg))]
    arcs = [(a.x, a.n_genome, a.tot_cnt, a.avg_dist, a.s1, a.s2) for a in (q.arc[i] for i in range(q.n_arc))]
    return segs, arcs


def first_diff(a, b, n=5):
    out = []
    sa, sb = sorted(a.splitlines()), sorted(b.splitlines())
    only_a, only_b = sorted(set(sa) - set(sb)), sorted(set(sb) - set(sa))
    for x in only_a[:n]: out.append("- " + x.decode(errors="replace"))
    for x in only_b[:n]: out.append("+ " + x.decode(errors="replace"))
    return "\n".join(out) + "\n(%d lines only in expected, %d only in got)" % (len(only_a), len(only_b))


def tokens_from_hits(d):
    """Walk tokens (include/pgx_walk.h) built from MATERIALISED hits the way format.c:183-225 picks them: the unfiltered hits of a genome,
    contig-major, by cm inside a contig, ties in array order.  Returns (ctypes token array, genome_off array, n)."""
    import numpy as np
    from pangene_b200 import engine
    dd = d.contents
    HIT = np.dtype([(n, "<i4") for n in ("pid", "qs", "qe", "cid", "mlen", "blen", "lof", "rank", "score_ori", "score_adj", "score_dom", "n_exon", "off_exon", "pid_dom", "pid_dom0")]
                   + [("bits", "<u4"), ("cs", "<i8"), ("cm", "<i8"), ("ce", "<i8")])
    gid = np.array([dd.prot[i].gid for i in range(dd.n_prot)], dtype=np.int64)
    parts, off = [], [0]
    for j in range(dd.n_genome):
        g = dd.genome[j]
        if g.n_hit == 0:
            off.append(off[-1]); continue
        h = np.frombuffer((C.c_char * (g.n_hit * 88)).from_address(C.addressof(g.hit.contents)), dtype=HIT)
        keep = np.nonzero((h["bits"] & abi.PG_HB_FLT) == 0)[0]
        order = keep[np.lexsort((h["cm"][keep], h["cid"][keep]))]   # lexsort is stable: ties keep array order
        t = np.zeros(len(order), dtype=[("cid", "<i4"), ("gene_strand", "<u4"), ("lof", "<i4")])
        t["cid"] = h["cid"][order]; t["gene_strand"] = (gid[h["pid"][order]] << 1) | (h["bits"][order] & 1); t["lof"] = h["lof"][order]
        parts.append(t); off.append(off[-1] + len(t))
    n = off[-1]
    buf = (engine.pgx_walk_token_t * max(n, 1))()
    if n:
        allt = np.concatenate(parts)
        C.memmove(buf, allt.ctypes.data, n * 12)
    return buf, (C.c_int64 * len(off))(*off), n


def gfa_bytes_tokens(graph, tok, off, walk=True):
    with tempfile.NamedTemporaryFile(suffix=".gfa") as t:
        hostio.write_gfa_tokens(graph, tok, off, t.name, True, walk)
        return open(t.name, "rb").read()


def matrix_bytes_tokens(graph, tok, off, copy_number=False):
    with tempfile.NamedTemporaryFile(suffix=".tsv") as t:
        hostio.matrix_from_tokens(graph, tok, off, t.name, copy_number)
        return open(t.name, "rb").read()


def matrix_bytes_gfa(gfa_data, copy_number=False):
    with tempfile.NamedTemporaryFile(suffix=".gfa") as a, tempfile.NamedTemporaryFile(suffix=".tsv") as b:
        a.write(gfa_data); a.flush()
        hostio.lib().pgio_gfa2matrix(os.fsencode(a.name), os.fsencode(b.name), 1 if copy_number else 0)
        return open(b.name, "rb").read()

"""ctypes binding of libpgio.so -- PAF reader, GFA/BED writers, synthetic generator (include/pgio.h)."""
import ctypes as C
import os
from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "csrc", "libpgio.so")
        if not os.path.exists(path):
            raise RuntimeError("libpgio.so is not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(path, mode=C.RTLD_LOCAL)
        P = C.POINTER
        L.pgio_opt_init.argtypes = [P(abi.pg_opt_t)]
        L.pgio_data_init.restype = P(abi.pg_data_t)
        L.pgio_data_destroy.argtypes = [P(abi.pg_data_t)]
        L.pgio_read_paf.argtypes = [P(abi.pg_opt_t), P(abi.pg_data_t), C.c_char_p]
        L.pgio_read_paf.restype = C.c_int32
        L.pgio_read_pafs.argtypes = [P(abi.pg_opt_t), P(abi.pg_data_t), P(C.c_char_p), C.c_int32, C.c_int32]
        L.pgio_read_pafs.restype = C.c_int32
        L.pgio_read_pafs_packed.argtypes = [P(abi.pg_opt_t), P(abi.pg_data_t), P(abi.pgx_packed_t), P(C.c_char_p), C.c_int32, C.c_int32]
        L.pgio_read_pafs_packed.restype = C.c_int32
        L.pgio_packed_destroy.argtypes = [P(abi.pgx_packed_t)]
        L.pgio_packed_total_hits.argtypes = [P(abi.pgx_packed_t)]
        L.pgio_packed_total_hits.restype = C.c_int64
        L.pgio_name_set.argtypes = [C.c_char_p]
        L.pgio_name_set.restype = C.c_void_p
        L.pgio_name_set_destroy.argtypes = [C.c_void_p]
        L.pgio_write_gfa.argtypes = [P(abi.pg_graph_t), C.c_char_p, C.c_int32]
        L.pgio_write_gfa.restype = C.c_int32
        L.pgio_write_bed.argtypes = [P(abi.pg_data_t), C.c_int32, C.c_char_p]
        L.pgio_write_bed.restype = C.c_int32
        L.pgio_write_gfa_tokens.argtypes = [P(abi.pg_graph_t), C.c_void_p, P(C.c_int64), C.c_char_p, C.c_int32]
        L.pgio_write_gfa_tokens.restype = C.c_int32
        L.pgio_matrix_from_tokens.argtypes = [P(abi.pg_graph_t), C.c_void_p, P(C.c_int64), C.c_char_p, C.c_int32]
        L.pgio_matrix_from_tokens.restype = C.c_int32
        L.pgio_gfa2matrix.argtypes = [C.c_char_p, C.c_char_p, C.c_int32]
        L.pgio_gfa2matrix.restype = C.c_int32
        L.pgio_synth_default.argtypes = [P(abi.pgio_synth_t), C.c_int32]
        L.pgio_synth.argtypes = [P(abi.pg_opt_t), P(abi.pgio_synth_t)]
        L.pgio_synth.restype = P(abi.pg_data_t)
        L.pgio_synth_write_paf.argtypes = [P(abi.pg_opt_t), P(abi.pgio_synth_t), C.c_int32, C.c_char_p]
        L.pgio_synth_write_paf.restype = C.c_int32
        L.pgio_data_clone.argtypes = [P(abi.pg_data_t)]
        L.pgio_data_clone.restype = P(abi.pg_data_t)
        L.pgio_data_clone_destroy.argtypes = [P(abi.pg_data_t)]
        L.pgio_total_hits.argtypes = [P(abi.pg_data_t)]
        L.pgio_total_hits.restype = C.c_int64
        _LIB = L
    return _LIB


def default_opt(**kw):
    """pg_opt_t with the reference defaults (option.c:6-26); keyword overrides by field name."""
    o = abi.pg_opt_t()
    lib().pgio_opt_init(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def read_pafs(opt, files, threads=1):
    """threads=1: file by file (pgio_read_paf); otherwise the parallel reader (pgio_read_pafs, 0 = all cores)"""
    d = lib().pgio_data_init()
    if threads == 1:
        for f in files:
            lib().pgio_read_paf(C.byref(opt), d, os.fsencode(f))
    else:
        arr = (C.c_char_p * len(files))(*[os.fsencode(f) for f in files])
        lib().pgio_read_pafs(C.byref(opt), d, arr, len(files), threads)
    return d


def read_pafs_packed(opt, files, threads=0):
    """the parallel reader emitting device-ready records (include/pgx_packed.h): returns (pg_data_t* without hit arrays, pgx_packed_t)"""
    d = lib().pgio_data_init()
    pk = abi.pgx_packed_t()
    arr = (C.c_char_p * len(files))(*[os.fsencode(f) for f in files])
    lib().pgio_read_pafs_packed(C.byref(opt), d, C.byref(pk), arr, len(files), threads)
    return d, pk


def synth_params(config=2, **kw):
    p = abi.pgio_synth_t()
    lib().pgio_synth_default(C.byref(p), config)
    for k, v in kw.items():
        setattr(p, k, v)
    if "n_genome" in kw and "n_genome_total" not in kw:
        p.n_genome_total = max(p.n_genome_total if "genome_begin" in kw else 0, p.genome_begin + p.n_genome)
    return p


def synth(opt, params):
    return lib().pgio_synth(C.byref(opt), C.byref(params))


def synth_write_pafs(opt, params, outdir):
    os.makedirs(outdir, exist_ok=True)
    files = []
    for j in range(params.genome_begin, params.genome_begin + params.n_genome):
        f = os.path.join(outdir, "S%04d.paf" % j)
        lib().pgio_synth_write_paf(C.byref(opt), C.byref(params), j, os.fsencode(f))
        files.append(f)
    return files


def total_hits(d):
    return int(lib().pgio_total_hits(d))


def write_gfa(graph, path, seg_arc=True, walk=True):
    return lib().pgio_write_gfa(graph, os.fsencode(path), (1 if seg_arc else 0) | (2 if walk else 0))


def write_gfa_tokens(graph, tok, off, path, seg_arc=True, walk=True):
    return lib().pgio_write_gfa_tokens(graph, C.cast(tok, C.c_void_p), off, os.fsencode(path), (1 if seg_arc else 0) | (2 if walk else 0))


def matrix_from_tokens(graph, tok, off, path, copy_number=False):
    return lib().pgio_matrix_from_tokens(graph, C.cast(tok, C.c_void_p), off, os.fsencode(path), 1 if copy_number else 0)


def write_bed(d, path, is_walk=False):
    return lib().pgio_write_bed(d, 1 if is_walk else 0, os.fsencode(path))


def gfa_digest(graph):
    """(md5 of the S and L lines in writer order, the W lines of the local genomes as bytes) -- what bench.py and the tests
    compare with tests/golden/c3_golden.json; a sharded run concatenates the W bytes in rank order before hashing."""
    import hashlib
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".gfa") as t:
        write_gfa(graph, t.name, True, False)
        sl = hashlib.md5(open(t.name, "rb").read()).hexdigest()
        write_gfa(graph, t.name, False, True)
        w = open(t.name, "rb").read()
    return sl, w


def gfa_digest_tokens(graph, tok, off):
    """gfa_digest() of the same GFA, formatted from walk tokens (pgio_write_gfa_tokens) instead of materialised hits"""
    import hashlib
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".gfa") as t:
        write_gfa_tokens(graph, tok, off, t.name, True, False)
        sl = hashlib.md5(open(t.name, "rb").read()).hexdigest()
        write_gfa_tokens(graph, tok, off, t.name, False, True)
        w = open(t.name, "rb").read()
    return sl, w


def destroy(d):
    lib().pgio_data_destroy(d)

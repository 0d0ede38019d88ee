"""ctypes binding of libpgx.so -- the sm_100a hot path behind the reference's own entry points (include/pgx.h).

There is no Python-side computation and no fallback: if the CUDA extension is missing or no Blackwell device is
visible, the calls fail loudly (RuntimeError here, abort() with a message inside the library).
"""
import ctypes as C
import os
from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# PGX_LIB: another build of the SAME library source -- an experimental libpgx.so for A/B timing, or the test suite's host build on the SIMT shim
# (tests/emu, tests/test_emulated_path.py).  Nothing in the product sets it; bench.py prints the path it ran (detail.library), the CLI links csrc/libpgx.so.
LIB_PATH = os.environ.get("PGX_LIB") or os.path.join(_HERE, "csrc", "libpgx.so")
_LIB = None


class pgx_walk_token_t(C.Structure):
    """include/pgx_walk.h: one printed hit of a W line"""
    _fields_ = [("cid", C.c_int32), ("gene_strand", C.c_uint32), ("lof", C.c_int32)]


class pgx_stats_t(C.Structure):
    _fields_ = [("launches", C.c_uint64), ("sweep_ms", C.c_double), ("sweep_bytes", C.c_double),
                ("run_ms", C.c_double * 3), ("n_hit", C.c_int64), ("n_active", C.c_int64),
                ("tie", C.c_int64 * 4), ("n_seg", C.c_int32), ("n_arc", C.c_int32),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("fast_rounds", C.c_int32), ("redo", C.c_int32)]


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libpgx.so (the CUDA hot path) is not built: run __graft_entry__.build(). "
                               "There is no CPU fallback.")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
        P = C.POINTER
        L.pg_post_process.argtypes = [P(abi.pg_opt_t), P(abi.pg_data_t)]
        L.pg_post_process.restype = None
        L.pg_graph_init.argtypes = [P(abi.pg_data_t)]
        L.pg_graph_init.restype = P(abi.pg_graph_t)
        L.pg_graph_gen.argtypes = [P(abi.pg_opt_t), P(abi.pg_graph_t)]
        L.pg_graph_gen.restype = None
        L.pg_graph_destroy.argtypes = [P(abi.pg_graph_t)]
        L.pgx_data_release.argtypes = [P(abi.pg_data_t)]
        L.pgx_walk_tokens.argtypes = [P(abi.pg_data_t), P(P(pgx_walk_token_t)), P(C.c_int64)]
        L.pgx_walk_tokens.restype = C.c_int64
        L.pgx_fetch_hits.argtypes = [P(abi.pg_data_t)]
        L.pgx_post_process_packed.argtypes = [P(abi.pg_opt_t), P(abi.pg_data_t), P(abi.pgx_packed_t)]
        L.pgx_device_count.restype = C.c_int
        L.pgx_set_device.argtypes = [C.c_int]
        L.pgx_set_lazy.argtypes = [C.c_int]
        L.pgx_set_sweep_stat.argtypes = [C.c_int]
        L.pgx_comm_unique_id.argtypes = [C.c_void_p]
        L.pgx_comm_init.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int32, C.c_int32]
        L.pgx_comm_init_loopback.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int32, C.c_int32]
        L.pgx_upload.argtypes = [P(abi.pg_data_t)]
        L.pgx_upload.restype = C.c_void_p
        L.pgx_run.argtypes = [C.c_void_p, P(abi.pg_opt_t), C.c_int]
        L.pgx_download.argtypes = [C.c_void_p, P(abi.pg_data_t), P(abi.pg_graph_t)]
        L.pgx_session_free.argtypes = [C.c_void_p]
        L.pgx_get_stats.argtypes = [P(pgx_stats_t)]
        L.pgx_trim.restype = None
        L.pgx_test_radix_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        L.pgx_test_exclusive_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.pgx_test_segmented_max.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.pgx_test_flag_sort_replay.argtypes = [C.c_void_p, C.c_int64]
        L.pgx_test_choose_representatives.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        L.pgx_test_choose_representatives.restype = C.c_int
        _LIB = L
    return _LIB


def set_verbose(v):
    C.c_int.in_dll(lib(), "pg_verbose").value = v


def hot_path(opt, d, stop_after_post=False, lazy=False):
    """The reference's call sequence (main.c:121-127) on the GPU: pg_post_process, pg_graph_init, pg_graph_gen."""
    L = lib()
    L.pgx_set_lazy(1 if lazy else 0)
    L.pg_post_process(C.byref(opt), d)
    if stop_after_post:
        L.pgx_data_release(d)
        return None
    g = L.pg_graph_init(d)
    L.pg_graph_gen(C.byref(opt), g)
    return g


def graph_and_tokens(opt, d, packed=None):
    """The GFA-only route (lazy mode 2): the graph comes back, the hits stay on the device and only the walk tokens of the W lines
    (12 bytes per printed hit) are read.  packed: the hits as device-ready records (abi.pgx_packed_t from hostio.read_pafs_packed)
    instead of d's pg_hit_t arrays.  Returns (graph, tokens pointer, genome_off array, n_tokens); free with release_tokens()."""
    L = lib()
    L.pgx_set_lazy(2)
    try:
        if packed is not None:
            L.pgx_post_process_packed(C.byref(opt), d, C.byref(packed))
        else:
            L.pg_post_process(C.byref(opt), d)
        g = L.pg_graph_init(d)
        L.pg_graph_gen(C.byref(opt), g)
        tok = C.POINTER(pgx_walk_token_t)()
        off = (C.c_int64 * (d.contents.n_genome + 1))()
        n = L.pgx_walk_tokens(d, C.byref(tok), off)
    finally:
        L.pgx_set_lazy(0)
    return g, tok, off, n


def release_tokens(d, tok):
    C.CDLL(None).free(tok)
    lib().pgx_data_release(d)


def stats():
    s = pgx_stats_t()
    lib().pgx_get_stats(C.byref(s))
    return s


def destroy(g):
    lib().pg_graph_destroy(g)

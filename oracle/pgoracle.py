"""TEST INFRASTRUCTURE: ctypes binding of oracle/libpgoracle.so (the plain-C restatement, pg_oracle.c).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from pangene_b200 import abi  # noqa: E402

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(os.path.join(_HERE, "libpgoracle.so"), mode=C.RTLD_LOCAL)
        P = C.POINTER
        L.pgo_post_process.argtypes = [P(abi.pg_opt_t), P(abi.pg_data_t)]
        L.pgo_graph_init.argtypes = [P(abi.pg_data_t)]
        L.pgo_graph_init.restype = P(abi.pg_graph_t)
        L.pgo_graph_gen.argtypes = [P(abi.pg_opt_t), P(abi.pg_graph_t)]
        L.pgo_graph_destroy.argtypes = [P(abi.pg_graph_t)]
        L.pgo_radix_replay128.argtypes = [C.c_void_p, C.c_int64]
        L.pgo_tie_report.argtypes = [P(C.c_int64)]
        _LIB = L
    return _LIB


def hot_path(opt, d, stop_after_post=False):
    """block of read.c:243-260 + pg_post_process + pg_graph_init + pg_graph_gen, restated; returns the graph"""
    L = lib()
    L.pgo_post_process(C.byref(opt), d)
    if stop_after_post:
        return None
    g = L.pgo_graph_init(d)
    L.pgo_graph_gen(C.byref(opt), g)
    return g


def tie_report():
    out = (C.c_int64 * 4)()
    lib().pgo_tie_report(out)
    return list(out)


def destroy(g):
    lib().pgo_graph_destroy(g)

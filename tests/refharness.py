"""TEST INFRASTRUCTURE: access to the compiled, unmodified reference under oracle/_ref/.

  * ref_cli()   runs the stock `pangene` binary (oracle/_ref/pangene) on PAF files
  * RefHot      ctypes view of oracle/_ref/libpgref.so: the reference's own hot-path functions driven
                by oracle/ref_hot.c on a pg_data_t produced by OUR reader/generator

Nothing in the product imports this module.
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pangene_b200 import abi  # noqa: E402

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_BIN = os.path.join(REF_DIR, "pangene")
REF_LIB = os.path.join(REF_DIR, "libpgref.so")


def available():
    return os.path.exists(REF_BIN) and os.path.exists(REF_LIB)


def ref_cli(files, args=(), stdout_path=None):
    """stdout of `pangene [args] files...` (the unmodified reference CLI)."""
    cmd = [REF_BIN] + list(args) + list(files)
    if stdout_path is None:
        return subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    with open(stdout_path, "wb") as fp:
        subprocess.run(cmd, check=True, stdout=fp, stderr=subprocess.DEVNULL)
    return None


class RefHot:
    def __init__(self):
        L = C.CDLL(REF_LIB, mode=C.RTLD_LOCAL)
        P = C.POINTER
        L.pgref_hot_path.argtypes = [P(abi.pg_opt_t), P(abi.pg_data_t), P(C.c_double), C.c_int]
        L.pgref_hot_path.restype = P(abi.pg_graph_t)
        L.pgref_graph_destroy.argtypes = [P(abi.pg_graph_t)]
        L.pgref_write.argtypes = [P(abi.pg_graph_t), P(abi.pg_data_t), C.c_int, C.c_char_p]
        L.pgref_set_verbose.argtypes = [C.c_int]
        self.L = L
        L.pgref_set_verbose(0)

    def hot_path(self, opt, d, stop_after_post=False):
        t = (C.c_double * 3)()
        # the two unconditional stderr lines of pg_post_process (graph.c:26,31) are silenced by pointing fd 2 away
        sys.stderr.flush()
        saved = os.dup(2)
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 2)
        try:
            g = self.L.pgref_hot_path(C.byref(opt), d, t, 1 if stop_after_post else 0)
        finally:
            os.dup2(saved, 2)
            os.close(saved)
            os.close(devnull)
        return g, list(t)

    def write(self, g, d, mode, path):
        sys.stdout.flush()
        return self.L.pgref_write(g, d, mode, os.fsencode(path))

    def destroy(self, g):
        self.L.pgref_graph_destroy(g)


def sorted_lines(path_or_bytes):
    data = path_or_bytes if isinstance(path_or_bytes, bytes) else open(path_or_bytes, "rb").read()
    return sorted(data.splitlines())

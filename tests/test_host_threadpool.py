"""CPU suite: the host thread pool of libpgx (pgx_hostpipe.hpp) under ThreadSanitizer -- back-to-back tiny parallel_for calls
(a late-waking worker must never pair the cursor of one call with the function object of another)."""
import os
import subprocess

import common

SRC = os.path.join(common.ROOT, "tests", "host", "threadpool_stress.cpp")
INC = ["-I" + os.path.join(common.ROOT, "pangene_b200", "csrc"), "-I" + os.path.join(common.ROOT, "include"), "-I/usr/local/cuda/include"]


def _build(tmp_path, extra):
    exe = str(tmp_path / "tp")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g"] + extra + INC + [SRC, "-o", exe, "-L/usr/local/cuda/lib64", "-lcudart", "-lpthread",
                        "-Wl,-rpath,/usr/local/cuda/lib64"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return exe if r.returncode == 0 else None, r.stdout.decode()


def test_threadpool_back_to_back_small_jobs(tmp_path):
    exe, log = _build(tmp_path, ["-fsanitize=thread"])
    tsan = exe is not None
    if exe is None:                      # no libtsan in the image: the plain build still checks exactly-once execution
        exe, log = _build(tmp_path, [])
    assert exe, log
    env = dict(os.environ, PGX_THREADS="8", TSAN_OPTIONS="halt_on_error=1")
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env, timeout=300)
    out = r.stdout.decode()
    assert r.returncode == 0 and "bad=0" in out and (not tsan or "WARNING: ThreadSanitizer" not in out), out[-3000:]

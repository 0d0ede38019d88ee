""""racecheck" of the tiled sweep kernels without a GPU: tools/tsan_sweep.sh builds the kernel text of pangene_b200/csrc/pgx_kernels.cuh on the
SIMT shim (OS threads, real barriers, __atomic builtins) under ThreadSanitizer and runs every variant in every mode -- a shared-memory access
pair that no __syncthreads() orders is a reported data race (checked by mutation: dropping one barrier of the cluster pass gives 17 reports
while the results stay right)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sweep_kernels_have_no_shared_memory_race():
    tsan = subprocess.run(["gcc", "-print-file-name=libtsan.so"], stdout=subprocess.PIPE, text=True).stdout.strip()
    if not shutil.which("g++") or not os.path.exists(tsan):
        pytest.skip("ThreadSanitizer runtime not installed")
    r = subprocess.run(["bash", os.path.join(ROOT, "tools", "tsan_sweep.sh"), "700", "5"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "TSAN_SWEEP PASS" in r.stdout and "ThreadSanitizer reports: 0" in r.stdout, r.stdout[-3000:]

"""The device primitives' control flow, checked on the CPU: tests/emu/ compiles the radix-sort, prefix-sum and segmented-max kernels
of pangene_b200/csrc/pgx_sort.cuh / pgx_scan.cuh for the host (SIMT shim: one OS thread per CUDA thread, one CTA at a time) and this
test runs them, with the launch sequences of the product, against numpy.  The GPU suite checks the same primitives on the device
(tests/test_gpu_primitives.py); here the point is that tile boundaries, ragged tails, the stability of the sort (ties keep their
incoming order: DESIGN.md "tie order") and the chained look-back of the single-pass scan are exercised without a GPU."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu import build as emu_build  # noqa: E402


@pytest.fixture(scope="module")
def emu():
    L = C.CDLL(emu_build.build_prim())
    L.prim_emu_scan.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
    L.prim_emu_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int]
    L.prim_emu_segmax.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    return L


@pytest.mark.parametrize("variant,n", [(0, 1), (0, 2047), (0, 2048), (0, 5000), (1, 1), (1, 2048), (1, 4095), (1, 70001), (2, 0), (2, 1), (2, 1023), (2, 1024), (2, 3000)])
def test_exclusive_sum(emu, variant, n):
    """variant 0: three launches (tiles, aggregates, add); 1: single pass with decoupled look-back; 2: one-block scan for small inputs"""
    rng = np.random.default_rng(100 * variant + n)
    a = rng.integers(0, 7, size=n, dtype=np.uint32)
    out = np.full(n + 1 + 8, 0xdeadbeef, dtype=np.uint32)
    assert emu.prim_emu_scan(variant, a.ctypes.data, out.ctypes.data, n) == 0
    want = np.concatenate([[0], np.cumsum(a, dtype=np.uint64)]).astype(np.uint32)
    assert np.array_equal(out[:n + 1], want)
    assert np.all(out[n + 1:] == 0xdeadbeef)            # nothing written past out[n]


def test_exclusive_sum_in_place(emu):
    a = np.arange(1, 6001, dtype=np.uint32)
    buf = np.concatenate([a, np.zeros(1, np.uint32)])
    assert emu.prim_emu_scan(1, buf.ctypes.data, buf.ctypes.data, len(a)) == 0   # the radix sort scans its histogram in place
    assert np.array_equal(buf, np.concatenate([[0], np.cumsum(a, dtype=np.uint64)]).astype(np.uint32))


@pytest.mark.parametrize("n,n_bits,key_bits", [(1, 8, 8), (33, 8, 3), (4096, 16, 16), (4097, 20, 5), (10000, 38, 38), (12289, 64, 64), (9000, 32, 1)])
def test_radix_sort_pairs_is_a_stable_sort(emu, n, n_bits, key_bits):
    rng = np.random.default_rng(n + n_bits)
    keys = rng.integers(0, 1 << min(key_bits, 62), size=n, dtype=np.uint64)
    if key_bits == 64:
        keys |= rng.integers(0, 4, size=n, dtype=np.uint64) << np.uint64(62)
    k = keys.copy()
    v = np.zeros(n, dtype=np.uint32)                   # "iota" values: the first pass ignores vals_in
    assert emu.prim_emu_sort(k.ctypes.data, v.ctypes.data, n, n_bits, 1) == 0
    want = np.argsort(keys, kind="stable")
    assert np.array_equal(v, want.astype(np.uint32))   # the permutation, not just sortedness: equal keys keep their order
    assert np.array_equal(k, keys[want])


def test_radix_sort_low_bits_only(emu):
    """n_bits smaller than the key width sorts by the low bits only and keeps the rest in input order (LSD passes are stable)"""
    rng = np.random.default_rng(7)
    keys = rng.integers(0, 1 << 24, size=6000, dtype=np.uint64)
    k, v = keys.copy(), np.zeros(6000, dtype=np.uint32)
    assert emu.prim_emu_sort(k.ctypes.data, v.ctypes.data, 6000, 8, 0) == 0
    assert np.array_equal(v, np.argsort(keys & np.uint64(0xff), kind="stable").astype(np.uint32))


@pytest.mark.parametrize("n", [1, 2048, 2049, 9000])
def test_segmented_running_max(emu, n):
    rng = np.random.default_rng(n)
    seg = np.sort(rng.integers(0, max(2, n // 700), size=n)).astype(np.int32)
    if n > 4100:
        seg[2040:4100] = seg[2040]                     # one segment spanning three tiles
        seg = np.sort(seg)
    val = rng.integers(-1000, 1000, size=n).astype(np.int32)
    out = np.zeros(n, dtype=np.int32)
    assert emu.prim_emu_segmax(seg.ctypes.data, val.ctypes.data, out.ctypes.data, n) == 0
    want = np.empty(n, dtype=np.int32)
    for i in range(n):
        want[i] = val[i] if i == 0 or seg[i] != seg[i - 1] else max(want[i - 1], val[i])
    assert np.array_equal(out, want)

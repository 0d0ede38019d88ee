"""GPU suite, part 1: the hand-written device primitives (radix sort, scans) against numpy."""
import ctypes as C

import numpy as np
import pytest

from pangene_b200 import engine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,bits", [(1, 8), (37, 8), (4096, 16), (4097, 24), (100000, 40), (1500000, 52), (300000, 1)])
def test_radix_sort_is_a_stable_sort(n, bits):
    rng = np.random.default_rng(n + bits)
    keys = rng.integers(0, 2 ** bits, size=n, dtype=np.uint64)
    if n > 1000:
        keys[:: 7] = keys[0]           # many duplicates: stability must show
    k = keys.copy()
    v = np.zeros(n, dtype=np.uint32)
    engine.lib().pgx_test_radix_sort(k.ctypes.data, v.ctypes.data, n, bits)
    order = np.argsort(keys, kind="stable").astype(np.uint32)
    assert np.array_equal(v, order)
    assert np.array_equal(k, keys[order])


@pytest.mark.parametrize("n", [1, 2047, 2048, 2049, 100000, 5000000])
def test_exclusive_sum(n):
    rng = np.random.default_rng(n)
    a = rng.integers(0, 9, size=n, dtype=np.uint32)
    out = np.zeros(n + 1, dtype=np.uint32)
    engine.lib().pgx_test_exclusive_sum(a.ctypes.data, out.ctypes.data, n)
    ref = np.concatenate([[0], np.cumsum(a, dtype=np.uint64)]).astype(np.uint32)
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("n,nseg", [(1, 1), (5000, 1), (5000, 700), (3000000, 2400), (3000000, 900000)])
def test_segmented_running_max(n, nseg):
    rng = np.random.default_rng(n + nseg)
    seg = np.sort(rng.integers(0, nseg, size=n)).astype(np.int32)
    val = rng.integers(0, 2 ** 30, size=n).astype(np.int32)
    out = np.zeros(n, dtype=np.int32)
    engine.lib().pgx_test_segmented_max(seg.ctypes.data, val.ctypes.data, out.ctypes.data, n)
    ref = val.copy()
    start = np.flatnonzero(np.concatenate([[True], seg[1:] != seg[:-1]]))
    for s, e in zip(start, list(start[1:]) + [n]):
        ref[s:e] = np.maximum.accumulate(val[s:e])
    assert np.array_equal(out, ref)

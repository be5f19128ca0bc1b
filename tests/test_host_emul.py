"""CPU tier: the kernels' per-element math (csrc/devmath.cuh compiled for the host) against the oracle,
glibc and libstdc++.  This is what lets kernel arithmetic be checked in a container without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul():
    from _emul import build_and_load
    L = build_and_load()
    L.emul_fast_x2.restype = C.c_uint32
    L.emul_fast_x2.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.emul_atan2.restype = C.c_float
    L.emul_atan2.argtypes = [C.c_float, C.c_float]
    L.emul_sincos_mismatch.restype = C.c_long
    L.emul_sincos_mismatch.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
    L.emul_pretest_violations.restype = C.c_long
    L.emul_pretest_violations.argtypes = [C.c_uint32, C.c_long, C.c_int, C.c_int, C.POINTER(C.c_long)]
    L.emul_logf_mismatch.restype = C.c_long
    L.emul_logf_mismatch.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
    for f in ("emul_std_sort", "emul_heap_sort", "ref_std_sort", "ref_heap_sort"):
        getattr(L, f).argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.emul_path_key.restype = C.c_uint32
    L.emul_path_key.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int]
    return L


def test_fast_score_x2_matches_oracle(emul):
    img = synth.frame(160, 120, 5)
    rng = np.random.default_rng(0)
    # add extreme-contrast pixels so the +/-255 corners of the packed arithmetic are exercised
    img[rng.integers(0, 120, 300), rng.integers(0, 160, 300)] = 255
    img[rng.integers(0, 120, 300), rng.integers(0, 160, 300)] = 0
    win = np.ascontiguousarray(img)
    ref = {}
    cands = po.fast_cell(win, -300)  # threshold below any score: every tested local max
    for y in range(3, 117, 7):
        for x in range(3, 155):
            v = emul.emul_fast_x2(win.ctypes.data, 160, x, y)
            s0, s1 = (v & 0xffff) - 256, (v >> 16) - 256
            ring = [int(win[y + dy, x + dx]) for dx, dy in zip(
                [0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1],
                [-3, -3, -2, -1, 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3])]
            d = [int(win[y, x]) - r for r in ring]
            best = max(max(min(d[(s + j) % 16] for j in range(9)), -max(d[(s + j) % 16] for j in range(9)))
                       for s in range(16)) - 1
            assert s0 == best, (x, y, s0, best)
            ref[(x + 1, y)] = s1
    for (x, y), s1 in ref.items():
        if (x, y) in ref:
            pass
    # lane 1 must equal lane 0 of the next pair
    for y in range(3, 117, 7):
        for x in range(3, 154):
            a = emul.emul_fast_x2(win.ctypes.data, 160, x, y) >> 16
            b = emul.emul_fast_x2(win.ctypes.data, 160, x + 1, y) & 0xffff
            assert a == b
    assert len(cands) > 0


def test_fast_score_raw_form_equals_subtracted_form(emul):
    """k_fast_cells_v2 scores from the raw ring values (max-of-min / min-of-max, centre subtracted once at the end)."""
    emul.emul_fast_raw_mismatch.restype = C.c_long
    emul.emul_fast_raw_mismatch.argtypes = [C.c_uint32, C.c_long]
    assert emul.emul_fast_raw_mismatch(7, 2_000_000) == 0


def test_atan2_matches_oracle(emul):
    rng = np.random.default_rng(1)
    for _ in range(20000):
        y = float(rng.integers(-3_000_000, 3_000_000))
        x = float(rng.integers(-3_000_000, 3_000_000))
        assert emul.emul_atan2(y, x) == po.atan2_deg(y, x)
    for y, x in [(0, 0), (0, 1), (1, 0), (0, -1), (-1, 0), (5, 5), (-5, 5), (5, -5), (-5, -5)]:
        assert emul.emul_atan2(float(y), float(x)) == po.atan2_deg(float(y), float(x))


def test_sincos_sampled_matches_glibc(emul):
    # every 97th float of [0, 6.3]; the exhaustive sweep is the `slow` test below
    lo, hi = 0, int(np.float32(6.3).view(np.uint32))
    bad = 0
    step = 1 << 22
    for a in range(lo, hi, step):
        bad += emul.emul_sincos_mismatch(a, min(a + step // 97, hi), 4)
    assert bad == 0


def test_fast_high_speed_test_is_a_necessary_condition(emul):
    """k_fast_cells computes the full score only for pixel pairs that pass the packed antipodal-pair test: a corner at
    threshold T (score >= T) must never be rejected by it, at both thresholds and over flat / noisy / high-contrast pixels."""
    total_pass = 0
    for T in (20, 7):
        for contrast in (3, 12, 40, 255):
            passed = C.c_long(0)
            assert emul.emul_pretest_violations(1000 + T + contrast, 400000, T, contrast, C.byref(passed)) == 0, (T, contrast)
            assert 0 < passed.value < 800000           # it does reject something and does accept something
            total_pass += passed.value
    assert total_pass > 0


def test_logf_matches_glibc(emul):
    """MapPoint::PredictScale's logf: every float of [2^-6, 2^10] (the range of mfMaxDistance / dist) and a
    stride-211 sweep of all positive floats incl. subnormals, inf."""
    lo, hi = int(np.float32(2.0 ** -6).view(np.uint32)), int(np.float32(2.0 ** 10).view(np.uint32))
    assert emul.emul_logf_mismatch(lo, hi, 1, os.cpu_count() or 4) == 0
    assert emul.emul_logf_mismatch(1, 0x7f800000, 211, os.cpu_count() or 4) == 0


@pytest.mark.slow
def test_sincos_exhaustive_matches_glibc(emul):
    hi = int(np.float32(6.3).view(np.uint32))
    assert emul.emul_sincos_mismatch(0, hi, os.cpu_count() or 4) == 0


def _sort_cases():
    rng = np.random.default_rng(2)
    for n in [0, 1, 2, 3, 15, 16, 17, 18, 31, 32, 33, 64, 100, 257, 1000, 4097]:
        for kind in range(6):
            if kind == 0:
                k = rng.integers(0, 4, n)           # many ties
            elif kind == 1:
                k = rng.integers(0, 1 << 20, n)
            elif kind == 2:
                k = np.arange(n)
            elif kind == 3:
                k = np.arange(n)[::-1]
            elif kind == 4:
                k = np.concatenate([np.arange(n // 2), np.arange(n - n // 2)[::-1]])  # organ pipe
            else:
                k = (rng.integers(2, 40, n) << 16) | rng.integers(0, 6, n) * 37  # (count, ulx)-like
            yield np.ascontiguousarray(k, np.uint32)


def test_introsort_transcription_matches_libstdcxx(emul):
    for k in _sort_cases():
        n = len(k)
        v = np.arange(n, dtype=np.uint32)
        k1, v1, k2, v2 = k.copy(), v.copy(), k.copy(), v.copy()
        emul.emul_std_sort(k1.ctypes.data, v1.ctypes.data, n)
        emul.ref_std_sort(k2.ctypes.data, v2.ctypes.data, n)
        assert (k1 == k2).all() and (v1 == v2).all(), n
        k1, v1, k2, v2 = k.copy(), v.copy(), k.copy(), v.copy()
        emul.emul_heap_sort(k1.ctypes.data, v1.ctypes.data, n)
        emul.ref_heap_sort(k2.ctypes.data, v2.ctypes.data, n)
        assert (k1 == k2).all() and (v1 == v2).all(), n


def test_introsort_median_of_3_killer(emul):
    # Musser's median-of-3 killer drives introsort into its heapsort fallback
    for n in [64, 256, 1024, 2048]:
        k = np.zeros(n, np.uint32)
        h = n // 2
        for i in range(h):
            if i % 2 == 0:
                k[i] = i + 1
            else:
                k[i] = h + i + (1 if h % 2 else 0)
            k[h + i] = 2 * (i + 1)
        v = np.arange(n, dtype=np.uint32)
        k1, v1, k2, v2 = k.copy(), v.copy(), k.copy(), v.copy()
        emul.emul_std_sort(k1.ctypes.data, v1.ctypes.data, n)
        emul.ref_std_sort(k2.ctypes.data, v2.ctypes.data, n)
        assert (k1 == k2).all() and (v1 == v2).all()

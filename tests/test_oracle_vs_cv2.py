"""CPU tier: pin the oracle's OpenCV-backed stages against the real cv2 build (the only
executable ground truth in this image).  Skipped when cv2 is not importable."""
import math

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import pyoracle as po  # noqa: E402
from orb_slam3_detailed_comments_b200 import synth  # noqa: E402

cv2.setNumThreads(1)


@pytest.mark.parametrize("w,h", [(640, 480), (752, 480), (1280, 720), (321, 243)])
def test_resize_chain_matches_cv2(w, h):
    img = synth.frame(w, h, seed=w)
    ex = po.OracleExtractor(500, 1.2, 8, 20, 7)
    ex(img)
    prev = img
    for l in range(1, 8):
        lw, lh = ex.level_size(l)
        prev = cv2.resize(prev, (lw, lh), interpolation=cv2.INTER_LINEAR)
        assert (ex.level_pyramid(l) == prev).all(), (w, h, l)


def test_resize_exact_2x_uses_area_path():
    img = synth.frame(640, 480, 3)
    assert (po.resize(img, 320, 240) == cv2.resize(img, (320, 240), interpolation=cv2.INTER_LINEAR)).all()
    assert (po.resize(img, 300, 240) == cv2.resize(img, (300, 240), interpolation=cv2.INTER_LINEAR)).all()


@pytest.mark.parametrize("shape", [(97, 131), (480, 640), (134, 179), (31, 40)])
def test_gaussian_blur_matches_cv2(shape):
    img = np.random.default_rng(shape[0]).integers(0, 256, shape, dtype=np.uint8)
    ref = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    assert (po.blur(img) == ref).all()


def _cv2_cells(level, ini=20, mn=7):
    H, W = level.shape
    minB, maxBX, maxBY = 16, W - 16, H - 16
    width, height = np.float32(maxBX - minB), np.float32(maxBY - minB)
    nCols, nRows = int(width / np.float32(35)), int(height / np.float32(35))
    wCell, hCell = int(math.ceil(width / nCols)), int(math.ceil(height / nRows))
    f_ini, f_min = cv2.FastFeatureDetector_create(ini, True), cv2.FastFeatureDetector_create(mn, True)
    out, fallback = [], 0
    for i in range(nRows):
        iniY = minB + i * hCell
        maxY = min(iniY + hCell + 6, maxBY)
        if iniY >= maxBY - 3:
            continue
        for j in range(nCols):
            iniX = minB + j * wCell
            maxX = min(iniX + wCell + 6, maxBX)
            if iniX >= maxBX - 6:
                continue
            cell = np.ascontiguousarray(level[iniY:maxY, iniX:maxX])
            k = f_ini.detect(cell)
            if not k:
                k = f_min.detect(cell)
                fallback += 1 if k else 0
            out += [(int(p.pt[0]) + j * wCell, int(p.pt[1]) + i * hCell, int(p.response)) for p in k]
    return np.array(out, np.int32).reshape(-1, 3), fallback


@pytest.mark.parametrize("w,h,seed,sigma,nrect", [(640, 480, 1, 1.5, 60), (640, 480, 2, 6.0, 10), (752, 480, 3, 1.5, 60)])
def test_per_cell_fast_matches_cv2(w, h, seed, sigma, nrect):
    img = synth.frame(w, h, seed, sigma, nrect)
    ex = po.OracleExtractor(1200, 1.2, 8, 20, 7)
    ex(img)
    fallbacks = 0
    for l in range(8):
        ref, fb = _cv2_cells(ex.level_pyramid(l))
        mine = ex.level_cands(l)
        assert mine.shape == ref.shape and (mine == ref).all(), (seed, l)
        fallbacks += fb
    if sigma > 3:
        assert fallbacks > 0          # the low-texture frame must exercise the 20 -> 7 fallback


def test_fast_atan2_matches_cv2():
    rng = np.random.default_rng(0)
    for _ in range(20000):
        y, x = float(rng.integers(-3_000_000, 3_000_000)), float(rng.integers(-3_000_000, 3_000_000))
        assert po.atan2_deg(y, x) == np.float32(cv2.fastAtan2(y, x))


def test_hamming_matches_cv2_norm():
    rng = np.random.default_rng(1)
    a, b = rng.integers(0, 256, (50, 32), dtype=np.uint8), rng.integers(0, 256, (50, 32), dtype=np.uint8)
    for i in range(50):
        assert po.hamming(a[i], b[i]) == int(cv2.norm(a[i], b[i], cv2.NORM_HAMMING))

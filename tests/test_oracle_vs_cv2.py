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


def _orb(nfeatures=3000):
    return cv2.ORB_create(nfeatures=nfeatures, scaleFactor=1.2, nlevels=1, edgeThreshold=19, firstLevel=0, WTA_K=2,
                          scoreType=cv2.ORB_FAST_SCORE, patchSize=31, fastThreshold=20)


@pytest.mark.parametrize("seed,sigma,nrect", [(1, 1.5, 60), (2, 6.0, 10), (9, 1.0, 120)])
def test_ic_angle_matches_opencv_orb(seed, sigma, nrect):
    """IC_Angle (ORBextractor.cc:91-138: intensity-centroid moments over the umax disc + cv::fastAtan2) against the orientation
    OpenCV's own ORB assigns to the keypoints it detects on the same image (one level): bit-identical floats."""
    img = synth.frame(640, 480, seed, sigma, nrect)
    um = po.OracleExtractor(1000, 1.2, 8, 20, 7).umax
    kps = _orb().detect(img, None)
    assert len(kps) > 30
    for p in kps:
        x, y = int(round(p.pt[0])), int(round(p.pt[1]))
        assert np.float32(po.ic_angle(img, x, y, um)) == np.float32(p.angle), (x, y)


def test_descriptor_matches_opencv_orb_up_to_its_blur():
    """computeOrbDescriptor (ORBextractor.cc:150-203: bit_pattern_31_, steering by the keypoint angle, cvRound, bit order) against
    cv2.ORB.compute on the oracle's own level-0 keypoints and angles.  OpenCV's ORB blurs a SUB-MATRIX of its pyramid image, which takes
    GaussianBlur's float path; ORB-SLAM3 blurs a clone (ORBextractor.cc:1629-1632), which takes the bit-exact fixed-point path the
    oracle follows -- the blurred images differ by at most 1 grey level.  So: most descriptors are identical, and every differing bit
    compares two samples whose (fixed-point) blurred intensities are within 2 of each other; a wrong pattern entry, rotation, rounding
    or bit order would flip bits on well-separated samples."""
    img = synth.frame(640, 480, 1)
    ex = po.OracleExtractor(1000, 1.2, 1, 20, 7)          # one level: level-0 coordinates are image coordinates
    _, k, d = ex(img)
    kps = [cv2.KeyPoint(float(q["x"]), float(q["y"]), 31.0, float(q["angle"]), float(q["response"]), 0, -1) for q in k]
    kp2, dcv = _orb(5000).compute(img, kps)
    assert len(kp2) == len(k) and all(a.pt == b.pt for a, b in zip(kps, kp2))
    same = (dcv == d).all(1)
    assert same.mean() > 0.7
    blur = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    pat = po.pattern().reshape(-1, 2).astype(np.float32)
    f32 = np.float32
    nbits = 0
    for i in np.nonzero(~same)[0]:
        x0, y0 = int(k[i]["x"]), int(k[i]["y"])
        ang = f32(k[i]["angle"]) * f32(f32(np.pi) / f32(180.0))
        a, b = f32(np.cos(np.float64(ang))), f32(np.sin(np.float64(ang)))
        diff = np.unpackbits((dcv[i] ^ d[i])[:, None], axis=1, bitorder="little")
        for byte, bit in zip(*np.nonzero(diff)):
            v = []
            for p in (16 * byte + 2 * bit, 16 * byte + 2 * bit + 1):
                px, py = pat[p]
                xr, yr = f32(f32(px * a) - f32(py * b)), f32(f32(px * b) + f32(py * a))
                v.append(int(blur[y0 + int(np.rint(yr)), x0 + int(np.rint(xr))]))
            assert abs(v[0] - v[1]) <= 2, (i, byte, bit, v)
            nbits += 1
    assert nbits < 0.005 * d.size * 8


def test_simd_fast_route_equals_the_scalar_cell_by_cell_one():
    """The oracle's FAST runs OpenCV-style (SSE2 corner test over the level, SSE2 cornerScore) since the straw-man guard showed the scalar
    version to be ~4x slower than cv2's; the cell-by-cell scalar route (the one test_per_cell_fast_matches_cv2 pins to cv2.FAST) stays
    in the library, and both must give identical candidates, keypoints and descriptors -- including cells that fall back to minThFAST."""
    import ctypes as C
    L = po.lib()
    L.orc_set_fast_simd.restype = None
    L.orc_fast_scores.restype = None
    L.orc_fast_scores.argtypes = [C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(0)
    for t in range(4000):
        p = [rng.integers(0, 256, 49), rng.integers(100, 110, 49), np.where(rng.random(49) < 0.5, 0, 255), rng.integers(0, 3, 49) * 127][t % 4].astype(np.uint8)
        o = np.zeros(2, np.int32)
        L.orc_fast_scores(p.ctypes.data, o.ctypes.data)
        assert o[0] == o[1]
    try:
        for (w, h, seed, sig, nr, nf) in [(640, 480, 1, 1.5, 60, 1200), (640, 480, 2, 6.0, 10, 1200), (752, 480, 3, 1.5, 60, 1200), (320, 240, 4, 3.0, 20, 500)]:
            img = synth.frame(w, h, seed, sig, nr)
            out = []
            for simd in (0, 1):
                L.orc_set_fast_simd(simd)
                ex = po.OracleExtractor(nf, 1.2, 8, 20, 7)
                mono, k, d = ex(img)
                out.append((mono, k, d, [ex.level_cands(l) for l in range(8)]))
            a, b = out
            assert a[0] == b[0] and (a[1].view(np.uint8) == b[1].view(np.uint8)).all() and (a[2] == b[2]).all()
            assert all(x.shape == y.shape and (x == y).all() for x, y in zip(a[3], b[3]))
    finally:
        L.orc_set_fast_simd(1)

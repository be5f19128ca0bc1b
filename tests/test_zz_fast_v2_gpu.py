"""GPU tier: both per-cell FAST kernels against the oracle -- k_fast_cells_v2 (host-built cell table, warp-private pair lists,
list-driven NMS on a byte score map, direct warp-aggregated emission; ORB_FAST_VARIANT=1) and the round-1 k_fast_cells
(ORB_FAST_VARIANT=0).  Stage-wise: candidate lists per level (coordinates + responses), then keypoints and descriptors."""
import numpy as np
import pytest

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import ORBextractor, synth
from test_extractor_gpu import _check_stages

pytestmark = pytest.mark.gpu

CASES = [  # w, h, seed, sigma, nrect, nfeatures
    (640, 480, 1, 1.5, 60, 1200), (640, 480, 2, 6.0, 10, 1200),     # low texture: cells empty at iniThFAST run again at minThFAST
    (752, 480, 3, 1.5, 60, 1200), (320, 240, 4, 3.0, 20, 500), (1280, 720, 5, 1.5, 60, 2000), (640, 480, 6, 12.0, 2, 1200),
    (500, 377, 8, 1.0, 80, 1500),                                   # odd size: ragged last cells, pitch padding
]


@pytest.mark.parametrize("variant", ["1", "0"])
@pytest.mark.parametrize("w,h,seed,sigma,nrect,nf", CASES)
def test_fast_variant_is_bit_exact(monkeypatch, variant, w, h, seed, sigma, nrect, nf):
    monkeypatch.setenv("ORB_FAST_VARIANT", variant)       # read by orbx_create
    img = synth.frame(w, h, seed, sigma, nrect)
    ex = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
    ref = po.OracleExtractor(nf, 1.2, 8, 20, 7)
    mono, kps, desc = ex(img)
    rmono, rk, rd = ref(img)
    _check_stages(ex, ref)
    assert mono == rmono and len(kps) == len(rk)
    assert (kps.view(np.uint8) == rk.view(np.uint8)).all() and (desc == rd).all()
    ex.close()


@pytest.mark.parametrize("variant", ["1", "0"])
def test_fast_variant_extreme_images(monkeypatch, variant):
    """Plateaus (equal scores suppress each other -> a cell with corners but no keypoint falls back to minThFAST), a checkerboard
    (every pixel passes both sides of the high-speed test), pure noise (list capacity) and a flat image (nothing at all)."""
    monkeypatch.setenv("ORB_FAST_VARIANT", variant)
    rng = np.random.default_rng(3)
    w, h = 640, 480
    yy, xx = np.mgrid[0:h, 0:w]
    imgs = [((xx // 2 + yy // 2) % 2 * 200 + 20).astype(np.uint8), ((xx + yy) % 2 * 255).astype(np.uint8),
            rng.integers(0, 256, (h, w)).astype(np.uint8), np.full((h, w), 77, np.uint8),
            (((xx // 7) % 2) * 90 + ((yy // 5) % 2) * 90 + rng.integers(0, 3, (h, w))).astype(np.uint8)]
    ex = ORBextractor(1200, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
    ref = po.OracleExtractor(1200, 1.2, 8, 20, 7)
    for img in imgs:
        mono, kps, desc = ex(img)
        rmono, rk, rd = ref(img)
        _check_stages(ex, ref)
        assert mono == rmono and len(kps) == len(rk)
        assert (kps.view(np.uint8) == rk.view(np.uint8)).all() and (desc == rd).all()
    ex.close()


@pytest.mark.parametrize("variant", ["1", "0"])
def test_fast_variant_batch(monkeypatch, variant):
    monkeypatch.setenv("ORB_FAST_VARIANT", variant)
    imgs = np.stack([synth.frame(640, 480, 30 + i, 1.5 if i % 2 else 5.0, 60 if i % 2 else 8) for i in range(6)])
    ex = ORBextractor(1200, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=6)
    ex.extract_batch(imgs)
    n, mono, off, kps, desc = ex.download(6)
    ref = po.OracleExtractor(1200, 1.2, 8, 20, 7)
    for i in range(6):
        rmono, rk, rd = ref(imgs[i])
        a, b = int(off[i]), int(off[i + 1])
        assert b - a == len(rk) and (kps[a:b].view(np.uint8) == rk.view(np.uint8)).all() and (desc[a:b] == rd).all()
    ex.close()


@pytest.mark.parametrize("w,h,seed,sigma,nrect,nf", CASES[:4] + CASES[6:])
def test_fast_tma_staging_is_bit_exact(monkeypatch, w, h, seed, sigma, nrect, nf):
    """ORB_FAST_TMA=1: the window rows arrive by cp.async.bulk + mbarrier (k_fast_cells_v2<29, true>) where the cell geometry
    allows it (64-byte rows: the 35-37 px cells of these sizes; 16-byte aligned level images), the default staging elsewhere --
    same candidates, keypoints and descriptors either way."""
    monkeypatch.setenv("ORB_FAST_TMA", "1")               # read by orbx_create
    img = synth.frame(w, h, seed, sigma, nrect)
    ex = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
    ref = po.OracleExtractor(nf, 1.2, 8, 20, 7)
    mono, kps, desc = ex(img)
    rmono, rk, rd = ref(img)
    _check_stages(ex, ref)
    assert mono == rmono and len(kps) == len(rk)
    assert (kps.view(np.uint8) == rk.view(np.uint8)).all() and (desc == rd).all()
    ex.close()

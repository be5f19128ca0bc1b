"""GPU tier: the CUDA extractor (through the C ABI) against the CPU oracle, stage by stage and end to end.
Bit-exact bar: keypoint records (28 bytes each), descriptor bytes, order."""
import numpy as np
import pytest

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import ORBextractor, synth

pytestmark = pytest.mark.gpu

CASES = [
    # w, h, seed, sigma, nrect, nfeatures
    (640, 480, 1, 1.5, 60, 1000),   # config 1
    (640, 480, 2, 6.0, 10, 1200),   # low texture: 20->7 fallback cells
    (752, 480, 3, 1.5, 60, 1200),   # config 2 (two quadtree roots)
    (320, 240, 4, 3.0, 20, 500),
    (1280, 720, 5, 1.5, 60, 2000),  # config 5
]


def _check_stages(ex, ref, b=0):
    for l in range(ex.nlevels):
        assert ex.level_size(l) == ref.level_size(l)
        assert (ex.image_pyramid(b, l) == ref.level_pyramid(l)).all(), f"pyramid level {l}"
        c, rc = ex.candidates(b, l), ref.level_cands(l)
        assert c.shape == rc.shape and (c == rc).all(), f"FAST candidates level {l}: {c.shape} vs {rc.shape}"
        rb = ref.level_blurred(l)
        if rb is not None:
            assert (ex.image_pyramid(b, l, blurred=True) == rb).all(), f"blur level {l}"
        k, rk = ex.level_keypoints(b, l), ref.level_kps(l)
        rk3 = np.stack([rk["x"], rk["y"], rk["response"]], 1).astype(np.int32) if len(rk) else np.zeros((0, 3), np.int32)
        assert k.shape == rk3.shape and (k == rk3).all(), f"quadtree level {l}"


@pytest.mark.parametrize("w,h,seed,sigma,nrect,nf", CASES)
def test_single_image_bit_exact(w, h, seed, sigma, nrect, nf):
    img = synth.frame(w, h, seed, sigma, nrect)
    ex = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
    ref = po.OracleExtractor(nf, 1.2, 8, 20, 7)
    for lap in [(0, 0), (0, 1000)]:   # stereo/RGB-D call site and the monocular call site (Frame.cc:380)
        mono, kps, desc = ex(img, lap)
        rmono, rk, rd = ref(img, lap)
        _check_stages(ex, ref)
        assert mono == rmono
        assert len(kps) == len(rk)
        assert (kps.view(np.uint8) == rk.view(np.uint8)).all()
        assert (desc == rd).all()
    ex.close()


def test_getters_match_oracle():
    ex = ORBextractor(1200, 1.2, 8, 20, 7, max_width=640, max_height=480)
    ref = po.OracleExtractor(1200, 1.2, 8, 20, 7)
    assert (ex.GetScaleFactors() == ref.scale_factors).all()
    assert (ex.GetInverseScaleFactors() == ref.inv_scale_factors).all()
    assert (ex.GetScaleSigmaSquares() == ref.level_sigma2).all()
    assert (ex.GetInverseScaleSigmaSquares() == ref.inv_level_sigma2).all()
    assert (ex.mnFeaturesPerLevel == ref.features_per_level).all()
    assert (ex.umax == ref.umax).all()
    assert ex.GetLevels() == 8
    ex.close()


def test_empty_image_returns_minus_one():
    ex = ORBextractor(500, 1.2, 8, 20, 7, max_width=320, max_height=240)
    mono, kps, desc = ex(np.zeros((0, 0), np.uint8))
    assert mono == -1 and len(kps) == 0
    ex.close()


def test_flat_image_has_no_keypoints():
    ex = ORBextractor(500, 1.2, 8, 20, 7, max_width=320, max_height=240)
    mono, kps, desc = ex(np.full((240, 320), 77, np.uint8))
    assert mono == 0 and len(kps) == 0 and desc.shape == (0, 32)
    ex.close()


def test_strided_input():
    big = synth.frame(400, 240, 9)
    img = big[:, 40:360]          # non-contiguous rows: stride 400, width 320
    ex = ORBextractor(500, 1.2, 8, 20, 7, max_width=320, max_height=240)
    ref = po.OracleExtractor(500, 1.2, 8, 20, 7)
    mono, kps, desc = ex(img)
    rmono, rk, rd = ref(np.ascontiguousarray(img))
    assert mono == rmono and (kps.view(np.uint8) == rk.view(np.uint8)).all() and (desc == rd).all()
    ex.close()


def test_batch_matches_per_image_oracle():
    B, w, h = 6, 640, 480
    imgs = synth.frame_batch(B, w, h, seed=20)
    imgs[3] = synth.frame(w, h, 99, 6.0, 10)      # ragged: a low-texture frame inside the batch
    imgs[4] = 128                                 # and a flat one (zero keypoints)
    ex = ORBextractor(1200, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    n, mono = ex.extract_batch(imgs)
    n2, mono2, off, kps, desc = ex.download(B)
    assert (n == n2).all() and off[-1] == n.sum()
    ref = po.OracleExtractor(1200, 1.2, 8, 20, 7)
    for b in range(B):
        rmono, rk, rd = ref(imgs[b])
        assert n[b] == len(rk) and mono[b] == rmono, b
        assert (kps[off[b]:off[b + 1]].view(np.uint8) == rk.view(np.uint8)).all(), b
        assert (desc[off[b]:off[b + 1]] == rd).all(), b
    ex.close()


def test_handle_reuse_across_sizes():
    ex = ORBextractor(800, 1.2, 8, 20, 7, max_width=752, max_height=480)
    ref = po.OracleExtractor(800, 1.2, 8, 20, 7)
    for (w, h, seed) in [(752, 480, 1), (640, 480, 2), (376, 240, 3), (752, 480, 4)]:
        img = synth.frame(w, h, seed)
        mono, kps, desc = ex(img)
        rmono, rk, rd = ref(img)
        assert mono == rmono and (kps.view(np.uint8) == rk.view(np.uint8)).all() and (desc == rd).all(), (w, h)
    ex.close()


def test_other_pyramid_parameters():
    # scaleFactor 2.0 takes cv::resize's INTER_AREA fast path; 4 levels; different thresholds
    for (nf, sf, nl, ini, mn, w, h) in [(600, 2.0, 3, 20, 7, 640, 480), (900, 1.5, 5, 30, 10, 640, 480),
                                        (5000, 1.2, 8, 20, 7, 752, 480)]:
        img = synth.frame(w, h, 31)
        ex = ORBextractor(nf, sf, nl, ini, mn, max_width=w, max_height=h)
        ref = po.OracleExtractor(nf, sf, nl, ini, mn)
        mono, kps, desc = ex(img)
        rmono, rk, rd = ref(img)
        assert mono == rmono and len(kps) == len(rk), (nf, sf)
        assert (kps.view(np.uint8) == rk.view(np.uint8)).all() and (desc == rd).all(), (nf, sf)
        ex.close()


def test_unsupported_geometry_is_reported():
    from orb_slam3_detailed_comments_b200 import OrbError
    with pytest.raises(OrbError):
        ORBextractor(500, 1.2, 8, 20, 7, max_width=160, max_height=120)  # level 7 smaller than one FAST cell

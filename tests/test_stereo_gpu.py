"""GPU tier: Frame::ComputeStereoMatches on the device against the CPU oracle (bit-exact float32)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import ORBextractor, synth

pytestmark = pytest.mark.gpu

BF, B = 47.9, 0.11   # EuRoC-like: fx 435 * baseline 0.11


def _oracle_pair(left, right, nf=1200):
    eL, eR = po.OracleExtractor(nf, 1.2, 8, 20, 7), po.OracleExtractor(nf, 1.2, 8, 20, 7)
    _, kL, dL = eL(left)
    _, kR, dR = eR(right)
    uR, dep, kept = po.stereo_matches(eL, eR, kL, dL, kR, dR, BF, B)
    return kL, uR, dep, kept


def _same_bits(a, b):
    return a.shape == b.shape and (a.view(np.uint32) == b.view(np.uint32)).all()


def test_stereo_batch_bit_exact():
    P, w, h = 3, 640, 480
    imgs = np.zeros((2 * P, h, w), np.uint8)
    for p in range(P):
        l, r, _ = synth.stereo_pair(w, h, seed=40 + p, dmin=2.0, dmax=60.0 if p != 2 else 20.0)
        imgs[2 * p], imgs[2 * p + 1] = l, r
    ex = ORBextractor(1200, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=2 * P)
    n, mono = ex.extract_batch(imgs)
    ex.stereo_batch(P, BF, B)
    _, _, off, kps, desc = ex.download(2 * P)
    uR, dep = ex.stereo_download(int(off[-1]))
    total_matched = 0
    for p in range(P):
        kL, ruR, rdep, kept = _oracle_pair(imgs[2 * p], imgs[2 * p + 1])
        a, b = off[2 * p], off[2 * p + 1]
        assert (kps[a:b].view(np.uint8) == kL.view(np.uint8)).all()
        assert _same_bits(uR[a:b], ruR), p
        assert _same_bits(dep[a:b], rdep), p
        assert kept == int((ruR >= 0).sum())
        total_matched += kept
    assert total_matched > 300      # the synthetic disparity field is matchable
    ex.close()


def test_stereo_pair_two_extractor_objects():
    # the reference's own arrangement: mpORBextractorLeft / mpORBextractorRight (Frame.cc:136-141)
    w, h = 640, 480
    l, r, _ = synth.stereo_pair(w, h, seed=77)
    exL = ORBextractor(1200, 1.2, 8, 20, 7, max_width=w, max_height=h)
    exR = ORBextractor(1200, 1.2, 8, 20, 7, max_width=w, max_height=h)
    _, kL, dL = exL(l)
    _, kR, dR = exR(r)
    uR, dep = exL.stereo_pair(exR, len(kL), BF, B)
    _, ruR, rdep, kept = _oracle_pair(l, r)
    assert _same_bits(uR, ruR) and _same_bits(dep, rdep) and kept > 100
    exL.close()
    exR.close()


def test_stereo_no_texture_and_identical_images():
    w, h = 320, 240
    ex = ORBextractor(500, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=4)
    img = synth.frame(w, h, 5)
    imgs = np.stack([np.full((h, w), 90, np.uint8), np.full((h, w), 90, np.uint8), img, img])
    ex.extract_batch(imgs)
    ex.stereo_batch(2, BF, B)
    _, _, off, kps, _ = ex.download(4)
    uR, dep = ex.stereo_download(int(off[-1]))
    assert off[1] == 0 and off[2] == 0                      # flat pair: no keypoints at all
    kL, ruR, rdep, kept = _oracle_pair(img, img, 500)       # zero disparity everywhere
    assert _same_bits(uR[off[2]:off[3]], ruR) and _same_bits(dep[off[2]:off[3]], rdep)
    ex.close()

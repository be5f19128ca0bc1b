"""GPU tier: k_stereo_match_v1 (one thread per left keypoint, right keypoints bucketed by row once per CTA;
csrc/stereo_core.cuh, CPU-validated by tests/test_stereo_emul.py) against the oracle, bit-exact float32.  ORB_STEREO_VARIANT is read at
every stereo call, so the switch works inside a process that already ran variant 0."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import ORBextractor, synth

pytestmark = pytest.mark.gpu
BF, B = 47.9, 0.11


@pytest.mark.parametrize("variant", ["1", "0"])
@pytest.mark.parametrize("w,h,nf", [(640, 480, 1200), (752, 480, 1200), (1280, 720, 2000), (320, 240, 500)])
def test_variant_batch_bit_exact(monkeypatch, variant, w, h, nf):
    monkeypatch.setenv("ORB_STEREO_VARIANT", variant)     # 1 (k_stereo_match_v1) is the default since round 2
    P = 3
    imgs = np.zeros((2 * P, h, w), np.uint8)
    for p in range(P):
        imgs[2 * p], imgs[2 * p + 1], _ = synth.stereo_pair(w, h, seed=60 + p, dmin=2.0, dmax=60.0 if p != 2 else 20.0)
    imgs[5] = 128                                         # a flat right image: no right keypoints for the last pair
    ex = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=2 * P)
    ex.extract_batch(imgs)
    ex.stereo_batch(P, BF, B)
    _, _, off, kps, desc = ex.download(2 * P)
    uR, dep = ex.stereo_download(int(off[-1]))
    matched = 0
    for p in range(P):
        eL, eR = po.OracleExtractor(nf, 1.2, 8, 20, 7), po.OracleExtractor(nf, 1.2, 8, 20, 7)
        _, kL, dL = eL(imgs[2 * p])
        _, kR, dR = eR(imgs[2 * p + 1])
        ruR, rdep, _ = po.stereo_matches(eL, eR, kL, dL, kR, dR, BF, B)
        a, b = off[2 * p], off[2 * p + 1]
        assert (uR[a:b].view(np.uint32) == ruR.view(np.uint32)).all() and (dep[a:b].view(np.uint32) == rdep.view(np.uint32)).all(), p
        matched += int((rdep > 0).sum())
    assert matched > 200
    ex.close()

"""GPU tier: both DistributeOctTree kernels against the oracle: k_quadtree_v1 (default since its first device run in round 2: multi-stage
bitonic passes, the ordered phase's std::sort spread over the CTA; CPU-validated by tests/test_quadtree_emul.py) and the round-1
k_quadtree (ORB_QT_VARIANT=0), which keeps its machine code bit for bit."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import ORBextractor, synth

pytestmark = pytest.mark.gpu

CASES = [(640, 480, 1, 1.5, 60, 1200), (640, 480, 2, 6.0, 10, 1200), (752, 480, 3, 1.5, 60, 1200), (320, 240, 4, 3.0, 20, 500),
         (1280, 720, 5, 1.5, 60, 2000), (640, 480, 7, 1.5, 60, 5000)]


@pytest.mark.parametrize("variant", ["1", "0"])
@pytest.mark.parametrize("w,h,seed,sigma,nrect,nf", CASES)
def test_variant_is_bit_exact(monkeypatch, variant, w, h, seed, sigma, nrect, nf):
    monkeypatch.setenv("ORB_QT_VARIANT", variant)        # read by orbx_create; 1 is the default since round 2
    img = synth.frame(w, h, seed, sigma, nrect)
    ex = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
    ref = po.OracleExtractor(nf, 1.2, 8, 20, 7)
    mono, kps, desc = ex(img)
    rmono, rk, rd = ref(img)
    for l in range(ex.nlevels):                           # name the first level whose distribution differs, before the final compare
        k, rl = ex.level_keypoints(0, l), ref.level_kps(l)
        rl3 = np.stack([rl["x"], rl["y"], rl["response"]], 1).astype(np.int32) if len(rl) else np.zeros((0, 3), np.int32)
        assert k.shape == rl3.shape and (k == rl3).all(), f"quadtree level {l}: {k.shape} vs {rl3.shape}"
    assert mono == rmono and len(kps) == len(rk)
    assert (kps.view(np.uint8) == rk.view(np.uint8)).all() and (desc == rd).all()
    ex.close()


@pytest.mark.parametrize("variant", ["1", "0"])
def test_variant_batch(monkeypatch, variant):
    monkeypatch.setenv("ORB_QT_VARIANT", variant)
    imgs = np.stack([synth.frame(640, 480, 20 + i) for i in range(6)])
    ex = ORBextractor(1200, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=6)
    ex.extract_batch(imgs)
    n, mono, off, kps, desc = ex.download(6)
    ref = po.OracleExtractor(1200, 1.2, 8, 20, 7)
    for i in range(6):
        rmono, rk, rd = ref(imgs[i])
        a, b = int(off[i]), int(off[i + 1])
        assert b - a == len(rk) and (kps[a:b].view(np.uint8) == rk.view(np.uint8)).all() and (desc[a:b] == rd).all()
    ex.close()


def test_many_textures_final_keypoints(monkeypatch):
    """The ordered phase's std::sort emulation (warp-cooperative partitions of the long segments, one lane per short one) decides the
    order in which tied nodes are divided: 24 textures x 8 levels of different candidate densities against the oracle, final output."""
    monkeypatch.setenv("ORB_QT_VARIANT", "1")
    rng = np.random.default_rng(11)
    imgs = np.stack([synth.frame(640, 480, 200 + i, float(rng.choice([1.0, 1.5, 2.5, 3.5, 5.0, 8.0])), int(rng.choice([5, 20, 60, 120]))) for i in range(24)])
    ex = ORBextractor(1200, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=24)
    ex.extract_batch(imgs)
    n, mono, off, kps, desc = ex.download(24)
    ref = po.OracleExtractor(1200, 1.2, 8, 20, 7)
    for i in range(24):
        rmono, rk, rd = ref(imgs[i])
        a, b = int(off[i]), int(off[i + 1])
        assert b - a == len(rk) and (kps[a:b].view(np.uint8) == rk.view(np.uint8)).all() and (desc[a:b] == rd).all(), i
    ex.close()

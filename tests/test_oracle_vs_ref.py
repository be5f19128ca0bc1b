"""CPU tier: oracle/extractor_oracle.cpp (the restatement every GPU parity test compares against) against oracle/_ref --
the reference's OWN src/ORBextractor.cc compiled unmodified over the cv2-pinned pixel models (oracle/Makefile `ref`,
oracle/ref_shim/).  This pins to the reference source: the constructor tables (ORBextractor.cc:468-571), ComputePyramid's
chaining (:1687-1738), the per-cell FAST loop with its borders and 20->7 fallback (:1061-1166), DistributeOctTree's node
order, tie-breaking and early exits (:711-1057), IC_Angle (:91-138), computeOrbDescriptor (:150-203) and operator()'s output
ordering / vLappingArea fill / rescale (:1557-1682)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import synth

pytestmark = pytest.mark.skipif(po.build_ref() is None, reason="oracle/_ref not built and /root/reference absent")

CASES = [  # tests/test_extractor_gpu.py::CASES (w, h, seed, sigma, nrect, nfeatures) ...
    (640, 480, 1, 1.5, 60, 1000), (640, 480, 2, 6.0, 10, 1200), (752, 480, 3, 1.5, 60, 1200), (320, 240, 4, 3.0, 20, 500),
    (1280, 720, 5, 1.5, 60, 2000),
    # ... plus shapes that stress the quadtree exits: very few candidates, a quota larger than the candidates, 3 roots
    (640, 480, 6, 12.0, 2, 1200), (400, 300, 7, 1.5, 60, 5000), (960, 240, 8, 2.0, 40, 800),
]


def test_constructor_tables():
    for nf, sf, nl in [(1000, 1.2, 8), (1200, 1.2, 8), (2000, 1.2, 8), (500, 1.5, 5), (3000, 2.0, 4), (1500, 1.1, 12)]:
        a, b = po.OracleExtractor(nf, sf, nl, 20, 7), po.RefExtractor(nf, sf, nl, 20, 7)
        for f in ("scale_factors", "inv_scale_factors", "level_sigma2", "inv_level_sigma2"):
            assert (getattr(a, f).view(np.uint32) == getattr(b, f).view(np.uint32)).all(), f
        assert (a.features_per_level == b.features_per_level).all() and (a.umax == b.umax).all()


@pytest.mark.parametrize("w,h,seed,sigma,nrect,nf", CASES)
def test_operator_call_bit_exact(w, h, seed, sigma, nrect, nf):
    img = synth.frame(w, h, seed, sigma, nrect)
    orc, ref = po.OracleExtractor(nf, 1.2, 8, 20, 7), po.RefExtractor(nf, 1.2, 8, 20, 7)
    for lap in [(0, 0), (0, 1000), (100, 300)]:
        mono, kps, desc = orc(img, lap)
        rmono, rk, rd = ref(img, lap)
        assert mono == rmono and len(kps) == len(rk) and len(kps) > 0
        assert (kps.view(np.uint8) == rk.view(np.uint8)).all()
        assert (desc == rd).all()
    for l in range(8):
        assert orc.level_size(l) == ref.level_size(l)
        assert (orc.level_pyramid(l) == ref.level_pyramid(l)).all()
    # the stage before descriptors: per-level keypoints in level coordinates, order and angles
    per_level = ref.keypoints_octtree(img)
    orc(img, (0, 0))
    for l in range(8):
        assert (orc.level_kps(l).view(np.uint8) == per_level[l].view(np.uint8)).all(), l


def test_other_pyramid_shapes_bit_exact():
    for (nf, sf, nl, w, h, seed) in [(800, 1.5, 5, 640, 480, 11), (600, 2.0, 4, 512, 384, 12), (1500, 1.1, 10, 480, 360, 13)]:
        img = synth.frame(w, h, seed)
        orc, ref = po.OracleExtractor(nf, sf, nl, 20, 7), po.RefExtractor(nf, sf, nl, 20, 7)
        (mono, kps, desc), (rmono, rk, rd) = orc(img), ref(img)
        assert mono == rmono and (kps.view(np.uint8) == rk.view(np.uint8)).all() and (desc == rd).all()


def _random_case(rng):
    """A random (candidate set, rectangle, N): the shapes DistributeOctTree meets (one to three roots; clustered, duplicated and
    equal-response candidates; N from 1 to beyond the candidate count)."""
    W = int(rng.integers(40, 1300))
    H = int(rng.integers(40, 760))
    if rng.random() < 0.3:
        W = max(W, 2 * H)           # several root nodes
    n = int(rng.choice([1, 2, 3, 5, 17, 60, 200, 700, 2500]))
    kind = rng.integers(0, 4)
    if kind == 0:
        x, y = rng.integers(0, W, n), rng.integers(0, H, n)
    elif kind == 1:                 # clusters: deep subdivision, single-point leaves
        cx, cy = rng.integers(0, W, 4), rng.integers(0, H, 4)
        k = rng.integers(0, 4, n)
        x = np.clip(cx[k] + rng.normal(0, 6, n).astype(int), 0, W - 1)
        y = np.clip(cy[k] + rng.normal(0, 6, n).astype(int), 0, H - 1)
    elif kind == 2:                 # coincident points (a leaf that can never split to size 1)
        x, y = rng.integers(0, W, n) // 8 * 8, rng.integers(0, H, n) // 8 * 8
    else:                           # points on the split lines
        x, y = rng.integers(0, 8, n) * (W // 8), rng.integers(0, 8, n) * (H // 8)
    resp = rng.integers(7, 12 if rng.random() < 0.5 else 255, n)
    cand = np.stack([x, y, resp], 1).astype(np.int32)
    N = int(rng.choice([1, 2, 5, 30, 72, 151, 261, 434, n, 4 * n + 1]))
    return cand, W, H, N


def test_distribute_random_cases():
    rng = np.random.default_rng(2024)
    ref = po.RefExtractor(1000, 1.2, 8, 20, 7)
    done = 0
    while done < 10000:
        cand, W, H, N = _random_case(rng)
        if round(W / H) < 1:        # nIni == 0: the reference divides by zero (hX = inf); not a geometry it can process
            continue
        r = ref.distribute(cand, 0, W, 0, H, N)
        o = po.oracle_distribute(cand, 0, W, 0, H, N)
        assert len(r) == len(o), (done, len(cand), W, H, N)
        # the restatement returns input indices; the reference returns copies carrying class_id = input index
        assert (r[:, 3] == o).all(), (done, len(cand), W, H, N)
        assert (cand[o] == r[:, :3]).all()
        done += 1

"""GPU tier: batched MapPoint::ComputeDistinctiveDescriptors / UpdateNormalAndDepth against the CPU oracle (bit-exact)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import ORBextractor, ComputeDistinctiveDescriptors, UpdateNormalAndDepth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ex():
    e = ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=1)
    yield e
    e.close()


def test_distinctive_descriptors(ex):
    rng = np.random.default_rng(0)
    counts = [0, 1, 2, 3, 4, 5, 8, 13, 31, 32, 33, 64, 100, 7, 2, 0, 9] + list(rng.integers(1, 40, 300))
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    descs = []
    for c in counts:                       # a cluster: one base descriptor, each observation with a few to many flipped bits
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        for _ in range(c):
            d = base.copy()
            for b in rng.integers(0, 256, rng.integers(0, 60)):
                d[b // 8] ^= np.uint8(1 << (b % 8))
            descs.append(d)
    descs = np.stack(descs)
    descs[off[10]:off[10] + 5] = descs[off[10]]      # exact duplicates: ties in the medians, first minimum wins
    got = ComputeDistinctiveDescriptors(ex, off, descs)
    for i, c in enumerate(counts):
        assert got[i] == po.distinctive_descriptor(descs[off[i]:off[i + 1]]), (i, c)


def test_update_normal_and_depth(ex):
    rng = np.random.default_rng(1)
    sf = ex.GetScaleFactors()
    counts = [0, 1, 2, 5, 17] + list(rng.integers(1, 30, 200))
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    n = len(counts)
    pos = rng.normal(0, 5, (n, 3)).astype(np.float32)
    centers = rng.normal(0, 3, (int(off[-1]), 3)).astype(np.float32)
    refc = rng.normal(0, 3, (n, 3)).astype(np.float32)
    lvl = rng.integers(0, 8, n).astype(np.int32)
    n0, mx0, mn0 = rng.normal(size=(n, 3)).astype(np.float32), rng.uniform(1, 2, n).astype(np.float32), rng.uniform(0, 1, n).astype(np.float32)
    nrm, mx, mn = UpdateNormalAndDepth(ex, off, centers, pos, refc, lvl, n0, mx0, mn0)
    for i, c in enumerate(counts):
        if c == 0:
            assert (nrm[i] == n0[i]).all() and mx[i] == mx0[i] and mn[i] == mn0[i]
            continue
        rn, rmx, rmn = po.update_normal_and_depth(centers[off[i]:off[i + 1]], pos[i], refc[i], lvl[i], sf)
        assert (nrm[i].view(np.uint32) == rn.view(np.uint32)).all(), i
        assert np.float32(mx[i]).view(np.uint32) == np.float32(rmx).view(np.uint32) and np.float32(mn[i]).view(np.uint32) == np.float32(rmn).view(np.uint32)

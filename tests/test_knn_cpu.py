"""CPU tier: the brute-force Hamming 2-NN oracle (orc_hamming_knn2) pinned against OpenCV itself --
cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(query, train, k=2) is the very call Frame::ComputeStereoFishEyeMatches makes (Frame.cc:1553) --
including the order of tied distances and train sets with fewer than two rows."""
import numpy as np
import pytest

from oracle import pyoracle as po

cv2 = pytest.importorskip("cv2")


def cv_knn2(q, t):
    idx, dist = np.full((len(q), 2), -1, np.int32), np.full((len(q), 2), -1, np.int32)
    if len(q) == 0 or len(t) == 0:
        return idx, dist
    for i, mm in enumerate(cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, 2)):
        for k, m in enumerate(mm):
            assert m.queryIdx == i
            idx[i, k], dist[i, k] = m.trainIdx, int(m.distance)
    return idx, dist


@pytest.mark.parametrize("nq,nt,bits", [(300, 500, 2), (257, 129, 8), (40, 1, 8), (5, 2, 1), (1200, 1200, 8), (3, 0, 8), (0, 7, 8)])
def test_oracle_matches_opencv(nq, nt, bits):
    rng = np.random.default_rng(nq * 1000 + nt)
    q = rng.integers(0, 1 << bits, (nq, 32)).astype(np.uint8)      # few bits per byte -> many tied distances
    t = rng.integers(0, 1 << bits, (nt, 32)).astype(np.uint8)
    if nt > 10 and nq > 10:
        t[[3, 7, nt - 1]] = q[0]                                   # exact duplicates at several train indices
        t[5] = q[1]
        t[5, 0] ^= 1
        t[9] = q[1]
        t[9, 31] ^= 0x80                                           # two rows at distance 1
    i0, d0 = po.hamming_knn2(q, t)
    i1, d1 = cv_knn2(q, t)
    assert (i0 == i1).all() and (d0 == d1).all()
    if nt > 10 and nq > 10:
        assert i0[0].tolist() == [3, 7] and d0[0].tolist() == [0, 0] and i0[1].tolist() == [5, 9]


def test_ratio_test_on_extracted_descriptors():
    """The consumer's rule (Frame.cc:1560): m[0].distance < m[1].distance * 0.7 on real ORB descriptors of a stereo pair."""
    from orb_slam3_detailed_comments_b200 import synth
    L, R = synth.stereo_pair(320, 240, 3)[:2]
    ex = po.OracleExtractor(500, 1.2, 8, 20, 7)
    _, _, dl = ex(L)
    _, _, dr = ex(R)
    i0, d0 = po.hamming_knn2(dl, dr)
    i1, d1 = cv_knn2(dl, dr)
    assert (i0 == i1).all() and (d0 == d1).all()
    good = d0[:, 0] < d0[:, 1] * np.float32(0.7)
    assert good.sum() > 50


def test_fuzz_against_opencv():
    """Random shapes and bit densities (hypothesis): the oracle equals cv2.BFMatcher.knnMatch on every draw."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(0, 70), st.integers(0, 70), st.integers(1, 8), st.integers(0, 2 ** 31 - 1))
    def run(nq, nt, bits, seed):
        rng = np.random.default_rng(seed)
        q = rng.integers(0, 1 << bits, (nq, 32)).astype(np.uint8)
        t = rng.integers(0, 1 << bits, (nt, 32)).astype(np.uint8)
        if nq and nt > 3:
            t[rng.integers(0, nt, 3)] = q[rng.integers(0, nq)]
        i0, d0 = po.hamming_knn2(q, t)
        i1, d1 = cv_knn2(q, t)
        assert (i0 == i1).all() and (d0 == d1).all()
    run()

"""CPU tier, build container only: the oracle's KeyFrame-typed matchers against the REFERENCE's own member functions.
oracle/_ref/liborb_ref2.so holds ORBmatcher::SearchByBoW (both overloads), SearchForInitialization and SearchForTriangulation cut out
of /root/reference/src/ORBmatcher.cc at build time and compiled verbatim over skeleton Frame / KeyFrame / MapPoint classes, with the
reference's own DBoW2::FeatureVector and Pinhole::epipolarConstrain.  Inputs: oracle-extracted features of two synthetic frames; the
vocabulary node of a feature is a hash of its descriptor (only the partition into nodes matters to these functions)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import synth

pytestmark = pytest.mark.skipif(po.build_ref2() is None, reason="oracle/_ref part 2 not built and /root/reference absent")

W, H = 640, 480
FX, FY, CX, CY, BF, B = 435.2, 435.2, 320.0, 240.0, 47.9, 0.11
CAM6 = np.array([FX, FY, CX, CY, BF, B], np.float32)
BOUNDS = np.array([0, W, 0, H], np.float32)


def node_of(d, m):
    return (((d[:, 0].astype(np.int32) >> 3) * 7 + (d[:, 5].astype(np.int32) >> 4) * 3 + (d[:, 17].astype(np.int32) >> 5)) % m).astype(np.int32)


@pytest.fixture(scope="module")
def two_frames():
    rng = np.random.default_rng(9)
    l, r, _ = synth.stereo_pair(W, H, seed=510)
    out = []
    for k in range(2):
        eL, eR = po.OracleExtractor(1200, 1.2, 8, 20, 7), po.OracleExtractor(1200, 1.2, 8, 20, 7)
        if k == 1:
            noise = rng.integers(-3, 4, (H, W))
            l = np.clip(np.roll(l, (4, 1), (1, 0)).astype(int) + noise, 0, 255).astype(np.uint8)
            r = np.clip(np.roll(r, (4, 1), (1, 0)).astype(int) + noise, 0, 255).astype(np.uint8)
        _, kL, dL = eL(l)
        _, kR, dR = eR(r)
        uR, dep, _ = po.stereo_matches(eL, eR, kL, dL, kR, dR, BF, B)
        out.append((kL, dL, uR, dep))
    sf = np.array([1.2 ** i for i in range(8)], np.float32)
    return out[0], out[1], sf


def _merge_order(node, keep):
    """Indices of the features with keep[i] and node[i] >= 0 in FeatureVector merge order (node, then feature index)."""
    sel = np.nonzero(keep & (node >= 0))[0]
    return sel[np.lexsort((sel, node[sel]))]


@pytest.mark.parametrize("nnratio,check", [(0.7, True), (0.9, False), (0.75, True)])
def test_search_by_bow_keyframe_to_frame(two_frames, nnratio, check):
    (k1, d1, u1, dep1), (k2, d2, u2, dep2), sf = two_frames
    rng = np.random.default_rng(int(nnratio * 100))
    nodeF = node_of(d2, 97); nodeF[::17] = -1                 # F = the second frame; a few features without a BoW entry
    nodeK = node_of(d1, 97); nodeK[5::23] = -1
    has_mp = dep1 > 0
    bad = has_mp & (rng.random(len(k1)) < 0.05)               # MapPoint::isBad(): skipped
    F = po.RefFrame(k2, d2, u2, BOUNDS, sf, CAM6)
    K = po.RefKeyFrame(k1, d1, u1, nodeK, has_mp, bad, sf, sf * sf, CAM6[:4])
    rfm, rn = po.ref2_search_bow(F, nodeF, K, nnratio, check)
    q = _merge_order(nodeK, has_mp & ~bad)
    ofm, on = po.search_bow(k2, d2, nodeF, nodeK[q], k1["angle"][q], d1[q], nnratio, check)
    got = np.where(ofm >= 0, q[np.maximum(ofm, 0)], -1)      # the oracle names the query; the reference the KF feature
    assert on == rn and (got == rfm).all()
    assert rn > 40


@pytest.mark.parametrize("nnratio,check", [(0.8, True), (0.95, False)])
def test_search_by_bow_keyframe_to_keyframe(two_frames, nnratio, check):
    (k1, d1, u1, dep1), (k2, d2, u2, dep2), sf = two_frames
    rng = np.random.default_rng(3)
    node1, node2 = node_of(d1, 61), node_of(d2, 61)
    node1[3::19] = -1; node2[::13] = -1
    mp1, mp2 = dep1 > 0, rng.random(len(k2)) < 0.7
    bad1, bad2 = mp1 & (rng.random(len(k1)) < 0.04), mp2 & (rng.random(len(k2)) < 0.04)
    K1 = po.RefKeyFrame(k1, d1, u1, node1, mp1, bad1, sf, sf * sf, CAM6[:4])
    K2 = po.RefKeyFrame(k2, d2, u2, node2, mp2, bad2, sf, sf * sf, CAM6[:4])
    rm, rn = po.ref2_search_bow_kf(K1, K2, nnratio, check)
    q = _merge_order(node1, mp1 & ~bad1)
    om, on = po.search_bow_kf(k2, d2, node2, (mp2 & ~bad2).astype(np.uint8), node1[q], k1["angle"][q], d1[q], nnratio, check)
    got = np.full(len(k1), -1, np.int32)
    got[q] = om
    assert on == rn and (got == rm).all()
    assert rn > 30


@pytest.mark.parametrize("window,nnratio,check", [(100, 0.9, True), (30, 0.8, False)])
def test_search_for_initialization(two_frames, window, nnratio, check):
    (k1, d1, u1, _), (k2, d2, u2, _), sf = two_frames
    F1 = po.RefFrame(k1, d1, None, BOUNDS, sf, CAM6)
    F2 = po.RefFrame(k2, d2, None, BOUNDS, sf, CAM6)
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)          # Tracking.cc:2452: mvbPrevMatched[i] = mInitialFrame.mvKeysUn[i].pt
    rm, rn, rprev = po.ref2_search_initialization(F1, F2, prev, window, nnratio, check)
    om, on = po.search_initialization(k1, d1, prev, k2, d2, BOUNDS, window, nnratio, check)
    assert on == rn and (om == rm).all()
    want = prev.copy()                                                  # ORBmatcher.cc:884-887: matched entries move to the F2 keypoint
    want[rm >= 0] = np.stack([k2["x"], k2["y"]], 1)[rm[rm >= 0]]
    assert (rprev == want).all()
    assert rn > 50 and (k1["octave"][rm >= 0] == 0).all()


def _quat_pose(yaw_deg, t):
    a = np.deg2rad(yaw_deg) / 2
    return np.array([0, np.sin(a), 0, np.cos(a), *t], np.float32)


@pytest.mark.parametrize("only_stereo,coarse,check,mono", [(False, False, True, False), (False, True, True, False), (True, False, True, False),
                                                           (False, False, False, True)])
def test_search_for_triangulation(two_frames, only_stereo, coarse, check, mono):
    (k1, d1, u1, dep1), (k2, d2, u2, dep2), sf = two_frames
    rng = np.random.default_rng(5)
    node1, node2 = node_of(d1, 61), node_of(d2, 61)
    node2[::13] = -1
    mp1, mp2 = rng.random(len(k1)) < 0.4, rng.random(len(k2)) < 0.3      # features that already hold a map point are skipped on both sides
    ur1, ur2 = (np.full(len(k1), -1, np.float32), np.full(len(k2), -1, np.float32)) if mono else (u1, u2)
    T1, T2 = _quat_pose(0.0, [0, 0, 0]), _quat_pose(0.4, [-0.03, -0.002, -0.01])
    K1 = po.RefKeyFrame(k1, d1, ur1, node1, mp1, None, sf, sf * sf, CAM6[:4], T1)
    K2 = po.RefKeyFrame(k2, d2, ur2, node2, mp2, None, sf, sf * sf, CAM6[:4], T2)
    rm, rn, F12, ep = po.ref2_search_triangulation(K1, K2, only_stereo, coarse, check)
    st1, st2 = (ur1 >= 0), (ur2 >= 0)
    q = _merge_order(node1, ~mp1 & (st1 if only_stereo else True))
    valid2 = ~mp2 & (st2 if only_stereo else True)
    om, on = po.search_triangulation(k1[q], d1[q], node1[q], st1[q].astype(np.uint8), k2, d2, node2, valid2.astype(np.uint8), st2.astype(np.uint8),
                                     F12, ep, sf, sf * sf, coarse, check)
    got = np.full(len(k1), -1, np.int32)
    got[q] = om
    assert on == rn and (got == rm).all()
    assert rn > (5 if not coarse else 40)


def test_compute_distinctive_descriptors():
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:438-529): least median Hamming distance to the other observations; ties keep
    the first; bad keyframes are left out."""
    rng = np.random.default_rng(8)
    base = rng.integers(0, 256, 32, dtype=np.uint8)
    for n in (1, 2, 3, 4, 7, 12, 30):
        for rep in range(6):
            d = np.repeat(base[None], n, 0).copy()
            for r in range(n):
                for b in rng.integers(0, 256, int(rng.integers(0, 60))):
                    d[r, b // 8] ^= np.uint8(1 << (b % 8))
            if rep == 5 and n > 2:
                d[1] = d[0]                                       # exact ties
            bad = (rng.random(n) < 0.25).astype(np.uint8) if rep % 2 else np.zeros(n, np.uint8)
            good = d[bad == 0]
            want = po.distinctive_descriptor(good)
            got = po.ref2_distinctive_descriptor(d, bad)
            if len(good) == 0:
                assert got is None and want == -1
            else:
                assert (got == good[want]).all(), (n, rep)


def test_update_normal_and_depth():
    """MapPoint::UpdateNormalAndDepth (MapPoint.cc:567-643): float bits of the mean viewing direction and of the two distance bounds."""
    rng = np.random.default_rng(4)
    sf = np.array([1.2 ** i for i in range(8)], np.float32)
    for n in (1, 2, 5, 17):
        for _ in range(8):
            centers = rng.normal(0, 2.0, (n, 3)).astype(np.float32)
            pos = (rng.normal(0, 1.0, 3) + np.array([0, 0, 8.0])).astype(np.float32)
            ref, level = int(rng.integers(0, n)), int(rng.integers(0, 8))
            a = po.update_normal_and_depth(centers, pos, centers[ref], level, sf)
            b = po.ref2_update_normal_and_depth(centers, pos, ref, level, sf)
            assert (a[0].view(np.uint32) == b[0].view(np.uint32)).all()
            assert np.float32(a[1]).view(np.uint32) == np.float32(b[1]).view(np.uint32) and np.float32(a[2]).view(np.uint32) == np.float32(b[2]).view(np.uint32)

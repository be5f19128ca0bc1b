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

import os

if os.environ.get("ORB_PIN_GEOMETRY") == "720p":      # tests/test_oracle_vs_ref_matcher2_720p.py re-runs this file at BASELINE config 5's geometry
    W, H, NFEAT = 1280, 720, 2000
    FX, FY, CX, CY, BF, B = 870.4, 870.4, 640.0, 360.0, 95.7, 0.11
else:
    W, H, NFEAT = 640, 480, 1200
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
        eL, eR = po.OracleExtractor(NFEAT, 1.2, 8, 20, 7), po.OracleExtractor(NFEAT, 1.2, 8, 20, 7)
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


# ---- projections into a keyframe: Fuse x2, SearchByProjection(KF, Scw) x2, SearchByProjection(F, KF, set), SearchBySim3 ----------------
def _unproject(k, depth):
    return np.stack([(k["x"] - CX) * depth / FX, (k["y"] - CY) * depth / FY, depth], 1).astype(np.float32)


def _camera_center(T):
    q, t = T[:4].astype(np.float64), T[4:].astype(np.float64)
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return (-R.T @ t).astype(np.float32), R, t


def _queries(k1, d1, dep1, rng, sf, dup):
    """Map points seen by frame 1 (= world), with competing duplicates, noisy normals and a few points seen from behind."""
    sel = np.nonzero(dep1 > 0)[0]
    sel = np.concatenate([sel, rng.choice(sel, int(dup * len(sel)))])
    pts = _unproject(k1[sel], dep1[sel]) + rng.normal(0, 0.002, (len(sel), 3)).astype(np.float32)
    dist = np.linalg.norm(pts, axis=1).astype(np.float32)
    nrm = pts / dist[:, None] + rng.normal(0, 0.2, pts.shape).astype(np.float32)
    nrm = (nrm / np.linalg.norm(nrm, axis=1)[:, None]).astype(np.float32)
    nrm[rng.random(len(sel)) < 0.05] *= -1
    maxd = (dist * sf[k1["octave"][sel]]).astype(np.float32)
    return dict(world_pos=pts, normal=nrm, max_dist=maxd, min_dist=(maxd / sf[7]).astype(np.float32), desc=d1[sel],
                angle=k1["angle"][sel].astype(np.float32), bad=(rng.random(len(sel)) < 0.04).astype(np.uint8))


def _keep(q, mask):
    return {k: (v[mask] if v is not None else None) for k, v in q.items()}


@pytest.fixture(scope="module")
def kf_target(two_frames):
    (k1, d1, u1, dep1), (k2, d2, u2, dep2), sf = two_frames
    zmid = float(np.median(dep1[dep1 > 0]))
    T = _quat_pose(0.05, [4 * zmid / FX, 1 * zmid / FY, 0.0])          # frame 2 = frame 1 shifted by (4, 1) px at the median depth
    isg = (1.0 / (sf * sf)).astype(np.float32)
    F = po.RefFrame(k2, d2, u2, BOUNDS, sf, CAM6, T)
    return dict(T=T, isg=isg, F=F, logsf=po.logf(1.2))


@pytest.mark.parametrize("th", [3.0, 8.0])
def test_fuse_pose(two_frames, kf_target, th):
    (k1, d1, u1, dep1), (k2, d2, u2, dep2), sf = two_frames
    rng = np.random.default_rng(int(th))
    T, isg, F = kf_target["T"], kf_target["isg"], kf_target["F"]
    has_mp = rng.random(len(k2)) < 0.3                                   # features that already hold a map point: the Replace branch
    K = po.RefKeyFrame(k2, d2, u2, None, has_mp, None, sf, sf * sf, CAM6[:4], T)
    po.ref2_kf_set_geometry(K, F, isg, BF)
    q = _queries(k1, d1, dep1, rng, sf, 0.3)
    in_kf = (rng.random(len(q["desc"])) < 0.05).astype(np.uint8)
    rf, rn = po.ref2_fuse(K, q, th, in_kf)
    ok = (q["bad"] == 0) & (in_kf == 0)
    Ow, _, _ = _camera_center(T)
    qq = _keep(q, ok)
    om, on, _ = po.search_keyframe(0, k2, d2, u2, BOUNDS, sf, isg, kf_target["logsf"], CAM6, T, Ow, qq["world_pos"], qq["normal"], qq["max_dist"],
                                   qq["min_dist"], qq["desc"], qq["angle"], None, th, 50.0)
    got = np.full(len(ok), -1, np.int32)
    got[ok] = om
    assert on == rn and (got == rf).all()
    assert rn > 100


def _sim3(q_unit, t, s):
    q = (np.asarray(q_unit, np.float64) * np.sqrt(s)).astype(np.float32)
    sc = np.float32(np.float32(q[0] * q[0] + q[1] * q[1]) + np.float32(q[2] * q[2] + q[3] * q[3]))
    return np.concatenate([q, np.asarray(t, np.float32), [sc]]).astype(np.float32)


@pytest.mark.parametrize("which,th,ratio", [("fuse", 4.0, 1.0), ("search", 6.0, 1.0), ("search", 10.0, 0.75), ("search_kfs", 8.0, 1.0)])
def test_projection_with_sim3_pose(two_frames, kf_target, which, th, ratio):
    (k1, d1, u1, dep1), (k2, d2, u2, dep2), sf = two_frames
    rng = np.random.default_rng(int(th * 10))
    T, isg, F = kf_target["T"], kf_target["isg"], kf_target["F"]
    s = 1.03
    S = _sim3(T[:4], T[4:].astype(np.float64) * s, s)                    # Scw: Tcw = (R, t / s) is the target pose again
    has_mp = rng.random(len(k2)) < 0.3
    K = po.RefKeyFrame(k2, d2, u2, None, has_mp, None, sf, sf * sf, CAM6[:4], T)
    po.ref2_kf_set_geometry(K, F, isg, BF)
    q = _queries(k1, d1, dep1, rng, sf, 1.0 if which != "fuse" else 0.3)
    ok = q["bad"] == 0
    qq = _keep(q, ok)
    if which == "fuse":
        rm, rn, Tcw7, Ow = po.ref2_fuse_sim3(K, S, q, th)
        om, on, _ = po.search_keyframe(1, k2, d2, u2, BOUNDS, sf, isg, kf_target["logsf"], CAM6, Tcw7, Ow, qq["world_pos"], qq["normal"],
                                       qq["max_dist"], qq["min_dist"], qq["desc"], qq["angle"], None, th, 50.0)
    else:
        claimed = (rng.random(len(k2)) < 0.2).astype(np.uint8)
        rm, rn, Tcw7, Ow = po.ref2_search_kf_sim3(K, S, q, claimed, th, ratio, with_kfs=(which == "search_kfs"))
        om, on, _ = po.search_keyframe(2, k2, d2, u2, BOUNDS, sf, isg, kf_target["logsf"], CAM6, Tcw7, Ow, qq["world_pos"], qq["normal"],
                                       qq["max_dist"], qq["min_dist"], qq["desc"], qq["angle"], claimed, th, np.float32(50) * np.float32(ratio))
    got = np.full(len(ok), -1, np.int32)
    got[ok] = om
    assert on == rn and (got == rm).all()
    assert rn > 100


@pytest.mark.parametrize("th,orb_dist,check", [(10.0, 100, True), (3.0, 64, True), (10.0, 100, False)])
def test_search_by_projection_frame_from_keyframe(two_frames, kf_target, th, orb_dist, check):
    """Relocalisation's SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist): the queries are the keyframe's own map points."""
    (k1, d1, u1, dep1), (k2, d2, u2, dep2), sf = two_frames
    rng = np.random.default_rng(orb_dist)
    T, isg, F = kf_target["T"], kf_target["isg"], kf_target["F"]
    has_mp = dep1 > 0
    bad = has_mp & (rng.random(len(k1)) < 0.04)
    pts = _unproject(k1, np.where(has_mp, dep1, 1.0))
    dist = np.linalg.norm(pts, axis=1).astype(np.float32)
    maxd = (dist * sf[k1["octave"]]).astype(np.float32)
    mind = (maxd / sf[7]).astype(np.float32)
    K1 = po.RefKeyFrame(k1, d1, u1, None, has_mp, bad, sf, sf * sf, CAM6[:4])
    po.ref2_kf_set_mappoints(K1, pts, maxd, mind, d1)
    already = (rng.random(len(k1)) < 0.1).astype(np.uint8)
    claimed = (rng.random(len(k2)) < 0.1).astype(np.uint8)
    rfm, rn = po.ref2_search_frame_kf(F, T, K1, already, claimed, th, orb_dist, check)
    qsel = np.nonzero(has_mp & ~bad & (already == 0))[0]
    Ow, _, _ = _camera_center(T)
    om, on, _ = po.search_keyframe(3, k2, d2, u2, BOUNDS, sf, isg, kf_target["logsf"], CAM6, T, Ow, pts[qsel], None, maxd[qsel], mind[qsel], d1[qsel],
                                   k1["angle"][qsel].astype(np.float32), claimed, th, float(orb_dist), check_ori=check)
    got = np.full(len(k2), -1, np.int32)
    got[om[om >= 0]] = qsel[om >= 0]
    assert on == rn and (got == rfm).all()
    assert rn > 50


def test_search_by_sim3(two_frames, kf_target):
    """SearchBySim3: both directions, the mutual-consistency check (ORBmatcher.cc:1920-1945) and pre-existing matches."""
    (k1, d1, u1, dep1), (k2, d2, u2, dep2), sf = two_frames
    rng = np.random.default_rng(21)
    isg, logsf = kf_target["isg"], kf_target["logsf"]
    T1 = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    T2 = kf_target["T"]
    Ow2, R2, t2 = _camera_center(T2)
    s = 1.02
    q12 = np.array([-T2[0], -T2[1], -T2[2], T2[3]], np.float64)
    S12 = _sim3(q12, -(R2.T @ t2) * s, s)
    S21 = _sim3(T2[:4], t2 / s, 1.0 / s)
    dep2c = np.where(u2 > 0, BF / np.maximum(k2["x"] - u2, 1e-3), -1).astype(np.float32)

    def points(k, d, dep, R, t, frac):
        has = (dep > 0) & (rng.random(len(k)) < frac)
        pc = _unproject(k, np.where(dep > 0, dep, 1.0)).astype(np.float64)
        pw = ((pc - t) @ R).astype(np.float32)
        dist = np.linalg.norm(pc, axis=1).astype(np.float32)
        maxd = (dist * sf[k["octave"]]).astype(np.float32)
        return has, pw, maxd, (maxd / sf[7]).astype(np.float32)

    has1, pw1, mx1, mn1 = points(k1, d1, dep1, np.eye(3), np.zeros(3), 0.8)
    has2, pw2, mx2, mn2 = points(k2, d2, dep2c, R2, t2, 0.8)
    F1 = po.RefFrame(k1, d1, u1, BOUNDS, sf, CAM6, T1)
    K1 = po.RefKeyFrame(k1, d1, u1, None, has1, None, sf, sf * sf, CAM6[:4], T1)
    K2 = po.RefKeyFrame(k2, d2, u2, None, has2, None, sf, sf * sf, CAM6[:4], T2)
    po.ref2_kf_set_geometry(K1, F1, isg, BF)
    po.ref2_kf_set_geometry(K2, kf_target["F"], isg, BF)
    po.ref2_kf_set_mappoints(K1, pw1, mx1, mn1, d1)
    po.ref2_kf_set_mappoints(K2, pw2, mx2, mn2, d2)
    for pre in (False, True):
        m_in = np.full(len(k1), -1, np.int32)
        if pre:                                                          # a few matches exist already (vpMatches12 from SearchByBoW)
            i1 = rng.choice(np.nonzero(has1)[0], 40, replace=False)
            m_in[i1] = rng.choice(np.nonzero(has2)[0], 40, replace=False)
        rm, rn = po.ref2_search_by_sim3(K1, K2, S12, S21, 7.5, m_in)
        a1 = has1 & (m_in < 0)
        a2 = has2.copy()
        a2[m_in[m_in >= 0]] = False
        q1, q2 = np.nonzero(a1)[0], np.nonzero(a2)[0]
        zero = np.zeros(3, np.float32)
        r12, _, _ = po.search_keyframe(4, k2, d2, None, BOUNDS, sf, isg, logsf, CAM6, T1, zero, pw1[q1], None, mx1[q1], mn1[q1], d1[q1], None, None,
                                       7.5, 100.0, sim3=S21)
        r21, _, _ = po.search_keyframe(4, k1, d1, None, BOUNDS, sf, isg, logsf, CAM6, T2, zero, pw2[q2], None, mx2[q2], mn2[q2], d2[q2], None, None,
                                       7.5, 100.0, sim3=S12)
        vn1 = {int(i1): int(i2) for i1, i2 in zip(q1, r12) if i2 >= 0}
        vn2 = {int(i2): int(i1) for i2, i1 in zip(q2, r21) if i1 >= 0}
        want = m_in.copy()
        nfound = 0
        for i1, i2 in vn1.items():
            if vn2.get(i2, -1) == i1:
                want[i1] = i2
                nfound += 1
        assert rn == nfound and (rm == want).all()
        assert nfound > 20

"""CPU tier: the C++ oracle against plain-Python restatements written independently from the reference (tests/pyref.py):
SearchForInitialization, SearchByBoW(KF, KF), Fuse x2.  Small inputs (pure-Python loops)."""
import numpy as np
import pytest

import pyref
from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import synth

W, H = 320, 240
FX, FY, CX, CY, BF, B = 217.6, 217.6, 160.0, 120.0, 24.0, 0.11
BOUNDS = [0, W, 0, H]


@pytest.fixture(scope="module")
def frames():
    l, r, _ = synth.stereo_pair(W, H, seed=77)
    eL, eR = po.OracleExtractor(500, 1.2, 8, 20, 7), po.OracleExtractor(500, 1.2, 8, 20, 7)
    _, kL, dL = eL(l)
    _, kR, dR = eR(r)
    uR, dep, _ = po.stereo_matches(eL, eR, kL, dL, kR, dR, BF, B)
    _, k2, d2 = eL(np.clip(np.roll(l, (2, 1), (1, 0)).astype(int) + np.random.default_rng(0).integers(-2, 3, (H, W)), 0, 255).astype(np.uint8))
    return eL, kL, dL, uR, dep, k2, d2


def test_search_for_initialization(frames):
    eL, kL, dL, uR, dep, k2, d2 = frames
    rng = np.random.default_rng(1)
    l0 = np.nonzero(kL["octave"] == 0)[0]
    dup = rng.choice(l0, len(l0) // 3, replace=False)          # duplicated queries: match stealing
    kp1, d1 = np.concatenate([kL[dup], kL]), np.concatenate([dL[dup], dL])
    prev = np.stack([kp1["x"], kp1["y"]], 1).astype(np.float32)
    for win, ratio, check in [(100, 0.9, True), (20, 0.8, False)]:
        m, nm = po.search_initialization(kp1, d1, prev, k2, d2, BOUNDS, win, ratio, check)
        rm, rnm = pyref.search_for_initialization(kp1, d1, prev, k2, d2, BOUNDS, win, ratio, check)
        assert nm == rnm and (m == rm).all() and nm > 20


def test_search_by_bow_keyframes(frames):
    eL, kL, dL, uR, dep, k2, d2 = frames
    rng = np.random.default_rng(2)
    node_of = lambda d: ((d[:, 0].astype(np.int32) >> 3) * 7 + (d[:, 5].astype(np.int32) >> 4) * 3 + (d[:, 17].astype(np.int32) >> 5)) % 23
    q = np.nonzero(rng.random(len(kL)) < 0.8)[0]
    q = np.concatenate([q, q[::4]])
    nd = node_of(dL[q])
    order = np.lexsort((q, nd))
    q, nd = q[order], nd[order]
    node2 = node_of(d2)
    node2[::13] = -1
    valid2 = (rng.random(len(k2)) < 0.7).astype(np.uint8)
    for ratio, check in [(0.75, True), (0.9, False)]:
        m, nm = po.search_bow_kf(k2, d2, node2, valid2, nd, kL["angle"][q], dL[q], ratio, check)
        rm, rnm = pyref.search_bow_keyframes(k2, d2, node2, valid2, nd, kL["angle"][q], dL[q], ratio, check)
        assert nm == rnm and (m == rm).all() and nm > 10


@pytest.mark.parametrize("variant", [0, 1])
def test_fuse(frames, variant):
    eL, kL, dL, uR, dep, k2, d2 = frames
    rng = np.random.default_rng(3 + variant)
    sf = eL.scale_factors
    isg = (1.0 / (sf * sf)).astype(np.float32)
    logsf = po.logf(1.2)
    sel = np.nonzero(dep > 0)[0]
    z = dep[sel]
    pts = np.stack([(kL["x"][sel] - CX) * z / FX, (kL["y"][sel] - CY) * z / FY, z], 1).astype(np.float32)
    dist = np.linalg.norm(pts, axis=1).astype(np.float32)
    nrm = (pts / dist[:, None] + rng.normal(0, 0.3, pts.shape)).astype(np.float32)
    nrm = (nrm / np.linalg.norm(nrm, axis=1)[:, None]).astype(np.float32)
    maxd = np.float32(1.2) * (dist * sf[kL["octave"][sel]]).astype(np.float32)
    mind = np.float32(0.8) * (maxd / np.float32(1.2) / sf[7]).astype(np.float32)
    zmid = float(np.median(z))
    T = np.array([0, 0.001, 0, 1, 2 * zmid / FX, 1 * zmid / FY, 0], np.float32)
    T[:4] /= np.linalg.norm(T[:4])
    Ow = (-T[4:]).astype(np.float32)
    ur2 = np.where(rng.random(len(k2)) < 0.6, k2["x"] - 5, -1).astype(np.float32)
    cam6 = [FX, FY, CX, CY, BF, B]
    m, nm, _ = po.search_keyframe(variant, k2, d2, ur2, BOUNDS, sf, isg, logsf, cam6, T, Ow, pts, nrm, maxd, mind, dL[sel], None, None, 4.0, 50.0)
    rm = pyref.fuse(variant, k2, d2, ur2, BOUNDS, sf, isg, logsf, cam6, T, Ow, pts, nrm, maxd, mind, dL[sel], 4.0)
    assert (m == rm).all() and nm == int((rm >= 0).sum()) and nm > 20


def test_search_local_points(frames):
    """The per-frame local-map search (greedy claims, level rule, stereo gate) -- the headline matcher."""
    eL, kL, dL, uR, dep, k2, d2 = frames
    rng = np.random.default_rng(9)
    sf = eL.scale_factors
    ur2 = np.where(rng.random(len(k2)) < 0.6, k2["x"] - 4, -1).astype(np.float32)
    sel = np.concatenate([np.arange(len(kL)), rng.integers(0, len(kL), 2 * len(kL))])     # several map points compete for a feature
    jit = rng.normal(0, 1.5, (len(sel), 2)).astype(np.float32)
    px, py = (kL["x"][sel] + 2 + jit[:, 0]).astype(np.float32), (kL["y"][sel] + 1 + jit[:, 1]).astype(np.float32)
    pxr = (px - 4 + rng.normal(0, 2, len(sel))).astype(np.float32)
    lvl = np.clip(kL["octave"][sel] + rng.integers(-1, 2, len(sel)), 0, 7).astype(np.int32)
    vc = rng.uniform(0.99, 1.0, len(sel)).astype(np.float32)
    claimed = (rng.random(len(k2)) < 0.1).astype(np.uint8)
    for th, ratio in [(1.0, 0.8), (3.0, 0.8), (5.0, 0.7)]:
        m, nm = po.search_local(k2, d2, ur2, BOUNDS, sf, px, py, pxr, lvl, vc, dL[sel], th, ratio, claimed=claimed)
        rm, rnm = pyref.search_local_points(k2, d2, ur2, BOUNDS, sf, px, py, pxr, lvl, vc, dL[sel], th, ratio, claimed)
        assert nm == rnm and (m == rm).all() and nm > 30, (th, nm, rnm)


def test_search_last_frame(frames):
    eL, kL, dL, uR, dep, k2, d2 = frames
    rng = np.random.default_rng(10)
    sf = eL.scale_factors
    ur2 = np.where(rng.random(len(k2)) < 0.6, k2["x"] - 4, -1).astype(np.float32)
    sel = np.nonzero(dep > 0)[0]
    sel = np.concatenate([sel, sel[::3]])
    z = dep[sel]
    pts = np.stack([(kL["x"][sel] - CX) * z / FX, (kL["y"][sel] - CY) * z / FY, z], 1).astype(np.float32)
    zmid = float(np.median(z))
    T = np.array([0, 0.001, 0, 1, 2 * zmid / FX, 1 * zmid / FY, 0], np.float32)
    T[:4] /= np.linalg.norm(T[:4])
    obs = (rng.random(len(sel)) < 0.7).astype(np.uint8)
    cam6 = np.float32([FX, FY, CX, CY, BF, B])
    for th, direction, check in [(15.0, 0, True), (7.0, 1, True), (7.0, 2, False)]:
        fm, nm = po.search_last(k2, d2, ur2, BOUNDS, sf, cam6, T, direction, pts, kL["octave"][sel], kL["angle"][sel], dL[sel], obs, th, check)
        rfm, rnm = pyref.search_last_frame(k2, d2, ur2, BOUNDS, sf, cam6, T, direction, pts, kL["octave"][sel], kL["angle"][sel], dL[sel], obs, th, check)
        assert nm == rnm and (fm == rfm).all() and nm > 20, (th, direction, nm, rnm)


def test_compute_stereo_matches():
    """Frame::ComputeStereoMatches: row-band candidates, Hamming coarse match, 11 x 11 SAD refinement over 11 shifts, parabola,
    median outlier filter -- bit-exact mvuRight / mvDepth."""
    l, r, _ = synth.stereo_pair(W, H, seed=78)
    eL, eR = po.OracleExtractor(500, 1.2, 8, 20, 7), po.OracleExtractor(500, 1.2, 8, 20, 7)
    _, kL, dL = eL(l)
    _, kR, dR = eR(r)
    uR, dep, _ = po.stereo_matches(eL, eR, kL, dL, kR, dR, BF, B)
    pyrL, pyrR = [eL.level_pyramid(i) for i in range(8)], [eR.level_pyramid(i) for i in range(8)]
    ruR, rdep = pyref.compute_stereo_matches(kL, dL, kR, dR, pyrL, pyrR, eL.scale_factors, eL.inv_scale_factors, BF, B)
    assert (uR.view(np.uint32) == ruR.view(np.uint32)).all() and (dep.view(np.uint32) == rdep.view(np.uint32)).all()
    assert (dep > 0).sum() > 100


def test_search_by_bow_frame(frames):
    eL, kL, dL, uR, dep, k2, d2 = frames
    rng = np.random.default_rng(12)
    node_of = lambda d: ((d[:, 0].astype(np.int32) >> 3) * 7 + (d[:, 5].astype(np.int32) >> 4) * 3 + (d[:, 17].astype(np.int32) >> 5)) % 19
    fnode = node_of(d2)
    fnode[::11] = -1
    q = np.nonzero(rng.random(len(kL)) < 0.8)[0]
    q = np.concatenate([q, q[::3]])
    nd = node_of(dL[q])
    order = np.lexsort((q, nd))
    q, nd = q[order], nd[order]
    for ratio, check in [(0.7, True), (0.9, False)]:
        fm, nm = po.search_bow(k2, d2, fnode, nd, kL["angle"][q], dL[q], ratio, check)
        rfm, rnm = pyref.search_bow_frame(k2, d2, fnode, nd, kL["angle"][q], dL[q], ratio, check)
        assert nm == rnm and (fm == rfm).all() and nm > 10


def test_search_for_triangulation(frames):
    eL, kL, dL, uR, dep, k2, d2 = frames
    rng = np.random.default_rng(8)
    node_of = lambda d: ((d[:, 0].astype(np.int32) >> 3) * 7 + (d[:, 5].astype(np.int32) >> 4) * 3 + (d[:, 17].astype(np.int32) >> 5)) % 19
    sel = np.nonzero(rng.random(len(kL)) < 0.7)[0]
    nd = node_of(dL[sel])
    node2 = node_of(d2).astype(np.int32)
    node2[::11] = -1
    valid2 = (rng.random(len(k2)) < 0.8).astype(np.uint8)
    K = np.array([[FX, 0, CX], [0, FY, CY], [0, 0, 1]], np.float32)
    t = np.array([0.03, 0.002, 0.01], np.float32)
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]], np.float32)
    F12 = (np.linalg.inv(K.T).astype(np.float32) @ tx @ np.eye(3, dtype=np.float32) @ np.linalg.inv(K).astype(np.float32)).astype(np.float32)
    ep = np.array([150.0, 118.0], np.float32)
    sf, s2 = eL.scale_factors, eL.level_sigma2
    ur2 = np.where(rng.random(len(k2)) < 0.5, 1.0, -1.0)
    total = 0
    for coarse, check, mono_all in [(False, True, False), (True, True, False), (False, False, True), (True, False, True)]:
        st1 = np.zeros(len(sel), np.uint8) if mono_all else (uR[sel] >= 0).astype(np.uint8)
        st2 = np.zeros(len(k2), np.uint8) if mono_all else (ur2 >= 0).astype(np.uint8)
        m, nm = po.search_triangulation(kL[sel], dL[sel], nd, st1, k2, d2, node2, valid2, st2, F12, ep, sf, s2, coarse, check)
        rm, rnm = pyref.search_for_triangulation(kL[sel], dL[sel], nd, st1, k2, d2, node2, valid2, st2, F12, ep, sf, s2, coarse, check)
        assert nm == rnm and (m == rm).all(), (coarse, check, mono_all)
        total += nm
    assert total > 40


def test_search_by_projection_sim3(frames):
    """SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (ORBmatcher.cc:495-618): claims carried from query to query,
    some features already matched on entry."""
    eL, kL, dL, uR, dep, k2, d2 = frames
    rng = np.random.default_rng(12)
    sf = eL.scale_factors
    isg = (1.0 / (sf * sf)).astype(np.float32)
    logsf = po.logf(1.2)
    sel = np.nonzero(dep > 0)[0]
    sel = np.concatenate([sel, sel[::3]])                      # duplicated map points compete for the same features
    z = dep[sel]
    pts = np.stack([(kL["x"][sel] - CX) * z / FX, (kL["y"][sel] - CY) * z / FY, z], 1).astype(np.float32)
    dist = np.linalg.norm(pts, axis=1).astype(np.float32)
    nrm = (pts / dist[:, None] + rng.normal(0, 0.3, pts.shape)).astype(np.float32)
    nrm = (nrm / np.linalg.norm(nrm, axis=1)[:, None]).astype(np.float32)
    maxd = np.float32(1.2) * (dist * sf[kL["octave"][sel]]).astype(np.float32)
    mind = np.float32(0.8) * (maxd / np.float32(1.2) / sf[7]).astype(np.float32)
    zmid = float(np.median(z))
    T = np.array([0, 0.001, 0, 1, 2 * zmid / FX, 1 * zmid / FY, 0], np.float32)
    T[:4] /= np.linalg.norm(T[:4])
    Ow = (-T[4:]).astype(np.float32)
    cam6 = [FX, FY, CX, CY, BF, B]
    pre = (rng.random(len(k2)) < 0.15).astype(np.uint8)        # vpMatched[idx] != NULL on entry
    for th, ratio in [(8.0, 1.5), (4.0, 1.0)]:
        thr = float(np.float32(50) * np.float32(ratio))
        m, nm, cl = po.search_keyframe(2, k2, d2, None, BOUNDS, sf, isg, logsf, cam6, T, Ow, pts, nrm, maxd, mind, dL[sel], None, pre, th, thr)
        rm = pyref.fuse(2, k2, d2, None, BOUNDS, sf, isg, logsf, cam6, T, Ow, pts, nrm, maxd, mind, dL[sel], th, thr, pre)
        assert (m == rm).all() and nm == int((rm >= 0).sum()) and nm > 20
        assert len(set(m[m >= 0].tolist())) == nm and not pre[m[m >= 0]].any()     # one map point per feature, none on a pre-matched one


def test_search_by_projection_reloc(frames):
    """SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:2196-2330)."""
    eL, kL, dL, uR, dep, k2, d2 = frames
    rng = np.random.default_rng(21)
    sf = eL.scale_factors
    isg = (1.0 / (sf * sf)).astype(np.float32)
    logsf = po.logf(1.2)
    sel = np.nonzero(dep > 0)[0]
    sel = np.concatenate([sel, sel[::4]])
    z = dep[sel]
    pts = np.stack([(kL["x"][sel] - CX) * z / FX, (kL["y"][sel] - CY) * z / FY, z], 1).astype(np.float32)
    pts[::17, 2] *= -1                                           # behind the camera: this overload has no depth-sign test
    dist = np.linalg.norm(pts, axis=1).astype(np.float32)
    maxd = np.float32(1.2) * (dist * sf[kL["octave"][sel]]).astype(np.float32)
    mind = np.float32(0.8) * (maxd / np.float32(1.2) / sf[7]).astype(np.float32)
    zmid = float(np.median(np.abs(z)))
    T = np.array([0, 0.001, 0.002, 1, 2 * zmid / FX, 1 * zmid / FY, 0], np.float32)
    T[:4] /= np.linalg.norm(T[:4])
    Ow = (-T[4:]).astype(np.float32)
    cam6 = [FX, FY, CX, CY, BF, B]
    pre = (rng.random(len(k2)) < 0.2).astype(np.uint8)          # CurrentFrame.mvpMapPoints[i2] != NULL on entry
    ang = (kL["angle"][sel] + rng.choice([0.0, 0.0, 0.0, 90.0], len(sel))).astype(np.float32) % np.float32(360)
    total = 0
    for th, orb_dist, check in [(10.0, 100, True), (3.0, 64, True), (10.0, 100, False)]:
        m, nm, cl = po.search_keyframe(3, k2, d2, None, BOUNDS, sf, isg, logsf, cam6, T, Ow, pts, None, maxd, mind, dL[sel], ang, pre, th,
                                       float(orb_dist), check)
        rm, rcl = pyref.search_by_projection_reloc(k2, d2, BOUNDS, sf, logsf, cam6, T, Ow, pts, maxd, mind, dL[sel], ang, pre, th, orb_dist, check)
        assert (m == rm).all() and nm == int((rm >= 0).sum()) and (cl.astype(bool) == rcl).all(), (th, orb_dist, check)
        total += nm
    assert total > 60


def test_search_by_sim3_one_direction(frames):
    """One direction of SearchBySim3 (ORBmatcher.cc:1719-1790) with a non-trivial similarity (rotation, translation, scale 1.07)."""
    eL, kL, dL, uR, dep, k2, d2 = frames
    rng = np.random.default_rng(31)
    sf = eL.scale_factors
    isg = (1.0 / (sf * sf)).astype(np.float32)
    logsf = po.logf(1.2)
    sel = np.nonzero(dep > 0)[0]
    z = dep[sel]
    pts = np.stack([(kL["x"][sel] - CX) * z / FX, (kL["y"][sel] - CY) * z / FY, z], 1).astype(np.float32)
    T = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)             # the other keyframe's pose: identity (points are in its camera frame)
    s = 1.07
    q = np.array([0.004, -0.006, 0.003, 1.0])
    q = q / np.linalg.norm(q) * np.sqrt(s)                      # RxSO3 stores sqrt(scale) * unit quaternion
    zmid = float(np.median(z))
    S8 = np.array([*q, 2 * zmid / FX, 1 * zmid / FY, -0.07 * zmid, 0], np.float32)
    S8[7] = np.float32(np.dot(S8[:4], S8[:4]))                  # scale = quaternion().squaredNorm() in float
    pc = s * pts + S8[4:7]
    dist = np.linalg.norm(pc, axis=1).astype(np.float32)
    maxd = np.float32(1.2) * (dist * sf[kL["octave"][sel]]).astype(np.float32)
    mind = np.float32(0.8) * (maxd / np.float32(1.2) / sf[7]).astype(np.float32)
    cam6 = [FX, FY, CX, CY, BF, B]
    total = 0
    for th in (7.5, 3.0):
        m, nm, _ = po.search_keyframe(4, k2, d2, None, BOUNDS, sf, isg, logsf, cam6, T, np.zeros(3, np.float32), pts, None, maxd, mind,
                                      dL[sel], None, None, th, 100.0, True, S8)
        rm = pyref.search_by_sim3_oneway(k2, d2, BOUNDS, sf, logsf, cam6, T, S8, pts, maxd, mind, dL[sel], th)
        assert (m == rm).all() and nm == int((rm >= 0).sum()), th
        total += nm
    assert total > 40

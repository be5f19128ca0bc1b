"""GPU tier: SearchByProjection (local map points / last frame) on the device against the CPU oracle.
Bar: identical match pairs (bit-exact integer arrays) and identical nmatches."""
import numpy as np
import pytest

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import ORBextractor, ORBmatcher, camera, synth

pytestmark = pytest.mark.gpu

W, H = 640, 480
FX, FY, CX, CY, BF, B = 435.2, 435.2, 320.0, 240.0, 47.9, 0.11
CAM6 = np.array([FX, FY, CX, CY, BF, B], np.float32)
BOUNDS = np.array([0, W, 0, H], np.float32)


def quat_pose(yaw_deg, t):
    a = np.deg2rad(yaw_deg) / 2
    return np.array([0, np.sin(a), 0, np.cos(a), *t], np.float32)   # qx qy qz qw tx ty tz


def make_scene(P, seed):
    """P stereo frames: 'last' = pair seed, 'cur' = the same pair shifted by a few pixels + noise.
    Returns the interleaved image batch for the CURRENT frames and, per frame, the last frame's oracle features."""
    rng = np.random.default_rng(seed)
    imgs = np.zeros((2 * P, H, W), np.uint8)
    lasts = []
    for p in range(P):
        l, r, _ = synth.stereo_pair(W, H, seed=seed + 10 * p)
        eL, eR = po.OracleExtractor(1200, 1.2, 8, 20, 7), po.OracleExtractor(1200, 1.2, 8, 20, 7)
        _, kL, dL = eL(l)
        _, kR, dR = eR(r)
        uR, dep, _ = po.stereo_matches(eL, eR, kL, dL, kR, dR, BF, B)
        lasts.append((kL, dL, uR, dep))
        sh = (3 + p, 1)
        noise = rng.integers(-3, 4, (H, W))
        imgs[2 * p] = np.clip(np.roll(l, sh, (1, 0)).astype(int) + noise, 0, 255)
        imgs[2 * p + 1] = np.clip(np.roll(r, sh, (1, 0)).astype(int) + noise, 0, 255)
    return imgs, lasts


def unproject(k, depth):
    z = depth
    return np.stack([(k["x"] - CX) * z / FX, (k["y"] - CY) * z / FY, z], 1).astype(np.float32)


@pytest.fixture(scope="module")
def scene():
    P = 3
    imgs, lasts = make_scene(P, 500)
    ex = ORBextractor(1200, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=2 * P)
    ex.extract_batch(imgs)
    ex.stereo_batch(P, BF, B)
    n, mono, off, kps, desc = ex.download(2 * P)
    uR, dep = ex.stereo_download(int(off[-1]))
    yield dict(P=P, ex=ex, lasts=lasts, off=off, kps=kps, desc=desc, uR=uR)
    ex.close()


def test_search_local_points_matches_oracle(scene):
    P, ex, off = scene["P"], scene["ex"], scene["off"]
    rng = np.random.default_rng(7)
    sf = ex.GetScaleFactors()
    for th, nnratio in [(1.0, 0.8), (3.0, 0.8), (5.0, 0.7)]:
        qoff, px, py, pxr, lv, vc, qd = [0], [], [], [], [], [], []
        claimed = np.zeros(int(off[-1]), np.uint8)
        for p in range(P):
            kL, dL, luR, ldep = scene["lasts"][p]
            sel = np.nonzero(ldep > 0)[0]
            # the local map: last frame's stereo points + a few thousand repeats with jitter (several map
            # points competing for the same feature exercises the greedy claims)
            sel = np.concatenate([sel, rng.choice(sel, 3 * len(sel))])
            jit = rng.normal(0, 1.5, (len(sel), 2)).astype(np.float32)
            x = (kL["x"][sel] + 3 + p + jit[:, 0]).astype(np.float32)
            y = (kL["y"][sel] + 1 + jit[:, 1]).astype(np.float32)
            px.append(x); py.append(y)
            pxr.append((x - BF / ldep[sel]).astype(np.float32))
            lv.append(np.clip(kL["octave"][sel] + rng.integers(-1, 2, len(sel)), 0, 7).astype(np.int32))
            vc.append(rng.uniform(0.99, 1.0, len(sel)).astype(np.float32))
            qd.append(dL[sel])
            qoff.append(qoff[-1] + len(sel))
            a, b = off[2 * p], off[2 * p + 1]
            claimed[a:b] = rng.random(b - a) < 0.15     # features already matched by the motion model
        px, py, pxr, lv, vc, qd = map(np.concatenate, (px, py, pxr, lv, vc, qd))
        m = ORBmatcher(nnratio, True)
        match, nm = m.SearchByProjection(ex, camera(FX, FY, CX, CY, BF, B, W, H), [2 * p for p in range(P)], qoff,
                                         px, py, pxr, lv, vc, qd, th=th, feature_claimed=claimed)
        for p in range(P):
            a, b = off[2 * p], off[2 * p + 1]
            s = slice(qoff[p], qoff[p + 1])
            rmatch, rnm = po.search_local(scene["kps"][a:b], scene["desc"][a:b], scene["uR"][a:b], BOUNDS, sf, px[s], py[s],
                                          pxr[s], lv[s], vc[s], qd[s], th, nnratio, claimed=claimed[a:b])
            assert rnm == nm[p], (th, p, rnm, nm[p])
            assert (match[s] == rmatch).all(), (th, p)
            assert rnm > 50


def test_search_last_frame_matches_oracle(scene):
    P, ex, off = scene["P"], scene["ex"], scene["off"]
    rng = np.random.default_rng(11)
    sf = ex.GetScaleFactors()
    total = int(off[-1])
    for th, direction_all, check in [(15.0, 0, True), (7.0, 1, True), (7.0, 2, False)]:
        qoff, xw, lo, la, qd, ob, Tcw, dirs = [0], [], [], [], [], [], [], []
        for p in range(P):
            kL, dL, luR, ldep = scene["lasts"][p]
            sel = np.nonzero(ldep > 0)[0]
            sel = np.concatenate([sel, sel[: len(sel) // 3]])          # duplicated map points: zero-observation overwrite path
            pts = unproject(kL[sel], ldep[sel])
            # camera moved so that the points shift ~ (3+p, 1) px at mid depth
            zmid = float(np.median(ldep[sel]))
            T = quat_pose(0.05 * p, [(3 + p) * zmid / FX, 1 * zmid / FY, 0.0])
            xw.append(pts); lo.append(kL["octave"][sel].astype(np.int32)); la.append(kL["angle"][sel].astype(np.float32))
            qd.append(dL[sel]); ob.append((rng.random(len(sel)) < 0.7).astype(np.uint8))
            Tcw.append(T); dirs.append(direction_all)
            qoff.append(qoff[-1] + len(sel))
        xw, lo, la, qd, ob = map(np.concatenate, (xw, lo, la, qd, ob))
        m = ORBmatcher(0.9, check)
        fm, nm = m.SearchByProjectionLastFrame(ex, camera(FX, FY, CX, CY, BF, B, W, H), [2 * p for p in range(P)], qoff,
                                               np.stack(Tcw), dirs, xw, lo, la, qd, ob, th, total)
        for p in range(P):
            a, b = off[2 * p], off[2 * p + 1]
            s = slice(qoff[p], qoff[p + 1])
            rfm, rnm = po.search_last(scene["kps"][a:b], scene["desc"][a:b], scene["uR"][a:b], BOUNDS, sf, CAM6, Tcw[p],
                                      dirs[p], xw[s], lo[s], la[s], qd[s], ob[s], th, check)
            got = fm[a:b].copy()
            got[got >= 0] -= qoff[p]
            assert rnm == nm[p], (th, p, rnm, nm[p])
            assert (got == rfm).all(), (th, p)
            assert rnm > 50
            assert (fm[off[2 * p + 1]:off[2 * p + 2]] == -1).all()      # right-eye rows untouched


def test_search_with_no_queries_and_monocular(scene):
    ex, off = scene["ex"], scene["off"]
    m = ORBmatcher(0.8, True)
    match, nm = m.SearchByProjection(ex, camera(FX, FY, CX, CY, BF, B, W, H), [0], [0, 0], np.zeros(0), np.zeros(0),
                                     np.zeros(0), np.zeros(0, np.int32), np.zeros(0), np.zeros((0, 32), np.uint8))
    assert len(match) == 0 and nm[0] == 0


def test_search_by_bow_matches_oracle(scene):
    P, ex, off = scene["P"], scene["ex"], scene["off"]
    total = int(off[-1])
    # a stand-in vocabulary: node = a hash of three descriptor bytes into 97 buckets (what matters to the matcher is only
    # the partition into nodes; the real ids come from DBoW2's tree, out of scope)
    def node_of(d):
        return ((d[:, 0].astype(np.int32) >> 3) * 7 + (d[:, 5].astype(np.int32) >> 4) * 3 + (d[:, 17].astype(np.int32) >> 5)) % 97
    feat_node = np.full(total, -1, np.int32)
    for p in range(P):
        a, b = off[2 * p], off[2 * p + 1]
        feat_node[a:b] = node_of(scene["desc"][a:b])
        feat_node[a:b][::17] = -1                       # a few features without a BoW entry
    for nnratio, check in [(0.7, True), (0.9, False)]:
        qoff, qn, qa, qd = [0], [], [], []
        for p in range(P):
            kL, dL, luR, ldep = scene["lasts"][p]
            sel = np.nonzero(ldep > 0)[0]                # keyframe features holding a map point
            nd = node_of(dL[sel])
            order = np.lexsort((sel, nd))               # FeatureVector merge order: node, then feature index
            sel, nd = sel[order], nd[order]
            qn.append(nd.astype(np.int32)); qa.append(kL["angle"][sel].astype(np.float32)); qd.append(dL[sel])
            qoff.append(qoff[-1] + len(sel))
        qn, qa, qd = map(np.concatenate, (qn, qa, qd))
        m = ORBmatcher(nnratio, check)
        fm, nm = m.SearchByBoW(ex, [2 * p for p in range(P)], qoff, qn, qa, qd, feat_node, total)
        for p in range(P):
            a, b = off[2 * p], off[2 * p + 1]
            s = slice(qoff[p], qoff[p + 1])
            rfm, rnm = po.search_bow(scene["kps"][a:b], scene["desc"][a:b], feat_node[a:b], qn[s], qa[s], qd[s], nnratio, check)
            got = fm[a:b].copy()
            got[got >= 0] -= qoff[p]
            assert rnm == nm[p], (nnratio, p, rnm, nm[p])
            assert (got == rfm).all(), (nnratio, p)
            assert rnm > 20


def test_search_for_triangulation_matches_oracle(scene):
    # KF1 = the last frame of sequence 0 (oracle features), KF2 = the current left image of the same sequence
    ex, off = scene["ex"], scene["off"]
    rng = np.random.default_rng(5)
    kL, dL, luR, ldep = scene["lasts"][0]
    a, b = off[0], off[1]
    k2, d2, ur2 = scene["kps"][a:b], scene["desc"][a:b], scene["uR"][a:b]
    node_of = lambda d: ((d[:, 0].astype(np.int32) >> 3) * 7 + (d[:, 5].astype(np.int32) >> 4) * 3 + (d[:, 17].astype(np.int32) >> 5)) % 61
    sel = np.nonzero(rng.random(len(kL)) < 0.6)[0]                 # KF1 features without a map point
    nd = node_of(dL[sel])
    order = np.lexsort((sel, nd))
    sel, nd = sel[order], nd[order]
    node2 = node_of(d2).astype(np.int32)
    node2[::13] = -1
    valid2 = (rng.random(len(k2)) < 0.7).astype(np.uint8)
    # fundamental matrix of a small sideways + forward motion (float32, as the shim would compute it with Eigen)
    K = np.array([[FX, 0, CX], [0, FY, CY], [0, 0, 1]], np.float32)
    t = np.array([0.03, 0.002, 0.01], np.float32)
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]], np.float32)
    R = np.eye(3, dtype=np.float32)
    F12 = (np.linalg.inv(K.T).astype(np.float32) @ tx @ R @ np.linalg.inv(K).astype(np.float32)).astype(np.float32)
    ep = np.array([310.0, 236.0], np.float32)
    sf, s2 = ex.GetScaleFactors(), ex.GetScaleSigmaSquares()
    for coarse, check, mono_all in [(False, True, False), (True, True, False), (False, False, True)]:
        st1 = (np.zeros(len(sel), np.uint8) if mono_all else (luR[sel] >= 0).astype(np.uint8))
        st2 = (np.zeros(len(k2), np.uint8) if mono_all else (ur2 >= 0).astype(np.uint8))
        m = ORBmatcher(0.6, check)
        got, nm = m.SearchForTriangulation(ex, kL[sel], dL[sel], nd, st1, k2, d2, node2, valid2, st2, F12, ep, bCoarse=coarse)
        ref, rnm = po.search_triangulation(kL[sel], dL[sel], nd, st1, k2, d2, node2, valid2, st2, F12, ep, sf, s2, coarse, check)
        assert rnm == nm and (got == ref).all(), (coarse, check, mono_all, rnm, nm)
        assert coarse is False or rnm > 20


def test_search_local_points_one_map_point_per_feature(scene):
    # the benchmark's query shape: every feature is the projection of its own map point (+ unrelated points)
    P, ex, off = scene["P"], scene["ex"], scene["off"]
    sf = ex.GetScaleFactors()
    rng = np.random.default_rng(3)
    qoff, parts = [0], []
    for p in range(P):
        a, b = off[2 * p], off[2 * p + 1]
        k, d, ur = scene["kps"][a:b], scene["desc"][a:b], scene["uR"][a:b]
        n = len(k)
        z = np.where(ur > 0, BF / np.maximum(k["x"] - ur, 0.01), 5.0).astype(np.float32)
        jit = rng.normal(0, 1.5, (n, 2)).astype(np.float32)
        x = np.concatenate([k["x"] + jit[:, 0], rng.uniform(20, W - 20, n)]).astype(np.float32)
        y = np.concatenate([k["y"] + jit[:, 1], rng.uniform(20, H - 20, n)]).astype(np.float32)
        zq = np.concatenate([z, rng.uniform(2, 15, n)]).astype(np.float32)
        lvl = np.concatenate([k["octave"], rng.integers(0, 8, n)]).astype(np.int32)
        vc = rng.uniform(0.99, 1.0, 2 * n).astype(np.float32)
        dq = np.concatenate([d, rng.integers(0, 256, (n, 32), dtype=np.uint8)])
        parts.append((x, y, (x - BF / zq).astype(np.float32), lvl, vc, dq))
        qoff.append(qoff[-1] + 2 * n)
    px, py, pxr, lv, vc, qd = (np.concatenate([pt[i] for pt in parts]) for i in range(6))
    m = ORBmatcher(0.8, True)
    match, nm = m.SearchByProjection(ex, camera(FX, FY, CX, CY, BF, B, W, H), [2 * p for p in range(P)], qoff, px, py, pxr, lv, vc, qd, th=3.0)
    for p in range(P):
        a, b = off[2 * p], off[2 * p + 1]
        s = slice(qoff[p], qoff[p + 1])
        rmatch, rnm = po.search_local(scene["kps"][a:b], scene["desc"][a:b], scene["uR"][a:b], BOUNDS, sf, px[s], py[s], pxr[s], lv[s],
                                      vc[s], qd[s], 3.0, 0.8)
        assert rnm == nm[p] and (match[s] == rmatch).all(), (p, rnm, nm[p])
        assert rnm > 0.9 * (b - a)


# ---- projection searches into keyframes: Fuse x2, SearchByProjection(KF, Scw), SearchByProjection(F, KF, set) -------------------
def _camera_center(T):
    q, t = T[:4].astype(np.float64), T[4:].astype(np.float64)
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return (-R.T @ t).astype(np.float32)


def _kf_queries(scene, p, rng, sf, dup):
    """Map points seen by the 'last' frame of pair p, expressed in that frame (= world), with the MapPoint fields the
    searches read: normal, scale-invariance distances (MapPoint::UpdateNormalAndDepth, MapPoint.cc:600-660), descriptor."""
    kL, dL, luR, ldep = scene["lasts"][p]
    sel = np.nonzero(ldep > 0)[0]
    sel = np.concatenate([sel, rng.choice(sel, int(dup * len(sel)))])   # competing map points: greedy claims
    pts = unproject(kL[sel], ldep[sel]) + rng.normal(0, 0.002, (len(sel), 3)).astype(np.float32)
    dist = np.linalg.norm(pts, axis=1).astype(np.float32)
    nrm = pts / dist[:, None] + rng.normal(0, 0.2, pts.shape).astype(np.float32)
    nrm = (nrm / np.linalg.norm(nrm, axis=1)[:, None]).astype(np.float32)
    nrm[rng.random(len(sel)) < 0.05] *= -1                               # a few points seen from behind: viewing-angle gate
    maxd = (dist * sf[kL["octave"][sel]]).astype(np.float32)
    mind = (maxd / sf[7]).astype(np.float32)
    return dict(world_pos=pts, normal=nrm, max_dist=maxd, min_dist=mind,   # raw mfMaxDistance / mfMinDistance
                desc=dL[sel],
                angle=kL["angle"][sel].astype(np.float32))


@pytest.mark.parametrize("variant,th,thr", [(0, 3.0, 50.0), (0, 8.0, 50.0), (1, 4.0, 50.0), (2, 6.0, 50.0), (2, 10.0, 37.5), (3, 10.0, 100.0),
                                            (3, 3.0, 64.0)])
def test_search_keyframe_matches_oracle(scene, variant, th, thr):
    P, ex, off = scene["P"], scene["ex"], scene["off"]
    rng = np.random.default_rng(100 + variant)
    sf = ex.GetScaleFactors()
    isg = ex.GetInverseScaleSigmaSquares()
    logsf = po.logf(1.2)
    cam = camera(FX, FY, CX, CY, BF, B, W, H)
    targets, queries = [], []
    for p in range(P):
        a, b = off[2 * p], off[2 * p + 1]
        zmid = float(np.median(scene["lasts"][p][3][scene["lasts"][p][3] > 0]))
        T = quat_pose(0.05 * p, [(3 + p) * zmid / FX, 1 * zmid / FY, 0.0])
        cl = (rng.random(b - a) < 0.2).astype(np.uint8) if variant >= 2 else None
        targets.append(dict(kp=scene["kps"][a:b], desc=scene["desc"][a:b], uright=scene["uR"][a:b], claimed=cl, Tcw=T,
                            Ow=_camera_center(T)))
        queries.append(_kf_queries(scene, p, rng, sf, dup=1.0 if variant >= 2 else 0.3))
    m = ORBmatcher(0.9, True)
    got, nm = m._search_keyframe(ex, cam, variant, targets, queries, th, thr)
    total = 0
    for p in range(P):
        t, q = targets[p], queries[p]
        rm, rnm, _ = po.search_keyframe(variant, t["kp"], t["desc"], t["uright"], BOUNDS, sf, isg, logsf, CAM6, t["Tcw"], t["Ow"],
                                        q["world_pos"], q["normal"], q["max_dist"], q["min_dist"], q["desc"], q["angle"], t["claimed"],
                                        th, thr)
        assert rnm == nm[p], (variant, p, rnm, nm[p])
        assert (got[p] == rm).all(), (variant, p, int((got[p] != rm).sum()))
        total += rnm
    assert total > 100, total


def test_search_keyframe_reloc_on_device_frame(scene):
    """SearchByProjection(CurrentFrame, pKF, ...) with the frame still on the device (image 2p of the last batch)."""
    P, ex, off = scene["P"], scene["ex"], scene["off"]
    rng = np.random.default_rng(5)
    sf, isg, logsf = ex.GetScaleFactors(), ex.GetInverseScaleSigmaSquares(), po.logf(1.2)
    cam = camera(FX, FY, CX, CY, BF, B, W, H)
    claimed_rows = (rng.random(int(off[-1])) < 0.1).astype(np.uint8)
    targets, queries = [], []
    for p in range(P):
        zmid = float(np.median(scene["lasts"][p][3][scene["lasts"][p][3] > 0]))
        T = quat_pose(0.05 * p, [(3 + p) * zmid / FX, 1 * zmid / FY, 0.0])
        targets.append(dict(Tcw=T, Ow=_camera_center(T), claimed_rows=claimed_rows))
        queries.append(_kf_queries(scene, p, rng, sf, dup=0.5))
    for check in (True, False):
        m = ORBmatcher(0.9, check)
        got, nm = m.SearchByProjectionReloc(ex, cam, targets, queries, 10.0, 100, target_images=[2 * p for p in range(P)])
        for p in range(P):
            a, b = off[2 * p], off[2 * p + 1]
            t, q = targets[p], queries[p]
            rm, rnm, _ = po.search_keyframe(3, scene["kps"][a:b], scene["desc"][a:b], scene["uR"][a:b], BOUNDS, sf, isg, logsf, CAM6,
                                            t["Tcw"], t["Ow"], q["world_pos"], None, q["max_dist"], q["min_dist"], q["desc"], q["angle"],
                                            claimed_rows[a:b], 10.0, 100.0, check_ori=check)
            assert rnm == nm[p] and (got[p] == rm).all(), (check, p, rnm, nm[p])
            assert rnm > 50


def _quat_to_R(q):
    x, y, z, w = [float(v) for v in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _sim3(q_unit, t, s):
    """8 floats of a Sophus::Sim3f: RxSO3 quaternion (norm^2 = scale), translation, scale = squaredNorm in float32."""
    q = (np.asarray(q_unit, np.float64) * np.sqrt(s)).astype(np.float32)
    sc = np.float32(np.float32(q[0] * q[0] + q[1] * q[1]) + np.float32(q[2] * q[2] + q[3] * q[3]))
    return np.concatenate([q, np.asarray(t, np.float32), [sc]]).astype(np.float32)


def test_search_by_sim3_matches_oracle(scene):
    P, ex, off = scene["P"], scene["ex"], scene["off"]
    rng = np.random.default_rng(21)
    sf, isg, logsf = ex.GetScaleFactors(), ex.GetInverseScaleSigmaSquares(), po.logf(1.2)
    cam = camera(FX, FY, CX, CY, BF, B, W, H)
    for p in range(P):
        kL, dL, luR, ldep = scene["lasts"][p]                       # keyframe 1: world frame = its camera frame
        a, b = off[2 * p], off[2 * p + 1]
        k2, d2, u2 = scene["kps"][a:b], scene["desc"][a:b], scene["uR"][a:b]
        dep2 = np.where(u2 > 0, BF / np.maximum(k2["x"] - u2, 1e-3), -1).astype(np.float32)
        zmid = float(np.median(ldep[ldep > 0]))
        T1 = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
        T2 = quat_pose(0.05 * p, [(3 + p) * zmid / FX, 1 * zmid / FY, 0.0])
        R2, t2 = _quat_to_R(T2[:4]), T2[4:].astype(np.float64)
        s = 1.0 + 0.01 * p
        q12 = np.array([-T2[0], -T2[1], -T2[2], T2[3]], np.float64)  # R12 = R2^T
        t12 = -(R2.T @ t2) * s
        S12 = _sim3(q12, t12, s)
        S21 = _sim3(T2[:4], t2 / s * 1.0, 1.0 / s)                   # p2 = (1/s) R2 (p1 - t12) ~= R2 p1/s + t2/s

        def points(k, d, dep, Tcw, frac):
            sel = np.nonzero(dep > 0)[0]
            sel = sel[rng.random(len(sel)) < frac]                    # the rest: no map point / already matched
            pc = unproject(k[sel], dep[sel]).astype(np.float64)
            R, t = _quat_to_R(Tcw[:4]), Tcw[4:].astype(np.float64)
            pw = ((pc - t) @ R).astype(np.float32)
            dist = np.linalg.norm(pc, axis=1).astype(np.float32)
            maxd = (dist * sf[k["octave"][sel]]).astype(np.float32)
            return dict(index=sel, world_pos=pw, max_dist=maxd, min_dist=(maxd / sf[7]).astype(np.float32),
                        desc=d[sel])

        mp1, mp2 = points(kL, dL, ldep, T1, 0.8), points(k2, d2, dep2, T2, 0.8)
        kf1, kf2 = dict(kp=kL, desc=dL, Tcw=T1), dict(kp=k2, desc=d2, Tcw=T2)
        m = ORBmatcher(0.9, True)
        found, nfound = m.SearchBySim3(ex, cam, kf1, kf2, S12, S21, mp1, mp2, 7.5)
        zero = np.zeros(3, np.float32)
        r12, _, _ = po.search_keyframe(4, k2, d2, None, BOUNDS, sf, isg, logsf, CAM6, T1, zero, mp1["world_pos"], None, mp1["max_dist"],
                                       mp1["min_dist"], mp1["desc"], None, None, 7.5, 100.0, sim3=S21)
        r21, _, _ = po.search_keyframe(4, kL, dL, None, BOUNDS, sf, isg, logsf, CAM6, T2, zero, mp2["world_pos"], None, mp2["max_dist"],
                                       mp2["min_dist"], mp2["desc"], None, None, 7.5, 100.0, sim3=S12)
        vn1 = {int(i1): int(i2) for i1, i2 in zip(mp1["index"], r12) if i2 >= 0}
        vn2 = {int(i2): int(i1) for i2, i1 in zip(mp2["index"], r21) if i1 >= 0}
        ref = {i1: i2 for i1, i2 in vn1.items() if vn2.get(i2, -1) == i1}
        assert found == ref and nfound == len(ref)
        assert len(vn1) > 100 and len(vn2) > 100 and nfound > 20, (len(vn1), len(vn2), nfound)


def test_search_by_bow_keyframes_matches_oracle(scene):
    """SearchByBoW(pKF1, pKF2): KF1 = the last frame of each pair (oracle features), KF2 = the current left image."""
    P, ex, off = scene["P"], scene["ex"], scene["off"]
    rng = np.random.default_rng(31)
    node_of = lambda d: ((d[:, 0].astype(np.int32) >> 3) * 7 + (d[:, 5].astype(np.int32) >> 4) * 3 + (d[:, 17].astype(np.int32) >> 5)) % 97
    pairs = []
    for p in range(P):
        kL, dL, luR, ldep = scene["lasts"][p]
        a, b = off[2 * p], off[2 * p + 1]
        k2, d2 = scene["kps"][a:b], scene["desc"][a:b]
        node2 = node_of(d2)
        node2[::19] = -1
        valid2 = (rng.random(len(k2)) < 0.7).astype(np.uint8)
        sel = np.nonzero(rng.random(len(kL)) < 0.75)[0]             # KF1 features with a good map point
        sel = np.concatenate([sel, sel[:: 5]])                      # duplicates: the vbMatched2 claims matter
        nd = node_of(dL[sel])
        order = np.lexsort((sel, nd))
        sel, nd = sel[order], nd[order]
        pairs.append(dict(kp2=k2, desc2=d2, node2=node2, valid2=valid2, query_node=nd.astype(np.int32),
                          query_angle=kL["angle"][sel].astype(np.float32), desc1=dL[sel]))
    for nnratio, check in [(0.75, True), (0.9, False)]:
        m = ORBmatcher(nnratio, check)
        got, nm = m.SearchByBoWKeyFrames(ex, pairs)
        for p, pr in enumerate(pairs):
            rm, rnm = po.search_bow_kf(pr["kp2"], pr["desc2"], pr["node2"], pr["valid2"], pr["query_node"], pr["query_angle"], pr["desc1"],
                                       nnratio, check)
            assert rnm == nm[p] and (got[p] == rm).all(), (nnratio, p, rnm, nm[p])
            assert rnm > 50


def test_search_for_initialization_matches_oracle(scene):
    """SearchForInitialization(F1, F2, vbPrevMatched, ...): F1 = the last frame (oracle features), F2 = the current left image,
    on the device and as host arrays; duplicated F1 features make later, better matches steal a feature (:812-817)."""
    P, ex, off = scene["P"], scene["ex"], scene["off"]
    cam = camera(FX, FY, CX, CY, BF, B, W, H)
    rng = np.random.default_rng(41)
    for p in range(P):
        kL, dL, _, _ = scene["lasts"][p]
        a, b = off[2 * p], off[2 * p + 1]
        k2, d2 = scene["kps"][a:b], scene["desc"][a:b]
        # F1: the last frame plus noisy copies of a third of its level-0 features (their descriptors get a few flipped bits, so the
        # copy or the original can be the closer one)
        l0 = np.nonzero(kL["octave"] == 0)[0]
        dup = rng.choice(l0, len(l0) // 3, replace=False)
        kp1 = np.concatenate([kL[dup], kL])
        dd = dL[dup].copy()
        flips = rng.integers(0, 256, (len(dup), 3))
        for j in range(3):
            dd[np.arange(len(dup)), flips[:, j] // 8] ^= (1 << (flips[:, j] % 8)).astype(np.uint8)
        d1 = np.concatenate([dd, dL])
        prev = np.stack([kp1["x"], kp1["y"]], 1).astype(np.float32)        # vbPrevMatched = F1 keypoint positions (Tracking.cc:2560)
        for win, nnratio, check in [(100, 0.9, True), (30, 0.8, False)]:
            m = ORBmatcher(nnratio, check)
            g_dev, n_dev = m.SearchForInitialization(ex, cam, kp1, d1, prev, win, target_image=2 * p)
            g_host, n_host = m.SearchForInitialization(ex, cam, kp1, d1, prev, win, kp2=k2, desc2=d2)
            rm, rnm = po.search_initialization(kp1, d1, prev, k2, d2, BOUNDS, win, nnratio, check)
            assert n_dev == rnm and (g_dev == rm).all(), (p, win, n_dev, rnm, int((g_dev != rm).sum()))
            assert n_host == rnm and (g_host == rm).all(), (p, win)
            assert rnm > 50, rnm


def test_is_in_frustum_and_masked_local_search(scene):
    """Frame::isInFrustum on the device (bit-exact float fields vs the oracle), then SearchByProjection(F, vpMapPoints) fed by its
    outputs through the in_view mask == the oracle search over the compacted visible points."""
    from orb_slam3_detailed_comments_b200 import isInFrustum
    P, ex, off = scene["P"], scene["ex"], scene["off"]
    rng = np.random.default_rng(61)
    sf, logsf = ex.GetScaleFactors(), po.logf(1.2)
    cam = camera(FX, FY, CX, CY, BF, B, W, H)
    poff, Rs, ts, Os, qs = [0], [], [], [], []
    for p in range(P):
        zmid = float(np.median(scene["lasts"][p][3][scene["lasts"][p][3] > 0]))
        T = quat_pose(0.05 * p, [(3 + p) * zmid / FX, 1 * zmid / FY, 0.0])
        R = _quat_to_R(T[:4]).astype(np.float32)
        t = T[4:].astype(np.float32)
        q = _kf_queries(scene, p, rng, sf, dup=0.5)
        # some points behind the camera / far outside the image / out of the scale-invariance range
        n = len(q["world_pos"])
        q["world_pos"][rng.random(n) < 0.05, 2] *= -1
        q["world_pos"][rng.random(n) < 0.05, 0] += 30
        q["max_dist"][rng.random(n) < 0.05] *= 0.3
        Rs.append(R); ts.append(t); Os.append((-R.T.astype(np.float64) @ t.astype(np.float64)).astype(np.float32)); qs.append(q)
        poff.append(poff[-1] + n)
    cat = lambda k: np.concatenate([q[k] for q in qs])
    got = isInFrustum(ex, cam, poff, np.stack(Rs), np.stack(ts), np.stack(Os), cat("world_pos"), cat("normal"), cat("max_dist"), cat("min_dist"))
    for p in range(P):
        s = slice(poff[p], poff[p + 1])
        r = po.is_in_frustum(Rs[p], ts[p], Os[p], BOUNDS, CAM6, 8, logsf, qs[p]["world_pos"], qs[p]["normal"], qs[p]["max_dist"], qs[p]["min_dist"])
        for k in ("in_view", "level"):
            assert (got[k][s] == r[k]).all(), (p, k)
        for k in ("proj_x", "proj_y", "proj_xr", "view_cos", "depth"):
            assert (got[k][s].view(np.uint32) == r[k].view(np.uint32)).all(), (p, k)
        assert 50 < r["in_view"].sum() < len(r["in_view"])
    m = ORBmatcher(0.8, True)
    match, nm = m.SearchByProjection(ex, cam, [2 * p for p in range(P)], poff, got["proj_x"], got["proj_y"], got["proj_xr"], got["level"],
                                     got["view_cos"], cat("desc"), th=3.0, in_view=got["in_view"])
    for p in range(P):
        a, b = off[2 * p], off[2 * p + 1]
        s = slice(poff[p], poff[p + 1])
        vis = np.nonzero(got["in_view"][s])[0]
        rmatch, rnm = po.search_local(scene["kps"][a:b], scene["desc"][a:b], scene["uR"][a:b], BOUNDS, sf, got["proj_x"][s][vis], got["proj_y"][s][vis],
                                      got["proj_xr"][s][vis], got["level"][s][vis], got["view_cos"][s][vis], qs[p]["desc"][vis], 3.0, 0.8)
        assert rnm == nm[p] and (match[s][vis] == rmatch).all() and (np.delete(match[s], vis) == -1).all(), (p, rnm, nm[p])
        assert rnm > 50

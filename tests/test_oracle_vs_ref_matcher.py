"""CPU tier: the oracle's matcher / grid / stereo / isInFrustum restatements against oracle/_ref part 2 -- member functions of the
REFERENCE cut out of /root/reference/src/{ORBmatcher.cc, Frame.cc, MapPoint.cc, CameraModels/Pinhole.cpp} at build time and compiled
verbatim over skeleton classes (oracle/Makefile `ref2`, oracle/tools/extract_functions.py, oracle/ref_shim/ref_capi2.cpp):

    ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints)      ORBmatcher.cc:45-239
    ORBmatcher::SearchByProjection(Frame& cur, const Frame& last, th, bMono)                           ORBmatcher.cc:1950-2184
    ORBmatcher::DescriptorDistance / ComputeThreeMaxima / RadiusByViewingCos / TH_LOW, TH_HIGH, HISTO_LENGTH
    Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea, Frame::isInFrustum, Frame::ComputeStereoMatches
    MapPoint::PredictScale(const float&, Frame*), Pinhole::project

This pins the control flow that the oracle restated by hand (greedy claims, level rule, stereo gates, rotation histogram, the
median filter of the stereo matcher, the distance-invariance gates) to the reference's own text; the float arithmetic under
Eigen / Sophus expressions is the skeleton's model (see ref_frame_skel.h), the same orders the oracle states."""
import numpy as np
import pytest

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import synth
import pyref

pytestmark = pytest.mark.skipif(po.build_ref2() is None, reason="oracle/_ref part 2 not built and /root/reference absent")

FX, FY, CX, CY, BF, B = 435.2, 435.2, 320.0, 240.0, 47.9, 0.11
CAM6 = np.float32([FX, FY, CX, CY, BF, B])


@pytest.fixture(scope="module")
def scenes():
    out = []
    for (w, h, seed, nf) in [(640, 480, 21, 1200), (752, 480, 22, 1000), (1280, 720, 23, 2000)]:
        l, r, _ = synth.stereo_pair(w, h, seed=seed)
        eL, eR = po.OracleExtractor(nf, 1.2, 8, 20, 7), po.OracleExtractor(nf, 1.2, 8, 20, 7)
        _, kL, dL = eL(l)
        _, kR, dR = eR(r)
        uR, dep, _ = po.stereo_matches(eL, eR, kL, dL, kR, dR, BF, B)
        out.append(dict(w=w, h=h, nf=nf, l=l, r=r, kL=kL, dL=dL, kR=kR, dR=dR, uR=uR, dep=dep, sf=eL.scale_factors,
                        bounds=np.float32([0, w, 0, h])))
    return out


def test_constants_hamming_and_three_maxima():
    c = po.ref2_constants()
    assert (c["TH_LOW"], c["TH_HIGH"], c["HISTO_LENGTH"]) == (50, 100, 30) and c["radius_close"] == 2.5 and c["radius_far"] == 4.0
    rng = np.random.default_rng(0)
    for _ in range(300):
        a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
        if rng.random() < 0.3:
            b = a ^ (rng.random(32) < 0.1).astype(np.uint8)
        assert po.ref2_descriptor_distance(a, b) == po.hamming(a, b) == int(np.unpackbits(a ^ b).sum())
    for t in range(400):
        L = 30
        sizes = rng.integers(0, [3, 12, 60, 200][t % 4], L)
        if t % 5 == 0:
            sizes[rng.integers(0, L, 4)] = sizes.max()        # ties
        if t % 7 == 0:
            sizes[:] = 0
            sizes[rng.integers(0, L)] = 50                    # one dominant bin: second / third dropped by the 10 % rule
        hist = [list(range(s)) for s in sizes]
        assert po.ref2_three_maxima(sizes) == tuple(pyref.three_maxima(hist))


def test_grid_get_features_in_area(scenes):
    rng = np.random.default_rng(1)
    for s in scenes[:2]:
        F = po.RefFrame(s["kL"], s["dL"], s["uR"], s["bounds"], s["sf"], CAM6)
        for _ in range(200):
            x, y = rng.uniform(-20, s["w"] + 20), rng.uniform(-20, s["h"] + 20)
            r = float(rng.choice([1.0, 7.5, 15.0, 40.0]))
            lo, hi = [(-1, -1), (0, 2), (3, -1), (2, 3), (0, 0)][int(rng.integers(0, 5))]
            got = F.features_in_area(np.float32(x), np.float32(y), np.float32(r), lo, hi)
            k = s["kL"]
            inside = (np.abs(k["x"] - np.float32(x)) < np.float32(r)) & (np.abs(k["y"] - np.float32(y)) < np.float32(r))
            if lo > 0 or hi >= 0:
                inside &= k["octave"] >= lo
                if hi >= 0:
                    inside &= k["octave"] <= hi
            # the grid only limits WHICH cells are scanned; every feature inside the window lies in a scanned cell unless it fell
            # outside the grid at assignment -- so the sets agree, in cell-major order
            assert sorted(got.tolist()) == np.nonzero(inside)[0].tolist()


@pytest.mark.parametrize("th,nnratio,far", [(1.0, 0.8, False), (3.0, 0.8, False), (3.0, 0.6, True), (6.0, 0.9, False)])
def test_search_by_projection_local_map_points(scenes, th, nnratio, far):
    rng = np.random.default_rng(int(th * 10) + int(far))
    for s in scenes:
        k, d, dep = s["kL"], s["dL"], s["dep"]
        n = len(k)
        zz = np.where(dep > 0, dep, 5.0).astype(np.float32)
        x = np.concatenate([k["x"] + rng.normal(0, 1.5, n), rng.uniform(10, s["w"] - 10, n)]).astype(np.float32)
        y = np.concatenate([k["y"] + rng.normal(0, 1.5, n), rng.uniform(10, s["h"] - 10, n)]).astype(np.float32)
        zq = np.concatenate([zz, rng.uniform(2, 15, n)]).astype(np.float32)
        xr = (x - np.float32(BF) / zq).astype(np.float32)
        lvl = np.concatenate([k["octave"], rng.integers(0, 8, n)]).astype(np.int32)
        vc = rng.uniform(0.99, 1.0, 2 * n).astype(np.float32)            # both radii of RadiusByViewingCos
        qd = np.concatenate([d, rng.integers(0, 256, (n, 32), dtype=np.uint8)])
        flip = rng.random((2 * n, 32)) < 0.04
        qd = qd ^ np.packbits(flip.reshape(2 * n, 32, 1).repeat(8, 2) & (rng.random((2 * n, 32, 8)) < 0.2), axis=2).reshape(2 * n, 32)
        perm = rng.permutation(2 * n)                                       # vpMapPoints order is arbitrary: greedy claims depend on it
        x, y, xr, lvl, vc, qd, zq = x[perm], y[perm], xr[perm], lvl[perm], vc[perm], qd[perm], zq[perm]
        claimed = (rng.random(n) < 0.15).astype(np.uint8)
        F = po.RefFrame(k, d, s["uR"], s["bounds"], s["sf"], CAM6)
        want, wn = F.search_local(x, y, xr, lvl, vc, qd, th, nnratio, claimed=claimed, trackdepth=zq, far=far, th_far=9.0)
        got, gn = po.search_local(k, d, s["uR"], s["bounds"], s["sf"], x, y, xr, lvl, vc, qd, th, nnratio, claimed=claimed, trackdepth=zq,
                                  far=far, th_far=9.0)
        assert gn == wn and (got == want).all() and wn > n // 3


def _quat(axis, ang):
    a = np.asarray(axis, np.float64)
    a = a / np.linalg.norm(a)
    return np.concatenate([a * np.sin(ang / 2), [np.cos(ang / 2)]])


@pytest.mark.parametrize("th,check,move", [(7.0, True, 0.0), (15.0, True, 0.4), (15.0, False, -0.4), (30.0, True, 0.0)])
def test_search_by_projection_last_frame(scenes, th, check, move):
    rng = np.random.default_rng(int(th) + int(check))
    for s in scenes:
        k, d, dep = s["kL"], s["dL"], s["dep"]
        sel = np.nonzero(dep > 0)[0]
        z = dep[sel]
        fx, fy, cx, cy = (FX, FY, CX, CY) if s["w"] != 1280 else (FX, FY, CX, CY)
        pts = np.stack([(k["x"][sel] - cx) * z / fx, (k["y"][sel] - cy) * z / fy, z], 1).astype(np.float32)
        Tcw = np.concatenate([_quat([0.2, 1, 0.1], 0.004), [0.003, -0.002, 0.001]]).astype(np.float32)
        Tlw = np.concatenate([_quat([0, 1, 0], 0.0), [0.0, 0.0, move]]).astype(np.float32)          # tlc.z vs mb decides forward / backward
        ang = (k["angle"][sel] + rng.normal(0, 4, len(sel))).astype(np.float32) % np.float32(360.0)
        ang[rng.random(len(sel)) < 0.1] = rng.uniform(0, 360, int((rng.random(len(sel)) < 0.1).sum()) or 1)[0]
        obs = (rng.random(len(sel)) < 0.85).astype(np.uint8)
        F = po.RefFrame(k, d, s["uR"], s["bounds"], s["sf"], CAM6)
        want, wn, direction = F.search_last(Tcw, Tlw, pts, k["octave"][sel], ang, d[sel], obs, th, check_ori=check)
        assert direction == (1 if move > B else (2 if -move > B else 0))
        got, gn = po.search_last(k, d, s["uR"], s["bounds"], s["sf"], CAM6, Tcw, direction, pts, k["octave"][sel], ang, d[sel], obs, th, check)
        assert gn == wn and (got == want).all() and wn > 30


def test_is_in_frustum(scenes):
    rng = np.random.default_rng(5)
    s = scenes[0]
    F = po.RefFrame(s["kL"], s["dL"], s["uR"], s["bounds"], s["sf"], CAM6)
    n = 4000
    q = _quat([0.3, 1, -0.2], 0.05)
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]).astype(np.float32)
    t = np.float32([0.05, -0.02, 0.1])
    Ow = (-(R.T.astype(np.float64) @ t.astype(np.float64))).astype(np.float32)
    pc = np.stack([rng.uniform(-8, 8, n), rng.uniform(-6, 6, n), rng.uniform(-1, 14, n)], 1)
    xw = ((pc - t) @ R.astype(np.float64)).astype(np.float32)
    dist = np.linalg.norm(xw - Ow, axis=1).astype(np.float32)
    lv = rng.integers(0, 8, n)
    maxd = (dist * s["sf"][lv] * rng.uniform(0.7, 1.4, n)).astype(np.float32)          # some fail the 1.2f / 0.8f gates, some sit between raw and gate
    mind = (maxd / s["sf"][7]).astype(np.float32)
    nrm = (Ow - xw) / np.maximum(dist[:, None], 1e-6)
    nrm = -(nrm + rng.normal(0, 0.5, (n, 3)))
    nrm = (nrm / np.linalg.norm(nrm, axis=1)[:, None]).astype(np.float32)
    want = F.is_in_frustum(R, t, Ow, xw, nrm, maxd, mind, 0.5)
    got = po.is_in_frustum(R, t, Ow, s["bounds"], CAM6, 8, np.float32(np.log(np.float32(1.2))), xw, nrm, maxd, mind, 0.5)
    vis = want["in_view"].astype(bool)
    assert 300 < vis.sum() < n - 300 and (got["in_view"] == want["in_view"]).all()
    for key in ("proj_x", "proj_y", "proj_xr", "view_cos", "depth"):
        assert (got[key][vis].view(np.uint32) == want[key][vis].view(np.uint32)).all(), key
    assert (got["level"][vis] == want["level"][vis]).all() and len(set(want["level"][vis].tolist())) >= 6


def test_compute_stereo_matches(scenes):
    for s in scenes:
        rL, rR = po.RefExtractor(s["nf"], 1.2, 8, 20, 7), po.RefExtractor(s["nf"], 1.2, 8, 20, 7)
        _, kL, dL = rL(s["l"])
        _, kR, dR = rR(s["r"])
        assert (kL.view(np.uint8) == s["kL"].view(np.uint8)).all() and (dR == s["dR"]).all()
        F = po.RefFrame(kL, dL, None, s["bounds"], s["sf"], CAM6)
        uR, dep = F.stereo_matches(rL, rR, kR, dR)
        assert (uR.view(np.uint32) == s["uR"].view(np.uint32)).all() and (dep.view(np.uint32) == s["dep"].view(np.uint32)).all()
        assert (dep > 0).sum() > 200

"""CPU tier: the oracle functions added after golden_v1 against (i) the committed golden vectors and (ii) independent plain-numpy
restatements where the routine is simple enough to be written twice (isInFrustum, the DBoW2 tree walk and BowVector,
ComputeDistinctiveDescriptors, UpdateNormalAndDepth, the PoseOptimization fixed point)."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import synth, synthetic_vocabulary

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "golden_v2.npz"))
_spec = importlib.util.spec_from_file_location("make_golden_v2", os.path.join(HERE, "golden", "make_golden_v2.py"))
_mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mg)


def test_oracle_reproduces_golden_v2():
    out, _ = _mg.cases()
    for k, v in out.items():
        g = G[k]
        v = np.asarray(v)
        assert g.shape == v.shape, k
        if v.dtype.kind == "f":
            assert (np.ascontiguousarray(v).view(np.uint8) == np.ascontiguousarray(g).view(np.uint8)).all(), k     # bit-exact floats
        else:
            assert (v == g).all(), k
    assert int(G["kf0_n"]) > 100 and int(G["kf2_n"]) > 50 and int(G["kf3_n"]) > 50 and int(G["init_n"]) > 100 and int(G["bowkf_n"]) > 50


def test_is_in_frustum_against_numpy_float32():
    pts, nrm, maxd, mind, T, Ow = (G["in_" + k] for k in ("pts", "nrm", "maxd", "mind", "T", "Ow"))
    f32 = np.float32
    t = T[4:]
    Pc = (pts + t).astype(f32)                                     # R = I: r0.P = x + (0 + 0)
    u = (f32(435.2) * Pc[:, 0] / Pc[:, 2] + f32(320.0)).astype(f32)
    v = (f32(435.2) * Pc[:, 1] / Pc[:, 2] + f32(240.0)).astype(f32)
    PO = (pts - Ow).astype(f32)
    dist = np.sqrt((PO[:, 0] * PO[:, 0] + (PO[:, 1] * PO[:, 1] + PO[:, 2] * PO[:, 2])).astype(f32)).astype(f32)
    vc = ((PO[:, 0] * nrm[:, 0] + (PO[:, 1] * nrm[:, 1] + PO[:, 2] * nrm[:, 2])).astype(f32) / dist).astype(f32)
    vis = (Pc[:, 2] >= 0) & (u >= 0) & (u <= 640) & (v >= 0) & (v <= 480) & (dist >= (f32(0.8) * mind).astype(f32)) & (dist <= (f32(1.2) * maxd).astype(f32)) & (vc >= f32(0.5))   # Get{Min,Max}DistanceInvariance
    assert (G["fr_in_view"].astype(bool) == vis).all()
    assert (G["fr_proj_x"][vis] == u[vis]).all() and (G["fr_proj_y"][vis] == v[vis]).all() and (G["fr_view_cos"][vis] == vc[vis]).all()
    lvl = np.ceil((np.log((maxd / dist).astype(np.float64)) / np.log(1.2))).astype(int).clip(0, 7)   # double log: may differ at exact powers only
    assert (np.abs(G["fr_level"][vis] - lvl[vis]) <= 1).all() and (G["fr_level"][vis] == lvl[vis]).mean() > 0.99
    assert (G["fr_level"][~vis] == -1).all()


def test_bow_transform_against_numpy():
    voc = synthetic_vocabulary(k=7, L=3, seed=11)
    rng = np.random.default_rng(0)
    desc = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    r = po.bow_transform(voc, desc, 2)
    pop = np.array([bin(i).count("1") for i in range(256)])
    acc = {}
    for i, d in enumerate(desc):
        node, level, nid = 0, 0, 0
        while voc["child_offset"][node] != voc["child_offset"][node + 1]:
            ch = voc["child_ids"][voc["child_offset"][node]:voc["child_offset"][node + 1]]
            dd = pop[np.bitwise_xor(voc["node_desc"][ch], d)].sum(1)
            node = int(ch[int(np.argmin(dd))])            # argmin = first minimum
            level += 1
            if level == voc["L"] - 2:
                nid = node
        assert r["word"][i] == voc["node_word"][node] and r["node"][i] == nid and r["weight"][i] == voc["node_weight"][node]
        if voc["node_weight"][node] > 0:
            acc[int(voc["node_word"][node])] = acc.get(int(voc["node_word"][node]), 0.0) + float(voc["node_weight"][node])
    words = sorted(acc)
    norm = 0.0
    for w in words:
        norm += abs(acc[w])
    assert r["bow_word"].tolist() == words
    assert (r["bow_weight"] == np.array([acc[w] / norm for w in words])).all()


def test_distinctive_descriptor_and_normal_depth_against_numpy():
    rng = np.random.default_rng(2)
    pop = np.array([bin(i).count("1") for i in range(256)])
    for n in [1, 2, 3, 6, 17, 40]:
        d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        d[n // 2] = d[0]                                   # a duplicate: ties
        D = pop[np.bitwise_xor(d[:, None, :], d[None, :, :])].sum(2)
        med = np.sort(D, axis=1)[:, int(0.5 * (n - 1))]
        assert po.distinctive_descriptor(d) == int(np.argmin(med))
    assert po.distinctive_descriptor(np.zeros((0, 32), np.uint8)) == -1
    f32 = np.float32
    sf = np.array([f32(1.2) ** i for i in range(8)], f32)
    c, p, rc = rng.normal(size=(5, 3)).astype(f32), rng.normal(size=3).astype(f32) * 4, rng.normal(size=3).astype(f32)
    nrm, mx, mn = po.update_normal_and_depth(c, p, rc, 3, sf)
    acc = np.zeros(3, f32)
    for o in range(5):
        v = (p - c[o]).astype(f32)
        nn = np.sqrt(f32(v[0] * v[0] + f32(v[1] * v[1] + v[2] * v[2])))
        acc = (acc + (v / nn).astype(f32)).astype(f32)
    pc = (p - rc).astype(f32)
    dist = np.sqrt(f32(pc[0] * pc[0] + f32(pc[1] * pc[1] + pc[2] * pc[2])))
    assert (nrm == (acc / f32(5)).astype(f32)).all() and mx == f32(dist * sf[3]) and mn == f32(f32(dist * sf[3]) / sf[7])


def test_pose_optimization_recovers_a_known_pose_and_flags_the_planted_outliers():
    rng = np.random.default_rng(4)
    n = 500
    Xc = np.stack([rng.uniform(-3, 3, n), rng.uniform(-2, 2, n), rng.uniform(2, 12, n)], 1)
    ang, ax = 0.02, np.array([0.3, 1.0, 0.2]) / np.linalg.norm([0.3, 1.0, 0.2])
    q = np.concatenate([ax * np.sin(ang / 2), [np.cos(ang / 2)]])
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    t = np.array([0.04, -0.03, 0.06])
    Xw = (Xc - t) @ R
    u, v = 435.2 * Xc[:, 0] / Xc[:, 2] + 320, 435.2 * Xc[:, 1] / Xc[:, 2] + 240
    obs = np.stack([u, v, u - 47.9 / Xc[:, 2]], 1)
    bad = rng.random(n) < 0.1
    obs[bad, :2] += rng.choice([-1, 1], (int(bad.sum()), 2)) * rng.uniform(15, 40, (int(bad.sum()), 2))
    obs[rng.random(n) < 0.3, 2] = -1
    r = po.pose_optimization([0, 0, 0, 1, 0, 0, 0], Xw, obs, np.ones(n), [435.2, 435.2, 320, 240, 47.9])
    assert np.abs(r["pose"][:4] - q).max() < 1e-6 and np.abs(r["pose"][4:] - t).max() < 1e-5       # noise-free inliers: exact fixed point
    assert (r["outlier"].astype(bool) == bad).all() and r["inliers"] == n - bad.sum() and r["rounds"] == 4
    few = po.pose_optimization([0, 0, 0, 1, 0, 0, 0], Xw[:2], obs[:2], np.ones(2), [435.2, 435.2, 320, 240, 47.9])
    assert few["inliers"] == 0 and few["rounds"] == 0 and (few["pose"] == [0, 0, 0, 1, 0, 0, 0]).all()


@pytest.mark.parametrize("lap", [(0, 0), (0, 1000), (100, 300), (250, 251)])       # std::vector<int> vLappingArea
def test_emission_order_and_lapping_area_against_numpy(lap):
    """ORBextractor::operator() (ORBextractor.cc:1646-1681): level by level, a keypoint whose scaled x lies in [vLappingArea[0],
    vLappingArea[1]] is written from the BACK of the output (stereoIndex--), the others from the front (monoIndex++); the return value
    is monoIndex.  Restated in numpy from the oracle's own per-level lists (level_kps, in the order DistributeOctTree + the per-level
    loop produce them) and compared with what the oracle's operator() returns."""
    img = synth.frame(320, 240, 6)
    ex = po.OracleExtractor(500, 1.2, 8, 20, 7)
    mono, kps, desc = ex(img, lapping=lap)
    lv = [ex.level_kps(l) for l in range(8)]
    n = sum(len(a) for a in lv)
    assert n == len(kps) and n > 100
    out = np.zeros(n, kps.dtype)
    mi, si = 0, n - 1
    for l in range(8):
        sc = np.float32(ex.scale_factors[l])
        for k in lv[l]:
            k = k.copy()
            if l != 0:
                k["x"], k["y"] = np.float32(k["x"]) * sc, np.float32(k["y"]) * sc
            if k["x"] >= np.float32(lap[0]) and k["x"] <= np.float32(lap[1]):
                out[si] = k
                si -= 1
            else:
                out[mi] = k
                mi += 1
    assert mono == mi
    for f in ("x", "y", "octave", "response"):
        assert (out[f] == kps[f]).all(), f

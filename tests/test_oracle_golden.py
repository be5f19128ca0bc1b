"""CPU tier: the oracle against the committed golden vectors and the known-answer tables of SURVEY.md §8."""
import hashlib
import os

import numpy as np

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import synth

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_constructor_tables_known_answers():
    ex = po.OracleExtractor(1000, 1.2, 8, 20, 7)
    assert (ex.umax == G["kat_umax"]).all()
    assert (ex.features_per_level == G["kat_quota_1000"]).all()
    assert (po.OracleExtractor(1200, 1.2, 8, 20, 7).features_per_level == G["kat_quota_1200"]).all()
    ex(synth.frame(640, 480, 1))
    assert [list(ex.level_size(l)) for l in range(8)] == G["kat_levels_640x480"].tolist()
    # scale tables: float32 recurrences of the constructor (ORBextractor.cc:478-500)
    s = np.float32(1.0)
    for l in range(8):
        assert ex.scale_factors[l] == s and ex.inv_scale_factors[l] == np.float32(1.0) / s
        s = np.float32(s * np.float32(1.2))


def test_pattern_table_checksum():
    vals = po.pattern().tolist()
    assert len(vals) == 1024 and vals[:4] == [8, -3, 9, 5] and vals[-4:] == [-1, -6, 0, -11]
    digest = hashlib.sha256(",".join(str(v) for v in vals).encode()).hexdigest()
    assert digest == "88df8ca875cc8db56799edd57bb914edad8acb2d48c202b7a464a575b55dbdb8"


def test_config1_matches_golden():
    img = synth.frame(640, 480, 1)
    ex = po.OracleExtractor(1000, 1.2, 8, 20, 7)
    mono, k, d = ex(img, (0, 0))
    assert mono == int(G["cfg1_mono"]) == len(k)
    assert (k.view(np.uint8) == G["cfg1_kps"].view(np.uint8)).all() and (d == G["cfg1_desc"]).all()
    assert [sha(ex.level_pyramid(l)) for l in range(8)] == G["cfg1_pyr_sha"].tolist()
    assert [sha(ex.level_blurred(l)) for l in range(8)] == G["cfg1_blur_sha"].tolist()
    assert [len(ex.level_cands(l)) for l in range(8)] == G["cfg1_ncand"].tolist()
    # monocular call site {0,1000}: everything is back-filled, return value 0, order reversed (Frame.cc:380)
    mono2, k2, d2 = ex(img, (0, 1000))
    assert mono2 == 0 == int(G["cfg1_mono_ret"])
    assert (k2.view(np.uint8) == G["cfg1_mono_kps"].view(np.uint8)).all() and (d2 == G["cfg1_mono_desc"]).all()
    assert (np.ascontiguousarray(k2[::-1]).view(np.uint8) == k.view(np.uint8)).all() and (d2[::-1] == d).all()


def test_low_texture_and_stereo_match_golden():
    ex = po.OracleExtractor(1200, 1.2, 8, 20, 7)
    _, k, d = ex(synth.frame(640, 480, 2, 6.0, 10))
    assert (k.view(np.uint8) == G["low_kps"].view(np.uint8)).all() and (d == G["low_desc"]).all()
    l, r, _ = synth.stereo_pair(640, 480, seed=40)
    eL, eR = po.OracleExtractor(1200, 1.2, 8, 20, 7), po.OracleExtractor(1200, 1.2, 8, 20, 7)
    _, kL, dL = eL(l)
    _, kR, dR = eR(r)
    uR, dep, kept = po.stereo_matches(eL, eR, kL, dL, kR, dR, 47.9, 0.11)
    assert [sha(kL), sha(dL)] == G["st_kpsL_sha"].tolist()
    assert (uR.view(np.uint32) == G["st_uright"].view(np.uint32)).all() and (dep.view(np.uint32) == G["st_depth"].view(np.uint32)).all()
    assert kept == int(G["st_kept"]) > 100
    m = uR >= 0
    assert ((kL["x"][m] - uR[m]) >= 0).all() and np.allclose(dep[m], 47.9 / np.maximum(kL["x"][m] - uR[m], 0.01), rtol=1e-5)


def test_lba_golden_and_fixed_point():
    pr = {k[7:]: G[k] for k in G.files if k.startswith("lba_in_")}
    r = po.lba(pr["pose"], pr["fixed"], pr["point"], pr["edge_kf"], pr["edge_mp"], pr["obs"], pr["inv_sigma2"], pr["cam5"], 0.0, 10)
    assert np.abs(r["pose"] - G["lba_pose"]).max() < 1e-9 and np.abs(r["point"] - G["lba_point"]).max() < 1e-9
    assert [r["iterations"], r["trials"]] == G["lba_stats"][:2].astype(int).tolist()
    # fixed point: observations generated from the estimate itself => zero residual => zero update
    pose, point = pr["pose"].copy(), pr["point"].copy()
    obs = pr["obs"].copy()
    fx, fy, cx, cy, bf = pr["cam5"]
    for e in range(len(obs)):
        q, t = pose[pr["edge_kf"][e], :4], pose[pr["edge_kf"][e], 4:]
        X = point[pr["edge_mp"][e]]
        x, y, z, w = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        pc = R @ X + t
        u = fx * pc[0] / pc[2] + cx
        obs[e] = [u, fy * pc[1] / pc[2] + cy, -1.0]       # monocular edges: exact double projection
    r0 = po.lba(pose, pr["fixed"], point, pr["edge_kf"], pr["edge_mp"], obs, pr["inv_sigma2"], pr["cam5"], 0.0, 10)
    assert r0["chi2_init"] < 1e-12 and np.abs(r0["pose"] - pose).max() < 1e-9 and np.abs(r0["point"] - point).max() < 1e-9


def test_lba_jacobians_by_finite_differences():
    # g2o's numeric scheme (base_binary_edge.hpp:131-205) is central differences on oplus; the analytic
    # Jacobians of both edge types must agree with it
    import ctypes as C
    L = po.lib()
    L.orc_edge_residual.restype = C.c_int
    rng = np.random.default_rng(3)
    cam = np.array([435.2, 435.2, 320.0, 240.0, 47.9])
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    for trial in range(20):
        q = rng.normal(0, 1, 4); q /= np.linalg.norm(q); q = q if q[3] > 0 else -q
        pose = np.concatenate([q, rng.normal(0, 0.5, 3)])
        # a point in front of the camera: X = R^T (pc - t)
        x, y, z, w = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        pc = np.array([rng.uniform(-2, 2), rng.uniform(-1.5, 1.5), rng.uniform(2, 12)])
        X = R.T @ (pc - pose[4:])
        for D, obs, delta, tol in [(2, np.array([300.0, 200.0, -1.0]), 1e-6, 1e-5), (3, np.array([300.0, 200.0, 280.0]), 1e-3, 2e-3)]:
            A, B = np.zeros(9), np.zeros(18)
            L.orc_edge_jacobians(P(pose), P(X), D, P(cam), P(A), P(B))
            A, B = A.reshape(3, 3)[:D], B.reshape(3, 6)[:D]

            def res(pose_, X_):
                r = np.zeros(3)
                L.orc_edge_residual(P(np.ascontiguousarray(pose_)), P(np.ascontiguousarray(X_)), P(obs), P(cam), P(r))
                return r[:D].copy()
            for j in range(3):
                d = np.zeros(3); d[j] = delta
                num = (res(pose, X + d) - res(pose, X - d)) / (2 * delta)
                assert np.allclose(num, A[:, j], rtol=tol, atol=tol * 50), (D, "A", j)
            for j in range(6):
                up, um = pose.copy(), pose.copy()
                d = np.zeros(6); d[j] = delta
                L.orc_pose_oplus(P(up), P(d)); L.orc_pose_oplus(P(um), P(-d))
                num = (res(up, X) - res(um, X)) / (2 * delta)
                assert np.allclose(num, B[:, j], rtol=tol, atol=tol * 50), (D, "B", j)


def test_descriptor_distance_known_answers():
    rng = np.random.default_rng(2)
    for _ in range(100):
        a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
        want = sum(int(x).bit_count() for x in np.bitwise_xor(a, b))
        assert po.hamming(a, b) == want
    z = np.zeros(32, np.uint8)
    assert po.hamming(z, z) == 0 and po.hamming(z, np.full(32, 255, np.uint8)) == 256

#!/usr/bin/env python3
"""Golden vectors (v2) for the oracle functions added after golden_v1: keyframe-target searches, SearchForInitialization,
SearchByBoW(KF, KF), PoseOptimization, isInFrustum, the DBoW2 transform and the two MapPoint routines.  Produced by the
oracle itself (the reference cannot be built here and ships no vectors): they pin the oracle against drift and give the
CUDA tests a fixture that does not depend on how the oracle was compiled on the GPU box.

  python tests/golden/make_golden_v2.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402
from orb_slam3_detailed_comments_b200 import synth, synthetic_vocabulary  # noqa: E402

W, H = 640, 480
FX, FY, CX, CY, BF, B = 435.2, 435.2, 320.0, 240.0, 47.9, 0.11


def scene():
    l, r, _ = synth.stereo_pair(W, H, seed=900)
    eL, eR = po.OracleExtractor(1000, 1.2, 8, 20, 7), po.OracleExtractor(1000, 1.2, 8, 20, 7)
    _, kL, dL = eL(l)
    _, kR, dR = eR(r)
    uR, dep, _ = po.stereo_matches(eL, eR, kL, dL, kR, dR, BF, B)
    _, k2, d2 = eL(np.roll(l, (3, 1), (1, 0)))
    return eL, kL, dL, uR, dep, k2, d2


def cases():
    """Every case as (name, callable -> dict of arrays); the same callables are used by tests/test_oracle_golden_v2.py."""
    eL, kL, dL, uR, dep, k2, d2 = scene()
    sf = eL.scale_factors
    isg = (1.0 / (sf * sf)).astype(np.float32)
    logsf = po.logf(1.2)
    rng = np.random.default_rng(5)
    sel = np.nonzero(dep > 0)[0]
    sel = np.concatenate([sel, rng.choice(sel, len(sel) // 2)])
    z = dep[sel]
    pts = np.stack([(kL["x"][sel] - CX) * z / FX, (kL["y"][sel] - CY) * z / FY, z], 1).astype(np.float32)
    dist = np.linalg.norm(pts, axis=1).astype(np.float32)
    nrm = (pts / dist[:, None] + rng.normal(0, 0.2, pts.shape)).astype(np.float32)
    nrm = (nrm / np.linalg.norm(nrm, axis=1)[:, None]).astype(np.float32)
    maxd = (dist * sf[kL["octave"][sel]]).astype(np.float32)          # raw mfMaxDistance / mfMinDistance (the gates apply 1.2f / 0.8f)
    mind = (maxd / sf[7]).astype(np.float32)
    zmid = float(np.median(z))
    T = np.array([0, 0.0004, 0, 1, 3 * zmid / FX, 1 * zmid / FY, 0], np.float32)
    T[:4] /= np.linalg.norm(T[:4])
    Ow = -T[4:]                                   # rotation ~ identity
    bounds, cam6 = [0, W, 0, H], [FX, FY, CX, CY, BF, B]
    out = {}
    for v, th, thr in [(0, 3.0, 50.0), (1, 4.0, 50.0), (2, 6.0, 40.0), (3, 10.0, 100.0)]:
        cl = (rng.random(len(k2)) < 0.2).astype(np.uint8) if v >= 2 else None
        m, nm, _ = po.search_keyframe(v, k2, d2, None if v else np.full(len(k2), -1, np.float32), bounds, sf, isg, logsf, cam6, T, Ow, pts, nrm,
                                      maxd, mind, dL[sel], kL["angle"][sel], cl, th, thr)
        out[f"kf{v}_match"], out[f"kf{v}_n"] = m, np.int32(nm)
        out[f"kf{v}_claimed"] = np.zeros(0, np.uint8) if cl is None else cl
    prev = np.stack([kL["x"], kL["y"]], 1).astype(np.float32)
    m, nm = po.search_initialization(kL, dL, prev, k2, d2, bounds, 100, 0.9, True)
    out["init_match"], out["init_n"] = m, np.int32(nm)
    node_of = lambda d: ((d[:, 0].astype(np.int32) >> 3) * 7 + (d[:, 5].astype(np.int32) >> 4) * 3 + (d[:, 17].astype(np.int32) >> 5)) % 97
    q = np.nonzero(rng.random(len(kL)) < 0.8)[0]
    nd = node_of(dL[q])
    order = np.lexsort((q, nd))
    q, nd = q[order], nd[order]
    valid2 = (rng.random(len(k2)) < 0.7).astype(np.uint8)
    m, nm = po.search_bow_kf(k2, d2, node_of(d2), valid2, nd, kL["angle"][q], dL[q], 0.75, True)
    out["bowkf_match"], out["bowkf_n"], out["bowkf_valid2"], out["bowkf_q"] = m, np.int32(nm), valid2, q.astype(np.int32)
    # PoseOptimization: the last frame's stereo points seen from a camera that moved
    obs = np.stack([kL["x"][sel] + 3, kL["y"][sel] + 1, uR[sel] + 3], 1).astype(np.float32) + rng.normal(0, 0.5, (len(sel), 3)).astype(np.float32)
    obs[rng.random(len(sel)) < 0.2, 2] = -1
    r = po.pose_optimization(np.array([0, 0, 0, 1, 0, 0, 0], np.float32), pts, obs, isg[kL["octave"][sel]], np.float32([FX, FY, CX, CY, BF]))
    out["po_pose"], out["po_outlier"], out["po_stats"] = r["pose"], r["outlier"], np.int32([r["inliers"], r["rounds"], r["iterations"], r["trials"]])
    out["po_obs"] = obs
    # the same frame with the stereo projection evaluated as the device kernels do (oracle/lba_oracle.cpp g_stereo_form): what this file held
    # before the oracle was pinned against the reference's own edges (tests/test_oracle_vs_ref_edges.py)
    prev = po.set_stereo_form(1)
    r = po.pose_optimization(np.array([0, 0, 0, 1, 0, 0, 0], np.float32), pts, obs, isg[kL["octave"][sel]], np.float32([FX, FY, CX, CY, BF]))
    po.set_stereo_form(prev)
    out["po_pose_devform"], out["po_outlier_devform"] = r["pose"], r["outlier"]
    out["po_stats_devform"] = np.int32([r["inliers"], r["rounds"], r["iterations"], r["trials"]])
    f = po.is_in_frustum(np.eye(3, dtype=np.float32), T[4:], Ow, bounds, cam6, 8, logsf, pts, nrm, maxd, mind)
    for k, v in f.items():
        out["fr_" + k] = v
    voc = synthetic_vocabulary(k=10, L=3, seed=3)
    bt = po.bow_transform(voc, dL, 2)
    for k, v in bt.items():
        out["bow_" + k] = v
    out["dd_best"] = np.int32([po.distinctive_descriptor(dL[a:a + n]) for a, n in [(0, 1), (3, 2), (10, 7), (40, 32), (100, 61)]])
    n3, mx, mn = po.update_normal_and_depth(pts[:9], pts[20], pts[3], 4, sf)
    out["und"] = np.concatenate([n3, [mx, mn]]).astype(np.float32)
    inputs = dict(sel=sel.astype(np.int32), pts=pts, nrm=nrm, maxd=maxd, mind=mind, T=T, Ow=Ow.astype(np.float32))
    return out, inputs


def main():
    out, inputs = cases()
    out.update({"in_" + k: v for k, v in inputs.items()})
    path = os.path.join(HERE, "golden_v2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()

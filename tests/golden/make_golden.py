#!/usr/bin/env python3
"""Generate the committed golden vectors from the CPU oracle (run in the build container).

The reference ships no tests or golden vectors, and it cannot be compiled here, so these
vectors are produced by the oracle AFTER the oracle's OpenCV-backed stages were checked against cv2 4.13
(tests/test_oracle_vs_cv2.py).  They pin the oracle against drift (CPU tier) and give the CUDA path a
fixture that does not depend on the oracle being rebuilt on the GPU box (GPU tier).

  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402
from orb_slam3_detailed_comments_b200 import synth  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    out = {}
    # config 1: one 640x480 frame, nFeatures = 1000, both call sites (stereo {0,0}, monocular {0,1000})
    img = synth.frame(640, 480, 1)
    ex = po.OracleExtractor(1000, 1.2, 8, 20, 7)
    mono, k, d = ex(img, (0, 0))
    out["cfg1_kps"], out["cfg1_desc"], out["cfg1_mono"] = k, d, np.int32(mono)
    out["cfg1_quota"] = ex.features_per_level.copy()
    out["cfg1_pyr_sha"] = np.array([sha(ex.level_pyramid(l)) for l in range(8)])
    out["cfg1_blur_sha"] = np.array([sha(ex.level_blurred(l)) for l in range(8)])
    out["cfg1_ncand"] = np.array([len(ex.level_cands(l)) for l in range(8)], np.int32)
    mono2, k2, d2 = ex(img, (0, 1000))
    out["cfg1_mono_kps"], out["cfg1_mono_desc"], out["cfg1_mono_ret"] = k2, d2, np.int32(mono2)
    # low texture (20 -> 7 fallback)
    img2 = synth.frame(640, 480, 2, 6.0, 10)
    ex2 = po.OracleExtractor(1200, 1.2, 8, 20, 7)
    _, k, d = ex2(img2)
    out["low_kps"], out["low_desc"] = k, d
    # config 3: stereo pair + matches
    l, r, _ = synth.stereo_pair(640, 480, seed=40)
    eL, eR = po.OracleExtractor(1200, 1.2, 8, 20, 7), po.OracleExtractor(1200, 1.2, 8, 20, 7)
    _, kL, dL = eL(l)
    _, kR, dR = eR(r)
    uR, dep, kept = po.stereo_matches(eL, eR, kL, dL, kR, dR, 47.9, 0.11)
    out["st_kpsL_sha"] = np.array([sha(kL), sha(dL)])
    out["st_uright"], out["st_depth"], out["st_kept"] = uR, dep, np.int32(kept)
    # config 4: local BA (small instance; the full-size one is compared live against the oracle on the GPU)
    pr = synth.lba_problem(n_kf=6, n_fixed=1, n_mp=120, seed=3)
    res = po.lba(pr["pose"], pr["fixed"], pr["point"], pr["edge_kf"], pr["edge_mp"], pr["obs"], pr["inv_sigma2"], pr["cam5"], 0.0, 10)
    for key in ("pose", "fixed", "point", "edge_kf", "edge_mp", "obs", "inv_sigma2", "cam5"):
        out["lba_in_" + key] = pr[key]
    out["lba_pose"], out["lba_point"] = res["pose"], res["point"]
    out["lba_stats"] = np.array([res["iterations"], res["trials"], res["chi2_init"], res["chi2"], res["lambda_"]])
    # the same problem with the stereo projection evaluated as the device kernels do (oracle/lba_oracle.cpp g_stereo_form): the vectors
    # this file held before the oracle was pinned against the reference's own edges (tests/test_oracle_vs_ref_edges.py)
    prev = po.set_stereo_form(1)
    res = po.lba(pr["pose"], pr["fixed"], pr["point"], pr["edge_kf"], pr["edge_mp"], pr["obs"], pr["inv_sigma2"], pr["cam5"], 0.0, 10)
    po.set_stereo_form(prev)
    out["lba_pose_devform"], out["lba_point_devform"] = res["pose"], res["point"]
    # known-answer values derived independently of the oracle
    out["kat_umax"] = np.array([15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3], np.int32)
    out["kat_quota_1000"] = np.array([217, 181, 151, 126, 105, 87, 73, 60], np.int32)
    out["kat_quota_1200"] = np.array([261, 217, 181, 151, 126, 105, 87, 72], np.int32)
    out["kat_levels_640x480"] = np.array([[640, 480], [533, 400], [444, 333], [370, 278], [309, 231], [257, 193], [214, 161], [179, 134]], np.int32)
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote", os.path.join(HERE, "golden_v1.npz"), os.path.getsize(os.path.join(HERE, "golden_v1.npz")), "bytes")


if __name__ == "__main__":
    main()

"""CPU tier: csrc/stereo_core.cuh -- the per-left-keypoint body of k_stereo_match_v1 (row-bucketed candidates, SAD by the same thread)
-- compiled for the host and run against the oracle's Frame::ComputeStereoMatches, after applying the reference's median-SAD filter
(Frame.cc:1338-1357) to the emulated pre-filter output."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import synth

BF, B = 47.9, 0.11


@pytest.fixture(scope="module")
def emul():
    from _emul import build_and_load
    L = build_and_load()
    L.emul_stereo_v1.restype = None
    L.emul_stereo_v1.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def run_v1(emul, eL, eR, kL, dL, kR, dR, H, bf, b):
    nl = eL.nlevels
    pl = [np.ascontiguousarray(eL.level_pyramid(l)) for l in range(nl)]
    pr = [np.ascontiguousarray(eR.level_pyramid(l)) for l in range(nl)]
    ptr = lambda arrs: (C.c_void_p * nl)(*[a.ctypes.data for a in arrs])
    wl, wr = np.array([a.shape[1] for a in pl], np.int32), np.array([a.shape[1] for a in pr], np.int32)
    kL, kR = np.ascontiguousarray(kL), np.ascontiguousarray(kR)
    dL, dR = np.ascontiguousarray(dL), np.ascontiguousarray(dR)
    u, d, sad = np.zeros(len(kL), np.float32), np.zeros(len(kL), np.float32), np.zeros(len(kL), np.int32)
    emul.emul_stereo_v1(kL.ctypes.data, dL.ctypes.data, len(kL), kR.ctypes.data, dR.ctypes.data, len(kR), nl, ptr(pl), ptr(pr),
                        wl.ctypes.data, wr.ctypes.data, eL.scale_factors.ctypes.data, eL.inv_scale_factors.ctypes.data, H, bf, b,
                        u.ctypes.data, d.ctypes.data, sad.ctypes.data)
    ok = sad >= 0                                        # Frame.cc:1338-1357: sort the SADs, median = element size / 2, cut at 1.5 * 1.4 * median
    if ok.any():
        med = np.sort(sad[ok])[ok.sum() // 2]
        th = np.float32(np.float32(1.5) * np.float32(1.4)) * np.float32(med)
        drop = ok & ~(sad.astype(np.float32) < th)
        u[drop], d[drop] = -1.0, -1.0
    return u, d


@pytest.mark.parametrize("w,h,seed,nf", [(640, 480, 1, 1200), (752, 480, 3, 1200), (320, 240, 4, 500), (1280, 720, 5, 2000)])
def test_v1_matches_oracle(emul, w, h, seed, nf):
    l, r, _ = synth.stereo_pair(w, h, seed)
    eL, eR = po.OracleExtractor(nf, 1.2, 8, 20, 7), po.OracleExtractor(nf, 1.2, 8, 20, 7)
    _, kL, dL = eL(l)
    _, kR, dR = eR(r)
    uR, dep, kept = po.stereo_matches(eL, eR, kL, dL, kR, dR, BF, B)
    u, d = run_v1(emul, eL, eR, kL, dL, kR, dR, h, BF, B)
    assert (u.view(np.uint32) == uR.view(np.uint32)).all() and (d.view(np.uint32) == dep.view(np.uint32)).all()
    assert (dep > 0).sum() > 100


def test_v1_edge_cases(emul):
    l, r, _ = synth.stereo_pair(320, 240, 9)
    eL, eR = po.OracleExtractor(500, 1.2, 8, 20, 7), po.OracleExtractor(500, 1.2, 8, 20, 7)
    _, kL, dL = eL(l)
    _, kR, dR = eR(r)
    # no right keypoints at all; and a right image that is flat (no keypoints) vs the same left
    u, d = run_v1(emul, eL, eR, kL, dL, kR[:0], dR[:0], 240, BF, B)
    assert (u == -1).all() and (d == -1).all()
    # duplicated right keypoints (equal distances): the lowest right index must win, whatever the bucket order
    kR2, dR2 = np.concatenate([kR, kR]), np.concatenate([dR, dR])
    uR, dep, _ = po.stereo_matches(eL, eR, kL, dL, kR2, dR2, BF, B)
    u, d = run_v1(emul, eL, eR, kL, dL, kR2, dR2, 240, BF, B)
    assert (u.view(np.uint32) == uR.view(np.uint32)).all() and (d.view(np.uint32) == dep.view(np.uint32)).all()


def test_v1_on_synthetic_right_keypoint_sets(emul):
    """Right keypoint sets built to stress the candidate gates rather than coming from an extraction: every left keypoint gets several
    right twins at random disparities, rows up to +-9 px away, octaves +-2 and a few flipped descriptor bits (many equal distances:
    the lowest right index must win), plus unrelated keypoints; coordinates at the image border included."""
    l, r, _ = synth.stereo_pair(640, 480, 12)
    eL, eR = po.OracleExtractor(800, 1.2, 8, 20, 7), po.OracleExtractor(800, 1.2, 8, 20, 7)
    _, kL, dL = eL(l)
    eR(r)                                                   # only its pyramid is used
    for seed in range(6):
        rng = np.random.default_rng(seed)
        reps = int(rng.integers(1, 4))
        kR = np.concatenate([kL] * reps).copy()
        dR = np.concatenate([dL] * reps).copy()
        kR["x"] = np.clip(kR["x"] - rng.uniform(-5, 70, len(kR)).astype(np.float32), 0, 639)
        kR["y"] = np.clip(kR["y"] + rng.integers(-9, 10, len(kR)).astype(np.float32), 0, 479)
        kR["octave"] = np.clip(kR["octave"] + rng.integers(-2, 3, len(kR)), 0, 7)
        flips = rng.integers(0, 256, (len(kR), 3))
        for c in range(3):
            m = rng.random(len(kR)) < 0.5
            dR[m, flips[m, c] // 8] ^= (1 << (flips[m, c] % 8)).astype(np.uint8)
        perm = rng.permutation(len(kR))
        kR, dR = np.ascontiguousarray(kR[perm]), np.ascontiguousarray(dR[perm])
        uR, dep, _ = po.stereo_matches(eL, eR, kL, dL, kR, dR, BF, B)
        u, d = run_v1(emul, eL, eR, kL, dL, kR, dR, 480, BF, B)
        assert (u.view(np.uint32) == uR.view(np.uint32)).all() and (d.view(np.uint32) == dep.view(np.uint32)).all(), seed

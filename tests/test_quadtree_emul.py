"""CPU tier: the kernel's sorted-path quadtree (csrc/quadtree_core.cuh compiled for the host) against the
oracle's std::list / std::sort restatement of DistributeOctTree, on real FAST candidates and on adversarial
point sets (ties in score, clusters, near-empty levels, two root nodes)."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul():
    d = os.path.join(HERE, "host_emul")
    so = os.path.join(d, "libemul.so")
    csrc = os.path.join(HERE, "..", "orb_slam3_detailed_comments_b200", "csrc")
    srcs = [os.path.join(d, "emul.cpp"), os.path.join(csrc, "devmath.cuh"), os.path.join(csrc, "quadtree_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-o", so, srcs[0], "-lpthread"])
    L = C.CDLL(so)
    L.emul_distribute.restype = C.c_int
    L.emul_distribute.argtypes = [C.c_void_p, C.c_int] + [C.c_int] * 6 + [C.c_void_p, C.c_int]
    return L


def cell_geom(w, h):
    width, height = np.float32(w - 32), np.float32(h - 32)
    nCols, nRows = int(width / np.float32(35)), int(height / np.float32(35))
    return int(math.ceil(width / nCols)), int(math.ceil(height / nRows)), nCols


def run(emul, c, W, H, N):
    wC, hC, nC = cell_geom(W, H)
    out = np.zeros((N + 32, 3), np.int32)
    c = np.ascontiguousarray(c, np.int32)
    S = emul.emul_distribute(c.ctypes.data, len(c), W - 32, H - 32, N, wC, hC, nC, out.ctypes.data, len(out))
    assert S >= 0, "workspace capacity (N + 20 nodes) exceeded"
    return out[:S]


@pytest.mark.parametrize("w,h,seed,sigma,nrect,nf", [(640, 480, 1, 1.5, 60, 1200), (640, 480, 2, 6.0, 10, 1200),
                                                      (752, 480, 3, 1.5, 60, 1200), (1280, 720, 5, 1.5, 60, 2000),
                                                      (640, 480, 6, 1.5, 60, 300), (640, 480, 7, 1.5, 60, 5000)])
def test_real_candidates(emul, w, h, seed, sigma, nrect, nf):
    ex = po.OracleExtractor(nf, 1.2, 8, 20, 7)
    ex(synth.frame(w, h, seed, sigma, nrect))
    for l in range(8):
        c = ex.level_cands(l)
        W, H = ex.level_size(l)
        N = int(ex.features_per_level[l])
        ref = po.distribute(c, 16, W - 16, 16, H - 16, N)
        got = run(emul, c, W, H, N)
        assert len(got) == len(ref) and (got == c[ref]).all(), (seed, l)
        assert len(got) <= N + 3 or len(got) <= 16     # the capacity argument of quadtree_core.cuh


def _grid_order(pts, W, H):
    # candidates must arrive in the reference order: cell row, cell column, then row-major inside the cell
    wC, hC, nC = cell_geom(W, H)
    p64 = pts.astype(np.int64)
    key = ((p64[:, 1] - 3) // hC * nC + (p64[:, 0] - 3) // wC) * (1 << 24) + p64[:, 1] * 4096 + p64[:, 0]
    return pts[np.argsort(key, kind="stable")]


@pytest.mark.parametrize("case", range(12))
def test_adversarial_point_sets(emul, case):
    rng = np.random.default_rng(100 + case)
    W, H = [(640, 480), (752, 480), (309, 231), (1280, 720)][case % 4]
    rw, rh = W - 32, H - 32
    n = [5, 40, 300, 3000, 9000, 1][case % 6]
    if case % 3 == 0:      # uniform
        xy = np.stack([rng.integers(3, rw - 3, n), rng.integers(3, rh - 3, n)], 1)
    elif case % 3 == 1:    # a few tight clusters
        ctr = np.stack([rng.integers(20, rw - 20, 4), rng.integers(20, rh - 20, 4)], 1)
        xy = ctr[rng.integers(0, 4, n)] + rng.integers(-15, 16, (n, 2))
    else:                  # a line along the root boundary + a corner blob
        xy = np.concatenate([np.stack([np.full(n // 2, rw // 2), rng.integers(3, rh - 3, n // 2)], 1),
                             np.stack([rng.integers(3, 40, n - n // 2), rng.integers(3, 40, n - n // 2)], 1)])
    xy = np.clip(xy, 3, [rw - 4, rh - 4])
    xy = np.unique(xy, axis=0)                       # FAST keypoints are distinct pixels
    score = rng.integers(7, 12 if case % 2 else 200, len(xy))     # few distinct scores => many ties
    c = _grid_order(np.concatenate([xy, score[:, None]], 1).astype(np.int32), W, H)
    for N in [1, 7, 60, 261]:
        ref = po.distribute(c, 16, W - 16, 16, H - 16, N)
        got = run(emul, c, W, H, N)
        assert len(got) == len(ref) and (got == c[ref]).all(), (case, N)

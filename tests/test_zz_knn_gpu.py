"""GPU tier: orbm_hamming_knn2 (csrc/knn.cu, K9 brute-force Hamming 2-NN) against the oracle, which
tests/test_knn_cpu.py pins to cv2.BFMatcher.  Written after round 1's last GPU run: opt-in until its first green device run."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import ORBextractor, knnMatch2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tma", ["0", "1"])
def test_batched_sets_match_oracle(monkeypatch, tma):
    monkeypatch.setenv("ORB_KNN_TMA", tma)      # 1: train tiles by cp.async.bulk + mbarrier (read at every call)
    rng = np.random.default_rng(7)
    shapes = [(300, 500, 2), (257, 129, 8), (40, 1, 8), (5, 2, 1), (1200, 1200, 8), (3, 0, 8), (0, 7, 8), (513, 385, 3)]
    qs = [rng.integers(0, 1 << b, (nq, 32)).astype(np.uint8) for nq, nt, b in shapes]
    ts = [rng.integers(0, 1 << b, (nt, 32)).astype(np.uint8) for nq, nt, b in shapes]
    ts[0][[3, 7, 499]] = qs[0][0]
    ex = ORBextractor(500, 1.2, 8, 20, 7, max_width=320, max_height=240, max_batch=1)
    got = knnMatch2(ex, qs, ts)
    for (gi, gd), q, t in zip(got, qs, ts):
        ri, rd = po.hamming_knn2(q, t)
        assert (gi == ri).all() and (gd == rd).all()
    ex.close()

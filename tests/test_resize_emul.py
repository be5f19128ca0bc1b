"""CPU tier: csrc/resize_core.cuh -- the per-thread body of k_resize_v3 (source rows by aligned word loads, PRMT tap pairs, IDP.2A) --
compiled for the host and run against the oracle's cv::resize(INTER_LINEAR) model on every level transition of the usual geometries,
with the source placed at every byte alignment, with pitch padding, and with the readable bytes ending exactly at the last pixel
(the byte-load path of the threads whose word loads would cross the end of a caller-owned image)."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as po


@pytest.fixture(scope="module")
def emul():
    from _emul import build_and_load
    L = build_and_load()
    L.emul_resize_v3.restype = None
    L.emul_resize_v3.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_int, C.c_int, C.c_int]
    return L


def _run(emul, img, dw, dh, align, pad, tight):
    sh, sw = img.shape
    spitch = sw + pad
    buf = np.full(spitch * sh + 64, 0xA5, np.uint8)
    view = buf[align:align + spitch * sh].reshape(sh, spitch)
    view[:, :sw] = img
    avail = (sh - 1) * spitch + sw if tight else spitch * sh + 32
    dpitch = (dw + 15) // 16 * 16
    out = np.full((dh, dpitch), 0xEE, np.uint8)
    emul.emul_resize_v3(buf.ctypes.data + align, sw, sh, spitch, avail, out.ctypes.data, dw, dh, dpitch)
    return out


@pytest.mark.parametrize("w,h,scale", [(640, 480, 1.2), (752, 480, 1.2), (1280, 720, 1.2), (500, 377, 1.2), (640, 480, 1.5), (320, 240, 1.9)])
def test_every_level_transition_matches_the_oracle(emul, w, h, scale):
    rng = np.random.default_rng(w + h)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    sw, sh = w, h
    for l in range(1, 8):
        s = np.float32(1.0)
        for _ in range(l):
            s = np.float32(s * np.float32(scale))
        inv = np.float32(1.0) / s
        dw, dh = int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))   # ORBextractor.cc:1691-1692 cvRound(float)
        if dw < 8 or dh < 8:
            break
        want = po.resize(img, dw, dh)
        for align, pad, tight in [(0, 0, False), (1, 0, True), (2, 16, False), (3, 5, True)]:
            got = _run(emul, img, dw, dh, align, pad, tight)
            assert (got[:, :dw] == want).all(), (l, align, pad, tight)
            assert (got[:, dw:(dw + 3) // 4 * 4] == 0).all()            # the partial last word is zero past the width, as k_resize writes it
        img, sw, sh = want, dw, dh

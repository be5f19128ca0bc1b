"""Seeded synthetic inputs for BASELINE.json's configs.

numpy only: the same generators run in the build container, on the GPU box, in tests/ and bench.py.
"""
import numpy as np


def _gauss_kernel(sigma):
    r = max(1, int(np.ceil(3 * sigma)))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum()


def _smooth(a, sigma):
    k = _gauss_kernel(sigma)
    r = len(k) // 2
    p = np.pad(a, ((0, 0), (r, r)), mode="reflect")
    a = sum(k[i] * p[:, i:i + a.shape[1]] for i in range(len(k)))
    p = np.pad(a, ((r, r), (0, 0)), mode="reflect")
    return sum(k[i] * p[i:i + a.shape[0], :] for i in range(len(k)))


def frame(w=640, h=480, seed=1, sigma=1.5, nrect=60):
    """Smoothed uniform noise normalised to 0..255 plus `nrect` filled rectangles (config 1).

    sigma=1.5/nrect=60 is the texture-rich frame (every level meets its quota); sigma=6/nrect=10 is the
    low-texture variant that exercises the 20->7 FAST fallback.
    """
    rng = np.random.default_rng(seed)
    a = _smooth(rng.random((h, w)), sigma)
    a = (a - a.min()) / (a.max() - a.min()) * 255.0
    for _ in range(nrect):
        x0 = int(rng.integers(0, w - 8))
        y0 = int(rng.integers(0, h - 8))
        rw = int(rng.integers(8, max(9, w // 6)))
        rh = int(rng.integers(8, max(9, h // 6)))
        a[y0:y0 + rh, x0:x0 + rw] = float(rng.integers(0, 256))
    return np.clip(np.rint(a), 0, 255).astype(np.uint8)


def stereo_pair(w=640, h=480, seed=1, dmin=2.0, dmax=60.0, sigma=1.5, nrect=60):
    """Left frame + right frame = left warped by a smooth disparity field d(x,y) in [dmin,dmax] px
    (right(x - d) = left(x), rectified, zero distortion).  Returns (left, right, disparity)."""
    rng = np.random.default_rng(seed + 7919)
    pad = int(np.ceil(dmax)) + 2
    big = frame(w + pad, h, seed, sigma, nrect).astype(np.float64)
    left = big[:, :w]
    # disparity as a function of the RIGHT pixel: smooth in x, piecewise over rows
    base = _smooth(rng.random((h, w)), 25.0)
    base = (base - base.min()) / (base.max() - base.min())
    disp = dmin + (dmax - dmin) * base
    xs = np.arange(w)[None, :] + disp  # sample position in the left/big image
    x0 = np.floor(xs).astype(np.int64)
    fx = xs - x0
    x0 = np.clip(x0, 0, w + pad - 2)
    rows = np.arange(h)[:, None]
    right = (1 - fx) * big[rows, x0] + fx * big[rows, x0 + 1]
    return (np.clip(np.rint(left), 0, 255).astype(np.uint8),
            np.clip(np.rint(right), 0, 255).astype(np.uint8), disp.astype(np.float32))


def frame_batch(n, w=640, h=480, seed=0, sigma=1.5, nrect=60):
    return np.stack([frame(w, h, seed + i, sigma, nrect) for i in range(n)])

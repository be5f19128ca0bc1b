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


# ------------------------------------------------------------------------------------------------
# Local bundle adjustment problem (BASELINE.json config 4)
# ------------------------------------------------------------------------------------------------
def _quat_from_rotvec(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.array([0, 0, 0, 1.0])
    a = w / th
    return np.array([*(a * np.sin(th / 2)), np.cos(th / 2)])


def _quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def _quat_rot(q, v):
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return v @ R.T


def lba_problem(n_kf=20, n_fixed=2, n_mp=3000, seed=0, mono_frac=0.1, outlier_frac=0.03, fx=435.2, fy=435.2, cx=320.0,
                cy=240.0, bf=47.9, width=640, height=480, scale_factor=1.2, n_levels=8):
    """20 keyframes (first `n_fixed` fixed) on a forward-moving trajectory, `n_mp` points each seen by 4-8
    keyframes, stereo edges with 10 % monocular (ur = -1), pixel noise sigma = scale[octave], 3 % gross
    outliers (20 px), poses perturbed by 1 cm / 0.2 deg, points by 2 cm.  Returns a dict of flat arrays
    (the lba_problem layout of include/orbslam3_b200.h)."""
    rng = np.random.default_rng(seed)
    # ground-truth camera-from-world poses: translation along +x/+z, slow yaw
    Tq, Tt = [], []
    for k in range(n_kf):
        q = _quat_from_rotvec(np.array([0.0, np.deg2rad(0.2 * k), 0.0]))
        c = np.array([0.02 * 5 * k, 0.0, 0.01 * 5 * k])          # camera centre in the world
        Tq.append(q)
        Tt.append(-_quat_rot(q, c))
    Tq, Tt = np.array(Tq), np.array(Tt)
    scales = scale_factor ** np.arange(n_levels)
    pts, ekf, emp, obs, inv_s2 = [], [], [], [], []
    tries = 0
    while len(pts) < n_mp and tries < 50 * n_mp:
        tries += 1
        k0 = int(rng.integers(0, n_kf))
        z = rng.uniform(2.0, 15.0)
        u, v = rng.uniform(30, width - 30), rng.uniform(30, height - 30)
        xc = np.array([(u - cx) * z / fx, (v - cy) * z / fy, z])
        # world point: X = R^T (xc - t)
        qinv = Tq[k0] * np.array([-1, -1, -1, 1])
        X = _quat_rot(qinv, xc - Tt[k0])
        nobs = int(rng.integers(4, 9))
        cand = np.arange(max(0, k0 - 6), min(n_kf, k0 + 7))
        rng.shuffle(cand)
        seen = []
        for k in cand:
            pc = _quat_rot(Tq[k], X) + Tt[k]
            if pc[2] <= 0.5:
                continue
            uu, vv = fx * pc[0] / pc[2] + cx, fy * pc[1] / pc[2] + cy
            ur = uu - bf / pc[2]
            if 0 < uu < width and 0 < vv < height and ur > 0:
                seen.append((int(k), uu, vv, ur))
            if len(seen) == nobs:
                break
        if len(seen) < 2:
            continue
        idx = len(pts)
        pts.append(X)
        for (k, uu, vv, ur) in sorted(seen):
            octave = int(rng.integers(0, n_levels))
            sig = scales[octave]
            n = rng.normal(0, sig, 3)
            if rng.random() < outlier_frac:
                n += rng.choice([-1, 1], 3) * 20.0
            mono = rng.random() < mono_frac
            ekf.append(k)
            emp.append(idx)
            obs.append([np.float32(uu + n[0]), np.float32(vv + n[1]), -1.0 if mono else np.float32(ur + n[2])])
            inv_s2.append(np.float32(1.0 / np.float32(sig * sig)))
    pts = np.array(pts)
    # perturbed initial estimates (float32 like the map, promoted to double)
    pose = np.zeros((n_kf, 7))
    for k in range(n_kf):
        if k < n_fixed:
            q, t = Tq[k], Tt[k]
        else:
            dq = _quat_from_rotvec(rng.normal(0, np.deg2rad(0.2), 3))
            q = _quat_mul(dq, Tq[k])
            t = Tt[k] + rng.normal(0, 0.01, 3)
        q = q / np.linalg.norm(q)
        if q[3] < 0:
            q = -q
        pose[k, :4], pose[k, 4:] = q, t
    point = (pts + rng.normal(0, 0.02, pts.shape)).astype(np.float32).astype(np.float64)
    fixed = np.zeros(n_kf, np.uint8)
    fixed[:n_fixed] = 1
    return dict(pose=pose.astype(np.float32).astype(np.float64), fixed=fixed, point=point,
                edge_kf=np.array(ekf, np.int32), edge_mp=np.array(emp, np.int32), obs=np.array(obs, np.float64),
                inv_sigma2=np.array(inv_s2, np.float64), cam5=np.array([np.float32(fx), np.float32(fy), np.float32(cx),
                                                                         np.float32(cy), np.float32(bf)], np.float64),
                pose_gt=np.concatenate([Tq, Tt], 1), point_gt=pts)

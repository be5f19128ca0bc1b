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


def _expm(w):
    w = np.asarray(w, float)
    t = np.linalg.norm(w)
    W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if t < 1e-9:
        return np.eye(3) + W
    return np.eye(3) + np.sin(t) / t * W + (1 - np.cos(t)) / t ** 2 * W @ W


def _hat32(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], np.float32)


def preintegrate(acc, gyr, dt, bias6=None, ng=1.7e-4, na=2.0e-3, ngw=1.9393e-5, naw=3.0e-3, freq=200.0):
    """IMU::Preintegrated::IntegrateNewMeasurement over a run of samples (src/ImuTypes.cc:247-320, IntegratedRotation :125-152,
    Calib::Set :565-580 with Tracking's sqrt(freq) scaling of the EuRoC.yaml noise densities), in float32 like the reference.
    acc / gyr: [n][3]; bias6 = bax bay baz bwx bwy bwz.  Returns the members a liba_link needs (dR dV dP JRg JVg JVa JPg JPa, C
    15 x 15, dT, bias).  NormalizeRotation (an SVD re-orthonormalisation) is replaced by nothing: n <= a few hundred steps."""
    f = np.float32
    b = np.zeros(6, f) if bias6 is None else np.asarray(bias6, f)
    sf = np.sqrt(freq)
    Nga = np.diag(np.array([(ng * sf) ** 2] * 3 + [(na * sf) ** 2] * 3, f))
    NgaWalk = np.diag(np.array([(ngw / sf) ** 2] * 3 + [(naw / sf) ** 2] * 3, f))
    dR, dV, dP = np.eye(3, dtype=f), np.zeros(3, f), np.zeros(3, f)
    JRg, JVg, JVa, JPg, JPa = (np.zeros((3, 3), f) for _ in range(5))
    C = np.zeros((15, 15), f)
    dT = f(0)
    dt = f(dt)
    for a_m, w_m in zip(np.asarray(acc, f), np.asarray(gyr, f)):
        A, B = np.eye(9, dtype=f), np.zeros((9, 6), f)
        a = a_m - b[:3]
        dP = dP + dV * dt + f(0.5) * (dR @ a) * dt * dt
        dV = dV + (dR @ a) * dt
        Wacc = _hat32(a)
        A[3:6, 0:3] = -dR * dt @ Wacc
        A[6:9, 0:3] = f(-0.5) * dR * dt * dt @ Wacc
        A[6:9, 3:6] = np.eye(3, dtype=f) * dt
        B[3:6, 3:6] = dR * dt
        B[6:9, 3:6] = f(0.5) * dR * dt * dt
        JPa = JPa + JVa * dt - f(0.5) * dR * dt * dt
        JPg = JPg + JVg * dt - f(0.5) * dR * dt * dt @ Wacc @ JRg
        JVa = JVa - dR * dt
        JVg = JVg - dR * dt @ Wacc @ JRg
        v = (w_m - b[3:]) * dt
        d2 = f(v @ v)
        d = np.sqrt(d2)
        W = _hat32(v)
        if d < 1e-4:
            dRi, rJ = np.eye(3, dtype=f) + W, np.eye(3, dtype=f)
        else:
            dRi = np.eye(3, dtype=f) + W * (np.sin(d) / d) + W @ W * ((f(1) - np.cos(d)) / d2)
            rJ = np.eye(3, dtype=f) - W * ((f(1) - np.cos(d)) / d2) + W @ W * ((d - np.sin(d)) / (d2 * d))
        dR = (dR @ dRi).astype(f)
        A[0:3, 0:3] = dRi.T
        B[0:3, 0:3] = rJ * dt
        C[0:9, 0:9] = A @ C[0:9, 0:9] @ A.T + B @ Nga @ B.T
        C[9:15, 9:15] += NgaWalk
        JRg = dRi.T @ JRg - rJ * dt
        dT = dT + dt
    return dict(dR=dR, dV=dV, dP=dP, JRg=JRg, JVg=JVg, JVa=JVa, JPg=JPg, JPa=JPa, C=C, dT=float(dT), bias=b)


def inertial_window(n_opt=10, n_cov_fixed=3, n_mp=2000, seed=0, noise=0.5, perturb=True, dt=0.25, keep_pre=False):
    """A synthetic Optimizer::LocalInertialBA window (Optimizer.cc:2217-2340): a temporal chain of n_opt optimisable keyframes
    plus the fixed keyframe before them (index 0), joined by inertial links preintegrated from synthetic 200 Hz IMU samples
    (preintegrate(), information matrices through liba_link_information), then n_cov_fixed fixed covisible keyframes without links; stereo / mono observations of n_mp points.  Returns the dict orb_slam3_detailed_comments_b200.InertialOptimizer takes
    (state [nKF][21] = Rwb twb v bg ba, fixed, point, edge_kf, edge_mp, obs, inv_sigma2, links, Tcb, cam5)."""
    from ._native import LIBA_LINK
    from .optimizer import link_information
    rng = np.random.default_rng(seed)
    G = np.array([0, 0, -float(np.float32(9.81))])
    n_chain = n_opt + 1
    states, links, pres = [], np.zeros(n_chain - 1, LIBA_LINK), []
    R, p, v = _expm(rng.normal(0, 0.2, 3)), rng.normal(0, 1, 3), rng.normal(0, 0.3, 3)
    for k in range(n_chain):
        states.append(np.concatenate([R.reshape(-1), p, v, np.zeros(6)]))
        if k == n_chain - 1:
            break
        # 200 Hz IMU samples over the interval: slowly varying body rates and a specific force near -g (a hovering platform),
        # preintegrated exactly as the reference does; the next keyframe state is the one these deltas predict
        n_s = int(round(dt * 200))
        tt = np.linspace(0, 1, n_s)[:, None]
        w0, w1 = rng.normal(0, 0.15, 3), rng.normal(0, 0.15, 3)
        a0, a1 = rng.normal(0, 0.5, 3), rng.normal(0, 0.5, 3)
        gyr = w0 * (1 - tt) + w1 * tt
        acc = a0 * (1 - tt) + a1 * tt - (R.T @ G) * 0.97
        pre = preintegrate(acc, gyr, dt / n_s)
        pres.append(pre)
        dT = pre["dT"]
        dR = pre["dR"].astype(np.float64)
        v2 = v + G * dT + R @ pre["dV"].astype(np.float64)
        p2 = p + v * dT + 0.5 * G * dT * dT + R @ pre["dP"].astype(np.float64)
        lk = links[k]
        lk["k1"], lk["k2"], lk["robust"], lk["dt"] = k, k + 1, int(k == 0), dT
        for name in ("dR", "dV", "dP", "JRg", "JVg", "JVa", "JPg", "JPa"):
            lk[name] = pre[name].reshape(-1)
        lk["bias"] = pre["bias"]
        info, ig, ia = link_information(pre["C"], oldest=(k == 0))
        lk["info"], lk["infoG"], lk["infoA"] = info.reshape(-1), ig.reshape(-1), ia.reshape(-1)
        R, p, v = R @ dR, p2, v2
    for _ in range(n_cov_fixed):       # covisible fixed keyframes: poses near the chain, no inertial links
        s = states[rng.integers(0, n_chain)].copy()
        s[:9] = (s[:9].reshape(3, 3) @ _expm(rng.normal(0, 0.05, 3))).reshape(-1)
        s[9:12] += rng.normal(0, 0.3, 3)
        states.append(s)
    states = np.array(states)
    n_kf = len(states)
    FX, FY, CX, CY, BF = 435.2, 435.2, 320.0, 240.0, 47.9
    Rcb, tcb = _expm([0.01, -0.02, 0.015]), np.array([0.05, -0.01, 0.02])
    Rwb = states[:, :9].reshape(n_kf, 3, 3)
    twb = states[:, 9:12]
    k0 = rng.integers(0, n_kf, n_mp)
    Xc0 = np.stack([rng.uniform(-2, 2, n_mp), rng.uniform(-1.5, 1.5, n_mp), rng.uniform(3, 10, n_mp)], 1)
    Xw = np.einsum("nij,nj->ni", Rwb[k0], (Xc0 - tcb) @ Rcb) + twb[k0]
    Xb = np.einsum("kji,nkj->nki", Rwb, Xw[:, None, :] - twb[None])          # [n_mp][n_kf][3]
    Xc = Xb @ Rcb.T + tcb
    z = Xc[..., 2]
    zs = np.where(z > 0.5, z, 1.0)
    u, vv = FX * Xc[..., 0] / zs + CX, FY * Xc[..., 1] / zs + CY
    vis = (z > 0.5) & (u >= 0) & (u < 640) & (vv >= 0) & (vv < 480)
    keep = vis.sum(1) >= 2
    Xw, vis, u, vv, z = Xw[keep], vis[keep], u[keep], vv[keep], z[keep]
    emp, ekf = np.nonzero(vis)
    ur = np.where(rng.random(len(emp)) < 0.7, u[emp, ekf] - BF / z[emp, ekf] + rng.normal(0, noise, len(emp)), -1.0)
    obs = np.stack([u[emp, ekf] + rng.normal(0, noise, len(emp)), vv[emp, ekf] + rng.normal(0, noise, len(emp)), ur], 1)
    level = rng.integers(0, 8, len(emp))
    inv_sigma2 = (1.0 / np.float32(1.2) ** (2 * level)).astype(np.float64)
    fixed = np.zeros(n_kf, np.uint8)
    fixed[0] = 1
    fixed[n_chain:] = 1
    point = Xw.copy()
    if perturb:
        for k in range(1, n_chain):
            Rk = states[k, :9].reshape(3, 3)
            states[k, 9:12] += Rk @ rng.normal(0, 0.02, 3)
            states[k, :9] = (Rk @ _expm(rng.normal(0, 0.005, 3))).reshape(-1)
            states[k, 12:15] += rng.normal(0, 0.03, 3)
            states[k, 15:18] += rng.normal(0, 1e-3, 3)
            states[k, 18:21] += rng.normal(0, 1e-2, 3)
        point += rng.normal(0, 0.03, point.shape)
    out = dict(state=states, fixed=fixed, point=point, edge_kf=ekf.astype(np.int32), edge_mp=emp.astype(np.int32), obs=obs,
               inv_sigma2=inv_sigma2, links=links, Tcb=np.concatenate([Rcb.reshape(-1), tcb]), cam5=[FX, FY, CX, CY, BF])
    if keep_pre:
        out["pre"] = pres       # the IMU::Preintegrated members behind links[k] (tests/test_host_liba_vs_ref.py builds a mock map from them)
    return out

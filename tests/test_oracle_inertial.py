"""CPU tier: the LocalInertialBA oracle (the device side is csrc/liba_core.cuh, checked against this oracle by tests/test_liba_emul.py).  Pins available without the reference:
finite differences of EdgeInertial's analytic Jacobians under the vertices' own update rules, the noise-free fixed point, and
recovery of a perturbed trajectory."""
import numpy as np
import pytest

from oracle import pyoracle as po

G = np.array([0, 0, -float(np.float32(9.81))])


def expm(w):
    w = np.asarray(w, float)
    t = np.linalg.norm(w)
    W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if t < 1e-9:
        return np.eye(3) + W
    return np.eye(3) + np.sin(t) / t * W + (1 - np.cos(t)) / t ** 2 * W @ W


def make_trajectory(n_kf, rng, dt=0.25):
    """Smooth body trajectory + exact preintegrated deltas between consecutive keyframes (zero bias, no noise)."""
    states, links = [], []
    R, p, v = expm(rng.normal(0, 0.2, 3)), rng.normal(0, 1, 3), rng.normal(0, 0.5, 3)
    for k in range(n_kf):
        states.append(np.concatenate([R.reshape(-1), p, v, np.zeros(3), np.zeros(3)]))
        if k == n_kf - 1:
            break
        w, a = rng.normal(0, 0.3, 3), rng.normal(0, 1.0, 3)           # constant body rates over the interval
        dR = expm(w * dt)
        R2 = R @ dR
        acc_w = R @ a                                                 # crude: world acceleration constant over the interval
        v2 = v + (acc_w + G) * dt
        p2 = p + v * dt + 0.5 * (acc_w + G) * dt * dt
        dV = R.T @ (v2 - v - G * dt)
        dP = R.T @ (p2 - p - v * dt - 0.5 * G * dt * dt)
        lk = np.zeros(1, po.LIBA_LINK)[0]
        lk["k1"], lk["k2"], lk["robust"], lk["dt"] = k, k + 1, 0, dt
        lk["dR"], lk["dV"], lk["dP"] = dR.reshape(-1), dV, dP
        lk["JRg"] = (-np.eye(3) * dt).reshape(-1)                     # plausible first-order bias Jacobians
        lk["JVg"] = (rng.normal(0, 0.05, (3, 3))).reshape(-1)
        lk["JVa"] = (-np.eye(3) * dt).reshape(-1)
        lk["JPg"] = (rng.normal(0, 0.01, (3, 3))).reshape(-1)
        lk["JPa"] = (-np.eye(3) * 0.5 * dt * dt).reshape(-1)
        lk["blin"] = 0
        A = rng.normal(size=(9, 9))
        lk["info"] = (A @ A.T + 9 * np.eye(9)).reshape(-1) * 50
        lk["infoG"] = (np.eye(3) * 1e4).reshape(-1)
        lk["infoA"] = (np.eye(3) * 1e3).reshape(-1)
        links.append(lk)
        R, p, v = R2, p2, v2
    return np.array(states), np.array(links, po.LIBA_LINK)


def test_inertial_edge_jacobians_by_finite_differences():
    rng = np.random.default_rng(0)
    states, links = make_trajectory(2, rng)
    states[0, 15:21] = rng.normal(0, 0.01, 6)          # non-zero biases so that the bias columns are exercised
    states[1, 12:15] += rng.normal(0, 0.1, 3)          # off the fixed point: non-zero residual
    states[1, :9] = (states[1, :9].reshape(3, 3) @ expm(rng.normal(0, 0.05, 3))).reshape(-1)
    e0, J = po.inertial_edge(states, links[0])
    assert np.abs(e0).max() > 1e-3
    h = 1e-6
    Jn = np.zeros((9, 24))
    for c in range(24):
        d1, d2 = np.zeros(15), np.zeros(15)
        for sgn in (+1, -1):
            s = states.copy()
            if c < 15:
                d = np.zeros(15); d[c] = sgn * h
                s[0] = po.kf_oplus(s[0], d)
            else:
                d = np.zeros(15); d[c - 15] = sgn * h
                s[1] = po.kf_oplus(s[1], d)
            e, _ = po.inertial_edge(s, links[0])
            Jn[:, c] += sgn * e / (2 * h)
    # the float evaluation of the preintegrated deltas (ImuTypes.cc) limits the bias columns to ~1e-2 relative
    assert np.abs(J[:, :9] - Jn[:, :9]).max() < 1e-5            # pose 1, velocity 1
    assert np.abs(J[:, 15:24] - Jn[:, 15:24]).max() < 1e-5      # pose 2, velocity 2
    assert np.abs(J[:, 9:15] - Jn[:, 9:15]).max() < 5e-2 * max(1.0, np.abs(J[:, 9:15]).max())


def scene(n_kf=8, n_mp=400, seed=1, noise=0.0):
    rng = np.random.default_rng(seed)
    states, links = make_trajectory(n_kf, rng)
    FX, FY, CX, CY, BF = 435.2, 435.2, 320.0, 240.0, 47.9
    Rcb, tcb = expm([0.01, -0.02, 0.015]), np.array([0.05, -0.01, 0.02])
    Tcb = np.concatenate([Rcb.reshape(-1), tcb])
    pts, ekf, emp, obs = [], [], [], []
    for j in range(n_mp):
        k0 = rng.integers(0, n_kf)
        Rwb, twb = states[k0, :9].reshape(3, 3), states[k0, 9:12]
        Xc = np.array([rng.uniform(-2, 2), rng.uniform(-1.5, 1.5), rng.uniform(3, 10)])
        Xb = Rcb.T @ (Xc - tcb)
        Xw = Rwb @ Xb + twb
        seen = 0
        for k in range(n_kf):
            Rwb, twb = states[k, :9].reshape(3, 3), states[k, 9:12]
            Xc = Rcb @ (Rwb.T @ (Xw - twb)) + tcb
            if Xc[2] < 0.5:
                continue
            u, v = FX * Xc[0] / Xc[2] + CX, FY * Xc[1] / Xc[2] + CY
            if not (0 <= u < 640 and 0 <= v < 480):
                continue
            ur = u - BF / Xc[2] if rng.random() < 0.7 else -1.0
            ekf.append(k); emp.append(len(pts)); obs.append([u + rng.normal(0, noise), v + rng.normal(0, noise), ur if ur < 0 else ur + rng.normal(0, noise)])
            seen += 1
        if seen >= 2:
            pts.append(Xw)
        else:
            del ekf[len(ekf) - seen:], emp[len(emp) - seen:], obs[len(obs) - seen:]
    fixed = np.zeros(n_kf, np.uint8)
    fixed[0] = 1
    return dict(state=states, links=links, Tcb=Tcb, cam5=[FX, FY, CX, CY, BF], point=np.array(pts), edge_kf=np.array(ekf, np.int32),
                edge_mp=np.array(emp, np.int32), obs=np.array(obs), inv_sigma2=np.ones(len(ekf)), fixed=fixed)


def test_noise_free_problem_is_a_fixed_point():
    s = scene()
    r = po.liba(s["state"], s["fixed"], s["point"], s["edge_kf"], s["edge_mp"], s["obs"], s["inv_sigma2"], s["Tcb"], s["cam5"], s["links"], 1.0, 10)
    assert r["chi2_init"] < 1e-3                       # float storage of the preintegrated deltas leaves ~1e-7 residuals
    assert np.abs(r["state"] - s["state"]).max() < 1e-5 and np.abs(r["point"] - s["point"]).max() < 1e-4


def test_recovers_a_perturbed_trajectory():
    s = scene(seed=2)
    rng = np.random.default_rng(3)
    st = s["state"].copy()
    for k in range(1, len(st)):
        st[k] = po.kf_oplus(st[k], np.concatenate([rng.normal(0, 0.01, 3), rng.normal(0, 0.03, 3), rng.normal(0, 0.05, 3), np.zeros(6)]))
    pt = s["point"] + rng.normal(0, 0.05, s["point"].shape)
    r = po.liba(st, s["fixed"], pt, s["edge_kf"], s["edge_mp"], s["obs"], s["inv_sigma2"], s["Tcb"], s["cam5"], s["links"], 1.0, 25)
    assert r["chi2"] < 1e-3 * r["chi2_init"]
    assert np.abs(r["state"][:, 9:12] - s["state"][:, 9:12]).max() < 2e-3          # positions
    assert np.abs(r["state"][:, :9] - s["state"][:, :9]).max() < 1e-3              # rotations
    assert np.abs(r["state"][:, 12:15] - s["state"][:, 12:15]).max() < 2e-2        # velocities
    assert (r["state"][0] == s["state"][0]).all()                                  # the fixed keyframe is untouched


@pytest.mark.parametrize("stereo", [False, True])
def test_reprojection_jacobians_by_finite_differences(stereo):
    rng = np.random.default_rng(5)
    s = scene(n_kf=2, n_mp=5, seed=6)
    st = s["state"][1]
    Xw = s["point"][0] + rng.normal(0, 0.05, 3)
    obs = np.array([300.0, 200.0, 290.0 if stereo else -1.0])
    D, r0, Jp, Jx = po.liba_reproj(st, Xw, obs, s["Tcb"], s["cam5"])
    assert D == (3 if stereo else 2)
    h = 1e-6
    for c in range(3):
        d = np.zeros(3); d[c] = h
        rp = po.liba_reproj(st, Xw + d, obs, s["Tcb"], s["cam5"])[1]
        rm = po.liba_reproj(st, Xw - d, obs, s["Tcb"], s["cam5"])[1]
        assert np.abs((rp - rm)[:D] / (2 * h) - Jp[:, c]).max() < 1e-4
    for c in range(6):
        d = np.zeros(15); d[c] = h
        rp = po.liba_reproj(po.kf_oplus(st, d), Xw, obs, s["Tcb"], s["cam5"])[1]
        rm = po.liba_reproj(po.kf_oplus(st, -d), Xw, obs, s["Tcb"], s["cam5"])[1]
        assert np.abs((rp - rm)[:D] / (2 * h) - Jx[:, c]).max() < 1e-4


def test_golden_window():
    """The committed window (tests/golden/golden_liba.npz): the oracle still produces what it produced when the fixture was made, and
    the generator still produces the fixture's inputs."""
    import os
    from orb_slam3_detailed_comments_b200 import synth
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_liba.npz"))
    links = np.frombuffer(z["in_links"].tobytes(), po.LIBA_LINK)
    r = po.liba(z["in_state"], z["in_fixed"], z["in_point"], z["in_edge_kf"], z["in_edge_mp"], z["in_obs"], z["in_inv_sigma2"], z["in_Tcb"],
                z["in_cam5"], links, 1.0, 10)
    sc = z["scalars"]
    assert r["iterations"] == int(sc[0]) and r["trials"] == int(sc[1])
    assert np.allclose([r["lambda_"], r["chi2"], r["chi2_init"], r["chi2_last"]], sc[2:], rtol=1e-9)
    assert np.abs(r["state"] - z["state"]).max() < 1e-9 and np.abs(r["point"] - z["point"]).max() < 1e-9
    assert np.allclose(r["edge_chi2"], z["edge_chi2"], rtol=1e-7, atol=1e-9) and np.allclose(r["link_chi2"], z["link_chi2"], rtol=1e-7, atol=1e-9)
    assert (r["edge_depth_pos"] == z["edge_depth_pos"]).all()
    s = synth.inertial_window(n_opt=5, n_cov_fixed=2, n_mp=120, seed=17)
    assert np.allclose(s["state"], z["in_state"], rtol=0, atol=1e-9) and np.allclose(s["obs"], z["in_obs"], rtol=0, atol=1e-6)
    assert np.allclose(s["links"]["info"], links["info"], rtol=1e-6) and np.allclose(s["links"]["dP"], links["dP"], rtol=0, atol=1e-7)


def test_bias_correction_agrees_with_reintegration():
    """EdgeInertial evaluates the preintegrated deltas at the current bias estimate through the first-order correction
    dR Exp(JRg dbg), dV + JVg dbg + JVa dba, dP + JPg dbg + JPa dba (ImuTypes.cc:383-408).  Independent check: re-preintegrate the same
    IMU samples AT the new bias (no correction needed) -- the two residuals must agree to second order in the bias change."""
    from orb_slam3_detailed_comments_b200 import synth
    rng = np.random.default_rng(11)
    n, dt = 50, 0.005
    gyr = rng.normal(0, 0.2, (n, 3)) + np.array([0.1, -0.2, 0.05])
    acc = rng.normal(0, 0.5, (n, 3)) + np.array([0.2, 0.1, 9.7])

    def link_of(pre):
        lk = np.zeros(1, po.LIBA_LINK)[0]
        lk["k1"], lk["k2"], lk["dt"] = 0, 1, pre["dT"]
        for name in ("dR", "dV", "dP", "JRg", "JVg", "JVa", "JPg", "JPa"):
            lk[name] = pre[name].reshape(-1)
        lk["blin"] = pre["bias"]
        lk["info"] = np.eye(9).reshape(-1)
        return lk

    st = np.zeros((2, 21))
    st[0, :9] = expm(rng.normal(0, 0.3, 3)).reshape(-1)
    st[1, :9] = expm(rng.normal(0, 0.3, 3)).reshape(-1)
    st[:, 9:15] = rng.normal(0, 1, (2, 6))
    pre0 = synth.preintegrate(acc, gyr, dt)
    errs = []
    for scale in (1.0, 0.5):
        dbg, dba = scale * np.array([2e-3, -1e-3, 1.5e-3]), scale * np.array([2e-2, 1e-2, -1.5e-2])
        s = st.copy()
        s[0, 15:18], s[0, 18:21] = dbg, dba
        e_corr, _ = po.inertial_edge(s, link_of(pre0))                              # deltas at bias 0, corrected to (dba, dbg)
        e_true, _ = po.inertial_edge(s, link_of(synth.preintegrate(acc, gyr, dt, np.concatenate([dba, dbg]))))
        errs.append(np.abs(e_corr - e_true).max())
        first_order = np.abs(po.inertial_edge(st, link_of(pre0))[0] - e_true).max()  # ignoring the bias change altogether
        assert errs[-1] < 0.02 * first_order
    assert errs[1] < 0.4 * errs[0] + 2e-6        # halving the bias change quarters the discrepancy (float32 deltas: 1e-6 floor)


def _log_so3(R):
    c = min(1.0, max(-1.0, (np.trace(R) - 1) / 2))
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2
    th = np.arccos(c)
    return w if abs(np.sin(th)) < 1e-5 else th * w / np.sin(th)


def numpy_cost(s, state, point):
    """The objective LocalInertialBA hands to g2o, in plain numpy from the reference's edge definitions (G2oTypes.cc:390-490, 592-611,
    G2oTypes.h:736-800; Huber sqrt(5.991) / sqrt(7.815) / sqrt(16.92) where installed)."""
    fx, fy, cx, cy, bf = s["cam5"]
    Rcb, tcb = np.asarray(s["Tcb"][:9]).reshape(3, 3), np.asarray(s["Tcb"][9:])
    hub = lambda c, d: c if c <= float(np.float32(d * d)) else 2 * np.sqrt(c) * d - float(np.float32(d * d))
    dM, dS, dI = float(np.float32(np.sqrt(5.991))), float(np.float32(np.sqrt(7.815))), np.sqrt(16.92)
    total = 0.0
    for e in range(len(s["edge_kf"])):
        st = state[s["edge_kf"][e]]
        Rwb, twb = st[:9].reshape(3, 3), st[9:12]
        Xc = Rcb @ (Rwb.T @ (point[s["edge_mp"][e]] - twb)) + tcb
        u, v = fx * Xc[0] / Xc[2] + cx, fy * Xc[1] / Xc[2] + cy
        z = s["obs"][e]
        r = [z[0] - u, z[1] - v] + ([] if z[2] < 0 else [z[2] - (u - bf / Xc[2])])
        total += hub(s["inv_sigma2"][e] * float(np.dot(r, r)), dM if z[2] < 0 else dS)
    links = np.asarray(s["links"]).view(po.LIBA_LINK).reshape(-1)
    for L in links:
        s1, s2 = state[L["k1"]], state[L["k2"]]
        R1, R2 = s1[:9].reshape(3, 3), s2[:9].reshape(3, 3)
        dbg, dba = s1[15:18] - L["blin"][3:].astype(float), s1[18:21] - L["blin"][:3].astype(float)
        m = lambda k: L[k].astype(float).reshape(3, 3)
        dR = m("dR") @ expm(m("JRg") @ dbg)
        dV = L["dV"].astype(float) + m("JVg") @ dbg + m("JVa") @ dba
        dP = L["dP"].astype(float) + m("JPg") @ dbg + m("JPa") @ dba
        dt = L["dt"]
        e9 = np.concatenate([_log_so3(dR.T @ R1.T @ R2), R1.T @ (s2[12:15] - s1[12:15] - G * dt) - dV,
                             R1.T @ (s2[9:12] - s1[9:12] - s1[12:15] * dt - G * dt * dt / 2) - dP])
        c = float(e9 @ L["info"].reshape(9, 9) @ e9)
        total += hub(c, dI) if L["robust"] else c
        eg, ea = s2[15:18] - s1[15:18], s2[18:21] - s1[18:21]
        total += float(eg @ L["infoG"].reshape(3, 3) @ eg) + float(ea @ L["infoA"].reshape(3, 3) @ ea)
    return total


def test_objective_and_stationarity_against_numpy():
    """Independent of the oracle's Jacobians, Schur complement and LM control: (i) the chi2 the oracle reports IS the numpy objective at
    the initial and at the final estimate; (ii) run to convergence, the final estimate is a stationary point of that objective along
    positions, velocities, biases of the free keyframes and the coordinates of a sample of map points."""
    from orb_slam3_detailed_comments_b200 import synth
    s = synth.inertial_window(n_opt=4, n_cov_fixed=2, n_mp=90, seed=23)
    r = po.liba(s["state"], s["fixed"], s["point"], s["edge_kf"], s["edge_mp"], s["obs"], s["inv_sigma2"], s["Tcb"], s["cam5"],
                s["links"].view(po.LIBA_LINK), 1.0, 200)
    c0, c1 = numpy_cost(s, s["state"], s["point"]), numpy_cost(s, r["state"], r["point"])
    assert abs(r["chi2_init"] - c0) <= 2e-5 * c0 and abs(r["chi2"] - c1) <= 2e-5 * c1 + 1e-6     # float32 delta storage inside the edge
    assert c1 < 0.01 * c0

    def grad(state, point, h):
        g = []
        for k in np.nonzero(s["fixed"] == 0)[0]:
            for c in range(9, 21):
                a, b = state.copy(), state.copy()
                a[k, c] += h[c]
                b[k, c] -= h[c]
                g.append((numpy_cost(s, a, point) - numpy_cost(s, b, point)) / (2 * h[c]))
        for l in range(0, len(point), 9):
            for c in range(3):
                a, b = point.copy(), point.copy()
                a[l, c] += 1e-6
                b[l, c] -= 1e-6
                g.append((numpy_cost(s, state, a) - numpy_cost(s, state, b)) / 2e-6)
        return np.array(g)
    h = np.full(21, 1e-6)
    h[15:18] = 1e-8            # gyro bias: information ~1e10, keep the probe inside the quadratic region
    g0, g1 = grad(s["state"], s["point"], h), grad(r["state"], r["point"], h)
    scale = np.maximum(np.abs(g0), 1e-3 * np.abs(g0).max())
    assert (np.abs(g1) / scale).max() < 5e-3, (np.abs(g1) / scale).max()

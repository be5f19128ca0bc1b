"""CPU tier: an independent check of the LocalBundleAdjustment oracle's OBJECTIVE.  The cost g2o minimises -- sum over edges of
Huber(inv_sigma2 * |z - pi(R X + t)|^2), stereo z = (u, v, u - bf / Z), deltas sqrt(5.991) / sqrt(7.815) -- is written here in plain
numpy straight from the reference (Optimizer.cc:1957-2090, types_six_dof_expmap.cpp:190-197, robust_kernel_impl.cpp:78-91); run to
convergence, the oracle's estimate must be a stationary point of THAT function: numerical directional derivatives along every map
point coordinate and every free keyframe's translation vanish (relative to their size at the starting estimate), and the cost the
oracle reports equals the numpy cost.  This is independent of the oracle's own Jacobians, Schur complement and LM control."""
import numpy as np

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import synth


def quat_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def cost(pose, point, pr):
    fx, fy, cx, cy, bf = pr["cam5"]
    dM, dS = float(np.float32(np.sqrt(5.991))), float(np.float32(np.sqrt(7.815)))
    R = np.stack([quat_R(p[:4]) for p in pose])
    Xc = np.einsum("eij,ej->ei", R[pr["edge_kf"]], point[pr["edge_mp"]]) + pose[pr["edge_kf"], 4:7]
    u, v = fx * Xc[:, 0] / Xc[:, 2] + cx, fy * Xc[:, 1] / Xc[:, 2] + cy
    ur = u - bf / Xc[:, 2]
    obs = pr["obs"]
    mono = obs[:, 2] < 0
    e2 = (obs[:, 0] - u) ** 2 + (obs[:, 1] - v) ** 2 + np.where(mono, 0.0, (obs[:, 2] - ur) ** 2)
    chi = pr["inv_sigma2"] * e2
    d = np.where(mono, dM, dS)
    return float(np.where(chi <= d * d, chi, 2 * np.sqrt(chi) * d - d * d).sum())


def directional(pose, point, pr, h=1e-6):
    g = []
    for l in range(0, len(point), 7):                      # a sample of the map points, all three coordinates
        for c in range(3):
            p1, p2 = point.copy(), point.copy()
            p1[l, c] += h
            p2[l, c] -= h
            g.append((cost(pose, p1, pr) - cost(pose, p2, pr)) / (2 * h))
    for k in np.nonzero(pr["fixed"] == 0)[0]:              # every free keyframe: the three translation directions
        for c in range(3):
            q1, q2 = pose.copy(), pose.copy()
            q1[k, 4 + c] += h
            q2[k, 4 + c] -= h
            g.append((cost(q1, point, pr) - cost(q2, point, pr)) / (2 * h))
    return np.array(g)


def test_converged_estimate_is_a_stationary_point_of_the_reference_cost():
    pr = synth.lba_problem(n_kf=8, n_fixed=2, n_mp=250, seed=5)
    pr = {k: (np.asarray(v) if not np.isscalar(v) else v) for k, v in pr.items()}
    r = po.lba(pr["pose"], pr["fixed"], pr["point"], pr["edge_kf"], pr["edge_mp"], pr["obs"], pr["inv_sigma2"], pr["cam5"], 0.0, 100)
    c0, c1 = cost(np.asarray(pr["pose"], float), np.asarray(pr["point"], float), pr), cost(r["pose"], r["point"], pr)
    assert abs(r["chi2_init"] - c0) <= 1e-6 * c0                 # the oracle's activeRobustChi2 is this function
    assert c1 < 0.5 * c0 and abs(r["chi2"] - c1) <= 1e-6 * c1
    g0 = directional(np.asarray(pr["pose"], float), np.asarray(pr["point"], float), pr)
    g1 = directional(r["pose"], r["point"], pr)
    assert np.abs(g1).max() < 2e-4 * np.abs(g0).max(), (np.abs(g1).max(), np.abs(g0).max())

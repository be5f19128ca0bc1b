"""CPU tier, build container only: the oracle's Levenberg-Marquardt loop against g2o's OWN.
oracle/_ref/liborb_ref4.so holds OptimizationAlgorithmLevenberg::solve / computeLambdaInit / computeScale (and the constructor's
constants), SparseOptimizer::optimize and RobustKernelHuber::setDelta / robustify, cut out of /root/reference/Thirdparty/g2o at build
time and compiled verbatim; the abstract Solver / SparseOptimizer they call are the oracle's LbaEngine (errors, buildSystem, solve,
update, push / pop).  So both runs share the numerics and differ only in who decides: g2o's control flow or orc_lba's restatement of
it -- accept / reject by rho, the lambda update and its clamps, _ni doubling, the ten-trial limit, ORB-SLAM3's nBad stop, the
iteration loop.  Equality is exact: poses, points, per-edge chi2 (bits), iterations, trials, lambda."""
import numpy as np
import pytest

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import synth

pytestmark = pytest.mark.skipif(po.build_ref4() is None, reason="oracle/_ref part 4 not built and /root/reference absent")


def _same(a, b):
    assert a["iterations"] == b["iterations"] and a["trials"] == b["trials"]
    assert np.float64(a["lambda_"]).view(np.uint64) == np.float64(b["lambda_"]).view(np.uint64)
    # (orc_lba reports the chi2 of the last ACCEPTED state, the g2o skeleton that of the last computeActiveErrors: equal unless the run
    # ended on a rejected trial; the per-edge chi2 below are the last computeActiveErrors' in both)
    for k in ("pose", "point", "edge_chi2"):
        assert (a[k].view(np.uint64) == b[k].view(np.uint64)).all(), k
    assert (a["edge_depth_pos"] == b["edge_depth_pos"]).all()


def _args(p):
    return (p["pose"], p["fixed"], p["point"], p["edge_kf"], p["edge_mp"], p["obs"], p["inv_sigma2"], p["cam5"])


@pytest.mark.parametrize("seed,n_kf,n_mp,lam,iters", [(0, 8, 300, 100.0, 10), (1, 8, 300, 0.0, 10), (2, 12, 500, 100.0, 5), (3, 6, 150, 1e-3, 10),
                                                      (4, 20, 800, 100.0, 10)])
def test_levenberg_loop_on_local_ba_problems(seed, n_kf, n_mp, lam, iters):
    p = synth.lba_problem(n_kf=n_kf, n_fixed=2, n_mp=n_mp, seed=seed)
    a, b = po.lba(*_args(p), lambda_init=lam, max_iters=iters), po.ref4_lba(*_args(p), lambda_init=lam, max_iters=iters)
    _same(a, b)
    assert np.float64(a["chi2"]).view(np.uint64) == np.float64(b["chi2"]).view(np.uint64)
    assert a["iterations"] >= 3 and a["chi2"] < a["chi2_init"]


@pytest.mark.parametrize("seed,sig_p,sig_t,lam", [(0, 2.0, 0.5, 1e-9), (0, 1.5, 0.4, 1e-12), (1, 4.0, 1.0, 1e-9), (3, 1.5, 0.4, 1e-12), (2, 3.0, 0.0, 1e-6)])
def test_rejected_steps_and_lambda_growth(seed, sig_p, sig_t, lam):
    """A tiny initial lambda on a badly perturbed problem: trials are rejected (rho < 0), lambda grows by _ni = 2, 4, 8, ..., a step is
    accepted and lambda shrinks by max(1/3, 1 - (2 rho - 1)^3) clamped at 2/3; some runs end on the ten-trial limit."""
    p = synth.lba_problem(n_kf=8, n_fixed=2, n_mp=300, seed=7 + seed)
    rng = np.random.default_rng(seed)
    p["point"] = p["point"] + rng.normal(0, sig_p, p["point"].shape)
    p["pose"][2:, 4:] += rng.normal(0, sig_t, (6, 3))
    a, b = po.lba(*_args(p), lambda_init=lam, max_iters=10), po.ref4_lba(*_args(p), lambda_init=lam, max_iters=10)
    _same(a, b)
    assert a["trials"] >= a["iterations"] + 3                  # several rejected trials


def test_nbad_stop_of_the_orbslam3_fork():
    """Started at the optimum of a noise-free problem the chi2 cannot improve by 0.1 % three times in a row: Terminate after 3 iterations
    (optimization_algorithm_levenberg.cpp:163-171)."""
    p = synth.lba_problem(n_kf=8, n_fixed=2, n_mp=300, seed=5, outlier_frac=0.0)
    first = po.lba(*_args(p), lambda_init=100.0, max_iters=10)
    p2 = dict(p, pose=first["pose"], point=first["point"])
    a, b = po.lba(*_args(p2), lambda_init=100.0, max_iters=10), po.ref4_lba(*_args(p2), lambda_init=100.0, max_iters=10)
    _same(a, b)
    assert a["iterations"] <= 4


def test_ten_trial_limit():
    """Every trial rejected (NaN-free but hopeless: observations moved by hundreds of pixels with lambda tiny): qmax reaches 10 -> Terminate."""
    p = synth.lba_problem(n_kf=6, n_fixed=2, n_mp=120, seed=9)
    q = dict(p)
    q["point"] = p["point"] * np.array([1.0, 1.0, -1.0])      # points behind the cameras: steps never reduce the robust chi2 reliably
    a, b = po.lba(*_args(q), lambda_init=1e-12, max_iters=10), po.ref4_lba(*_args(q), lambda_init=1e-12, max_iters=10)
    _same(a, b)


def test_huber_kernel_with_its_float_member():
    """RobustKernelHuber keeps dsqr as a float (robust_kernel_impl.h:84): the inlier / outlier switch and rho use the rounded value."""
    rng = np.random.default_rng(0)
    for delta in (np.sqrt(5.991), np.sqrt(7.815), 1.0, 0.3):
        d = np.float32(delta)
        for e in np.concatenate([rng.uniform(0, 30, 200), [float(np.float32(d * d)), float(np.float64(d) * np.float64(d)), 5.991, 7.815, 0.0]]):
            r, o = po.ref4_huber(e, d), po.huber(e, d)
            assert np.float64(r[0]).view(np.uint64) == np.float64(o[0]).view(np.uint64) and np.float64(r[1]).view(np.uint64) == np.float64(o[1]).view(np.uint64)


# ---- Optimizer::PoseOptimization: the reference's own function (graph construction, four rounds, float chi2 classification, levels,
# kernel removal, return value) over the oracle's PoseEngine, against orc_pose_optimization ------------------------------------------------
def _quat(axis, ang):
    a = np.asarray(axis, float)
    a /= np.linalg.norm(a)
    return np.concatenate([a * np.sin(ang / 2), [np.cos(ang / 2)]])


def _qR(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


FX, FY, CX, CY, BF = 435.2, 435.2, 320.0, 240.0, 47.9


def _frame(seed, N, mp_frac, mono_frac, outlier_frac, motion=1.0, noise=0.7):
    rng = np.random.default_rng(seed)
    Xc = np.stack([rng.uniform(-3, 3, N), rng.uniform(-2, 2, N), rng.uniform(2, 12, N)], 1)
    qt, tt = _quat(rng.normal(size=3), 0.03 * motion), rng.normal(0, 0.05 * motion, 3)
    Xw = ((Xc - tt) @ _qR(qt)).astype(np.float32)
    u, v = FX * Xc[:, 0] / Xc[:, 2] + CX, FY * Xc[:, 1] / Xc[:, 2] + CY
    xy = (np.stack([u, v], 1) + rng.normal(0, noise, (N, 2))).astype(np.float32)
    ur = (u - BF / Xc[:, 2] + rng.normal(0, noise, N)).astype(np.float32)
    ur[rng.random(N) < mono_frac] = -1
    bad = rng.random(N) < outlier_frac
    xy[bad] += rng.normal(0, 30, (int(bad.sum()), 2)).astype(np.float32)
    octave = rng.integers(0, 8, N).astype(np.int32)
    has_mp = (rng.random(N) < mp_frac).astype(np.uint8)
    isg = (1.0 / (np.float32(1.2) ** np.arange(8, dtype=np.float32)) ** 2).astype(np.float32)
    return dict(pose=np.array([0, 0, 0, 1, 0, 0, 0], np.float32), has_mp=has_mp, world_pos=Xw, kp_xy=xy, octave=octave, uright=ur, isg=isg)


@pytest.mark.skipif(po.build_ref5() is None, reason="oracle/_ref part 5 not built and /root/reference absent")
@pytest.mark.parametrize("seed,N,mp_frac,mono_frac,outlier_frac,motion", [(1, 1200, 0.5, 0.3, 0.15, 1.0), (2, 600, 0.9, 0.0, 0.05, 0.5), (3, 800, 0.4, 1.0, 0.2, 1.0),
                                                                          (4, 60, 0.5, 0.2, 0.3, 2.0), (5, 30, 0.3, 0.5, 0.1, 1.0), (6, 12, 0.2, 0.5, 0.0, 1.0),
                                                                          (7, 1500, 0.6, 0.3, 0.45, 3.0)])
def test_pose_optimization_against_the_reference_function(seed, N, mp_frac, mono_frac, outlier_frac, motion):
    f = _frame(seed, N, mp_frac, mono_frac, outlier_frac, motion)
    cam5 = np.array([FX, FY, CX, CY, BF], np.float32)
    r = po.ref5_pose_optimization(f["pose"], f["has_mp"], f["world_pos"], f["kp_xy"], f["octave"], f["uright"], f["isg"], cam5)
    sel = np.nonzero(f["has_mp"])[0]                      # the oracle takes the correspondences: features with a map point, in order
    obs = np.concatenate([f["kp_xy"][sel], f["uright"][sel, None]], 1)
    o = po.pose_optimization(f["pose"], f["world_pos"][sel], obs, f["isg"][f["octave"][sel]], cam5)
    assert r["inliers"] == o["inliers"]
    assert (r["outlier"][sel] == o["outlier"]).all() and not r["outlier"][f["has_mp"] == 0].any()
    if len(sel) >= 3:                                     # fewer: the function returns before it touches the pose
        assert (r["pose"].view(np.uint32) == o["pose"].astype(np.float32).view(np.uint32)).all()   # SetPose takes the estimate as float
    else:
        assert r["inliers"] == 0

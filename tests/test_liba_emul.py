"""CPU tier: the DEVICE algorithm of LocalInertialBA (csrc/liba_core.cuh -- the very source nvcc compiles into k_liba) built for
the host and run against the oracle (oracle/lba_oracle.cpp orc_liba).  Same phases, same packing (csrc/liba_pack.h), same C ABI
structs; what this cannot see are device-only hazards (barrier placement, atomics), which the GPU test covers once it has run."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import _native as N
from orb_slam3_detailed_comments_b200.optimizer import finish_inertial_result, pack_inertial_problem
from test_oracle_inertial import scene

TOL = 1e-4   # BASELINE.json north_star tolerance for the BA rows
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul():
    from _emul import build_and_load
    L = build_and_load()
    L.emul_liba.restype = C.c_int
    L.emul_liba.argtypes = [C.POINTER(N.liba_problem), C.POINTER(N.liba_result)]
    L.emul_liba_inertial.restype = None
    L.emul_liba_inertial.argtypes = [C.c_void_p] * 4
    return L


def run_emul(L, pr, lambda_init, max_iters):
    keep = []
    p, r, out = pack_inertial_problem(pr, lambda_init, max_iters, keep)
    L.emul_liba(C.byref(p), C.byref(r))
    return finish_inertial_result(keep[0][0], out, r)


def perturbed(seed, n_kf=8, n_mp=300, noise=0.5):
    s = scene(n_kf=n_kf, n_mp=n_mp, seed=seed, noise=noise)
    rng = np.random.default_rng(seed + 100)
    st = s["state"].copy()
    for k in range(1, len(st)):
        st[k] = po.kf_oplus(st[k], np.concatenate([rng.normal(0, 0.01, 3), rng.normal(0, 0.03, 3), rng.normal(0, 0.05, 3),
                                                   rng.normal(0, 1e-3, 3), rng.normal(0, 1e-2, 3)]))
    s["state"] = st
    s["point"] = s["point"] + rng.normal(0, 0.05, s["point"].shape)
    return s


def test_link_layouts_agree():
    assert N.LIBA_LINK.itemsize == po.LIBA_LINK.itemsize == 1080
    for (n1, n2) in zip(N.LIBA_LINK.names, po.LIBA_LINK.names):
        assert N.LIBA_LINK.fields[n1][1] == po.LIBA_LINK.fields[n2][1]


def test_inertial_edge_matches_oracle(emul):
    s = perturbed(5)
    for l in range(len(s["links"])):
        lk = np.ascontiguousarray(s["links"][l:l + 1])
        st = np.ascontiguousarray(s["state"][[int(lk["k1"][0]), int(lk["k2"][0])]])
        e_o, J_o = po.inertial_edge(st, lk[0])
        e, J = np.zeros(9), np.zeros((9, 24))
        emul.emul_liba_inertial(st.ctypes.data, lk.ctypes.data, e.ctypes.data, J.ctypes.data)
        assert np.array_equal(e, e_o) and np.array_equal(J, J_o)


@pytest.mark.parametrize("seed,lam,iters,n_fixed", [(11, 1.0, 10, 1), (12, 1e-2, 4, 1), (13, 1.0, 10, 3), (14, 0.0, 6, 2)])
def test_device_algorithm_matches_oracle(emul, seed, lam, iters, n_fixed):
    s = perturbed(seed)
    s["fixed"][:] = 0
    s["fixed"][:n_fixed] = 1
    if seed == 13:
        s["links"]["robust"] = 1
    ref = po.liba(s["state"], s["fixed"], s["point"], s["edge_kf"], s["edge_mp"], s["obs"], s["inv_sigma2"], s["Tcb"], s["cam5"], s["links"], lam, iters)
    got = run_emul(emul, s, lam, iters)
    assert got["iterations"] == ref["iterations"] and got["trials"] == ref["trials"]
    assert abs(got["chi2_init"] - ref["chi2_init"]) <= 1e-9 * max(1.0, ref["chi2_init"])
    assert abs(got["chi2"] - ref["chi2"]) <= 1e-6 * max(1.0, ref["chi2"])
    assert abs(got["lambda_"] - ref["lambda_"]) <= 1e-6 * ref["lambda_"]
    assert np.abs(got["state"] - ref["state"]).max() < TOL
    assert np.abs(got["point"] - ref["point"]).max() < TOL
    assert np.allclose(got["edge_chi2"], ref["edge_chi2"], rtol=1e-6, atol=1e-8)
    assert np.allclose(got["link_chi2"], ref["link_chi2"], rtol=1e-6, atol=1e-8)
    assert abs(got["chi2_last"] - ref["chi2_last"]) <= 1e-6 * max(1.0, ref["chi2_last"])      # err_end of Optimizer.cc:2685
    assert (got["edge_depth_pos"] == ref["edge_depth_pos"]).all() and ref["edge_depth_pos"].all()
    assert (got["state"][:n_fixed] == s["state"][:n_fixed]).all()
    assert ref["chi2"] < ref["chi2_init"]


def test_noise_free_window_is_a_fixed_point(emul):
    s = scene()
    got = run_emul(emul, s, 1.0, 10)
    assert got["chi2_init"] < 1e-3
    assert np.abs(got["state"] - s["state"]).max() < 1e-5 and np.abs(got["point"] - s["point"]).max() < 1e-4


def test_mono_only_and_all_fixed_but_one(emul):
    s = perturbed(21, n_kf=5, n_mp=120)
    s["obs"][:, 2] = -1.0
    s["fixed"][:] = 1
    s["fixed"][-1] = 0
    ref = po.liba(s["state"], s["fixed"], s["point"], s["edge_kf"], s["edge_mp"], s["obs"], s["inv_sigma2"], s["Tcb"], s["cam5"], s["links"], 1.0, 10)
    got = run_emul(emul, s, 1.0, 10)
    assert got["iterations"] == ref["iterations"] and got["trials"] == ref["trials"]
    assert np.abs(got["state"] - ref["state"]).max() < TOL and np.abs(got["point"] - ref["point"]).max() < TOL


# ---- the same source run by N host threads with real barriers under ThreadSanitizer ---------------------------------------------


@pytest.fixture(scope="module")
def mt_binary():
    d = os.path.join(HERE, "host_emul")
    exe = os.path.join(d, "liba_mt_tsan")
    csrc = os.path.join(HERE, "..", "orb_slam3_detailed_comments_b200", "csrc")
    srcs = [os.path.join(d, "liba_mt.cpp"), os.path.join(csrc, "liba_core.cuh"), os.path.join(csrc, "liba_pack.h"),
            os.path.join(HERE, "..", "include", "orbslam3_b200.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(s) > os.path.getmtime(exe) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=thread", "-ffp-contract=off", "-std=c++20", "-o", exe, srcs[0], "-lpthread"])
    return exe


def run_mt(exe, s, lam, iters, threads, tmp_path, ctas=1):
    c = np.ascontiguousarray
    links = c(s["links"]).view(N.LIBA_LINK).reshape(-1)
    blob = b"".join([
        np.array([len(s["state"]), len(s["point"]), len(s["edge_kf"]), len(links), iters, 0], np.int32).tobytes(),
        np.concatenate([[lam], np.asarray(s["Tcb"], np.float64).reshape(-1), np.asarray(s["cam5"], np.float64)]).tobytes(),
        c(s["state"], np.float64).tobytes(), c(s["point"], np.float64).tobytes(), c(s["obs"], np.float64).tobytes(),
        c(s["inv_sigma2"], np.float64).tobytes(), c(s["edge_kf"], np.int32).tobytes(), c(s["edge_mp"], np.int32).tobytes(),
        links.tobytes(), c(s["fixed"], np.uint8).tobytes()])
    fin, fout = tmp_path / "p.bin", tmp_path / "r.bin"
    fin.write_bytes(blob)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66")
    pr = subprocess.run([exe, str(fin), str(fout), str(threads), str(ctas)], env=env, capture_output=True, text=True, timeout=600)
    assert pr.returncode == 0, "ThreadSanitizer / driver failure:\n" + pr.stderr[-4000:]
    o = np.frombuffer(fout.read_bytes(), np.float64)
    nk, nm, ne, nl = len(s["state"]), len(s["point"]), len(s["edge_kf"]), len(links)
    a = 5
    st = o[a:a + 21 * nk].reshape(nk, 21); a += 21 * nk
    pt = o[a:a + 3 * nm].reshape(nm, 3); a += 3 * nm
    chi = o[a:a + ne]; a += ne
    lchi = o[a:a + 3 * nl].reshape(nl, 3)
    return dict(iterations=int(o[0]), trials=int(o[1]), lambda_=o[2], chi2=o[3], chi2_init=o[4], state=st, point=pt, edge_chi2=chi, link_chi2=lchi)


@pytest.mark.parametrize("threads,ctas", [(2, 1), (7, 1), (16, 1), (4, 2), (3, 8)])
def test_threaded_run_is_race_free_and_matches_oracle(mt_binary, tmp_path, threads, ctas):
    """N threads, real barriers, real atomics, ThreadSanitizer on: a missing LIBA_SYNC in liba_core.cuh fails here.  Atomic
    accumulation order differs from run to run, so values match the oracle to the BA tolerance, not bit for bit."""
    s = perturbed(31, n_kf=6, n_mp=150)
    s["fixed"][:] = 0
    s["fixed"][:2] = 1
    ref = po.liba(s["state"], s["fixed"], s["point"], s["edge_kf"], s["edge_mp"], s["obs"], s["inv_sigma2"], s["Tcb"], s["cam5"], s["links"], 1.0, 6)
    got = run_mt(mt_binary, s, 1.0, 6, threads, tmp_path, ctas)
    assert got["iterations"] == ref["iterations"] and abs(got["trials"] - ref["trials"]) <= 1
    assert abs(got["chi2"] - ref["chi2"]) <= 1e-5 * max(1.0, ref["chi2"])
    assert np.abs(got["state"] - ref["state"]).max() < TOL and np.abs(got["point"] - ref["point"]).max() < TOL


def test_bench_window_matches_oracle(emul):
    """The window bench.py times (synth.inertial_window: 10 optimisable + 4 fixed keyframes, covisible fixed ones without links,
    level-dependent information, a robust oldest link)."""
    from orb_slam3_detailed_comments_b200 import synth
    s = synth.inertial_window(seed=3, n_mp=600)
    ref = po.liba(s["state"], s["fixed"], s["point"], s["edge_kf"], s["edge_mp"], s["obs"], s["inv_sigma2"], s["Tcb"], s["cam5"],
                  s["links"].view(po.LIBA_LINK), 1.0, 10)
    got = run_emul(emul, s, 1.0, 10)
    assert got["iterations"] == ref["iterations"] and got["trials"] == ref["trials"]
    assert np.abs(got["state"] - ref["state"]).max() < TOL and np.abs(got["point"] - ref["point"]).max() < TOL
    assert ref["chi2"] < 0.1 * ref["chi2_init"]
    assert (got["state"][s["fixed"] == 1] == s["state"][s["fixed"] == 1]).all()


def test_duplicate_links_are_rejected():
    from orb_slam3_detailed_comments_b200 import synth
    s = synth.inertial_window(seed=4, n_mp=100)
    s["links"] = np.concatenate([s["links"], s["links"][:1]])
    keep = []
    p, r, _ = pack_inertial_problem(s, 1.0, 10, keep)
    from _emul import build_and_load
    L = build_and_load()
    L.emul_liba_layout_total.restype = C.c_size_t
    L.emul_liba_layout_total.argtypes = [C.POINTER(N.liba_problem)]
    assert L.emul_liba_layout_total(C.byref(p)) == 0
    s["links"] = s["links"][:-1]
    p, r, _ = pack_inertial_problem(s, 1.0, 10, keep)
    assert L.emul_liba_layout_total(C.byref(p)) > 0


def test_link_information_helper_matches_numpy():
    """liba_link_information (host-only entry of the C ABI): EdgeInertial's information from IMU::Preintegrated::C."""
    from orb_slam3_detailed_comments_b200.optimizer import link_information
    rng = np.random.default_rng(0)
    for trial in range(12):
        A = rng.normal(size=(15, 15))
        Cm = (A @ A.T * 1e-4 + np.diag(rng.uniform(1e-6, 1e-3, 15))).astype(np.float32)
        if trial % 4 == 0:
            Cm[:9, :9] *= 1e8            # information eigenvalues around 1e-6..1e-12: some fall under the 1e-12 clamp
        info, ig, ia = link_information(Cm, oldest=(trial % 2 == 0))
        M = np.linalg.inv(Cm[:9, :9].astype(np.float64))
        w, V = np.linalg.eigh((M + M.T) / 2)
        w[w < 1e-12] = 0
        ref = V @ np.diag(w) @ V.T * (1e-2 if trial % 2 == 0 else 1.0)
        assert np.abs(info - ref).max() <= 1e-9 * np.abs(ref).max()
        assert np.allclose(ig, np.linalg.inv(Cm[9:12, 9:12].astype(np.float64)), rtol=1e-12)
        assert np.allclose(ia, np.linalg.inv(Cm[12:15, 12:15].astype(np.float64)), rtol=1e-12)
    with pytest.raises(Exception):
        link_information(np.zeros((15, 15), np.float32))


def test_threaded_run_on_the_bench_window(mt_binary, tmp_path):
    """Realistic conditioning (preintegrated links: information ~1e8 on rotation, 1e10 on the gyro random walk) under a different
    reduction order (a team of 4 "CTAs" x 4 threads): same Levenberg path, values inside the BA tolerance, and no data race."""
    from orb_slam3_detailed_comments_b200 import synth
    s = synth.inertial_window(seed=5, n_mp=500)
    ref = po.liba(s["state"], s["fixed"], s["point"], s["edge_kf"], s["edge_mp"], s["obs"], s["inv_sigma2"], s["Tcb"], s["cam5"],
                  s["links"].view(po.LIBA_LINK), 1.0, 10)
    got = run_mt(mt_binary, s, 1.0, 10, 4, tmp_path, 4)
    assert got["iterations"] == ref["iterations"] and abs(got["trials"] - ref["trials"]) <= 1
    assert np.abs(got["state"] - ref["state"]).max() < TOL and np.abs(got["point"] - ref["point"]).max() < TOL


def test_preintegration_generator_is_self_consistent():
    """synth.preintegrate (the float32 restatement of IMU::Preintegrated::IntegrateNewMeasurement that feeds the synthetic windows):
    constant rates -> closed forms; bias Jacobians -> finite differences of the deltas with respect to the bias."""
    from orb_slam3_detailed_comments_b200 import synth
    n, dt = 50, 0.005
    w, a = np.array([0.2, -0.1, 0.3]), np.array([0.5, -0.3, 9.6])
    pre = synth.preintegrate(np.tile(a, (n, 1)), np.tile(w, (n, 1)), dt)
    assert np.abs(pre["dR"] - synth._expm(w * n * dt)).max() < 2e-5
    assert abs(pre["dT"] - n * dt) < 1e-6
    eps = 1e-3
    for axis in range(3):
        for kind, sl in (("a", 0), ("g", 3)):
            b = np.zeros(6)
            b[sl + axis] = eps
            p2 = synth.preintegrate(np.tile(a, (n, 1)), np.tile(w, (n, 1)), dt, b)
            JV, JP = pre["JVa" if kind == "a" else "JVg"], pre["JPa" if kind == "a" else "JPg"]
            assert np.abs((p2["dV"] - pre["dV"]) / eps - JV[:, axis]).max() < 5e-3
            assert np.abs((p2["dP"] - pre["dP"]) / eps - JP[:, axis]).max() < 5e-3
            if kind == "g":
                dphi = (pre["dR"].astype(np.float64).T @ p2["dR"].astype(np.float64))
                rot = np.array([dphi[2, 1] - dphi[1, 2], dphi[0, 2] - dphi[2, 0], dphi[1, 0] - dphi[0, 1]]) / 2
                assert np.abs(rot / eps - pre["JRg"][:, axis]).max() < 5e-3
    Cm = pre["C"].astype(np.float64)
    assert np.allclose(Cm, Cm.T, atol=1e-9) and (np.linalg.eigvalsh(Cm[:9, :9]) > 0).all()


def test_depth_sign_and_last_trial_chi2(emul):
    """A point pushed behind one of its cameras: isDepthPositive() is reported per edge at the final estimate; chi2_last is the
    robust chi2 of the last Levenberg trial (what the reference reads as err_end), not necessarily the accepted one."""
    s = perturbed(51, n_kf=5, n_mp=80)
    k = int(s["edge_kf"][0])
    Rwb, twb = s["state"][k, :9].reshape(3, 3), s["state"][k, 9:12]
    Rcb, tcb = np.asarray(s["Tcb"][:9]).reshape(3, 3), np.asarray(s["Tcb"][9:])
    s["point"][int(s["edge_mp"][0])] = Rwb @ (Rcb.T @ (np.array([0.1, 0.1, -2.0]) - tcb)) + twb      # z = -2 in that camera
    ref = po.liba(s["state"], s["fixed"], s["point"], s["edge_kf"], s["edge_mp"], s["obs"], s["inv_sigma2"], s["Tcb"], s["cam5"], s["links"], 1.0, 0)
    got = run_emul(emul, s, 1.0, 0)                 # zero iterations: the estimate stays where it was put
    assert ref["edge_depth_pos"][0] == 0 and (got["edge_depth_pos"] == ref["edge_depth_pos"]).all()
    ref = po.liba(s["state"], s["fixed"], s["point"], s["edge_kf"], s["edge_mp"], s["obs"], s["inv_sigma2"], s["Tcb"], s["cam5"], s["links"], 1.0, 6)
    got = run_emul(emul, s, 1.0, 6)
    assert (got["edge_depth_pos"] == ref["edge_depth_pos"]).all()
    assert abs(got["chi2_last"] - ref["chi2_last"]) <= 1e-6 * max(1.0, ref["chi2_last"]) and ref["chi2_last"] >= ref["chi2"] - 1e-9


def test_large_window_blarge_settings(emul):
    """bLarge (Optimizer.cc:2211-2216, 2347): 25 optimisable keyframes (a 375 x 375 reduced system), lambda 1e-2, 4 iterations."""
    from orb_slam3_detailed_comments_b200 import synth
    s = synth.inertial_window(n_opt=25, n_cov_fixed=5, n_mp=900, seed=9)
    ref = po.liba(s["state"], s["fixed"], s["point"], s["edge_kf"], s["edge_mp"], s["obs"], s["inv_sigma2"], s["Tcb"], s["cam5"],
                  s["links"].view(po.LIBA_LINK), 1e-2, 4)
    got = run_emul(emul, s, 1e-2, 4)
    assert got["iterations"] == ref["iterations"] and got["trials"] == ref["trials"]
    assert np.abs(got["state"] - ref["state"]).max() < TOL and np.abs(got["point"] - ref["point"]).max() < TOL
    assert ref["chi2"] < 0.05 * ref["chi2_init"]


def test_solver_failure_path(emul):
    """A landmark whose only edge carries zero information, with a lambda so small that det(H_ll + lambda I) underflows to 0: the
    block solver fails, g2o's Levenberg raises lambda and retries (optimization_algorithm_levenberg.cpp:96-128).  The device algorithm
    must take the same path as the oracle (failure flag, restored estimate, same trial count) and recover once lambda is large enough."""
    s = perturbed(61, n_kf=5, n_mp=80)
    s["point"] = np.concatenate([s["point"], [[0.3, -0.2, 6.0]]])
    s["edge_kf"] = np.concatenate([s["edge_kf"], [1]]).astype(np.int32)
    s["edge_mp"] = np.concatenate([s["edge_mp"], [len(s["point"]) - 1]]).astype(np.int32)
    s["obs"] = np.concatenate([s["obs"], [[300.0, 200.0, -1.0]]])
    s["inv_sigma2"] = np.concatenate([s["inv_sigma2"], [0.0]])
    for lam in (1e-200, 1e-120):
        ref = po.liba(s["state"], s["fixed"], s["point"], s["edge_kf"], s["edge_mp"], s["obs"], s["inv_sigma2"], s["Tcb"], s["cam5"], s["links"], lam, 6)
        got = run_emul(emul, s, lam, 6)
        assert got["iterations"] == ref["iterations"] and got["trials"] == ref["trials"], lam
        assert np.array_equal(np.isfinite(got["state"]), np.isfinite(ref["state"])) and np.isfinite(ref["state"]).all()
        assert np.abs(got["state"] - ref["state"]).max() < TOL and np.abs(got["point"] - ref["point"]).max() < TOL
        assert ref["trials"] > ref["iterations"]          # at least one failed / rejected trial happened


def test_solver_failure_path_threaded(mt_binary, tmp_path):
    """The failure flags (set by several threads, read by all after a barrier) under ThreadSanitizer, two "CTAs"."""
    s = perturbed(61, n_kf=5, n_mp=80)
    s["point"] = np.concatenate([s["point"], [[0.3, -0.2, 6.0]]])
    s["edge_kf"] = np.concatenate([s["edge_kf"], [1]]).astype(np.int32)
    s["edge_mp"] = np.concatenate([s["edge_mp"], [len(s["point"]) - 1]]).astype(np.int32)
    s["obs"] = np.concatenate([s["obs"], [[300.0, 200.0, -1.0]]])
    s["inv_sigma2"] = np.concatenate([s["inv_sigma2"], [0.0]])
    ref = po.liba(s["state"], s["fixed"], s["point"], s["edge_kf"], s["edge_mp"], s["obs"], s["inv_sigma2"], s["Tcb"], s["cam5"], s["links"], 1e-200, 3)
    got = run_mt(mt_binary, s, 1e-200, 3, 4, tmp_path, 2)
    assert got["iterations"] == ref["iterations"] and got["trials"] == ref["trials"] == 10
    assert np.abs(got["state"] - ref["state"]).max() < 1e-12


def test_window_shape_fuzz(emul):
    """hypothesis: window sizes, which keyframes are fixed (incl. both ends of a link, or all but one), robust flags, mono / stereo mixes,
    dropped links -- the device algorithm follows the oracle's Levenberg path on every draw."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.integers(2, 8), st.integers(12, 90), st.integers(0, 2 ** 31 - 1), st.sampled_from([1.0, 1e-2, 0.0, 100.0]), st.integers(1, 6))
    def run(n_kf, n_mp, seed, lam, iters):
        rng = np.random.default_rng(seed)
        s = perturbed(int(rng.integers(0, 1000)), n_kf=n_kf, n_mp=n_mp)
        if len(s["edge_kf"]) == 0:
            return
        fx = (rng.random(n_kf) < 0.4).astype(np.uint8)
        if fx.all():
            fx[rng.integers(0, n_kf)] = 0
        s["fixed"] = fx
        s["links"] = s["links"].copy()
        s["links"]["robust"] = rng.integers(0, 2, len(s["links"]))
        if len(s["links"]) > 1 and rng.random() < 0.5:
            s["links"] = np.delete(s["links"], rng.integers(0, len(s["links"])))
        s["obs"] = s["obs"].copy()
        s["obs"][rng.random(len(s["obs"])) < 0.3, 2] = -1.0
        ref = po.liba(s["state"], s["fixed"], s["point"], s["edge_kf"], s["edge_mp"], s["obs"], s["inv_sigma2"], s["Tcb"], s["cam5"], s["links"], lam, iters)
        got = run_emul(emul, s, lam, iters)
        assert got["iterations"] == ref["iterations"] and abs(got["trials"] - ref["trials"]) <= 1
        if got["trials"] == ref["trials"]:
            assert np.abs(got["state"] - ref["state"]).max() < TOL and np.abs(got["point"] - ref["point"]).max() < TOL
        assert (got["state"][fx == 1] == s["state"][fx == 1]).all()
    run()


def test_threaded_window_shape_fuzz(mt_binary, tmp_path):
    """The same kind of draws run by T threads x C "CTAs" under ThreadSanitizer: no race on any window shape."""
    rng = np.random.default_rng(123)
    for trial in range(10):
        n_kf, n_mp = int(rng.integers(2, 8)), int(rng.integers(12, 80))
        s = perturbed(int(rng.integers(0, 1000)), n_kf=n_kf, n_mp=n_mp)
        if len(s["edge_kf"]) == 0:
            continue
        fx = (rng.random(n_kf) < 0.4).astype(np.uint8)
        if fx.all():
            fx[rng.integers(0, n_kf)] = 0
        s["fixed"] = fx
        lam = [1.0, 0.0, 1e-2][trial % 3]
        ref = po.liba(s["state"], s["fixed"], s["point"], s["edge_kf"], s["edge_mp"], s["obs"], s["inv_sigma2"], s["Tcb"], s["cam5"], s["links"], lam, 4)
        got = run_mt(mt_binary, s, lam, 4, int(rng.integers(1, 6)), tmp_path, int(rng.choice([1, 2, 4, 8])))
        assert got["iterations"] == ref["iterations"] and abs(got["trials"] - ref["trials"]) <= 1
        if got["trials"] == ref["trials"]:
            assert np.abs(got["state"] - ref["state"]).max() < TOL

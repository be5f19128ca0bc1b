"""GPU tier: liba_solve (LocalInertialBA's numeric core, csrc/liba.cu) against the oracle, through the C ABI.
The kernel's source (csrc/liba_core.cuh) is validated on the CPU by tests/test_liba_emul.py (single-thread run == oracle; N-thread
run under ThreadSanitizer race-free).  First device run: round 2 (compute-sanitizer memcheck + racecheck clean, all cases green)."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from test_liba_emul import TOL, perturbed

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def opt():
    from orb_slam3_detailed_comments_b200.optimizer import InertialOptimizer
    o = InertialOptimizer(0)
    yield o
    o.close()


def oracle(s, lam, iters):
    return po.liba(s["state"], s["fixed"], s["point"], s["edge_kf"], s["edge_mp"], s["obs"], s["inv_sigma2"], s["Tcb"], s["cam5"], s["links"], lam, iters)


def check(got, ref):
    assert got["iterations"] == ref["iterations"] and abs(got["trials"] - ref["trials"]) <= 1
    assert abs(got["chi2"] - ref["chi2"]) <= 1e-5 * max(1.0, ref["chi2"])
    assert np.abs(got["state"] - ref["state"]).max() < TOL and np.abs(got["point"] - ref["point"]).max() < TOL


@pytest.mark.parametrize("cluster", [1, 8, 2, 16])
@pytest.mark.parametrize("seed,lam,iters,n_fixed", [(11, 1.0, 10, 1), (12, 1e-2, 4, 1), (13, 1.0, 10, 3)])
def test_matches_oracle(opt, monkeypatch, seed, lam, iters, n_fixed, cluster):
    monkeypatch.setenv("ORB_LIBA_CLUSTER", str(cluster))      # CTAs per window (liba.cu reads it at every call)
    s = perturbed(seed)
    s["fixed"][:] = 0
    s["fixed"][:n_fixed] = 1
    check(opt.LocalInertialBA(s, lam, iters), oracle(s, lam, iters))


def test_batch_of_windows(opt):
    ss = [perturbed(40 + i, n_kf=5 + i, n_mp=150) for i in range(4)]
    for got, s in zip(opt.LocalInertialBABatch(ss, 1.0, 10), ss):
        check(got, oracle(s, 1.0, 10))


def test_bench_window(opt):
    """The window bench.py times: realistic preintegrated links (information up to 1e10), 27 k edges, a cluster of 8 CTAs."""
    from orb_slam3_detailed_comments_b200 import synth
    s = synth.inertial_window(seed=2)
    ref = po.liba(s["state"], s["fixed"], s["point"], s["edge_kf"], s["edge_mp"], s["obs"], s["inv_sigma2"], s["Tcb"], s["cam5"],
                  s["links"].view(po.LIBA_LINK), 1.0, 10)
    check(opt.LocalInertialBA(s, 1.0, 10), ref)


def test_large_window(opt):
    """bLarge: 25 optimisable keyframes (375 x 375 reduced system in global memory), lambda 1e-2, 4 iterations."""
    from orb_slam3_detailed_comments_b200 import synth
    s = synth.inertial_window(n_opt=25, n_cov_fixed=5, n_mp=900, seed=9)
    ref = po.liba(s["state"], s["fixed"], s["point"], s["edge_kf"], s["edge_mp"], s["obs"], s["inv_sigma2"], s["Tcb"], s["cam5"],
                  s["links"].view(po.LIBA_LINK), 1e-2, 4)
    check(opt.LocalInertialBA(s, 1e-2, 4), ref)

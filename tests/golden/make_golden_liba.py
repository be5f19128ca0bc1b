"""tests/golden/golden_liba.npz: a small LocalInertialBA window (inputs) with the oracle's outputs at generation time -- final states and
points, per-edge and per-link chi2, iteration / trial counts, lambda, chi2 trace ends -- so that a later edit of oracle/lba_oracle.cpp
(orc_liba) or of the generator cannot drift unnoticed.  python tests/golden/make_golden_liba.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import pyoracle as po  # noqa: E402
from orb_slam3_detailed_comments_b200 import synth  # noqa: E402

if __name__ == "__main__":
    s = synth.inertial_window(n_opt=5, n_cov_fixed=2, n_mp=120, seed=17)
    r = po.liba(s["state"], s["fixed"], s["point"], s["edge_kf"], s["edge_mp"], s["obs"], s["inv_sigma2"], s["Tcb"], s["cam5"],
                s["links"].view(po.LIBA_LINK), 1.0, 10)
    np.savez_compressed(os.path.join(HERE, "golden_liba.npz"),
                        in_state=s["state"], in_fixed=s["fixed"], in_point=s["point"], in_edge_kf=s["edge_kf"], in_edge_mp=s["edge_mp"],
                        in_obs=s["obs"], in_inv_sigma2=s["inv_sigma2"], in_Tcb=np.asarray(s["Tcb"]), in_cam5=np.asarray(s["cam5"], np.float64),
                        in_links=np.frombuffer(np.ascontiguousarray(s["links"]).tobytes(), np.uint8),
                        state=r["state"], point=r["point"], edge_chi2=r["edge_chi2"], link_chi2=r["link_chi2"],
                        scalars=np.array([r["iterations"], r["trials"], r["lambda_"], r["chi2"], r["chi2_init"], r["chi2_last"]]),
                        edge_depth_pos=r["edge_depth_pos"])
    print(r["iterations"], r["trials"], r["chi2_init"], r["chi2"], os.path.getsize(os.path.join(HERE, "golden_liba.npz")))

#!/usr/bin/env python3
"""One LocalBundleAdjustment (config 4) and one LocalInertialBA window after a warm-up each: the ncu target for k_lba / k_liba."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam3_detailed_comments_b200 import Optimizer, InertialOptimizer, synth

opt = Optimizer(0)
pr = synth.lba_problem(seed=0)
for _ in range(3):
    g = opt.LocalBundleAdjustment(pr, lambda_init=100.0)
print("lba", g["iterations"], g.get("trials"))
iopt = InertialOptimizer(0)
w = synth.inertial_window(seed=0)
for _ in range(3):
    gi = iopt.LocalInertialBA(w, 1.0, 10)
print("liba", gi["iterations"], gi.get("trials"))

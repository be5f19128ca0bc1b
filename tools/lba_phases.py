import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ORB_LBA_PHASES"] = "1"
from orb_slam3_detailed_comments_b200 import Optimizer, synth
opt = Optimizer(0)
pr = synth.lba_problem(seed=0)
for i in range(2):
    t = time.perf_counter(); g = opt.LocalBundleAdjustment(pr, lambda_init=100.0); print("solve ms", 1e3 * (time.perf_counter() - t), g["iterations"], g["trials"], file=sys.stderr)
prs = [synth.lba_problem(seed=s) for s in range(4)]
t = time.perf_counter(); opt.LocalBundleAdjustmentBatch(prs, lambda_init=100.0); print("batch4 ms", 1e3 * (time.perf_counter() - t), file=sys.stderr)

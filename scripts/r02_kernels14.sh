#!/bin/bash
# launch priority for the latency-bound kernels (ORB_PRIO=1) + column IC_Angle with immediate disc bounds: parity, bench A/B
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_extractor_gpu.py tests/test_zz_replay_step_gpu.py tests/test_zz_host_boundary_gpu.py tests/test_stereo_gpu.py tests/test_matcher_gpu.py -x -q 2>&1 | tail -3 | tee gpurun_out/r02_k14_tests.log
ORB_PRIO=1 timeout 900 python -m pytest tests/test_extractor_gpu.py tests/test_zz_replay_step_gpu.py tests/test_zz_chain_gpu.py tests/test_zz_graph_gpu.py -x -q 2>&1 | tail -3 | tee -a gpurun_out/r02_k14_tests.log
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --e2e-repeats 1 --latency-frames 5 > gpurun_out/r02_k14_bench_$name.json 2> gpurun_out/r02_k14_bench_$name.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_k14_bench_$name.json"))
s = d["roofline"]["stage_ms_per_batch"]
print("$name: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v, 3) for k, v in s.items()}, "parity", d["parity"]["ok"], "chain", round(d["chained_flow"]["e2e_frames_per_s"]))
PY
}
run prio0 ORB_PRIO=0
run prio1 ORB_PRIO=1
run prio0b ORB_PRIO=0
run prio1b ORB_PRIO=1

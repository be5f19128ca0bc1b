#!/bin/bash
# counting-sort k_frame_grid + predicated group scan: parity, then queries-per-CTA sweep
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_matcher_gpu.py tests/test_zz_replay_step_gpu.py tests/test_zz_chain_gpu.py tests/test_zz_host_boundary_gpu.py tests/test_zz_config5_gpu.py -x -q 2>&1 | tail -6 | tee gpurun_out/r02_k11_tests.log
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --e2e-repeats 1 --latency-frames 5 > gpurun_out/r02_k11_bench_$name.json 2> gpurun_out/r02_k11_bench_$name.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_k11_bench_$name.json"))
print("$name: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "stereo+search", round(d["roofline"]["stage_ms_per_batch"]["stereo+search"], 4), "parity", d["parity"]["ok"], "chain", round(d["chained_flow"]["e2e_frames_per_s"]))
PY
}
run default ORB_X=0
run q32 ORB_PROJ_QPB=32
run q64 ORB_PROJ_QPB=64
run q128 ORB_PROJ_QPB=128
run l4q64 ORB_PROJ_LANES=4 ORB_PROJ_QPB=64
run l8q64 ORB_PROJ_LANES=8 ORB_PROJ_QPB=64

#!/bin/bash
# k_proj_candidates_grp<G> (G lanes per query for the two per-frame searches) against the warp-per-query kernel:
# matcher / replay / chain / host-boundary parity with the default lanes, then the bench line per variant.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_matcher_gpu.py tests/test_zz_replay_step_gpu.py tests/test_zz_chain_gpu.py tests/test_zz_host_boundary_gpu.py tests/test_zz_config5_gpu.py -x -q 2>&1 | tail -6 | tee gpurun_out/r02_k10_tests.log
for l in 4 16; do
ORB_PROJ_LANES=$l timeout 600 python -m pytest tests/test_matcher_gpu.py tests/test_zz_replay_step_gpu.py -x -q 2>&1 | tail -2 | tee -a gpurun_out/r02_k10_tests.log
done
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --e2e-repeats 1 --latency-frames 5 > gpurun_out/r02_k10_bench_$name.json 2> gpurun_out/r02_k10_bench_$name.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_k10_bench_$name.json"))
print("$name: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "stereo+search", round(d["roofline"]["stage_ms_per_batch"]["stereo+search"], 4), "parity", d["parity"]["ok"], "chain", round(d["chained_flow"]["e2e_frames_per_s"]))
PY
}
run warp ORB_PROJ_LANES=32
run default ORB_X=0
run l4 ORB_PROJ_LANES=4
run l8 ORB_PROJ_LANES=8
run l16 ORB_PROJ_LANES=16
run default_q128 ORB_PROJ_QPB=128
run default_q512 ORB_PROJ_QPB=512

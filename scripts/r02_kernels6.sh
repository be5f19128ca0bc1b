#!/bin/bash
# k_resize (word loads + funnel shifts), k_fast_cells_v2 at 12 CTAs per SM (alt build)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_extractor_gpu.py tests/test_zz_fast_v2_gpu.py tests/test_stereo_gpu.py tests/test_zz_graph_gpu.py tests/test_zz_replay_step_gpu.py tests/test_liba_gpu.py -x -q 2>&1 | tail -6 | tee gpurun_out/r02_k6_tests.log
ORB_LIB_PATH=$PWD/orb_slam3_detailed_comments_b200/lib/liborbslam3_b200_alt.so timeout 600 python -m pytest tests/test_zz_fast_v2_gpu.py -x -q 2>&1 | tail -3
run() {
  timeout 300 python bench.py --extract-only --no-cpu-baseline --e2e-repeats 1 > gpurun_out/r02_k6_bench_$1.json 2> gpurun_out/r02_k6_bench_$1.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_k6_bench_$1.json"))
print("$1: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(x, 3) for k, x in d["roofline"]["stage_ms_per_batch"].items()})
PY
}
run main
ORB_LIB_PATH=$PWD/orb_slam3_detailed_comments_b200/lib/liborbslam3_b200_alt.so run alt_fast12

#!/bin/bash
# IC_Angle by rows with IDP.4A in k_orient_describe; k_resize_v2 with 8 rows per thread: extractor parity, then the bench lines
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_extractor_gpu.py tests/test_zz_fast_v2_gpu.py tests/test_zz_replay_step_gpu.py tests/test_zz_host_boundary_gpu.py tests/test_zz_config5_gpu.py tests/test_stereo_gpu.py -x -q 2>&1 | tail -6 | tee gpurun_out/r02_k13_tests.log
ORB_RESIZE_VARIANT=2 timeout 900 python -m pytest tests/test_extractor_gpu.py -x -q 2>&1 | tail -2 | tee -a gpurun_out/r02_k13_tests.log
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --e2e-repeats 1 --latency-frames 5 > gpurun_out/r02_k13_bench_$name.json 2> gpurun_out/r02_k13_bench_$name.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_k13_bench_$name.json"))
s = d["roofline"]["stage_ms_per_batch"]
print("$name: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v, 3) for k, v in s.items()}, "parity", d["parity"]["ok"])
PY
}
run default ORB_X=0
run rows8 ORB_RESIZE_VARIANT=2
run default2 ORB_X=0

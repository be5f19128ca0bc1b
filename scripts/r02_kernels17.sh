#!/bin/bash
# k_resize_v3 (source rows by word loads + PRMT + IDP.2A, ORB_RESIZE_VARIANT=3): extractor / replay / host-boundary parity, bench A/B
set -u
mkdir -p gpurun_out
ORB_RESIZE_VARIANT=3 timeout 900 python -m pytest tests/test_extractor_gpu.py tests/test_zz_fast_v2_gpu.py tests/test_zz_replay_step_gpu.py tests/test_zz_host_boundary_gpu.py tests/test_zz_config5_gpu.py tests/test_zz_graph_gpu.py tests/test_stereo_gpu.py -x -q 2>&1 | tail -3 | tee gpurun_out/r02_k17_tests.log
for v in 3 1; do
ORB_RESIZE_VARIANT=$v timeout 600 python bench.py --no-cpu-baseline --e2e-repeats 1 --latency-frames 5 > gpurun_out/r02_k17_bench_$v.json 2> gpurun_out/r02_k17_bench_$v.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_k17_bench_$v.json"))
s = d["roofline"]["stage_ms_per_batch"]
print("resize variant $v: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(x, 3) for k, x in s.items()}, d["parity"]["ok"])
PY
done

#!/bin/bash
# ORB_FAST_TMA=1 (k_fast_cells_v2<29, true>: window rows by cp.async.bulk + mbarrier) against the default staging:
# extractor parity with the variant on, then the bench line of each.
set -u
mkdir -p gpurun_out
ORB_FAST_TMA=1 timeout 900 python -m pytest tests/test_extractor_gpu.py tests/test_zz_fast_v2_gpu.py tests/test_zz_replay_step_gpu.py tests/test_zz_host_boundary_gpu.py -x -q 2>&1 | tail -6 | tee gpurun_out/r02_k9_tests.log
for v in 0 1; do
ORB_FAST_TMA=$v timeout 600 python bench.py --no-cpu-baseline --latency-frames 20 > gpurun_out/r02_k9_bench_tma$v.json 2> gpurun_out/r02_k9_bench_tma$v.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_k9_bench_tma$v.json"))
print("tma$v: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["roofline"]["stage_ms_per_batch"], d["parity"])
PY
done

#!/bin/bash
# packed SAD (VABSDIFF4 with accumulate) in k_stereo_match_v1: stereo parity, then the bench line
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stereo_gpu.py tests/test_zz_stereo_v1_gpu.py tests/test_zz_replay_step_gpu.py tests/test_zz_host_boundary_gpu.py tests/test_zz_config5_gpu.py -x -q 2>&1 | tail -3 | tee gpurun_out/r02_k15_tests.log
timeout 600 python bench.py --no-cpu-baseline --e2e-repeats 1 --latency-frames 20 > gpurun_out/r02_k15_bench.json 2> gpurun_out/r02_k15_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_k15_bench.json"))
s = d["roofline"]["stage_ms_per_batch"]
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v, 3) for k, v in s.items()}, "parity", d["parity"]["ok"], "lat", d["latency_b1"]["graph_replay_device_resident"]["p50_ms"], d["latency_b1"]["e2e_host_buffers"]["p50_ms"])
PY

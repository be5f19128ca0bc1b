#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_quadtree_v1_gpu.py tests/test_extractor_gpu.py tests/test_zz_config5_gpu.py tests/test_zz_graph_gpu.py -x -q 2>&1 | tail -6 | tee gpurun_out/r02_k7_tests.log
timeout 300 python tools/qt_phases.py 2>&1 | grep -E "level|sweeps|std::sort" | tee gpurun_out/r02_qt_phases5.txt
timeout 300 python bench.py --extract-only --no-cpu-baseline --e2e-repeats 1 > gpurun_out/r02_k7_bench.json 2> gpurun_out/r02_k7_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_k7_bench.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(x, 3) for k, x in d["roofline"]["stage_ms_per_batch"].items()})
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_lba|k_liba|k_hamming|k_pose_opt|k_pose_edges|k_in_frustum|k_chain|k_pack" -c 200 --csv --log-file gpurun_out/r02_launches_side.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-repeats 1 --latency-frames 5 > gpurun_out/r02_launches_side.log 2>&1
echo "side launches rc=$?"; wc -l gpurun_out/r02_launches_side.csv

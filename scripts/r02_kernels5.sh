#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_quadtree_v1_gpu.py tests/test_zz_fast_v2_gpu.py tests/test_extractor_gpu.py tests/test_zz_graph_gpu.py -x -q 2>&1 | tail -6 | tee gpurun_out/r02_k5_tests.log
timeout 300 python tools/qt_phases.py > gpurun_out/r02_qt_phases4.txt 2>&1
timeout 300 python tools/qt_phases.py --sigma 3.5 --nrect 40 >> gpurun_out/r02_qt_phases4.txt 2>&1
grep -E "level|std::sort|bitonic" gpurun_out/r02_qt_phases4.txt
timeout 300 python bench.py --extract-only --no-cpu-baseline --e2e-repeats 1 > gpurun_out/r02_k5_bench.json 2> gpurun_out/r02_k5_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_k5_bench.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(x, 3) for k, x in d["roofline"]["stage_ms_per_batch"].items()})
PY

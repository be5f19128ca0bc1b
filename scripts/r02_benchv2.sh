#!/bin/bash
# bench v2 first contact: replay-step tests, then the bench line.   gpurun --timeout 1500 -- 'bash scripts/r02_benchv2.sh'
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_replay_step_gpu.py tests/test_zz_graph_gpu.py -x -q 2>&1 | tail -25 > gpurun_out/r02_replay_tests.log
tail -5 gpurun_out/r02_replay_tests.log
timeout 900 python bench.py > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err
echo "bench rc=$?"; tail -5 gpurun_out/r02_bench_b.err; head -c 900 gpurun_out/r02_bench_b.json

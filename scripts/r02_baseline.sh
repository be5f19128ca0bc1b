#!/bin/bash
# round-2 baseline: gpu tests, bench line, ncu launch list.   gpurun --timeout 1500 -- 'bash scripts/r02_baseline.sh'
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_pytest.log
timeout 600 python bench.py --steps 40 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --no-cpu-baseline --steps 4 --warmup 4 > gpurun_out/r02_launches.log 2>&1
tail -3 gpurun_out/r02_pytest.log; head -c 600 gpurun_out/r02_bench_a.json

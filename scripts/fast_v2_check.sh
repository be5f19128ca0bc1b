#!/bin/bash
# k_fast_cells_v2: parity, per-stage effect, instruction counts.   gpurun --timeout 900 -- 'bash scripts/fast_v2_check.sh'
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_fast_v2_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/fast_v2_tests.log
ORB_FAST_VARIANT=1 timeout 600 python -m pytest tests/test_extractor_gpu.py tests/test_stereo_gpu.py -x -q 2>&1 | tail -5 | tee -a gpurun_out/fast_v2_tests.log
for v in 1; do
  ORB_FAST_VARIANT=$v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --extract-only 2>gpurun_out/fast_v${v}_bench.err | tee gpurun_out/fast_v${v}_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('variant $v', d['value'], d['roofline']['stage_ms_per_step'])"
done
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,smsp__average_warp_latency_issue_stalled_barrier.ratio,smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_wait.ratio,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_mio_throttle.ratio,smsp__average_warp_latency_issue_stalled_not_selected.ratio,smsp__thread_inst_executed.sum
ORB_FAST_VARIANT=1 timeout 600 ncu --metrics $M --clock-control none -k regex:k_fast_cells -c 3 --csv --log-file gpurun_out/fast_v1_ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --extract-only > gpurun_out/fast_v1_ncu.log 2>&1

#!/bin/bash
# ncu --set full with source for the kernels of the stereo + search stage.   gpurun --timeout 900 -- 'bash scripts/r02_ncu_search.sh'
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --e2e-repeats 1 --latency-frames 2"
timeout 700 ncu --set full --import-source on --clock-control none -k regex:"k_proj_candidates|k_frame_grid|k_proj_resolve_par|k_stereo_match_v1|k_stereo_median" --launch-skip 16 -c 10 -o gpurun_out/r02_search -f $B > gpurun_out/r02_search.log 2>&1
echo "search rc=$?"; ls -la gpurun_out/r02_search.ncu-rep

#!/bin/bash
# ncu --set full with source for the kernels that changed on the last day of round 2.   gpurun --timeout 900 -- 'bash scripts/r02_ncu_final.sh'
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --e2e-repeats 1 --latency-frames 2"
timeout 800 ncu --set full --import-source on --clock-control none -k regex:"k_resize_v2|k_blur|k_orient_describe|k_stereo_match_v1|k_proj_candidates_grp|k_frame_grid|k_proj_resolve_par" --launch-skip 60 -c 16 -o gpurun_out/r02_final_kernels -f $B > gpurun_out/r02_final_kernels.log 2>&1
echo "rc=$?"; ls -la gpurun_out/r02_final_kernels.ncu-rep

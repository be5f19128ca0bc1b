#!/bin/bash
# ncu --set full with source for the kernels VERDICT names.   gpurun --timeout 1700 -- 'bash scripts/r02_ncu_full.sh'
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --extract-only --no-graph --no-cpu-baseline --e2e-repeats 1"
timeout 700 ncu --set full --import-source on --clock-control none -k regex:"k_fast_cells_v2|k_quadtree_v1" --launch-skip 40 -c 4 -o gpurun_out/r02_fq -f $B > gpurun_out/r02_fq.log 2>&1
echo "fq rc=$?"; ls -la gpurun_out/r02_fq.ncu-rep
timeout 500 ncu --set full --import-source on --clock-control none -k regex:"k_lba|k_liba" --launch-skip 4 -c 2 -o gpurun_out/r02_ba -f python tools/run_ba_once.py > gpurun_out/r02_ba.log 2>&1
echo "ba rc=$?"; ls -la gpurun_out/r02_ba.ncu-rep
timeout 500 ncu --set full --import-source on --clock-control none -k regex:"k_blur|k_orient_describe|k_resize|k_stereo_match_v1|k_proj_candidates" --launch-skip 60 -c 12 -o gpurun_out/r02_rest -f $B > gpurun_out/r02_rest.log 2>&1
echo "rest rc=$?"; ls -la gpurun_out/r02_rest.ncu-rep
# launch list of the graph-replay bench (kernel nodes of the graphs are profiled individually)
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 2500 --csv --log-file gpurun_out/r02_launches_v2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-repeats 1 --latency-frames 5 > gpurun_out/r02_launches_v2.log 2>&1
echo "launches rc=$?"; wc -l gpurun_out/r02_launches_v2.csv

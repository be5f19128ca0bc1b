#!/bin/bash
set -u
mkdir -p gpurun_out
ORB_FAST_VARIANT=1 timeout 800 ncu --set full --import-source on --clock-control none -k regex:k_fast_cells_v2 -s 2 -c 1 -o gpurun_out/fast_v2_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --extract-only > gpurun_out/fast_v2_full.log 2>&1
ls -la gpurun_out/fast_v2_full.ncu-rep

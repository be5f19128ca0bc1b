#!/bin/bash
# k_lba: 16-CTA clusters + pose-blocked triangular solve (parity, then phase clocks: 16 vs 8 CTAs, 256 vs 512 threads)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lba_gpu.py tests/test_zz_host_boundary_gpu.py -x -q 2>&1 | tail -8 | tee gpurun_out/r02_lba_tests.log
{
echo "== cluster 16, 256 threads"; timeout 200 python tools/lba_phases.py 2>&1 | tail -4
echo "== cluster 8, 256 threads"; ORB_LBA_CLUSTER=8 timeout 200 python tools/lba_phases.py 2>&1 | tail -4
echo "== cluster 16, 512 threads"; ORB_LIB_PATH=$PWD/orb_slam3_detailed_comments_b200/lib/liborbslam3_b200_lba512.so timeout 200 python tools/lba_phases.py 2>&1 | tail -4
echo "== cluster 16, 512 threads: parity"; ORB_LIB_PATH=$PWD/orb_slam3_detailed_comments_b200/lib/liborbslam3_b200_lba512.so timeout 300 python -m pytest tests/test_lba_gpu.py -x -q 2>&1 | tail -3
} > gpurun_out/r02_lba_phases2.txt 2>&1
cat gpurun_out/r02_lba_phases2.txt

#!/bin/bash
# First device run of k_liba (LocalInertialBA).  Run through gpurun from the repo root:
#   gpurun --timeout 900 -- 'bash scripts/first_contact.sh'
# Writes gpurun_out/liba_*.  The kernel is CPU-validated (tests/test_liba_emul.py); until this script has been green once the GPU
# test and the bench section stay opt-in behind ORB_LIBA_GPU=1.
set -u
mkdir -p gpurun_out
export ORB_LIBA_GPU=1
export ORB_FIRST_CONTACT=1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_liba_gpu.py -x -q -k "matches_oracle and 11-1.0" > gpurun_out/liba_memcheck.log 2>&1
echo "memcheck exit $?" | tee -a gpurun_out/liba_memcheck.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_liba_gpu.py -x -q -k "matches_oracle and 12-0.01" > gpurun_out/liba_racecheck.log 2>&1
echo "racecheck exit $?" | tee -a gpurun_out/liba_racecheck.log
timeout 600 python -m pytest tests/test_liba_gpu.py -x -q 2>&1 | tee gpurun_out/liba_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/liba_bench.err | tee gpurun_out/liba_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_liba -c 20 --csv --log-file gpurun_out/liba_launches.csv \
    python -m pytest tests/test_liba_gpu.py -x -q -k batch > gpurun_out/liba_ncu.log 2>&1
# k_quadtree_v1 (CTA-parallel ordered-phase sort): parity, then the per-stage effect
timeout 600 python -m pytest tests/test_zz_quadtree_v1_gpu.py -x -q 2>&1 | tee gpurun_out/qt_v1_tests.log
ORB_QT_VARIANT=1 timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/qt_v1_bench.err | tee gpurun_out/qt_v1_bench.json
# K9 brute-force Hamming 2-NN (orbm_hamming_knn2)
timeout 300 python -m pytest tests/test_zz_knn_gpu.py -x -q 2>&1 | tee gpurun_out/knn_tests.log

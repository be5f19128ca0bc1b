#!/bin/bash
# First device run of the kernels written after round 1's GPU budget was spent.  From the repo root:
#   gpurun --timeout 1500 -- 'bash scripts/first_contact.sh'
# Every step runs in its own process (a faulting kernel poisons only its own CUDA context) and logs under gpurun_out/.
# The kernels are CPU-validated (tests/test_liba_emul.py, tests/test_quadtree_emul.py, tests/test_knn_cpu.py); until a step has been
# green its GPU tests / bench section stay opt-in behind ORB_FIRST_CONTACT=1.
set -u
mkdir -p gpurun_out
export ORB_FIRST_CONTACT=1
# 1. k_quadtree_v1 (multi-stage bitonic passes + CTA-parallel ordered-phase sort): parity, then the per-stage effect on the headline step
timeout 600 python -m pytest tests/test_zz_quadtree_v1_gpu.py -x -q 2>&1 | tee gpurun_out/qt_v1_tests.log
ORB_QT_VARIANT=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/qt_v1_bench.err | tee gpurun_out/qt_v1_bench.json
# 1b. k_stereo_match_v1 (thread per left keypoint, row buckets): parity, then the bench with both variants on
timeout 600 python -m pytest tests/test_zz_stereo_v1_gpu.py -x -q 2>&1 | tee gpurun_out/stereo_v1_tests.log
ORB_QT_VARIANT=1 ORB_STEREO_VARIANT=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/v1_bench.err | tee gpurun_out/v1_bench.json
# 2. K9 brute-force Hamming 2-NN (orbm_hamming_knn2)
timeout 300 python -m pytest tests/test_zz_knn_gpu.py -x -q 2>&1 | tee gpurun_out/knn_tests.log
# 3. k_liba (LocalInertialBA): sanitizers on one small window, parity at 1 / 2 / 8 CTAs per window, timing, launch list
export ORB_LIBA_GPU=1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_liba_gpu.py -x -q -k "matches_oracle and 11-1.0" > gpurun_out/liba_memcheck.log 2>&1
echo "memcheck exit $?" | tee -a gpurun_out/liba_memcheck.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_liba_gpu.py -x -q -k "matches_oracle and 12-0.01" > gpurun_out/liba_racecheck.log 2>&1
echo "racecheck exit $?" | tee -a gpurun_out/liba_racecheck.log
timeout 600 python -m pytest tests/test_liba_gpu.py -x -q 2>&1 | tee gpurun_out/liba_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/liba_bench.err | tee gpurun_out/liba_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_liba -c 20 --csv --log-file gpurun_out/liba_launches.csv \
    python -m pytest tests/test_liba_gpu.py -x -q -k batch > gpurun_out/liba_ncu.log 2>&1

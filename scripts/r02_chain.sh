#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_chain_gpu.py -x -q 2>&1 | tail -30 > gpurun_out/r02_chain_test.log
cat gpurun_out/r02_chain_test.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02_pytest3.log
cat gpurun_out/r02_pytest3.log

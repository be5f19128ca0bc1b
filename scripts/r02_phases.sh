#!/bin/bash
# full GPU tier after the PredictScale fix, then the phase clocks of k_quadtree_v1 (profile build) and k_lba.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r02_pytest2.log
tail -4 gpurun_out/r02_pytest2.log
timeout 300 python tools/qt_phases.py > gpurun_out/r02_qt_phases.txt 2>&1
timeout 300 python tools/qt_phases.py --sigma 3.5 --nrect 40 >> gpurun_out/r02_qt_phases.txt 2>&1
cat gpurun_out/r02_qt_phases.txt
timeout 300 python tools/lba_phases.py > gpurun_out/r02_lba_phases.txt 2>&1
cat gpurun_out/r02_lba_phases.txt

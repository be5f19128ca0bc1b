#!/bin/bash
set -u
mkdir -p gpurun_out
{ for c in 1 2 4 8 16; do echo "== cluster $c"; ORB_LBA_CLUSTER=$c timeout 200 python tools/lba_phases.py 2>&1 | grep -E "lba 0|solve ms" | tail -2; done; } > gpurun_out/r02_lba_scale.txt 2>&1
cat gpurun_out/r02_lba_scale.txt

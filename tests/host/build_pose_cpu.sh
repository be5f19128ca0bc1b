#!/bin/bash
# Builds tests/host/pose_cpu_mine (CPU tier): Optimizer::PoseOptimization through host/Optimizer_pose_b200.cc with orbo_pose_optimization
# answered by the oracle.  The reference's own function for the comparison is oracle/_ref part 5 (make -C oracle ref5).  Needs the
# reference checkout (its Optimizer.h): build container only.
set -e
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
REF="${ORB_REFERENCE_ROOT:-/root/reference}"
H="$ROOT/orb_slam3_detailed_comments_b200/host"
T="$ROOT/tests/host"
[ -f "$REF/include/Optimizer.h" ] || { echo "reference checkout not present: $REF" >&2; exit 3; }
make -C "$ROOT/oracle" -s
g++ -std=c++14 -O1 -DORB_REFSHIM_POSE -Wall -Wno-unused-function -Wno-comment -include "$H/refshim/ref_skeleton.h" -I "$H/refshim" -I "$REF/include" -I "$REF" \
    -I "$ROOT/include" -I "$H" "$T/pose_cpu.cc" "$H/Optimizer_pose_b200.cc" "$T/pose_stub.cc" -L "$ROOT/oracle/_build" -lorb_oracle \
    -Wl,-rpath,"$ROOT/oracle/_build" -lpthread -o "$T/pose_cpu_mine"

#!/bin/bash
# Builds tests/host/voc_cpu_mine (CPU tier): Frame::ComputeBoW through host/Frame_bow_b200.cc with orbv_* answered by the oracle, over a mock
# vocabulary object whose nodes are protected members like DBoW2's.  No reference header is involved (Frame.h is the skeleton's).
set -e
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
REF="${ORB_REFERENCE_ROOT:-/root/reference}"
H="$ROOT/orb_slam3_detailed_comments_b200/host"
T="$ROOT/tests/host"
[ -f "$REF/include/ORBextractor.h" ] || { echo "reference checkout not present: $REF" >&2; exit 3; }
make -C "$ROOT/oracle" -s
mkdir -p "$T/_gen/voc_inc"
: > "$T/_gen/voc_inc/Frame.h"     # `#include "Frame.h"` resolves to the pre-included skeleton (its guard is already defined)
g++ -std=c++14 -O1 -DORB_REFSHIM_VOC -Wall -Wno-unused-function -Wno-comment -include "$H/refshim/ref_skeleton.h" -I "$T/_gen/voc_inc" -I "$H/refshim" -I "$REF/include" \
    -I "$ROOT/include" -I "$H" "$T/voc_cpu.cc" "$H/Frame_bow_b200.cc" "$T/voc_stub.cc" -L "$ROOT/oracle/_build" -lorb_oracle \
    -Wl,-rpath,"$ROOT/oracle/_build" -lpthread -o "$T/voc_cpu_mine"

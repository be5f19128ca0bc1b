#!/bin/bash
# Builds tests/host/bow_cpu_mine (CPU tier): both ORBmatcher::SearchByBoW overloads and SearchForTriangulation through host/ORBmatcher_bow_b200.cc / ORBmatcher_triangulation_b200.cc (+ ORBmatcher_b200.cc
# for the constructor) with the searches answered by the oracle.  The reference's own functions for the comparison are oracle/_ref part 2.
# Needs the reference checkout (its ORBmatcher.h): build container only.
set -e
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
REF="${ORB_REFERENCE_ROOT:-/root/reference}"
H="$ROOT/orb_slam3_detailed_comments_b200/host"
T="$ROOT/tests/host"
[ -f "$REF/include/ORBmatcher.h" ] || { echo "reference checkout not present: $REF" >&2; exit 3; }
make -C "$ROOT/oracle" -s
g++ -std=c++14 -O1 -DORB_REFSHIM_BOW -DORB_REFSHIM_TRI -Wall -Wno-unused-function -Wno-comment -include "$H/refshim/ref_skeleton.h" -I "$H/refshim" -I "$REF/include" -I "$REF" \
    -I "$ROOT/include" -I "$H" "$T/bow_cpu.cc" "$H/ORBmatcher_bow_b200.cc" "$H/ORBmatcher_triangulation_b200.cc" "$H/ORBmatcher_init_b200.cc" "$T/bow_stub.cc" "$T/bow_ctor.cc" -L "$ROOT/oracle/_build" -lorb_oracle \
    -Wl,-rpath,"$ROOT/oracle/_build" -lpthread -o "$T/bow_cpu_mine"

#!/bin/bash
# Builds tests/host/fuse_cpu_mine (CPU tier): ORBmatcher::Fuse(pKF, vpMapPoints, th) through host/ORBmatcher_fuse_b200.cc with the search answered by
# the oracle.  The reference's own function for the comparison is oracle/_ref part 2.  Needs the reference checkout: build container only.
set -e
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
REF="${ORB_REFERENCE_ROOT:-/root/reference}"
H="$ROOT/orb_slam3_detailed_comments_b200/host"
T="$ROOT/tests/host"
[ -f "$REF/include/ORBmatcher.h" ] || { echo "reference checkout not present: $REF" >&2; exit 3; }
make -C "$ROOT/oracle" -s
g++ -std=c++14 -O1 -DORB_REFSHIM_FUSE -Wall -Wno-unused-function -Wno-comment -Wno-reorder -include "$H/refshim/ref_skeleton.h" -I "$H/refshim" -I "$REF/include" -I "$REF" \
    -I "$ROOT/include" -I "$H" "$T/fuse_cpu.cc" "$H/ORBmatcher_fuse_b200.cc" "$H/ORBmatcher_reloc_b200.cc" "$H/ORBmatcher_sim3_b200.cc" "$T/fuse_stub.cc" "$T/bow_ctor.cc" -L "$ROOT/oracle/_build" -lorb_oracle \
    -Wl,-rpath,"$ROOT/oracle/_build" -lpthread -o "$T/fuse_cpu_mine"

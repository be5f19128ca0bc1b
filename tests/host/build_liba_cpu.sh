#!/bin/bash
# Builds tests/host/liba_cpu_mine and tests/host/liba_cpu_ref (CPU tier): Optimizer::LocalInertialBA over the mock map, once through
# host/Optimizer_liba_b200.cc with liba_solve answered by the oracle, once through the REFERENCE's own function cut out of
# $REF/src/Optimizer.cc into tests/host/_gen/ -- a build directory, git-ignored -- and compiled verbatim over graph stand-ins.
# liba_link_information (host-only helper of the product library) is linked from liborbslam3_b200.so in both.  Needs the reference
# checkout: build container only.
set -e
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
REF="${ORB_REFERENCE_ROOT:-/root/reference}"
H="$ROOT/orb_slam3_detailed_comments_b200/host"
T="$ROOT/tests/host"
[ -f "$REF/src/Optimizer.cc" ] || { echo "reference checkout not present: $REF" >&2; exit 3; }
make -C "$ROOT/oracle" -s
mkdir -p "$T/_gen"
CXXF="-std=c++14 -O1 -DORB_REFSHIM_LIBA -DORB_REFSHIM_REF_LBA -Wall -Wno-unused-function -Wno-comment -Wno-unused-variable -Wno-unused-but-set-variable -Wno-sign-compare -include $H/refshim/ref_skeleton.h -I $H/refshim -I $REF/include -I $REF -I $ROOT/include -I $H -I $T"
LD="-L $ROOT/oracle/_build -lorb_oracle -Wl,-rpath,$ROOT/oracle/_build -L $ROOT/orb_slam3_detailed_comments_b200/lib -lorbslam3_b200 -Wl,-rpath,$ROOT/orb_slam3_detailed_comments_b200/lib -lpthread"
g++ $CXXF "$T/liba_cpu.cc" "$H/Optimizer_liba_b200.cc" "$T/liba_stub.cc" $LD -o "$T/liba_cpu_mine"
if [ -f "$T/liba_ref.cc" ]; then
    python3 "$ROOT/oracle/tools/extract_functions.py" "$REF/src/Optimizer.cc" "$T/_gen/liba_opt.inc" "void Optimizer::LocalInertialBA("
    g++ $CXXF "$T/liba_cpu.cc" "$T/liba_ref.cc" "$T/liba_stub.cc" $LD -o "$T/liba_cpu_ref"
fi

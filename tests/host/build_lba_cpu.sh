#!/bin/bash
# Builds tests/host/lba_cpu_mine and tests/host/lba_cpu_ref (CPU tier): Optimizer::LocalBundleAdjustment over the mock map, once
# through host/Optimizer_lba_b200.cc with lba_solve_bool answered by the oracle, once through the REFERENCE's own function cut out
# of $REF/src/Optimizer.cc (with g2o's Levenberg functions out of $REF/Thirdparty/g2o) into tests/host/_gen/ -- a build directory,
# git-ignored -- and compiled verbatim over graph stand-ins.  Needs the reference checkout: build container only.
set -e
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
REF="${ORB_REFERENCE_ROOT:-/root/reference}"
H="$ROOT/orb_slam3_detailed_comments_b200/host"
T="$ROOT/tests/host"
G2O="$REF/Thirdparty/g2o/g2o/core"
[ -f "$REF/src/Optimizer.cc" ] || { echo "reference checkout not present: $REF" >&2; exit 3; }
make -C "$ROOT/oracle" -s
mkdir -p "$T/_gen"
X="$ROOT/oracle/tools/extract_functions.py"
python3 "$X" "$REF/src/Optimizer.cc" "$T/_gen/lba_opt.inc" "void Optimizer::LocalBundleAdjustment(KeyFrame *pKF, bool *pbStopFlag, Map *pMap, int &num_fixedKF"
python3 "$X" "$G2O/optimization_algorithm_levenberg.cpp" "$T/_gen/g1.inc" "OptimizationAlgorithmLevenberg::OptimizationAlgorithmLevenberg(Solver* solver)" \
    "OptimizationAlgorithm::SolverResult OptimizationAlgorithmLevenberg::solve(" "double OptimizationAlgorithmLevenberg::computeLambdaInit() const" \
    "double OptimizationAlgorithmLevenberg::computeScale() const"
python3 "$X" "$G2O/sparse_optimizer.cpp" "$T/_gen/g2.inc" "int SparseOptimizer::optimize(int iterations, bool online)"
python3 "$X" "$G2O/robust_kernel_impl.cpp" "$T/_gen/g3.inc" "void RobustKernelHuber::setDelta(double delta)"
cat "$T/_gen/g1.inc" "$T/_gen/g2.inc" "$T/_gen/g3.inc" > "$T/_gen/lba_g2o.inc"
CXXF="-std=c++14 -O1 -DORB_REFSHIM_REF_LBA -Wall -Wno-unused-function -Wno-comment -Wno-unused-variable -Wno-unused-but-set-variable -include $H/refshim/ref_skeleton.h -I $H/refshim -I $REF/include -I $REF -I $ROOT/include -I $H -I $T"
LD="-L $ROOT/oracle/_build -lorb_oracle -Wl,-rpath,$ROOT/oracle/_build -lpthread"
g++ $CXXF "$T/lba_cpu.cc" "$H/Optimizer_lba_b200.cc" "$T/lba_stub.cc" $LD -o "$T/lba_cpu_mine"
g++ $CXXF "$T/lba_cpu.cc" "$T/lba_ref.cc" $LD -o "$T/lba_cpu_ref"

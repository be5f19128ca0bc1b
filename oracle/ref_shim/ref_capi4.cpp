// oracle/ref_shim/ref_capi4.cpp -- TEST INFRASTRUCTURE.
//
// g2o's own Levenberg-Marquardt control flow (OptimizationAlgorithmLevenberg::solve / computeLambdaInit / computeScale, its
// constructor's constants, SparseOptimizer::optimize) and RobustKernelHuber, cut out of /root/reference/Thirdparty/g2o at build time
// (oracle/tools/extract_functions.py -> oracle/_ref/gen/ref4_gen.inc, never committed) and compiled verbatim over g2o_skel.h, whose
// Solver / SparseOptimizer forward to the oracle's LbaEngine.  ref4_lba == Optimizer::LocalBundleAdjustment's optimisation call
// (Optimizer.cc:1859-1876, 2100: setUserLambdaInit(100) when a stop flag is given, optimize(10)) with g2o deciding every step.
#include "g2o_skel/g2o_skel.h"

using namespace std;

namespace g2o {
G2OBatchStatistics* G2OBatchStatistics::_g = nullptr;
#include "../_ref/gen/ref4_gen.inc"
}  // namespace g2o

extern "C" {
void* orc_lba_engine_create(int nKF, int nMP, int nE, const double* pose, const uint8_t* fixed, const double* point, const int* ekf, const int* emp,
                            const double* obs, const double* invs2, const double* cam5);
void orc_lba_engine_destroy(void* h);
void orc_lba_engine_finish(void* h, double* pose, double* point, double* edge_chi2, uint8_t* edge_depth_pos);

// same contract as orc_lba (oracle/lba_oracle.cpp); stats = {outer iterations, final lambda, final robust chi2, total LM trials, -}
int ref4_lba(int nKF, int nMP, int nE, double* pose, const uint8_t* fixed, double* point, const int* ekf, const int* emp, const double* obs,
             const double* invs2, const double* cam5, double lambdaInit, int maxIters, double* edge_chi2, uint8_t* edge_depth_pos, double* stats) {
    void* E = orc_lba_engine_create(nKF, nMP, nE, pose, fixed, point, ekf, emp, obs, invs2, cam5);
    int iters;
    {
        g2o::Solver solver(E);
        g2o::SparseOptimizer optimizer(E);
        solver._optimizer = &optimizer;
        struct Counting : public g2o::OptimizationAlgorithmLevenberg {   // sums the trials of every solve(), nothing else
            explicit Counting(g2o::Solver* s) : g2o::OptimizationAlgorithmLevenberg(s) {}
            virtual SolverResult solve(int iteration, bool online = false) {
                const SolverResult r = g2o::OptimizationAlgorithmLevenberg::solve(iteration, online);
                totalTrials += levenbergIteration();
                return r;
            }
        } alg(&solver);
        if (lambdaInit > 0) alg.setUserLambdaInit(lambdaInit);
        optimizer.setAlgorithm(&alg);
        iters = optimizer.optimize(maxIters);
        if (stats) { stats[0] = optimizer.outerIterations; stats[1] = alg.currentLambda(); stats[2] = optimizer.activeRobustChi2(); stats[3] = alg.totalTrials; stats[4] = 0; }
    }
    orc_lba_engine_finish(E, pose, point, edge_chi2, edge_depth_pos);
    orc_lba_engine_destroy(E);
    return iters;
}

// RobustKernelHuber as Optimizer.cc sets it up: rk->setDelta(thHuberMono / thHuberStereo) with `const float` thresholds
void ref4_huber(double e, float delta_f, double* rho3) {
    g2o::RobustKernelHuber rk;
    rk.setDelta(delta_f);
    Eigen::Vector3d rho;
    rk.robustify(e, rho);
    rho3[0] = rho[0]; rho3[1] = rho[1]; rho3[2] = rho[2];
}
}  // extern "C"

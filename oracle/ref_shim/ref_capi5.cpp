// oracle/ref_shim/ref_capi5.cpp -- TEST INFRASTRUCTURE.
//
// The reference's own Optimizer::PoseOptimization(Frame*) (src/Optimizer.cc:55-401), cut out at build time and compiled VERBATIM over
// the stand-ins of g2o_skel/po_skel.h, together with g2o's Levenberg control flow and Huber::setDelta (as oracle/_ref part 4).  The
// numerics (edge errors, the 6 x 6 system, the pose update) are the oracle's PoseEngine; graph construction, the four rounds, the
// float chi2 classification, the levels, the kernel removal and the return value are the reference's code.
#include "g2o_skel/po_skel.h"

using namespace std;

namespace ORB_SLAM3 {
class GeometricCamera {
   public:
    virtual ~GeometricCamera() {}
};
class MapPoint {
   public:
    static std::mutex mGlobalMutex;
    Eigen::Vector3f GetWorldPos() { return mWorldPos; }
    Eigen::Vector3f mWorldPos;
};
std::mutex MapPoint::mGlobalMutex;

class Frame {
   public:
    Sophus::SE3<float> GetPose() const { return mTcw; }
    void SetPose(const Sophus::SE3<float>& T) { mTcw = T; }
    Sophus::SE3<float> GetRelativePoseTrl() const { return Sophus::SE3<float>(); }   // two-camera rigs: not exercised
    int N = 0, Nleft = -1;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<float> mvuRight, mvInvLevelSigma2;
    std::vector<bool> mvbOutlier;
    std::vector<cv::KeyPoint> mvKeysUn, mvKeys, mvKeysRight;
    GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0;
    Sophus::SE3<float> mTcw;
};

// the edge classes of include/OptimizableTypes.h / g2o types_six_dof_expmap.h as far as PoseOptimization touches them
class EdgeSE3ProjectXYZOnlyPose : public g2o::OptimizableGraph::Edge {
   public:
    int dim() const { return 2; }
    void setMeasurement(const Eigen::Matrix<double, 2, 1>& m) { meas[0] = m(0); meas[1] = m(1); meas[2] = -1.0; }
    void setInformation(const Eigen::Matrix2d& I) { assert(I(0, 0) == I(1, 1) && I(0, 1) == 0); info00 = I(0, 0); }
    GeometricCamera* pCamera = nullptr;
};
class EdgeSE3ProjectXYZOnlyPoseToBody : public EdgeSE3ProjectXYZOnlyPose {
   public:
    g2o::SE3Quat mTrl;
};
}  // namespace ORB_SLAM3

namespace g2o {
class EdgeStereoSE3ProjectXYZOnlyPose : public OptimizableGraph::Edge {
   public:
    int dim() const { return 3; }
    void setMeasurement(const Eigen::Matrix<double, 3, 1>& m) { meas[0] = m(0); meas[1] = m(1); meas[2] = m(2); }
    void setInformation(const Eigen::Matrix3d& I) { assert(I(0, 0) == I(1, 1) && I(1, 1) == I(2, 2) && I(0, 1) == 0); info00 = I(0, 0); }
    double fx = 0, fy = 0, cx = 0, cy = 0, bf = 0;
};

const double* SparseOptimizer::g_cam5 = nullptr;
int SparseOptimizer::g_iters = 0, SparseOptimizer::g_trials = 0;
void SparseOptimizer::postIteration(int) {
    ++g_iters;
    g_trials += static_cast<OptimizationAlgorithmLevenberg*>(_algorithm)->levenbergIteration();
}

double OptimizableGraph::Vertex::hessian(int i, int j) const { assert(i == j); return orc_po_engine_hessian_diag(_opt->_e, j); }
void OptimizableGraph::Edge::setRobustKernel(RobustKernel* rk) {
    if (rk) { _rk = rk; return; }
    robustCleared = true;
    if (_opt && _opt->_e) orc_po_engine_set_robust(_opt->_e, index, 0);
}
void OptimizableGraph::Edge::setLevel(int l) {
    _level = l;
    if (_opt && _opt->_e) orc_po_engine_set_level(_opt->_e, index, l);
}
void OptimizableGraph::Edge::computeError() { orc_po_engine_compute_error(_opt->_e, index); }
double OptimizableGraph::Edge::chi2() const { return orc_po_engine_chi2(_opt->_e, index); }

void VertexSE3Expmap::setEstimate(const SE3Quat& T) {
    _est = T;
    if (_opt && _opt->_e) {
        const double p[7] = {T.q.x, T.q.y, T.q.z, T.q.w, T.t(0), T.t(1), T.t(2)};
        orc_po_engine_set_estimate(_opt->_e, p);
    }
}
SE3Quat VertexSE3Expmap::estimate() const {
    if (!(_opt && _opt->_e)) return _est;
    double p[7];
    orc_po_engine_get_estimate(_opt->_e, p);
    SE3Quat T;
    T.q = Eigen::Quaterniond(p[3], p[0], p[1], p[2]);
    T.t(0) = p[4]; T.t(1) = p[5]; T.t(2) = p[6];
    return T;
}

bool Solver::buildSystem() { orc_po_engine_build(_optimizer->_e); return true; }
bool Solver::solve() { return orc_po_engine_solve(_optimizer->_e, _lambda) != 0; }
const double* Solver::x() const { return orc_po_engine_x(_optimizer->_e); }
const double* Solver::b() const { return orc_po_engine_b(_optimizer->_e); }

// initializeOptimization(0): the active set = edges of level 0 (the engine keeps the levels).  The first call hands the oracle's engine
// what the reference code stored in the graph: measurements, information, world points, the vertex estimate.
bool SparseOptimizer::initializeOptimization(int level) {
    assert(level == 0);
    if (!_e) {
        const int n = (int)_edges.size();
        std::vector<double>&Xw = _Xw, &obs = _obs, &invs2 = _invs2;   // the engine keeps pointers: members, not locals
        Xw.assign(3 * (size_t)n, 0.0); obs.assign(3 * (size_t)n, 0.0); invs2.assign(n, 0.0);
        for (int i = 0; i < n; ++i) {
            const OptimizableGraph::Edge* e = _edges[i];
            for (int k = 0; k < 3; ++k) { Xw[3 * i + k] = e->Xw(k); obs[3 * i + k] = e->meas[k]; }
            invs2[i] = e->info00;
            const RobustKernelHuber* rk = dynamic_cast<const RobustKernelHuber*>(e->_rk);
            const float want = e->dim() == 2 ? (float)sqrt(5.991) : (float)sqrt(7.815);   // the engine's deltas (Optimizer.cc:105-107)
            if (!rk || (float)rk->delta() != want || rk->dsqr_value() != (float)((double)want * (double)want)) {
                std::cerr << "po_skel: unexpected robust kernel on edge " << i << std::endl;
                std::abort();
            }
        }
        const SE3Quat& T = static_cast<VertexSE3Expmap*>(_vertex)->_est;
        const double p[7] = {T.q.x, T.q.y, T.q.z, T.q.w, T.t(0), T.t(1), T.t(2)};
        _e = orc_po_engine_create(n, Xw.data(), obs.data(), invs2.data(), g_cam5, p);
        for (int i = 0; i < n; ++i) {
            if (_edges[i]->_level) orc_po_engine_set_level(_e, i, _edges[i]->_level);
            if (_edges[i]->robustCleared) orc_po_engine_set_robust(_e, i, 0);
        }
    }
    _ivMap.clear();
    if (orc_po_engine_active(_e) > 0) _ivMap.push_back(_vertex);
    return true;
}
SparseOptimizer::~SparseOptimizer() {
    for (size_t i = 0; i < _edges.size(); ++i) { delete _edges[i]->_rk; delete _edges[i]; }
    delete _vertex;
    delete _algorithm;
    if (_e) orc_po_engine_destroy(_e);
}

#include "../_ref/gen/ref5_g2o.inc"
}  // namespace g2o

namespace ORB_SLAM3 {
class Optimizer {
   public:
    static int PoseOptimization(Frame* pFrame);
};
#include "../_ref/gen/ref5_opt.inc"
}  // namespace ORB_SLAM3

extern "C" {

// Optimizer::PoseOptimization on a frame given as flat arrays: N features; has_mp[i] != 0: mvpMapPoints[i] is a point at xw[3 i ..]
// (float); kp_xy / octave = mvKeysUn; uright = mvuRight; pose7 = Tcw (float, qx qy qz qw tx ty tz) in / out; outlier[N] = mvbOutlier.
int ref5_pose_optimization(int N, const unsigned char* has_mp, const float* xw, const float* kp_xy, const int* octave, const float* uright,
                           const float* invLevelSigma2, int nlevels, const float* cam5, float* pose7, unsigned char* outlier) {
    using namespace ORB_SLAM3;
    Frame F;
    GeometricCamera cam;
    std::vector<MapPoint> mps(N);
    F.N = N;
    F.mvpMapPoints.assign(N, static_cast<MapPoint*>(NULL));
    F.mvKeysUn.resize(N);
    F.mvuRight.assign(uright, uright + N);
    F.mvbOutlier.assign(N, false);
    F.mvInvLevelSigma2.assign(invLevelSigma2, invLevelSigma2 + nlevels);
    for (int i = 0; i < N; ++i) {
        F.mvKeysUn[i].pt.x = kp_xy[2 * i]; F.mvKeysUn[i].pt.y = kp_xy[2 * i + 1]; F.mvKeysUn[i].octave = octave[i];
        if (has_mp[i]) {
            mps[i].mWorldPos(0) = xw[3 * i]; mps[i].mWorldPos(1) = xw[3 * i + 1]; mps[i].mWorldPos(2) = xw[3 * i + 2];
            F.mvpMapPoints[i] = &mps[i];
        }
    }
    F.mvKeys = F.mvKeysUn;
    F.mpCamera = &cam;
    F.fx = cam5[0]; F.fy = cam5[1]; F.cx = cam5[2]; F.cy = cam5[3]; F.mbf = cam5[4];
    F.mTcw.q = Eigen::Quaternionf(pose7[3], pose7[0], pose7[1], pose7[2]);
    F.mTcw.t(0) = pose7[4]; F.mTcw.t(1) = pose7[5]; F.mTcw.t(2) = pose7[6];
    const double cam5d[5] = {cam5[0], cam5[1], cam5[2], cam5[3], cam5[4]};
    g2o::SparseOptimizer::g_cam5 = cam5d;
    g2o::SparseOptimizer::g_iters = g2o::SparseOptimizer::g_trials = 0;
    const int r = Optimizer::PoseOptimization(&F);
    for (int i = 0; i < N; ++i) outlier[i] = F.mvbOutlier[i] ? 1 : 0;
    pose7[0] = F.mTcw.q.x; pose7[1] = F.mTcw.q.y; pose7[2] = F.mTcw.q.z; pose7[3] = F.mTcw.q.w;
    pose7[4] = F.mTcw.t(0); pose7[5] = F.mTcw.t(1); pose7[6] = F.mTcw.t(2);
    return r;
}
void ref5_last_counts(int* iters_trials) { iters_trials[0] = g2o::SparseOptimizer::g_iters; iters_trials[1] = g2o::SparseOptimizer::g_trials; }

}  // extern "C"

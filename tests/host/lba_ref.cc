// tests/host/lba_ref.cc -- TEST INFRASTRUCTURE (CPU tier): the REFERENCE's own Optimizer::LocalBundleAdjustment, cut out of
// /root/reference/src/Optimizer.cc at build time (tests/host/build_lba_cpu.sh -> tests/host/_gen/, git-ignored) and compiled verbatim
// over the skeleton map of host/refshim and the graph stand-ins of lba_ref_graph.h.  See that header.
#include "lba_ref_graph.h"
#include "Optimizer.h"   // the reference's header (parsed against the skeletons)

#include <mutex>

using namespace std;

namespace ORB_SLAM3 {
class Verbose {   // include/Tracking.h's console helper: the function reports "0 fixed KF" through it
   public:
    enum eLevel { VERBOSITY_QUIET = 0, VERBOSITY_NORMAL = 1, VERBOSITY_VERBOSE = 2, VERBOSITY_VERY_VERBOSE = 3, VERBOSITY_DEBUG = 4 };
    static void PrintMess(std::string, eLevel) {}
};
// the edge classes of include/OptimizableTypes.h as far as LocalBundleAdjustment touches them
class EdgeSE3ProjectXYZ : public g2o::OptimizableGraph::Edge {
   public:
    int dim() const { return 2; }
    void setMeasurement(const Eigen::Matrix<double, 2, 1>& m) { meas[0] = m(0); meas[1] = m(1); meas[2] = -1.0; }
    void setInformation(const Eigen::Matrix2d& I) { assert(I(0, 0) == I(1, 1) && I(0, 1) == 0); info00 = I(0, 0); }
    GeometricCamera* pCamera = nullptr;
};
class EdgeSE3ProjectXYZToBody : public EdgeSE3ProjectXYZ {
   public:
    g2o::SE3Quat mTrl;
};
}  // namespace ORB_SLAM3

namespace g2o {
class EdgeStereoSE3ProjectXYZ : public OptimizableGraph::Edge {
   public:
    int dim() const { return 3; }
    void setMeasurement(const Eigen::Matrix<double, 3, 1>& m) { meas[0] = m(0); meas[1] = m(1); meas[2] = m(2); }
    void setInformation(const Eigen::Matrix3d& I) { assert(I(0, 0) == I(1, 1) && I(1, 1) == I(2, 2) && I(0, 1) == 0); info00 = I(0, 0); }
    double fx = 0, fy = 0, cx = 0, cy = 0, bf = 0;
};

double OptimizableGraph::Vertex::hessian(int i, int j) const { assert(i == j); return orc_lba_engine_hessian_diag(_opt->_e, _index, j); }
double OptimizableGraph::Edge::chi2() const { _opt->sync_results(); return _opt->_chi2[index]; }
bool OptimizableGraph::Edge::isDepthPositive() { _opt->sync_results(); return _opt->_depthPos[index] != 0; }

SE3Quat VertexSE3Expmap::estimate() const {
    if (!(_opt && _opt->_e)) return _est;
    _opt->sync_results();
    int k = 0;
    for (; k < (int)_opt->_poseV.size(); ++k) if (_opt->_poseV[k] == this) break;
    const double* p = &_opt->_poseOut[7 * (size_t)k];
    SE3Quat T;
    T.q = Eigen::Quaterniond(p[3], p[0], p[1], p[2]);
    T.t = Eigen::Vector3d(p[4], p[5], p[6]);
    return T;
}
Eigen::Vector3d VertexSBAPointXYZ::estimate() const {
    if (!(_opt && _opt->_e)) return _est;
    _opt->sync_results();
    int k = 0;
    for (; k < (int)_opt->_pointV.size(); ++k) if (_opt->_pointV[k] == this) break;
    const double* p = &_opt->_pointOut[3 * (size_t)k];
    return Eigen::Vector3d(p[0], p[1], p[2]);
}

bool Solver::buildSystem() { orc_lba_engine_build(_optimizer->_e); return true; }
bool Solver::solve() { return orc_lba_engine_solve(_optimizer->_e, _lambda) != 0; }
const double* Solver::x() const { return orc_lba_engine_x(_optimizer->_e); }
const double* Solver::b() const { return orc_lba_engine_b(_optimizer->_e); }
size_t Solver::vectorSize() const { return (size_t)orc_lba_engine_vector_size(_optimizer->_e); }

extern "C" const double* lba_ref_cam5();   // fx fy cx cy bf of the window (the engine takes them at creation): set by the driver

// initializeOptimization(): hands the oracle's engine what the reference code stored -- pose vertices in insertion order (local
// keyframes, then fixed ones), point vertices, edges in creation order with the indices of their two vertices.
bool SparseOptimizer::initializeOptimization(int) {
    if (_e) return true;
    for (size_t i = 0; i < _vertices.size(); ++i) {
        if (dynamic_cast<VertexSE3Expmap*>(_vertices[i])) _poseV.push_back(_vertices[i]);
        else _pointV.push_back(_vertices[i]);
    }
    const int nKF = (int)_poseV.size(), nMP = (int)_pointV.size(), nE = (int)_edges.size();
    std::map<OptimizableGraph::Vertex*, int> idx;
    _pose.assign(7 * (size_t)nKF, 0.0); _fixed.assign(nKF, 0); _point.assign(3 * (size_t)nMP, 0.0);
    int nFree = 0;
    for (int k = 0; k < nKF; ++k) {
        const SE3Quat& T = static_cast<VertexSE3Expmap*>(_poseV[k])->_est;
        double* p = &_pose[7 * (size_t)k];
        p[0] = T.q.x(); p[1] = T.q.y(); p[2] = T.q.z(); p[3] = T.q.w(); p[4] = T.t(0); p[5] = T.t(1); p[6] = T.t(2);
        _fixed[k] = _poseV[k]->fixed() ? 1 : 0;
        idx[_poseV[k]] = k;
        if (!_fixed[k]) { _poseV[k]->_index = nFree++; _ivMap.push_back(_poseV[k]); }
    }
    for (int l = 0; l < nMP; ++l) {
        const Eigen::Vector3d& X = static_cast<VertexSBAPointXYZ*>(_pointV[l])->_est;
        for (int c = 0; c < 3; ++c) _point[3 * (size_t)l + c] = X(c);
        idx[_pointV[l]] = l;
        _pointV[l]->_index = nFree + l;
        _ivMap.push_back(_pointV[l]);
    }
    _ekf.assign(nE, 0); _emp.assign(nE, 0); _obs.assign(3 * (size_t)nE, 0.0); _invs2.assign(nE, 0.0);
    for (int e = 0; e < nE; ++e) {
        const OptimizableGraph::Edge* E = _edges[e];
        _emp[e] = idx[E->_v[0]];     // setVertex(0, the point), setVertex(1, the keyframe)
        _ekf[e] = idx[E->_v[1]];
        for (int c = 0; c < 3; ++c) _obs[3 * (size_t)e + c] = E->meas[c];
        _invs2[e] = E->info00;
        const RobustKernelHuber* rk = dynamic_cast<const RobustKernelHuber*>(E->_rk);
        const float want = E->dim() == 2 ? (float)sqrt(5.991) : (float)sqrt(7.815);   // Optimizer.cc:1957-1958 (the engine's deltas)
        if (!rk || (float)rk->delta() != want || rk->dsqr_value() != (float)((double)want * (double)want)) {
            std::cerr << "lba_ref_graph: unexpected robust kernel on edge " << e << std::endl;
            std::abort();
        }
    }
    const double* c5 = lba_ref_cam5();
    for (int c = 0; c < 5; ++c) _cam5[c] = c5[c];
    _e = orc_lba_engine_create(nKF, nMP, nE, _pose.data(), _fixed.data(), _point.data(), _ekf.data(), _emp.data(), _obs.data(), _invs2.data(), _cam5);
    return true;
}
void SparseOptimizer::sync_results() {
    if (_synced || !_e) return;
    _poseOut.assign(_pose.size(), 0.0); _pointOut.assign(_point.size(), 0.0);
    _chi2.assign(_edges.size() + 1, 0.0); _depthPos.assign(_edges.size() + 1, 0);
    orc_lba_engine_finish(_e, _poseOut.data(), _pointOut.data(), _chi2.data(), _depthPos.data());
    _synced = true;
}
SparseOptimizer::~SparseOptimizer() {
    for (size_t i = 0; i < _edges.size(); ++i) { delete _edges[i]->_rk; delete _edges[i]; }
    for (size_t i = 0; i < _vertices.size(); ++i) delete _vertices[i];
    delete _algorithm;
    if (_e) orc_lba_engine_destroy(_e);
}

#include "_gen/lba_g2o.inc"
}  // namespace g2o

namespace ORB_SLAM3 {
#include "_gen/lba_opt.inc"
}  // namespace ORB_SLAM3

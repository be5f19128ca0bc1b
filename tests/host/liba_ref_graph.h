// tests/host/liba_ref_graph.h -- TEST INFRASTRUCTURE (CPU tier).
//
// Stand-ins for the g2o graph classes and the IMU vertex / edge types (include/G2oTypes.h) that the reference's own
// Optimizer::LocalInertialBA (src/Optimizer.cc:2203-2812) builds its problem with, so that the function -- cut out of the reference at
// build time and compiled verbatim (tests/host/build_liba_cpu.sh) -- runs over the skeleton map of host/refshim.  The stand-ins only
// RECORD what the reference code stores (vertex estimates, fixed flags, which vertices an edge joins, measurements, information
// matrices, robust kernels and their deltas, the algorithm's lambda); optimize(n) flattens the recorded graph -- keyframes, points and
// edges in insertion order -- into the arrays of liba_solve and runs it (tests/host/liba_stub.cc: the oracle's orc_liba), then the
// estimates and per-edge chi2 the function reads back come from that result.  Restated here, not the reference's: the one-line vertex
// constructors of G2oTypes.cc (VertexVelocity(pKF) = pKF->GetVelocity() ...), ImuCamPose's constructor as far as one camera goes
// (:30-70) and its Update rule Rcw = Rcb Rbw (:236-243), EdgeInertial's information (G2oTypes.cc:575-586, through the product's host
// helper liba_link_information, which both binaries share).  tests/test_host_liba_vs_ref.py compares the result with
// host/Optimizer_liba_b200.cc run over the same solver.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <vector>

#include "Eigen/Core"
#include "orbslam3_b200.h"

namespace ORB_SLAM3 {
class KeyFrame;
class GeometricCamera {   // Pinhole / KannalaBrandt8::uncertainty2 both return 1.0 (Pinhole.h, KannalaBrandt8.h)
   public:
    float uncertainty2(const Eigen::Matrix<double, 2, 1>&) { return 1.0f; }
};
}  // namespace ORB_SLAM3

namespace g2o {

class RobustKernel {
   public:
    virtual ~RobustKernel() {}
    void setDelta(double d) { _delta = d; }
    double _delta = 1.0;
};
class RobustKernelHuber : public RobustKernel {};

class SparseOptimizer;
class HyperGraph {
   public:
    class Vertex {
       public:
        virtual ~Vertex() {}
        void setId(int id) { _id = id; }
        int id() const { return _id; }
        void setFixed(bool f) { _fixed = f; }
        bool fixed() const { return _fixed; }
        void setMarginalized(bool) {}
        int _id = -1;
        bool _fixed = false;
        SparseOptimizer* _opt = nullptr;
        ORB_SLAM3::KeyFrame* _kf = nullptr;   // the keyframe a keyframe-bound vertex was made from
    };
};
class OptimizableGraph : public HyperGraph {
   public:
    typedef HyperGraph::Vertex Vertex;
    class Edge {
       public:
        virtual ~Edge() {}
        void setVertex(size_t i, HyperGraph::Vertex* v) { if (!v) { std::cerr << "liba_ref_graph: null vertex" << std::endl; std::abort(); } if (_v.size() <= i) _v.resize(i + 1, nullptr); _v[i] = v; }
        void setRobustKernel(RobustKernel* rk) { _rk = rk; }
        std::vector<HyperGraph::Vertex*> _v;
        RobustKernel* _rk = nullptr;
        SparseOptimizer* _opt = nullptr;
        int _index = -1;     // among the reprojection edges (creation order), or among the inertial edges
    };
};

class VertexSBAPointXYZ : public HyperGraph::Vertex {
   public:
    void setEstimate(const Eigen::Vector3d& X) { _est = X; }
    const Eigen::Vector3d& estimate() const { return _est; }
    Eigen::Vector3d _est;
};

template <typename M> class LinearSolver { public: virtual ~LinearSolver() {} };
template <typename M> class LinearSolverEigen : public LinearSolver<M> {};
class BlockSolverX {
   public:
    typedef int PoseMatrixType;
    typedef LinearSolver<PoseMatrixType> LinearSolverType;
    explicit BlockSolverX(LinearSolverType* ls) : _ls(ls) {}
    ~BlockSolverX() { delete _ls; }
    LinearSolverType* _ls;
};
class OptimizationAlgorithmLevenberg {
   public:
    explicit OptimizationAlgorithmLevenberg(BlockSolverX* s) : _solver(s) {}
    ~OptimizationAlgorithmLevenberg() { delete _solver; }
    void setUserLambdaInit(double l) { _lambdaInit = l; }
    BlockSolverX* _solver;
    double _lambdaInit = 0;
};

}  // namespace g2o

namespace ORB_SLAM3 {

typedef Eigen::Matrix<double, 6, 1> Vector6d;

struct ImuCamPose {   // include/G2oTypes.h:61-103 as far as one camera goes
    Eigen::Vector3d twb;
    Eigen::Matrix3d Rwb;
    std::vector<Eigen::Matrix3d> Rcw, Rcb;
    std::vector<Eigen::Vector3d> tcw, tcb;
    double bf = 0, fx = 0, fy = 0, cx = 0, cy = 0;
};
class VertexPose : public g2o::HyperGraph::Vertex {
   public:
    explicit VertexPose(KeyFrame* pKF);
    const ImuCamPose& estimate() const { return _est; }
    ImuCamPose _est;
};
class VertexVelocity : public g2o::HyperGraph::Vertex {
   public:
    explicit VertexVelocity(KeyFrame* pKF);
    const Eigen::Vector3d& estimate() const { return _est; }
    Eigen::Vector3d _est;
};
class VertexGyroBias : public VertexVelocity { public: explicit VertexGyroBias(KeyFrame* pKF); };
class VertexAccBias : public VertexVelocity { public: explicit VertexAccBias(KeyFrame* pKF); };

class EdgeInertial : public g2o::OptimizableGraph::Edge {
   public:
    explicit EdgeInertial(IMU::Preintegrated* pInt);
    const Eigen::Matrix<double, 9, 9>& information() const { return _info; }
    void setInformation(const Eigen::Matrix<double, 9, 9>& I) { _info = I; }
    IMU::Preintegrated* mpInt;
    Eigen::Matrix<double, 9, 9> _info;
};
class EdgeGyroRW : public g2o::OptimizableGraph::Edge {
   public:
    void setInformation(const Eigen::Matrix3d& I) { _info = I; }
    Eigen::Matrix3d _info;
};
class EdgeAccRW : public EdgeGyroRW {};

class EdgeReproj : public g2o::OptimizableGraph::Edge {
   public:
    double chi2() const;
    bool isDepthPositive() const;
    double meas[3] = {0, 0, -1};
    double info00 = 0;
    int cam = 0;
};
class EdgeMono : public EdgeReproj {
   public:
    explicit EdgeMono(int cam_idx = 0) { cam = cam_idx; }
    void setMeasurement(const Eigen::Matrix<double, 2, 1>& m) { meas[0] = m(0); meas[1] = m(1); meas[2] = -1.0; }
    void setInformation(const Eigen::Matrix2d& I) { assert(I(0, 0) == I(1, 1) && I(0, 1) == 0); info00 = I(0, 0); }
};
class EdgeStereo : public EdgeReproj {
   public:
    explicit EdgeStereo(int cam_idx = 0) { cam = cam_idx; }
    void setMeasurement(const Eigen::Matrix<double, 3, 1>& m) { meas[0] = m(0); meas[1] = m(1); meas[2] = m(2); }
    void setInformation(const Eigen::Matrix3d& I) { assert(I(0, 0) == I(1, 1) && I(1, 1) == I(2, 2) && I(0, 1) == 0); info00 = I(0, 0); }
};

}  // namespace ORB_SLAM3

namespace g2o {

class SparseOptimizer {
   public:
    ~SparseOptimizer();
    void setAlgorithm(OptimizationAlgorithmLevenberg* a) { _algorithm = a; }
    void setForceStopFlag(bool* f) { _forceStopFlag = f; }
    bool addVertex(HyperGraph::Vertex* v) { v->_opt = this; _byId[v->id()] = v; _vertices.push_back(v); return true; }
    bool addEdge(OptimizableGraph::Edge* e) { e->_opt = this; _edges.push_back(e); return true; }
    HyperGraph::Vertex* vertex(int id) { std::map<int, HyperGraph::Vertex*>::iterator it = _byId.find(id); return it == _byId.end() ? nullptr : it->second; }
    bool initializeOptimization(int = 0) { return true; }
    void computeActiveErrors() { if (!_solved) solve(-1); }   // before optimize(): only the initial chi2 is read (one throw-away iteration computes it)
    double activeRobustChi2() const { return _solved > 1 ? _chiLast : _chiInit; }
    int optimize(int iterations) { solve(iterations); return _iterations; }
    void solve(int iterations);       // flatten -> liba_solve -> estimates / chi2 back into the graph
    OptimizationAlgorithmLevenberg* _algorithm = nullptr;
    bool* _forceStopFlag = nullptr;
    std::map<int, HyperGraph::Vertex*> _byId;
    std::vector<HyperGraph::Vertex*> _vertices;
    std::vector<OptimizableGraph::Edge*> _edges;
    int _solved = 0, _iterations = 0;
    double _chiInit = 0, _chiLast = 0;
    std::vector<double> _chi2;
    std::vector<uint8_t> _depthPos;
};

}  // namespace g2o

// tests/host/lba_ref_graph.h -- TEST INFRASTRUCTURE (CPU tier).
//
// Stand-ins for the g2o graph classes that the reference's own Optimizer::LocalBundleAdjustment (src/Optimizer.cc:1740-2188) builds
// its problem with, so that the function -- cut out of the reference at build time and compiled verbatim (tests/host/build_lba_cpu.sh)
// -- runs over the skeleton map of host/refshim: window selection, which observers become fixed cameras, vertex ids, which
// observations become monocular / stereo edges, the early exits, optimize(10) through g2o's own Levenberg functions (also extracted),
// the chi2 / depth test of every edge, the erasures, the write-back.  The numerics (errors, Jacobians, Schur system, LDL^T, updates)
// are the oracle's LbaEngine (oracle/lba_oracle.cpp), created from what the reference code stored in the vertices and edges.
// tests/test_host_lba_vs_ref.py compares the result with host/Optimizer_lba_b200.cc run over the same engine.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include "Eigen/Core"

#define FIXED(x) x
#define g2o_isfinite(x) std::isfinite(x)

extern "C" {   // oracle/lba_oracle.cpp
void* orc_lba_engine_create(int nKF, int nMP, int nE, const double* pose, const uint8_t* fixed, const double* point, const int* ekf, const int* emp,
                            const double* obs, const double* invs2, const double* cam5);
void orc_lba_engine_destroy(void* h);
double orc_lba_engine_errors(void* h);
void orc_lba_engine_build(void* h);
int orc_lba_engine_solve(void* h, double lambda);
void orc_lba_engine_update(void* h, const double* u);
void orc_lba_engine_push(void* h);
void orc_lba_engine_pop(void* h);
void orc_lba_engine_discard_top(void* h);
int orc_lba_engine_vector_size(void* h);
const double* orc_lba_engine_x(void* h);
const double* orc_lba_engine_b(void* h);
int orc_lba_engine_vertices(void* h);
int orc_lba_engine_vertex_dim(void* h, int k);
double orc_lba_engine_hessian_diag(void* h, int k, int j);
void orc_lba_engine_finish(void* h, double* pose, double* point, double* edge_chi2, uint8_t* edge_depth_pos);
}

namespace g2o {

inline double get_monotonic_time() { return 0.0; }
struct G2OBatchStatistics {
    int iteration = 0, numVertices = 0, numEdges = 0, levenbergIterations = 0;
    double chi2 = 0, timeResiduals = 0, timeQuadraticForm = 0, timeLinearSolution = 0, timeUpdate = 0, timeIteration = 0;
    static G2OBatchStatistics* globalStats() { return nullptr; }
    static void setGlobalStats(G2OBatchStatistics*) {}
};
template <typename T>
class Property {
   public:
    Property(const std::string&, const T& v) : _v(v) {}
    const T& value() const { return _v; }
    void setValue(const T& v) { _v = v; }
   private:
    T _v;
};
class PropertyMap {
   public:
    template <typename P, typename V> P* makeProperty(const std::string& name, const V& v) { return new P(name, v); }
};

struct SE3Quat {
    Eigen::Quaterniond q;
    Eigen::Vector3d t;
    SE3Quat() {}
    SE3Quat(const Eigen::Quaterniond& q_, const Eigen::Vector3d& t_) : q(q_), t(t_) {}
    const Eigen::Quaterniond& rotation() const { return q; }
    const Eigen::Vector3d& translation() const { return t; }
};

class RobustKernel {
   public:
    RobustKernel() : _delta(1.) {}
    virtual ~RobustKernel() {}
    virtual void setDelta(double delta) { _delta = delta; }
    double delta() const { return _delta; }
   protected:
    double _delta;
};
class RobustKernelHuber : public RobustKernel {
   public:
    virtual void setDelta(double delta);   // extracted from robust_kernel_impl.cpp
    float dsqr_value() const { return dsqr; }
   private:
    float dsqr;
};

class SparseOptimizer;

class OptimizableGraph {
   public:
    class Vertex {
       public:
        virtual ~Vertex() {}
        virtual int dimension() const = 0;
        double hessian(int i, int j) const;
        void setId(int id) { _id = id; }
        int id() const { return _id; }
        void setFixed(bool f) { _fixed = f; }
        bool fixed() const { return _fixed; }
        void setMarginalized(bool) {}
        int _id = -1, _index = -1;     // _index: position among the engine's vertices (poses first)
        bool _fixed = false;
        SparseOptimizer* _opt = nullptr;
    };
    class Edge {
       public:
        virtual ~Edge() {}
        virtual int dim() const = 0;
        void setVertex(size_t i, Vertex* v) { assert(v); _v[i] = v; }
        void setRobustKernel(RobustKernel* rk) { _rk = rk; }
        void setLevel(int) {}
        double chi2() const;
        bool isDepthPositive();
        double meas[3] = {0, 0, -1};
        double info00 = 0;
        RobustKernel* _rk = nullptr;
        Vertex* _v[2] = {nullptr, nullptr};
        int index = -1;
        SparseOptimizer* _opt = nullptr;
    };
};

class VertexSE3Expmap : public OptimizableGraph::Vertex {
   public:
    int dimension() const { return 6; }
    void setEstimate(const SE3Quat& T) { _est = T; }
    SE3Quat estimate() const;
    SE3Quat _est;
};
class VertexSBAPointXYZ : public OptimizableGraph::Vertex {
   public:
    int dimension() const { return 3; }
    void setEstimate(const Eigen::Vector3d& X) { _est = X; }
    Eigen::Vector3d estimate() const;
    Eigen::Vector3d _est;
};

class Solver {
   public:
    Solver() : _optimizer(nullptr), _lambda(0) {}
    virtual ~Solver() {}
    bool buildStructure() { return true; }
    bool buildSystem();
    bool setLambda(double lambda, bool = false) { _lambda = lambda; return true; }
    bool solve();
    void restoreDiagonal() {}
    const double* x() const;
    const double* b() const;
    size_t vectorSize() const;
    SparseOptimizer* optimizer() const { return _optimizer; }
    SparseOptimizer* _optimizer;
    double _lambda;
};
template <typename M> class LinearSolver { public: virtual ~LinearSolver() {} };
template <typename M> class LinearSolverEigen : public LinearSolver<M> {};
template <typename M> class LinearSolverDense : public LinearSolver<M> {};
class BlockSolver_6_3 : public Solver {
   public:
    typedef int PoseMatrixType;
    typedef LinearSolver<PoseMatrixType> LinearSolverType;
    explicit BlockSolver_6_3(LinearSolverType* ls) : _ls(ls) {}
    ~BlockSolver_6_3() { delete _ls; }
    LinearSolverType* _ls;
};

class OptimizationAlgorithm {
   public:
    enum SolverResult { Terminate = 2, OK = 1, Fail = -1 };
    OptimizationAlgorithm() : _optimizer(nullptr) {}
    virtual ~OptimizationAlgorithm() {}
    virtual bool init(bool online = false) = 0;
    virtual SolverResult solve(int iteration, bool online = false) = 0;
    virtual void printVerbose(std::ostream&) const {}
    SparseOptimizer* _optimizer;
    PropertyMap _properties;
};
class OptimizationAlgorithmWithHessian : public OptimizationAlgorithm {
   public:
    explicit OptimizationAlgorithmWithHessian(Solver* solver) : _solver(solver) {}
    ~OptimizationAlgorithmWithHessian() { delete _solver; }
    virtual bool init(bool = false) { return true; }
    Solver* _solver;
};
class OptimizationAlgorithmLevenberg : public OptimizationAlgorithmWithHessian {
   public:
    explicit OptimizationAlgorithmLevenberg(Solver* solver);
    virtual ~OptimizationAlgorithmLevenberg() {}
    virtual SolverResult solve(int iteration, bool online = false);
    void setUserLambdaInit(double lambda) { _userLambdaInit->setValue(lambda); }
    int levenbergIteration() { return _levenbergIterations; }
   protected:
    Property<int>* _maxTrialsAfterFailure;
    Property<double>* _userLambdaInit;
    double _currentLambda, _tau, _goodStepLowerScale, _goodStepUpperScale, _ni;
    int _levenbergIterations, _nBad;
    double computeLambdaInit() const;
    double computeScale() const;
};

class SparseOptimizer {
   public:
    typedef std::vector<OptimizableGraph::Vertex*> VertexContainer;
    SparseOptimizer() : _e(nullptr), _algorithm(nullptr), _computeBatchStatistics(false), _forceStopFlag(nullptr), _chi(0) {}
    ~SparseOptimizer();
    void setAlgorithm(OptimizationAlgorithm* a) { _algorithm = a; a->_optimizer = this; static_cast<OptimizationAlgorithmWithHessian*>(a)->_solver->_optimizer = this; }
    void setVerbose(bool) {}
    void setForceStopFlag(bool* f) { _forceStopFlag = f; }
    bool addVertex(OptimizableGraph::Vertex* v) { v->_opt = this; _byId[v->id()] = v; _vertices.push_back(v); return true; }
    bool addEdge(OptimizableGraph::Edge* e) { e->_opt = this; e->index = (int)_edges.size(); _edges.push_back(e); return true; }
    OptimizableGraph::Vertex* vertex(int id) { std::map<int, OptimizableGraph::Vertex*>::iterator it = _byId.find(id); return it == _byId.end() ? nullptr : it->second; }
    bool initializeOptimization(int level = 0);
    int optimize(int iterations, bool online = false);   // body extracted from g2o/core/sparse_optimizer.cpp
    void computeActiveErrors() { _chi = orc_lba_engine_errors(_e); }
    double activeRobustChi2() const { return _chi; }
    void push() { orc_lba_engine_push(_e); }
    void pop() { orc_lba_engine_pop(_e); }
    void discardTop() { orc_lba_engine_discard_top(_e); }
    void update(const double* u) { orc_lba_engine_update(_e, u); }
    bool terminate() { return _forceStopFlag ? (*_forceStopFlag) : false; }
    const VertexContainer& indexMapping() const { return _ivMap; }
    bool verbose() const { return false; }
    void preIteration(int) {}
    void postIteration(int) {}
    void sync_results();                       // engine estimates / per-edge chi2 and depth signs, fetched once after optimize()
    void* _e;
    VertexContainer _ivMap;                     // the optimisable vertices in the engine's order: free poses, then points
    std::vector<int> _activeEdges, _activeVertices;
    OptimizationAlgorithm* _algorithm;
    std::vector<G2OBatchStatistics> _batchStatistics;
    bool _computeBatchStatistics;
    bool* _forceStopFlag;
    double _chi;
    std::map<int, OptimizableGraph::Vertex*> _byId;
    std::vector<OptimizableGraph::Vertex*> _vertices;
    std::vector<OptimizableGraph::Edge*> _edges;
    // the engine's inputs (it keeps pointers) and the results fetched back
    std::vector<double> _pose, _point, _obs, _invs2, _poseOut, _pointOut, _chi2;
    std::vector<uint8_t> _fixed, _depthPos;
    std::vector<int> _ekf, _emp;
    double _cam5[5];
    bool _synced = false;
    std::vector<OptimizableGraph::Vertex*> _poseV, _pointV;
};

}  // namespace g2o

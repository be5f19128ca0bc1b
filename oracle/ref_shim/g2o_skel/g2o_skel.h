// oracle/ref_shim/g2o_skel/g2o_skel.h -- TEST INFRASTRUCTURE.
//
// Skeleton of the g2o classes around the functions oracle/Makefile (target ref4) cuts out of the reference's Thirdparty/g2o at build
// time and compiles VERBATIM:
//     g2o/core/optimization_algorithm_levenberg.cpp : the constructor (tau, the good-step bounds, maxTrialsAfterFailure = 10),
//                                                     solve(int, bool), computeLambdaInit(), computeScale()
//     g2o/core/sparse_optimizer.cpp                 : SparseOptimizer::optimize(int, bool)
//     g2o/core/robust_kernel_impl.cpp               : RobustKernelHuber::setDelta, robustify  (with the `float dsqr` member of the header)
// None of those touches Eigen: they call an abstract Solver / SparseOptimizer.  Here both forward to the oracle's LbaEngine
// (oracle/lba_oracle.cpp: errors + robust chi2, buildSystem, setLambda + solve, update, push / pop), so what runs is g2o's own
// Levenberg control flow -- rho, the lambda update, _ni, qmax, the nBad stop of ORB-SLAM3's fork, the iteration loop -- over the
// oracle's numerics, to be compared with orc_lba's restated loop (tests/test_oracle_vs_ref_g2o.py).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <iostream>
#include <limits>
#include <string>
#include <vector>

#define FIXED(x) x
#define g2o_isfinite(x) std::isfinite(x)

namespace Eigen {
struct Vector3d {
    double v[3];
    double& operator[](int i) { return v[i]; }
    const double& operator[](int i) const { return v[i]; }
};
}  // namespace Eigen

extern "C" {   // oracle/lba_oracle.cpp
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
}

namespace g2o {

inline double get_monotonic_time() { return 0.0; }

struct G2OBatchStatistics {
    int iteration = 0, numVertices = 0, numEdges = 0, levenbergIterations = 0;
    double chi2 = 0, timeResiduals = 0, timeQuadraticForm = 0, timeLinearSolution = 0, timeUpdate = 0, timeIteration = 0;
    static G2OBatchStatistics* globalStats() { return _g; }
    static void setGlobalStats(G2OBatchStatistics* b) { _g = b; }
    static G2OBatchStatistics* _g;
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
    template <typename P, typename V>
    P* makeProperty(const std::string& name, const V& v) { return new P(name, v); }   // leaked like a map that is never cleared: test code
};

class SparseOptimizer;

class Solver {   // BlockSolver<6,3> + LinearSolverEigen as the engine implements them
   public:
    Solver(void* engine) : _e(engine), _optimizer(nullptr), _lambda(0) {}
    bool buildStructure() { return true; }
    bool buildSystem() { orc_lba_engine_build(_e); return true; }
    bool setLambda(double lambda, bool = false) { _lambda = lambda; return true; }
    bool solve() { return orc_lba_engine_solve(_e, _lambda) != 0; }
    void restoreDiagonal() {}
    const double* x() const { return orc_lba_engine_x(_e); }
    const double* b() const { return orc_lba_engine_b(_e); }
    size_t vectorSize() const { return (size_t)orc_lba_engine_vector_size(_e); }
    SparseOptimizer* optimizer() const { return _optimizer; }
    void* _e;
    SparseOptimizer* _optimizer;
    double _lambda;
};

class OptimizableGraph {
   public:
    class Vertex {
       public:
        Vertex(void* e, int k) : _e(e), _k(k) {}
        int dimension() const { return orc_lba_engine_vertex_dim(_e, _k); }
        double hessian(int i, int j) const { assert(i == j); return orc_lba_engine_hessian_diag(_e, _k, j); }
        void* _e;
        int _k;
    };
};

class OptimizationAlgorithm {
   public:
    enum SolverResult { Terminate = 2, OK = 1, Fail = -1 };   // optimization_algorithm.h:54
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
    virtual bool init(bool = false) { return true; }
    Solver* _solver;
};
class OptimizationAlgorithmLevenberg : public OptimizationAlgorithmWithHessian {   // members: optimization_algorithm_levenberg.h:70-79
   public:
    explicit OptimizationAlgorithmLevenberg(Solver* solver);
    virtual ~OptimizationAlgorithmLevenberg() {}
    virtual SolverResult solve(int iteration, bool online = false);
    double currentLambda() const { return _currentLambda; }
    void setMaxTrialsAfterFailure(int max_trials) { _maxTrialsAfterFailure->setValue(max_trials); }
    void setUserLambdaInit(double lambda) { _userLambdaInit->setValue(lambda); }
    int levenbergIteration() { return _levenbergIterations; }
    int totalTrials = 0;   // not in g2o: summed by the wrapper after every solve()
   protected:
    Property<int>* _maxTrialsAfterFailure;
    Property<double>* _userLambdaInit;
    double _currentLambda;
    double _tau;
    double _goodStepLowerScale;
    double _goodStepUpperScale;
    double _ni;
    int _levenbergIterations;
    int _nBad;
    double computeLambdaInit() const;
    double computeScale() const;
};

class SparseOptimizer {
   public:
    typedef std::vector<OptimizableGraph::Vertex*> VertexContainer;
    explicit SparseOptimizer(void* engine) : _e(engine), _algorithm(nullptr), _computeBatchStatistics(false), _forceStopFlag(nullptr), _chi(0), outerIterations(0) {
        const int n = orc_lba_engine_vertices(engine);
        for (int k = 0; k < n; ++k) _ivMap.push_back(new OptimizableGraph::Vertex(engine, k));
    }
    ~SparseOptimizer() { for (size_t i = 0; i < _ivMap.size(); ++i) delete _ivMap[i]; }
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
    void postIteration(int) { ++outerIterations; }
    void setAlgorithm(OptimizationAlgorithm* a) { _algorithm = a; a->_optimizer = this; }
    void* _e;
    VertexContainer _ivMap;
    std::vector<int> _activeEdges, _activeVertices;
    OptimizationAlgorithm* _algorithm;
    std::vector<G2OBatchStatistics> _batchStatistics;
    bool _computeBatchStatistics;
    bool* _forceStopFlag;
    double _chi;
    int outerIterations;
};

class RobustKernel {
   public:
    RobustKernel() : _delta(1.) {}
    virtual ~RobustKernel() {}
    virtual void robustify(double squaredError, Eigen::Vector3d& rho) const = 0;
    virtual void setDelta(double delta) { _delta = delta; }
    double delta() const { return _delta; }
   protected:
    double _delta;
};
class RobustKernelHuber : public RobustKernel {   // robust_kernel_impl.h:76-85
   public:
    virtual void setDelta(double delta);
    virtual void robustify(double e2, Eigen::Vector3d& rho) const;
   private:
    float dsqr;
};

}  // namespace g2o

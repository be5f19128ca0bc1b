// oracle/ref_shim/g2o_skel/po_skel.h -- TEST INFRASTRUCTURE.
//
// Skeleton classes under the reference's own Optimizer::PoseOptimization(Frame*) (src/Optimizer.cc:55-401), which oracle/Makefile
// (target ref5) cuts out at build time and compiles VERBATIM together with g2o's Levenberg control flow (as oracle/_ref part 4).
// The function builds its graph through these stand-ins -- vertex estimate from the frame pose, one edge per feature with a map
// point (monocular when mvuRight < 0, else stereo), measurement, information = invSigma2 * I, Huber kernel with delta = sqrt(5.991) /
// sqrt(7.815) as float -- and then runs its four rounds: optimize(10), chi2 classification as float against 5.991 / 7.815, levels,
// kernel removal after the third round, the `edges().size() < 10` exit, nInitialCorrespondences - nBad.  The numerics behind
// computeError / chi2 / the linear system are the oracle's PoseEngine (oracle/lba_oracle.cpp), created from what the reference code
// stored in the edges.  Compared with orc_pose_optimization in tests/test_oracle_vs_ref_g2o.py.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

#define FIXED(x) x
#define g2o_isfinite(x) std::isfinite(x)

extern "C" {   // oracle/lba_oracle.cpp
void* orc_po_engine_create(int n, const double* Xw, const double* obs, const double* invs2, const double* cam5, const double* pose7);
void orc_po_engine_destroy(void* h);
void orc_po_engine_set_estimate(void* h, const double* pose7);
void orc_po_engine_get_estimate(void* h, double* pose7);
void orc_po_engine_set_level(void* h, int e, int level);
void orc_po_engine_set_robust(void* h, int e, int on);
void orc_po_engine_compute_error(void* h, int e);
double orc_po_engine_chi2(void* h, int e);
int orc_po_engine_active(void* h);
double orc_po_engine_errors(void* h);
void orc_po_engine_build(void* h);
int orc_po_engine_solve(void* h, double lambda);
void orc_po_engine_update(void* h);
void orc_po_engine_push(void* h);
void orc_po_engine_pop(void* h);
void orc_po_engine_discard_top(void* h);
double orc_po_engine_hessian_diag(void* h, int j);
const double* orc_po_engine_x(void* h);
const double* orc_po_engine_b(void* h);
}

namespace Eigen {
template <typename T, int R, int C>
struct Matrix {
    T v[R * C];
    Matrix() { for (int i = 0; i < R * C; ++i) v[i] = T(0); }
    T& operator()(int i) { return v[i]; }
    const T& operator()(int i) const { return v[i]; }
    T& operator[](int i) { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
    T& operator()(int r, int c) { return v[r * C + c]; }
    const T& operator()(int r, int c) const { return v[r * C + c]; }
    struct CommaInit {
        Matrix* m; int i;
        template <typename U> CommaInit& operator,(U x) { m->v[i++] = (T)x; return *this; }
    };
    template <typename U> CommaInit operator<<(U x) { v[0] = (T)x; return CommaInit{this, 1}; }
    static Matrix Identity() { Matrix m; for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = T(1); return m; }
    template <typename S> Matrix operator*(S s) const { Matrix o; for (int i = 0; i < R * C; ++i) o.v[i] = v[i] * (T)s; return o; }
    template <typename U> Matrix<U, R, C> cast() const { Matrix<U, R, C> o; for (int i = 0; i < R * C; ++i) o.v[i] = (U)v[i]; return o; }
};
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<float, 3, 1> Vector3f;
template <typename T>
struct Quaternion {
    T x, y, z, w;
    Quaternion() : x(0), y(0), z(0), w(1) {}
    Quaternion(T w_, T x_, T y_, T z_) : x(x_), y(y_), z(z_), w(w_) {}
    template <typename U> Quaternion<U> cast() const { return Quaternion<U>((U)w, (U)x, (U)y, (U)z); }
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;
}  // namespace Eigen

namespace Sophus {
template <typename T>
class SE3 {
   public:
    Eigen::Quaternion<T> q;
    Eigen::Matrix<T, 3, 1> t;
    SE3() {}
    SE3(const Eigen::Quaternion<T>& q_, const Eigen::Matrix<T, 3, 1>& t_) : q(q_), t(t_) {}   // (Sophus normalises: inputs are unit here)
    const Eigen::Quaternion<T>& unit_quaternion() const { return q; }
    const Eigen::Matrix<T, 3, 1>& translation() const { return t; }
};
}  // namespace Sophus

namespace cv {
struct Point2f { float x, y; };
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; };
}  // namespace cv

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
class RobustKernelHuber : public RobustKernel {   // robust_kernel_impl.h:76-85; setDelta extracted from robust_kernel_impl.cpp
   public:
    virtual void setDelta(double delta);
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
        int dimension() const { return 6; }
        double hessian(int i, int j) const;
        SparseOptimizer* _opt = nullptr;
    };
    class Edge {
       public:
        virtual ~Edge() {}
        virtual int dim() const = 0;
        void setVertex(size_t, Vertex* v) { assert(v); }
        void setRobustKernel(RobustKernel* rk);
        void setLevel(int l);
        void computeError();
        double chi2() const;
        // what the reference code stored
        double meas[3] = {0, 0, -1};
        double info00 = 0;
        RobustKernel* _rk = nullptr;
        bool robustCleared = false;
        int _level = 0;
        Eigen::Vector3d Xw;
        int index = -1;
        SparseOptimizer* _opt = nullptr;
    };
};

class VertexSE3Expmap : public OptimizableGraph::Vertex {
   public:
    void setEstimate(const SE3Quat& T);
    SE3Quat estimate() const;
    void setId(int) {}
    void setFixed(bool f) { assert(!f); }
    SE3Quat _est;
};

class Solver {
   public:
    Solver() : _optimizer(nullptr), _lambda(0) {}
    bool buildStructure() { return true; }
    bool buildSystem();
    bool setLambda(double lambda, bool = false) { _lambda = lambda; return true; }
    bool solve();
    void restoreDiagonal() {}
    const double* x() const;
    const double* b() const;
    size_t vectorSize() const { return 6; }
    SparseOptimizer* optimizer() const { return _optimizer; }
    SparseOptimizer* _optimizer;
    double _lambda;
};
template <typename M> class LinearSolver { public: virtual ~LinearSolver() {} };
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
    double currentLambda() const { return _currentLambda; }
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
    SparseOptimizer() : _e(nullptr), _algorithm(nullptr), _computeBatchStatistics(false), _chi(0), _vertex(nullptr) {}
    ~SparseOptimizer();
    void setAlgorithm(OptimizationAlgorithm* a) { _algorithm = a; a->_optimizer = this; static_cast<OptimizationAlgorithmWithHessian*>(a)->_solver->_optimizer = this; }
    bool addVertex(OptimizableGraph::Vertex* v) { v->_opt = this; _vertex = v; return true; }
    bool addEdge(OptimizableGraph::Edge* e) { e->_opt = this; e->index = (int)_edges.size(); _edges.push_back(e); return true; }
    OptimizableGraph::Vertex* vertex(int) { return _vertex; }
    const std::vector<OptimizableGraph::Edge*>& edges() const { return _edges; }
    bool initializeOptimization(int level = 0);
    int optimize(int iterations, bool online = false);   // body extracted from g2o/core/sparse_optimizer.cpp
    void computeActiveErrors() { _chi = orc_po_engine_errors(_e); }
    double activeRobustChi2() const { return _chi; }
    void push() { orc_po_engine_push(_e); }
    void pop() { orc_po_engine_pop(_e); }
    void discardTop() { orc_po_engine_discard_top(_e); }
    void update(const double*) { orc_po_engine_update(_e); }
    bool terminate() { return false; }
    const VertexContainer& indexMapping() const { return _ivMap; }
    bool verbose() const { return false; }
    void preIteration(int) {}
    void postIteration(int);
    void* _e;                                   // the oracle's PoseEngine, created by the first initializeOptimization
    VertexContainer _ivMap;
    std::vector<int> _activeEdges, _activeVertices;
    OptimizationAlgorithm* _algorithm;
    std::vector<G2OBatchStatistics> _batchStatistics;
    bool _computeBatchStatistics;
    double _chi;
    OptimizableGraph::Vertex* _vertex;
    std::vector<OptimizableGraph::Edge*> _edges;
    std::vector<double> _Xw, _obs, _invs2;      // the engine's inputs (it keeps pointers into them)
    int totalOuterIterations = 0, totalTrials = 0;
    static const double* g_cam5;
    static int g_iters, g_trials;              // summed over every optimize() of the last PoseOptimization call                // fx fy cx cy bf of the frame under optimisation (the engine wants them at creation)
};

}  // namespace g2o

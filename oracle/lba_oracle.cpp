// oracle/lba_oracle.cpp -- TEST INFRASTRUCTURE (see oracle.h).
// CPU (fp64, single-threaded) restatement of Optimizer::LocalBundleAdjustment's numeric core
// (/root/reference/src/Optimizer.cc:1859-2150): g2o Levenberg-Marquardt over BlockSolver_6_3 with
// Huber kernels, restated from
//   Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-199   (LM control)
//   Thirdparty/g2o/g2o/core/block_solver.hpp:354-486, 502-604              (build, lambda, Schur, back-subst)
//   Thirdparty/g2o/g2o/core/base_binary_edge.hpp:55-120                    (constructQuadraticForm)
//   Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:65-91                   (Huber, float dsqr)
//   Thirdparty/g2o/g2o/types/types_six_dof_expmap.cpp:190-275, se3quat.h:217-285   (stereo edge, exp map)
//   src/OptimizableTypes.cpp:175-197, src/CameraModels/Pinhole.cpp:47-54,119-130   (mono edge)
// Eigen's summation order is not reproduced: parity with the GPU path is at 1e-4 on the final
// estimate (BASELINE.json north_star), not bit-exact.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

namespace {

struct Cam { double fx, fy, cx, cy, bf; };

inline void quat_to_R(const double* q, double R[9]) {  // q = x y z w (unit)
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}

inline void R_to_quat(const double R[9], double* q) {  // Eigen's matrix -> quaternion
    const double t = R[0] + R[4] + R[8];
    if (t > 0) {
        double s = std::sqrt(t + 1.0);
        q[3] = 0.5 * s;
        s = 0.5 / s;
        q[0] = (R[7] - R[5]) * s; q[1] = (R[2] - R[6]) * s; q[2] = (R[3] - R[1]) * s;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        double s = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
        q[i] = 0.5 * s;
        s = 0.5 / s;
        q[3] = (R[3 * k + j] - R[3 * j + k]) * s;
        q[j] = (R[3 * j + i] + R[3 * i + j]) * s;
        q[k] = (R[3 * k + i] + R[3 * i + k]) * s;
    }
}

inline void normalize_rotation(double* q) {  // se3quat.h:280-285
    if (q[3] < 0) for (int i = 0; i < 4; ++i) q[i] = -q[i];
    const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] /= n;
}

// T <- exp(update) * T   (VertexSE3Expmap::oplusImpl, types_six_dof_expmap.h:73-76; se3quat.h:223-255)
void pose_oplus(double* T, const double* upd) {
    const double w[3] = {upd[0], upd[1], upd[2]}, v[3] = {upd[3], upd[4], upd[5]};
    const double theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const double O[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double O2[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) O2[3 * i + j] = O[3 * i] * O[j] + O[3 * i + 1] * O[3 + j] + O[3 * i + 2] * O[6 + j];
    double Rd[9], V[9];
    for (int i = 0; i < 9; ++i) {
        const double I = (i % 4 == 0) ? 1.0 : 0.0;
        if (theta < 0.00001) {
            Rd[i] = I + O[i] + O2[i];
            V[i] = Rd[i];
        } else {
            Rd[i] = I + std::sin(theta) / theta * O[i] + (1 - std::cos(theta)) / (theta * theta) * O2[i];
            V[i] = I + (1 - std::cos(theta)) / (theta * theta) * O[i] + (theta - std::sin(theta)) / std::pow(theta, 3) * O2[i];
        }
    }
    double qd[4];
    R_to_quat(Rd, qd);
    normalize_rotation(qd);
    const double td[3] = {V[0] * v[0] + V[1] * v[1] + V[2] * v[2], V[3] * v[0] + V[4] * v[1] + V[5] * v[2],
                          V[6] * v[0] + V[7] * v[1] + V[8] * v[2]};
    // result = (qd, td) * (q, t):  t' = td + Rd_q * t,  q' = qd * q   (se3quat.h operator*)
    double Rq[9];
    quat_to_R(qd, Rq);
    const double* t = T + 4;
    const double tn[3] = {td[0] + Rq[0] * t[0] + Rq[1] * t[1] + Rq[2] * t[2], td[1] + Rq[3] * t[0] + Rq[4] * t[1] + Rq[5] * t[2],
                          td[2] + Rq[6] * t[0] + Rq[7] * t[1] + Rq[8] * t[2]};
    const double ax = qd[0], ay = qd[1], az = qd[2], aw = qd[3], bx = T[0], by = T[1], bz = T[2], bw = T[3];
    double qn[4] = {aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                    aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz};
    normalize_rotation(qn);
    for (int i = 0; i < 4; ++i) T[i] = qn[i];
    for (int i = 0; i < 3; ++i) T[4 + i] = tn[i];
}

inline void se3_map(const double* T, const double* X, double* Xc) {  // q * X + t, Eigen _transformVector
    const double qx = T[0], qy = T[1], qz = T[2], qw = T[3];
    double uv[3] = {qy * X[2] - qz * X[1], qz * X[0] - qx * X[2], qx * X[1] - qy * X[0]};
    for (int i = 0; i < 3; ++i) uv[i] += uv[i];
    const double c[3] = {qy * uv[2] - qz * uv[1], qz * uv[0] - qx * uv[2], qx * uv[1] - qy * uv[0]};
    for (int i = 0; i < 3; ++i) Xc[i] = X[i] + qw * uv[i] + c[i] + T[4 + i];
}

struct Problem {
    int nKF, nMP, nE;
    std::vector<double> pose, point;  // 7 / 3 per vertex
    const uint8_t* fixed;
    const int *ekf, *emp;
    const double *obs, *invs2;
    Cam cam;
    double deltaMono, deltaStereo, dsqrMono, dsqrStereo;
};

// The stereo projections divide as `const float invz = 1.0f/trans_xyz[2]` (types_six_dof_expmap.cpp:191, :340): the float literal is
// promoted, the division is a DOUBLE division and its quotient is rounded to float.  The binary edge of LocalBundleAdjustment then
// multiplies `bf*invz` with bf a `const float&` parameter -- a FLOAT product (:196) -- while the pose-only edge uses its double member
// (:345), a double product.  g_stereo_form = 1 reproduces what the device kernels compute instead (float(z) first, a float division,
// double products everywhere; csrc/lba.cu:105, csrc/poseopt.cu:62), so that the tests can bound what that costs
// (tests/test_oracle_vs_ref_g2o.py::test_device_stereo_form_*).  Pinned against the reference source by oracle/_ref part 7.
int g_stereo_form = 0;

// residual of edge e at the current estimate; returns dimension (2 mono, 3 stereo).  unary: the pose-only edges of PoseOptimization
inline int edge_error(const Problem& P, int e, double r[3], double Xc[3], bool unary = false) {
    se3_map(&P.pose[7 * P.ekf[e]], &P.point[3 * P.emp[e]], Xc);
    const double* z = &P.obs[3 * e];
    if (z[2] < 0) {  // mono: EdgeSE3ProjectXYZ::computeError, Pinhole::project(Vector3d)
        r[0] = z[0] - (P.cam.fx * Xc[0] / Xc[2] + P.cam.cx);
        r[1] = z[1] - (P.cam.fy * Xc[1] / Xc[2] + P.cam.cy);
        r[2] = 0;
        return 2;
    }
    const float invzf = g_stereo_form ? 1.0f / (float)Xc[2] : (float)(1.0 / Xc[2]);
    const double invz = (double)invzf;
    const double u = Xc[0] * invz * P.cam.fx + P.cam.cx;
    r[0] = z[0] - u;
    r[1] = z[1] - (Xc[1] * invz * P.cam.fy + P.cam.cy);
    const double disp = (unary || g_stereo_form) ? P.cam.bf * invz : (double)((float)P.cam.bf * invzf);
    r[2] = z[2] - (u - disp);
    return 3;
}

inline double huber_rho(double e, double delta, double dsqr, double* w) {
    if (e <= dsqr) { *w = 1.0; return e; }
    const double s = std::sqrt(e);
    *w = delta / s;
    return 2 * s * delta - dsqr;
}

// analytic Jacobians of one edge: A = d r / d X (D x 3), B = d r / d (omega, upsilon) (D x 6), row-major with
// 3 / 6 columns.  types_six_dof_expmap.cpp:228-275 (stereo), OptimizableTypes.cpp:175-197 + Pinhole.cpp:119-130 (mono)
inline void edge_jacobians(const Cam& cam, int D, const double R[9], const double Xc[3], double A[9], double B[18]) {
    const double fx = cam.fx, fy = cam.fy, bf = cam.bf;
    const double xx = Xc[0], yy = Xc[1], zz = Xc[2], z2 = zz * zz;
    if (D == 3) {
        for (int c = 0; c < 3; ++c) {
            A[c] = -fx * R[c] / zz + fx * xx * R[6 + c] / z2;
            A[3 + c] = -fy * R[3 + c] / zz + fy * yy * R[6 + c] / z2;
            A[6 + c] = A[c] - bf * R[6 + c] / z2;
        }
        B[0] = xx * yy / z2 * fx; B[1] = -(1 + (xx * xx / z2)) * fx; B[2] = yy / zz * fx; B[3] = -1. / zz * fx; B[4] = 0; B[5] = xx / z2 * fx;
        B[6] = (1 + yy * yy / z2) * fy; B[7] = -xx * yy / z2 * fy; B[8] = -xx / zz * fy; B[9] = 0; B[10] = -1. / zz * fy; B[11] = yy / z2 * fy;
        B[12] = B[0] - bf * yy / z2; B[13] = B[1] + bf * xx / z2; B[14] = B[2]; B[15] = B[3]; B[16] = 0; B[17] = B[5] - bf / z2;
    } else {
        // projectJac = -Pinhole::projectJac(Xc);  A = projectJac * R;  B = projectJac * SE3deriv
        const double J[6] = {-(fx / zz), -0.0, fx * xx / z2, -0.0, -(fy / zz), fy * yy / z2};
        for (int c = 0; c < 3; ++c) {
            A[c] = J[0] * R[c] + J[1] * R[3 + c] + J[2] * R[6 + c];
            A[3 + c] = J[3] * R[c] + J[4] * R[3 + c] + J[5] * R[6 + c];
        }
        const double S[18] = {0, zz, -yy, 1, 0, 0, -zz, 0, xx, 0, 1, 0, yy, -xx, 0, 0, 0, 1};
        for (int c = 0; c < 6; ++c) {
            B[c] = J[0] * S[c] + J[1] * S[6 + c] + J[2] * S[12 + c];
            B[6 + c] = J[3] * S[c] + J[4] * S[6 + c] + J[5] * S[12 + c];
        }
    }
}

// the pose-only edges' Jacobian B = d r / d (omega, upsilon).  The stereo one is written with reciprocals
// (EdgeStereoSE3ProjectXYZOnlyPose::linearizeOplus, types_six_dof_expmap.cpp:352-404: invz = 1.0/z, invz_2 = invz*invz), unlike the
// binary edge's quotients; the monocular one is the binary edge's (OptimizableTypes.cpp:58-73).
inline void pose_edge_jacobian(const Cam& cam, int D, const double R[9], const double Xc[3], double B[18]) {
    if (D == 2 || g_stereo_form) {
        double A[9];
        edge_jacobians(cam, D, R, Xc, A, B);
        return;
    }
    const double fx = cam.fx, fy = cam.fy, bf = cam.bf;
    const double x = Xc[0], y = Xc[1], invz = 1.0 / Xc[2], invz_2 = invz * invz;
    B[0] = x * y * invz_2 * fx; B[1] = -(1 + (x * x * invz_2)) * fx; B[2] = y * invz * fx; B[3] = -invz * fx; B[4] = 0; B[5] = x * invz_2 * fx;
    B[6] = (1 + y * y * invz_2) * fy; B[7] = -x * y * invz_2 * fy; B[8] = -x * invz * fy; B[9] = 0; B[10] = -invz * fy; B[11] = y * invz_2 * fy;
    B[12] = B[0] - bf * y * invz_2; B[13] = B[1] + bf * x * invz_2; B[14] = B[2]; B[15] = B[3]; B[16] = 0; B[17] = B[5] - bf * invz_2;
}

// dense LDL^T (no pivoting) of an n x n symmetric matrix given by its upper triangle (row-major full
// storage, lower part ignored).  Returns false on a zero / non-finite pivot (LinearSolverEigen fails).
bool ldlt_solve(std::vector<double>& A, int n, const double* b, double* x) {
    std::vector<double> L((size_t)n * n, 0.0), D(n);
    for (int j = 0; j < n; ++j) {
        double d = A[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) d -= L[(size_t)j * n + k] * L[(size_t)j * n + k] * D[k];
        if (!(std::fabs(d) > 0) || !std::isfinite(d)) return false;
        D[j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = A[(size_t)j * n + i];  // upper triangle: A(j,i), i > j
            for (int k = 0; k < j; ++k) s -= L[(size_t)i * n + k] * L[(size_t)j * n + k] * D[k];
            L[(size_t)i * n + j] = s / d;
        }
    }
    std::vector<double> y(n);
    for (int i = 0; i < n; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= L[(size_t)i * n + k] * y[k];
        y[i] = s;
    }
    for (int i = 0; i < n; ++i) y[i] /= D[i];
    for (int i = n - 1; i >= 0; --i) {
        double s = y[i];
        for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * x[k];
        x[i] = s;
    }
    return true;
}

inline bool inv3(const double* M, double* Mi) {
    const double a = M[0], b = M[1], c = M[2], d = M[3], e = M[4], f = M[5], g = M[6], h = M[7], i = M[8];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    const double id = 1.0 / det;
    Mi[0] = (e * i - f * h) * id; Mi[1] = (c * h - b * i) * id; Mi[2] = (b * f - c * e) * id;
    Mi[3] = (f * g - d * i) * id; Mi[4] = (a * i - c * g) * id; Mi[5] = (c * d - a * f) * id;
    Mi[6] = (d * h - e * g) * id; Mi[7] = (b * g - a * h) * id; Mi[8] = (a * e - b * d) * id;
    return std::isfinite(id);
}

}  // namespace

// ONE restatement of g2o's Levenberg-Marquardt control flow for every optimiser of this file: SparseOptimizer::optimize(maxIters)
// (sparse_optimizer.cpp:354-416) around OptimizationAlgorithmLevenberg::solve (optimization_algorithm_levenberg.cpp:61-172, with
// ORB-SLAM3's nBad stop), computeLambdaInit (:174-188, tau = 1e-5) and computeScale (:190-197).  oracle/_ref part 4 compiles those
// functions of the reference verbatim and runs them over LbaEngine: tests/test_oracle_vs_ref_g2o.py holds this template to them.
// Ops: compute_errors() -> robust chi2 (computeActiveErrors + activeRobustChi2), build_system(), solve(lambda) -> bool (setLambda +
// solve + restoreDiagonal), apply_update() (SparseOptimizer::update(solver->x())), push() / pop() / discard_top(), max_diagonal(),
// scale(lambda) = sum_j x_j (lambda x_j + b_j).
struct LmOutcome {
    int iters = 0, trials = 0;
    double lambda = -1, currentChi = 0, iniChi0 = 0, lastChi = 0;
};
template <class Ops>
static LmOutcome lm_optimize(Ops& E, int maxIters, double userLambdaInit, const volatile int* stop_flag) {
    LmOutcome R;
    double lambda = userLambdaInit > 0 ? userLambdaInit : -1, ni = 2;
    int nBad = 0;
    bool ok = true;
    for (int it = 0; it < maxIters && ok && !(stop_flag && *stop_flag); ++it) {
        double currentChi = E.compute_errors();
        R.lastChi = currentChi;
        if (it == 0) R.iniChi0 = currentChi;
        double tempChi = currentChi;
        const double iniChi = currentChi;
        E.build_system();
        if (it == 0) {
            if (!(userLambdaInit > 0)) lambda = 1e-5 * E.max_diagonal();
            ni = 2;
            nBad = 0;
        }
        double rho = 0;
        int qmax = 0;
        do {
            E.push();
            const bool ok2 = E.solve(lambda);
            E.apply_update();
            tempChi = E.compute_errors();
            R.lastChi = tempChi;
            if (!ok2) tempChi = std::numeric_limits<double>::max();
            rho = currentChi - tempChi;
            double scale = E.scale(lambda);
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && std::isfinite(tempChi)) {
                double alpha = 1. - std::pow((2 * rho - 1), 3);
                alpha = std::min(alpha, 2. / 3.);
                const double sf = std::max(1. / 3., alpha);
                lambda *= sf;
                ni = 2;
                currentChi = tempChi;
                E.discard_top();
            } else {
                lambda *= ni;
                ni *= 2;
                E.pop();
            }
            ++qmax;
            ++R.trials;
        } while (rho < 0 && qmax < 10 && !(stop_flag && *stop_flag));
        ++R.iters;
        R.currentChi = currentChi;
        R.lambda = lambda;
        if (qmax == 10 || rho == 0) { ok = false; break; }
        if ((iniChi - currentChi) * 1e3 < iniChi) ++nBad; else nBad = 0;
        if (nBad >= 3) { ok = false; break; }
    }
    R.lambda = lambda;
    return R;
}

struct LmFnOps {   // the same interface from closures (PoseOptimization and LocalInertialBA keep their state in locals)
    std::function<double()> errors_, max_diagonal_;
    std::function<void()> build_, update_, push_, pop_, discard_;
    std::function<bool(double)> solve_;
    std::function<double(double)> scale_;
    double compute_errors() { return errors_(); }
    void build_system() { build_(); }
    bool solve(double lambda) { return solve_(lambda); }
    void apply_update() { update_(); }
    void push() { push_(); }
    void pop() { pop_(); }
    void discard_top() { discard_(); }
    double max_diagonal() { return max_diagonal_(); }
    double scale(double lambda) { return scale_(lambda); }
};

// The numeric engine of LocalBundleAdjustment as g2o splits it between SparseOptimizer (errors, robust chi2, update, push / pop) and
// BlockSolver<6,3> + LinearSolverEigen (buildSystem, setLambda + solve + restoreDiagonal).  orc_lba drives it with the restated
// Levenberg loop below; oracle/_ref part 4 drives the SAME engine with g2o's own OptimizationAlgorithmLevenberg::solve and
// SparseOptimizer::optimize compiled verbatim (tests/test_oracle_vs_ref_g2o.py), which pins the loop.
struct LbaEngine {
    Problem P;
    int nKF, nMP, nE, nP, sp, sl;
    const int *ekf, *emp;
    const double* invs2;
    std::vector<int> pidx;
    std::vector<double> Hpp, Hll, Hpl, b, x, err;
    std::vector<std::vector<int>> byMp;   // edges grouped by landmark for the Schur complement
    std::vector<std::vector<double>> stackPose, stackPoint;   // SparseOptimizer::push / pop / discardTop

    LbaEngine(int nKF_, int nMP_, int nE_, const double* pose, const uint8_t* fixed, const double* point, const int* ekf_, const int* emp_,
              const double* obs, const double* invs2_, const double* cam5)
        : nKF(nKF_), nMP(nMP_), nE(nE_), ekf(ekf_), emp(emp_), invs2(invs2_) {
        P.nKF = nKF; P.nMP = nMP; P.nE = nE;
        P.pose.assign(pose, pose + 7 * (size_t)nKF);
        P.point.assign(point, point + 3 * (size_t)nMP);
        P.fixed = fixed; P.ekf = ekf; P.emp = emp; P.obs = obs; P.invs2 = invs2;
        P.cam = {cam5[0], cam5[1], cam5[2], cam5[3], cam5[4]};
        P.deltaMono = (double)(float)std::sqrt(5.991);   // Optimizer.cc:1957-1958 (const float)
        P.deltaStereo = (double)(float)std::sqrt(7.815);
        P.dsqrMono = (double)(float)(P.deltaMono * P.deltaMono);        // RobustKernelHuber::dsqr is a float
        P.dsqrStereo = (double)(float)(P.deltaStereo * P.deltaStereo);
        pidx.assign(nKF, -1);
        nP = 0;
        for (int k = 0; k < nKF; ++k) if (!fixed[k]) pidx[k] = nP++;
        sp = 6 * nP; sl = 3 * nMP;
        Hpp.assign((size_t)nP * 36, 0.0); Hll.assign((size_t)nMP * 9, 0.0); Hpl.assign((size_t)nE * 18, 0.0);
        b.assign(sp + sl, 0.0); x.assign(sp + sl, 0.0); err.assign(nE, 0.0);
        byMp.resize(nMP);
        for (int e = 0; e < nE; ++e) byMp[emp[e]].push_back(e);
    }

    double compute_errors() {  // computeActiveErrors + activeRobustChi2
        double chi = 0;
        for (int e = 0; e < nE; ++e) {
            double r[3], Xc[3];
            const int D = edge_error(P, e, r, Xc);
            const double c = invs2[e] * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
            err[e] = c;
            double w;
            chi += huber_rho(c, D == 2 ? P.deltaMono : P.deltaStereo, D == 2 ? P.dsqrMono : P.dsqrStereo, &w);
        }
        return chi;
    }

    void build_system() {  // BlockSolver::buildSystem (linearizeOplus + constructQuadraticForm)
        std::fill(Hpp.begin(), Hpp.end(), 0.0);
        std::fill(Hll.begin(), Hll.end(), 0.0);
        std::fill(Hpl.begin(), Hpl.end(), 0.0);
        std::fill(b.begin(), b.end(), 0.0);
        for (int e = 0; e < nE; ++e) {
            double r[3], Xc[3];
            const int D = edge_error(P, e, r, Xc);
            double R[9];
            quat_to_R(&P.pose[7 * ekf[e]], R);
            double A[9] = {0}, B[18] = {0};
            edge_jacobians(P.cam, D, R, Xc, A, B);
            const double c2 = invs2[e] * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
            double w;
            huber_rho(c2, D == 2 ? P.deltaMono : P.deltaStereo, D == 2 ? P.dsqrMono : P.dsqrStereo, &w);
            const double om = w * invs2[e];  // weighted information (scalar * I)
            const int mp = emp[e], pi = pidx[ekf[e]];
            double* hl = &Hll[9 * (size_t)mp];
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) {
                    double s = 0;
                    for (int d = 0; d < D; ++d) s += A[3 * d + i] * A[3 * d + j];
                    hl[3 * i + j] += om * s;
                }
                double s = 0;
                for (int d = 0; d < D; ++d) s += A[3 * d + i] * r[d];
                b[sp + 3 * mp + i] += -om * s;
            }
            if (pi >= 0) {
                double* hp = &Hpp[36 * (size_t)pi];
                double* hx = &Hpl[18 * (size_t)e];
                for (int i = 0; i < 6; ++i) {
                    for (int j = 0; j < 6; ++j) {
                        double s = 0;
                        for (int d = 0; d < D; ++d) s += B[6 * d + i] * B[6 * d + j];
                        hp[6 * i + j] += om * s;
                    }
                    for (int j = 0; j < 3; ++j) {
                        double s = 0;
                        for (int d = 0; d < D; ++d) s += B[6 * d + i] * A[3 * d + j];
                        hx[3 * i + j] = om * s;
                    }
                    double s = 0;
                    for (int d = 0; d < D; ++d) s += B[6 * d + i] * r[d];
                    b[6 * pi + i] += -om * s;
                }
            }
        }
    }

    bool solve(double lambda) {  // setLambda + BlockSolver::solve + restoreDiagonal
        const int n = sp;
        std::vector<double> Hs((size_t)n * n, 0.0), bs(b.begin(), b.begin() + n), Dinv((size_t)nMP * 9);
        for (int p = 0; p < nP; ++p)
            for (int i = 0; i < 6; ++i)
                for (int j = 0; j < 6; ++j) Hs[(size_t)(6 * p + i) * n + 6 * p + j] = Hpp[36 * (size_t)p + 6 * i + j] + (i == j ? lambda : 0.0);
        for (int l = 0; l < nMP; ++l) {
            double Dm[9];
            for (int i = 0; i < 9; ++i) Dm[i] = Hll[9 * (size_t)l + i] + ((i % 4 == 0) ? lambda : 0.0);
            double* Di = &Dinv[9 * (size_t)l];
            inv3(Dm, Di);
            const double* bl = &b[sp + 3 * l];
            const double db[3] = {Di[0] * bl[0] + Di[1] * bl[1] + Di[2] * bl[2], Di[3] * bl[0] + Di[4] * bl[1] + Di[5] * bl[2],
                                  Di[6] * bl[0] + Di[7] * bl[1] + Di[8] * bl[2]};
            for (int e1 : byMp[l]) {
                const int p1 = pidx[ekf[e1]];
                if (p1 < 0) continue;
                const double* B1 = &Hpl[18 * (size_t)e1];
                double BD[18];
                for (int i = 0; i < 6; ++i)
                    for (int j = 0; j < 3; ++j) BD[3 * i + j] = B1[3 * i] * Di[j] + B1[3 * i + 1] * Di[3 + j] + B1[3 * i + 2] * Di[6 + j];
                for (int i = 0; i < 6; ++i) bs[6 * p1 + i] -= B1[3 * i] * db[0] + B1[3 * i + 1] * db[1] + B1[3 * i + 2] * db[2];
                for (int e2 : byMp[l]) {
                    const int p2 = pidx[ekf[e2]];
                    if (p2 < p1) continue;  // upper block triangle only (block_solver.hpp:419-430)
                    const double* B2 = &Hpl[18 * (size_t)e2];
                    for (int i = 0; i < 6; ++i)
                        for (int j = 0; j < 6; ++j)
                            Hs[(size_t)(6 * p1 + i) * n + 6 * p2 + j] -= BD[3 * i] * B2[3 * j] + BD[3 * i + 1] * B2[3 * j + 1] + BD[3 * i + 2] * B2[3 * j + 2];
                }
            }
        }
        if (n > 0 && !ldlt_solve(Hs, n, bs.data(), x.data())) return false;
        for (int l = 0; l < nMP; ++l) {  // x_l = Dinv (b_l - B^T x_p)
            double c[3] = {b[sp + 3 * l], b[sp + 3 * l + 1], b[sp + 3 * l + 2]};
            for (int e : byMp[l]) {
                const int p = pidx[ekf[e]];
                if (p < 0) continue;
                const double* Bm = &Hpl[18 * (size_t)e];
                for (int j = 0; j < 3; ++j)
                    for (int i = 0; i < 6; ++i) c[j] -= Bm[3 * i + j] * x[6 * p + i];
            }
            const double* Di = &Dinv[9 * (size_t)l];
            for (int i = 0; i < 3; ++i) x[sp + 3 * l + i] = Di[3 * i] * c[0] + Di[3 * i + 1] * c[1] + Di[3 * i + 2] * c[2];
        }
        return true;
    }

    void update(const double* u) {  // SparseOptimizer::update: VertexSE3Expmap::oplusImpl / VertexSBAPointXYZ::oplusImpl
        for (int k = 0; k < nKF; ++k) if (pidx[k] >= 0) pose_oplus(&P.pose[7 * k], &u[6 * pidx[k]]);
        for (int l = 0; l < nMP; ++l) for (int i = 0; i < 3; ++i) P.point[3 * l + i] += u[sp + 3 * l + i];
    }
    void apply_update() { update(x.data()); }
    double scale(double lambda) const {
        double sc = 0;
        for (int j = 0; j < sp + sl; ++j) sc += x[j] * (lambda * x[j] + b[j]);
        return sc;
    }
    void push() { stackPose.push_back(P.pose); stackPoint.push_back(P.point); }
    void pop() { P.pose = stackPose.back(); P.point = stackPoint.back(); stackPose.pop_back(); stackPoint.pop_back(); }
    void discard_top() { stackPose.pop_back(); stackPoint.pop_back(); }
    double max_diagonal() const {  // OptimizationAlgorithmLevenberg::computeLambdaInit's scan over the vertices' Hessian diagonals
        double md = 0;
        for (int p = 0; p < nP; ++p) for (int j = 0; j < 6; ++j) md = std::max(std::fabs(Hpp[36 * (size_t)p + 7 * j]), md);
        for (int l = 0; l < nMP; ++l) for (int j = 0; j < 3; ++j) md = std::max(std::fabs(Hll[9 * (size_t)l + 4 * j]), md);
        return md;
    }
    // post-processing inputs, Optimizer.cc:2107-2150; stats as orc_lba documents them
    void finish(double* pose, double* point, double* edge_chi2, uint8_t* edge_depth_pos) {
        for (int e = 0; e < nE; ++e) {
            edge_chi2[e] = err[e];
            double Xc[3];
            se3_map(&P.pose[7 * ekf[e]], &P.point[3 * emp[e]], Xc);
            edge_depth_pos[e] = Xc[2] > 0.0 ? 1 : 0;
        }
        memcpy(pose, P.pose.data(), sizeof(double) * 7 * (size_t)nKF);
        memcpy(point, P.point.data(), sizeof(double) * 3 * (size_t)nMP);
    }
};

extern "C" {

// pose: nKF x 7 (qx qy qz qw tx ty tz), in/out.  fixed: nKF.  point: nMP x 3, in/out.  obs: nE x 3 (ur < 0 =>
// monocular edge).  invs2: nE (information = invSigma2 * I).  cam5: fx fy cx cy bf (float members promoted).
// stop_flag: polled like pbStopFlag.  Outputs: edge_chi2[nE] = e->chi2() as left by the LAST computeActiveErrors,
// edge_depth_pos[nE] at the final estimate, stats[0..] = {outer iterations, final lambda, final robust chi2,
// total LM trials, initial robust chi2}.  Returns number of outer iterations run.
int orc_lba(int nKF, int nMP, int nE, double* pose, const uint8_t* fixed, double* point, const int* ekf, const int* emp,
            const double* obs, const double* invs2, const double* cam5, double lambdaInit, int maxIters,
            const volatile int* stop_flag, double* edge_chi2, uint8_t* edge_depth_pos, double* stats) {
    LbaEngine E(nKF, nMP, nE, pose, fixed, point, ekf, emp, obs, invs2, cam5);
    const LmOutcome R = lm_optimize(E, maxIters, lambdaInit, stop_flag);
    E.finish(pose, point, edge_chi2, edge_depth_pos);
    if (stats) {
        stats[0] = R.iters; stats[1] = R.lambda; stats[2] = R.currentChi; stats[3] = R.trials; stats[4] = R.iniChi0;
    }
    return R.iters;
}

// ---- the engine behind a C interface, for oracle/_ref part 4 (g2o's own Levenberg loop drives it) ---------------------------------
void* orc_lba_engine_create(int nKF, int nMP, int nE, const double* pose, const uint8_t* fixed, const double* point, const int* ekf, const int* emp,
                            const double* obs, const double* invs2, const double* cam5) {
    return new LbaEngine(nKF, nMP, nE, pose, fixed, point, ekf, emp, obs, invs2, cam5);
}
void orc_lba_engine_destroy(void* h) { delete (LbaEngine*)h; }
double orc_lba_engine_errors(void* h) { return ((LbaEngine*)h)->compute_errors(); }
void orc_lba_engine_build(void* h) { ((LbaEngine*)h)->build_system(); }
int orc_lba_engine_solve(void* h, double lambda) { return ((LbaEngine*)h)->solve(lambda) ? 1 : 0; }
void orc_lba_engine_update(void* h, const double* u) { ((LbaEngine*)h)->update(u); }
void orc_lba_engine_push(void* h) { ((LbaEngine*)h)->push(); }
void orc_lba_engine_pop(void* h) { ((LbaEngine*)h)->pop(); }
void orc_lba_engine_discard_top(void* h) { ((LbaEngine*)h)->discard_top(); }
double orc_lba_engine_max_diagonal(void* h) { return ((LbaEngine*)h)->max_diagonal(); }
int orc_lba_engine_vector_size(void* h) { return ((LbaEngine*)h)->sp + ((LbaEngine*)h)->sl; }
const double* orc_lba_engine_x(void* h) { return ((LbaEngine*)h)->x.data(); }
const double* orc_lba_engine_b(void* h) { return ((LbaEngine*)h)->b.data(); }
// the optimisable vertices as g2o indexes them for computeLambdaInit: free poses (dimension 6), then points (dimension 3)
int orc_lba_engine_vertices(void* h) { return ((LbaEngine*)h)->nP + ((LbaEngine*)h)->nMP; }
int orc_lba_engine_vertex_dim(void* h, int k) { return k < ((LbaEngine*)h)->nP ? 6 : 3; }
double orc_lba_engine_hessian_diag(void* h, int k, int j) {
    LbaEngine* E = (LbaEngine*)h;
    return k < E->nP ? E->Hpp[36 * (size_t)k + 7 * j] : E->Hll[9 * (size_t)(k - E->nP) + 4 * j];
}
void orc_lba_engine_finish(void* h, double* pose, double* point, double* edge_chi2, uint8_t* edge_depth_pos) {
    ((LbaEngine*)h)->finish(pose, point, edge_chi2, edge_depth_pos);
}
// RobustKernelHuber::robustify as the engine applies it: rho[0] and rho[1] for squared error e and delta = sqrt(chi2 threshold) as float
void orc_huber(double e, float delta_f, double* rho2) {
    const double delta = (double)delta_f, dsqr = (double)(float)(delta * delta);
    rho2[0] = huber_rho(e, delta, dsqr, &rho2[1]);
}
}

// ---- small exports for the finite-difference Jacobian test (tests/test_oracle_golden.py) -------------------
extern "C" {
// residual r (3, third = 0 for mono) of one edge
int orc_edge_residual(const double* pose7, const double* X, const double* obs3, const double* cam5, double* r) {
    Problem P;
    P.pose.assign(pose7, pose7 + 7);
    P.point.assign(X, X + 3);
    static const int zero = 0;
    P.ekf = &zero; P.emp = &zero; P.obs = obs3;
    P.cam = {cam5[0], cam5[1], cam5[2], cam5[3], cam5[4]};
    double Xc[3];
    return edge_error(P, 0, r, Xc);
}
// the same for a pose-only edge (PoseOptimization), with its Jacobian B (D x 6)
int orc_pose_edge(const double* pose7, const double* Xw, const double* obs3, const double* cam5, double* B18, double* r) {
    Problem P;
    P.pose.assign(pose7, pose7 + 7);
    P.point.assign(Xw, Xw + 3);
    static const int zero = 0;
    P.ekf = &zero; P.emp = &zero; P.obs = obs3;
    P.cam = {cam5[0], cam5[1], cam5[2], cam5[3], cam5[4]};
    double Xc[3], R[9];
    const int D = edge_error(P, 0, r, Xc, true);
    quat_to_R(pose7, R);
    for (int i = 0; i < 18; ++i) B18[i] = 0;
    pose_edge_jacobian(P.cam, D, R, Xc, B18);
    return D;
}
int orc_set_stereo_form(int device_form) { const int prev = g_stereo_form; g_stereo_form = device_form; return prev; }
void orc_pose_oplus(double* pose7, const double* upd6) { pose_oplus(pose7, upd6); }
void orc_edge_jacobians(const double* pose7, const double* X, int D, const double* cam5, double* A9, double* B18) {
    Cam cam = {cam5[0], cam5[1], cam5[2], cam5[3], cam5[4]};
    double R[9], Xc[3];
    quat_to_R(pose7, R);
    se3_map(pose7, X, Xc);
    for (int i = 0; i < 9; ++i) A9[i] = 0;
    for (int i = 0; i < 18; ++i) B18[i] = 0;
    edge_jacobians(cam, D, R, Xc, A9, B18);
}
}

// ------------------------------------------------------------------------------------------------
// Optimizer::PoseOptimization(Frame*)  src/Optimizer.cc:55-412, Nleft == -1 (one camera or rectified stereo).
// One VertexSE3Expmap, unary edges EdgeSE3ProjectXYZOnlyPose (OptimizableTypes.cpp) / EdgeStereoSE3ProjectXYZOnlyPose
// (Thirdparty/g2o types_six_dof_expmap.cpp:339-404), Levenberg over LinearSolverDense, 4 rounds of 10 iterations that
// restart from the frame's pose, chi2 classification (5.991 / 7.815, compared as floats) after every round, the robust
// kernel dropped after round 2.
//   Xw[n][3], obs[n][3] (obs[2] < 0 => monocular edge), invs2[n], cam5 = fx fy cx cy bf.
//   pose7 in: frame pose (qx qy qz qw tx ty tz), out: SE3quat_recov.  outlier[n] out: pFrame->mvbOutlier.
//   stats: {rounds run, total LM iterations, total LM trials, final lambda}.  Returns nInitialCorrespondences - nBad.
struct PoseEngine {   // the one-vertex problem of PoseOptimization as g2o splits it (see LbaEngine)
    Problem P;
    int n;
    std::vector<int> ekf, emp;
    const double *obs, *invs2;
    std::vector<double> err;
    std::vector<uint8_t> level, robust;
    double H[36], b[6], x[6];
    std::vector<std::vector<double>> stack;

    PoseEngine(int n_, const double* Xw, const double* obs_, const double* invs2_, const double* cam5, const double* pose7)
        : n(n_), ekf(n_, 0), emp(n_), obs(obs_), invs2(invs2_), err(n_, 0.0), level(n_, 0), robust(n_, 1) {
        P.nKF = 1; P.nMP = n; P.nE = n;
        P.pose.assign(pose7, pose7 + 7);
        P.point.assign(Xw, Xw + 3 * (size_t)n);
        for (int i = 0; i < n; ++i) emp[i] = i;
        P.fixed = nullptr; P.ekf = ekf.data(); P.emp = emp.data(); P.obs = obs; P.invs2 = invs2;
        P.cam = {cam5[0], cam5[1], cam5[2], cam5[3], cam5[4]};
        P.deltaMono = (double)(float)std::sqrt(5.991);          // Optimizer.cc:101-102 (const float)
        P.deltaStereo = (double)(float)std::sqrt(7.815);
        P.dsqrMono = (double)(float)(P.deltaMono * P.deltaMono);
        P.dsqrStereo = (double)(float)(P.deltaStereo * P.deltaStereo);
        std::fill(H, H + 36, 0.0); std::fill(b, b + 6, 0.0); std::fill(x, x + 6, 0.0);
    }
    double edge_chi2(int e) {
        double r[3], Xc[3];
        edge_error(P, e, r, Xc, true);
        return invs2[e] * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    }
    double compute_errors() {                  // computeActiveErrors + activeRobustChi2
        double chi = 0;
        for (int e = 0; e < n; ++e) {
            if (level[e]) continue;
            const double c = edge_chi2(e);
            err[e] = c;
            const bool mono = obs[3 * e + 2] < 0;
            double w;
            chi += robust[e] ? huber_rho(c, mono ? P.deltaMono : P.deltaStereo, mono ? P.dsqrMono : P.dsqrStereo, &w) : c;
        }
        return chi;
    }
    void build_system() {
        std::fill(H, H + 36, 0.0);
        std::fill(b, b + 6, 0.0);
        double R[9];
        quat_to_R(P.pose.data(), R);
        for (int e = 0; e < n; ++e) {
            if (level[e]) continue;
            double r[3], Xc[3], B[18] = {0};
            const int D = edge_error(P, e, r, Xc, true);
            pose_edge_jacobian(P.cam, D, R, Xc, B);
            const double c2 = invs2[e] * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
            double w = 1.0;
            if (robust[e]) huber_rho(c2, D == 2 ? P.deltaMono : P.deltaStereo, D == 2 ? P.dsqrMono : P.dsqrStereo, &w);
            const double om = w * invs2[e];
            for (int i = 0; i < 6; ++i) {
                for (int j = 0; j < 6; ++j) {
                    double s = 0;
                    for (int d = 0; d < D; ++d) s += B[6 * d + i] * B[6 * d + j];
                    H[6 * i + j] += om * s;
                }
                double s = 0;
                for (int d = 0; d < D; ++d) s += B[6 * d + i] * r[d];
                b[i] += -om * s;
            }
        }
    }
    bool solve(double lambda) {
        std::vector<double> Hl(H, H + 36);
        for (int j = 0; j < 6; ++j) Hl[7 * j] += lambda;
        return ldlt_solve(Hl, 6, b, x);
    }
    void apply_update() { pose_oplus(P.pose.data(), x); }
    void push() { stack.push_back(P.pose); }
    void pop() { P.pose = stack.back(); stack.pop_back(); }
    void discard_top() { stack.pop_back(); }
    double max_diagonal() const { double md = 0; for (int j = 0; j < 6; ++j) md = std::max(std::fabs(H[7 * j]), md); return md; }
    double scale(double lambda) const { double sc = 0; for (int j = 0; j < 6; ++j) sc += x[j] * (lambda * x[j] + b[j]); return sc; }
    int active() const { int a = 0; for (int e = 0; e < n; ++e) a += level[e] == 0; return a; }
};

extern "C" int orc_pose_optimization(int n, const double* Xw, const double* obs, const double* invs2, const double* cam5,
                                     double* pose7, uint8_t* outlier, double* stats) {
    if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
    if (n < 3) return 0;                                    // Optimizer.cc:292-293
    PoseEngine E(n, Xw, obs, invs2, cam5, pose7);
    const std::vector<double> pose0(pose7, pose7 + 7);
    for (int i = 0; i < n; ++i) outlier[i] = 0;
    int nBadEdges = 0, rounds = 0, totalIters = 0, totalTrials = 0;
    double lambda = -1;
    for (int it = 0; it < 4; ++it) {
        E.P.pose = pose0;                                    // vSE3->setEstimate(pFrame->GetPose())
        if (E.active() > 0) {                                // optimize(10): initializeOptimization(0) found the vertex
            const LmOutcome R = lm_optimize(E, 10, 0.0, nullptr);
            lambda = R.lambda;
            totalIters += R.iters;
            totalTrials += R.trials;
        }
        nBadEdges = 0;
        for (int e = 0; e < n; ++e) {
            if (outlier[e]) E.err[e] = E.edge_chi2(e);        // e->computeError() for the edges the optimiser did not touch
            const float chi2 = (float)E.err[e];
            const float th = obs[3 * e + 2] < 0 ? 5.991f : 7.815f;
            if (chi2 > th) { outlier[e] = 1; E.level[e] = 1; ++nBadEdges; }
            else { outlier[e] = 0; E.level[e] = 0; }
            if (it == 2) E.robust[e] = 0;                    // e->setRobustKernel(0)
        }
        ++rounds;
        if (n < 10) break;                                   // optimizer.edges().size() < 10
    }
    memcpy(pose7, E.P.pose.data(), sizeof(double) * 7);
    if (stats) { stats[0] = rounds; stats[1] = totalIters; stats[2] = totalTrials; stats[3] = lambda; }
    return n - nBadEdges;
}

// ---- PoseEngine behind a C interface, for oracle/_ref part 4 (the reference's own Optimizer::PoseOptimization drives it) ------------------
extern "C" {
void* orc_po_engine_create(int n, const double* Xw, const double* obs, const double* invs2, const double* cam5, const double* pose7) {
    return new PoseEngine(n, Xw, obs, invs2, cam5, pose7);
}
void orc_po_engine_destroy(void* h) { delete (PoseEngine*)h; }
void orc_po_engine_set_estimate(void* h, const double* pose7) { ((PoseEngine*)h)->P.pose.assign(pose7, pose7 + 7); }
void orc_po_engine_get_estimate(void* h, double* pose7) { memcpy(pose7, ((PoseEngine*)h)->P.pose.data(), sizeof(double) * 7); }
void orc_po_engine_set_level(void* h, int e, int level) { ((PoseEngine*)h)->level[e] = (uint8_t)level; }
void orc_po_engine_set_robust(void* h, int e, int on) { ((PoseEngine*)h)->robust[e] = on ? 1 : 0; }
void orc_po_engine_compute_error(void* h, int e) { PoseEngine* E = (PoseEngine*)h; E->err[e] = E->edge_chi2(e); }
double orc_po_engine_chi2(void* h, int e) { return ((PoseEngine*)h)->err[e]; }
int orc_po_engine_active(void* h) { return ((PoseEngine*)h)->active(); }
double orc_po_engine_errors(void* h) { return ((PoseEngine*)h)->compute_errors(); }
void orc_po_engine_build(void* h) { ((PoseEngine*)h)->build_system(); }
int orc_po_engine_solve(void* h, double lambda) { return ((PoseEngine*)h)->solve(lambda) ? 1 : 0; }
void orc_po_engine_update(void* h) { ((PoseEngine*)h)->apply_update(); }
void orc_po_engine_push(void* h) { ((PoseEngine*)h)->push(); }
void orc_po_engine_pop(void* h) { ((PoseEngine*)h)->pop(); }
void orc_po_engine_discard_top(void* h) { ((PoseEngine*)h)->discard_top(); }
double orc_po_engine_hessian_diag(void* h, int j) { return ((PoseEngine*)h)->H[7 * j]; }
const double* orc_po_engine_x(void* h) { return ((PoseEngine*)h)->x; }
const double* orc_po_engine_b(void* h) { return ((PoseEngine*)h)->b; }
}

// ------------------------------------------------------------------------------------------------
// Optimizer::LocalInertialBA numeric core  src/Optimizer.cc:2203-2812 (one camera, Nleft == -1) -- SURVEY 8(f) N2.
// The device implementation is csrc/liba_core.cuh (checked against this restatement by tests/test_liba_emul.py and, on a GPU,
// tests/test_liba_gpu.py); this restatement is pinned by finite-difference checks and fixed-point tests (tests/test_oracle_inertial.py).
//
// Vertices per keyframe: VertexPose (ImuCamPose: Rwb, twb; update twb += Rwb ut, Rwb = Rwb Exp(ur), G2oTypes.cc:221-244),
// VertexVelocity, VertexGyroBias, VertexAccBias (additive); a keyframe is either optimised or fixed as a whole.  Marginalised
// VertexSBAPointXYZ per map point.  Edges: EdgeMono / EdgeStereo (G2oTypes.h, .cc:390-490) with Huber sqrt(5.991) /
// sqrt(7.815); EdgeInertial (G2oTypes.cc:563-687) between consecutive keyframes, optionally Huber sqrt(16.92);
// EdgeGyroRW / EdgeAccRW (G2oTypes.h:736-800).  g2o Levenberg with a user lambda (1e0, or 1e-2 when bLarge), BlockSolverX
// + LinearSolverEigen: dense here, landmarks eliminated by the Schur complement.
// Inputs the caller computes exactly as the reference does: the 9 x 9 EdgeInertial information (inverse of C.block<9,9>,
// symmetrised, eigenvalues < 1e-12 clamped, times 1e-2 for the oldest link), InfoG / InfoA (inverse 3 x 3 blocks of C), the
// preintegrated deltas and bias Jacobians (float members of IMU::Preintegrated) and the bias they were linearised at.
// NormalizeRotation (an SVD re-orthonormalisation of matrices that are orthonormal to rounding) is omitted: it moves values by
// ~1e-16 (double) / ~1e-7 (the float delta rotation), far inside the 1e-4 parity bar.
namespace {

struct Mat3 { double m[9]; };
inline Mat3 mat3_mul(const Mat3& a, const Mat3& b) {
    Mat3 c;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c.m[3 * i + j] = a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j] + a.m[3 * i + 2] * b.m[6 + j];
    return c;
}
inline Mat3 mat3_T(const Mat3& a) { return Mat3{{a.m[0], a.m[3], a.m[6], a.m[1], a.m[4], a.m[7], a.m[2], a.m[5], a.m[8]}}; }
inline void mat3_vec(const Mat3& a, const double* v, double* o) {
    for (int i = 0; i < 3; ++i) o[i] = a.m[3 * i] * v[0] + a.m[3 * i + 1] * v[1] + a.m[3 * i + 2] * v[2];
}
inline Mat3 skew(const double* w) { return Mat3{{0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0}}; }
inline Mat3 mat3_I() { return Mat3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }

// G2oTypes.cc:908-986
inline Mat3 exp_so3(const double* w) {
    const double d2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], d = std::sqrt(d2);
    const Mat3 W = skew(w), W2 = mat3_mul(W, W);
    Mat3 r = mat3_I();
    if (d < 1e-5) { for (int i = 0; i < 9; ++i) r.m[i] += W.m[i] + 0.5 * W2.m[i]; }
    else { for (int i = 0; i < 9; ++i) r.m[i] += W.m[i] * std::sin(d) / d + W2.m[i] * (1.0 - std::cos(d)) / d2; }
    return r;
}
inline void log_so3(const Mat3& R, double* w) {
    const double tr = R.m[0] + R.m[4] + R.m[8];
    w[0] = (R.m[7] - R.m[5]) / 2; w[1] = (R.m[2] - R.m[6]) / 2; w[2] = (R.m[3] - R.m[1]) / 2;
    const double costheta = (tr - 1.0) * 0.5f;
    if (costheta > 1 || costheta < -1) return;
    const double theta = std::acos(costheta), s = std::sin(theta);
    if (std::fabs(s) < 1e-5) return;
    for (int i = 0; i < 3; ++i) w[i] = theta * w[i] / s;
}
inline Mat3 right_jac_so3(const double* v, bool inverse) {
    const double d2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], d = std::sqrt(d2);
    Mat3 r = mat3_I();
    if (d < 1e-5) return r;
    const Mat3 W = skew(v), W2 = mat3_mul(W, W);
    if (inverse) { for (int i = 0; i < 9; ++i) r.m[i] += W.m[i] / 2 + W2.m[i] * (1.0 / d2 - (1.0 + std::cos(d)) / (2.0 * d * std::sin(d))); }
    else { for (int i = 0; i < 9; ++i) r.m[i] += -W.m[i] * (1.0 - std::cos(d)) / d2 + W2.m[i] * (d - std::sin(d)) / (d2 * d); }
    return r;
}

struct InertialLink {   // one EdgeInertial + EdgeGyroRW + EdgeAccRW between keyframes k1 (previous) and k2
    int k1, k2, robust;
    double dt;
    float dR[9], dV[3], dP[3], JRg[9], JVg[9], JVa[9], JPg[9], JPa[9], blin[6];   // blin: bax bay baz bwx bwy bwz of the preintegration
    double info[81], infoG[9], infoA[9];
};

struct KfState { Mat3 Rwb; double twb[3], v[3], bg[3], ba[3]; };

struct InertialProblem {
    int nKF, nMP, nE, nL;
    std::vector<KfState> kf;
    std::vector<double> point;
    const uint8_t* fixed;
    const int *ekf, *emp;
    const double *obs, *invs2;
    const InertialLink* links;
    Mat3 Rcb, Rbc;
    double tcb[3], tbc[3];
    double fx, fy, cx, cy, bf;
};

inline void cam_pose(const InertialProblem& P, const KfState& s, Mat3& Rcw, double* tcw) {   // ImuCamPose::Update, G2oTypes.cc:236-243
    const Mat3 Rbw = mat3_T(s.Rwb);
    double tbw[3];
    mat3_vec(Rbw, s.twb, tbw);
    for (int i = 0; i < 3; ++i) tbw[i] = -tbw[i];
    Rcw = mat3_mul(P.Rcb, Rbw);
    mat3_vec(P.Rcb, tbw, tcw);
    for (int i = 0; i < 3; ++i) tcw[i] += P.tcb[i];
}

// reprojection residual (EdgeMono / EdgeStereo::computeError) and, optionally, Jacobians Jp (D x 3, point) and Jx (D x 6, pose)
inline int reproj(const InertialProblem& P, int e, double* r, double* Jp, double* Jx) {
    const KfState& s = P.kf[P.ekf[e]];
    Mat3 Rcw;
    double tcw[3], Xc[3];
    cam_pose(P, s, Rcw, tcw);
    mat3_vec(Rcw, &P.point[3 * P.emp[e]], Xc);
    for (int i = 0; i < 3; ++i) Xc[i] += tcw[i];
    const double* z = P.obs + 3 * e;
    const int D = z[2] < 0 ? 2 : 3;
    const double u = P.fx * Xc[0] / Xc[2] + P.cx, v = P.fy * Xc[1] / Xc[2] + P.cy;   // Pinhole::project(Vector3d)
    r[0] = z[0] - u;
    r[1] = z[1] - v;
    r[2] = 0;
    if (D == 3) r[2] = z[2] - (u - P.bf * (1 / Xc[2]));
    if (Jp) {
        double pj[9] = {P.fx / Xc[2], 0, -P.fx * Xc[0] / (Xc[2] * Xc[2]), 0, P.fy / Xc[2], -P.fy * Xc[1] / (Xc[2] * Xc[2]), 0, 0, 0};
        if (D == 3) { pj[6] = pj[0]; pj[7] = pj[1]; pj[8] = pj[2] + P.bf * (1.0 / (Xc[2] * Xc[2])); }
        double Xb[3];
        mat3_vec(P.Rbc, Xc, Xb);
        for (int i = 0; i < 3; ++i) Xb[i] += P.tbc[i];
        const double S[18] = {0, Xb[2], -Xb[1], 1, 0, 0, -Xb[2], 0, Xb[0], 0, 1, 0, Xb[1], -Xb[0], 0, 0, 0, 1};
        for (int d = 0; d < D; ++d) {
            for (int c = 0; c < 3; ++c) Jp[3 * d + c] = -(pj[3 * d] * Rcw.m[c] + pj[3 * d + 1] * Rcw.m[3 + c] + pj[3 * d + 2] * Rcw.m[6 + c]);
            double pr[3];   // proj_jac * Rcb
            for (int c = 0; c < 3; ++c) pr[c] = pj[3 * d] * P.Rcb.m[c] + pj[3 * d + 1] * P.Rcb.m[3 + c] + pj[3 * d + 2] * P.Rcb.m[6 + c];
            for (int c = 0; c < 6; ++c) Jx[6 * d + c] = pr[0] * S[c] + pr[1] * S[6 + c] + pr[2] * S[12 + c];
        }
    }
    return D;
}

// EdgeInertial::computeError / linearizeOplus.  e9 = (er, ev, ep).  J (optional): 9 x 30 over [pose1 6 | v1 3 | bg1 3 | ba1 3 | pose2 6 | v2 3 | 6 unused]
inline void inertial(const InertialProblem& P, const InertialLink& L, double* e9, double* J) {
    const KfState &s1 = P.kf[L.k1], &s2 = P.kf[L.k2];
    const double g[3] = {0, 0, -(double)9.81f};   // IMU::GRAVITY_VALUE is a float (ImuTypes.h:41)
    // IMU::Bias holds floats; the deltas are evaluated in float (ImuTypes.cc:388-430)
    float dbg[3], dba[3];
    for (int i = 0; i < 3; ++i) { dba[i] = (float)s1.ba[i] - L.blin[i]; dbg[i] = (float)s1.bg[i] - L.blin[3 + i]; }
    float wv[3];
    for (int i = 0; i < 3; ++i) wv[i] = L.JRg[3 * i] * dbg[0] + L.JRg[3 * i + 1] * dbg[1] + L.JRg[3 * i + 2] * dbg[2];
    const double wd[3] = {wv[0], wv[1], wv[2]};
    const Mat3 E = exp_so3(wd);
    Mat3 dR;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            dR.m[3 * i + j] = (double)(float)((double)L.dR[3 * i] * E.m[j] + (double)L.dR[3 * i + 1] * E.m[3 + j] + (double)L.dR[3 * i + 2] * E.m[6 + j]);
    double dV[3], dP[3];
    for (int i = 0; i < 3; ++i) {
        dV[i] = (double)(float)(L.dV[i] + (L.JVg[3 * i] * dbg[0] + L.JVg[3 * i + 1] * dbg[1] + L.JVg[3 * i + 2] * dbg[2]) +
                                (L.JVa[3 * i] * dba[0] + L.JVa[3 * i + 1] * dba[1] + L.JVa[3 * i + 2] * dba[2]));
        dP[i] = (double)(float)(L.dP[i] + (L.JPg[3 * i] * dbg[0] + L.JPg[3 * i + 1] * dbg[1] + L.JPg[3 * i + 2] * dbg[2]) +
                                (L.JPa[3 * i] * dba[0] + L.JPa[3 * i + 1] * dba[1] + L.JPa[3 * i + 2] * dba[2]));
    }
    const Mat3 Rbw1 = mat3_T(s1.Rwb);
    const Mat3 eR = mat3_mul(mat3_mul(mat3_T(dR), Rbw1), s2.Rwb);
    double er[3];
    log_so3(eR, er);
    const double dt = L.dt;
    double a[3], b[3], ra[3], rb[3];
    for (int i = 0; i < 3; ++i) {
        a[i] = s2.v[i] - s1.v[i] - g[i] * dt;
        b[i] = s2.twb[i] - s1.twb[i] - s1.v[i] * dt - g[i] * dt * dt / 2;
    }
    mat3_vec(Rbw1, a, ra);
    mat3_vec(Rbw1, b, rb);
    for (int i = 0; i < 3; ++i) { e9[i] = er[i]; e9[3 + i] = ra[i] - dV[i]; e9[6 + i] = rb[i] - dP[i]; }
    if (!J) return;
    std::fill(J, J + 9 * 30, 0.0);
    auto put = [&](int row0, int col0, const Mat3& M, double sgn) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) J[(row0 + i) * 30 + col0 + j] = sgn * M.m[3 * i + j];
    };
    const Mat3 invJr = right_jac_so3(er, true);
    put(0, 0, mat3_mul(mat3_mul(invJr, mat3_T(s2.Rwb)), s1.Rwb), -1.0);          // d er / d r1
    put(3, 0, skew(ra), 1.0);                                                   // d ev / d r1
    put(6, 0, skew(rb), 1.0);                                                   // d ep / d r1
    put(6, 3, mat3_I(), -1.0);                                                  // d ep / d t1
    put(3, 6, Rbw1, -1.0);                                                      // velocity 1
    { Mat3 m = Rbw1; for (double& x : m.m) x *= dt; put(6, 6, m, -1.0); }
    Mat3 JRgd, JVgd, JPgd, JVad, JPad;
    for (int i = 0; i < 9; ++i) { JRgd.m[i] = L.JRg[i]; JVgd.m[i] = L.JVg[i]; JPgd.m[i] = L.JPg[i]; JVad.m[i] = L.JVa[i]; JPad.m[i] = L.JPa[i]; }
    const double dbgd[3] = {dbg[0], dbg[1], dbg[2]};
    double jw[3];
    mat3_vec(JRgd, dbgd, jw);
    put(0, 9, mat3_mul(mat3_mul(mat3_mul(invJr, mat3_T(eR)), right_jac_so3(jw, false)), JRgd), -1.0);   // gyro bias 1
    put(3, 9, JVgd, -1.0);
    put(6, 9, JPgd, -1.0);
    put(3, 12, JVad, -1.0);                                                     // acc bias 1
    put(6, 12, JPad, -1.0);
    put(0, 15, invJr, 1.0);                                                     // pose 2
    put(6, 18, mat3_mul(Rbw1, s2.Rwb), 1.0);
    put(3, 21, Rbw1, 1.0);                                                      // velocity 2
}

inline void kf_oplus(KfState& s, const double* d15) {   // pose (6), velocity, gyro bias, acc bias
    double dt[3];
    mat3_vec(s.Rwb, d15 + 3, dt);
    for (int i = 0; i < 3; ++i) s.twb[i] += dt[i];
    s.Rwb = mat3_mul(s.Rwb, exp_so3(d15));
    for (int i = 0; i < 3; ++i) { s.v[i] += d15[6 + i]; s.bg[i] += d15[9 + i]; s.ba[i] += d15[12 + i]; }
}

}  // namespace

extern "C" {

// state: nKF x 21 doubles (Rwb row-major 9, twb 3, v 3, bg 3, ba 3), in/out.  fixed[nKF].  point nMP x 3 in/out.
// Tcb: Rcb (9) tcb (3).  cam5: fx fy cx cy bf.  links: nL records (see liba_link in pyoracle.py).  Outputs as orc_lba; stats[5] =
// optimizer.activeRobustChi2() as the reference reads it right after optimize() (Optimizer.cc:2685: the errors of the LAST trial, accepted
// or not); edge_depth_positive[e] = EdgeMono / EdgeStereo::isDepthPositive() at the final estimate (G2oTypes.cc:212-215).
int orc_liba(int nKF, int nMP, int nE, int nL, double* state, const uint8_t* fixed, double* point, const int* ekf, const int* emp,
             const double* obs, const double* invs2, const double* Tcb12, const double* cam5, const void* links_raw, double lambdaInit,
             int maxIters, double* edge_chi2, double* link_chi2, double* stats, uint8_t* edge_depth_positive) {
    InertialProblem P;
    P.nKF = nKF; P.nMP = nMP; P.nE = nE; P.nL = nL;
    P.kf.resize(nKF);
    for (int k = 0; k < nKF; ++k) {
        const double* s = state + 21 * k;
        for (int i = 0; i < 9; ++i) P.kf[k].Rwb.m[i] = s[i];
        for (int i = 0; i < 3; ++i) { P.kf[k].twb[i] = s[9 + i]; P.kf[k].v[i] = s[12 + i]; P.kf[k].bg[i] = s[15 + i]; P.kf[k].ba[i] = s[18 + i]; }
    }
    P.point.assign(point, point + 3 * (size_t)nMP);
    P.fixed = fixed; P.ekf = ekf; P.emp = emp; P.obs = obs; P.invs2 = invs2;
    P.links = static_cast<const InertialLink*>(links_raw);
    for (int i = 0; i < 9; ++i) P.Rcb.m[i] = Tcb12[i];
    for (int i = 0; i < 3; ++i) P.tcb[i] = Tcb12[9 + i];
    P.Rbc = mat3_T(P.Rcb);
    mat3_vec(P.Rbc, P.tcb, P.tbc);
    for (int i = 0; i < 3; ++i) P.tbc[i] = -P.tbc[i];
    P.fx = cam5[0]; P.fy = cam5[1]; P.cx = cam5[2]; P.cy = cam5[3]; P.bf = cam5[4];
    const double dM = (double)(float)std::sqrt(5.991), dS = (double)(float)std::sqrt(7.815), dI = std::sqrt(16.92);
    const double sqM = (double)(float)(dM * dM), sqS = (double)(float)(dS * dS), sqI = (double)(float)(dI * dI);

    std::vector<int> pidx(nKF, -1);
    int nP = 0;
    for (int k = 0; k < nKF; ++k) if (!fixed[k]) pidx[k] = nP++;
    const int sp = 15 * nP, sl = 3 * nMP;
    std::vector<double> Hpp((size_t)sp * sp), Hll((size_t)nMP * 9), W((size_t)nE * 18), b(sp + sl), x(sp + sl), err(nE, 0.0), lerr(nL * 3, 0.0);

    auto link_errors = [&](const InertialLink& L, double* c3) {   // chi2 of the inertial, gyro-RW and acc-RW edges
        double e9[9];
        inertial(P, L, e9, nullptr);
        double c = 0;
        for (int i = 0; i < 9; ++i) for (int j = 0; j < 9; ++j) c += e9[i] * L.info[9 * i + j] * e9[j];
        c3[0] = c;
        double cg = 0, ca = 0;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                cg += (P.kf[L.k2].bg[i] - P.kf[L.k1].bg[i]) * L.infoG[3 * i + j] * (P.kf[L.k2].bg[j] - P.kf[L.k1].bg[j]);
                ca += (P.kf[L.k2].ba[i] - P.kf[L.k1].ba[i]) * L.infoA[3 * i + j] * (P.kf[L.k2].ba[j] - P.kf[L.k1].ba[j]);
            }
        c3[1] = cg; c3[2] = ca;
    };
    auto compute_errors = [&]() -> double {
        double chi = 0;
        for (int e = 0; e < nE; ++e) {
            double r[3];
            const int D = reproj(P, e, r, nullptr, nullptr);
            const double c = invs2[e] * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
            err[e] = c;
            double w;
            chi += huber_rho(c, D == 2 ? dM : dS, D == 2 ? sqM : sqS, &w);
        }
        for (int l = 0; l < nL; ++l) {
            double c3[3];
            link_errors(P.links[l], c3);
            for (int i = 0; i < 3; ++i) lerr[3 * l + i] = c3[i];
            double w;
            chi += (P.links[l].robust ? huber_rho(c3[0], dI, sqI, &w) : c3[0]) + c3[1] + c3[2];
        }
        return chi;
    };
    auto add_pp = [&](int r0, int c0, int nr, int nc, const double* M, int ld) {
        for (int i = 0; i < nr; ++i) for (int j = 0; j < nc; ++j) Hpp[(size_t)(r0 + i) * sp + c0 + j] += M[i * ld + j];
    };
    auto build_system = [&]() {
        std::fill(Hpp.begin(), Hpp.end(), 0.0);
        std::fill(Hll.begin(), Hll.end(), 0.0);
        std::fill(W.begin(), W.end(), 0.0);
        std::fill(b.begin(), b.end(), 0.0);
        for (int e = 0; e < nE; ++e) {
            double r[3], Jp[9], Jx[18];
            const int D = reproj(P, e, r, Jp, Jx);
            const double c2 = invs2[e] * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
            double w;
            huber_rho(c2, D == 2 ? dM : dS, D == 2 ? sqM : sqS, &w);
            const double om = w * invs2[e];
            const int mp = emp[e], pi = pidx[ekf[e]];
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) { double s = 0; for (int d = 0; d < D; ++d) s += Jp[3 * d + i] * Jp[3 * d + j]; Hll[9 * (size_t)mp + 3 * i + j] += om * s; }
                double s = 0; for (int d = 0; d < D; ++d) s += Jp[3 * d + i] * r[d];
                b[sp + 3 * mp + i] += -om * s;
            }
            if (pi >= 0) {
                double M[36];
                for (int i = 0; i < 6; ++i) {
                    for (int j = 0; j < 6; ++j) { double s = 0; for (int d = 0; d < D; ++d) s += Jx[6 * d + i] * Jx[6 * d + j]; M[6 * i + j] = om * s; }
                    for (int j = 0; j < 3; ++j) { double s = 0; for (int d = 0; d < D; ++d) s += Jx[6 * d + i] * Jp[3 * d + j]; W[18 * (size_t)e + 3 * i + j] = om * s; }
                    double s = 0; for (int d = 0; d < D; ++d) s += Jx[6 * d + i] * r[d];
                    b[15 * pi + i] += -om * s;
                }
                add_pp(15 * pi, 15 * pi, 6, 6, M, 6);
            }
        }
        for (int l = 0; l < nL; ++l) {
            const InertialLink& L = P.links[l];
            double e9[9], J[9 * 30];
            inertial(P, L, e9, J);
            double c = 0;
            for (int i = 0; i < 9; ++i) for (int j = 0; j < 9; ++j) c += e9[i] * L.info[9 * i + j] * e9[j];
            double w = 1.0;
            if (L.robust) huber_rho(c, dI, sqI, &w);
            // columns: k1 -> [0, 15), k2 pose -> [15, 21), k2 velocity -> [21, 24)
            const int p1 = pidx[L.k1], p2 = pidx[L.k2];
            auto col_of = [&](int c30) -> int {
                if (c30 < 15) return p1 >= 0 ? 15 * p1 + c30 : -1;
                if (c30 < 21) return p2 >= 0 ? 15 * p2 + (c30 - 15) : -1;
                if (c30 < 24) return p2 >= 0 ? 15 * p2 + 6 + (c30 - 21) : -1;
                return -1;
            };
            double OJ[9 * 30], Oe[9];   // w * Info * J, w * Info * e
            for (int i = 0; i < 9; ++i) {
                double s = 0; for (int k = 0; k < 9; ++k) s += L.info[9 * i + k] * e9[k];
                Oe[i] = w * s;
                for (int cidx = 0; cidx < 24; ++cidx) { double t = 0; for (int k = 0; k < 9; ++k) t += L.info[9 * i + k] * J[k * 30 + cidx]; OJ[i * 30 + cidx] = w * t; }
            }
            for (int ca = 0; ca < 24; ++ca) {
                const int ga = col_of(ca);
                if (ga < 0) continue;
                double s = 0; for (int k = 0; k < 9; ++k) s += J[k * 30 + ca] * Oe[k];
                b[ga] += -s;
                for (int cb = 0; cb < 24; ++cb) {
                    const int gb = col_of(cb);
                    if (gb < 0) continue;
                    double t = 0; for (int k = 0; k < 9; ++k) t += J[k * 30 + ca] * OJ[k * 30 + cb];
                    Hpp[(size_t)ga * sp + gb] += t;
                }
            }
            // random walks: e = b2 - b1, J1 = -I, J2 = I
            for (int which = 0; which < 2; ++which) {
                const double* info = which == 0 ? L.infoG : L.infoA;
                const int off = which == 0 ? 9 : 12;
                double e3[3], Oe3[3];
                for (int i = 0; i < 3; ++i) e3[i] = which == 0 ? P.kf[L.k2].bg[i] - P.kf[L.k1].bg[i] : P.kf[L.k2].ba[i] - P.kf[L.k1].ba[i];
                for (int i = 0; i < 3; ++i) Oe3[i] = info[3 * i] * e3[0] + info[3 * i + 1] * e3[1] + info[3 * i + 2] * e3[2];
                for (int i = 0; i < 3; ++i) {
                    if (p1 >= 0) b[15 * p1 + off + i] += Oe3[i];      // -J1^T Omega e = +Omega e
                    if (p2 >= 0) b[15 * p2 + off + i] += -Oe3[i];
                    for (int j = 0; j < 3; ++j) {
                        if (p1 >= 0) Hpp[(size_t)(15 * p1 + off + i) * sp + 15 * p1 + off + j] += info[3 * i + j];
                        if (p2 >= 0) Hpp[(size_t)(15 * p2 + off + i) * sp + 15 * p2 + off + j] += info[3 * i + j];
                        if (p1 >= 0 && p2 >= 0) {
                            Hpp[(size_t)(15 * p1 + off + i) * sp + 15 * p2 + off + j] -= info[3 * i + j];
                            Hpp[(size_t)(15 * p2 + off + i) * sp + 15 * p1 + off + j] -= info[3 * i + j];
                        }
                    }
                }
            }
        }
    };
    std::vector<std::vector<int>> byPoint(nMP);
    for (int e = 0; e < nE; ++e) byPoint[emp[e]].push_back(e);
    auto solve = [&](double lambda) -> bool {
        std::vector<double> Hs(Hpp), bs(b.begin(), b.begin() + sp), Dinv((size_t)nMP * 9);
        for (int j = 0; j < sp; ++j) Hs[(size_t)j * sp + j] += lambda;
        for (int l = 0; l < nMP; ++l) {
            double D[9];
            for (int i = 0; i < 9; ++i) D[i] = Hll[9 * (size_t)l + i];
            D[0] += lambda; D[4] += lambda; D[8] += lambda;
            if (!inv3(D, &Dinv[9 * (size_t)l])) return false;
            const double* Di = &Dinv[9 * (size_t)l];
            const double* bl = &b[sp + 3 * l];
            double Dib[3];
            for (int i = 0; i < 3; ++i) Dib[i] = Di[3 * i] * bl[0] + Di[3 * i + 1] * bl[1] + Di[3 * i + 2] * bl[2];
            for (int e1 : byPoint[l]) {
                const int p1 = pidx[ekf[e1]];
                if (p1 < 0) continue;
                const double* W1 = &W[18 * (size_t)e1];
                double WD[18];   // W1 (6x3) * Dinv
                for (int i = 0; i < 6; ++i) for (int j = 0; j < 3; ++j) WD[3 * i + j] = W1[3 * i] * Di[j] + W1[3 * i + 1] * Di[3 + j] + W1[3 * i + 2] * Di[6 + j];
                for (int i = 0; i < 6; ++i) bs[15 * p1 + i] -= W1[3 * i] * Dib[0] + W1[3 * i + 1] * Dib[1] + W1[3 * i + 2] * Dib[2];
                for (int e2 : byPoint[l]) {
                    const int p2 = pidx[ekf[e2]];
                    if (p2 < 0) continue;
                    const double* W2 = &W[18 * (size_t)e2];
                    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j)
                        Hs[(size_t)(15 * p1 + i) * sp + 15 * p2 + j] -= WD[3 * i] * W2[3 * j] + WD[3 * i + 1] * W2[3 * j + 1] + WD[3 * i + 2] * W2[3 * j + 2];
                }
            }
        }
        if (sp > 0 && !ldlt_solve(Hs, sp, bs.data(), x.data())) return false;
        for (int l = 0; l < nMP; ++l) {
            double c[3] = {b[sp + 3 * l], b[sp + 3 * l + 1], b[sp + 3 * l + 2]};
            for (int e : byPoint[l]) {
                const int p = pidx[ekf[e]];
                if (p < 0) continue;
                const double* Wm = &W[18 * (size_t)e];
                for (int j = 0; j < 3; ++j) for (int i = 0; i < 6; ++i) c[j] -= Wm[3 * i + j] * x[15 * p + i];
            }
            const double* Di = &Dinv[9 * (size_t)l];
            for (int i = 0; i < 3; ++i) x[sp + 3 * l + i] = Di[3 * i] * c[0] + Di[3 * i + 1] * c[1] + Di[3 * i + 2] * c[2];
        }
        return true;
    };

    std::vector<std::vector<KfState>> stackKf;
    std::vector<std::vector<double>> stackPoint;
    LmFnOps ops;
    ops.errors_ = [&]() { return compute_errors(); };
    ops.build_ = [&]() { build_system(); };
    ops.solve_ = [&](double lam) { return solve(lam); };
    ops.update_ = [&]() {
        for (int k = 0; k < nKF; ++k) if (pidx[k] >= 0) kf_oplus(P.kf[k], &x[15 * pidx[k]]);
        for (int l = 0; l < nMP; ++l) for (int i = 0; i < 3; ++i) P.point[3 * l + i] += x[sp + 3 * l + i];
    };
    ops.push_ = [&]() { stackKf.push_back(P.kf); stackPoint.push_back(P.point); };
    ops.pop_ = [&]() { P.kf = stackKf.back(); P.point = stackPoint.back(); stackKf.pop_back(); stackPoint.pop_back(); };
    ops.discard_ = [&]() { stackKf.pop_back(); stackPoint.pop_back(); };
    ops.max_diagonal_ = [&]() {
        double md = 0;
        for (int j = 0; j < sp; ++j) md = std::max(std::fabs(Hpp[(size_t)j * sp + j]), md);
        for (int l = 0; l < nMP; ++l) for (int j = 0; j < 3; ++j) md = std::max(std::fabs(Hll[9 * (size_t)l + 4 * j]), md);
        return md;
    };
    ops.scale_ = [&](double lam) { double sc = 0; for (int j = 0; j < sp + sl; ++j) sc += x[j] * (lam * x[j] + b[j]); return sc; };
    const LmOutcome R = lm_optimize(ops, maxIters, lambdaInit, nullptr);
    const int iters = R.iters, trials = R.trials;
    const double lambda = R.lambda, currentChi = R.currentChi, iniChi0 = R.iniChi0, lastChi = R.lastChi;
    for (int e = 0; e < nE; ++e) if (edge_chi2) edge_chi2[e] = err[e];
    for (int l = 0; l < 3 * nL; ++l) if (link_chi2) link_chi2[l] = lerr[l];
    for (int k = 0; k < nKF; ++k) {
        double* s = state + 21 * k;
        for (int i = 0; i < 9; ++i) s[i] = P.kf[k].Rwb.m[i];
        for (int i = 0; i < 3; ++i) { s[9 + i] = P.kf[k].twb[i]; s[12 + i] = P.kf[k].v[i]; s[15 + i] = P.kf[k].bg[i]; s[18 + i] = P.kf[k].ba[i]; }
    }
    memcpy(point, P.point.data(), sizeof(double) * 3 * (size_t)nMP);
    if (stats) { stats[0] = iters; stats[1] = lambda; stats[2] = currentChi; stats[3] = trials; stats[4] = iniChi0; stats[5] = lastChi; }
    if (edge_depth_positive)
        for (int e = 0; e < nE; ++e) {
            Mat3 Rcw;
            double tcw[3];
            cam_pose(P, P.kf[ekf[e]], Rcw, tcw);
            const double* X = &P.point[3 * emp[e]];
            edge_depth_positive[e] = (Rcw.m[6] * X[0] + Rcw.m[7] * X[1] + Rcw.m[8] * X[2] + tcw[2]) > 0.0 ? 1 : 0;
        }
    return iters;
}

// test hooks: one EdgeInertial evaluated at the given states (error 9, Jacobian 9 x 30), and the ImuCamPose update
int orc_inertial_link_size() { return (int)sizeof(InertialLink); }
void orc_inertial_edge(const double* state2x21, const void* link_raw, double* e9, double* J270) {
    InertialProblem P;
    P.kf.resize(2);
    for (int k = 0; k < 2; ++k) {
        const double* s = state2x21 + 21 * k;
        for (int i = 0; i < 9; ++i) P.kf[k].Rwb.m[i] = s[i];
        for (int i = 0; i < 3; ++i) { P.kf[k].twb[i] = s[9 + i]; P.kf[k].v[i] = s[12 + i]; P.kf[k].bg[i] = s[15 + i]; P.kf[k].ba[i] = s[18 + i]; }
    }
    InertialLink L = *static_cast<const InertialLink*>(link_raw);
    L.k1 = 0; L.k2 = 1;
    inertial(P, L, e9, J270);
}
void orc_kf_oplus(double* state21, const double* d15) {
    KfState s;
    for (int i = 0; i < 9; ++i) s.Rwb.m[i] = state21[i];
    for (int i = 0; i < 3; ++i) { s.twb[i] = state21[9 + i]; s.v[i] = state21[12 + i]; s.bg[i] = state21[15 + i]; s.ba[i] = state21[18 + i]; }
    kf_oplus(s, d15);
    for (int i = 0; i < 9; ++i) state21[i] = s.Rwb.m[i];
    for (int i = 0; i < 3; ++i) { state21[9 + i] = s.twb[i]; state21[12 + i] = s.v[i]; state21[15 + i] = s.bg[i]; state21[18 + i] = s.ba[i]; }
}
}

// test hook: one EdgeMono / EdgeStereo (inertial parametrisation) at a keyframe state: residual (3), d r / d X (9), d r / d pose (18)
extern "C" int orc_liba_reproj(const double* state21, const double* Xw, const double* obs3, const double* Tcb12, const double* cam5,
                               double* r3, double* Jp9, double* Jx18) {
    InertialProblem P;
    P.kf.resize(1);
    for (int i = 0; i < 9; ++i) P.kf[0].Rwb.m[i] = state21[i];
    for (int i = 0; i < 3; ++i) { P.kf[0].twb[i] = state21[9 + i]; P.kf[0].v[i] = 0; P.kf[0].bg[i] = 0; P.kf[0].ba[i] = 0; }
    P.point.assign(Xw, Xw + 3);
    const int ekf = 0, emp = 0;
    const double w = 1.0;
    P.ekf = &ekf; P.emp = &emp; P.obs = obs3; P.invs2 = &w;
    for (int i = 0; i < 9; ++i) P.Rcb.m[i] = Tcb12[i];
    for (int i = 0; i < 3; ++i) P.tcb[i] = Tcb12[9 + i];
    P.Rbc = mat3_T(P.Rcb);
    mat3_vec(P.Rbc, P.tcb, P.tbc);
    for (int i = 0; i < 3; ++i) P.tbc[i] = -P.tbc[i];
    P.fx = cam5[0]; P.fy = cam5[1]; P.cx = cam5[2]; P.cy = cam5[3]; P.bf = cam5[4];
    return reproj(P, 0, r3, Jp9, Jx18);
}

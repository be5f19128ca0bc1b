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

// residual of edge e at the current estimate; returns dimension (2 mono, 3 stereo)
inline int edge_error(const Problem& P, int e, double r[3], double Xc[3]) {
    se3_map(&P.pose[7 * P.ekf[e]], &P.point[3 * P.emp[e]], Xc);
    const double* z = &P.obs[3 * e];
    if (z[2] < 0) {  // mono: EdgeSE3ProjectXYZ::computeError, Pinhole::project(Vector3d)
        r[0] = z[0] - (P.cam.fx * Xc[0] / Xc[2] + P.cam.cx);
        r[1] = z[1] - (P.cam.fy * Xc[1] / Xc[2] + P.cam.cy);
        r[2] = 0;
        return 2;
    }
    const double invz = (double)(1.0f / (float)Xc[2]);  // `const float invz = 1.0f/trans_xyz[2]` (.cpp:191)
    const double u = Xc[0] * invz * P.cam.fx + P.cam.cx;
    r[0] = z[0] - u;
    r[1] = z[1] - (Xc[1] * invz * P.cam.fy + P.cam.cy);
    r[2] = z[2] - (u - P.cam.bf * invz);
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

extern "C" {

// pose: nKF x 7 (qx qy qz qw tx ty tz), in/out.  fixed: nKF.  point: nMP x 3, in/out.  obs: nE x 3 (ur < 0 =>
// monocular edge).  invs2: nE (information = invSigma2 * I).  cam5: fx fy cx cy bf (float members promoted).
// stop_flag: polled like pbStopFlag.  Outputs: edge_chi2[nE] = e->chi2() as left by the LAST computeActiveErrors,
// edge_depth_pos[nE] at the final estimate, stats[0..] = {outer iterations, final lambda, final robust chi2,
// total LM trials, initial robust chi2}.  Returns number of outer iterations run.
int orc_lba(int nKF, int nMP, int nE, double* pose, const uint8_t* fixed, double* point, const int* ekf, const int* emp,
            const double* obs, const double* invs2, const double* cam5, double lambdaInit, int maxIters,
            const volatile int* stop_flag, double* edge_chi2, uint8_t* edge_depth_pos, double* stats) {
    Problem P;
    P.nKF = nKF; P.nMP = nMP; P.nE = nE;
    P.pose.assign(pose, pose + 7 * (size_t)nKF);
    P.point.assign(point, point + 3 * (size_t)nMP);
    P.fixed = fixed; P.ekf = ekf; P.emp = emp; P.obs = obs; P.invs2 = invs2;
    P.cam = {cam5[0], cam5[1], cam5[2], cam5[3], cam5[4]};
    P.deltaMono = (double)(float)std::sqrt(5.991);   // Optimizer.cc:1957-1958 (const float)
    P.deltaStereo = (double)(float)std::sqrt(7.815);
    P.dsqrMono = (double)(float)(P.deltaMono * P.deltaMono);        // RobustKernelHuber::dsqr is a float
    P.dsqrStereo = (double)(float)(P.deltaStereo * P.deltaStereo);

    std::vector<int> pidx(nKF, -1);
    int nP = 0;
    for (int k = 0; k < nKF; ++k) if (!fixed[k]) pidx[k] = nP++;
    const int sp = 6 * nP, sl = 3 * nMP;
    std::vector<double> Hpp((size_t)nP * 36), Hll((size_t)nMP * 9), Hpl((size_t)nE * 18), b(sp + sl), x(sp + sl);
    std::vector<double> err(nE, 0.0);

    auto compute_errors = [&]() -> double {  // computeActiveErrors + activeRobustChi2
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
    };

    auto build_system = [&]() {  // BlockSolver::buildSystem (linearizeOplus + constructQuadraticForm)
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
    };

    // edges grouped by landmark for the Schur complement
    std::vector<std::vector<int>> byMp(nMP);
    for (int e = 0; e < nE; ++e) byMp[emp[e]].push_back(e);

    auto solve = [&](double lambda) -> bool {  // setLambda + BlockSolver::solve + restoreDiagonal
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
    };

    double lambda = -1, ni = 2;
    int nBad = 0, iters = 0, trials = 0;
    double currentChi = 0, iniChi0 = 0;
    bool ok = true;
    for (int it = 0; it < maxIters && ok && !(stop_flag && *stop_flag); ++it) {
        currentChi = compute_errors();
        if (it == 0) iniChi0 = currentChi;
        double tempChi = currentChi;
        const double iniChi = currentChi;
        build_system();
        if (it == 0) {
            if (lambdaInit > 0) lambda = lambdaInit;
            else {
                double md = 0;
                for (int p = 0; p < nP; ++p) for (int j = 0; j < 6; ++j) md = std::max(std::fabs(Hpp[36 * (size_t)p + 7 * j]), md);
                for (int l = 0; l < nMP; ++l) for (int j = 0; j < 3; ++j) md = std::max(std::fabs(Hll[9 * (size_t)l + 4 * j]), md);
                lambda = 1e-5 * md;
            }
            ni = 2;
            nBad = 0;
        }
        double rho = 0;
        int qmax = 0;
        do {
            const std::vector<double> savedPose = P.pose, savedPoint = P.point;  // push
            const bool ok2 = solve(lambda);
            for (int k = 0; k < nKF; ++k) if (pidx[k] >= 0) pose_oplus(&P.pose[7 * k], &x[6 * pidx[k]]);
            for (int l = 0; l < nMP; ++l) for (int i = 0; i < 3; ++i) P.point[3 * l + i] += x[sp + 3 * l + i];
            tempChi = compute_errors();
            if (!ok2) tempChi = std::numeric_limits<double>::max();
            rho = currentChi - tempChi;
            double scale = 0;
            for (int j = 0; j < sp + sl; ++j) scale += x[j] * (lambda * x[j] + b[j]);
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && std::isfinite(tempChi)) {
                double alpha = 1. - std::pow((2 * rho - 1), 3);
                alpha = std::min(alpha, 2. / 3.);
                const double sf = std::max(1. / 3., alpha);
                lambda *= sf;
                ni = 2;
                currentChi = tempChi;
            } else {
                lambda *= ni;
                ni *= 2;
                P.pose = savedPose;   // pop
                P.point = savedPoint;
            }
            ++qmax;
            ++trials;
        } while (rho < 0 && qmax < 10 && !(stop_flag && *stop_flag));
        ++iters;
        if (qmax == 10 || rho == 0) { ok = false; break; }
        if ((iniChi - currentChi) * 1e3 < iniChi) ++nBad; else nBad = 0;
        if (nBad >= 3) { ok = false; break; }
    }
    // post-processing inputs, Optimizer.cc:2107-2150
    for (int e = 0; e < nE; ++e) {
        edge_chi2[e] = err[e];
        double Xc[3];
        se3_map(&P.pose[7 * ekf[e]], &P.point[3 * emp[e]], Xc);
        edge_depth_pos[e] = Xc[2] > 0.0 ? 1 : 0;
    }
    memcpy(pose, P.pose.data(), sizeof(double) * 7 * (size_t)nKF);
    memcpy(point, P.point.data(), sizeof(double) * 3 * (size_t)nMP);
    if (stats) {
        stats[0] = iters; stats[1] = lambda; stats[2] = currentChi; stats[3] = trials; stats[4] = iniChi0;
    }
    return iters;
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
extern "C" int orc_pose_optimization(int n, const double* Xw, const double* obs, const double* invs2, const double* cam5,
                                     double* pose7, uint8_t* outlier, double* stats) {
    if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
    if (n < 3) return 0;                                    // Optimizer.cc:292-293
    Problem P;
    P.nKF = 1; P.nMP = n; P.nE = n;
    P.pose.assign(pose7, pose7 + 7);
    P.point.assign(Xw, Xw + 3 * (size_t)n);
    std::vector<int> ekf(n, 0), emp(n);
    for (int i = 0; i < n; ++i) emp[i] = i;
    P.fixed = nullptr; P.ekf = ekf.data(); P.emp = emp.data(); P.obs = obs; P.invs2 = invs2;
    P.cam = {cam5[0], cam5[1], cam5[2], cam5[3], cam5[4]};
    P.deltaMono = (double)(float)std::sqrt(5.991);          // Optimizer.cc:101-102 (const float)
    P.deltaStereo = (double)(float)std::sqrt(7.815);
    P.dsqrMono = (double)(float)(P.deltaMono * P.deltaMono);
    P.dsqrStereo = (double)(float)(P.deltaStereo * P.deltaStereo);
    std::vector<double> pose0(pose7, pose7 + 7), err(n, 0.0);
    std::vector<uint8_t> level(n, 0);
    bool robust = true;
    for (int i = 0; i < n; ++i) outlier[i] = 0;
    double H[36], b[6], x[6];

    auto edge_chi2 = [&](int e) {
        double r[3], Xc[3];
        edge_error(P, e, r, Xc);
        return invs2[e] * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    };
    auto compute_errors = [&]() -> double {                  // computeActiveErrors + activeRobustChi2
        double chi = 0;
        for (int e = 0; e < n; ++e) {
            if (level[e]) continue;
            const double c = edge_chi2(e);
            err[e] = c;
            const bool mono = obs[3 * e + 2] < 0;
            double w;
            chi += robust ? huber_rho(c, mono ? P.deltaMono : P.deltaStereo, mono ? P.dsqrMono : P.dsqrStereo, &w) : c;
        }
        return chi;
    };
    auto build_system = [&]() {
        std::fill(H, H + 36, 0.0);
        std::fill(b, b + 6, 0.0);
        double R[9];
        quat_to_R(P.pose.data(), R);
        for (int e = 0; e < n; ++e) {
            if (level[e]) continue;
            double r[3], Xc[3], A[9] = {0}, B[18] = {0};
            const int D = edge_error(P, e, r, Xc);
            edge_jacobians(P.cam, D, R, Xc, A, B);
            const double c2 = invs2[e] * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
            double w = 1.0;
            if (robust) huber_rho(c2, D == 2 ? P.deltaMono : P.deltaStereo, D == 2 ? P.dsqrMono : P.dsqrStereo, &w);
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
    };

    int nBadEdges = 0, rounds = 0, totalIters = 0, totalTrials = 0;
    double lambda = -1;
    for (int it = 0; it < 4; ++it) {
        P.pose = pose0;                                      // vSE3->setEstimate(pFrame->GetPose())
        int nActive = 0;
        for (int e = 0; e < n; ++e) nActive += level[e] == 0;
        if (nActive > 0) {                                   // optimize(10): initializeOptimization(0) found the vertex
            double ni = 2, currentChi = 0;
            int nBadIt = 0;
            bool ok = true;
            for (int iter = 0; iter < 10 && ok; ++iter) {
                currentChi = compute_errors();
                double tempChi = currentChi;
                const double iniChi = currentChi;
                build_system();
                if (iter == 0) {
                    double md = 0;
                    for (int j = 0; j < 6; ++j) md = std::max(std::fabs(H[7 * j]), md);
                    lambda = 1e-5 * md;
                    ni = 2;
                    nBadIt = 0;
                }
                double rho = 0;
                int qmax = 0;
                do {
                    const std::vector<double> saved = P.pose;    // push
                    std::vector<double> Hl(H, H + 36);
                    for (int j = 0; j < 6; ++j) Hl[7 * j] += lambda;
                    const bool ok2 = ldlt_solve(Hl, 6, b, x);
                    pose_oplus(P.pose.data(), x);
                    tempChi = compute_errors();
                    if (!ok2) tempChi = std::numeric_limits<double>::max();
                    rho = currentChi - tempChi;
                    double scale = 0;
                    for (int j = 0; j < 6; ++j) scale += x[j] * (lambda * x[j] + b[j]);
                    scale += 1e-3;
                    rho /= scale;
                    if (rho > 0 && std::isfinite(tempChi)) {
                        double alpha = 1. - std::pow((2 * rho - 1), 3);
                        alpha = std::min(alpha, 2. / 3.);
                        lambda *= std::max(1. / 3., alpha);
                        ni = 2;
                        currentChi = tempChi;
                    } else {
                        lambda *= ni;
                        ni *= 2;
                        P.pose = saved;                          // pop
                    }
                    ++qmax;
                    ++totalTrials;
                } while (rho < 0 && qmax < 10);
                ++totalIters;
                if (qmax == 10 || rho == 0) { ok = false; break; }
                if ((iniChi - currentChi) * 1e3 < iniChi) ++nBadIt; else nBadIt = 0;
                if (nBadIt >= 3) { ok = false; break; }
            }
        }
        nBadEdges = 0;
        for (int e = 0; e < n; ++e) {
            if (outlier[e]) err[e] = edge_chi2(e);            // e->computeError() for the edges the optimiser did not touch
            const float chi2 = (float)err[e];
            const float th = obs[3 * e + 2] < 0 ? 5.991f : 7.815f;
            if (chi2 > th) { outlier[e] = 1; level[e] = 1; ++nBadEdges; }
            else { outlier[e] = 0; level[e] = 0; }
        }
        if (it == 2) robust = false;
        ++rounds;
        if (n < 10) break;                                   // optimizer.edges().size() < 10
    }
    memcpy(pose7, P.pose.data(), sizeof(double) * 7);
    if (stats) { stats[0] = rounds; stats[1] = totalIters; stats[2] = totalTrials; stats[3] = lambda; }
    return n - nBadEdges;
}

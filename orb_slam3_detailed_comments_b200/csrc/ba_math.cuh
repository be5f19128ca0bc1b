// ba_math.cuh -- fp64 device helpers shared by the local bundle adjustment (lba.cu) and the pose-only optimiser
// (poseopt.cu): g2o::SE3Quat map / exponential-map update, Huber kernel, pose Jacobians of the projection edges, and
// the dense 6 x 6 LDL^T solve of g2o::LinearSolverDense.
#pragma once
#include <cuda_runtime.h>

namespace orb {

__device__ __forceinline__ void quat_to_R(const double* q, double R[9]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}

__device__ __forceinline__ void se3_map(const double* T, const double* X, double* Xc) {
    const double qx = T[0], qy = T[1], qz = T[2], qw = T[3];
    double uv0 = qy * X[2] - qz * X[1], uv1 = qz * X[0] - qx * X[2], uv2 = qx * X[1] - qy * X[0];
    uv0 += uv0; uv1 += uv1; uv2 += uv2;
    Xc[0] = X[0] + qw * uv0 + (qy * uv2 - qz * uv1) + T[4];
    Xc[1] = X[1] + qw * uv1 + (qz * uv0 - qx * uv2) + T[5];
    Xc[2] = X[2] + qw * uv2 + (qx * uv1 - qy * uv0) + T[6];
}

__device__ __forceinline__ double huber_rho(double e, double delta, double dsqr, double* w) {
    if (e <= dsqr) { *w = 1.0; return e; }
    const double s = sqrt(e);
    *w = delta / s;
    return 2 * s * delta - dsqr;
}

// VertexSE3Expmap::oplusImpl: T <- exp(update) * T
__device__ inline void pose_oplus(double* T, const double* upd) {
    const double w0 = upd[0], w1 = upd[1], w2 = upd[2];
    const double theta = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
    const double O[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0};
    double O2[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) O2[3 * i + j] = O[3 * i] * O[j] + O[3 * i + 1] * O[3 + j] + O[3 * i + 2] * O[6 + j];
    double Rd[9], V[9];
    double a, b, c;
    if (theta < 0.00001) { a = 1; b = 1; c = 0; }
    else {
        a = sin(theta) / theta;
        b = (1 - cos(theta)) / (theta * theta);
        c = (theta - sin(theta)) / (theta * theta * theta);
    }
    for (int i = 0; i < 9; ++i) {
        const double I = (i % 4 == 0) ? 1.0 : 0.0;
        Rd[i] = I + a * O[i] + b * O2[i];
        V[i] = (theta < 0.00001) ? Rd[i] : (I + b * O[i] + c * O2[i]);
    }
    // Quaterniond(Rd)
    double qd[4];
    const double tr = Rd[0] + Rd[4] + Rd[8];
    if (tr > 0) {
        double s = sqrt(tr + 1.0);
        qd[3] = 0.5 * s;
        s = 0.5 / s;
        qd[0] = (Rd[7] - Rd[5]) * s; qd[1] = (Rd[2] - Rd[6]) * s; qd[2] = (Rd[3] - Rd[1]) * s;
    } else {
        int i = 0;
        if (Rd[4] > Rd[0]) i = 1;
        if (Rd[8] > Rd[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        double s = sqrt(Rd[4 * i] - Rd[4 * j] - Rd[4 * k] + 1.0);
        qd[i] = 0.5 * s;
        s = 0.5 / s;
        qd[3] = (Rd[3 * k + j] - Rd[3 * j + k]) * s;
        qd[j] = (Rd[3 * j + i] + Rd[3 * i + j]) * s;
        qd[k] = (Rd[3 * k + i] + Rd[3 * i + k]) * s;
    }
    {
        if (qd[3] < 0) for (int i = 0; i < 4; ++i) qd[i] = -qd[i];
        const double nrm = sqrt(qd[0] * qd[0] + qd[1] * qd[1] + qd[2] * qd[2] + qd[3] * qd[3]);
        for (int i = 0; i < 4; ++i) qd[i] /= nrm;
    }
    const double v0 = upd[3], v1 = upd[4], v2 = upd[5];
    const double td[3] = {V[0] * v0 + V[1] * v1 + V[2] * v2, V[3] * v0 + V[4] * v1 + V[5] * v2, V[6] * v0 + V[7] * v1 + V[8] * v2};
    double Rq[9];
    quat_to_R(qd, Rq);
    const double t0 = T[4], t1 = T[5], t2 = T[6];
    const double tn[3] = {td[0] + Rq[0] * t0 + Rq[1] * t1 + Rq[2] * t2, td[1] + Rq[3] * t0 + Rq[4] * t1 + Rq[5] * t2,
                          td[2] + Rq[6] * t0 + Rq[7] * t1 + Rq[8] * t2};
    const double ax = qd[0], ay = qd[1], az = qd[2], aw = qd[3], bx = T[0], by = T[1], bz = T[2], bw = T[3];
    double qn[4] = {aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                    aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz};
    if (qn[3] < 0) for (int i = 0; i < 4; ++i) qn[i] = -qn[i];
    const double nrm = sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
    for (int i = 0; i < 4; ++i) T[i] = qn[i] / nrm;
    T[4] = tn[0]; T[5] = tn[1]; T[6] = tn[2];
}


// d r / d (omega, upsilon) of one projection edge (rows of 6), r = obs - project(T X):
//   D == 3  EdgeStereoSE3ProjectXYZ[OnlyPose]::linearizeOplus  (types_six_dof_expmap.cpp:228-275, 375-404)
//   D == 2  -Pinhole::projectJac(Xc) * SE3deriv                 (OptimizableTypes.cpp:175-197, Pinhole.cpp:119-130)
__device__ __forceinline__ void pose_jacobian(int D, double fx, double fy, double bf, const double Xc[3], double B[18]) {
    const double xx = Xc[0], yy = Xc[1], zz = Xc[2];
    if (D == 3) {
        const double iz = 1.0 / zz, iz2 = iz * iz;
        B[0] = xx * yy * iz2 * fx; B[1] = -(1 + (xx * xx * iz2)) * fx; B[2] = yy * iz * fx; B[3] = -iz * fx; B[4] = 0; B[5] = xx * iz2 * fx;
        B[6] = (1 + yy * yy * iz2) * fy; B[7] = -xx * yy * iz2 * fy; B[8] = -xx * iz * fy; B[9] = 0; B[10] = -iz * fy; B[11] = yy * iz2 * fy;
        B[12] = B[0] - bf * yy * iz2; B[13] = B[1] + bf * xx * iz2; B[14] = B[2]; B[15] = B[3]; B[16] = 0; B[17] = B[5] - bf * iz2;
    } else {
        const double z2 = zz * zz;
        const double j0 = -(fx / zz), j2 = fx * xx / z2, j4 = -(fy / zz), j5 = fy * yy / z2;
        B[0] = j2 * yy;  B[1] = j0 * zz - j2 * xx; B[2] = -j0 * yy; B[3] = j0; B[4] = 0;  B[5] = j2;
        B[6] = -j4 * zz + j5 * yy; B[7] = -j5 * xx; B[8] = j4 * xx; B[9] = 0; B[10] = j4; B[11] = j5;
        B[12] = B[13] = B[14] = B[15] = B[16] = B[17] = 0;
    }
}

// (H + lambda I) x = b for a symmetric 6 x 6 H by LDL^T without pivoting (what Eigen's LDLT reduces to on a positive
// definite system, up to its pivot order); false on a zero / non-finite pivot.
__device__ __forceinline__ bool ldlt6_solve(const double H[36], double lambda, const double b[6], double x[6]) {
    double L[36], Dg[6], y[6];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = H[7 * j] + lambda;
#pragma unroll
        for (int k = 0; k < j; ++k) d -= L[6 * j + k] * L[6 * j + k] * Dg[k];
        if (!(fabs(d) > 0) || !isfinite(d)) { ok = false; d = 1.0; }
        Dg[j] = d;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double s = H[6 * j + i];
#pragma unroll
            for (int k = 0; k < j; ++k) s -= L[6 * i + k] * L[6 * j + k] * Dg[k];
            L[6 * i + j] = s / d;
        }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double s = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s -= L[6 * i + k] * y[k];
        y[i] = s;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] /= Dg[i];
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double s = y[i];
#pragma unroll
        for (int k = i + 1; k < 6; ++k) s -= L[6 * k + i] * x[k];
        x[i] = s;
    }
    return ok;
}

}  // namespace orb

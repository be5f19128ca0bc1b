// lba.cu -- Optimizer::LocalBundleAdjustment's numeric core on the device
// (/root/reference/src/Optimizer.cc:1859-2150: g2o OptimizationAlgorithmLevenberg over BlockSolver_6_3,
// Huber kernels, 10 iterations; math spec: SURVEY.md Appendix C).
//
// One CTA per problem runs the WHOLE Levenberg-Marquardt loop on the device (no host round trips):
//   E  per-edge reprojection residuals + robust chi2            (sparse_optimizer.cpp:61-113)
//   B  point-parallel linearisation: every thread owns map points, walks their edges, keeps H_ll / b_l in
//      registers, writes H_pl per edge and the per-edge pose terms; pose blocks are then reduced per pose in
//      a fixed order (deterministic -- LM accept/reject decisions must not depend on atomics ordering)
//      (base_binary_edge.hpp:55-120, types_six_dof_expmap.cpp:228-275, OptimizableTypes.cpp:175-197)
//   S  Schur complement: D^-1 per landmark, then one warp per upper pose-block pair sums B_i D^-1 B_j^T over
//      the landmarks the two poses share (block_solver.hpp:367-439)
//   L  dense LDL^T of the reduced pose system in shared memory (linear_solver_eigen.h:94-122: SimplicialLDLT;
//      at <= 24 poses the sparse ordering is irrelevant)
//   X  back-substitution of the landmarks, exp-map pose update (se3quat.h:223-255), gain ratio, lambda schedule
//      (optimization_algorithm_levenberg.cpp:61-169).
// All arithmetic is fp64 (the stereo edge keeps g2o's float32 reciprocal depth).
#include <algorithm>
#include <chrono>
#include <thread>
#include <vector>

#include "common.cuh"

using namespace orb;

namespace orb {

#define LBA_THREADS 1024
#define LBA_SMEM_N 144   // reduced systems up to 144 x 144 (24 free keyframes) are factorised in shared memory

struct LbaDev {
    int nKF, nMP, nE, nP, n;   // n = 6 * nP
    // inputs
    const double* pose0;   // [nKF][7] qx qy qz qw tx ty tz
    const double* point0;  // [nMP][3]
    const int* pidx;       // [nKF] pose index or -1 (fixed)
    const int* ekf;
    const int* emp;
    const double* obs;     // [nE][3], ur < 0 => monocular edge
    const double* invs2;
    const int* mpStart;    // [nMP + 1]
    const int* mpEdge;     // edges grouped by map point
    const int* poseStart;  // [nP + 1]
    const int* poseEdge;   // edges grouped by free pose
    const int* pairStart;  // [nPairBlocks + 1], block (p1 <= p2) at index blockIndex[p1 * nP + p2]
    const int2* pairItem;  // (e1, e2): same landmark, pose(e1) = p1, pose(e2) = p2
    const int* blockP1;    // [nPairBlocks]
    const int* blockP2;
    int nPairBlocks;
    double fx, fy, cx, cy, bf, lambdaInit;
    int maxIters;
    const volatile int* stop;   // mapped host flag or null
    // state + scratch
    double* pose;     // current estimate
    double* point;
    double* poseSave;
    double* pointSave;
    double* Hpl;      // [nE][18]
    double* W;        // [nE][27] per-edge pose terms (21 upper H + 6 b)
    double* Hll;      // [nMP][9]
    double* bl;       // [nMP][3]
    double* Dinv;     // [nMP][9]
    double* db;       // [nMP][3]
    double* Hpp;      // [nP][36]
    double* bp;       // [nP][6]
    double* xl;       // [nMP][3]
    double* HsGlobal; // [n][n] when n > LBA_SMEM_N
    double* err;      // [nE] chi2 per edge (last evaluation)
    // outputs
    uint8_t* depthPos;   // [nE]
    double* stats;       // iterations, lambda, chi2, trials, initial chi2
};

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

__device__ __forceinline__ int edge_error(const LbaDev& P, int e, double r[3], double Xc[3]) {
    se3_map(P.pose + 7 * P.ekf[e], P.point + 3 * P.emp[e], Xc);
    const double* z = P.obs + 3 * e;
    if (z[2] < 0) {   // EdgeSE3ProjectXYZ::computeError + Pinhole::project(Vector3d)
        r[0] = z[0] - (P.fx * Xc[0] / Xc[2] + P.cx);
        r[1] = z[1] - (P.fy * Xc[1] / Xc[2] + P.cy);
        r[2] = 0;
        return 2;
    }
    const double invz = (double)__fdiv_rn(1.0f, (float)Xc[2]);   // const float invz = 1.0f/trans_xyz[2]
    const double u = Xc[0] * invz * P.fx + P.cx;
    r[0] = z[0] - u;
    r[1] = z[1] - (Xc[1] * invz * P.fy + P.cy);
    r[2] = z[2] - (u - P.bf * invz);
    return 3;
}

struct Huber {
    double dM, dS, sqM, sqS;
};

__device__ __forceinline__ double huber_rho(double e, double delta, double dsqr, double* w) {
    if (e <= dsqr) { *w = 1.0; return e; }
    const double s = sqrt(e);
    *w = delta / s;
    return 2 * s * delta - dsqr;
}

// deterministic block sum (fixed tree); result broadcast to every thread
__device__ double block_sum(double v, double* red) {
    const int t = threadIdx.x;
    red[t] = v;
    __syncthreads();
    for (int s = LBA_THREADS / 2; s > 0; s >>= 1) {
        if (t < s) red[t] += red[t + s];
        __syncthreads();
    }
    const double out = red[0];
    __syncthreads();
    return out;
}

__device__ double block_max(double v, double* red) {
    const int t = threadIdx.x;
    red[t] = v;
    __syncthreads();
    for (int s = LBA_THREADS / 2; s > 0; s >>= 1) {
        if (t < s) red[t] = fmax(red[t], red[t + s]);
        __syncthreads();
    }
    const double out = red[0];
    __syncthreads();
    return out;
}

__device__ double compute_errors(const LbaDev& P, const Huber& H, double* red) {
    double chi = 0;
    for (int e = threadIdx.x; e < P.nE; e += LBA_THREADS) {
        double r[3], Xc[3];
        const int D = edge_error(P, e, r, Xc);
        const double c = P.invs2[e] * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        P.err[e] = c;
        double w;
        chi += huber_rho(c, D == 2 ? H.dM : H.dS, D == 2 ? H.sqM : H.sqS, &w);
    }
    return block_sum(chi, red);
}

// VertexSE3Expmap::oplusImpl: T <- exp(update) * T
__device__ void pose_oplus(double* T, const double* upd) {
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

// B phase, part 1: one thread per map point
__device__ void build_points(const LbaDev& P, const Huber& H) {
    for (int l = threadIdx.x; l < P.nMP; l += LBA_THREADS) {
        double hl[6] = {0, 0, 0, 0, 0, 0};   // symmetric 3x3: 00 01 02 11 12 22
        double b0 = 0, b1 = 0, b2 = 0;
        for (int k = P.mpStart[l]; k < P.mpStart[l + 1]; ++k) {
            const int e = P.mpEdge[k];
            double r[3], Xc[3];
            const int D = edge_error(P, e, r, Xc);
            const double xx = Xc[0], yy = Xc[1], zz = Xc[2], z2 = zz * zz;
            double R[9];
            quat_to_R(P.pose + 7 * P.ekf[e], R);
            double A[9], B[18];
            if (D == 3) {   // EdgeStereoSE3ProjectXYZ::linearizeOplus
                for (int c = 0; c < 3; ++c) {
                    A[c] = -P.fx * R[c] / zz + P.fx * xx * R[6 + c] / z2;
                    A[3 + c] = -P.fy * R[3 + c] / zz + P.fy * yy * R[6 + c] / z2;
                    A[6 + c] = A[c] - P.bf * R[6 + c] / z2;
                }
                B[0] = xx * yy / z2 * P.fx; B[1] = -(1 + (xx * xx / z2)) * P.fx; B[2] = yy / zz * P.fx;
                B[3] = -1. / zz * P.fx; B[4] = 0; B[5] = xx / z2 * P.fx;
                B[6] = (1 + yy * yy / z2) * P.fy; B[7] = -xx * yy / z2 * P.fy; B[8] = -xx / zz * P.fy;
                B[9] = 0; B[10] = -1. / zz * P.fy; B[11] = yy / z2 * P.fy;
                B[12] = B[0] - P.bf * yy / z2; B[13] = B[1] + P.bf * xx / z2; B[14] = B[2];
                B[15] = B[3]; B[16] = 0; B[17] = B[5] - P.bf / z2;
            } else {        // ORB_SLAM3::EdgeSE3ProjectXYZ::linearizeOplus with Pinhole::projectJac
                const double j0 = -(P.fx / zz), j2 = P.fx * xx / z2, j4 = -(P.fy / zz), j5 = P.fy * yy / z2;
                for (int c = 0; c < 3; ++c) {
                    A[c] = j0 * R[c] + j2 * R[6 + c];
                    A[3 + c] = j4 * R[3 + c] + j5 * R[6 + c];
                    A[6 + c] = 0;
                }
                B[0] = j2 * yy;  B[1] = j0 * zz - j2 * xx; B[2] = -j0 * yy; B[3] = j0; B[4] = 0;  B[5] = j2;
                B[6] = -j4 * zz + j5 * yy; B[7] = -j5 * xx; B[8] = j4 * xx; B[9] = 0;  B[10] = j4; B[11] = j5;
                for (int c = 12; c < 18; ++c) B[c] = 0;
            }
            const double c2 = P.invs2[e] * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
            double w;
            huber_rho(c2, D == 2 ? H.dM : H.dS, D == 2 ? H.sqM : H.sqS, &w);
            const double om = w * P.invs2[e];
            // H_ll += A^T om A ; b_l += -A^T om r
            hl[0] += om * (A[0] * A[0] + A[3] * A[3] + A[6] * A[6]);
            hl[1] += om * (A[0] * A[1] + A[3] * A[4] + A[6] * A[7]);
            hl[2] += om * (A[0] * A[2] + A[3] * A[5] + A[6] * A[8]);
            hl[3] += om * (A[1] * A[1] + A[4] * A[4] + A[7] * A[7]);
            hl[4] += om * (A[1] * A[2] + A[4] * A[5] + A[7] * A[8]);
            hl[5] += om * (A[2] * A[2] + A[5] * A[5] + A[8] * A[8]);
            b0 += -om * (A[0] * r[0] + A[3] * r[1] + A[6] * r[2]);
            b1 += -om * (A[1] * r[0] + A[4] * r[1] + A[7] * r[2]);
            b2 += -om * (A[2] * r[0] + A[5] * r[1] + A[8] * r[2]);
            if (P.pidx[P.ekf[e]] >= 0) {
                double* hx = P.Hpl + 18 * (size_t)e;
                double* wv = P.W + 27 * (size_t)e;
                int u = 0;
                for (int i = 0; i < 6; ++i) {
                    for (int j = 0; j < 3; ++j) hx[3 * i + j] = om * (B[i] * A[j] + B[6 + i] * A[3 + j] + B[12 + i] * A[6 + j]);
                    for (int j = i; j < 6; ++j) wv[u++] = om * (B[i] * B[j] + B[6 + i] * B[6 + j] + B[12 + i] * B[12 + j]);
                    wv[21 + i] = -om * (B[i] * r[0] + B[6 + i] * r[1] + B[12 + i] * r[2]);
                }
            }
        }
        double* o = P.Hll + 9 * (size_t)l;
        o[0] = hl[0]; o[1] = hl[1]; o[2] = hl[2]; o[3] = hl[1]; o[4] = hl[3]; o[5] = hl[4]; o[6] = hl[2]; o[7] = hl[4]; o[8] = hl[5];
        P.bl[3 * l] = b0; P.bl[3 * l + 1] = b1; P.bl[3 * l + 2] = b2;
    }
}

// B phase, part 2: H_pp / b_p = per-pose sums of the per-edge terms, fixed order (one warp per (pose, term))
__device__ void build_poses(const LbaDev& P) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = LBA_THREADS / 32;
    for (int task = warp; task < P.nP * 27; task += nw) {
        const int p = task / 27, c = task - p * 27;
        double s = 0;
        for (int k = P.poseStart[p] + lane; k < P.poseStart[p + 1]; k += 32) s += P.W[27 * (size_t)P.poseEdge[k] + c];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            if (c < 21) {
                int i = 0, rem = c;
                while (rem >= 6 - i) { rem -= 6 - i; ++i; }
                const int j = i + rem;
                P.Hpp[36 * (size_t)p + 6 * i + j] = s;
                P.Hpp[36 * (size_t)p + 6 * j + i] = s;
            } else {
                P.bp[6 * p + (c - 21)] = s;
            }
        }
    }
}

__global__ void __launch_bounds__(LBA_THREADS) k_lba(const LbaDev* __restrict__ problems) {
    extern __shared__ __align__(16) unsigned char lba_smem[];
    __shared__ LbaDev Ps;
    __shared__ int s_stop;
    if (threadIdx.x == 0) Ps = problems[blockIdx.x];
    __syncthreads();
    const LbaDev& P = Ps;
    double* red = reinterpret_cast<double*>(lba_smem);               // LBA_THREADS doubles
    double* xp = red + LBA_THREADS;                                   // n (solution) -- up to 960
    double* bs = xp + 960;                                            // n (rhs / scratch)
    double* Hs = (P.n <= LBA_SMEM_N) ? (bs + 960) : P.HsGlobal;       // n x n
    __shared__ int s_ok;
    const int tid = threadIdx.x, n = P.n;
    Huber H;
    H.dM = (double)(float)sqrt(5.991);     // Optimizer.cc:1957-1958: const float thHuberMono / thHuberStereo
    H.dS = (double)(float)sqrt(7.815);
    H.sqM = (double)(float)(H.dM * H.dM);  // RobustKernelHuber::dsqr is a float member
    H.sqS = (double)(float)(H.dS * H.dS);

    for (int i = tid; i < 7 * P.nKF; i += LBA_THREADS) P.pose[i] = P.pose0[i];
    for (int i = tid; i < 3 * P.nMP; i += LBA_THREADS) P.point[i] = P.point0[i];
    __syncthreads();

    double lambda = -1, ni = 2, currentChi = 0, iniChi0 = 0;
    int nBad = 0, iters = 0, trials = 0;
    bool ok = true;
    for (int it = 0; it < P.maxIters && ok; ++it) {
        if (tid == 0) s_stop = (P.stop && *P.stop) ? 1 : 0;   // one reader: the decision must be uniform
        __syncthreads();
        if (s_stop) break;
        currentChi = compute_errors(P, H, red);
        if (it == 0) iniChi0 = currentChi;
        const double iniChi = currentChi;
        double tempChi = currentChi;
        build_points(P, H);
        __syncthreads();
        build_poses(P);
        __syncthreads();
        if (it == 0) {
            if (P.lambdaInit > 0) lambda = P.lambdaInit;
            else {   // computeLambdaInit: tau * max |H_jj|
                double md = 0;
                for (int i = tid; i < 6 * P.nP; i += LBA_THREADS) md = fmax(md, fabs(P.Hpp[36 * (size_t)(i / 6) + 7 * (i % 6)]));
                for (int i = tid; i < 3 * P.nMP; i += LBA_THREADS) md = fmax(md, fabs(P.Hll[9 * (size_t)(i / 3) + 4 * (i % 3)]));
                lambda = 1e-5 * block_max(md, red);
            }
            ni = 2;
            nBad = 0;
        }
        double rho = 0;
        int qmax = 0;
        bool stopped = false;
        do {
            // push
            for (int i = tid; i < 7 * P.nKF; i += LBA_THREADS) P.poseSave[i] = P.pose[i];
            for (int i = tid; i < 3 * P.nMP; i += LBA_THREADS) P.pointSave[i] = P.point[i];
            // ---- S: Schur complement -------------------------------------------------------------
            for (int l = tid; l < P.nMP; l += LBA_THREADS) {
                const double* M = P.Hll + 9 * (size_t)l;
                const double a = M[0] + lambda, b = M[1], c = M[2], d = M[3], e = M[4] + lambda, f = M[5], g = M[6], h = M[7], i = M[8] + lambda;
                const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
                const double id = 1.0 / det;
                double* Di = P.Dinv + 9 * (size_t)l;
                Di[0] = (e * i - f * h) * id; Di[1] = (c * h - b * i) * id; Di[2] = (b * f - c * e) * id;
                Di[3] = (f * g - d * i) * id; Di[4] = (a * i - c * g) * id; Di[5] = (c * d - a * f) * id;
                Di[6] = (d * h - e * g) * id; Di[7] = (b * g - a * h) * id; Di[8] = (a * e - b * d) * id;
                const double* bl = P.bl + 3 * l;
                P.db[3 * l] = Di[0] * bl[0] + Di[1] * bl[1] + Di[2] * bl[2];
                P.db[3 * l + 1] = Di[3] * bl[0] + Di[4] * bl[1] + Di[5] * bl[2];
                P.db[3 * l + 2] = Di[6] * bl[0] + Di[7] * bl[1] + Di[8] * bl[2];
            }
            for (int i = tid; i < n * n; i += LBA_THREADS) Hs[i] = 0.0;
            __syncthreads();
            {
                const int warp = tid >> 5, lane = tid & 31, nw = LBA_THREADS / 32;
                for (int blk = warp; blk < P.nPairBlocks; blk += nw) {
                    const int p1 = P.blockP1[blk], p2 = P.blockP2[blk];
                    // lane handles entries lane and lane + 32 of the 6x6 block
                    const int e0 = lane, e1 = lane + 32;
                    const int i0 = e0 / 6, j0 = e0 % 6, i1 = (e1 < 36) ? e1 / 6 : 0, j1 = (e1 < 36) ? e1 % 6 : 0;
                    double acc0 = 0, acc1 = 0;
                    for (int k = P.pairStart[blk]; k < P.pairStart[blk + 1]; ++k) {
                        const int2 it2 = P.pairItem[k];
                        const double* B1 = P.Hpl + 18 * (size_t)it2.x;
                        const double* B2 = P.Hpl + 18 * (size_t)it2.y;
                        const double* Di = P.Dinv + 9 * (size_t)P.emp[it2.x];
                        {
                            const double t0 = B1[3 * i0] * Di[0] + B1[3 * i0 + 1] * Di[3] + B1[3 * i0 + 2] * Di[6];
                            const double t1 = B1[3 * i0] * Di[1] + B1[3 * i0 + 1] * Di[4] + B1[3 * i0 + 2] * Di[7];
                            const double t2 = B1[3 * i0] * Di[2] + B1[3 * i0 + 1] * Di[5] + B1[3 * i0 + 2] * Di[8];
                            acc0 += t0 * B2[3 * j0] + t1 * B2[3 * j0 + 1] + t2 * B2[3 * j0 + 2];
                        }
                        if (e1 < 36) {
                            const double t0 = B1[3 * i1] * Di[0] + B1[3 * i1 + 1] * Di[3] + B1[3 * i1 + 2] * Di[6];
                            const double t1 = B1[3 * i1] * Di[1] + B1[3 * i1 + 1] * Di[4] + B1[3 * i1 + 2] * Di[7];
                            const double t2 = B1[3 * i1] * Di[2] + B1[3 * i1 + 1] * Di[5] + B1[3 * i1 + 2] * Di[8];
                            acc1 += t0 * B2[3 * j1] + t1 * B2[3 * j1 + 1] + t2 * B2[3 * j1 + 2];
                        }
                    }
                    const double* hp = P.Hpp + 36 * (size_t)p1;
                    const bool diag = p1 == p2;
                    Hs[(size_t)(6 * p1 + i0) * n + 6 * p2 + j0] = (diag ? hp[6 * i0 + j0] + (i0 == j0 ? lambda : 0.0) : 0.0) - acc0;
                    if (e1 < 36) Hs[(size_t)(6 * p1 + i1) * n + 6 * p2 + j1] = (diag ? hp[6 * i1 + j1] + (i1 == j1 ? lambda : 0.0) : 0.0) - acc1;
                }
                // b_schur = b_p - sum_e B_e (D^-1 b_l)
                for (int p = warp; p < P.nP; p += nw) {
                    double s[6] = {0, 0, 0, 0, 0, 0};
                    for (int k = P.poseStart[p] + lane; k < P.poseStart[p + 1]; k += 32) {
                        const int e = P.poseEdge[k];
                        const double* Bm = P.Hpl + 18 * (size_t)e;
                        const double* d = P.db + 3 * P.emp[e];
                        for (int i = 0; i < 6; ++i) s[i] += Bm[3 * i] * d[0] + Bm[3 * i + 1] * d[1] + Bm[3 * i + 2] * d[2];
                    }
                    for (int i = 0; i < 6; ++i)
                        for (int o = 16; o > 0; o >>= 1) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
                    if (lane == 0)
                        for (int i = 0; i < 6; ++i) bs[6 * p + i] = P.bp[6 * p + i] - s[i];
                }
            }
            __syncthreads();
            // ---- L: LDL^T of the upper-stored reduced system, right-looking, in place ---------------
            // after step j: row j of Hs holds D_j on the diagonal and L(i,j) * D_j ... we keep U = D L^T:
            // Hs(j,i) (i > j) <- L(i,j); trailing update Hs(i,k) -= L(i,j) L(k,j) D_j for k >= i > j
            if (tid == 0) s_ok = 1;
            __syncthreads();
            for (int j = 0; j < n; ++j) {
                const double dj = Hs[(size_t)j * n + j];
                if (!(fabs(dj) > 0) || !isfinite(dj)) {
                    if (tid == 0) s_ok = 0;
                    break;   // uniform: dj is read by every thread from the same location
                }
                for (int i = j + 1 + tid; i < n; i += LBA_THREADS) red[i - j - 1] = Hs[(size_t)j * n + i] / dj;   // L(i,j), n-j-1 <= 959... uses red as scratch
                __syncthreads();
                const int m = n - j - 1;
                for (int idx = tid; idx < m * m; idx += LBA_THREADS) {
                    const int a = idx / m, b = idx - a * m;   // i = j+1+a, k = j+1+b, need k >= i
                    if (b >= a) Hs[(size_t)(j + 1 + a) * n + (j + 1 + b)] -= red[a] * red[b] * dj;
                }
                __syncthreads();
                for (int i = j + 1 + tid; i < n; i += LBA_THREADS) Hs[(size_t)j * n + i] = red[i - j - 1];
                __syncthreads();
            }
            __syncthreads();
            const bool ok2 = s_ok != 0;
            if (ok2) {
                // forward: y = L^-1 b ; then y /= D ; backward: x = L^-T y   (warp 0, lanes over the dot product)
                if (tid < 32) {
                    for (int i = 0; i < n; ++i) {
                        double s = 0;
                        for (int k = tid; k < i; k += 32) s += Hs[(size_t)k * n + i] * xp[k];
                        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                        if (tid == 0) xp[i] = bs[i] - s;
                        __syncwarp();
                    }
                    for (int i = tid; i < n; i += 32) xp[i] /= Hs[(size_t)i * n + i];
                    __syncwarp();
                    for (int i = n - 1; i >= 0; --i) {
                        double s = 0;
                        for (int k = i + 1 + tid; k < n; k += 32) s += Hs[(size_t)i * n + k] * xp[k];
                        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                        if (tid == 0) xp[i] -= s;
                        __syncwarp();
                    }
                }
            } else {
                for (int i = tid; i < n; i += LBA_THREADS) xp[i] = 0.0;
            }
            __syncthreads();
            // ---- X: landmarks x_l = D^-1 (b_l - B^T x_p) -------------------------------------------
            for (int l = tid; l < P.nMP; l += LBA_THREADS) {
                double c0 = P.bl[3 * l], c1 = P.bl[3 * l + 1], c2 = P.bl[3 * l + 2];
                for (int k = P.mpStart[l]; k < P.mpStart[l + 1]; ++k) {
                    const int e = P.mpEdge[k];
                    const int p = P.pidx[P.ekf[e]];
                    if (p < 0) continue;
                    const double* Bm = P.Hpl + 18 * (size_t)e;
                    const double* x = xp + 6 * p;
                    for (int i = 0; i < 6; ++i) {
                        c0 -= Bm[3 * i] * x[i];
                        c1 -= Bm[3 * i + 1] * x[i];
                        c2 -= Bm[3 * i + 2] * x[i];
                    }
                }
                const double* Di = P.Dinv + 9 * (size_t)l;
                double x0 = Di[0] * c0 + Di[1] * c1 + Di[2] * c2, x1 = Di[3] * c0 + Di[4] * c1 + Di[5] * c2, x2 = Di[6] * c0 + Di[7] * c1 + Di[8] * c2;
                if (!ok2) { x0 = x1 = x2 = 0.0; }
                P.xl[3 * l] = x0; P.xl[3 * l + 1] = x1; P.xl[3 * l + 2] = x2;
            }
            __syncthreads();
            // ---- update (SparseOptimizer::update) --------------------------------------------------
            for (int k = tid; k < P.nKF; k += LBA_THREADS)
                if (P.pidx[k] >= 0) pose_oplus(P.pose + 7 * k, xp + 6 * P.pidx[k]);
            for (int i = tid; i < 3 * P.nMP; i += LBA_THREADS) P.point[i] += P.xl[i];
            __syncthreads();
            tempChi = compute_errors(P, H, red);
            if (!ok2) tempChi = 1.7976931348623157e308;
            // computeScale: sum_j x_j (lambda x_j + b_j)
            double sc = 0;
            for (int i = tid; i < n; i += LBA_THREADS) sc += xp[i] * (lambda * xp[i] + P.bp[i]);
            for (int i = tid; i < 3 * P.nMP; i += LBA_THREADS) sc += P.xl[i] * (lambda * P.xl[i] + P.bl[i]);
            const double scale = block_sum(sc, red) + 1e-3;
            rho = (currentChi - tempChi) / scale;
            if (rho > 0 && isfinite(tempChi)) {
                double alpha = 1. - pow((2 * rho - 1), 3);
                alpha = fmin(alpha, 2. / 3.);
                lambda *= fmax(1. / 3., alpha);
                ni = 2;
                currentChi = tempChi;
            } else {
                lambda *= ni;
                ni *= 2;
                for (int i = tid; i < 7 * P.nKF; i += LBA_THREADS) P.pose[i] = P.poseSave[i];   // pop
                for (int i = tid; i < 3 * P.nMP; i += LBA_THREADS) P.point[i] = P.pointSave[i];
            }
            if (tid == 0) s_stop = (P.stop && *P.stop) ? 1 : 0;
            __syncthreads();
            ++qmax;
            ++trials;
            stopped = s_stop != 0;
        } while (rho < 0 && qmax < 10 && !stopped);
        ++iters;
        if (qmax == 10 || rho == 0) { ok = false; break; }
        if ((iniChi - currentChi) * 1e3 < iniChi) ++nBad; else nBad = 0;
        if (nBad >= 3) { ok = false; break; }
    }
    __syncthreads();
    for (int e = tid; e < P.nE; e += LBA_THREADS) {
        double Xc[3];
        se3_map(P.pose + 7 * P.ekf[e], P.point + 3 * P.emp[e], Xc);
        P.depthPos[e] = Xc[2] > 0.0 ? 1 : 0;
    }
    if (tid == 0) {
        P.stats[0] = iters; P.stats[1] = lambda; P.stats[2] = currentChi; P.stats[3] = trials; P.stats[4] = iniChi0;
    }
}

}  // namespace orb

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct lba_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint8_t* d_buf = nullptr;
    size_t d_bytes = 0;
    uint8_t* h_buf = nullptr;      // pinned staging
    size_t h_bytes = 0;
    int* h_stop = nullptr;         // mapped pinned flag the kernel polls
    int* d_stop = nullptr;
};

static orb_status lba_ensure(lba_handle* h, size_t dbytes, size_t hbytes) {
    if (dbytes > h->d_bytes) {
        if (h->d_buf) cudaFree(h->d_buf);
        h->d_buf = nullptr; h->d_bytes = 0;
        const size_t want = dbytes + dbytes / 4 + (1 << 20);
        ORB_CUDA(cudaMalloc((void**)&h->d_buf, want));
        h->d_bytes = want;
    }
    if (hbytes > h->h_bytes) {
        if (h->h_buf) cudaFreeHost(h->h_buf);
        h->h_buf = nullptr; h->h_bytes = 0;
        const size_t want = hbytes + hbytes / 4 + (1 << 20);
        ORB_CUDA(cudaMallocHost((void**)&h->h_buf, want));
        h->h_bytes = want;
    }
    return ORB_OK;
}

extern "C" orb_status lba_create(int32_t device, lba_handle** out) {
    if (!out) return set_error(ORB_ERR_INVALID, "null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
        return set_error(ORB_ERR_NO_DEVICE, "no usable CUDA device (this library has no CPU fallback)");
    ORB_CUDA(cudaSetDevice(device));
    lba_handle* h = new lba_handle();
    h->device = device;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return set_error(ORB_ERR_CUDA, "cudaStreamCreate failed");
    }
    if (cudaHostAlloc((void**)&h->h_stop, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&h->d_stop, h->h_stop, 0) != cudaSuccess) {
        cudaStreamDestroy(h->stream);
        delete h;
        return set_error(ORB_ERR_CUDA, "mapped stop flag allocation failed");
    }
    *h->h_stop = 0;
    ORB_CUDA(cudaFuncSetAttribute(k_lba, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)((LBA_THREADS + 960 + 960 + LBA_SMEM_N * LBA_SMEM_N) * sizeof(double))));
    *out = h;
    return ORB_OK;
}

extern "C" void lba_destroy(lba_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->d_buf) cudaFree(h->d_buf);
    if (h->h_buf) cudaFreeHost(h->h_buf);
    if (h->h_stop) cudaFreeHost(h->h_stop);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

namespace {
struct Carver {
    size_t off = 0;
    size_t take(size_t bytes) {
        off = (off + 255) / 256 * 256;
        const size_t o = off;
        off += bytes;
        return o;
    }
};

struct Prep {   // host-side index structures of one problem (offsets into the staging buffer)
    int nP = 0, nBlocks = 0;
    std::vector<int> pidx, mpStart, mpEdge, poseStart, poseEdge, pairStart, blockP1, blockP2;
    std::vector<int2> pairItem;
};

orb_status prepare(const lba_problem& in, Prep& R) {
    const int nKF = in.n_kf, nMP = in.n_mp, nE = in.n_edges;
    R.pidx.assign(nKF, -1);
    R.nP = 0;
    for (int k = 0; k < nKF; ++k)
        if (!in.fixed[k]) R.pidx[k] = R.nP++;
    if (R.nP > 160) return set_error(ORB_ERR_UNSUPPORTED, "more than 160 free keyframes in one local BA");
    R.mpStart.assign(nMP + 1, 0);
    R.poseStart.assign(R.nP + 1, 0);
    for (int e = 0; e < nE; ++e) {
        if (in.edge_kf[e] < 0 || in.edge_kf[e] >= nKF || in.edge_mp[e] < 0 || in.edge_mp[e] >= nMP)
            return set_error(ORB_ERR_INVALID, "edge index out of range");
        ++R.mpStart[in.edge_mp[e] + 1];
        const int p = R.pidx[in.edge_kf[e]];
        if (p >= 0) ++R.poseStart[p + 1];
    }
    for (int l = 0; l < nMP; ++l) R.mpStart[l + 1] += R.mpStart[l];
    for (int p = 0; p < R.nP; ++p) R.poseStart[p + 1] += R.poseStart[p];
    R.mpEdge.resize(nE);
    R.poseEdge.resize(R.poseStart[R.nP]);
    {
        std::vector<int> a(R.mpStart.begin(), R.mpStart.end() - 1), b(R.poseStart.begin(), R.poseStart.end() - 1);
        for (int e = 0; e < nE; ++e) {
            R.mpEdge[a[in.edge_mp[e]]++] = e;
            const int p = R.pidx[in.edge_kf[e]];
            if (p >= 0) R.poseEdge[b[p]++] = e;
        }
    }
    // upper pose-block pairs that share at least one landmark (the sparsity pattern of H_schur,
    // block_solver.hpp:199-238) + every diagonal block
    const int nP = R.nP;
    std::vector<int> cnt((size_t)nP * nP, 0);
    for (int l = 0; l < nMP; ++l)
        for (int a = R.mpStart[l]; a < R.mpStart[l + 1]; ++a) {
            const int p1 = R.pidx[in.edge_kf[R.mpEdge[a]]];
            if (p1 < 0) continue;
            for (int b = R.mpStart[l]; b < R.mpStart[l + 1]; ++b) {
                const int p2 = R.pidx[in.edge_kf[R.mpEdge[b]]];
                if (p2 < p1) continue;
                if (p2 == p1 && b != a) continue;   // a landmark has one edge per keyframe; guard duplicates
                ++cnt[(size_t)p1 * nP + p2];
            }
        }
    std::vector<int> blockOf((size_t)nP * nP, -1);
    R.blockP1.clear(); R.blockP2.clear(); R.pairStart.assign(1, 0);
    for (int p1 = 0; p1 < nP; ++p1)
        for (int p2 = p1; p2 < nP; ++p2)
            if (cnt[(size_t)p1 * nP + p2] > 0 || p1 == p2) {
                blockOf[(size_t)p1 * nP + p2] = (int)R.blockP1.size();
                R.blockP1.push_back(p1);
                R.blockP2.push_back(p2);
                R.pairStart.push_back(R.pairStart.back() + cnt[(size_t)p1 * nP + p2]);
            }
    R.nBlocks = (int)R.blockP1.size();
    R.pairItem.resize(R.pairStart.back());
    std::vector<int> fill(R.pairStart.begin(), R.pairStart.end() - 1);
    for (int l = 0; l < nMP; ++l)
        for (int a = R.mpStart[l]; a < R.mpStart[l + 1]; ++a) {
            const int e1 = R.mpEdge[a], p1 = R.pidx[in.edge_kf[e1]];
            if (p1 < 0) continue;
            for (int b = R.mpStart[l]; b < R.mpStart[l + 1]; ++b) {
                const int e2 = R.mpEdge[b], p2 = R.pidx[in.edge_kf[e2]];
                if (p2 < p1) continue;
                if (p2 == p1 && b != a) continue;
                R.pairItem[fill[blockOf[(size_t)p1 * nP + p2]]++] = make_int2(e1, e2);
            }
        }
    return ORB_OK;
}
}  // namespace

extern "C" orb_status lba_solve_batch(lba_handle* h, int32_t n_problems, const lba_problem* in, lba_result* out,
                                      const volatile int32_t* stop_flag) {
    if (!h || !in || !out || n_problems < 1) return set_error(ORB_ERR_INVALID, "bad arguments");
    ORB_CUDA(cudaSetDevice(h->device));
    std::vector<Prep> prep(n_problems);
    std::vector<LbaDev> dev(n_problems);
    // pass 1: sizes.  Every array lives in one device buffer; inputs are staged in one pinned buffer with the
    // same offsets so a single H2D moves them all.
    Carver in_c, work_c;
    struct Offs { size_t pose0, point0, pidx, ekf, emp, obs, invs2, mpStart, mpEdge, poseStart, poseEdge, pairStart, pairItem, bP1, bP2;
                  size_t pose, point, poseS, pointS, Hpl, W, Hll, bl, Dinv, db, Hpp, bp, xl, Hs, err, dpos, stats; };
    std::vector<Offs> offs(n_problems);
    for (int i = 0; i < n_problems; ++i) {
        const lba_problem& p = in[i];
        if (p.n_kf < 1 || p.n_mp < 1 || p.n_edges < 1 || !p.pose || !p.fixed || !p.point || !p.edge_kf || !p.edge_mp || !p.obs ||
            !p.inv_sigma2 || !out[i].pose || !out[i].point)
            return set_error(ORB_ERR_INVALID, "incomplete lba_problem / lba_result");
        orb_status s = prepare(p, prep[i]);
        if (s != ORB_OK) return s;
        const Prep& R = prep[i];
        Offs& o = offs[i];
        const size_t nKF = p.n_kf, nMP = p.n_mp, nE = p.n_edges, nP = R.nP, n = 6 * nP;
        o.pose0 = in_c.take(nKF * 56); o.point0 = in_c.take(nMP * 24); o.pidx = in_c.take(nKF * 4);
        o.ekf = in_c.take(nE * 4); o.emp = in_c.take(nE * 4); o.obs = in_c.take(nE * 24); o.invs2 = in_c.take(nE * 8);
        o.mpStart = in_c.take((nMP + 1) * 4); o.mpEdge = in_c.take(nE * 4);
        o.poseStart = in_c.take((nP + 1) * 4); o.poseEdge = in_c.take(R.poseEdge.size() * 4 + 4);
        o.pairStart = in_c.take((R.nBlocks + 1) * 4); o.pairItem = in_c.take(R.pairItem.size() * 8 + 8);
        o.bP1 = in_c.take(R.nBlocks * 4 + 4); o.bP2 = in_c.take(R.nBlocks * 4 + 4);
        o.pose = work_c.take(nKF * 56); o.point = work_c.take(nMP * 24); o.poseS = work_c.take(nKF * 56); o.pointS = work_c.take(nMP * 24);
        o.Hpl = work_c.take(nE * 144); o.W = work_c.take(nE * 216); o.Hll = work_c.take(nMP * 72); o.bl = work_c.take(nMP * 24);
        o.Dinv = work_c.take(nMP * 72); o.db = work_c.take(nMP * 24); o.Hpp = work_c.take(nP * 288 + 8); o.bp = work_c.take(nP * 48 + 8);
        o.xl = work_c.take(nMP * 24); o.Hs = work_c.take(n > LBA_SMEM_N ? n * n * 8 : 8); o.err = work_c.take(nE * 8);
        o.dpos = work_c.take(nE); o.stats = work_c.take(64);
    }
    const size_t in_bytes = (in_c.off + 255) / 256 * 256, work_bytes = work_c.off + 256;
    const size_t desc_off = in_bytes, desc_bytes = (sizeof(LbaDev) * n_problems + 255) / 256 * 256;
    orb_status s = lba_ensure(h, in_bytes + desc_bytes + work_bytes, in_bytes + desc_bytes);
    if (s != ORB_OK) return s;
    uint8_t* hb = h->h_buf;
    uint8_t* db = h->d_buf;
    uint8_t* dwork = db + in_bytes + desc_bytes;
    for (int i = 0; i < n_problems; ++i) {
        const lba_problem& p = in[i];
        const Prep& R = prep[i];
        const Offs& o = offs[i];
        memcpy(hb + o.pose0, p.pose, (size_t)p.n_kf * 56);
        memcpy(hb + o.point0, p.point, (size_t)p.n_mp * 24);
        memcpy(hb + o.pidx, R.pidx.data(), (size_t)p.n_kf * 4);
        memcpy(hb + o.ekf, p.edge_kf, (size_t)p.n_edges * 4);
        memcpy(hb + o.emp, p.edge_mp, (size_t)p.n_edges * 4);
        memcpy(hb + o.obs, p.obs, (size_t)p.n_edges * 24);
        memcpy(hb + o.invs2, p.inv_sigma2, (size_t)p.n_edges * 8);
        memcpy(hb + o.mpStart, R.mpStart.data(), R.mpStart.size() * 4);
        memcpy(hb + o.mpEdge, R.mpEdge.data(), R.mpEdge.size() * 4);
        memcpy(hb + o.poseStart, R.poseStart.data(), R.poseStart.size() * 4);
        if (!R.poseEdge.empty()) memcpy(hb + o.poseEdge, R.poseEdge.data(), R.poseEdge.size() * 4);
        memcpy(hb + o.pairStart, R.pairStart.data(), R.pairStart.size() * 4);
        if (!R.pairItem.empty()) memcpy(hb + o.pairItem, R.pairItem.data(), R.pairItem.size() * 8);
        if (R.nBlocks) {
            memcpy(hb + o.bP1, R.blockP1.data(), (size_t)R.nBlocks * 4);
            memcpy(hb + o.bP2, R.blockP2.data(), (size_t)R.nBlocks * 4);
        }
        LbaDev& D = dev[i];
        D.nKF = p.n_kf; D.nMP = p.n_mp; D.nE = p.n_edges; D.nP = R.nP; D.n = 6 * R.nP;
        D.pose0 = (const double*)(db + o.pose0); D.point0 = (const double*)(db + o.point0); D.pidx = (const int*)(db + o.pidx);
        D.ekf = (const int*)(db + o.ekf); D.emp = (const int*)(db + o.emp); D.obs = (const double*)(db + o.obs);
        D.invs2 = (const double*)(db + o.invs2); D.mpStart = (const int*)(db + o.mpStart); D.mpEdge = (const int*)(db + o.mpEdge);
        D.poseStart = (const int*)(db + o.poseStart); D.poseEdge = (const int*)(db + o.poseEdge);
        D.pairStart = (const int*)(db + o.pairStart); D.pairItem = (const int2*)(db + o.pairItem);
        D.blockP1 = (const int*)(db + o.bP1); D.blockP2 = (const int*)(db + o.bP2); D.nPairBlocks = R.nBlocks;
        D.fx = p.fx; D.fy = p.fy; D.cx = p.cx; D.cy = p.cy; D.bf = p.bf; D.lambdaInit = p.lambda_init; D.maxIters = p.max_iters;
        D.stop = stop_flag ? h->d_stop : nullptr;
        D.pose = (double*)(dwork + o.pose); D.point = (double*)(dwork + o.point); D.poseSave = (double*)(dwork + o.poseS);
        D.pointSave = (double*)(dwork + o.pointS); D.Hpl = (double*)(dwork + o.Hpl); D.W = (double*)(dwork + o.W);
        D.Hll = (double*)(dwork + o.Hll); D.bl = (double*)(dwork + o.bl); D.Dinv = (double*)(dwork + o.Dinv); D.db = (double*)(dwork + o.db);
        D.Hpp = (double*)(dwork + o.Hpp); D.bp = (double*)(dwork + o.bp); D.xl = (double*)(dwork + o.xl); D.HsGlobal = (double*)(dwork + o.Hs);
        D.err = (double*)(dwork + o.err); D.depthPos = dwork + o.dpos; D.stats = (double*)(dwork + o.stats);
    }
    memcpy(hb + desc_off, dev.data(), sizeof(LbaDev) * n_problems);
    *h->h_stop = (stop_flag && *stop_flag) ? 1 : 0;
    ORB_CUDA(cudaMemcpyAsync(db, hb, in_bytes + desc_bytes, cudaMemcpyHostToDevice, h->stream));
    int maxn = 0;
    for (int i = 0; i < n_problems; ++i) maxn = std::max(maxn, dev[i].n);
    const size_t smem = (LBA_THREADS + 960 + 960 + (size_t)(maxn <= LBA_SMEM_N ? maxn * maxn : 0)) * sizeof(double);
    k_lba<<<n_problems, LBA_THREADS, smem, h->stream>>>((const LbaDev*)(db + desc_off));
    ORB_LAUNCHED();
    ORB_CUDA(cudaGetLastError());
    // relay *pbStopFlag to the mapped flag while the kernel runs (LocalMapping::InterruptBA writes it asynchronously)
    if (stop_flag) {
        while (cudaStreamQuery(h->stream) == cudaErrorNotReady) {
            if (*stop_flag) *h->h_stop = 1;
            std::this_thread::sleep_for(std::chrono::microseconds(20));
        }
    }
    ORB_CUDA(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < n_problems; ++i) {
        const lba_problem& p = in[i];
        const Offs& o = offs[i];
        ORB_CUDA(cudaMemcpyAsync(out[i].pose, dwork + o.pose, (size_t)p.n_kf * 56, cudaMemcpyDeviceToHost, h->stream));
        ORB_CUDA(cudaMemcpyAsync(out[i].point, dwork + o.point, (size_t)p.n_mp * 24, cudaMemcpyDeviceToHost, h->stream));
        if (out[i].edge_chi2) ORB_CUDA(cudaMemcpyAsync(out[i].edge_chi2, dwork + o.err, (size_t)p.n_edges * 8, cudaMemcpyDeviceToHost, h->stream));
        if (out[i].edge_depth_positive)
            ORB_CUDA(cudaMemcpyAsync(out[i].edge_depth_positive, dwork + o.dpos, (size_t)p.n_edges, cudaMemcpyDeviceToHost, h->stream));
        ORB_CUDA(cudaMemcpyAsync(hb + 64 * i, dwork + o.stats, 40, cudaMemcpyDeviceToHost, h->stream));
    }
    ORB_CUDA(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < n_problems; ++i) {
        const double* st = (const double*)(hb + 64 * i);
        out[i].iterations = (int32_t)st[0];
        out[i].lambda = st[1];
        out[i].chi2 = st[2];
        out[i].trials = (int32_t)st[3];
        out[i].chi2_initial = st[4];
    }
    return ORB_OK;
}

extern "C" orb_status lba_solve(lba_handle* h, const lba_problem* in, lba_result* out, const volatile int32_t* stop_flag) {
    return lba_solve_batch(h, 1, in, out, stop_flag);
}

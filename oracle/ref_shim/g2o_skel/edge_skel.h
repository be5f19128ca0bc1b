// oracle/ref_shim/g2o_skel/edge_skel.h -- TEST INFRASTRUCTURE.
//
// Skeleton around the reprojection edges' own arithmetic, cut out of the reference at build time (oracle/Makefile target ref7) and
// compiled VERBATIM:
//     Thirdparty/g2o/g2o/types/types_six_dof_expmap.cpp : EdgeStereoSE3ProjectXYZ::cam_project / linearizeOplus,
//                                                         EdgeStereoSE3ProjectXYZOnlyPose::cam_project / linearizeOplus
//     src/OptimizableTypes.cpp                          : EdgeSE3ProjectXYZ::linearizeOplus, EdgeSE3ProjectXYZOnlyPose::linearizeOplus
//     src/CameraModels/Pinhole.cpp                      : projectJac(const Eigen::Vector3d&), project(const Eigen::Vector3d&)
// Those bodies are element-wise scalar expressions over (R, the transformed point, the intrinsics): which Jacobian entry gets which
// expression, in which association, with which float / double types.  What is NOT the reference's here: SE3Quat::map and
// Quaternion::toRotationMatrix (Eigen's formulas, restated like the oracle does) and the 2x3 * 3xN products of the monocular edges
// (terms summed in index order).  tests/test_oracle_vs_ref_g2o.py compares the oracle's edge_jacobians / edge residuals with these.
#pragma once
#include <cmath>
#include <vector>

namespace Eigen {
template <typename T, int R, int C>
struct Matrix {
    T v[R * C];
    Matrix() { for (int i = 0; i < R * C; ++i) v[i] = T(0); }
    T& operator()(int r, int c) { return v[r * C + c]; }
    const T& operator()(int r, int c) const { return v[r * C + c]; }
    T& operator[](int i) { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
    T& operator()(int i) { return v[i]; }
    const T& operator()(int i) const { return v[i]; }
    struct CommaInit {
        Matrix* m; int i;
        template <typename U> CommaInit& operator,(U x) { m->v[i++] = (T)x; return *this; }
    };
    template <typename U> CommaInit operator<<(U x) { v[0] = (T)x; return CommaInit{this, 1}; }
    Matrix operator-() const { Matrix o; for (int i = 0; i < R * C; ++i) o.v[i] = -v[i]; return o; }
    template <int K>
    Matrix<T, R, K> operator*(const Matrix<T, C, K>& b) const {   // terms in index order
        Matrix<T, R, K> o;
        for (int r = 0; r < R; ++r)
            for (int k = 0; k < K; ++k) {
                T s = (*this)(r, 0) * b(0, k);
                for (int c = 1; c < C; ++c) s += (*this)(r, c) * b(c, k);
                o(r, k) = s;
            }
        return o;
    }
};
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 3> Matrix3d;
struct Quaterniond {
    double x, y, z, w;
    Matrix3d toRotationMatrix() const {   // Eigen/src/Geometry/Quaternion.h, as oracle quat_to_R
        Matrix3d R;
        R(0, 0) = 1 - 2 * (y * y + z * z); R(0, 1) = 2 * (x * y - z * w);     R(0, 2) = 2 * (x * z + y * w);
        R(1, 0) = 2 * (x * y + z * w);     R(1, 1) = 1 - 2 * (x * x + z * z); R(1, 2) = 2 * (y * z - x * w);
        R(2, 0) = 2 * (x * z - y * w);     R(2, 1) = 2 * (y * z + x * w);     R(2, 2) = 1 - 2 * (x * x + y * y);
        return R;
    }
};
}  // namespace Eigen

namespace g2o {
using Eigen::Vector2d;
using Eigen::Vector3d;
using Eigen::Matrix3d;

struct SE3Quat {
    Eigen::Quaterniond _r;
    Vector3d _t;
    const Eigen::Quaterniond& rotation() const { return _r; }
    Vector3d map(const Vector3d& X) const {   // _r * xyz + _t: Eigen's _transformVector, as oracle se3_map
        const double uvx = _r.y * X[2] - _r.z * X[1], uvy = _r.z * X[0] - _r.x * X[2], uvz = _r.x * X[1] - _r.y * X[0];
        const double ux = uvx + uvx, uy = uvy + uvy, uz = uvz + uvz;
        const double cx = _r.y * uz - _r.z * uy, cy = _r.z * ux - _r.x * uz, cz = _r.x * uy - _r.y * ux;
        Vector3d o;
        o[0] = X[0] + _r.w * ux + cx + _t[0]; o[1] = X[1] + _r.w * uy + cy + _t[1]; o[2] = X[2] + _r.w * uz + cz + _t[2];
        return o;
    }
};
struct OptimizableGraphVertex { virtual ~OptimizableGraphVertex() {} };
struct VertexSE3Expmap : OptimizableGraphVertex { SE3Quat _e; const SE3Quat& estimate() const { return _e; } };
struct VertexSBAPointXYZ : OptimizableGraphVertex { Vector3d _e; const Vector3d& estimate() const { return _e; } };

inline Vector2d project2d(const Vector3d& v) { Vector2d r; r[0] = v[0] / v[2]; r[1] = v[1] / v[2]; return r; }   // types_six_dof_expmap.cpp:36-41

class EdgeStereoSE3ProjectXYZ {
   public:
    Vector3d cam_project(const Vector3d& trans_xyz, const float& bf) const;
    void linearizeOplus();
    double fx, fy, cx, cy, bf;
    std::vector<OptimizableGraphVertex*> _vertices;
    Eigen::Matrix<double, 3, 3> _jacobianOplusXi;
    Eigen::Matrix<double, 3, 6> _jacobianOplusXj;
};
class EdgeStereoSE3ProjectXYZOnlyPose {
   public:
    Vector3d cam_project(const Vector3d& trans_xyz) const;
    void linearizeOplus();
    double fx, fy, cx, cy, bf;
    Vector3d Xw;
    std::vector<OptimizableGraphVertex*> _vertices;
    Eigen::Matrix<double, 3, 6> _jacobianOplusXi;
};
}  // namespace g2o

namespace ORB_SLAM3 {
class Pinhole {
   public:
    Eigen::Vector2d project(const Eigen::Vector3d& v3D);
    Eigen::Matrix<double, 2, 3> projectJac(const Eigen::Vector3d& v3D);
    std::vector<float> mvParameters;
};
typedef Pinhole GeometricCamera;
class EdgeSE3ProjectXYZ {
   public:
    void linearizeOplus();
    GeometricCamera* pCamera;
    std::vector<g2o::OptimizableGraphVertex*> _vertices;
    Eigen::Matrix<double, 2, 3> _jacobianOplusXi;
    Eigen::Matrix<double, 2, 6> _jacobianOplusXj;
};
class EdgeSE3ProjectXYZOnlyPose {
   public:
    void linearizeOplus();
    GeometricCamera* pCamera;
    Eigen::Vector3d Xw;
    std::vector<g2o::OptimizableGraphVertex*> _vertices;
    Eigen::Matrix<double, 2, 6> _jacobianOplusXi;
};
}  // namespace ORB_SLAM3

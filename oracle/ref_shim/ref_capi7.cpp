// oracle/ref_shim/ref_capi7.cpp -- TEST INFRASTRUCTURE: the reprojection edges' own Jacobians and projections (see g2o_skel/edge_skel.h),
// cut out of the reference at build time into oracle/_ref/gen/ref7_*.inc and compiled verbatim.
#include "g2o_skel/edge_skel.h"

namespace g2o {
#include "../_ref/gen/ref7_g2o.inc"
}
namespace ORB_SLAM3 {
#include "../_ref/gen/ref7_orb.inc"
}

namespace {
void fill(g2o::VertexSE3Expmap& v, const double* p7) {
    v._e._r.x = p7[0]; v._e._r.y = p7[1]; v._e._r.z = p7[2]; v._e._r.w = p7[3];
    v._e._t[0] = p7[4]; v._e._t[1] = p7[5]; v._e._t[2] = p7[6];
}
}  // namespace

extern "C" {
// binary edges of LocalBundleAdjustment: A = _jacobianOplusXi (D x 3, point), B = _jacobianOplusXj (D x 6, pose), r = obs - projection
void ref7_edge(const double* pose7, const double* X, const double* obs3, const float* cam5, int stereo, double* A9, double* B18, double* r3) {
    g2o::VertexSE3Expmap vp;
    g2o::VertexSBAPointXYZ vx;
    fill(vp, pose7);
    for (int i = 0; i < 3; ++i) vx._e[i] = X[i];
    if (stereo) {
        g2o::EdgeStereoSE3ProjectXYZ e;
        e.fx = cam5[0]; e.fy = cam5[1]; e.cx = cam5[2]; e.cy = cam5[3]; e.bf = cam5[4];   // Optimizer.cc:2038-2042: float members into doubles
        e._vertices.push_back(&vx); e._vertices.push_back(&vp);
        e.linearizeOplus();
        for (int i = 0; i < 9; ++i) A9[i] = e._jacobianOplusXi.v[i];
        for (int i = 0; i < 18; ++i) B18[i] = e._jacobianOplusXj.v[i];
        const float bf = cam5[4];
        const g2o::Vector3d p = e.cam_project(vp.estimate().map(vx.estimate()), bf);   // computeError, types_six_dof_expmap.h:160-166
        for (int i = 0; i < 3; ++i) r3[i] = obs3[i] - p[i];
    } else {
        ORB_SLAM3::Pinhole cam;
        cam.mvParameters = {cam5[0], cam5[1], cam5[2], cam5[3]};
        ORB_SLAM3::EdgeSE3ProjectXYZ e;
        e.pCamera = &cam;
        e._vertices.push_back(&vx); e._vertices.push_back(&vp);
        e.linearizeOplus();
        for (int i = 0; i < 6; ++i) A9[i] = e._jacobianOplusXi.v[i];
        for (int i = 0; i < 12; ++i) B18[i] = e._jacobianOplusXj.v[i];
        const Eigen::Vector2d p = cam.project(vp.estimate().map(vx.estimate()));        // computeError, OptimizableTypes.h:88-93
        r3[0] = obs3[0] - p[0]; r3[1] = obs3[1] - p[1]; r3[2] = 0;
    }
}
// unary edges of PoseOptimization: B = _jacobianOplusXi (D x 6)
void ref7_pose_edge(const double* pose7, const double* Xw, const double* obs3, const float* cam5, int stereo, double* B18, double* r3) {
    g2o::VertexSE3Expmap vp;
    fill(vp, pose7);
    if (stereo) {
        g2o::EdgeStereoSE3ProjectXYZOnlyPose e;
        e.fx = cam5[0]; e.fy = cam5[1]; e.cx = cam5[2]; e.cy = cam5[3]; e.bf = cam5[4];
        for (int i = 0; i < 3; ++i) e.Xw[i] = Xw[i];
        e._vertices.push_back(&vp);
        e.linearizeOplus();
        for (int i = 0; i < 18; ++i) B18[i] = e._jacobianOplusXi.v[i];
        const g2o::Vector3d p = e.cam_project(vp.estimate().map(e.Xw));
        for (int i = 0; i < 3; ++i) r3[i] = obs3[i] - p[i];
    } else {
        ORB_SLAM3::Pinhole cam;
        cam.mvParameters = {cam5[0], cam5[1], cam5[2], cam5[3]};
        ORB_SLAM3::EdgeSE3ProjectXYZOnlyPose e;
        e.pCamera = &cam;
        for (int i = 0; i < 3; ++i) e.Xw[i] = Xw[i];
        e._vertices.push_back(&vp);
        e.linearizeOplus();
        for (int i = 0; i < 12; ++i) B18[i] = e._jacobianOplusXi.v[i];
        const Eigen::Vector2d p = cam.project(vp.estimate().map(e.Xw));
        r3[0] = obs3[0] - p[0]; r3[1] = obs3[1] - p[1]; r3[2] = 0;
    }
}
}  // extern "C"

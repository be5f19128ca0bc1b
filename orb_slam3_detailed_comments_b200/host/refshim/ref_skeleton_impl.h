// refshim/ref_skeleton_impl.h -- bodies of the skeleton classes: a minimal in-memory map for the marshaling tests (tests/host/).
// Include in exactly one translation unit of a test binary.
#pragma once
#include "ref_skeleton.h"

namespace ORB_SLAM3 {

float Frame::fx = 0, Frame::fy = 0, Frame::cx = 0, Frame::cy = 0, Frame::mnMinX = 0, Frame::mnMaxX = 0, Frame::mnMinY = 0, Frame::mnMaxY = 0;

long unsigned int Map::GetInitKFid() { return mock_init_kf_id; }
void Map::IncreaseChangeIndex() { ++mock_change_index; }
bool Map::IsInertial() { return mock_inertial; }

void MapPoint::SetWorldPos(const Eigen::Vector3f& Pos) { mock_pos = Pos; }
Eigen::Vector3f MapPoint::GetWorldPos() { return mock_pos; }
Eigen::Vector3f MapPoint::GetNormal() { return mock_normal; }
std::map<KeyFrame*, std::tuple<int, int>> MapPoint::GetObservations() { return mock_obs; }
int MapPoint::Observations() { return (int)mock_obs.size(); }
void MapPoint::EraseObservation(KeyFrame* pKF, bool) { mock_obs.erase(pKF); }
bool MapPoint::isBad() { return mock_bad; }
cv::Mat MapPoint::GetDescriptor() { return mock_desc; }
void MapPoint::UpdateNormalAndDepth() { ++mock_normal_updates; }
Map* MapPoint::GetMap() { return mock_map; }

void KeyFrame::SetPose(const Sophus::SE3f& Tcw) { mock_Tcw = Tcw; ++mock_pose_sets; }
Sophus::SE3f KeyFrame::GetPose() { return mock_Tcw; }
std::vector<KeyFrame*> KeyFrame::GetVectorCovisibleKeyFrames() { return mock_covisible; }
void KeyFrame::EraseMapPointMatch(MapPoint* pMP) {
    for (auto& m : mock_matches)
        if (m == pMP) m = nullptr;
}
std::vector<MapPoint*> KeyFrame::GetMapPointMatches() { return mock_matches; }
bool KeyFrame::isBad() { return mock_bad; }
#ifdef ORB_REFSHIM_REF_LBA
Sophus::SE3f KeyFrame::GetRelativePoseTrl() { return Sophus::SE3f(); }   // two-camera rigs: not exercised
#endif
Map* KeyFrame::GetMap() { return mock_map; }
#ifdef ORB_REFSHIM_FUSE
std::vector<int> g_fuse_log;   // (op, acting point, other point or -1, keyframe feature or -1) per mutation, in call order
void MapPoint::AddObservation(KeyFrame* pKF, int idx) { mock_obs[pKF] = std::make_tuple(idx, -1); g_fuse_log.insert(g_fuse_log.end(), {3, mock_id, -1, idx}); }
bool MapPoint::IsInKeyFrame(KeyFrame* pKF) { return mock_obs.count(pKF) != 0; }
void MapPoint::Replace(MapPoint* pMP) { mock_bad = true; g_fuse_log.insert(g_fuse_log.end(), {1, mock_id, pMP->mock_id, -1}); }   // MapPoint.cc:323-384 marks this point bad
void KeyFrame::AddMapPoint(MapPoint* pMP, const size_t& idx) { mock_matches[idx] = pMP; }
MapPoint* KeyFrame::GetMapPoint(const size_t& idx) { return mock_matches[idx]; }
std::set<MapPoint*> KeyFrame::GetMapPoints() { std::set<MapPoint*> s; for (MapPoint* p : mock_matches) if (p && !p->isBad()) s.insert(p); return s; }   // KeyFrame.cc:370-385: the non-NULL, non-bad entries of mvpMapPoints
#endif
#ifdef ORB_REFSHIM_TRI
Sophus::SE3f KeyFrame::GetPoseInverse() { return mock_Tcw.inverse(); }
Eigen::Vector3f KeyFrame::GetCameraCenter() { return mock_Tcw.inverse().translation(); }
#endif
#ifdef ORB_REFSHIM_POSE
std::mutex MapPoint::mGlobalMutex;
void Frame::SetPose(const Sophus::SE3<float>& Tcw) { mTcw = Tcw; ++mock_pose_sets; }
#endif
#ifdef ORB_REFSHIM_LIBA
long unsigned Map::KeyFramesInMap() { return mock_n_keyframes; }
void IMU::Preintegrated::SetNewBias(const Bias& bu_) { mock_bu = bu_; ++mock_bias_sets; }
void KeyFrame::SetVelocity(const Eigen::Vector3f& Vw_) { mock_vel = Vw_; ++mock_vel_sets; }
Eigen::Matrix3f KeyFrame::GetRotation() { return mock_Tcw.rotationMatrix(); }
Eigen::Vector3f KeyFrame::GetTranslation() { return mock_Tcw.translation(); }
// KeyFrame::GetImuRotation / GetImuPosition (KeyFrame.cc:157-167): the rotation and translation of Twb = mTwc * mImuCalib.mTcb
Eigen::Matrix3f KeyFrame::GetImuRotation() { return (mock_Tcw.inverse() * mImuCalib.mTcb).rotationMatrix(); }
Eigen::Vector3f KeyFrame::GetImuPosition() { return (mock_Tcw.inverse() * mImuCalib.mTcb).translation(); }
Eigen::Vector3f KeyFrame::GetVelocity() { return mock_vel; }
void KeyFrame::SetNewBias(const IMU::Bias& b) { mock_bias = b; ++mock_bias_sets; }
Eigen::Vector3f KeyFrame::GetGyroBias() { return Eigen::Vector3f(mock_bias.bwx, mock_bias.bwy, mock_bias.bwz); }
Eigen::Vector3f KeyFrame::GetAccBias() { return Eigen::Vector3f(mock_bias.bax, mock_bias.bay, mock_bias.baz); }
IMU::Bias KeyFrame::GetImuBias() { return mock_bias; }
#endif

}  // namespace ORB_SLAM3

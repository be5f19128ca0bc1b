// tests/host/pose_cpu.cc -- TEST INFRASTRUCTURE (CPU tier): Optimizer::PoseOptimization (host/Optimizer_pose_b200.cc, orbo_pose_optimization
// answered by the oracle: pose_stub.cc) over a mock Frame read from raw arrays; tests/test_host_pose_vs_ref.py compares the frame it leaves
// behind with the one the reference's own function leaves (oracle/_ref part 5).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "ref_skeleton_impl.h"
#include "Optimizer.h"

using namespace ORB_SLAM3;

static std::string g_dir;
template <typename T>
static std::vector<T> rd(const std::string& name) {
    std::ifstream f(g_dir + "/" + name, std::ios::binary);
    if (!f) { std::fprintf(stderr, "missing input %s\n", name.c_str()); std::exit(2); }
    f.seekg(0, std::ios::end);
    const size_t n = (size_t)f.tellg();
    f.seekg(0);
    std::vector<T> v(n / sizeof(T));
    f.read((char*)v.data(), (std::streamsize)(v.size() * sizeof(T)));
    return v;
}
template <typename T>
static void wr(const std::string& name, const std::vector<T>& v) {
    std::ofstream f(g_dir + "/" + name, std::ios::binary);
    f.write((const char*)v.data(), (std::streamsize)(v.size() * sizeof(T)));
}

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: pose_cpu <dir>\n"); return 2; }
    g_dir = argv[1];
    auto cam = rd<float>("cam5.f32"), pose = rd<float>("pose.f32"), xw = rd<float>("world_pos.f32"), xy = rd<float>("kp_xy.f32"), ur = rd<float>("uright.f32"),
         isg = rd<float>("inv_level_sigma2.f32");
    auto has = rd<uint8_t>("has_mp.u8");
    auto oct = rd<int>("octave.i32");
    const int N = (int)has.size();
    Frame F;
    Frame::fx = cam[0]; Frame::fy = cam[1]; Frame::cx = cam[2]; Frame::cy = cam[3];
    F.mbf = cam[4];
    F.N = N;
    int dummy_extractor = 0;
    F.mpORBextractorLeft = reinterpret_cast<ORBextractor*>(&dummy_extractor);
    F.mTcw = Sophus::SE3f(Eigen::Quaternionf(pose[3], pose[0], pose[1], pose[2]), Eigen::Vector3f(pose[4], pose[5], pose[6]));
    F.mvInvLevelSigma2 = isg;
    std::vector<MapPoint> mps(N);
    F.mvKeysUn.resize(N); F.mvuRight = ur; F.mvpMapPoints.assign(N, nullptr); F.mvbOutlier.assign(N, true);   // true: the function must clear the flags it owns
    for (int i = 0; i < N; ++i) {
        F.mvKeysUn[i].pt.x = xy[2 * i]; F.mvKeysUn[i].pt.y = xy[2 * i + 1]; F.mvKeysUn[i].octave = oct[i];
        if (has[i]) { mps[i].mock_pos = Eigen::Vector3f(xw[3 * i], xw[3 * i + 1], xw[3 * i + 2]); F.mvpMapPoints[i] = &mps[i]; }
    }
    const int ret = Optimizer::PoseOptimization(&F);
    const Sophus::SE3f T = F.GetPose();
    std::vector<float> pout = {T.unit_quaternion().x(), T.unit_quaternion().y(), T.unit_quaternion().z(), T.unit_quaternion().w(),
                               T.translation()(0), T.translation()(1), T.translation()(2)};
    std::vector<uint8_t> outl(N);
    for (int i = 0; i < N; ++i) outl[i] = F.mvbOutlier[i] ? 1 : 0;
    wr("out_pose.f32", pout); wr("out_outlier.u8", outl); wr("out_ret.i32", std::vector<int>{ret, F.mock_pose_sets});
    std::printf("pose_cpu ok\n");
    return 0;
}

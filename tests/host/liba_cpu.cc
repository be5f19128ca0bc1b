// tests/host/liba_cpu.cc -- TEST INFRASTRUCTURE (CPU tier): Optimizer::LocalInertialBA over the mock map of host/refshim, without a GPU.
// Linked twice by build_liba_cpu.sh:
//   liba_cpu_mine = this + host/Optimizer_liba_b200.cc + liba_stub.cc (liba_solve -> the oracle's orc_liba)
//   liba_cpu_ref  = this + liba_ref.cc (the reference's own function, verbatim, over graph stand-ins that flatten into the same orc_liba)
// Same inputs (raw arrays in a directory), same outputs; tests/test_host_liba_vs_ref.py compares them.
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
static std::map<std::string, double> read_meta() {
    std::map<std::string, double> m;
    std::ifstream f(g_dir + "/meta.txt");
    std::string k;
    double v;
    while (f >> k >> v) m[k] = v;
    return m;
}
static Sophus::SE3f se3_of(const float* q7) {   // qx qy qz qw tx ty tz
    return Sophus::SE3f(Eigen::Quaternionf(q7[3], q7[0], q7[1], q7[2]), Eigen::Vector3f(q7[4], q7[5], q7[6]));
}
static void fill3x3(Eigen::Matrix3f& M, const float* p) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M(r, c) = p[3 * r + c]; }

extern "C" void* liba_mock_camera();   // liba_stub.cc

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: liba_cpu <dir>\n"); return 2; }
    g_dir = argv[1];
    auto M = read_meta();
    const float cam[5] = {(float)M["fx"], (float)M["fy"], (float)M["cx"], (float)M["cy"], (float)M["bf"]};
    const int nKF = (int)M["nkf"], nMP = (int)M["nmp"];
    auto pose = rd<float>("kf_pose.f32");                 // [nKF][7] Tcw
    auto vel = rd<float>("kf_vel.f32");                   // [nKF][3]
    auto bias = rd<float>("kf_bias.f32");                 // [nKF][6] bax bay baz bwx bwy bwz
    auto flags = rd<int>("kf_flags.i32");                 // [nKF][4] bImu, prev index (-1), bad, has preintegration
    auto pre = rd<float>("kf_pre.f32");                   // [nKF][292] dT, C (225), b (6), dR (9), dV (3), dP (3), JRg JVg JVa JPg JPa (45)
    auto tcb = rd<float>("tcb.f32");                      // [7]
    auto pts = rd<float>("point.f32");
    auto depth = rd<float>("point_track_depth.f32");
    auto badmp = rd<uint8_t>("point_bad.u8");
    auto ekf = rd<int>("edge_kf.i32"), emp = rd<int>("edge_mp.i32");
    auto eobs = rd<float>("edge_obs.f32"), ew = rd<float>("edge_inv_sigma2.f32");
    const int nE = (int)ekf.size();
    Map map;
    map.mock_n_keyframes = (long unsigned)M["keyframes_in_map"];
    std::vector<std::vector<cv::KeyPoint>> keys(nKF);
    std::vector<std::vector<float>> ur(nKF), sig(nKF);
    std::vector<int> featOf(nE);
    for (int e = 0; e < nE; ++e) {
        const int k = ekf[e];
        cv::KeyPoint kp;
        kp.pt.x = eobs[3 * e]; kp.pt.y = eobs[3 * e + 1]; kp.octave = (int)keys[k].size();
        featOf[e] = (int)keys[k].size();
        keys[k].push_back(kp); ur[k].push_back(eobs[3 * e + 2]); sig[k].push_back(ew[e]);
    }
    std::vector<KeyFrame*> kfs(nKF);
    std::vector<IMU::Preintegrated> pres(nKF);
    // one contiguous block: pointer order == index order, so every std::map<KeyFrame*, ...> iterates in keyframe-index order
    KeyFrame* block = (KeyFrame*)::operator new(sizeof(KeyFrame) * nKF);
    const Sophus::SE3f Tcb = se3_of(tcb.data());
    for (int k = 0; k < nKF; ++k) {
        kfs[k] = new (block + k) KeyFrame(100 + k, cam[0], cam[1], cam[2], cam[3], cam[4], (float)M["b"], keys[k], ur[k], sig[k]);
        KeyFrame* K = kfs[k];
        K->mock_Tcw = se3_of(&pose[7 * k]);
        K->mock_map = &map;
        K->mock_bad = flags[4 * k + 2] != 0;
        K->mock_matches.assign(keys[k].size(), nullptr);
        K->mpCamera = reinterpret_cast<GeometricCamera*>(liba_mock_camera());
        K->bImu = flags[4 * k] != 0;
        K->mImuCalib.mTcb = Tcb;
        K->mImuCalib.mTbc = Tcb.inverse();
        K->mock_vel = Eigen::Vector3f(vel[3 * k], vel[3 * k + 1], vel[3 * k + 2]);
        const float* b = &bias[6 * k];
        K->mock_bias = IMU::Bias(b[0], b[1], b[2], b[3], b[4], b[5]);
        if (flags[4 * k + 3]) {
            const float* p = &pre[292 * (size_t)k];
            IMU::Preintegrated& P = pres[k];
            P.dT = p[0];
            for (int r = 0; r < 15; ++r) for (int c = 0; c < 15; ++c) P.C(r, c) = p[1 + 15 * r + c];
            P.b = IMU::Bias(p[226], p[227], p[228], p[229], p[230], p[231]);
            fill3x3(P.dR, p + 232);
            P.dV = Eigen::Vector3f(p[241], p[242], p[243]);
            P.dP = Eigen::Vector3f(p[244], p[245], p[246]);
            fill3x3(P.JRg, p + 247); fill3x3(P.JVg, p + 256); fill3x3(P.JVa, p + 265); fill3x3(P.JPg, p + 274); fill3x3(P.JPa, p + 283);
            K->mpImuPreintegrated = &P;
        }
    }
    for (int k = 0; k < nKF; ++k)
        if (flags[4 * k + 1] >= 0) kfs[k]->mPrevKF = kfs[flags[4 * k + 1]];
    std::vector<MapPoint> mps(nMP);
    for (int p = 0; p < nMP; ++p) {
        mps[p].mnId = 1000 + p; mps[p].mock_pos = Eigen::Vector3f(pts[3 * p], pts[3 * p + 1], pts[3 * p + 2]);
        mps[p].mock_map = &map; mps[p].mock_bad = badmp[p] != 0; mps[p].mTrackDepth = depth[p];
    }
    for (int e = 0; e < nE; ++e) {
        mps[emp[e]].mock_obs[kfs[ekf[e]]] = std::make_tuple(featOf[e], -1);
        kfs[ekf[e]]->mock_matches[featOf[e]] = &mps[emp[e]];
    }
    KeyFrame* pKF = kfs[(int)M["ba_kf"]];
    bool stop = false;
    int nFixed = -7, nOpt = -7, nMPs = -7, nEdges = -7;
    Optimizer::LocalInertialBA(pKF, &stop, &map, nFixed, nOpt, nMPs, nEdges, M["large"] != 0, M["rec_init"] != 0);
    std::vector<float> pout((size_t)nKF * 7), vout((size_t)nKF * 3), bout((size_t)nKF * 6), xout((size_t)nMP * 3), bu((size_t)nKF * 6, 0.f);
    std::vector<int> counters = {nFixed, nOpt, nMPs, nEdges, map.mock_change_index};
    for (int k = 0; k < nKF; ++k) {
        const Sophus::SE3f T = kfs[k]->GetPose();
        float* o = &pout[(size_t)k * 7];
        o[0] = T.unit_quaternion().x(); o[1] = T.unit_quaternion().y(); o[2] = T.unit_quaternion().z(); o[3] = T.unit_quaternion().w();
        o[4] = T.translation()(0); o[5] = T.translation()(1); o[6] = T.translation()(2);
        for (int c = 0; c < 3; ++c) vout[3 * k + c] = kfs[k]->mock_vel(c);
        const IMU::Bias b = kfs[k]->mock_bias;
        const float bb[6] = {b.bax, b.bay, b.baz, b.bwx, b.bwy, b.bwz};
        for (int c = 0; c < 6; ++c) bout[6 * k + c] = bb[c];
        const IMU::Bias u = pres[k].mock_bu;
        const float uu[6] = {u.bax, u.bay, u.baz, u.bwx, u.bwy, u.bwz};
        for (int c = 0; c < 6; ++c) bu[6 * k + c] = uu[c];
        counters.push_back(kfs[k]->mock_pose_sets); counters.push_back(kfs[k]->mock_vel_sets); counters.push_back(kfs[k]->mock_bias_sets);
        counters.push_back(pres[k].mock_bias_sets); counters.push_back((int)kfs[k]->mnBALocalForKF); counters.push_back((int)kfs[k]->mnBAFixedForKF);
    }
    std::vector<int> alive(nE), updates(nMP);
    for (int p = 0; p < nMP; ++p) {
        const Eigen::Vector3f X = mps[p].GetWorldPos();
        xout[3 * p] = X(0); xout[3 * p + 1] = X(1); xout[3 * p + 2] = X(2);
        updates[p] = mps[p].mock_normal_updates;
    }
    for (int e = 0; e < nE; ++e)   // an edge survives iff both sides still know each other
        alive[e] = (kfs[ekf[e]]->mock_matches[featOf[e]] == &mps[emp[e]] ? 1 : 0) + (mps[emp[e]].mock_obs.count(kfs[ekf[e]]) ? 2 : 0);
    wr("out_pose.f32", pout); wr("out_vel.f32", vout); wr("out_bias.f32", bout); wr("out_pre_bu.f32", bu); wr("out_point.f32", xout);
    wr("out_counters.i32", counters); wr("out_alive.i32", alive); wr("out_updates.i32", updates);
    std::printf("liba_cpu ok\n");
    return 0;
}

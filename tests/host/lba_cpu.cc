// tests/host/lba_cpu.cc -- TEST INFRASTRUCTURE (CPU tier): Optimizer::LocalBundleAdjustment over the mock map of host/refshim, as
// section 5 of host_boundary.cc sets it up, without a GPU.  Linked twice by build_lba_cpu.sh:
//   lba_cpu_mine = this + host/Optimizer_lba_b200.cc + lba_stub.cc (lba_solve_bool -> the oracle's orc_lba)
//   lba_cpu_ref  = this + lba_ref.cc (the reference's own function, verbatim, over graph stand-ins and the oracle's LbaEngine)
// Same inputs (raw arrays in a directory), same outputs; tests/test_host_lba_vs_ref.py compares them.
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

static double g_cam5[5];
extern "C" const double* lba_ref_cam5() { return g_cam5; }

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: lba_cpu <dir>\n"); return 2; }
    g_dir = argv[1];
    auto M = read_meta();
    const float cam[5] = {(float)M["fx"], (float)M["fy"], (float)M["cx"], (float)M["cy"], (float)M["bf"]};
    for (int i = 0; i < 5; ++i) g_cam5[i] = cam[i];
    const int nKF = (int)M["lba_nkf"], nMP = (int)M["lba_nmp"];
    auto pose = rd<float>("lba_pose.f32");                // [nKF][7]
    auto role = rd<int>("lba_role.i32");                  // 0: the BA keyframe, 1: covisible neighbour, 2: other observer, 3: bad neighbour
    auto pts = rd<float>("lba_point.f32");
    auto ekf = rd<int>("lba_edge_kf.i32"), emp = rd<int>("lba_edge_mp.i32");
    auto eobs = rd<float>("lba_obs.f32"), ew = rd<float>("lba_inv_sigma2.f32");
    auto badmp = rd<uint8_t>("lba_badmp.u8");
    const int nE = (int)ekf.size();
    Map map;
    map.mock_init_kf_id = (long unsigned int)M["lba_init_kf_id"];
    map.mock_inertial = M["lba_inertial"] != 0;
    // per-keyframe feature arrays: feature j of keyframe k is its j-th edge; octave j indexes a per-feature sigma table
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
    // one contiguous block: pointer order == index order, so every std::map<KeyFrame*, ...> iterates in keyframe-index order
    KeyFrame* block = (KeyFrame*)::operator new(sizeof(KeyFrame) * nKF);
    for (int k = 0; k < nKF; ++k) {
        kfs[k] = new (block + k) KeyFrame(100 + k, cam[0], cam[1], cam[2], cam[3], cam[4], (float)M["b"], keys[k], ur[k], sig[k]);
        kfs[k]->mock_Tcw = se3_of(&pose[7 * k]);
        kfs[k]->mock_map = &map;
        kfs[k]->mock_bad = role[k] == 3;
        kfs[k]->mock_matches.assign(keys[k].size(), nullptr);
    }
    std::vector<MapPoint> mps(nMP);
    for (int p = 0; p < nMP; ++p) {
        mps[p].mnId = 1000 + p; mps[p].mock_pos = Eigen::Vector3f(pts[3 * p], pts[3 * p + 1], pts[3 * p + 2]);
        mps[p].mock_map = &map; mps[p].mock_bad = badmp[p] != 0;
    }
    for (int e = 0; e < nE; ++e) {
        mps[emp[e]].mock_obs[kfs[ekf[e]]] = std::make_tuple(featOf[e], -1);
        kfs[ekf[e]]->mock_matches[featOf[e]] = &mps[emp[e]];
    }
    KeyFrame* pKF = nullptr;
    for (int k = 0; k < nKF; ++k) if (role[k] == 0) pKF = kfs[k];
    for (int k = 0; k < nKF; ++k) if (role[k] == 1 || role[k] == 3) pKF->mock_covisible.push_back(kfs[k]);
    bool stop = M["lba_stop"] != 0;
    int nFixed = -1, nOpt = -1, nMPs = -1, nEdges = -1;
    Optimizer::LocalBundleAdjustment(pKF, M.count("lba_null_stop") && M["lba_null_stop"] != 0 ? nullptr : &stop, &map, nFixed, nOpt, nMPs, nEdges);
    std::vector<float> pout((size_t)nKF * 7), xout((size_t)nMP * 3);
    std::vector<int> counters = {nFixed, nOpt, nEdges, map.mock_change_index, (int)map.msOptKFs.size(), (int)map.msFixedKFs.size()};
    for (int k = 0; k < nKF; ++k) {
        const Sophus::SE3f T = kfs[k]->GetPose();
        float* o = &pout[(size_t)k * 7];
        o[0] = T.unit_quaternion().x(); o[1] = T.unit_quaternion().y(); o[2] = T.unit_quaternion().z(); o[3] = T.unit_quaternion().w();
        o[4] = T.translation()(0); o[5] = T.translation()(1); o[6] = T.translation()(2);
        counters.push_back(kfs[k]->mock_pose_sets);
    }
    std::vector<int> alive(nE), updates(nMP);
    for (int p = 0; p < nMP; ++p) {
        const Eigen::Vector3f X = mps[p].GetWorldPos();
        xout[3 * p] = X(0); xout[3 * p + 1] = X(1); xout[3 * p + 2] = X(2);
        updates[p] = mps[p].mock_normal_updates;
    }
    for (int e = 0; e < nE; ++e)   // an edge survives iff both sides still know each other
        alive[e] = (kfs[ekf[e]]->mock_matches[featOf[e]] == &mps[emp[e]] ? 1 : 0) + (mps[emp[e]].mock_obs.count(kfs[ekf[e]]) ? 2 : 0);
    wr("out_lba_pose.f32", pout); wr("out_lba_point.f32", xout); wr("out_lba_counters.i32", counters);
    wr("out_lba_alive.i32", alive); wr("out_lba_updates.i32", updates);
    std::printf("lba_cpu ok\n");
    return 0;
}

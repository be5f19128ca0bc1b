// tests/host/bow_cpu.cc -- TEST INFRASTRUCTURE (CPU tier): both ORBmatcher::SearchByBoW overloads (host/ORBmatcher_bow_b200.cc, the searches
// answered by the oracle: bow_stub.cc) over a mock keyframe pair / frame read from raw arrays; tests/test_host_bow_vs_ref.py compares the
// match vectors with those of the reference's own functions (oracle/_ref part 2).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "ref_skeleton_impl.h"
#include "ORBmatcher.h"
#include "orbslam3_b200.h"

using namespace ORB_SLAM3;
extern "C" void bow_stub_set_frame(const orbx_keypoint* kps, const uint8_t* desc, int n);

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
static std::vector<cv::KeyPoint> keys_of(const std::vector<orbx_keypoint>& k) {
    std::vector<cv::KeyPoint> o(k.size());
    for (size_t i = 0; i < k.size(); ++i) { o[i].pt.x = k[i].x; o[i].pt.y = k[i].y; o[i].size = k[i].size; o[i].angle = k[i].angle; o[i].response = k[i].response; o[i].octave = k[i].octave; o[i].class_id = k[i].class_id; }
    return o;
}
static void fill_featvec(DBoW2::FeatureVector& fv, const std::vector<int>& node) {
    for (size_t i = 0; i < node.size(); ++i) if (node[i] >= 0) fv[(DBoW2::NodeId)node[i]].push_back((unsigned int)i);
}

struct MockKF {
    std::vector<MapPoint> mps;
    KeyFrame* kf;
    MockKF(int id, const std::vector<orbx_keypoint>& k, const std::vector<uint8_t>& d, const std::vector<int>& node, const std::vector<uint8_t>& has, const std::vector<uint8_t>& bad,
           const std::vector<float>& ur, const std::vector<float>& cam4, const float* T7)
        : mps(k.size()) {
        const int n = (int)k.size();
        kf = new KeyFrame(id, cam4[0], cam4[1], cam4[2], cam4[3], 0, 0, keys_of(k), ur, std::vector<float>(8, 1.f));
        kf->mock_Tcw = Sophus::SE3f(Eigen::Quaternionf(T7[3], T7[0], T7[1], T7[2]), Eigen::Vector3f(T7[4], T7[5], T7[6]));
        cv::Mat D(n, 32, CV_8UC1);
        std::memcpy(D.ptr(0), d.data(), (size_t)n * 32);
        const_cast<cv::Mat&>(kf->mDescriptors) = D;
        fill_featvec(kf->mFeatVec, node);
        kf->mock_matches.assign(n, nullptr);
        for (int i = 0; i < n; ++i) if (has[i]) { mps[i].mock_bad = bad[i] != 0; mps[i].mnId = 1000 * id + i; kf->mock_matches[i] = &mps[i]; }
    }
    int index_of(MapPoint* p) const { return p ? (int)(p - mps.data()) : -1; }
};

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: bow_cpu <dir>\n"); return 2; }
    g_dir = argv[1];
    auto k1 = rd<orbx_keypoint>("k1.kp"), k2 = rd<orbx_keypoint>("k2.kp");
    auto d1 = rd<uint8_t>("d1.u8"), d2 = rd<uint8_t>("d2.u8");
    auto n1 = rd<int>("node1.i32"), n2 = rd<int>("node2.i32");
    auto has1 = rd<uint8_t>("has1.u8"), bad1 = rd<uint8_t>("bad1.u8"), has2 = rd<uint8_t>("has2.u8"), bad2 = rd<uint8_t>("bad2.u8");
    auto par = rd<float>("params.f32");   // nnratio, check orientation, only stereo, coarse, cam4, T1w (7), T2w (7)
    auto u1 = rd<float>("ur1.f32"), u2 = rd<float>("ur2.f32");
    const std::vector<float> cam4(par.begin() + 4, par.begin() + 8);
    MockKF K1(1, k1, d1, n1, has1, bad1, u1, cam4, &par[8]), K2(2, k2, d2, n2, has2, bad2, u2, cam4, &par[15]);
    ORBmatcher matcher(par[0], par[1] != 0);
    // (a) keyframe 1 against the FRAME made of the second feature set
    Frame F;
    F.N = (int)k2.size();
    int dummy_extractor = 0;
    F.mpORBextractorLeft = reinterpret_cast<ORBextractor*>(&dummy_extractor);
    F.mvKeys = keys_of(k2); F.mvKeysUn = F.mvKeys;
    fill_featvec(F.mFeatVec, n2);
    bow_stub_set_frame(k2.data(), d2.data(), (int)k2.size());
    std::vector<MapPoint*> vpMatches(3, nullptr);     // the function resizes it
    const int ra = matcher.SearchByBoW(K1.kf, F, vpMatches);
    std::vector<int> fa(vpMatches.size());
    for (size_t i = 0; i < vpMatches.size(); ++i) fa[i] = K1.index_of(vpMatches[i]);
    // (b) keyframe 1 against keyframe 2
    std::vector<MapPoint*> vp12;
    const int rb = matcher.SearchByBoW(K1.kf, K2.kf, vp12);
    std::vector<int> fb(vp12.size());
    for (size_t i = 0; i < vp12.size(); ++i) fb[i] = K2.index_of(vp12[i]);
    // (c) SearchForTriangulation between the two keyframes, with its own sets of features that already hold a map point
    {
        auto th1 = rd<uint8_t>("tri_has1.u8"), th2 = rd<uint8_t>("tri_has2.u8");
        for (size_t i = 0; i < th1.size(); ++i) K1.kf->mock_matches[i] = th1[i] ? &K1.mps[i] : nullptr;
        for (size_t i = 0; i < th2.size(); ++i) K2.kf->mock_matches[i] = th2[i] ? &K2.mps[i] : nullptr;
    }
    std::vector<std::pair<size_t, size_t> > pairs(2, std::make_pair((size_t)7, (size_t)7));   // the function clears it
    const int rc = matcher.SearchForTriangulation(K1.kf, K2.kf, pairs, par[2] != 0, par[3] != 0);
    std::vector<int> fc;
    for (size_t i = 0; i < pairs.size(); ++i) { fc.push_back((int)pairs[i].first); fc.push_back((int)pairs[i].second); }
    fc.push_back(rc);
    wr("out_tri_pairs.i32", fc);
    // (d) SearchForInitialization: the first feature set as the reference frame, the second as the current frame -- once resident
    //     (the registered frame), once from its own arrays (another frame is "on the device")
    auto bounds = rd<float>("bounds.f32");
    Frame::mnMinX = bounds[0]; Frame::mnMaxX = bounds[1]; Frame::mnMinY = bounds[2]; Frame::mnMaxY = bounds[3];
    Frame F1;
    F1.N = (int)k1.size(); F1.mvKeys = keys_of(k1); F1.mvKeysUn = F1.mvKeys;
    cv::Mat D1((int)k1.size(), 32, CV_8UC1), D2((int)k2.size(), 32, CV_8UC1);
    std::memcpy(D1.ptr(0), d1.data(), d1.size()); std::memcpy(D2.ptr(0), d2.data(), d2.size());
    F1.mDescriptors = D1; F.mDescriptors = D2;
    std::vector<int> init_out;
    std::vector<float> prev_out;
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) bow_stub_set_frame(k1.data(), d1.data(), 7);        // something else is resident now
        std::vector<cv::Point2f> prev(k1.size());
        for (size_t i = 0; i < k1.size(); ++i) prev[i] = F1.mvKeysUn[i].pt;
        std::vector<int> m12;
        const int rd_ = matcher.SearchForInitialization(F1, F, prev, m12, (int)par[22]);
        init_out.insert(init_out.end(), m12.begin(), m12.end());
        init_out.push_back(rd_);
        for (size_t i = 0; i < prev.size(); ++i) { prev_out.push_back(prev[i].x); prev_out.push_back(prev[i].y); }
    }
    wr("out_init.i32", init_out); wr("out_init_prev.f32", prev_out);
    wr("out_frame_match.i32", fa); wr("out_kf_match.i32", fb); wr("out_ret.i32", std::vector<int>{ra, rb});
    std::printf("bow_cpu ok\n");
    return 0;
}

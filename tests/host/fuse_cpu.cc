// tests/host/fuse_cpu.cc -- TEST INFRASTRUCTURE (CPU tier): ORBmatcher::Fuse(pKF, vpMapPoints, th) (host/ORBmatcher_fuse_b200.cc, the search answered
// by the oracle: fuse_stub.cc) over a mock keyframe and map points read from raw arrays; tests/test_host_fuse_vs_ref.py compares the
// mutation log with the best features the reference's own function finds (oracle/_ref part 2).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <set>
#include <string>
#include <vector>

#include "ref_skeleton_impl.h"
#include "ORBmatcher.h"
#include "orbslam3_b200.h"

using namespace ORB_SLAM3;
extern "C" void fuse_stub_set_log_scale_factor(float v);
extern "C" void fuse_stub_set_frame(const orbx_keypoint* kps, const uint8_t* desc, const float* ur, int n);
namespace ORB_SLAM3 { extern std::vector<int> g_fuse_log; }

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
    if (argc < 2) { std::fprintf(stderr, "usage: fuse_cpu <dir>\n"); return 2; }
    g_dir = argv[1];
    auto k2 = rd<orbx_keypoint>("kf.kp");
    auto d2 = rd<uint8_t>("kf_desc.u8"), has = rd<uint8_t>("kf_has_mp.u8");
    auto u2 = rd<float>("kf_uright.f32"), par = rd<float>("params.f32");   // fx fy cx cy bf b | minx maxx miny maxy | T (7) | th | logsf
    auto kfobs = rd<int>("kf_mp_obs.i32");                                 // Observations() of the keyframe's own map points
    auto xw = rd<float>("q_xw.f32"), nr = rd<float>("q_normal.f32"), mx = rd<float>("q_max.f32"), mn = rd<float>("q_min.f32");
    auto qd = rd<uint8_t>("q_desc.u8"), qbad = rd<uint8_t>("q_bad.u8"), qin = rd<uint8_t>("q_in_kf.u8");
    auto qobs = rd<int>("q_obs.i32"), qslot = rd<int>("q_slot.i32");       // Observations() of every query; slot: which MapPoint object a vpMapPoints entry points at (-1: NULL)
    fuse_stub_set_log_scale_factor(par[18]);
    const int N = (int)k2.size(), nq = (int)qslot.size(), nobj = (int)(xw.size() / 3);
    std::vector<cv::KeyPoint> keys(N);
    for (int i = 0; i < N; ++i) { keys[i].pt.x = k2[i].x; keys[i].pt.y = k2[i].y; keys[i].size = k2[i].size; keys[i].angle = k2[i].angle; keys[i].response = k2[i].response; keys[i].octave = k2[i].octave; keys[i].class_id = k2[i].class_id; }
    KeyFrame K(7, par[0], par[1], par[2], par[3], par[4], par[5], keys, u2, std::vector<float>(8, 1.f));
    const_cast<int&>(K.mnMinX) = (int)par[6]; const_cast<int&>(K.mnMaxX) = (int)par[7]; const_cast<int&>(K.mnMinY) = (int)par[8]; const_cast<int&>(K.mnMaxY) = (int)par[9];
    K.mock_Tcw = Sophus::SE3f(Eigen::Quaternionf(par[13], par[10], par[11], par[12]), Eigen::Vector3f(par[14], par[15], par[16]));
    cv::Mat D(N, 32, CV_8UC1);
    std::memcpy(D.ptr(0), d2.data(), (size_t)N * 32);
    const_cast<cv::Mat&>(K.mDescriptors) = D;
    std::vector<KeyFrame*> others;                                         // stand-ins for the observers that make up Observations()
    auto observers = [&](MapPoint& p, int n) { for (int k = 0; k < n; ++k) { while ((int)others.size() <= k) others.push_back((KeyFrame*)(0x1000 + 64 * others.size())); p.mock_obs[others[k]] = std::make_tuple(0, -1); } };
    std::vector<MapPoint> kfmp(N), obj(nobj);
    K.mock_matches.assign(N, nullptr);
    for (int i = 0; i < N; ++i) if (has[i]) { kfmp[i].mock_id = 100000 + i; observers(kfmp[i], kfobs[i]); kfmp[i].mock_obs[&K] = std::make_tuple(i, -1); K.mock_matches[i] = &kfmp[i]; }
    for (int j = 0; j < nobj; ++j) {
        obj[j].mock_id = j; obj[j].mock_pos = Eigen::Vector3f(xw[3 * j], xw[3 * j + 1], xw[3 * j + 2]); obj[j].mock_normal = Eigen::Vector3f(nr[3 * j], nr[3 * j + 1], nr[3 * j + 2]);
        obj[j].mock_set_distances(mn[j], mx[j]);
        obj[j].mock_desc.create(1, 32, CV_8UC1); std::memcpy(obj[j].mock_desc.ptr(0), &qd[(size_t)j * 32], 32);
        obj[j].mock_bad = qbad[j] != 0;
        observers(obj[j], qobs[j]);
        if (qin[j]) obj[j].mock_obs[&K] = std::make_tuple(0, -1);
    }
    std::vector<MapPoint*> vp(nq);
    for (int i = 0; i < nq; ++i) vp[i] = qslot[i] < 0 ? nullptr : &obj[qslot[i]];
    ORBmatcher matcher(0.6f, true);
    const int ret = matcher.Fuse(&K, vp, par[17], false);
    std::vector<int> out = g_fuse_log;
    out.push_back(ret);
    wr("out_log.i32", out);
    // (c) the Sim3 searches on a fresh copy of the keyframe and of the query objects
    {
        auto sp = rd<float>("s_params.f32");          // S8 (non-unit quaternion, translation, scale), th, ratioHamming
        auto sclaimed = rd<uint8_t>("s_claimed.u8");
        const float ss = sp[0] * sp[0] + sp[1] * sp[1] + sp[2] * sp[2] + sp[3] * sp[3], nn = std::sqrt(ss);
        Sophus::Sim3f Scw(ss, Eigen::Quaternionf(sp[3] / nn, sp[0] / nn, sp[1] / nn, sp[2] / nn), Eigen::Vector3f(sp[4], sp[5], sp[6]));
        std::vector<int> sout;
        for (int which = 0; which < 2; ++which) {
            KeyFrame K3(11 + which, par[0], par[1], par[2], par[3], par[4], par[5], keys, u2, std::vector<float>(8, 1.f));
            const_cast<int&>(K3.mnMinX) = (int)par[6]; const_cast<int&>(K3.mnMaxX) = (int)par[7]; const_cast<int&>(K3.mnMinY) = (int)par[8]; const_cast<int&>(K3.mnMaxY) = (int)par[9];
            const_cast<cv::Mat&>(K3.mDescriptors) = D;
            std::vector<MapPoint> kfmp3(N), obj3(nobj);
            K3.mock_matches.assign(N, nullptr);
            for (int i = 0; i < N; ++i) if (has[i]) { kfmp3[i].mock_id = 100000 + i; K3.mock_matches[i] = &kfmp3[i]; }
            std::vector<MapPoint*> vp3(nobj);
            for (int j = 0; j < nobj; ++j) {
                obj3[j].mock_id = j; obj3[j].mock_pos = Eigen::Vector3f(xw[3 * j], xw[3 * j + 1], xw[3 * j + 2]); obj3[j].mock_normal = Eigen::Vector3f(nr[3 * j], nr[3 * j + 1], nr[3 * j + 2]);
                obj3[j].mock_set_distances(mn[j], mx[j]);
                obj3[j].mock_desc.create(1, 32, CV_8UC1); std::memcpy(obj3[j].mock_desc.ptr(0), &qd[(size_t)j * 32], 32);
                obj3[j].mock_bad = qbad[j] != 0;
                vp3[j] = &obj3[j];
            }
            ORBmatcher m3(0.75f, true);
            g_fuse_log.clear();
            if (which == 0) {
                MapPoint occupied;
                occupied.mock_id = -2;
                std::vector<MapPoint*> vpMatched(N, nullptr);
                for (int i = 0; i < N; ++i) if (sclaimed[i]) vpMatched[i] = &occupied;
                std::vector<MapPoint*> vpMatchedB = vpMatched;
                const int r3 = m3.SearchByProjection(&K3, Scw, vp3, vpMatched, (int)sp[8], sp[9]);
                for (int i = 0; i < N; ++i) sout.push_back(vpMatched[i] ? vpMatched[i]->mock_id : -1);
                sout.push_back(r3);
                // the overload that also records the keyframe of every point: point j "comes from" the fake keyframe address 0x100000 + 64 j
                std::vector<KeyFrame*> fromKF(nobj), vpMatchedKF(N, nullptr);
                for (int j = 0; j < nobj; ++j) fromKF[j] = (KeyFrame*)(size_t)(0x100000 + 64 * j);
                const int r3b = m3.SearchByProjection(&K3, Scw, vp3, fromKF, vpMatchedB, vpMatchedKF, (int)sp[8], sp[9]);
                int same = r3b == r3 ? 1 : 0;
                for (int i = 0; i < N; ++i) {
                    if (vpMatchedB[i] != vpMatched[i]) same = 0;
                    const bool fresh = vpMatched[i] && vpMatched[i] != &occupied;
                    if (fresh != (vpMatchedKF[i] != nullptr)) same = 0;
                    if (fresh && vpMatchedKF[i] != fromKF[vpMatched[i]->mock_id]) same = 0;
                }
                sout.push_back(same);
            } else {
                std::vector<MapPoint*> repl(nobj, nullptr);
                const int r3 = m3.Fuse(&K3, Scw, vp3, sp[8], repl);
                for (int j = 0; j < nobj; ++j) sout.push_back(repl[j] ? repl[j]->mock_id : -1);
                sout.push_back(r3);
                sout.push_back((int)g_fuse_log.size() / 4);
                sout.insert(sout.end(), g_fuse_log.begin(), g_fuse_log.end());
            }
        }
        wr("out_sim3.i32", sout);
    }
    // (b) relocalisation: SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist).  The keyframe is made of the query objects
    //     (feature i holds object i where r_has[i]); the frame is the feature set above, resident on the "device".
    {
        auto rk = rd<orbx_keypoint>("r_kf.kp");
        auto rhas = rd<uint8_t>("r_has.u8"), rbad = rd<uint8_t>("r_bad.u8"), ralready = rd<uint8_t>("r_already.u8"), rclaimed = rd<uint8_t>("r_claimed.u8");
        auto rxw = rd<float>("r_xw.f32"), rmx = rd<float>("r_max.f32"), rmn = rd<float>("r_min.f32"), rpar = rd<float>("r_params.f32");   // th, ORBdist, check
        auto rdesc = rd<uint8_t>("r_desc.u8");
        const int n1 = (int)rk.size();
        std::vector<cv::KeyPoint> rkeys(n1);
        for (int i = 0; i < n1; ++i) { rkeys[i].pt.x = rk[i].x; rkeys[i].pt.y = rk[i].y; rkeys[i].angle = rk[i].angle; rkeys[i].octave = rk[i].octave; }
        KeyFrame K1(9, par[0], par[1], par[2], par[3], par[4], par[5], rkeys, std::vector<float>(n1, -1.f), std::vector<float>(8, 1.f));
        std::vector<MapPoint> mp1(n1);
        K1.mock_matches.assign(n1, nullptr);
        std::set<MapPoint*> found;
        for (int i = 0; i < n1; ++i) {
            if (!rhas[i]) continue;
            mp1[i].mock_id = i; mp1[i].mock_pos = Eigen::Vector3f(rxw[3 * i], rxw[3 * i + 1], rxw[3 * i + 2]); mp1[i].mock_set_distances(rmn[i], rmx[i]);
            mp1[i].mock_desc.create(1, 32, CV_8UC1); std::memcpy(mp1[i].mock_desc.ptr(0), &rdesc[(size_t)i * 32], 32);
            mp1[i].mock_bad = rbad[i] != 0;
            K1.mock_matches[i] = &mp1[i];
            if (ralready[i]) found.insert(&mp1[i]);
        }
        Frame F;
        F.N = N; F.mbf = par[4]; F.mb = par[5];
        Frame::fx = par[0]; Frame::fy = par[1]; Frame::cx = par[2]; Frame::cy = par[3];
        Frame::mnMinX = par[6]; Frame::mnMaxX = par[7]; Frame::mnMinY = par[8]; Frame::mnMaxY = par[9];
        int dummy_extractor = 0;
        F.mpORBextractorLeft = reinterpret_cast<ORBextractor*>(&dummy_extractor);
        F.mTcw = K.mock_Tcw;
        F.mvKeysUn = keys;
        MapPoint occupied;
        F.mvpMapPoints.assign(N, nullptr);
        for (int i = 0; i < N; ++i) if (rclaimed[i]) F.mvpMapPoints[i] = &occupied;
        fuse_stub_set_frame(k2.data(), d2.data(), u2.data(), N);
        ORBmatcher m2(0.9f, rpar[2] != 0);
        const int r2 = m2.SearchByProjection(F, &K1, found, rpar[0], (int)rpar[1]);
        std::vector<int> fm(N, -1);
        for (int i = 0; i < N; ++i) if (F.mvpMapPoints[i] && F.mvpMapPoints[i] != &occupied) fm[i] = F.mvpMapPoints[i]->mock_id;
        fm.push_back(r2);
        wr("out_reloc.i32", fm);
    }
    std::printf("fuse_cpu ok\n");
    return 0;
}

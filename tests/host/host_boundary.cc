// tests/host/host_boundary.cc -- runs the host translation units (orb_slam3_detailed_comments_b200/host/*.cc) the way
// Tracking / LocalMapping would: through the reference's own class declarations (ORB_SLAM3::ORBextractor, ORBmatcher,
// Frame::ComputeStereoMatches, Optimizer::LocalBundleAdjustment) over the skeleton map of host/refshim.  Inputs and outputs are
// raw arrays in a scratch directory; tests/test_zz_host_boundary_gpu.py writes the inputs and compares the outputs with the oracle.
//   host_boundary <dir>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "ref_skeleton_impl.h"   // also brings the reference's ORBextractor.h
#include "ORBmatcher.h"          // the reference's headers, parsed against the skeletons
#include "Optimizer.h"
#include "orb_b200_host.h"

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
static cv::Mat desc_row(const std::vector<uint8_t>& all, int i) {
    cv::Mat m(1, 32, CV_8U);
    std::memcpy(m.data, &all[(size_t)i * 32], 32);
    return m;
}
static Sophus::SE3f se3_of(const float* q7) {   // qx qy qz qw tx ty tz
    return Sophus::SE3f(Eigen::Quaternionf(q7[3], q7[0], q7[1], q7[2]), Eigen::Vector3f(q7[4], q7[5], q7[6]));
}

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: host_boundary <dir>\n"); return 2; }
    g_dir = argv[1];
    auto M = read_meta();
    const int W = (int)M["W"], H = (int)M["H"], NF = (int)M["nfeatures"];
    try {
        // ---- 1. ORBextractor::operator() on both eyes (Frame.cc:136-141 does this on two threads; one suffices here) ----
        std::vector<uint8_t> imgL = rd<uint8_t>("left.u8"), imgR = rd<uint8_t>("right.u8");
        ORBextractor exL(NF, 1.2f, 8, 20, 7), exR(NF, 1.2f, 8, 20, 7);
        cv::Mat mL(H, W, CV_8UC1, imgL.data(), (size_t)W), mR(H, W, CV_8UC1, imgR.data(), (size_t)W), dL, dR;
        std::vector<cv::KeyPoint> kL, kR;
        std::vector<int> lap = {0, 0};
        const int monoL = exL(mL, cv::Mat(), kL, dL, lap), monoR = exR(mR, cv::Mat(), kR, dR, lap);
        {
            std::vector<uint8_t> b((const uint8_t*)kL.data(), (const uint8_t*)kL.data() + kL.size() * 28);
            wr("out_kpsL.bin", b);
            std::vector<uint8_t> d((size_t)kL.size() * 32);
            for (size_t i = 0; i < kL.size(); ++i) std::memcpy(&d[i * 32], dL.ptr((int)i), 32);
            wr("out_descL.bin", d);
            std::vector<uint8_t> b2((const uint8_t*)kR.data(), (const uint8_t*)kR.data() + kR.size() * 28);
            wr("out_kpsR.bin", b2);
            std::vector<uint8_t> pyr;
            for (int l = 0; l < exL.GetLevels(); ++l) {
                const cv::Mat& m = exL.mvImagePyramid[l];
                for (int r = 0; r < m.rows; ++r) pyr.insert(pyr.end(), m.ptr(r), m.ptr(r) + m.cols);
            }
            wr("out_pyramidL.u8", pyr);
            std::vector<float> sc = exL.GetScaleFactors();
            wr("out_scale.f32", sc);
            wr("out_mono.i32", std::vector<int>{monoL, monoR, (int)kL.size(), (int)kR.size()});
        }
        // the mono call site (Frame.cc:380): vLappingArea {0, 1000}
        {
            std::vector<int> lap2 = {0, 1000};
            std::vector<cv::KeyPoint> k2;
            cv::Mat d2;
            ORBextractor exM(NF, 1.2f, 8, 20, 7);
            const int mono2 = exM(mL, cv::Mat(), k2, d2, lap2);
            std::vector<uint8_t> b((const uint8_t*)k2.data(), (const uint8_t*)k2.data() + k2.size() * 28);
            wr("out_kpsM.bin", b);
            wr("out_monoM.i32", std::vector<int>{mono2, (int)k2.size()});
            orb_b200_release(&exM);
            // an empty input returns -1 (ORBextractor.cc:1560-1561)
            cv::Mat none;
            std::vector<cv::KeyPoint> k3;
            cv::Mat d3;
            wr("out_empty.i32", std::vector<int>{exL(none, cv::Mat(), k3, d3, lap)});
            exL(mL, cv::Mat(), kL, dL, lap);   // restore exL's device-resident batch for the steps below
        }
        // ---- 2. Frame::ComputeStereoMatches ----
        Frame::fx = (float)M["fx"]; Frame::fy = (float)M["fy"]; Frame::cx = (float)M["cx"]; Frame::cy = (float)M["cy"];
        Frame::mnMinX = 0; Frame::mnMaxX = (float)W; Frame::mnMinY = 0; Frame::mnMaxY = (float)H;
        Frame F;
        F.mpORBextractorLeft = &exL; F.mpORBextractorRight = &exR;
        F.mbf = (float)M["bf"]; F.mb = (float)M["b"];
        F.N = (int)kL.size();
        F.mvKeys = kL; F.mvKeysUn = kL; F.mvKeysRight = kR;
        F.mDescriptors = dL;
        F.mvpMapPoints.assign(F.N, nullptr);
        F.mvbOutlier.assign(F.N, false);
        F.ComputeStereoMatches();
        wr("out_uright.f32", F.mvuRight);
        wr("out_depth.f32", F.mvDepth);

        // ---- 3. SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints) ----
        if (M.count("loc_n")) {
            const int nq = (int)M["loc_n"];
            auto px = rd<float>("loc_px.f32"), py = rd<float>("loc_py.f32"), pxr = rd<float>("loc_pxr.f32"), vc = rd<float>("loc_vcos.f32"),
                 dep = rd<float>("loc_depth.f32");
            auto lv = rd<int>("loc_level.i32");
            auto inview = rd<uint8_t>("loc_inview.u8"), bad = rd<uint8_t>("loc_bad.u8"), qd = rd<uint8_t>("loc_desc.u8"), claimed = rd<uint8_t>("loc_claimed.u8");
            std::vector<MapPoint> mps(nq);
            std::vector<MapPoint*> vp(nq);
            KeyFrame dummyKF(0, 1, 1, 0, 0, 1, 1, {}, {}, {});
            for (int i = 0; i < nq; ++i) {
                MapPoint& p = mps[i];
                p.mnId = i; p.mTrackProjX = px[i]; p.mTrackProjY = py[i]; p.mTrackProjXR = pxr[i]; p.mnTrackScaleLevel = lv[i];
                p.mTrackViewCos = vc[i]; p.mTrackDepth = dep[i]; p.mbTrackInView = inview[i] != 0; p.mock_bad = bad[i] != 0;
                p.mock_desc = desc_row(qd, i);
                p.mock_obs[&dummyKF] = std::make_tuple(0, -1);          // local map points always have observations
                vp[i] = &p;
            }
            MapPoint holder, fresh;                                      // holder: Observations() > 0; fresh: a point without observations
            holder.mock_obs[&dummyKF] = std::make_tuple(0, -1);
            for (int i = 0; i < F.N; ++i) F.mvpMapPoints[i] = claimed[i] == 1 ? &holder : (claimed[i] == 2 ? &fresh : nullptr);
            ORBmatcher matcher((float)M["loc_nnratio"], true);
            const int nm = matcher.SearchByProjection(F, vp, (float)M["loc_th"], M["loc_far"] != 0, (float)M["loc_thfar"]);
            std::vector<int> res(F.N + 1, -1);
            for (int i = 0; i < F.N; ++i) {
                MapPoint* p = F.mvpMapPoints[i];
                res[i] = !p ? -1 : (p == &holder ? -2 : (p == &fresh ? -3 : (int)(p - mps.data())));
            }
            res[F.N] = nm;
            wr("out_local.i32", res);
        }
        // ---- 4. SearchByProjection(CurrentFrame, LastFrame, th, bMono) ----
        if (M.count("last_n")) {
            const int nl = (int)M["last_n"];
            auto has = rd<uint8_t>("last_has.u8"), outl = rd<uint8_t>("last_outlier.u8"), obsp = rd<uint8_t>("last_obspos.u8"), qd = rd<uint8_t>("last_desc.u8");
            auto xw = rd<float>("last_xw.f32"), ang = rd<float>("last_angle.f32"), Tc = rd<float>("last_Tcw.f32"), Tl = rd<float>("last_Tlw.f32");
            auto oct = rd<int>("last_octave.i32");
            KeyFrame dummyKF(0, 1, 1, 0, 0, 1, 1, {}, {}, {});
            std::vector<MapPoint> mps(nl);
            Frame Last;
            Last.N = nl; Last.mb = F.mb; Last.mbf = F.mbf;
            Last.mvKeys.resize(nl); Last.mvKeysUn.resize(nl);
            Last.mvpMapPoints.assign(nl, nullptr); Last.mvbOutlier.assign(nl, false);
            for (int i = 0; i < nl; ++i) {
                Last.mvKeys[i].octave = oct[i]; Last.mvKeysUn[i].octave = oct[i]; Last.mvKeysUn[i].angle = ang[i];
                Last.mvbOutlier[i] = outl[i] != 0;
                if (has[i]) {
                    MapPoint& p = mps[i];
                    p.mnId = i; p.mock_pos = Eigen::Vector3f(xw[3 * i], xw[3 * i + 1], xw[3 * i + 2]); p.mock_desc = desc_row(qd, i);
                    if (obsp[i]) p.mock_obs[&dummyKF] = std::make_tuple(0, -1);
                    Last.mvpMapPoints[i] = &p;
                }
            }
            Last.mTcw = se3_of(Tl.data());
            Frame Cur = F;                                   // the same device-resident frame, no map points yet (Tracking.cc:3367)
            Cur.mvpMapPoints.assign(Cur.N, nullptr);
            Cur.mTcw = se3_of(Tc.data());
            ORBmatcher matcher(0.9f, M["last_checkori"] != 0);
            const int nm = matcher.SearchByProjection(Cur, Last, (float)M["last_th"], M["last_mono"] != 0);
            std::vector<int> res(Cur.N + 1, -1);
            for (int i = 0; i < Cur.N; ++i)
                if (Cur.mvpMapPoints[i]) res[i] = (int)(Cur.mvpMapPoints[i] - mps.data());
            res[Cur.N] = nm;
            wr("out_last.i32", res);
        }
        // ---- 5. Optimizer::LocalBundleAdjustment over a mock map ----
        if (M.count("lba_nkf")) {
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
                kfs[k] = new (block + k) KeyFrame(100 + k, (float)M["fx"], (float)M["fy"], (float)M["cx"], (float)M["cy"], (float)M["bf"], (float)M["b"], keys[k], ur[k], sig[k]);
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
            Optimizer::LocalBundleAdjustment(pKF, &stop, &map, nFixed, nOpt, nMPs, nEdges);
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
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "host_boundary: %s\n", e.what());
        return 1;
    }
    std::printf("host_boundary ok\n");
    return 0;
}

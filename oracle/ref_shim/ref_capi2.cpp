// oracle/ref_shim/ref_capi2.cpp -- TEST INFRASTRUCTURE.
//
// C entry points around member functions of the REFERENCE compiled verbatim: oracle/Makefile (target ref2) has
// oracle/tools/extract_functions.py cut them out of /root/reference/src/{ORBmatcher.cc, Frame.cc, MapPoint.cc,
// CameraModels/Pinhole.cpp} into oracle/_ref/gen/ref2_gen.inc (a build artefact, never committed) and compiles that text here,
// against the reference's own include/ORBmatcher.h and the skeleton classes of skel/ref_frame_skel.h.  The wrappers only marshal
// flat arrays into those objects.  tests/test_oracle_vs_ref_matcher.py compares the oracle's restatements with these.
#include "skel/ref_frame_skel.h"
#include "ORBmatcher.h"   // the reference's own header (its MapPoint.h / KeyFrame.h / Frame.h includes are guarded out by the skeleton)

#include <cstring>
#include <memory>

namespace ORB_SLAM3 {
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv, Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
MapPoint* MapPoint::g_last_asked = nullptr;
#include "../_ref/gen/ref2_gen.inc"
}  // namespace ORB_SLAM3

using namespace ORB_SLAM3;

namespace {
struct Kp28 { float x, y, size, angle, response; int octave, class_id; };

struct Holder {
    Frame F;
    Pinhole cam;
    cv::Mat desc;
    std::vector<MapPoint*> owned;
    ~Holder() { for (MapPoint* p : owned) delete p; }
};

void set_statics(const float* bounds4) {
    Frame::mnMinX = bounds4[0]; Frame::mnMaxX = bounds4[1]; Frame::mnMinY = bounds4[2]; Frame::mnMaxY = bounds4[3];
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (Frame::mnMaxX - Frame::mnMinX);     // Frame.cc:187-188
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (Frame::mnMaxY - Frame::mnMinY);
}

cv::Mat desc_mat(const unsigned char* d, int n) {
    cv::Mat m(std::max(n, 1), 32, CV_8UC1);
    if (n) std::memcpy(m.data, d, (size_t)n * 32);
    return m;
}

void set_pose(Frame& F, const float* T7) {
    if (!T7) return;
    for (int i = 0; i < 4; ++i) F.mTcw.q[i] = T7[i];
    for (int i = 0; i < 3; ++i) F.mTcw.t[i] = T7[4 + i];
}
}  // namespace

extern "C" {

// a Frame (Nleft == -1) from flat arrays: mvKeys = mvKeysUn (rectified), mDescriptors, mvuRight (or all -1), the grid
void* ref2_frame_create(const void* kps, const unsigned char* desc, const float* uright, int N, const float* bounds4,
                        const float* scaleFactors, int nlevels, const float* cam6, const float* Tcw7) {
    Holder* H = new Holder();
    Frame& F = H->F;
    set_statics(bounds4);
    const Kp28* k = (const Kp28*)kps;
    F.N = N;
    F.mvKeys.resize(N);
    for (int i = 0; i < N; ++i) {
        cv::KeyPoint& o = F.mvKeys[i];
        o.pt.x = k[i].x; o.pt.y = k[i].y; o.size = k[i].size; o.angle = k[i].angle; o.response = k[i].response; o.octave = k[i].octave; o.class_id = k[i].class_id;
    }
    F.mvKeysUn = F.mvKeys;
    F.mDescriptors = desc_mat(desc, N);
    F.mvuRight.assign(N, -1.0f);
    if (uright) for (int i = 0; i < N; ++i) F.mvuRight[i] = uright[i];
    F.mvDepth.assign(N, -1.0f);
    F.mvpMapPoints.assign(N, static_cast<MapPoint*>(NULL));
    F.mvbOutlier.assign(N, false);
    F.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    F.mvInvScaleFactors.resize(nlevels);
    for (int l = 0; l < nlevels; ++l) F.mvInvScaleFactors[l] = 1.0f / scaleFactors[l];
    F.mnScaleLevels = nlevels;
    F.mfLogScaleFactor = nlevels > 1 ? logf(scaleFactors[1]) : 0.f;        // Frame.cc:121 mfLogScaleFactor = log(mfScaleFactor)
    H->cam.mvParameters = {cam6[0], cam6[1], cam6[2], cam6[3]};
    F.mpCamera = &H->cam;
    F.mbf = cam6[4]; F.mb = cam6[5];
    set_pose(F, Tcw7);
    F.AssignFeaturesToGrid();
    return H;
}
void ref2_frame_destroy(void* h) { delete (Holder*)h; }

int ref2_descriptor_distance(const unsigned char* a, const unsigned char* b) {
    cv::Mat A = desc_mat(a, 1), B = desc_mat(b, 1);
    return ORBmatcher::DescriptorDistance(A, B);
}

struct OpenMatcher : public ORBmatcher {
    using ORBmatcher::ORBmatcher;
    using ORBmatcher::ComputeThreeMaxima;
    using ORBmatcher::RadiusByViewingCos;
};

void ref2_three_maxima(const int* sizes, int L, int* out3) {
    std::vector<std::vector<int>> h(L);
    for (int i = 0; i < L; ++i) h[i].assign(sizes[i], 0);
    OpenMatcher m(0.6f, true);
    int a = -1, b = -1, c = -1;
    m.ComputeThreeMaxima(h.data(), L, a, b, c);
    out3[0] = a; out3[1] = b; out3[2] = c;
}

void ref2_constants(int* out3, float* radius2) {
    out3[0] = ORBmatcher::TH_LOW; out3[1] = ORBmatcher::TH_HIGH; out3[2] = ORBmatcher::HISTO_LENGTH;
    OpenMatcher m(0.6f, true);
    radius2[0] = m.RadiusByViewingCos(0.9985f); radius2[1] = m.RadiusByViewingCos(0.99f);
}

int ref2_get_features_in_area(void* h, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap) {
    const std::vector<size_t> v = ((Holder*)h)->F.GetFeaturesInArea(x, y, r, minLevel, maxLevel);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int)v[i];
    return (int)v.size();
}

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints): one MapPoint per query carrying the
// fields Frame::isInFrustum wrote; claimed[idx] != 0 => F.mvpMapPoints[idx] holds a map point with Observations() > 0 on entry.
// match[q] = the feature that now holds query q, or -1.  Returns the reference's return value.
int ref2_search_local(void* h, int nq, const float* projx, const float* projy, const float* projxr, const int* level, const float* viewcos,
                      const float* trackdepth, const unsigned char* qdesc, const unsigned char* claimed, float th, float nnratio, int bFar,
                      float thFar, int* match) {
    Holder* H = (Holder*)h;
    Frame& F = H->F;
    MapPoint occupied;
    occupied.nObs = 1;
    for (int i = 0; i < F.N; ++i) F.mvpMapPoints[i] = (claimed && claimed[i]) ? &occupied : static_cast<MapPoint*>(NULL);
    std::vector<MapPoint> mps(nq);
    std::vector<MapPoint*> vp(nq);
    for (int q = 0; q < nq; ++q) {
        MapPoint& m = mps[q];
        m.mbTrackInView = true; m.mTrackProjX = projx[q]; m.mTrackProjY = projy[q]; m.mTrackProjXR = projxr[q]; m.mnTrackScaleLevel = level[q];
        m.mTrackViewCos = viewcos[q]; m.mTrackDepth = trackdepth ? trackdepth[q] : 0.f; m.mDescriptor = desc_mat(qdesc + 32 * (size_t)q, 1);
        m.nObs = 1; m.query_index = q;
        vp[q] = &m;
    }
    ORBmatcher matcher(nnratio, true);
    const int n = matcher.SearchByProjection(F, vp, th, bFar != 0, thFar);
    for (int q = 0; q < nq; ++q) match[q] = -1;
    for (int i = 0; i < F.N; ++i)
        if (F.mvpMapPoints[i] && F.mvpMapPoints[i] != &occupied) match[F.mvpMapPoints[i]->query_index] = i;
    for (int i = 0; i < F.N; ++i) F.mvpMapPoints[i] = NULL;
    return n;
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono): LastFrame = nq features, each with a map
// point (world position, descriptor, Observations() > 0 or not), mvKeys[i].octave / mvKeysUn[i].angle as given, pose Tlw.
// feat_match[idx] = query index now held by CurrentFrame.mvpMapPoints[idx], or -1; *direction = 0 / 1 (bForward) / 2 (bBackward) as
// the reference computes it (ORBmatcher.cc:1964-1971).
int ref2_search_last(void* h, const float* Tcw7, const float* Tlw7, int nq, const float* xw, const int* lastOctave, const float* lastAngle,
                     const unsigned char* qdesc, const unsigned char* obs_pos, float th, int checkOri, int bMono, int* feat_match, int* direction) {
    Holder* H = (Holder*)h;
    Frame& F = H->F;
    set_pose(F, Tcw7);
    for (int i = 0; i < F.N; ++i) F.mvpMapPoints[i] = NULL;   // Tracking.cc:3367
    Frame L;
    set_pose(L, Tlw7);
    L.N = nq;
    L.mvKeys.resize(nq); L.mvKeysUn.resize(nq);
    L.mvbOutlier.assign(nq, false);
    std::vector<MapPoint> mps(nq);
    L.mvpMapPoints.resize(nq);
    for (int q = 0; q < nq; ++q) {
        L.mvKeys[q].octave = lastOctave[q]; L.mvKeysUn[q].octave = lastOctave[q]; L.mvKeysUn[q].angle = lastAngle[q]; L.mvKeys[q].angle = lastAngle[q];
        MapPoint& m = mps[q];
        m.mWorldPos = Eigen::Vector3f(xw[3 * q], xw[3 * q + 1], xw[3 * q + 2]);
        m.mDescriptor = desc_mat(qdesc + 32 * (size_t)q, 1);
        m.nObs = (!obs_pos || obs_pos[q]) ? 1 : 0;
        m.query_index = q;
        L.mvpMapPoints[q] = &m;
    }
    {   // the two lines of ORBmatcher.cc:1964-1971 that decide the search direction, for the oracle's `direction` argument
        const Eigen::Vector3f twc = F.GetPose().inverse().translation();
        const Eigen::Vector3f tlc = L.GetPose() * twc;
        *direction = (tlc(2) > F.mb && !bMono) ? 1 : ((-tlc(2) > F.mb && !bMono) ? 2 : 0);
    }
    ORBmatcher matcher(0.9f, checkOri != 0);
    const int n = matcher.SearchByProjection(F, L, th, bMono != 0);
    for (int i = 0; i < F.N; ++i) {
        feat_match[i] = F.mvpMapPoints[i] ? F.mvpMapPoints[i]->query_index : -1;
        F.mvpMapPoints[i] = NULL;
    }
    return n;
}

// Frame::isInFrustum(MapPoint*, viewingCosLimit) for np points; maxDist / minDist = MapPoint::mfMaxDistance / mfMinDistance
int ref2_is_in_frustum(void* h, const float* Rcw9, const float* tcw, const float* Ow, int np, const float* xw, const float* normal,
                       const float* maxDist, const float* minDist, float limit, unsigned char* in_view, float* projx, float* projy, float* projxr,
                       int* level, float* viewcos, float* depth) {
    Frame& F = ((Holder*)h)->F;
    for (int i = 0; i < 9; ++i) F.mRcw.v[i] = Rcw9[i];
    for (int i = 0; i < 3; ++i) { F.mtcw.v[i] = tcw[i]; F.mOw.v[i] = Ow[i]; }
    int nvis = 0;
    for (int i = 0; i < np; ++i) {
        MapPoint m;
        m.mWorldPos = Eigen::Vector3f(xw[3 * i], xw[3 * i + 1], xw[3 * i + 2]);
        m.mNormalVector = Eigen::Vector3f(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
        m.mfMaxDistance = maxDist[i]; m.mfMinDistance = minDist[i];
        m.mnTrackScaleLevel = -1; m.mTrackProjXR = -1; m.mTrackViewCos = 0; m.mTrackDepth = 0;
        const bool vis = F.isInFrustum(&m, limit);
        in_view[i] = m.mbTrackInView ? 1 : 0;
        projx[i] = m.mTrackProjX; projy[i] = m.mTrackProjY; projxr[i] = m.mTrackProjXR; level[i] = m.mnTrackScaleLevel; viewcos[i] = m.mTrackViewCos;
        depth[i] = m.mTrackDepth;
        nvis += vis;
    }
    return nvis;
}

// Frame::ComputeStereoMatches: left features from the frame handle, right features given; the two extractors are the REFERENCE's own
// ORBextractor objects (ref_capi.cpp) after they extracted the two images (mvImagePyramid).
void ref2_stereo_matches(void* h, void* exL, void* exR, const void* kpsR, const unsigned char* descR, int NR, float* uright, float* depth) {
    Frame& F = ((Holder*)h)->F;
    F.mpORBextractorLeft = (ORBextractor*)exL;
    F.mpORBextractorRight = (ORBextractor*)exR;
    const Kp28* k = (const Kp28*)kpsR;
    F.mvKeysRight.resize(NR);
    for (int i = 0; i < NR; ++i) {
        cv::KeyPoint& o = F.mvKeysRight[i];
        o.pt.x = k[i].x; o.pt.y = k[i].y; o.size = k[i].size; o.angle = k[i].angle; o.response = k[i].response; o.octave = k[i].octave; o.class_id = k[i].class_id;
    }
    F.mDescriptorsRight = desc_mat(descR, NR);
    F.ComputeStereoMatches();
    for (int i = 0; i < F.N; ++i) { uright[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; }
}

}  // extern "C"

// ---- part 2b: the KeyFrame-typed matchers ------------------------------------------------------------------------------------
namespace {
struct KFHolder {
    KeyFrame K;
    Pinhole cam;
    std::unique_ptr<MapPoint[]> mps;   // one per feature; mvpMapPoints[i] points here when the feature holds a map point
};
void fill_featvec(DBoW2::FeatureVector& fv, const int* node, int n) {
    fv.clear();
    if (!node) return;
    for (int i = 0; i < n; ++i)
        if (node[i] >= 0) fv.addFeature((DBoW2::NodeId)node[i], (unsigned)i);   // ascending i, as TemplatedVocabulary::transform adds them
}
}  // namespace

extern "C" {

// A single-camera KeyFrame (NLeft == -1, mpCamera2 == NULL) from flat arrays.  node[i] = mFeatVec node of feature i or -1;
// has_mp[i] != 0: feature i holds a map point (bad[i] != 0: MapPoint::isBad()); the map point of feature i carries query_index = i.
void* ref2_kf_create(const void* kps, const unsigned char* desc, const float* uright, int N, const int* node, const unsigned char* has_mp,
                     const unsigned char* bad, const float* scaleFactors, const float* levelSigma2, int nlevels, const float* cam4, const float* Tcw7) {
    KFHolder* H = new KFHolder();
    KeyFrame& K = H->K;
    const Kp28* k = (const Kp28*)kps;
    K.N = N;
    K.mvKeys.resize(N);
    for (int i = 0; i < N; ++i) {
        cv::KeyPoint& o = K.mvKeys[i];
        o.pt.x = k[i].x; o.pt.y = k[i].y; o.size = k[i].size; o.angle = k[i].angle; o.response = k[i].response; o.octave = k[i].octave; o.class_id = k[i].class_id;
    }
    K.mvKeysUn = K.mvKeys;
    K.mDescriptors = desc_mat(desc, N);
    K.mvuRight.assign(N, -1.0f);
    if (uright) for (int i = 0; i < N; ++i) K.mvuRight[i] = uright[i];
    fill_featvec(K.mFeatVec, node, N);
    H->mps.reset(new MapPoint[std::max(N, 1)]);
    K.mvpMapPoints.assign(N, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < N; ++i) {
        H->mps[i].query_index = i;
        H->mps[i].mbBad = bad && bad[i];
        if (has_mp && has_mp[i]) K.mvpMapPoints[i] = &H->mps[i];
    }
    if (scaleFactors) K.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    if (levelSigma2) K.mvLevelSigma2.assign(levelSigma2, levelSigma2 + nlevels);
    H->cam.mvParameters = {cam4[0], cam4[1], cam4[2], cam4[3]};
    K.mpCamera = &H->cam;
    if (Tcw7) {
        for (int i = 0; i < 4; ++i) K.mTcw.q[i] = Tcw7[i];
        for (int i = 0; i < 3; ++i) K.mTcw.t[i] = Tcw7[4 + i];
    }
    return H;
}
void ref2_kf_destroy(void* h) { delete (KFHolder*)h; }

// ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches): feat_node = F.mFeatVec;
// feat_match[idxF] = the KF feature whose map point F's feature idxF now holds, or -1
int ref2_search_bow(void* hF, const int* feat_node, void* hKF, float nnratio, int checkOri, int* feat_match) {
    Frame& F = ((Holder*)hF)->F;
    KeyFrame& K = ((KFHolder*)hKF)->K;
    fill_featvec(F.mFeatVec, feat_node, F.N);
    std::vector<MapPoint*> vp;
    ORBmatcher matcher(nnratio, checkOri != 0);
    const int n = matcher.SearchByBoW(&K, F, vp);
    for (int i = 0; i < F.N; ++i) feat_match[i] = vp[i] ? vp[i]->query_index : -1;
    return n;
}

// ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12): match12[idx1] = the KF2 feature, or -1
int ref2_search_bow_kf(void* hKF1, void* hKF2, float nnratio, int checkOri, int* match12) {
    KeyFrame& K1 = ((KFHolder*)hKF1)->K;
    KeyFrame& K2 = ((KFHolder*)hKF2)->K;
    std::vector<MapPoint*> vp;
    ORBmatcher matcher(nnratio, checkOri != 0);
    const int n = matcher.SearchByBoW(&K1, &K2, vp);
    for (int i = 0; i < K1.N; ++i) match12[i] = vp[i] ? vp[i]->query_index : -1;
    return n;
}

// ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize): prevMatched = 2 floats per F1 feature, in / out
int ref2_search_initialization(void* hF1, void* hF2, float* prevMatched, int windowSize, float nnratio, int checkOri, int* matches12) {
    Frame& F1 = ((Holder*)hF1)->F;
    Frame& F2 = ((Holder*)hF2)->F;
    std::vector<cv::Point2f> prev(F1.N);
    for (int i = 0; i < F1.N; ++i) prev[i] = cv::Point2f(prevMatched[2 * i], prevMatched[2 * i + 1]);
    std::vector<int> m12;
    ORBmatcher matcher(nnratio, checkOri != 0);
    const int n = matcher.SearchForInitialization(F1, F2, prev, m12, windowSize);
    for (int i = 0; i < F1.N; ++i) { matches12[i] = m12[i]; prevMatched[2 * i] = prev[i].x; prevMatched[2 * i + 1] = prev[i].y; }
    return n;
}

// ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse): match12[idx1] = idx2 or -1.  F12 / ep = the
// fundamental matrix Pinhole::epipolarConstrain builds and the epipole, recomputed here with the same (miniature) Eigen / Sophus calls
// so that the test can hand the oracle the values its caller would.
int ref2_search_triangulation(void* hKF1, void* hKF2, int bOnlyStereo, int bCoarse, int checkOri, int* match12, float* F12_9, float* ep2) {
    KeyFrame& K1 = ((KFHolder*)hKF1)->K;
    KeyFrame& K2 = ((KFHolder*)hKF2)->K;
    std::vector<std::pair<size_t, size_t> > pairs;
    ORBmatcher matcher(0.6f, checkOri != 0);
    const int n = matcher.SearchForTriangulation(&K1, &K2, pairs, bOnlyStereo != 0, bCoarse != 0);
    for (int i = 0; i < K1.N; ++i) match12[i] = -1;
    for (size_t i = 0; i < pairs.size(); ++i) match12[pairs[i].first] = (int)pairs[i].second;
    {   // ORBmatcher.cc:1052-1070 and Pinhole.cpp:191-194
        Sophus::SE3f T1w = K1.GetPose(), T2w = K2.GetPose(), Tw2 = K2.GetPoseInverse();
        Eigen::Vector3f Cw = K1.GetCameraCenter();
        Eigen::Vector3f C2 = T2w * Cw;
        Eigen::Vector2f ep = K2.mpCamera->project(C2);
        Sophus::SE3f T12 = T1w * Tw2;
        Eigen::Matrix3f R12 = T12.rotationMatrix();
        Eigen::Vector3f t12 = T12.translation();
        Eigen::Matrix3f t12x = Sophus::SO3f::hat(t12);
        Eigen::Matrix3f Ka = K1.mpCamera->toK_(), Kb = K2.mpCamera->toK_();
        Eigen::Matrix3f F12 = Ka.transpose().inverse() * t12x * R12 * Kb.inverse();
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) F12_9[3 * r + c] = F12(r, c);
        ep2[0] = ep(0); ep2[1] = ep(1);
    }
    return n;
}

}  // extern "C"

// ---- part 2c: MapPoint maintenance ------------------------------------------------------------------------------------------------
extern "C" {

// MapPoint::ComputeDistinctiveDescriptors: observation i = feature 0 of keyframe i (descriptor descs[i], KeyFrame::isBad() = kf_bad[i]);
// the keyframes live in one array, so std::map<KeyFrame*, ...> walks them in index order.  Writes mDescriptor; returns 0 when the
// function returned before choosing (no observation / all keyframes bad), else 1.
int ref2_distinctive_descriptor(int nObs, const unsigned char* descs, const unsigned char* kf_bad, unsigned char* out32) {
    std::unique_ptr<KeyFrame[]> kfs(new KeyFrame[std::max(nObs, 1)]);
    MapPoint mp;
    for (int i = 0; i < nObs; ++i) {
        kfs[i].mDescriptors = desc_mat(descs + 32 * (size_t)i, 1);
        kfs[i].mbBad = kf_bad && kf_bad[i];
        mp.mObservations[&kfs[i]] = std::make_tuple(0, -1);
    }
    mp.mDescriptor = cv::Mat();
    mp.ComputeDistinctiveDescriptors();
    if (mp.mDescriptor.empty()) return 0;
    std::memcpy(out32, mp.mDescriptor.data, 32);
    return 1;
}

// MapPoint::UpdateNormalAndDepth: observation i is seen from camera centre centers[3 i ..] (a keyframe with identity rotation and
// translation -centre); the reference keyframe is observation ref, its feature's octave = level.
void ref2_update_normal_and_depth(int nObs, const float* centers, const float* pos, int ref, int level, const float* scaleFactors, int nlevels,
                                  float* normal3, float* maxmin2) {
    std::unique_ptr<KeyFrame[]> kfs(new KeyFrame[std::max(nObs, 1)]);
    MapPoint mp;
    for (int i = 0; i < nObs; ++i) {
        KeyFrame& K = kfs[i];
        for (int k = 0; k < 3; ++k) K.mTcw.t[k] = -centers[3 * i + k];
        K.mvKeysUn.resize(1);
        K.mvKeysUn[0].octave = level;
        K.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
        K.mnScaleLevels = nlevels;
        mp.mObservations[&K] = std::make_tuple(0, -1);
    }
    mp.mpRefKF = &kfs[ref];
    mp.mWorldPos = Eigen::Vector3f(pos[0], pos[1], pos[2]);
    mp.UpdateNormalAndDepth();
    for (int k = 0; k < 3; ++k) normal3[k] = mp.mNormalVector(k);
    maxmin2[0] = mp.mfMaxDistance; maxmin2[1] = mp.mfMinDistance;
}

}  // extern "C"

// ---- part 2d: projections into a keyframe ---------------------------------------------------------------------------------------------
namespace {
struct QuerySet {   // the map points handed to a search: world position, normal, distance members, descriptor
    std::unique_ptr<MapPoint[]> mp;
    std::vector<MapPoint*> ptr;
    QuerySet(int n, const float* xw, const float* normal, const float* maxD, const float* minD, const unsigned char* desc, const unsigned char* bad) : mp(new MapPoint[std::max(n, 1)]), ptr(n) {
        for (int q = 0; q < n; ++q) {
            MapPoint& m = mp[q];
            m.mWorldPos = Eigen::Vector3f(xw[3 * q], xw[3 * q + 1], xw[3 * q + 2]);
            if (normal) m.mNormalVector = Eigen::Vector3f(normal[3 * q], normal[3 * q + 1], normal[3 * q + 2]);
            m.mfMaxDistance = maxD[q]; m.mfMinDistance = minD[q];
            m.mDescriptor = desc_mat(desc + 32 * (size_t)q, 1);
            m.mbBad = bad && bad[q];
            m.query_index = q;
            ptr[q] = &m;
        }
    }
};
void set_sim3(Sophus::Sim3f& S, const float* s8) {
    for (int i = 0; i < 4; ++i) S.q[i] = s8[i];
    for (int i = 0; i < 3; ++i) S.t[i] = s8[4 + i];
}
void export_scw(const Sophus::Sim3f& Scw, float* Tcw7, float* Ow3) {   // ORBmatcher.cc:503-504 / 1554-1555
    Sophus::SE3f Tcw = Sophus::SE3f(Scw.rotationMatrix(), Scw.translation() / Scw.scale());
    Eigen::Vector3f Ow = Tcw.inverse().translation();
    for (int i = 0; i < 4; ++i) Tcw7[i] = Tcw.q[i];
    for (int i = 0; i < 3; ++i) { Tcw7[4 + i] = Tcw.t[i]; Ow3[i] = Ow(i); }
}
}  // namespace

extern "C" {

// what the KeyFrame constructor copies from its Frame (KeyFrame.cc:60-100): the grid, the image bounds (as int), the cell sizes, the
// scale tables, the intrinsics
void ref2_kf_set_geometry(void* hKF, void* hF, const float* invLevelSigma2, float bf) {
    KeyFrame& K = ((KFHolder*)hKF)->K;
    Frame& F = ((Holder*)hF)->F;
    K.mGrid.assign(FRAME_GRID_COLS, std::vector<std::vector<size_t> >(FRAME_GRID_ROWS));
    for (int i = 0; i < FRAME_GRID_COLS; ++i)
        for (int j = 0; j < FRAME_GRID_ROWS; ++j) K.mGrid[i][j] = F.mGrid[i][j];
    K.mnMinX = (int)Frame::mnMinX; K.mnMinY = (int)Frame::mnMinY; K.mnMaxX = (int)Frame::mnMaxX; K.mnMaxY = (int)Frame::mnMaxY;
    K.mfGridElementWidthInv = Frame::mfGridElementWidthInv; K.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
    K.mfLogScaleFactor = F.mfLogScaleFactor; K.mnScaleLevels = F.mnScaleLevels;
    K.mvScaleFactors = F.mvScaleFactors;
    K.mvInvLevelSigma2.assign(invLevelSigma2, invLevelSigma2 + F.mnScaleLevels);
    const std::vector<float>& c = ((KFHolder*)hKF)->cam.mvParameters;
    K.fx = c[0]; K.fy = c[1]; K.cx = c[2]; K.cy = c[3]; K.mbf = bf;
}

// the keyframe's own map points (features with has_mp): world position, distance members, descriptor
void ref2_kf_set_mappoints(void* hKF, const float* xw, const float* maxD, const float* minD, const unsigned char* desc) {
    KFHolder* H = (KFHolder*)hKF;
    for (int i = 0; i < H->K.N; ++i) {
        MapPoint& m = H->mps[i];
        m.mWorldPos = Eigen::Vector3f(xw[3 * i], xw[3 * i + 1], xw[3 * i + 2]);
        m.mfMaxDistance = maxD[i]; m.mfMinDistance = minD[i];
        m.mDescriptor = desc_mat(desc + 32 * (size_t)i, 1);
    }
}

// ORBmatcher::Fuse(pKF, vpMapPoints, th, bRight = false): fused[q] = the keyframe feature query q was fused into, or -1.
// in_kf[q] != 0: MapPoint::IsInKeyFrame(pKF).  Features with has_mp hold a map point on entry (Replace branch), the others are filled
// by AddMapPoint as the function goes.
int ref2_fuse(void* hKF, int nq, const float* xw, const float* normal, const float* maxD, const float* minD, const unsigned char* qdesc,
              const unsigned char* bad, const unsigned char* in_kf, float th, int* fused) {
    KeyFrame& K = ((KFHolder*)hKF)->K;
    QuerySet Q(nq, xw, normal, maxD, minD, qdesc, bad);
    for (int q = 0; q < nq; ++q)
        if (in_kf && in_kf[q]) Q.mp[q].mObservations[&K] = std::make_tuple(0, -1);
    std::vector<std::pair<int, int> > log;
    K.log_get = &log;
    std::vector<MapPoint*> saved = K.mvpMapPoints;
    ORBmatcher matcher(0.6f, true);
    const int n = matcher.Fuse(&K, Q.ptr, th, false);
    K.log_get = nullptr;
    K.mvpMapPoints = saved;
    for (int q = 0; q < nq; ++q) fused[q] = -1;
    for (size_t i = 0; i < log.size(); ++i) fused[log[i].first] = log[i].second;
    return n;
}

// ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint): S8 = Scw as (non-unit quaternion x y z w, translation, -); already[q] != 0:
// the point is one of pKF->GetMapPoints().  Tcw7 / Ow3: the SE3f and camera centre the function builds, for the oracle's caller-side inputs.
int ref2_fuse_sim3(void* hKF, const float* S8, int nq, const float* xw, const float* normal, const float* maxD, const float* minD,
                   const unsigned char* qdesc, const unsigned char* bad, float th, int* fused, float* Tcw7, float* Ow3) {
    KeyFrame& K = ((KFHolder*)hKF)->K;
    QuerySet Q(nq, xw, normal, maxD, minD, qdesc, bad);
    Sophus::Sim3f Scw;
    set_sim3(Scw, S8);
    export_scw(Scw, Tcw7, Ow3);
    std::vector<std::pair<int, int> > log;
    K.log_get = &log;
    std::vector<MapPoint*> saved = K.mvpMapPoints;
    std::vector<MapPoint*> repl(nq, static_cast<MapPoint*>(NULL));
    ORBmatcher matcher(0.6f, true);
    const int n = matcher.Fuse(&K, Scw, Q.ptr, th, repl);
    K.log_get = nullptr;
    K.mvpMapPoints = saved;
    for (int q = 0; q < nq; ++q) fused[q] = -1;
    for (size_t i = 0; i < log.size(); ++i) fused[log[i].first] = log[i].second;
    return n;
}

// ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (with_kfs == 0) and the overload with
// vpPointsKFs / vpMatchedKF (with_kfs != 0).  matched_in[idx] != 0: vpMatched[idx] is non-NULL on entry.  match[q] = feature or -1.
int ref2_search_kf_sim3(void* hKF, const float* S8, int nq, const float* xw, const float* normal, const float* maxD, const float* minD,
                        const unsigned char* qdesc, const unsigned char* bad, const unsigned char* matched_in, float th, float ratioHamming,
                        int with_kfs, int* match, float* Tcw7, float* Ow3) {
    KeyFrame& K = ((KFHolder*)hKF)->K;
    QuerySet Q(nq, xw, normal, maxD, minD, qdesc, bad);
    Sophus::Sim3f Scw;
    set_sim3(Scw, S8);
    export_scw(Scw, Tcw7, Ow3);
    MapPoint occupied;
    occupied.query_index = -1;
    std::vector<MapPoint*> vpMatched(K.N, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < K.N; ++i)
        if (matched_in && matched_in[i]) vpMatched[i] = &occupied;
    ORBmatcher matcher(0.75f, true);
    int n;
    if (with_kfs) {
        std::vector<KeyFrame*> kfs(nq, &K), mkf(K.N, static_cast<KeyFrame*>(NULL));
        n = matcher.SearchByProjection(&K, Scw, Q.ptr, kfs, vpMatched, mkf, th, ratioHamming);
    } else {
        n = matcher.SearchByProjection(&K, Scw, Q.ptr, vpMatched, th, ratioHamming);
    }
    for (int q = 0; q < nq; ++q) match[q] = -1;
    for (int i = 0; i < K.N; ++i)
        if (vpMatched[i] && vpMatched[i] != &occupied) match[vpMatched[i]->query_index] = i;
    return n;
}

// ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist): the keyframe's map points (ref2_kf_set_mappoints) are
// the queries; already[i] != 0: the point of KF feature i is in sAlreadyFound; claimed[idx]: CurrentFrame.mvpMapPoints[idx] non-NULL on
// entry.  feat_match[idx] = the KF feature whose point the frame feature now holds, or -1.
int ref2_search_frame_kf(void* hF, const float* Tcw7, void* hKF, const unsigned char* already, const unsigned char* claimed, float th, int ORBdist,
                         int checkOri, int* feat_match) {
    Frame& F = ((Holder*)hF)->F;
    KFHolder* KH = (KFHolder*)hKF;
    set_pose(F, Tcw7);
    MapPoint occupied;
    for (int i = 0; i < F.N; ++i) F.mvpMapPoints[i] = (claimed && claimed[i]) ? &occupied : static_cast<MapPoint*>(NULL);
    std::set<MapPoint*> found;
    for (int i = 0; i < KH->K.N; ++i)
        if (already && already[i] && KH->K.mvpMapPoints[i]) found.insert(KH->K.mvpMapPoints[i]);
    ORBmatcher matcher(0.9f, checkOri != 0);
    const int n = matcher.SearchByProjection(F, &KH->K, found, th, ORBdist);
    for (int i = 0; i < F.N; ++i) {
        feat_match[i] = (F.mvpMapPoints[i] && F.mvpMapPoints[i] != &occupied) ? F.mvpMapPoints[i]->query_index : -1;
        F.mvpMapPoints[i] = NULL;
    }
    return n;
}

// ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, S12, th): S21_8 = S12.inverse() as the caller's Sophus computes it (the oracle takes
// both directions from its caller).  matches12[i1] = KF2 feature or -1, in (pre-existing matches) / out (with the new ones).
int ref2_search_by_sim3(void* hKF1, void* hKF2, const float* S12_8, const float* S21_8, float th, int* matches12) {
    KFHolder* H1 = (KFHolder*)hKF1;
    KFHolder* H2 = (KFHolder*)hKF2;
    KeyFrame &K1 = H1->K, &K2 = H2->K;
    Sophus::Sim3f S12, S21;
    set_sim3(S12, S12_8);
    set_sim3(S21, S21_8);
    S12.inv = &S21; S21.inv = &S12;
    for (int i = 0; i < K2.N; ++i) H2->mps[i].mObservations[&K2] = std::make_tuple(i, -1);   // GetIndexInKeyFrame(pKF2)
    std::vector<MapPoint*> vp(K1.N, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < K1.N; ++i)
        if (matches12[i] >= 0) vp[i] = &H2->mps[matches12[i]];
    ORBmatcher matcher(0.75f, true);
    const int n = matcher.SearchBySim3(&K1, &K2, vp, S12, th);
    for (int i = 0; i < K1.N; ++i) matches12[i] = vp[i] ? vp[i]->query_index : -1;
    return n;
}

}  // extern "C"

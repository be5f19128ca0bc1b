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

namespace ORB_SLAM3 {
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv, Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
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

// oracle/ref_shim/ref_capi.cpp -- TEST INFRASTRUCTURE.
//
// C entry points around the REFERENCE's own ORB_SLAM3::ORBextractor, compiled from /root/reference/src/ORBextractor.cc
// as it lies (oracle/Makefile, target _ref/liborb_ref.so) against the miniature cv:: of ref_shim/opencv2.  Used by
// tests/test_oracle_vs_ref.py to pin oracle/extractor_oracle.cpp's restatement (quadtree order, operator() ordering,
// IC_Angle, computeOrbDescriptor, constructor tables) to the reference source, and by bench.py's CPU arm as the
// "reference" kind of baseline for the extractor stages.
#include <chrono>
#include <cstring>

#include "ORBextractor.h"   // the reference's header

namespace {
struct Open : public ORB_SLAM3::ORBextractor {   // the protected members the stage-wise comparisons need
    using ORB_SLAM3::ORBextractor::ORBextractor;
    using ORB_SLAM3::ORBextractor::ComputeKeyPointsOctTree;
    using ORB_SLAM3::ORBextractor::ComputePyramid;
    using ORB_SLAM3::ORBextractor::DistributeOctTree;
    using ORB_SLAM3::ORBextractor::mnFeaturesPerLevel;
    using ORB_SLAM3::ORBextractor::mvInvLevelSigma2;
    using ORB_SLAM3::ORBextractor::mvInvScaleFactor;
    using ORB_SLAM3::ORBextractor::mvLevelSigma2;
    using ORB_SLAM3::ORBextractor::mvScaleFactor;
    using ORB_SLAM3::ORBextractor::umax;
};
static_assert(sizeof(cv::KeyPoint) == 28, "cv::KeyPoint layout");
}  // namespace

extern "C" {

void* ref_extractor_create(int nf, float sf, int nl, int ini, int mn) { return new Open(nf, sf, nl, ini, mn); }
void ref_extractor_destroy(void* h) { delete (Open*)h; }

// ORBextractor::operator() (ORBextractor.cc:1557-1682).  Returns its return value; *n_out = keypoints written.
int ref_extract(void* h, const unsigned char* img, int w, int hh, int stride, int lap0, int lap1, void* kps_out,
                unsigned char* desc_out, int cap, int* n_out) {
    Open* e = (Open*)h;
    cv::Mat image(hh, w, CV_8UC1, (void*)img, (size_t)stride), desc;
    std::vector<cv::KeyPoint> kps;
    std::vector<int> lap = {lap0, lap1};
    const int mono = (*e)(image, cv::Mat(), kps, desc, lap);
    const int n = (int)kps.size();
    *n_out = n;
    if (n > cap) return -100;
    if (n) std::memcpy(kps_out, kps.data(), (size_t)n * 28);
    for (int i = 0; i < n; ++i) std::memcpy(desc_out + (size_t)i * 32, desc.ptr(i), 32);
    return mono;
}

void ref_tables(void* h, float* sc, float* isc, float* s2, float* is2, int* quota, int* um) {
    Open* e = (Open*)h;
    const int L = e->GetLevels();
    for (int l = 0; l < L; ++l) {
        sc[l] = e->mvScaleFactor[l]; isc[l] = e->mvInvScaleFactor[l]; s2[l] = e->mvLevelSigma2[l]; is2[l] = e->mvInvLevelSigma2[l];
        quota[l] = e->mnFeaturesPerLevel[l];
    }
    for (int i = 0; i < 16; ++i) um[i] = e->umax[i];
}

void ref_level_size(void* h, int l, int* w, int* hh) {
    Open* e = (Open*)h;
    *w = e->mvImagePyramid[l].cols; *hh = e->mvImagePyramid[l].rows;
}
void ref_level_pyramid(void* h, int l, unsigned char* out) {
    Open* e = (Open*)h;
    const cv::Mat& m = e->mvImagePyramid[l];
    for (int r = 0; r < m.rows; ++r) std::memcpy(out + (size_t)r * m.cols, m.ptr(r), (size_t)m.cols);
}

// ComputePyramid + ComputeKeyPointsOctTree alone (ORBextractor.cc:1061-1208): per-level keypoints in level coordinates
// with angles, before descriptors and rescaling.  level_offsets has nlevels+1 entries.  Returns the total or -100.
int ref_keypoints_octtree(void* h, const unsigned char* img, int w, int hh, int stride, void* kps_out, int cap, int* level_offsets) {
    Open* e = (Open*)h;
    cv::Mat image(hh, w, CV_8UC1, (void*)img, (size_t)stride);
    e->ComputePyramid(image);
    std::vector<std::vector<cv::KeyPoint>> all;
    e->ComputeKeyPointsOctTree(all);
    int n = 0;
    for (size_t l = 0; l < all.size(); ++l) {
        level_offsets[l] = n;
        if (n + (int)all[l].size() > cap) return -100;
        if (!all[l].empty()) std::memcpy((char*)kps_out + (size_t)n * 28, all[l].data(), all[l].size() * 28);
        n += (int)all[l].size();
    }
    level_offsets[all.size()] = n;
    return n;
}

// DistributeOctTree (ORBextractor.cc:711-1057) on an arbitrary candidate set: cand = n x (x, y, response) int32, in the
// order they enter vToDistributeKeys.  out_xy_resp = kept keypoints (x, y, response) int32 in the returned order.
int ref_distribute(void* h, const int* cand, int n, int minX, int maxX, int minY, int maxY, int N, int level, int* out, int cap) {
    Open* e = (Open*)h;
    std::vector<cv::KeyPoint> v(n);
    for (int i = 0; i < n; ++i) {
        v[i] = cv::KeyPoint((float)cand[3 * i], (float)cand[3 * i + 1], 7.f, -1.f, (float)cand[3 * i + 2]);
        v[i].class_id = i;   // carried through so the caller sees WHICH duplicate survived
    }
    std::vector<cv::KeyPoint> r = e->DistributeOctTree(v, minX, maxX, minY, maxY, N, level);
    if ((int)r.size() > cap) return -100;
    for (size_t i = 0; i < r.size(); ++i) {
        out[4 * i] = (int)r[i].pt.x; out[4 * i + 1] = (int)r[i].pt.y; out[4 * i + 2] = (int)r[i].response; out[4 * i + 3] = r[i].class_id;
    }
    return (int)r.size();
}

double ref_now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // extern "C"

// ORBextractor_b200.cc -- drop-in replacement for /root/reference/src/ORBextractor.cc.
//
// Compiled AGAINST THE REFERENCE'S OWN HEADER (include/ORBextractor.h, unchanged) and linked with
// liborbslam3_b200.so instead of the reference's ORBextractor.cc, it gives Tracking / Frame the same
// ORB_SLAM3::ORBextractor class: same constructor, same operator(), same getters, mvImagePyramid populated.
// All arithmetic happens on the B200 behind the C ABI (include/orbslam3_b200.h); this file only marshals.
//
//   g++ -std=c++14 -I<ORB_SLAM3>/include -I<this repo>/include $(pkg-config --cflags opencv4) \
//       -c ORBextractor_b200.cc            # then link -lorbslam3_b200 in place of ORBextractor.o
//
// The reference class has no spare member and an inline empty destructor, so the device handle lives in a
// side table keyed by `this` (three extractors per session: left, right, ini -- Tracking.cc:629-635).
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "ORBextractor.h"      // the reference's header
#include "orbslam3_b200.h"

#ifndef ORB_B200_MAX_WIDTH
#define ORB_B200_MAX_WIDTH 1280
#endif
#ifndef ORB_B200_MAX_HEIGHT
#define ORB_B200_MAX_HEIGHT 1024
#endif
#ifndef ORB_B200_MATERIALIZE_PYRAMID
#define ORB_B200_MATERIALIZE_PYRAMID 1   // Frame::ComputeStereoMatches reads mvImagePyramid (Frame.cc:1249,1275)
#endif

namespace {
std::mutex g_mu;
std::unordered_map<const ORB_SLAM3::ORBextractor*, orbx_handle*> g_handles;

orbx_handle* handle_of(const ORB_SLAM3::ORBextractor* self) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_handles.find(self);
    return it == g_handles.end() ? nullptr : it->second;
}

[[noreturn]] void die(const char* what) {
    std::fprintf(stderr, "ORBextractor (B200): %s: %s\n", what, orb_last_error());
    std::abort();   // the reference has no error channel here either; there is no CPU fallback
}
}  // namespace

namespace ORB_SLAM3 {

// ORBextractor.cc:468-571
ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
    orbx_config cfg;
    cfg.n_features = _nfeatures;
    cfg.scale_factor = _scaleFactor;
    cfg.n_levels = _nlevels;
    cfg.ini_th_fast = _iniThFAST;
    cfg.min_th_fast = _minThFAST;
    cfg.max_width = ORB_B200_MAX_WIDTH;
    cfg.max_height = ORB_B200_MAX_HEIGHT;
    cfg.max_batch = 1;
    cfg.device = 0;
    if (const char* d = std::getenv("ORB_B200_DEVICE")) cfg.device = std::atoi(d);
    orbx_handle* h = nullptr;
    if (orbx_create(&cfg, &h) != ORB_OK) die("orbx_create");
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_handles[this] = h;
    }
    mvScaleFactor.resize(nlevels);
    mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels);
    mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    umax.resize(16);
    orbx_get_tables(h, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(),
                    mnFeaturesPerLevel.data(), umax.data());
    mvImagePyramid.resize(nlevels);
}

// ORBextractor.cc:1557-1682
int ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
                             cv::OutputArray _descriptors, std::vector<int>& vLappingArea) {
    if (_image.empty()) return -1;
    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1);
    orbx_handle* h = handle_of(this);
    static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "cv::KeyPoint layout");
    const int cap = 4 * nfeatures + 16 * nlevels;
    std::vector<cv::KeyPoint> kps(cap);
    std::vector<unsigned char> desc((size_t)cap * 32);
    int n = 0, mono = 0;
    const int lap0 = vLappingArea.size() > 0 ? vLappingArea[0] : 0, lap1 = vLappingArea.size() > 1 ? vLappingArea[1] : 0;
    if (orbx_extract(h, image.data, image.cols, image.rows, (int)image.step, lap0, lap1,
                     reinterpret_cast<orbx_keypoint*>(kps.data()), desc.data(), cap, &n, &mono) != ORB_OK)
        die("orbx_extract");
    kps.resize(n);
    _keypoints.swap(kps);
    if (n == 0) {
        _descriptors.release();
    } else {
        _descriptors.create(n, 32, CV_8U);
        cv::Mat d = _descriptors.getMat();
        for (int i = 0; i < n; ++i) std::memcpy(d.ptr(i), &desc[(size_t)i * 32], 32);
    }
#if ORB_B200_MATERIALIZE_PYRAMID
    for (int l = 0; l < nlevels; ++l) {
        int w = 0, hh = 0;
        orbx_level_size(h, l, &w, &hh);
        mvImagePyramid[l].create(hh, w, CV_8UC1);
        if (orbx_download_level(h, 0, l, 0, mvImagePyramid[l].data, (int)mvImagePyramid[l].step) != ORB_OK) die("orbx_download_level");
    }
#endif
    return mono;
}

}  // namespace ORB_SLAM3

// For Frame::ComputeStereoMatches (Frame.cc:1102-1358) the replacement body is three lines once both extractors
// are ours -- see INTEGRATION.md:
//   orbm_stereo_pair(handle(mpORBextractorLeft), handle(mpORBextractorRight), mbf, mb, mvuRight.data(), mvDepth.data(), N);
extern "C" orbx_handle* orb_b200_handle_of(const void* extractor) {
    return handle_of(static_cast<const ORB_SLAM3::ORBextractor*>(extractor));
}

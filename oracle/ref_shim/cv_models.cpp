// oracle/ref_shim/cv_models.cpp -- TEST INFRASTRUCTURE.
//
// The five OpenCV routines src/ORBextractor.cc calls, over the miniature cv:: types of ref_shim/opencv2/, each one
// forwarding to the pixel-arithmetic model in oracle/extractor_oracle.cpp that tests/test_oracle_vs_cv2.py pins
// against the real cv2 4.13 build (call sites: ORBextractor.cc:1702 resize, :1712/:1734 copyMakeBorder,
// :1135/:1144 FAST, :1632 GaussianBlur, :137 fastAtan2).
#include "opencv2/core/core.hpp"

#include "../oracle.h"

namespace cv {

static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

void resize(InputArray src_, OutputArray dst_, Size dsize, double fx, double fy, int interpolation) {
    assert(interpolation == INTER_LINEAR && fx == 0 && fy == 0);
    (void)fx; (void)fy; (void)interpolation;
    Mat src = src_.getMat();
    dst_.create(dsize, src.type());      // a destination of the right shape is written in place (here: a ROI of `temp`)
    Mat dst = dst_.getMat();
    orb_oracle::resize_linear_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType, const Scalar&) {
    assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    (void)borderType;
    Mat src = src_.getMat();
    dst_.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = dst_.getMat();
    // interior first (src may BE the interior of dst: ORBextractor.cc:1712), then the reflected frame from the interior
    for (int r = 0; r < src.rows; ++r) {
        uchar* d = dst.ptr(r + top) + left;
        if (d != src.ptr(r)) std::memmove(d, src.ptr(r), (size_t)src.cols);
    }
    const int w = src.cols, h = src.rows;
    for (int r = 0; r < h; ++r) {
        uchar* row = dst.ptr(r + top);
        for (int x = 0; x < left; ++x) row[x] = row[left + reflect101(x - left, w)];
        for (int x = 0; x < right; ++x) row[left + w + x] = row[left + reflect101(w + x, w)];
    }
    for (int r = 0; r < top; ++r) std::memcpy(dst.ptr(r), dst.ptr(top + reflect101(r - top, h)), (size_t)dst.cols);
    for (int r = 0; r < bottom; ++r) std::memcpy(dst.ptr(top + h + r), dst.ptr(top + reflect101(h + r, h)), (size_t)dst.cols);
}

void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sigmaX, double sigmaY, int borderType) {
    assert(ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2 && borderType == BORDER_REFLECT_101);
    (void)ksize; (void)sigmaX; (void)sigmaY; (void)borderType;
    Mat src = src_.getMat();
    Mat out(src.rows, src.cols, src.type());
    orb_oracle::gaussian_blur7_u8(src.data, src.cols, src.rows, (int)src.step, out.data, (int)out.step);
    dst_.create(src.rows, src.cols, src.type());
    Mat dst = dst_.getMat();
    for (int r = 0; r < src.rows; ++r) std::memcpy(dst.ptr(r), out.ptr(r), (size_t)src.cols);
}

void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
    assert(nonmaxSuppression);
    (void)nonmaxSuppression;
    Mat im = image.getMat();
    std::vector<orb_oracle::Cand> c;
    orb_oracle::fast_cell(im.data, im.cols, im.rows, (int)im.step, threshold, c);
    keypoints.clear();
    keypoints.reserve(c.size());
    for (const auto& k : c) keypoints.push_back(KeyPoint((float)k.x, (float)k.y, 7.f, -1.f, (float)k.score));
}

float fastAtan2(float y, float x) { return orb_oracle::fast_atan2_deg(y, x); }

void KeyPointsFilter::retainBest(std::vector<KeyPoint>& kps, int n) {
    if (n >= 0 && (int)kps.size() > n) {
        if (n == 0) { kps.clear(); return; }
        std::nth_element(kps.begin(), kps.begin() + n - 1, kps.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
        const float ambiguous = kps[n - 1].response;
        auto e = std::partition(kps.begin() + n, kps.end(), [ambiguous](const KeyPoint& k) { return k.response >= ambiguous; });
        kps.resize(e - kps.begin());
    }
}

}  // namespace cv

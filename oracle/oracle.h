// oracle/oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement ("port") of the ORB-SLAM3 per-frame hot path.  Only tests/, bench.py's
// cpu_baseline / --impl reference leg and __graft_entry__.smoke() may load this library; the product
// library (liborbslam3_b200.so) never links, loads or calls it.
//
// Parity pin status: the reference (electech6/ORB_SLAM3_detailed_comments) ships no tests and cannot be built as a whole in this
// image (no OpenCV C++ / Eigen).  Pins: (1) the OpenCV-backed stages against the real cv2 4.13 build (tests/test_oracle_vs_cv2.py)
// and glibc sinf / cosf / logf exhaustively; (2) the reference's OWN SOURCE where it compiles from its files (oracle/_ref, DESIGN.md
// section 2): ORBextractor.cc unmodified; every function of ORBmatcher.cc, the Frame / KeyFrame grid, isInFrustum,
// ComputeStereoMatches, the MapPoint routines and Pinhole's projection / epipolar test cut out at build time and compiled verbatim
// over skeleton classes; Thirdparty/DBoW2 unmodified on the reference's ORBvoc.txt; g2o's Levenberg control flow and Huber kernel
// verbatim, driving this oracle's bundle-adjustment engine.  (3) The Eigen-expressed numerics under the optimisers (edge errors and
// Jacobians, Schur complement, LDL^T) stay "parity unpinned by the reference source": their pins are finite differences, the
// objective restated in numpy, stationarity and planted-solution recovery.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace orb_oracle {

// cv::KeyPoint layout (28 bytes) -- what ORBextractor::operator() hands back.
struct KeyPoint {
    float x, y;
    float size;
    float angle;
    float response;
    int32_t octave;
    int32_t class_id;
};

struct Plane {  // continuous u8 image (a pyramid ROI without its border)
    int w = 0, h = 0;
    std::vector<uint8_t> px;
    uint8_t at(int y, int x) const { return px[(size_t)y * w + x]; }
};

// A FAST candidate as it enters DistributeOctTree: integer pixel coords relative to (minBorderX,
// minBorderY) and the integer FAST score.
struct Cand {
    int x, y, score;
};

class Extractor {
   public:
    // /root/reference/src/ORBextractor.cc:468-571
    Extractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);

    // /root/reference/src/ORBextractor.cc:1557-1682.  Returns monoIndex, or -1 on empty input,
    // or -2 for geometry the reference itself cannot process (nIni==0, level smaller than a cell).
    int extract(const uint8_t* img, int w, int h, int stride, int lap0, int lap1,
                std::vector<KeyPoint>& kps, std::vector<uint8_t>& desc);

    int nfeatures, nlevels, iniThFAST, minThFAST;
    float scaleFactor;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<int> umax;

    // stage outputs kept for stage-wise parity tests
    std::vector<Plane> pyramid;             // mvImagePyramid[l] (ROI)
    std::vector<Plane> blurred;             // GaussianBlur(clone) per level
    std::vector<std::vector<Cand>> cands;   // vToDistributeKeys per level
    std::vector<std::vector<KeyPoint>> lvl; // per-level keypoints after quadtree + angle (level coords)
    double t_pyr = 0, t_fast = 0, t_tree = 0, t_angle = 0, t_blur = 0, t_desc = 0;  // seconds, last call

    void compute_pyramid(const uint8_t* img, int w, int h, int stride);
    void detect_level(int level, std::vector<Cand>& out) const;
    std::vector<int> distribute(const std::vector<Cand>& c, int minX, int maxX, int minY, int maxY,
                                int N) const;  // indices into c, in output order
};

// primitives (pinned against cv2 / glibc in the tests)
void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh,
                      int dstride);
void gaussian_blur7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);
int fast_score16(const uint8_t* p, int stride);
int fast_score16_scalar(const uint8_t* p, int stride);   // the same value without SIMD (tests)  // max T such that p is a FAST-9/16 corner, i.e. cv response
void fast_cell(const uint8_t* win, int cw, int ch, int stride, int thr, std::vector<Cand>& out);
float fast_atan2_deg(float y, float x);
void set_fast_simd(int on);   // 1 (default): level-wide SSE2 corner test; 0: cell-by-cell scalar path (tests compare the two)
float ic_angle(const Plane& im, int x, int y, const std::vector<int>& umax);
void orb_descriptor(const Plane& blurred, int x, int y, float angle_deg, uint8_t* desc32);
int descriptor_distance(const uint8_t* a, const uint8_t* b);
extern const int8_t kPattern[1024];

}  // namespace orb_oracle

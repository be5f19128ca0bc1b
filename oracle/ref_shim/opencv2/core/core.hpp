// oracle/ref_shim/opencv2/core/core.hpp -- TEST INFRASTRUCTURE.
//
// A miniature stand-in for the OpenCV C++ API, written from scratch, holding exactly what the reference's
// UNMODIFIED src/ORBextractor.cc + include/ORBextractor.h need in order to compile in an image that has no OpenCV
// C++ headers (cv2 exists here as a Python wheel only).  oracle/Makefile's `_ref` target compiles the reference
// file where it lies under /root/reference against this directory; nothing of the reference is copied.
//
// The arithmetic behind resize / GaussianBlur / FAST / fastAtan2 / cvRound is NOT OpenCV's code: it is the set of
// integer / float32 models in oracle/extractor_oracle.cpp that tests/test_oracle_vs_cv2.py pins against the real
// cv2 4.13 build.  So oracle/_ref = (reference control flow, verbatim) x (cv2-pinned pixel arithmetic).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <sstream>     // the real core.hpp pulls these in; Thirdparty/DBoW2's TemplatedVocabulary.h relies on it
#include <stdexcept>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_PI 3.1415926535897932384626433832795

typedef unsigned char uchar;

static inline int cvRound(double v) { return (int)lrint(v); }   // round-half-even, as the SSE2 cvtsd2si path
static inline int cvRound(float v) { return (int)lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }

namespace cv {

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U>
    Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }   // saturate_cast<float> of a float product
    Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

template <typename T>
struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <typename T>
struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;

struct Range {
    int start, end;
    Range(int s, int e) : start(s), end(e) {}
};

struct KeyPoint {   // 28 bytes, the layout ORB-SLAM3 hands around
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f p, float sz, float ang = -1, float resp = 0, int oct = 0, int cid = -1)
        : pt(p), size(sz), angle(ang), response(resp), octave(oct), class_id(cid) {}
    KeyPoint(float x, float y, float sz, float ang = -1, float resp = 0, int oct = 0, int cid = -1)
        : pt(x, y), size(sz), angle(ang), response(resp), octave(oct), class_id(cid) {}
};

// u8 single-channel matrix header over a shared buffer: sub-matrices alias their parent like cv::Mat does
// (ORBextractor::ComputePyramid relies on that: the level image is a ROI of a bordered temporary).
class Mat {
   public:
    int rows = 0, cols = 0;
    uchar* data = nullptr;
    size_t step = 0;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(int r, int c, int /*type*/, void* ext, size_t stp = 0) : rows(r), cols(c), data((uchar*)ext), step(stp ? stp : (size_t)c) {}

    void create(int r, int c, int type) {
        if (data && r == rows && c == cols) return;   // cv::Mat::create keeps a buffer of the right shape
        const size_t es = type == CV_32F ? 4 : 1;     // CV_32F: only DBoW2's FORB::toMat32F (never called) asks for it
        rows = r; cols = c; step = (size_t)c * es;
        store_.reset(new std::vector<uchar>((size_t)r * c * es));
        data = store_->data();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() { rows = cols = 0; data = nullptr; step = 0; store_.reset(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    int channels() const { return 1; }
    Size size() const { return Size(cols, rows); }
    size_t step1() const { return step; }
    bool isContinuous() const { return step == (size_t)cols || rows == 1; }

    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    template <typename T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
    template <typename T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }

    Mat operator()(const Rect& roi) const {
        assert(roi.x >= 0 && roi.y >= 0 && roi.x + roi.width <= cols && roi.y + roi.height <= rows);
        Mat m;
        m.rows = roi.height; m.cols = roi.width; m.step = step; m.data = data + (size_t)roi.y * step + roi.x; m.store_ = store_;
        return m;
    }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
    Mat rowRange(const Range& r) const { return rowRange(r.start, r.end); }
    Mat colRange(const Range& r) const { return colRange(r.start, r.end); }
    Mat row(int r) const { return rowRange(r, r + 1); }
    Mat col(int c) const { return colRange(c, c + 1); }

    Mat clone() const {
        Mat m;
        copyTo(m);
        return m;
    }
    void copyTo(Mat& dst) const {
        dst.create(rows, cols, CV_8UC1);
        for (int r = 0; r < rows; ++r) std::memmove(dst.ptr(r), ptr(r), (size_t)cols);
    }
    void copyTo(Mat&& dst) const { copyTo(dst); }   // `desc.row(i).copyTo(descriptors.row(k))`: the temporary aliases the parent
    static Mat zeros(int r, int c, int type) {
        Mat m(r, c, type);
        if (m.data) std::memset(m.data, 0, (size_t)r * c);
        return m;
    }
    static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }

   private:
    std::shared_ptr<std::vector<uchar>> store_;
};

// InputArray / OutputArray reduced to "a reference to a Mat"
class _InputArray {
   public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    Mat getMat() const { return m_ ? *m_ : Mat(); }
    bool empty() const { return !m_ || m_->empty(); }

   protected:
    Mat* m_;
};
class _OutputArray : public _InputArray {
   public:
    _OutputArray() {}
    _OutputArray(Mat& m) { m_ = &m; }
    void create(int r, int c, int t) const { if (m_) m_->create(r, c, t); }
    void create(Size s, int t) const { if (m_) m_->create(s, t); }
    void release() const { if (m_) m_->release(); }
    Mat& getMatRef() const { return *m_; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
static inline InputArray noArray() { static _InputArray none; return none; }

enum BorderTypes { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
                   BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };

struct Scalar {
    double v[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {}
};

enum NormTypes { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6 };
// cv::norm(a, b, NORM_L1) of two u8 matrices of the same size (Frame::ComputeStereoMatches' SAD): an exact integer sum
static inline double norm(const Mat& a, const Mat& b, int normType) {
    assert(normType == NORM_L1 && a.rows == b.rows && a.cols == b.cols);
    (void)normType;
    long s = 0;
    for (int r = 0; r < a.rows; ++r) {
        const uchar *pa = a.ptr(r), *pb = b.ptr(r);
        for (int c = 0; c < a.cols; ++c) s += pa[c] > pb[c] ? pa[c] - pb[c] : pb[c] - pa[c];
    }
    return (double)s;
}

// --- the five OpenCV routines ORBextractor.cc calls (ref_shim/cv_models.cpp) ---
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType,
                    const Scalar& value = Scalar());
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
float fastAtan2(float y, float x);

// cv::FileStorage / cv::FileNode: Thirdparty/DBoW2's TemplatedVocabulary has virtual save / load members over them (YAML
// vocabularies).  ORB-SLAM3 loads its vocabulary with loadFromTextFile, so these only have to compile; using them throws.
class FileNode {
   public:
    FileNode operator[](const char*) const { throw std::runtime_error("cv::FileNode is not modelled"); }
    FileNode operator[](const std::string&) const { throw std::runtime_error("cv::FileNode is not modelled"); }
    FileNode operator[](int) const { throw std::runtime_error("cv::FileNode is not modelled"); }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator double() const { return 0.0; }
    operator std::string() const { return std::string(); }
};
class FileStorage {
   public:
    enum { READ = 0, WRITE = 1 };
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    FileNode operator[](const std::string&) const { throw std::runtime_error("cv::FileStorage is not modelled"); }
    FileNode operator[](const char*) const { throw std::runtime_error("cv::FileStorage is not modelled"); }
};
template <typename T>
static inline FileStorage& operator<<(FileStorage& fs, const T&) { throw std::runtime_error("cv::FileStorage is not modelled"); return fs; }

struct KeyPointsFilter {   // only ComputeKeyPointsOld (dead code in the reference) uses it
    static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);
};

}  // namespace cv

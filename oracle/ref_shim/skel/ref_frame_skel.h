// oracle/ref_shim/skel/ref_frame_skel.h -- TEST INFRASTRUCTURE.
//
// Skeleton of the reference's Frame / MapPoint / camera classes, holding exactly the members that the member functions
// oracle/Makefile extracts from the reference sources at build time touch:
//     ORBmatcher.cc : ORBmatcher::ORBmatcher, SearchByProjection(Frame&, vector<MapPoint*>&, ...), SearchByProjection(Frame&, const
//                     Frame&, th, bMono), RadiusByViewingCos, ComputeThreeMaxima, DescriptorDistance, TH_LOW / TH_HIGH / HISTO_LENGTH
//     Frame.cc      : AssignFeaturesToGrid, PosInGrid, GetFeaturesInArea, isInFrustum(MapPoint*, float), ComputeStereoMatches
//                     SearchByBoW(KeyFrame*, Frame&, ...), SearchByBoW(KeyFrame*, KeyFrame*, ...), SearchForInitialization,
//                     SearchForTriangulation, Fuse (both overloads), SearchByProjection(KeyFrame*, Sim3f&, ...) (both overloads),
//                     SearchByProjection(Frame&, KeyFrame*, set<MapPoint*>&, ...), SearchBySim3
//     KeyFrame.cc   : GetFeaturesInArea, IsInImage
//     MapPoint.cc   : PredictScale(const float&, Frame*), PredictScale(const float&, KeyFrame*), ComputeDistinctiveDescriptors(), UpdateNormalAndDepth()
//     Pinhole.cpp   : project(const Eigen::Vector3f&), toK_(), epipolarConstrain(...)
// Those functions are compiled VERBATIM (oracle/tools/extract_functions.py writes them into oracle/_ref/gen/, a build directory)
// against the reference's own include/ORBmatcher.h and include/ORBextractor.h; this header defines the include guards of
// Frame.h / KeyFrame.h / MapPoint.h (which pull in DBoW2, g2o, boost, Pangolin: none exist in this image) and supplies the class
// members with the reference's names and types.  So oracle/_ref/liborb_ref2.so = (reference control flow of the matchers, the
// grid, the stereo matcher and isInFrustum, verbatim) x (the arithmetic models below).
//
// Arithmetic that is NOT the reference's own code and therefore pins nothing: the miniature Eigen (fixed-size float vectors /
// 3x3 matrix; matrix * vector and dot / norm in the summation order DESIGN.md section 2 states for Eigen's unrolled kernels) and the
// miniature Sophus::SE3f (quaternion action as Thirdparty/Sophus so3.hpp:358-367) -- the same orders the oracle restates.
#pragma once
#define FRAME_H
#define KEYFRAME_H
#define MAPPOINT_H

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <tuple>
#include <vector>

#include <opencv2/core/core.hpp>   // oracle/ref_shim/opencv2: the miniature cv::
#include "ORBextractor.h"          // the reference's own header
#include "DBoW2/BowVector.h"       // the reference's own Thirdparty/DBoW2 headers (std::map subclasses); boost stubs: ref_shim/dbow
#include "DBoW2/FeatureVector.h"

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {
template <typename T, int R, int C>
struct Matrix {
    T v[R * C];
    Matrix() { for (int i = 0; i < R * C; ++i) v[i] = T(0); }
    Matrix(T a, T b) { static_assert(R * C == 2, "2-vector"); v[0] = a; v[1] = b; }
    Matrix(T a, T b, T c) { static_assert(R * C == 3, "3-vector"); v[0] = a; v[1] = b; v[2] = c; }
    T& operator()(int i) { return v[i]; }
    const T& operator()(int i) const { return v[i]; }
    T& operator[](int i) { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
    T& operator()(int r, int c) { return v[r * C + c]; }
    const T& operator()(int r, int c) const { return v[r * C + c]; }
    Matrix operator+(const Matrix& b) const { Matrix o; for (int i = 0; i < R * C; ++i) o.v[i] = v[i] + b.v[i]; return o; }
    Matrix operator-(const Matrix& b) const { Matrix o; for (int i = 0; i < R * C; ++i) o.v[i] = v[i] - b.v[i]; return o; }
    Matrix operator-() const { Matrix o; for (int i = 0; i < R * C; ++i) o.v[i] = -v[i]; return o; }
    Matrix operator/(T s) const { Matrix o; for (int i = 0; i < R * C; ++i) o.v[i] = v[i] / s; return o; }   // scalar_quotient_op: a division per coefficient
    void setZero() { for (int i = 0; i < R * C; ++i) v[i] = T(0); }
    // 3-term sums in the order x + (y + z) (Eigen 3.3+'s unrolled fixed-size reduction)
    T dot(const Matrix& b) const { static_assert(R * C == 3, "3-vector"); return v[0] * b.v[0] + (v[1] * b.v[1] + v[2] * b.v[2]); }
    T norm() const { return std::sqrt(dot(*this)); }
    Matrix<T, R, 1> operator*(const Matrix<T, C, 1>& b) const {
        static_assert(R == 3 && C == 3, "3x3 * 3x1");
        Matrix<T, R, 1> o;
        for (int r = 0; r < 3; ++r) o.v[r] = (*this)(r, 0) * b.v[0] + ((*this)(r, 1) * b.v[1] + (*this)(r, 2) * b.v[2]);
        return o;
    }
    // --- only Pinhole::epipolarConstrain's F12 = K1^-T [t]x R12 K2^-1 needs what follows; its value is handed to the oracle as an
    // input by the test (the oracle takes F12 from its caller), so these models decide nothing that is compared
    template <int C2>
    Matrix<T, R, C2> operator*(const Matrix<T, C, C2>& b) const {
        static_assert(C2 > 1, "matrix * matrix");
        Matrix<T, R, C2> o;
        for (int r = 0; r < R; ++r)
            for (int c = 0; c < C2; ++c) {
                T a = T(0);
                for (int k = 0; k < C; ++k) a += (*this)(r, k) * b(k, c);
                o(r, c) = a;
            }
        return o;
    }
    Matrix<T, C, R> transpose() const {
        Matrix<T, C, R> o;
        for (int r = 0; r < R; ++r)
            for (int c = 0; c < C; ++c) o(c, r) = (*this)(r, c);
        return o;
    }
    Matrix inverse() const {   // cofactors / determinant
        static_assert(R == 3 && C == 3, "3x3");
        const Matrix& m = *this;
        Matrix o;
        const T det = m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0)) +
                      m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
        const T id = T(1) / det;
        o(0, 0) = (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) * id; o(0, 1) = (m(0, 2) * m(2, 1) - m(0, 1) * m(2, 2)) * id; o(0, 2) = (m(0, 1) * m(1, 2) - m(0, 2) * m(1, 1)) * id;
        o(1, 0) = (m(1, 2) * m(2, 0) - m(1, 0) * m(2, 2)) * id; o(1, 1) = (m(0, 0) * m(2, 2) - m(0, 2) * m(2, 0)) * id; o(1, 2) = (m(0, 2) * m(1, 0) - m(0, 0) * m(1, 2)) * id;
        o(2, 0) = (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0)) * id; o(2, 1) = (m(0, 1) * m(2, 0) - m(0, 0) * m(2, 1)) * id; o(2, 2) = (m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0)) * id;
        return o;
    }
    struct CommaInit {   // `K << a, b, c, ...;` row-major
        Matrix* m; int i;
        CommaInit& operator,(T x) { m->v[i++] = x; return *this; }
    };
    CommaInit operator<<(T x) { v[0] = x; return CommaInit{this, 1}; }
};
typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 3, 3> Matrix3f;
}  // namespace Eigen

namespace Sophus {
template <typename T>
class SE3 {   // unit quaternion (x y z w) + translation; point action as so3.hpp:358-367 (Eigen's _transformVector), then + t
   public:
    T q[4], t[3];
    SE3() : q{0, 0, 0, 1}, t{0, 0, 0} {}
    // SE3(rotation matrix, translation): Eigen's matrix -> quaternion conversion (Quaternion.h, quaternionbase_assign_impl<..., 3, 3>);
    // the test hands the resulting quaternion to the oracle, whose caller would have built it with the real Eigen
    SE3(const Eigen::Matrix<T, 3, 3>& m, const Eigen::Matrix<T, 3, 1>& tr) {
        T tq = m(0, 0) + m(1, 1) + m(2, 2);
        if (tq > T(0)) {
            tq = std::sqrt(tq + T(1.0));
            q[3] = T(0.5) * tq;
            tq = T(0.5) / tq;
            q[0] = (m(2, 1) - m(1, 2)) * tq; q[1] = (m(0, 2) - m(2, 0)) * tq; q[2] = (m(1, 0) - m(0, 1)) * tq;
        } else {
            int i = 0;
            if (m(1, 1) > m(0, 0)) i = 1;
            if (m(2, 2) > m(i, i)) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            tq = std::sqrt(m(i, i) - m(j, j) - m(k, k) + T(1.0));
            q[i] = T(0.5) * tq;
            tq = T(0.5) / tq;
            q[3] = (m(k, j) - m(j, k)) * tq; q[j] = (m(j, i) + m(i, j)) * tq; q[k] = (m(k, i) + m(i, k)) * tq;
        }
        const T n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);   // Sophus normalises in SO3(Matrix) -> setQuaternion
        for (int a = 0; a < 4; ++a) q[a] /= n;
        for (int a = 0; a < 3; ++a) t[a] = tr(a);
    }
    Eigen::Matrix<T, 3, 1> translation() const { return Eigen::Matrix<T, 3, 1>(t[0], t[1], t[2]); }
    static void rot(const T* q, const T* p, T* o) {
        const T uvx = q[1] * p[2] - q[2] * p[1], uvy = q[2] * p[0] - q[0] * p[2], uvz = q[0] * p[1] - q[1] * p[0];
        const T ux = uvx + uvx, uy = uvy + uvy, uz = uvz + uvz;
        const T cx = q[1] * uz - q[2] * uy, cy = q[2] * ux - q[0] * uz, cz = q[0] * uy - q[1] * ux;
        o[0] = (p[0] + q[3] * ux) + cx; o[1] = (p[1] + q[3] * uy) + cy; o[2] = (p[2] + q[3] * uz) + cz;
    }
    Eigen::Matrix<T, 3, 1> operator*(const Eigen::Matrix<T, 3, 1>& p) const {
        T o[3];
        rot(q, p.v, o);
        return Eigen::Matrix<T, 3, 1>(o[0] + t[0], o[1] + t[1], o[2] + t[2]);
    }
    Eigen::Matrix<T, 3, 3> rotationMatrix() const {   // Eigen::Quaternion::toRotationMatrix (values handed to the oracle by the test)
        Eigen::Matrix<T, 3, 3> m;
        const T tx = T(2) * q[0], ty = T(2) * q[1], tz = T(2) * q[2];
        const T twx = tx * q[3], twy = ty * q[3], twz = tz * q[3], txx = tx * q[0], txy = ty * q[0], txz = tz * q[0], tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
        m(0, 0) = T(1) - (tyy + tzz); m(0, 1) = txy - twz; m(0, 2) = txz + twy;
        m(1, 0) = txy + twz; m(1, 1) = T(1) - (txx + tzz); m(1, 2) = tyz - twx;
        m(2, 0) = txz - twy; m(2, 1) = tyz + twx; m(2, 2) = T(1) - (txx + tyy);
        return m;
    }
    SE3 operator*(const SE3& b) const {   // se3.hpp: rotation = q * b.q (renormalised by Sophus; inputs here are unit), t = t + q * b.t
        SE3 r;
        r.q[3] = q[3] * b.q[3] - q[0] * b.q[0] - q[1] * b.q[1] - q[2] * b.q[2];
        r.q[0] = q[3] * b.q[0] + q[0] * b.q[3] + q[1] * b.q[2] - q[2] * b.q[1];
        r.q[1] = q[3] * b.q[1] + q[1] * b.q[3] + q[2] * b.q[0] - q[0] * b.q[2];
        r.q[2] = q[3] * b.q[2] + q[2] * b.q[3] + q[0] * b.q[1] - q[1] * b.q[0];
        T o[3];
        rot(q, b.t, o);
        r.t[0] = o[0] + t[0]; r.t[1] = o[1] + t[1]; r.t[2] = o[2] + t[2];
        return r;
    }
    SE3 inverse() const {   // se3.hpp: SE3(so3().inverse(), so3().inverse() * (translation() * -1))
        SE3 r;
        r.q[0] = -q[0]; r.q[1] = -q[1]; r.q[2] = -q[2]; r.q[3] = q[3];
        const T nt[3] = {t[0] * T(-1), t[1] * T(-1), t[2] * T(-1)};
        rot(r.q, nt, r.t);
        return r;
    }
};
typedef SE3<float> SE3f;
template <typename T>
struct SO3 {
    static Eigen::Matrix<T, 3, 3> hat(const Eigen::Matrix<T, 3, 1>& w) {   // so3.hpp: [0 -c b; c 0 -a; -b a 0]
        Eigen::Matrix<T, 3, 3> m;
        m(0, 1) = -w(2); m(0, 2) = w(1); m(1, 0) = w(2); m(1, 2) = -w(0); m(2, 0) = -w(1); m(2, 1) = w(0);
        return m;
    }
};
typedef SO3<float> SO3f;
template <typename T>
class Sim3 {   // RxSO3 as a NON-unit quaternion (rotation * sqrt(scale)) + translation; point action p' = q p q* + t (rxso3.hpp:265-273,
   public:     // sim3.hpp:227-230 -- the same statement the oracle restates).  inverse() returns what the test stored: the oracle takes both
    T q[4], t[3];   // directions from its caller.
    const Sim3* inv = nullptr;
    Sim3() : q{0, 0, 0, 1}, t{0, 0, 0} {}
    T scale() const { return q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]; }   // quaternion().squaredNorm()
    Eigen::Matrix<T, 3, 1> translation() const { return Eigen::Matrix<T, 3, 1>(t[0], t[1], t[2]); }
    Eigen::Matrix<T, 3, 3> rotationMatrix() const {   // rxso3.hpp: the unit quaternion's matrix
        SE3<T> u;
        const T n = std::sqrt(scale());
        for (int a = 0; a < 4; ++a) u.q[a] = q[a] / n;
        return u.rotationMatrix();
    }
    Sim3 inverse() const { return *inv; }
    Eigen::Matrix<T, 3, 1> operator*(const Eigen::Matrix<T, 3, 1>& p) const {
        // q p q*: Eigen's Quaternion * Quaternion products with p as a pure quaternion (rxso3.hpp:268-272)
        const T px = p(0), py = p(1), pz = p(2);
        const T aw = -q[0] * px - q[1] * py - q[2] * pz, ax = q[3] * px + q[1] * pz - q[2] * py, ay = q[3] * py + q[2] * px - q[0] * pz,
                az = q[3] * pz + q[0] * py - q[1] * px;   // a = q * (0, p)
        const T bx = -aw * q[0] + ax * q[3] - ay * q[2] + az * q[1], by = -aw * q[1] + ay * q[3] - az * q[0] + ax * q[2],
                bz = -aw * q[2] + az * q[3] - ax * q[1] + ay * q[0];   // vec(a * conj(q))
        return Eigen::Matrix<T, 3, 1>(bx + t[0], by + t[1], bz + t[2]);
    }
};
typedef Sim3<float> Sim3f;
}  // namespace Sophus

using namespace std;   // the reference headers and sources rely on it

namespace ORB_SLAM3 {

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

class Frame;
class KeyFrame;

class GeometricCamera {
   public:
    virtual ~GeometricCamera() {}
    virtual Eigen::Vector2f project(const Eigen::Vector3f& v3D) = 0;
    virtual Eigen::Matrix3f toK_() = 0;
    virtual bool epipolarConstrain(GeometricCamera* otherCamera, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const Eigen::Matrix3f& R12,
                                   const Eigen::Vector3f& t12, const float sigmaLevel, const float unc) = 0;
};
class Pinhole : public GeometricCamera {
   public:
    // bodies extracted from src/CameraModels/Pinhole.cpp
    Eigen::Vector2f project(const Eigen::Vector3f& v3D);
    Eigen::Matrix3f toK_();
    bool epipolarConstrain(GeometricCamera* pCamera2, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const Eigen::Matrix3f& R12,
                           const Eigen::Vector3f& t12, const float sigmaLevel, const float unc);
    std::vector<float> mvParameters;
};

class MapPoint {
   public:
    MapPoint() : mTrackProjX(0), mTrackProjY(0), mTrackDepth(0), mTrackDepthR(0), mTrackProjXR(0), mTrackProjYR(0), mbTrackInView(false),
                 mbTrackInViewR(false), mnTrackScaleLevel(0), mnTrackScaleLevelR(0), mTrackViewCos(0), mTrackViewCosR(0), nObs(1),
                 mfMinDistance(0), mfMaxDistance(0) {}
    Eigen::Vector3f GetWorldPos() { return mWorldPos; }
    Eigen::Vector3f GetNormal() { return mNormalVector; }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    int Observations() { return nObs; }
    bool isBad() { g_last_asked = this; return mbBad; }     // the matchers ask every query point first: who is being processed
    bool mbBad = false;
    static MapPoint* g_last_asked;
    bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) != 0; }
    std::tuple<int, int> GetIndexInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) ? mObservations[pKF] : std::tuple<int, int>(-1, -1); }
    void AddObservation(KeyFrame* pKF, int idx) { mObservations[pKF] = std::tuple<int, int>(idx, -1); }
    void Replace(MapPoint*) {}                              // Fuse's map surgery: not part of what is compared
    int PredictScale(const float& currentDist, KeyFrame* pKF);
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }   // MapPoint.cc:658-672
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
    int PredictScale(const float& currentDist, Frame* pF);             // bodies extracted from src/MapPoint.cc
    void ComputeDistinctiveDescriptors();
    void UpdateNormalAndDepth();
    float mTrackProjX, mTrackProjY, mTrackDepth, mTrackDepthR, mTrackProjXR, mTrackProjYR;
    bool mbTrackInView, mbTrackInViewR;
    int mnTrackScaleLevel, mnTrackScaleLevelR;
    float mTrackViewCos, mTrackViewCosR;
    // state
    int nObs;
    Eigen::Vector3f mWorldPos, mNormalVector;
    cv::Mat mDescriptor;
    float mfMinDistance, mfMaxDistance;
    std::mutex mMutexPos, mMutexFeatures;
    std::map<KeyFrame*, std::tuple<int, int> > mObservations;
    KeyFrame* mpRefKF = nullptr;
    int query_index = -1;   // not in the reference: which query of the flat test arrays this object is
};

class Frame {
   public:
    Frame() : mpORBextractorLeft(nullptr), mpORBextractorRight(nullptr), mbf(0), mb(0), N(0), mnScaleLevels(8), mfLogScaleFactor(0),
              mpCamera(nullptr), Nleft(-1), Nright(-1) {}
    // extracted from src/Frame.cc
    void AssignFeaturesToGrid();
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1,
                                     const bool bRight = false) const;
    bool isInFrustum(MapPoint* pMP, float viewingCosLimit);
    bool isInFrustumChecks(MapPoint*, float, bool = false) { return false; }   // fisheye branch (Nleft != -1): not exercised
    void ComputeStereoMatches();
    inline Sophus::SE3<float> GetPose() const { return mTcw; }
    Sophus::SE3<float> GetRelativePoseTrl() const { return Sophus::SE3<float>(); }   // fisheye branch: not exercised
    ORBextractor *mpORBextractorLeft, *mpORBextractorRight;
    float mbf, mb;
    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors, mDescriptorsRight;
    std::vector<bool> mvbOutlier;
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    std::vector<std::size_t> mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    int mnScaleLevels;
    float mfLogScaleFactor;
    vector<float> mvScaleFactors, mvInvScaleFactors;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
    GeometricCamera* mpCamera;
    GeometricCamera* mpCamera2 = nullptr;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    int Nleft, Nright;
    std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;   // fisheye branch (Nleft != -1): not exercised
    Sophus::SE3<float> mTcw;
    Eigen::Matrix<float, 3, 3> mRcw;
    Eigen::Matrix<float, 3, 1> mtcw, mOw;
};

class KeyFrame {   // the members the extracted KeyFrame-typed matchers read (single camera: NLeft == -1, mpCamera2 == nullptr)
   public:
    KeyFrame() : N(0), NLeft(-1), NRight(-1), mpCamera(nullptr), mpCamera2(nullptr) {}
    vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    MapPoint* GetMapPoint(const size_t& idx) {   // Fuse asks for the map point at bestIdx once per fused query: log (query, feature)
        if (log_get && MapPoint::g_last_asked) log_get->push_back(std::make_pair(MapPoint::g_last_asked->query_index, (int)idx));
        return mvpMapPoints[idx];
    }
    std::vector<std::pair<int, int> >* log_get = nullptr;
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
    std::set<MapPoint*> GetMapPoints() {
        std::set<MapPoint*> s;
        for (size_t i = 0; i < mvpMapPoints.size(); ++i)
            if (mvpMapPoints[i] && !mvpMapPoints[i]->mbBad) s.insert(mvpMapPoints[i]);
        return s;
    }
    // bodies extracted from src/KeyFrame.cc
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const bool bRight = false) const;
    bool IsInImage(const float& x, const float& y) const;
    int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
    std::vector<std::vector<std::vector<size_t> > > mGrid, mGridRight;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0, mfLogScaleFactor = 0;
    Sophus::SE3f GetPose() { return mTcw; }
    Sophus::SE3f GetPoseInverse() { return mTcw.inverse(); }
    Eigen::Vector3f GetCameraCenter() { return mTcw.inverse().translation(); }
    Sophus::SE3f GetRightPose() { return Sophus::SE3f(); }          // two-camera rigs (mpCamera2): not exercised
    Sophus::SE3f GetRightPoseInverse() { return Sophus::SE3f(); }
    Eigen::Vector3f GetRightCameraCenter() { return Eigen::Vector3f(); }
    bool isBad() { return mbBad; }
    bool mbBad = false;
    int mnScaleLevels = 8;
    int N, NLeft, NRight;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    GeometricCamera *mpCamera, *mpCamera2;
    std::vector<MapPoint*> mvpMapPoints;
    Sophus::SE3f mTcw;
};

}  // namespace ORB_SLAM3

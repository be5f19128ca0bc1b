// oracle/matcher_oracle.cpp -- TEST INFRASTRUCTURE (see oracle.h).
// CPU restatement of the projection-search matchers for the non-fisheye case (Frame::Nleft == -1):
//   * Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea   (src/Frame.cc:469-504, 962-978, 859-951)
//   * ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>, th, bFar, thFar)   (src/ORBmatcher.cc:45-239)
//   * ORBmatcher::SearchByProjection(Frame& cur, const Frame& last, th, bMono)     (src/ORBmatcher.cc:1950-2184)
//   * ORBmatcher::ComputeThreeMaxima                                               (src/ORBmatcher.cc:2335-2377)
#include <cmath>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace orb_oracle {

static const int GRID_COLS = 64, GRID_ROWS = 48;  // include/Frame.h:46-47

struct FrameView {
    const KeyPoint* kps;      // mvKeysUn (== mvKeys for rectified stereo)
    const uint8_t* desc;
    const float* uright;      // mvuRight (may be null: monocular => -1)
    int N;
    float minX, maxX, minY, maxY;
    float invW, invH;         // mfGridElementWidthInv/HeightInv
    const float* scaleFactors;
    std::vector<size_t> grid[64][48];

    void build() {  // AssignFeaturesToGrid
        invW = (float)GRID_COLS / (maxX - minX);
        invH = (float)GRID_ROWS / (maxY - minY);
        for (int i = 0; i < N; ++i) {
            const int px = (int)std::round((kps[i].x - minX) * invW);
            const int py = (int)std::round((kps[i].y - minY) * invH);
            if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
            grid[px][py].push_back(i);
        }
    }

    std::vector<size_t> area(float x, float y, float r, int minLevel, int maxLevel) const {  // GetFeaturesInArea
        std::vector<size_t> out;
        const int c0 = std::max(0, (int)std::floor((x - minX - r) * invW));
        if (c0 >= GRID_COLS) return out;
        const int c1 = std::min(GRID_COLS - 1, (int)std::ceil((x - minX + r) * invW));
        if (c1 < 0) return out;
        const int r0 = std::max(0, (int)std::floor((y - minY - r) * invH));
        if (r0 >= GRID_ROWS) return out;
        const int r1 = std::min(GRID_ROWS - 1, (int)std::ceil((y - minY + r) * invH));
        if (r1 < 0) return out;
        const bool checkLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = c0; ix <= c1; ++ix)
            for (int iy = r0; iy <= r1; ++iy)
                for (size_t id : grid[ix][iy]) {
                    const KeyPoint& k = kps[id];
                    if (checkLevels) {
                        if (k.octave < minLevel) continue;
                        if (maxLevel >= 0 && k.octave > maxLevel) continue;
                    }
                    const float dx = k.x - x, dy = k.y - y;
                    if (std::fabs(dx) < r && std::fabs(dy) < r) out.push_back(id);
                }
        return out;
    }
};

}  // namespace orb_oracle

using namespace orb_oracle;

extern "C" {

// SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints).  claimed[idx] != 0 <=> F.mvpMapPoints[idx]
// already holds a map point with Observations() > 0 (updated as matches are made: local map points have
// observations).  match[q] = feature index or -1.  Returns nmatches.
int orc_search_local(const KeyPoint* kps, const uint8_t* desc, const float* uright, int N, const float* bounds4,
                     const float* scaleFactors, int nq, const float* projx, const float* projy, const float* projxr,
                     const int* level, const float* viewcos, const float* trackdepth, const uint8_t* qdesc,
                     uint8_t* claimed, float th, float nnratio, int bFar, float thFar, int* match) {
    FrameView F;
    F.kps = kps; F.desc = desc; F.uright = uright; F.N = N;
    F.minX = bounds4[0]; F.maxX = bounds4[1]; F.minY = bounds4[2]; F.maxY = bounds4[3];
    F.scaleFactors = scaleFactors;
    F.build();
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    for (int q = 0; q < nq; ++q) {
        match[q] = -1;
        if (bFar && trackdepth[q] > thFar) continue;
        const int lvl = level[q];
        float r = (viewcos[q] > 0.998) ? 2.5f : 4.0f;   // RadiusByViewingCos: literal 0.998 is a double
        if (bFactor) r *= th;
        const std::vector<size_t> ind = F.area(projx[q], projy[q], r * scaleFactors[lvl], lvl - 1, lvl);
        if (ind.empty()) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (size_t idx : ind) {
            if (claimed[idx]) continue;
            if (uright && uright[idx] > 0) {
                const float er = std::fabs(projxr[q] - uright[idx]);
                if (er > r * scaleFactors[lvl]) continue;
            }
            const int dist = descriptor_distance(qdesc + 32 * (size_t)q, desc + 32 * idx);
            if (dist < bestDist) {
                bestDist2 = bestDist;
                bestDist = dist;
                bestLevel2 = bestLevel;
                bestLevel = kps[idx].octave;
                bestIdx = (int)idx;
            } else if (dist < bestDist2) {
                bestLevel2 = kps[idx].octave;
                bestDist2 = dist;
            }
        }
        if (bestDist <= 100) {  // TH_HIGH
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            if (bestLevel != bestLevel2 || bestDist <= nnratio * bestDist2) {
                match[q] = bestIdx;
                claimed[bestIdx] = 1;
                ++nmatches;
            }
        }
    }
    return nmatches;
}

// Sophus::SE3f * Vector3f (so3.hpp:358-367, se3.hpp:321-324), float32, non-contracted
static void se3_act(const float* q /*x y z w*/, const float* t, const float* p, float* out) {
    const float uvx = q[1] * p[2] - q[2] * p[1], uvy = q[2] * p[0] - q[0] * p[2], uvz = q[0] * p[1] - q[1] * p[0];
    const float ux = uvx + uvx, uy = uvy + uvy, uz = uvz + uvz;
    const float cx = q[1] * uz - q[2] * uy, cy = q[2] * ux - q[0] * uz, cz = q[0] * uy - q[1] * ux;
    out[0] = ((p[0] + q[3] * ux) + cx) + t[0];
    out[1] = ((p[1] + q[3] * uy) + cy) + t[1];
    out[2] = ((p[2] + q[3] * uz) + cz) + t[2];
}

// SearchByProjection(CurrentFrame, LastFrame, th, bMono), Nleft == -1.  Queries = LastFrame features with a
// map point that is not an outlier, in LastFrame order.  direction: 0 none, 1 forward, 2 backward.
// obs_pos[q] != 0 <=> pMP->Observations() > 0.  feat_match[idx] = query index holding the feature or -1
// (CurrentFrame.mvpMapPoints is all-NULL on entry, Tracking.cc:3367).  Returns nmatches.
int orc_search_last(const KeyPoint* kps, const uint8_t* desc, const float* uright, int N, const float* bounds4,
                    const float* scaleFactors, const float* cam6 /*fx fy cx cy bf b*/, const float* Tcw7, int direction,
                    int nq, const float* xw, const int* lastOctave, const float* lastAngle, const uint8_t* qdesc,
                    const uint8_t* obs_pos, float th, int checkOri, int* feat_match) {
    FrameView F;
    F.kps = kps; F.desc = desc; F.uright = uright; F.N = N;
    F.minX = bounds4[0]; F.maxX = bounds4[1]; F.minY = bounds4[2]; F.maxY = bounds4[3];
    F.scaleFactors = scaleFactors;
    F.build();
    for (int i = 0; i < N; ++i) feat_match[i] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[30];
    const float factor = 1.0f / 30;
    const bool fwd = direction == 1, bwd = direction == 2;
    for (int q = 0; q < nq; ++q) {
        float pc[3];
        se3_act(Tcw7, Tcw7 + 4, xw + 3 * q, pc);
        const float invzc = (float)(1.0 / pc[2]);
        if (invzc < 0) continue;
        const float u = cam6[0] * pc[0] / pc[2] + cam6[2], v = cam6[1] * pc[1] / pc[2] + cam6[3];
        if (u < F.minX || u > F.maxX) continue;
        if (v < F.minY || v > F.maxY) continue;
        const int oct = lastOctave[q];
        const float radius = th * scaleFactors[oct];
        std::vector<size_t> ind;
        if (fwd) ind = F.area(u, v, radius, oct, -1);
        else if (bwd) ind = F.area(u, v, radius, 0, oct);
        else ind = F.area(u, v, radius, oct - 1, oct + 1);
        if (ind.empty()) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (size_t i2 : ind) {
            if (feat_match[i2] >= 0 && obs_pos[feat_match[i2]]) continue;
            if (uright && uright[i2] > 0) {
                const float ur = u - cam6[4] * invzc;
                const float er = std::fabs(ur - uright[i2]);
                if (er > radius) continue;
            }
            const int dist = descriptor_distance(qdesc + 32 * (size_t)q, desc + 32 * i2);
            if (dist < bestDist) {
                bestDist = dist;
                bestIdx2 = (int)i2;
            }
        }
        if (bestDist <= 100) {
            feat_match[bestIdx2] = q;
            ++nmatches;
            if (checkOri) {
                float rot = lastAngle[q] - kps[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == 30) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (checkOri) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < 30; ++i) {
            const int s = (int)rotHist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < 30; ++i)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int idx : rotHist[i]) {
                    feat_match[idx] = -1;
                    --nmatches;
                }
    }
    return nmatches;
}
}

// ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches), Nleft == -1
// (src/ORBmatcher.cc:259-493).  Queries = the keyframe's features that hold a good map point, in the order of the
// FeatureVector merge join (ascending vocabulary node id, ascending keyframe feature index inside a node) -- the
// caller provides them in that order with their node id.  feat_node[i] = node of frame feature i (F.mFeatVec), -1 if
// none.  feat_match[i] = query index now stored in vpMapPointMatches[i], or -1.  Returns nmatches.
extern "C" int orc_search_bow(const KeyPoint* kps, const uint8_t* desc, const int* feat_node, int N, int nq,
                              const int* qnode, const float* qangle, const uint8_t* qdesc, float nnratio, int checkOri,
                              int* feat_match) {
    for (int i = 0; i < N; ++i) feat_match[i] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[30];
    const float factor = 1.0f / 30;
    for (int q = 0; q < nq; ++q) {
        int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
        for (int i = 0; i < N; ++i) {            // vIndicesF: features of the same node, ascending index
            if (feat_node[i] != qnode[q] || feat_node[i] < 0) continue;
            if (feat_match[i] >= 0) continue;    // :331
            const int dist = descriptor_distance(qdesc + 32 * (size_t)q, desc + 32 * (size_t)i);
            if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = i; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist1 <= 50) {                   // TH_LOW
            if ((float)bestDist1 < nnratio * (float)bestDist2) {
                feat_match[bestIdxF] = q;
                if (checkOri) {
                    float rot = qangle[q] - kps[bestIdxF].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)std::round(rot * factor);
                    if (bin == 30) bin = 0;
                    rotHist[bin].push_back(bestIdxF);
                }
                ++nmatches;
            }
        }
    }
    if (checkOri) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < 30; ++i) {
            const int s = (int)rotHist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < 30; ++i)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int idx : rotHist[i]) { feat_match[idx] = -1; --nmatches; }
    }
    return nmatches;
}

// ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse), no second camera
// (src/ORBmatcher.cc:1045-1323; Pinhole::epipolarConstrain src/CameraModels/Pinhole.cpp:186-209).
// Queries = KF1 features WITHOUT a map point (and stereo when bOnlyStereo), in FeatureVector merge order, with their
// node.  KF2 candidates: features of the same node without a map point (valid2 != 0), ascending index.
// F12 = K1^-T [t12]x R12 K2^-1 is computed by the caller exactly as the reference does (Eigen, float).
// match12[q] = KF2 feature index or -1.  Returns nmatches.
extern "C" int orc_search_triangulation(int nq, const KeyPoint* kp1, const uint8_t* desc1, const int* node1, const uint8_t* stereo1,
                                        int N2, const KeyPoint* kp2, const uint8_t* desc2, const int* node2, const uint8_t* valid2,
                                        const uint8_t* stereo2, const float* F12, const float* ep2, const float* scaleFactors,
                                        const float* sigma2, int bCoarse, int checkOri, int* match12) {
    int nmatches = 0;
    std::vector<int> rotHist[30];
    const float factor = 1.0f / 30;
    for (int q = 0; q < nq; ++q) {
        match12[q] = -1;
        int bestDist = 50, bestIdx2 = -1;   // TH_LOW
        for (int i2 = 0; i2 < N2; ++i2) {
            if (node2[i2] != node1[q] || node2[i2] < 0) continue;
            if (!valid2[i2]) continue;
            const int dist = descriptor_distance(desc1 + 32 * (size_t)q, desc2 + 32 * (size_t)i2);
            if (dist > 50 || dist > bestDist) continue;
            if (!stereo1[q] && !stereo2[i2]) {
                const float distex = ep2[0] - kp2[i2].x, distey = ep2[1] - kp2[i2].y;
                if (distex * distex + distey * distey < 100 * scaleFactors[kp2[i2].octave]) continue;
            }
            bool ok = bCoarse != 0;
            if (!ok) {   // Pinhole::epipolarConstrain(pCamera2, kp1, kp2, R12, t12, sigmaLevel, unc = sigma2[kp2.octave])
                const float a = kp1[q].x * F12[0] + kp1[q].y * F12[3] + F12[6];
                const float b = kp1[q].x * F12[1] + kp1[q].y * F12[4] + F12[7];
                const float c = kp1[q].x * F12[2] + kp1[q].y * F12[5] + F12[8];
                const float num = a * kp2[i2].x + b * kp2[i2].y + c;
                const float den = a * a + b * b;
                if (den != 0) {
                    const float dsqr = num * num / den;
                    ok = dsqr < 3.84 * sigma2[kp2[i2].octave];
                }
            }
            if (ok) { bestIdx2 = i2; bestDist = dist; }
        }
        if (bestIdx2 >= 0) {
            match12[q] = bestIdx2;
            ++nmatches;
            if (checkOri) {
                float rot = kp1[q].angle - kp2[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == 30) bin = 0;
                rotHist[bin].push_back(q);
            }
        }
    }
    if (checkOri) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < 30; ++i) {
            const int s = (int)rotHist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < 30; ++i)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int q : rotHist[i]) { match12[q] = -1; --nmatches; }
    }
    return nmatches;
}

// ------------------------------------------------------------------------------------------------
// Projection searches into a KeyFrame (or, variant 3, a Frame) from a pose, Nleft == -1, pinhole:
//   variant 0  Fuse(pKF, vpMapPoints, th, bRight=false)                         src/ORBmatcher.cc:1325-1544
//   variant 1  Fuse(pKF, Scw, vpPoints, th, vpReplacePoint)                     src/ORBmatcher.cc:1546-1687
//   variant 2  SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratio)     src/ORBmatcher.cc:495-618
//              (and the vpPointsKFs overload :620-732, which only records one more array per match)
//   variant 3  SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) src/ORBmatcher.cc:2196-2330
//   variant 4  one direction of SearchBySim3(pKF1, pKF2, vpMatches12, S12, th)  src/ORBmatcher.cc:1689-1948:
//              p_c = S * (Tcw * p_w) with S = (non-unit quaternion x y z w, translation, scale = quaternion().squaredNorm()
//              as the caller's Eigen computes it) -- Sophus rxso3.hpp:265-273, sim3.hpp:227-230; queries = the map points
//              of the other keyframe that are good and not already matched; no claims; TH_HIGH.
// Queries = the map points that pass the caller-side skips (NULL, isBad, IsInKeyFrame / spAlreadyFound), in order.
// Tcw7: qx qy qz qw tx ty tz of the SE3f the function builds (for Scw: SE3f(Scw.rotationMatrix(),
// Scw.translation()/Scw.scale())); Ow = Tcw.inverse().translation() / pKF->GetCameraCenter().
// claimed[idx] (variants 2, 3): vpMatched[idx] / CurrentFrame.mvpMapPoints[i2] non-NULL on entry; updated.
// thr: 50 (TH_LOW), TH_LOW * ratioHamming (float product) or ORBdist.  match[q] = feature index or -1.
// Vector3f norm / dot follow Eigen's unrolled reduction x + (y + z) (Eigen/src/Core/Redux.h, not in the tree).
// MapPoint::PredictScale  src/MapPoint.cc:688-721 with std::log(float) == logf.
extern "C" int orc_search_keyframe(int variant, const KeyPoint* kps, const uint8_t* desc, const float* uright, int N,
                                   const float* bounds4, const float* scaleFactors, const float* invLevelSigma2, int nLevels,
                                   float logScaleFactor, const float* cam6, const float* Tcw7, const float* Ow,
                                   const float* S8, int nq,
                                   const float* xw, const float* normal, const float* maxDist, const float* minDist,
                                   const uint8_t* qdesc, const float* qangle, uint8_t* claimed, float th, float thr,
                                   int checkOri, int* match) {
    FrameView F;
    F.kps = kps; F.desc = desc; F.uright = uright; F.N = N;
    F.minX = bounds4[0]; F.maxX = bounds4[1]; F.minY = bounds4[2]; F.maxY = bounds4[3];
    F.scaleFactors = scaleFactors;
    F.build();
    int nmatches = 0;
    std::vector<int> rotHist[30];
    std::vector<int> rotQuery[30];
    const float factor = 1.0f / 30;
    const float fx = cam6[0], fy = cam6[1], cx = cam6[2], cy = cam6[3], bf = cam6[4];
    for (int q = 0; q < nq; ++q) {
        match[q] = -1;
        const float* p3Dw = xw + 3 * q;
        float pc[3];
        se3_act(Tcw7, Tcw7 + 4, p3Dw, pc);
        float u, v, invz;
        if (variant == 4) {
            const float* sq = S8;          // x y z w
            const float* p = pc;
            const float uvx = sq[1] * p[2] - sq[2] * p[1], uvy = sq[2] * p[0] - sq[0] * p[2], uvz = sq[0] * p[1] - sq[1] * p[0];
            const float ux = uvx + uvx, uy = uvy + uvy, uz = uvz + uvz;
            const float c0 = sq[1] * uz - sq[2] * uy, c1 = sq[2] * ux - sq[0] * uz, c2 = sq[0] * uy - sq[1] * ux;
            const float p2[3] = {(S8[7] * p[0] + (sq[3] * ux + c0)) + S8[4], (S8[7] * p[1] + (sq[3] * uy + c1)) + S8[5],
                                 (S8[7] * p[2] + (sq[3] * uz + c2)) + S8[6]};
            pc[0] = p2[0]; pc[1] = p2[1]; pc[2] = p2[2];
            if (pc[2] < 0.0) continue;
            invz = (float)(1.0 / pc[2]);
            const float x = pc[0] * invz, y = pc[1] * invz;
            u = fx * x + cx;
            v = fy * y + cy;
        } else {
            if (variant != 3 && pc[2] < 0.0f) continue;
            invz = 1 / pc[2];
            u = fx * pc[0] / pc[2] + cx;       // Pinhole::project
            v = fy * pc[1] / pc[2] + cy;
        }
        if (variant == 3) {
            if (u < F.minX || u > F.maxX) continue;
            if (v < F.minY || v > F.maxY) continue;
        } else if (!(u >= F.minX && u < F.maxX && v >= F.minY && v < F.maxY)) continue;   // KeyFrame::IsInImage
        const float ur = u - bf * invz;
        float PO[3] = {p3Dw[0] - Ow[0], p3Dw[1] - Ow[1], p3Dw[2] - Ow[2]};
        if (variant == 4) { PO[0] = pc[0]; PO[1] = pc[1]; PO[2] = pc[2]; }      // dist3D = p3Dc2.norm()
        const float dist3D = std::sqrt(PO[0] * PO[0] + (PO[1] * PO[1] + PO[2] * PO[2]));
        if (dist3D < 0.8f * minDist[q] || dist3D > 1.2f * maxDist[q]) continue;   // Get{Min,Max}DistanceInvariance (MapPoint.cc:658-672); minDist / maxDist = mfMinDistance / mfMaxDistance
        if (variant < 3) {
            const float* Pn = normal + 3 * q;
            const float d = PO[0] * Pn[0] + (PO[1] * Pn[1] + PO[2] * Pn[2]);
            if (d < 0.5 * dist3D) continue;
        }
        const float ratio = maxDist[q] / dist3D;
        int nPredictedLevel = (int)std::ceil(logf(ratio) / logScaleFactor);
        if (nPredictedLevel < 0) nPredictedLevel = 0;
        else if (nPredictedLevel >= nLevels) nPredictedLevel = nLevels - 1;
        const float radius = th * scaleFactors[nPredictedLevel];
        const std::vector<size_t> ind = variant == 3 ? F.area(u, v, radius, nPredictedLevel - 1, nPredictedLevel + 1)
                                                     : F.area(u, v, radius, -1, -1);
        if (ind.empty()) continue;
        int bestDist = variant == 0 || variant == 2 || variant == 3 ? 256 : 0x7fffffff, bestIdx = -1;
        for (size_t idx : ind) {
            if ((variant == 2 || variant == 3) && claimed[idx]) continue;
            const int kpLevel = kps[idx].octave;
            if (variant != 3 && (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel)) continue;
            if (variant == 0) {
                if (uright && uright[idx] >= 0) {
                    const float ex = u - kps[idx].x, ey = v - kps[idx].y, er = ur - uright[idx];
                    const float e2 = ex * ex + ey * ey + er * er;
                    if (e2 * invLevelSigma2[kpLevel] > 7.8) continue;
                } else {
                    const float ex = u - kps[idx].x, ey = v - kps[idx].y;
                    const float e2 = ex * ex + ey * ey;
                    if (e2 * invLevelSigma2[kpLevel] > 5.99) continue;
                }
            }
            const int dist = descriptor_distance(qdesc + 32 * (size_t)q, desc + 32 * idx);
            if (dist < bestDist) {
                bestDist = dist;
                bestIdx = (int)idx;
            }
        }
        if ((float)bestDist <= thr) {
            match[q] = bestIdx;
            if (variant == 2 || variant == 3) claimed[bestIdx] = 1;
            ++nmatches;
            if (variant == 3 && checkOri) {
                float rot = qangle[q] - kps[bestIdx].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == 30) bin = 0;
                rotHist[bin].push_back(bestIdx);
                rotQuery[bin].push_back(q);
            }
        }
    }
    if (variant == 3 && checkOri) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < 30; ++i) {
            const int s = (int)rotHist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < 30; ++i)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0; j < rotHist[i].size(); ++j) {
                    claimed[rotHist[i][j]] = 0;       // CurrentFrame.mvpMapPoints[..] = NULL
                    match[rotQuery[i][j]] = -1;
                    --nmatches;
                }
    }
    return nmatches;
}

// ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12)  src/ORBmatcher.cc:892-1043,
// single-camera keyframes.  Queries = KF1 features with a good map point in FeatureVector merge order; node2[i] = node of
// KF2 feature i (-1 none), valid2[i] = holds a good map point.  match12[q] = idx2 or -1.  Returns nmatches.
extern "C" int orc_search_bow_kf(const KeyPoint* kp2, const uint8_t* desc2, const int* node2, const uint8_t* valid2, int n2, int nq,
                                 const int* qnode, const float* qangle, const uint8_t* desc1, float nnratio, int checkOri,
                                 int* match12) {
    std::vector<char> vbMatched2(n2, 0);
    std::vector<int> rotHist[30];
    const float factor = 1.0f / 30;
    int nmatches = 0;
    for (int q = 0; q < nq; ++q) {
        match12[q] = -1;
        if (qnode[q] < 0) continue;
        int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
        for (int idx2 = 0; idx2 < n2; ++idx2) {          // f2it->second: ascending feature index inside the node
            if (node2[idx2] != qnode[q]) continue;
            if (vbMatched2[idx2] || !(valid2 ? valid2[idx2] : 1)) continue;
            const int dist = descriptor_distance(desc1 + 32 * (size_t)q, desc2 + 32 * (size_t)idx2);
            if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist1 < 50) {
            if ((float)bestDist1 < nnratio * (float)bestDist2) {
                match12[q] = bestIdx2;
                vbMatched2[bestIdx2] = 1;
                if (checkOri) {
                    float rot = qangle[q] - kp2[bestIdx2].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)std::round(rot * factor);
                    if (bin == 30) bin = 0;
                    rotHist[bin].push_back(q);
                }
                ++nmatches;
            }
        }
    }
    if (checkOri) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < 30; ++i) {
            const int s = (int)rotHist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < 30; ++i)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int q : rotHist[i]) {
                    match12[q] = -1;
                    --nmatches;
                }
    }
    return nmatches;
}

// ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)  src/ORBmatcher.cc:734-890.
// matches12[i1] = i2 or -1.  Returns nmatches.  (The caller refreshes vbPrevMatched, :880-884.)
extern "C" int orc_search_initialization(const KeyPoint* kp1, const uint8_t* desc1, int n1, const float* prevMatched,
                                         const KeyPoint* kp2, const uint8_t* desc2, int n2, const float* bounds4, int windowSize,
                                         float nnratio, int checkOri, int* matches12) {
    FrameView F;
    F.kps = kp2; F.desc = desc2; F.uright = nullptr; F.N = n2;
    F.minX = bounds4[0]; F.maxX = bounds4[1]; F.minY = bounds4[2]; F.maxY = bounds4[3];
    F.scaleFactors = nullptr;
    F.build();
    int nmatches = 0;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    std::vector<int> rotHist[30];
    const float factor = 1.0f / 30;
    std::vector<int> vMatchedDistance(n2, 0x7fffffff), vnMatches21(n2, -1);
    for (int i1 = 0; i1 < n1; ++i1) {
        const int level1 = kp1[i1].octave;
        if (level1 > 0) continue;
        const std::vector<size_t> ind = F.area(prevMatched[2 * i1], prevMatched[2 * i1 + 1], (float)windowSize, level1, level1);
        if (ind.empty()) continue;
        int bestDist = 0x7fffffff, bestDist2 = 0x7fffffff, bestIdx2 = -1;
        for (size_t i2 : ind) {
            const int dist = descriptor_distance(desc1 + 32 * (size_t)i1, desc2 + 32 * i2);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = (int)i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= 50) {
            if (bestDist < (float)bestDist2 * nnratio) {
                if (vnMatches21[bestIdx2] >= 0) {
                    matches12[vnMatches21[bestIdx2]] = -1;
                    --nmatches;
                }
                matches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                ++nmatches;
                if (checkOri) {
                    float rot = kp1[i1].angle - kp2[bestIdx2].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)std::round(rot * factor);
                    if (bin == 30) bin = 0;
                    rotHist[bin].push_back(i1);
                }
            }
        }
    }
    if (checkOri) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < 30; ++i) {
            const int s = (int)rotHist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < 30; ++i)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int idx1 : rotHist[i])
                    if (matches12[idx1] >= 0) {
                        matches12[idx1] = -1;
                        --nmatches;
                    }
    }
    return nmatches;
}

// Frame::isInFrustum(MapPoint*, viewingCosLimit)  src/Frame.cc:667-720, Nleft == -1.  Per map point: Pc = mRcw * P + mtcw
// (Eigen fixed-size product / norm / dot in the unrolled reduction order x + (y + z), Eigen/src/Core/Redux.h -- not in the
// tree), depth / image-bounds / scale-invariance-distance / viewing-angle gates, MapPoint::PredictScale.
// Outputs per point: in_view (mbTrackInView), mTrackProjX / Y / XR, mnTrackScaleLevel, mTrackViewCos, mTrackDepth.
// Points that fail keep proj = -1 / level -1 except that mTrackProjX / Y are already set when only the distance or
// viewing-angle gate fails (:690-691), as in the reference.
extern "C" int orc_is_in_frustum(int n, const float* Rcw9, const float* tcw, const float* Ow, const float* bounds4, const float* cam6,
                                 int nLevels, float logScaleFactor, float viewingCosLimit, const float* xw, const float* normal,
                                 const float* maxDist, const float* minDist, uint8_t* in_view, float* projx, float* projy, float* projxr,
                                 int* level, float* viewcos, float* depth) {
    int nvis = 0;
    for (int i = 0; i < n; ++i) {
        in_view[i] = 0; projx[i] = -1; projy[i] = -1; projxr[i] = -1; level[i] = -1; viewcos[i] = 0; depth[i] = 0;
        const float* P = xw + 3 * i;
        float Pc[3];
        for (int r = 0; r < 3; ++r) Pc[r] = (Rcw9[3 * r] * P[0] + (Rcw9[3 * r + 1] * P[1] + Rcw9[3 * r + 2] * P[2])) + tcw[r];
        const float Pc_dist = std::sqrt(Pc[0] * Pc[0] + (Pc[1] * Pc[1] + Pc[2] * Pc[2]));
        const float PcZ = Pc[2];
        const float invz = 1.0f / PcZ;
        if (PcZ < 0.0f) continue;
        const float u = cam6[0] * Pc[0] / Pc[2] + cam6[2], v = cam6[1] * Pc[1] / Pc[2] + cam6[3];
        if (u < bounds4[0] || u > bounds4[1]) continue;
        if (v < bounds4[2] || v > bounds4[3]) continue;
        projx[i] = u;
        projy[i] = v;
        const float PO[3] = {P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2]};
        const float dist = std::sqrt(PO[0] * PO[0] + (PO[1] * PO[1] + PO[2] * PO[2]));
        if (dist < 0.8f * minDist[i] || dist > 1.2f * maxDist[i]) continue;   // Get{Min,Max}DistanceInvariance; PredictScale below takes the raw mfMaxDistance
        const float* Pn = normal + 3 * i;
        const float vc = (PO[0] * Pn[0] + (PO[1] * Pn[1] + PO[2] * Pn[2])) / dist;
        if (vc < viewingCosLimit) continue;
        const float ratio = maxDist[i] / dist;
        int lvl = (int)std::ceil(logf(ratio) / logScaleFactor);
        if (lvl < 0) lvl = 0;
        else if (lvl >= nLevels) lvl = nLevels - 1;
        in_view[i] = 1;
        projxr[i] = u - cam6[4] * invz;
        depth[i] = Pc_dist;
        level[i] = lvl;
        viewcos[i] = vc;
        ++nvis;
    }
    return nvis;
}

// DBoW2 TemplatedVocabulary<FORB>::transform(features, BowVector&, FeatureVector&, levelsup) with TF_IDF + L1
// (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1260, BowVector.cpp:34-84; Frame::ComputeBoW, src/Frame.cc:984-997).
// Tree flattened as in include/orbslam3_b200.h (orbv_create).  Outputs per feature: word, node at levelsup, weight; and the
// BowVector (ascending words, L1-normalised weights).  Returns the BowVector size.
#include <map>
extern "C" int orc_bow_transform(int n_nodes, int L, const int* child_off, const int* child, const uint8_t* ndesc, const int* nword,
                                 const double* nweight, int levelsup, int n, const uint8_t* desc, int* word, int* node, double* weight,
                                 int* bow_word, double* bow_weight) {
    std::map<unsigned, double> v;
    const int nid_level = L - levelsup;
    for (int i = 0; i < n; ++i) {
        int nid = 0;
        int final_id = 0, current_level = 0;
        while (child_off[final_id] != child_off[final_id + 1]) {
            ++current_level;
            const int c0 = child_off[final_id], c1 = child_off[final_id + 1];
            int best = child[c0];
            double best_d = (double)descriptor_distance(desc + 32 * (size_t)i, ndesc + 32 * (size_t)best);
            for (int c = c0 + 1; c < c1; ++c) {
                const double d = (double)descriptor_distance(desc + 32 * (size_t)i, ndesc + 32 * (size_t)child[c]);
                if (d < best_d) { best_d = d; best = child[c]; }
            }
            final_id = best;
            if (current_level == nid_level) nid = final_id;
        }
        word[i] = nword[final_id];
        weight[i] = nweight[final_id];
        node[i] = nid;
        if (weight[i] > 0) {
            auto it = v.lower_bound((unsigned)word[i]);
            if (it != v.end() && !(v.key_comp()((unsigned)word[i], it->first))) it->second += weight[i];
            else v.insert(it, std::make_pair((unsigned)word[i], weight[i]));
        }
    }
    double norm = 0.0;
    for (auto& kv : v) norm += std::fabs(kv.second);
    if (norm > 0.0)
        for (auto& kv : v) kv.second /= norm;
    int k = 0;
    for (auto& kv : v) { bow_word[k] = (int)kv.first; bow_weight[k] = kv.second; ++k; }
    return k;
}

// MapPoint::ComputeDistinctiveDescriptors  src/MapPoint.cc:438-520 (one map point): returns the position of the chosen
// descriptor in the observation list, -1 when the list is empty.
#include <algorithm>
extern "C" int orc_distinctive_descriptor(int N, const uint8_t* descs) {
    if (N <= 0) return -1;
    std::vector<float> D((size_t)N * N);
    for (int i = 0; i < N; ++i) {
        D[(size_t)i * N + i] = 0;
        for (int j = i + 1; j < N; ++j) {
            const int d = descriptor_distance(descs + 32 * (size_t)i, descs + 32 * (size_t)j);
            D[(size_t)i * N + j] = (float)d;
            D[(size_t)j * N + i] = (float)d;
        }
    }
    int BestMedian = 0x7fffffff, BestIdx = 0;
    for (int i = 0; i < N; ++i) {
        std::vector<int> vDists(D.begin() + (size_t)i * N, D.begin() + (size_t)i * N + N);
        std::sort(vDists.begin(), vDists.end());
        const int median = vDists[(size_t)(0.5 * (N - 1))];
        if (median < BestMedian) { BestMedian = median; BestIdx = i; }
    }
    return BestIdx;
}

// MapPoint::UpdateNormalAndDepth  src/MapPoint.cc:567-640 (one map point, pinhole keyframes): out3 = mNormalVector,
// *maxd / *mind = mfMaxDistance / mfMinDistance.  Vector3f::norm in Eigen's unrolled order x^2 + (y^2 + z^2).
extern "C" void orc_update_normal_and_depth(int N, const float* centers, const float* pos, const float* refCenter, int level,
                                            const float* scaleFactors, int nLevels, float* out3, float* maxd, float* mind) {
    if (N <= 0) return;
    float normal[3] = {0, 0, 0};
    for (int o = 0; o < N; ++o) {
        const float v[3] = {pos[0] - centers[3 * o], pos[1] - centers[3 * o + 1], pos[2] - centers[3 * o + 2]};
        const float nrm = std::sqrt(v[0] * v[0] + (v[1] * v[1] + v[2] * v[2]));
        for (int c = 0; c < 3; ++c) normal[c] = normal[c] + v[c] / nrm;
    }
    const float PC[3] = {pos[0] - refCenter[0], pos[1] - refCenter[1], pos[2] - refCenter[2]};
    const float dist = std::sqrt(PC[0] * PC[0] + (PC[1] * PC[1] + PC[2] * PC[2]));
    *maxd = dist * scaleFactors[level];
    *mind = *maxd / scaleFactors[nLevels - 1];
    for (int c = 0; c < 3; ++c) out3[c] = normal[c] / (float)N;
}

// cv::BFMatcher(cv::NORM_HAMMING).knnMatch(query, train, matches, 2) as Frame::ComputeStereoFishEyeMatches uses it (src/Frame.cc:1553):
// per query the two nearest train rows, ascending distance, ties by ascending train index (what OpenCV returns; pinned against
// cv2 4.13 by tests/test_knn_cpu.py).  idx / dist: [nq][2], -1 where the train set has fewer rows.
extern "C" void orc_hamming_knn2(int nq, const uint8_t* q, int nt, const uint8_t* t, int* idx, int* dist) {
    for (int i = 0; i < nq; ++i) {
        int bi[2] = {-1, -1}, bd[2] = {-1, -1};
        for (int j = 0; j < nt; ++j) {
            int d = 0;
            for (int k = 0; k < 32; ++k) d += __builtin_popcount((unsigned)(q[32 * (size_t)i + k] ^ t[32 * (size_t)j + k]));
            if (bi[0] < 0 || d < bd[0]) { bi[1] = bi[0]; bd[1] = bd[0]; bi[0] = j; bd[0] = d; }
            else if (bi[1] < 0 || d < bd[1]) { bi[1] = j; bd[1] = d; }
        }
        idx[2 * i] = bi[0]; idx[2 * i + 1] = bi[1]; dist[2 * i] = bd[0]; dist[2 * i + 1] = bd[1];
    }
}

// oracle/stereo_oracle.cpp -- TEST INFRASTRUCTURE (see oracle.h).
// CPU restatement of Frame::ComputeStereoMatches (/root/reference/src/Frame.cc:1102-1358).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

#include "oracle.h"

namespace orb_oracle {

// kpsL/kpsR: operator() outputs of the two eyes; pyrL/pyrR: mvImagePyramid of the two extractors.
// Writes mvuRight / mvDepth (N entries, -1 = no match).  Returns the number of matches kept.
int compute_stereo_matches(const KeyPoint* kpsL, const uint8_t* descL, int N, const KeyPoint* kpsR,
                           const uint8_t* descR, int Nr, const std::vector<Plane>& pyrL,
                           const std::vector<Plane>& pyrR, const float* scaleFactors, const float* invScaleFactors,
                           float mbf, float mb, float* uRight, float* depth) {
    for (int i = 0; i < N; ++i) uRight[i] = depth[i] = -1.0f;
    const int thOrbDist = (100 + 50) / 2;  // (TH_HIGH + TH_LOW) / 2, ORBmatcher.cc:35-36
    const int nRows = pyrL[0].h;
    std::vector<std::vector<size_t>> rowIdx(nRows);
    for (int iR = 0; iR < Nr; ++iR) {  // Frame.cc:1134-1156
        const float kpY = kpsR[iR].y;
        const float r = 2.0f * scaleFactors[kpsR[iR].octave];
        const int maxr = (int)std::ceil(kpY + r);
        const int minr = (int)std::floor(kpY - r);
        for (int yi = minr; yi <= maxr; ++yi)
            if (yi >= 0 && yi < nRows) rowIdx[yi].push_back(iR);  // (the reference does not bound-check)
    }
    const float minZ = mb, minD = 0, maxD = mbf / minZ;
    std::vector<std::pair<int, int>> distIdx;
    for (int iL = 0; iL < N; ++iL) {
        const KeyPoint& kpL = kpsL[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y, uL = kpL.x;
        const std::vector<size_t>& cand = rowIdx[(size_t)vL];
        if (cand.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = 100;  // TH_HIGH
        size_t bestIdxR = 0;
        for (size_t iC = 0; iC < cand.size(); ++iC) {
            const size_t iR = cand[iC];
            const KeyPoint& kpR = kpsR[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = descriptor_distance(descL + 32 * (size_t)iL, descR + 32 * iR);
                if (dist < bestDist) {
                    bestDist = dist;
                    bestIdxR = iR;
                }
            }
        }
        if (bestDist >= thOrbDist) continue;
        // sub-pixel refinement by 11x11 SAD over 11 shifts, Frame.cc:1232-1330
        const float uR0 = kpsR[bestIdxR].x;
        const float sf = invScaleFactors[kpL.octave];
        const float scaleduL = std::round(kpL.x * sf), scaledvL = std::round(kpL.y * sf), scaleduR0 = std::round(uR0 * sf);
        const int w = 5, L = 5;
        const Plane& IL = pyrL[kpL.octave];
        const Plane& IR = pyrR[kpL.octave];
        const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
        if (iniu < 0 || endu >= IR.w) continue;
        int bestSad = INT_MAX, bestinc = 0;
        float dists[2 * 5 + 1];
        const int cy = (int)scaledvL, cxL = (int)scaleduL, cxR = (int)scaleduR0;
        for (int inc = -L; inc <= L; ++inc) {
            int sad = 0;
            for (int dy = -w; dy <= w; ++dy)
                for (int dx = -w; dx <= w; ++dx)
                    sad += std::abs((int)IL.at(cy + dy, cxL + dx) - (int)IR.at(cy + dy, cxR + inc + dx));
            const float dist = (float)sad;
            if (dist < (float)bestSad) {
                bestSad = (int)dist;
                bestinc = inc;
            }
            dists[L + inc] = dist;
        }
        if (bestinc == -L || bestinc == L) continue;
        const float d1 = dists[L + bestinc - 1], d2 = dists[L + bestinc], d3 = dists[L + bestinc + 1];
        const float deltaR = (d1 - d3) / (2.0f * (d1 + d3 - 2.0f * d2));
        if (deltaR < -1 || deltaR > 1) continue;
        float bestuR = scaleFactors[kpL.octave] * ((float)scaleduR0 + (float)bestinc + deltaR);
        float disparity = uL - bestuR;
        if (disparity >= minD && disparity < maxD) {
            if (disparity <= 0) {
                disparity = 0.01;
                bestuR = uL - 0.01;
            }
            depth[iL] = mbf / disparity;
            uRight[iL] = bestuR;
            distIdx.push_back(std::make_pair(bestSad, iL));
        }
    }
    if (distIdx.empty()) return 0;  // (the reference would index an empty vector)
    std::sort(distIdx.begin(), distIdx.end());
    const float median = (float)distIdx[distIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    int kept = (int)distIdx.size();
    for (int i = (int)distIdx.size() - 1; i >= 0; --i) {
        if ((float)distIdx[i].first < thDist) break;
        uRight[distIdx[i].second] = -1;
        depth[distIdx[i].second] = -1;
        --kept;
    }
    return kept;
}

}  // namespace orb_oracle

using namespace orb_oracle;
extern "C" int orc_stereo(void* hL, void* hR, const KeyPoint* kpsL, const uint8_t* descL, int N, const KeyPoint* kpsR,
                          const uint8_t* descR, int Nr, float mbf, float mb, float* uRight, float* depth) {
    Extractor* L = (Extractor*)hL;
    Extractor* R = (Extractor*)hR;
    return compute_stereo_matches(kpsL, descL, N, kpsR, descR, Nr, L->pyramid, R->pyramid, L->mvScaleFactor.data(),
                                  L->mvInvScaleFactor.data(), mbf, mb, uRight, depth);
}

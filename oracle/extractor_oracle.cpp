// oracle/extractor_oracle.cpp -- TEST INFRASTRUCTURE (see oracle.h).
// CPU restatement of ORBextractor (/root/reference/src/ORBextractor.cc) with the OpenCV stages
// replaced by the integer / float32 models of SURVEY.md Appendix A.  Build: -O3 -march=native
// -ffp-contract=off (non-contracted IEEE float32 is the oracle's definition).
#include "oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <list>
#include <utility>
#include <emmintrin.h>

namespace orb_oracle {

const int8_t kPattern[1024] = {
#include "orb_pattern.inc"
};

static const int kHalfPatch = 15;  // ORBextractor.cc:77
static const int kPatch = 31;      // ORBextractor.cc:76
static const int kEdge = 19;       // ORBextractor.cc:78

static inline int round_half_even(float v) { return (int)lrintf(v); }  // cvRound (SSE cvtss2si)
static inline int round_half_even(double v) { return (int)lrint(v); }

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------------------------------------
// cv::resize(u8, INTER_LINEAR) -- SURVEY App. A.1 (called at ORBextractor.cc:1702).
// ---------------------------------------------------------------------------------------------
static void linear_taps(int ssize, int dsize, std::vector<int>& ofs, std::vector<int16_t>& coef) {
    const double scale = (double)ssize / dsize;
    ofs.resize(dsize);
    coef.resize(2 * (size_t)dsize);
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)std::floor(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
        ofs[d] = s;
        coef[2 * d] = (int16_t)round_half_even((1.f - f) * 2048.f);
        coef[2 * d + 1] = (int16_t)round_half_even(f * 2048.f);
    }
}

void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh,
                      int dstride) {
    if (sw == 2 * dw && sh == 2 * dh) {
        // cv::resize silently switches INTER_LINEAR to the INTER_AREA 2x2 fast path when both
        // ratios are exactly 2 (only reachable with scaleFactor == 2.0).
        for (int y = 0; y < dh; ++y)
            for (int x = 0; x < dw; ++x) {
                const uint8_t* a = src + (size_t)(2 * y) * sstride + 2 * x;
                dst[(size_t)y * dstride + x] = (uint8_t)((a[0] + a[1] + a[sstride] + a[sstride + 1] + 2) >> 2);
            }
        return;
    }
    std::vector<int> xo, yo;
    std::vector<int16_t> xc, yc;
    linear_taps(sw, dw, xo, xc);
    linear_taps(sh, dh, yo, yc);
    std::vector<int32_t> r0(dw), r1(dw);
    for (int y = 0; y < dh; ++y) {
        const int sy0 = yo[y], sy1 = std::min(sy0 + 1, sh - 1);
        const uint8_t* a = src + (size_t)sy0 * sstride;
        const uint8_t* b = src + (size_t)sy1 * sstride;
        for (int x = 0; x < dw; ++x) {
            const int x0 = xo[x], x1 = std::min(x0 + 1, sw - 1);
            r0[x] = a[x0] * xc[2 * x] + a[x1] * xc[2 * x + 1];
            r1[x] = b[x0] * xc[2 * x] + b[x1] * xc[2 * x + 1];
        }
        const int b0 = yc[2 * y], b1 = yc[2 * y + 1];
        uint8_t* o = dst + (size_t)y * dstride;
        for (int x = 0; x < dw; ++x)
            o[x] = (uint8_t)((((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2);
    }
}

// ---------------------------------------------------------------------------------------------
// cv::GaussianBlur(u8, 7x7, sigma=2, BORDER_REFLECT_101) -- SURVEY App. A.2 (ORBextractor.cc:1632)
// ---------------------------------------------------------------------------------------------
static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
    return i;
}

void gaussian_blur7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
    // Q0.8 kernel {18,34,48,56,48,34,18} (sums to 256); horizontal pass keeps Q8.8, vertical pass
    // accumulates Q16.16 and rounds with +32768 >> 16.
    std::vector<uint16_t> hbuf((size_t)w * h);
    std::vector<uint8_t> row((size_t)w + 6);
    for (int y = 0; y < h; ++y) {
        const uint8_t* r = src + (size_t)y * sstride;
        for (int x = -3; x < w + 3; ++x) row[x + 3] = r[reflect101(x, w)];
        uint16_t* o = &hbuf[(size_t)y * w];
        const uint8_t* p = row.data();
        for (int x = 0; x < w; ++x)
            o[x] = (uint16_t)(18 * (p[x] + p[x + 6]) + 34 * (p[x + 1] + p[x + 5]) + 48 * (p[x + 2] + p[x + 4]) + 56 * p[x + 3]);
    }
    for (int y = 0; y < h; ++y) {
        const uint16_t* r[7];
        for (int t = -3; t <= 3; ++t) r[t + 3] = &hbuf[(size_t)reflect101(y + t, h) * w];
        uint8_t* o = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x) {
            const uint32_t acc = 18u * ((uint32_t)r[0][x] + r[6][x]) + 34u * ((uint32_t)r[1][x] + r[5][x]) +
                                 48u * ((uint32_t)r[2][x] + r[4][x]) + 56u * (uint32_t)r[3][x];
            o[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// cv::FAST(img, kps, T, true) type 9_16 -- SURVEY App. A.3 (ORBextractor.cc:1135,1144)
// ---------------------------------------------------------------------------------------------
static const int kRingX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kRingY[16] = {-3, -3, -2, -1, 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3};

// score = max over the 16 arcs of 9 of max(min d, -max d) - 1, eight arcs at a time (SSE2, int16), the way OpenCV's cornerScore<16>
// arranges it: with a = min(d[k+1 .. k+8]) the arcs starting at k and at k + 1 are min(a, d[k]) and min(a, d[k+9]).  Integer min / max
// only, so the value is exactly fast_score16_scalar's (tests compare them on random rings).
int fast_score16(const uint8_t* p, int stride) {
    alignas(16) short d[32];
    const int c = p[0];
    for (int k = 0; k < 25; ++k) d[k] = (short)(c - p[kRingY[k & 15] * stride + kRingX[k & 15]]);
    __m128i q0 = _mm_set1_epi16(-1000), q1 = _mm_set1_epi16(1000);
    for (int k = 0; k < 16; k += 8) {
        __m128i v0 = _mm_loadu_si128((const __m128i*)(d + k + 1)), v1 = _mm_loadu_si128((const __m128i*)(d + k + 2));
        __m128i a = _mm_min_epi16(v0, v1), b = _mm_max_epi16(v0, v1);
        for (int j = 3; j <= 8; ++j) {
            v0 = _mm_loadu_si128((const __m128i*)(d + k + j));
            a = _mm_min_epi16(a, v0);
            b = _mm_max_epi16(b, v0);
        }
        v0 = _mm_loadu_si128((const __m128i*)(d + k));
        q0 = _mm_max_epi16(q0, _mm_min_epi16(a, v0));
        q1 = _mm_min_epi16(q1, _mm_max_epi16(b, v0));
        v0 = _mm_loadu_si128((const __m128i*)(d + k + 9));
        q0 = _mm_max_epi16(q0, _mm_min_epi16(a, v0));
        q1 = _mm_min_epi16(q1, _mm_max_epi16(b, v0));
    }
    q0 = _mm_max_epi16(q0, _mm_sub_epi16(_mm_setzero_si128(), q1));
    q0 = _mm_max_epi16(q0, _mm_srli_si128(q0, 8));
    q0 = _mm_max_epi16(q0, _mm_srli_si128(q0, 4));
    q0 = _mm_max_epi16(q0, _mm_srli_si128(q0, 2));
    return (short)_mm_cvtsi128_si32(q0) - 1;
}

int fast_score16_scalar(const uint8_t* p, int stride) {
    int d[16 + 9];
    const int c = p[0];
    for (int k = 0; k < 16; ++k) d[k] = c - p[kRingY[k] * stride + kRingX[k]];
    for (int k = 0; k < 9; ++k) d[16 + k] = d[k];
    // min / max over every window of 9 consecutive ring pixels via windows of 3
    int lo3[16 + 6], hi3[16 + 6];
    for (int k = 0; k < 22; ++k) {
        lo3[k] = std::min(d[k], std::min(d[k + 1], d[k + 2]));
        hi3[k] = std::max(d[k], std::max(d[k + 1], d[k + 2]));
    }
    int best = -256;
    for (int s = 0; s < 16; ++s) {
        const int mn = std::min(lo3[s], std::min(lo3[s + 3], lo3[s + 6]));
        const int mx = std::max(hi3[s], std::max(hi3[s + 3], hi3[s + 6]));
        best = std::max(best, std::max(mn, -mx));
    }
    return best - 1;
}

// 9 contiguous set bits in a circular 16-bit mask
static inline bool arc9(uint32_t m) {
    m |= m << 16;
    uint32_t t = m & (m >> 1);
    t &= t >> 2;
    t &= t >> 4;
    t &= m >> 8;
    return (t & 0xffffu) != 0;
}

// One cv::FAST call on a cell window (cw x ch, top-left at win): appends row-major keypoints
// (window coords) whose score >= thr and that are strict 3x3 local maxima among tested pixels.
// A pixel is a corner at threshold thr iff 9 contiguous ring pixels are all > c+thr or all < c-thr
// (cheap bit test first, like OpenCV's own pre-test); only corners get the exact score.
void fast_cell(const uint8_t* win, int cw, int ch, int stride, int thr, std::vector<Cand>& out) {
    if (cw < 7 || ch < 7) return;
    std::vector<int> sc((size_t)cw * ch, 0);
    int off[16];
    for (int k = 0; k < 16; ++k) off[k] = kRingY[k] * stride + kRingX[k];
    for (int y = 3; y < ch - 3; ++y)
        for (int x = 3; x < cw - 3; ++x) {
            const uint8_t* p = win + (size_t)y * stride + x;
            const int c = p[0], hi = c + thr, lo = c - thr;
            // antipodal pre-test: an arc of 9 contains pixel k or pixel k+8 for every k, so a bright (dark) arc needs one bright
            // (dark) pixel in EVERY antipodal pair; four pairs, tested progressively like OpenCV's own high-speed test
            {
                const int a0 = p[off[0]], b0 = p[off[8]];
                int br = (a0 > hi) | (b0 > hi), dk = (a0 < lo) | (b0 < lo);
                if (!(br | dk)) continue;
                const int a1 = p[off[4]], b1 = p[off[12]];
                br &= (a1 > hi) | (b1 > hi); dk &= (a1 < lo) | (b1 < lo);
                if (!(br | dk)) continue;
                const int a2 = p[off[2]], b2 = p[off[10]];
                br &= (a2 > hi) | (b2 > hi); dk &= (a2 < lo) | (b2 < lo);
                if (!(br | dk)) continue;
                const int a3 = p[off[6]], b3 = p[off[14]];
                br &= (a3 > hi) | (b3 > hi); dk &= (a3 < lo) | (b3 < lo);
                if (!(br | dk)) continue;
            }
            uint32_t mb = 0, md = 0;
            for (int k = 0; k < 16; ++k) {
                const int v = p[off[k]];
                mb |= (uint32_t)(v > hi) << k;
                md |= (uint32_t)(v < lo) << k;
            }
            if (!arc9(mb) && !arc9(md)) continue;
            sc[(size_t)y * cw + x] = fast_score16(p, stride);  // >= thr by construction
        }
    for (int y = 3; y < ch - 3; ++y)
        for (int x = 3; x < cw - 3; ++x) {
            const int s = sc[(size_t)y * cw + x];
            if (s == 0) continue;
            bool mx = true;
            for (int dy = -1; dy <= 1 && mx; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    if (!dx && !dy) continue;
                    if (s <= sc[(size_t)(y + dy) * cw + (x + dx)]) { mx = false; break; }
                }
            if (mx) out.push_back({x, y, s});
        }
}

// ---------------------------------------------------------------------------------------------
// cv::fastAtan2 scalar path -- SURVEY App. A.4 (ORBextractor.cc:137)
// ---------------------------------------------------------------------------------------------
float fast_atan2_deg(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s,
                p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float eps = (float)2.2204460492503131e-16;
    const float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// IC_Angle, ORBextractor.cc:91-138
float ic_angle(const Plane& im, int x, int y, const std::vector<int>& umax) {
    int m01 = 0, m10 = 0;
    const uint8_t* c = &im.px[(size_t)y * im.w + x];
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
        int vs = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            const int p = c[u + v * im.w], m = c[u - v * im.w];
            vs += p - m;
            m10 += u * (p + m);
        }
        m01 += v * vs;
    }
    return fast_atan2_deg((float)m01, (float)m10);
}

// computeOrbDescriptor, ORBextractor.cc:150-203.  cosf/sinf are glibc's float routines
// (float argument under `using namespace std`).
void orb_descriptor(const Plane& im, int x, int y, float angle_deg, uint8_t* desc) {
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float ang = angle_deg * factorPI;
    const float a = cosf(ang), b = sinf(ang);
    const uint8_t* c = &im.px[(size_t)y * im.w + x];
    const int step = im.w;
    const int8_t* p = kPattern;
    for (int i = 0; i < 32; ++i, p += 32) {
        int val = 0;
        for (int k = 0; k < 8; ++k) {
            const float xa = p[4 * k], ya = p[4 * k + 1], xb = p[4 * k + 2], yb = p[4 * k + 3];
            const int t0 = c[round_half_even(xa * b + ya * a) * step + round_half_even(xa * a - ya * b)];
            const int t1 = c[round_half_even(xb * b + yb * a) * step + round_half_even(xb * a - yb * b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

// ORBmatcher::DescriptorDistance, ORBmatcher.cc:2383-2403
int descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t x, y;
        memcpy(&x, a + 4 * i, 4);
        memcpy(&y, b + 4 * i, 4);
        d += __builtin_popcount(x ^ y);
    }
    return d;
}

// ---------------------------------------------------------------------------------------------
// constructor tables, ORBextractor.cc:468-571
// ---------------------------------------------------------------------------------------------
Extractor::Extractor(int nf, float sf, int nl, int ini, int mn)
    : nfeatures(nf), nlevels(nl), iniThFAST(ini), minThFAST(mn), scaleFactor(sf) {
    mvScaleFactor.assign(nl, 1.f);
    mvLevelSigma2.assign(nl, 1.f);
    for (int i = 1; i < nl; ++i) {
        mvScaleFactor[i] = (float)(mvScaleFactor[i - 1] * (double)sf);  // `double scaleFactor` member (ORBextractor.h:99)
        mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
    }
    mvInvScaleFactor.resize(nl);
    mvInvLevelSigma2.resize(nl);
    for (int i = 0; i < nl; ++i) {
        mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i];
        mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i];
    }
    mnFeaturesPerLevel.resize(nl);
    const float factor = (float)(1.0f / (double)sf);
    float per = nf * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
    int sum = 0;
    for (int l = 0; l < nl - 1; ++l) {
        mnFeaturesPerLevel[l] = round_half_even(per);
        sum += mnFeaturesPerLevel[l];
        per *= factor;
    }
    mnFeaturesPerLevel[nl - 1] = std::max(nf - sum, 0);

    umax.assign(kHalfPatch + 1, 0);
    const int vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    const int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= vmax; ++v) umax[v] = round_half_even(std::sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
    pyramid.resize(nl);
    blurred.resize(nl);
    cands.resize(nl);
    lvl.resize(nl);
}

// ComputePyramid, ORBextractor.cc:1687-1738 (border omitted: dead weight for results, App. A.6)
void Extractor::compute_pyramid(const uint8_t* img, int w, int h, int stride) {
    for (int l = 0; l < nlevels; ++l) {
        const float s = mvInvScaleFactor[l];
        Plane& P = pyramid[l];
        P.w = round_half_even((float)w * s);
        P.h = round_half_even((float)h * s);
        P.px.resize((size_t)P.w * P.h);
        if (l == 0) {
            for (int y = 0; y < h; ++y) memcpy(&P.px[(size_t)y * w], img + (size_t)y * stride, w);
        } else {
            const Plane& Q = pyramid[l - 1];
            resize_linear_u8(Q.px.data(), Q.w, Q.h, Q.w, P.px.data(), P.w, P.h, P.w);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The corner TEST of cv::FAST for 16 pixels at a time (SSE2), the way OpenCV's own FAST_t<16> does it: a pixel is a corner at
// threshold T iff 9 consecutive ring pixels are all > c + T or all < c - T.  Run once over the part of a level that the cells'
// tested areas tile, it gives a score map -- exact score where the pixel is a corner at T, 0 elsewhere -- from
// which every cell's cv::FAST(T, nms = true) result follows (a cell's NMS ignores pixels outside its own tested area).  Purely a
// faster route to the same numbers: detect_level(simd = false) keeps the cell-by-cell scalar path and tests/ compares the two; it
// also makes the timed CPU baseline a fairer stand-in for OpenCV's SIMD FAST.
// ---------------------------------------------------------------------------------------------
static int g_fast_simd = 1;
void set_fast_simd(int on) { g_fast_simd = on; }

static void fast_score_map(const Plane& P, int x0, int x1, int y0, int y1, int thr, std::vector<uint8_t>& S) {   // tested pixels [x0, x1) x [y0, y1)
    S.assign((size_t)P.w * P.h, 0);
    int off[25];
    for (int k = 0; k < 25; ++k) off[k] = kRingY[k & 15] * P.w + kRingX[k & 15];
    const __m128i delta = _mm_set1_epi8((char)-128), t = _mm_set1_epi8((char)thr), K8 = _mm_set1_epi8(8);
    for (int y = y0; y < y1; ++y) {
        const uint8_t* row = &P.px[(size_t)y * P.w];
        for (int xb = x0; xb < x1; xb += 16) {
            const int x = std::min(xb, x1 - 16);                 // the last block overlaps the previous one
            const uint8_t* p = row + x;
            __m128i c = _mm_loadu_si128((const __m128i*)p);
            const __m128i v0 = _mm_xor_si128(_mm_adds_epu8(c, t), delta);   // c + T  (saturating), signed domain
            const __m128i v1 = _mm_xor_si128(_mm_subs_epu8(c, t), delta);   // c - T
            const __m128i a0 = _mm_xor_si128(_mm_loadu_si128((const __m128i*)(p + off[0])), delta);
            const __m128i a4 = _mm_xor_si128(_mm_loadu_si128((const __m128i*)(p + off[4])), delta);
            const __m128i a8 = _mm_xor_si128(_mm_loadu_si128((const __m128i*)(p + off[8])), delta);
            const __m128i a12 = _mm_xor_si128(_mm_loadu_si128((const __m128i*)(p + off[12])), delta);
            // an arc of 9 contains two neighbouring compass points
            __m128i m0 = _mm_and_si128(_mm_cmpgt_epi8(a0, v0), _mm_cmpgt_epi8(a4, v0));
            __m128i m1 = _mm_and_si128(_mm_cmpgt_epi8(v1, a0), _mm_cmpgt_epi8(v1, a4));
            m0 = _mm_or_si128(m0, _mm_and_si128(_mm_cmpgt_epi8(a4, v0), _mm_cmpgt_epi8(a8, v0)));
            m1 = _mm_or_si128(m1, _mm_and_si128(_mm_cmpgt_epi8(v1, a4), _mm_cmpgt_epi8(v1, a8)));
            m0 = _mm_or_si128(m0, _mm_and_si128(_mm_cmpgt_epi8(a8, v0), _mm_cmpgt_epi8(a12, v0)));
            m1 = _mm_or_si128(m1, _mm_and_si128(_mm_cmpgt_epi8(v1, a8), _mm_cmpgt_epi8(v1, a12)));
            m0 = _mm_or_si128(m0, _mm_and_si128(_mm_cmpgt_epi8(a12, v0), _mm_cmpgt_epi8(a0, v0)));
            m1 = _mm_or_si128(m1, _mm_and_si128(_mm_cmpgt_epi8(v1, a12), _mm_cmpgt_epi8(v1, a0)));
            if (_mm_movemask_epi8(_mm_or_si128(m0, m1)) == 0) continue;
            __m128i c0 = _mm_setzero_si128(), c1 = c0, max0 = c0, max1 = c0;
            for (int k = 0; k < 25; ++k) {
                const __m128i r = _mm_xor_si128(_mm_loadu_si128((const __m128i*)(p + off[k])), delta);
                m0 = _mm_cmpgt_epi8(r, v0);
                m1 = _mm_cmpgt_epi8(v1, r);
                c0 = _mm_and_si128(_mm_sub_epi8(c0, m0), m0);    // run length of "brighter", reset where the test fails
                c1 = _mm_and_si128(_mm_sub_epi8(c1, m1), m1);
                max0 = _mm_max_epu8(max0, c0);
                max1 = _mm_max_epu8(max1, c1);
            }
            int mask = _mm_movemask_epi8(_mm_cmpgt_epi8(_mm_max_epu8(max0, max1), K8));     // a run of 9 or more
            while (mask) {
                const int b = __builtin_ctz(mask);
                mask &= mask - 1;
                S[(size_t)y * P.w + x + b] = (uint8_t)fast_score16(p + b, P.w);              // >= thr by construction, <= 254
            }
        }
    }
}

// per-level cell loop, ORBextractor.cc:1069-1166.  Output coords are relative to minBorder.
void Extractor::detect_level(int level, std::vector<Cand>& out) const {
    const Plane& P = pyramid[level];
    out.clear();
    const int minBX = kEdge - 3, minBY = minBX;
    const int maxBX = P.w - kEdge + 3, maxBY = P.h - kEdge + 3;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / 35.f), nRows = (int)(height / 35.f);
    if (nCols < 1 || nRows < 1) return;
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    std::vector<Cand> cell;
    // tested pixels of all cells tile [minBX + 3, maxBX - 3) x [minBY + 3, maxBY - 3)
    const bool simd = g_fast_simd && (maxBX - minBX - 6) >= 16;
    std::vector<uint8_t> S;
    if (simd) fast_score_map(P, minBX + 3, maxBX - 3, minBY + 3, maxBY - 3, iniThFAST, S);
    for (int i = 0; i < nRows; ++i) {
        const float iniY = (float)(minBY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBY - 3) continue;
        if (maxY > maxBY) maxY = (float)maxBY;
        for (int j = 0; j < nCols; ++j) {
            const float iniX = (float)(minBX + j * wCell);
            float maxX = iniX + wCell + 6;
            if (iniX >= maxBX - 6) continue;
            if (maxX > maxBX) maxX = (float)maxBX;
            const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
            const uint8_t* win = &P.px[(size_t)y0 * P.w + x0];
            cell.clear();
            if (simd && cw >= 7 && ch >= 7) {
                // cv::FAST(iniThFAST, nms) of this window from the level's score map: strict 3 x 3 maximum among the window's own
                // tested pixels (columns [3, cw - 3), rows [3, ch - 3)); everything outside counts as 0
                const uint8_t* sm = &S[(size_t)y0 * P.w + x0];
                for (int y = 3; y < ch - 3; ++y)
                    for (int x = 3; x < cw - 3; ++x) {
                        const int sc0 = sm[(size_t)y * P.w + x];
                        if (sc0 == 0) continue;
                        bool mx = true;
                        for (int dy = -1; dy <= 1 && mx; ++dy)
                            for (int dx = -1; dx <= 1; ++dx) {
                                if (!dx && !dy) continue;
                                const int yy = y + dy, xx = x + dx;
                                const int nb = (yy >= 3 && yy < ch - 3 && xx >= 3 && xx < cw - 3) ? sm[(size_t)yy * P.w + xx] : 0;
                                if (sc0 <= nb) { mx = false; break; }
                            }
                        if (mx) cell.push_back({x, y, sc0});
                    }
            } else {
                fast_cell(win, cw, ch, P.w, iniThFAST, cell);
            }
            if (cell.empty()) fast_cell(win, cw, ch, P.w, minThFAST, cell);
            for (const Cand& c : cell) out.push_back({c.x + j * wCell, c.y + i * hCell, c.score});
        }
    }
}

// ---------------------------------------------------------------------------------------------
// DistributeOctTree, ORBextractor.cc:711-1057 (+ DivideNode :602-674, compareNodes :676-697).
// Restated with index lists instead of KeyPoint copies; list order, push_front order, the early
// break and libstdc++ std::sort tie behaviour are the reference's.
// ---------------------------------------------------------------------------------------------
namespace {
struct Node {
    int ulx, uly, urx, bry;  // UL=(ulx,uly) UR=(urx,uly) BL=(ulx,bry) BR=(urx,bry)
    std::vector<int> keys;   // candidate indices, original relative order
    bool no_more = false;
    std::list<Node>::iterator self;
};

void divide(const Node& p, const std::vector<Cand>& c, Node out[4]) {
    const int halfX = (int)std::ceil((float)(p.urx - p.ulx) / 2);
    const int halfY = (int)std::ceil((float)(p.bry - p.uly) / 2);
    const int mx = p.ulx + halfX, my = p.uly + halfY;
    out[0] = Node{p.ulx, p.uly, mx, my, {}, false, {}};
    out[1] = Node{mx, p.uly, p.urx, my, {}, false, {}};
    out[2] = Node{p.ulx, my, mx, p.bry, {}, false, {}};
    out[3] = Node{mx, my, p.urx, p.bry, {}, false, {}};
    for (int id : p.keys) {
        const float x = (float)c[id].x, y = (float)c[id].y;
        const int q = (x < (float)mx) ? ((y < (float)my) ? 0 : 2) : ((y < (float)my) ? 1 : 3);
        out[q].keys.push_back(id);
    }
    for (int q = 0; q < 4; ++q)
        if (out[q].keys.size() == 1) out[q].no_more = true;
}

bool node_less(std::pair<int, Node*>& a, std::pair<int, Node*>& b) {
    if (a.first < b.first) return true;
    if (a.first > b.first) return false;
    return a.second->ulx < b.second->ulx;
}
}  // namespace

std::vector<int> Extractor::distribute(const std::vector<Cand>& c, int minX, int maxX, int minY, int maxY,
                                       int N) const {
    std::vector<int> result;
    const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
    if (nIni < 1) return result;  // reference would divide by zero
    const float hX = (float)(maxX - minX) / nIni;
    std::list<Node> nodes;
    std::vector<Node*> roots(nIni);
    for (int i = 0; i < nIni; ++i) {
        Node n{(int)(hX * (float)i), 0, (int)(hX * (float)(i + 1)), maxY - minY, {}, false, {}};
        nodes.push_back(n);
        roots[i] = &nodes.back();
    }
    for (int i = 0; i < (int)c.size(); ++i) roots[(int)((float)c[i].x / hX)]->keys.push_back(i);
    for (auto it = nodes.begin(); it != nodes.end();) {
        if (it->keys.size() == 1) { it->no_more = true; ++it; }
        else if (it->keys.empty()) it = nodes.erase(it);
        else ++it;
    }
    bool finish = false;
    std::vector<std::pair<int, Node*>> expandable;
    auto push_children = [&](Node ch[4], int* n_expand) {
        for (int q = 0; q < 4; ++q) {
            if (ch[q].keys.empty()) continue;
            nodes.push_front(std::move(ch[q]));
            Node& f = nodes.front();
            if (f.keys.size() > 1) {
                if (n_expand) ++*n_expand;
                expandable.emplace_back((int)f.keys.size(), &f);
                f.self = nodes.begin();
            }
        }
    };
    while (!finish) {
        const int prev = (int)nodes.size();
        int nToExpand = 0;
        expandable.clear();
        for (auto it = nodes.begin(); it != nodes.end();) {
            if (it->no_more) { ++it; continue; }
            Node ch[4];
            divide(*it, c, ch);
            push_children(ch, &nToExpand);
            it = nodes.erase(it);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prev) {
            finish = true;
        } else if ((int)nodes.size() + nToExpand * 3 > N) {
            while (!finish) {
                const int prev2 = (int)nodes.size();
                std::vector<std::pair<int, Node*>> order = expandable;
                expandable.clear();
                std::sort(order.begin(), order.end(), node_less);
                for (int j = (int)order.size() - 1; j >= 0; --j) {
                    Node ch[4];
                    divide(*order[j].second, c, ch);
                    push_children(ch, nullptr);
                    nodes.erase(order[j].second->self);
                    if ((int)nodes.size() >= N) break;
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == prev2) finish = true;
            }
        }
    }
    result.reserve(nodes.size());
    for (const Node& n : nodes) {
        int best = n.keys[0];
        for (size_t k = 1; k < n.keys.size(); ++k)
            if (c[n.keys[k]].score > c[best].score) best = n.keys[k];
        result.push_back(best);
    }
    return result;
}

// ORBextractor::operator(), ORBextractor.cc:1557-1682
int Extractor::extract(const uint8_t* img, int w, int h, int stride, int lap0, int lap1,
                       std::vector<KeyPoint>& kps, std::vector<uint8_t>& desc) {
    kps.clear();
    desc.clear();
    if (!img || w <= 0 || h <= 0) return -1;
    double t0 = now_s();
    compute_pyramid(img, w, h, stride);
    double t1 = now_s();
    t_pyr = t1 - t0;
    t_fast = t_tree = t_angle = t_blur = t_desc = 0;
    int total = 0;
    for (int l = 0; l < nlevels; ++l) {
        t0 = now_s();
        detect_level(l, cands[l]);
        t1 = now_s();
        t_fast += t1 - t0;
        const Plane& P = pyramid[l];
        const int minB = kEdge - 3;
        std::vector<int> keep = distribute(cands[l], minB, P.w - kEdge + 3, minB, P.h - kEdge + 3,
                                           mnFeaturesPerLevel[l]);
        const int scaledPatch = (int)(kPatch * mvScaleFactor[l]);
        lvl[l].clear();
        for (int id : keep) {
            const Cand& cd = cands[l][id];
            lvl[l].push_back({(float)(cd.x + minB), (float)(cd.y + minB), (float)scaledPatch, -1.f,
                              (float)cd.score, l, -1});
        }
        t0 = now_s();
        t_tree += t0 - t1;
        total += (int)lvl[l].size();
    }
    t0 = now_s();
    for (int l = 0; l < nlevels; ++l)
        for (KeyPoint& k : lvl[l]) k.angle = ic_angle(pyramid[l], (int)k.x, (int)k.y, umax);
    t_angle = now_s() - t0;

    kps.resize(total);
    desc.resize((size_t)total * 32);
    int mono = 0, stereo = total - 1;
    for (int l = 0; l < nlevels; ++l) {
        if (lvl[l].empty()) { blurred[l] = Plane(); continue; }
        t0 = now_s();
        const Plane& P = pyramid[l];
        Plane& B = blurred[l];
        B.w = P.w; B.h = P.h;
        B.px.resize(P.px.size());
        gaussian_blur7_u8(P.px.data(), P.w, P.h, P.w, B.px.data(), B.w);
        t1 = now_s();
        t_blur += t1 - t0;
        const float scale = mvScaleFactor[l];
        for (const KeyPoint& k0 : lvl[l]) {
            uint8_t d[32];
            orb_descriptor(B, (int)k0.x, (int)k0.y, k0.angle, d);
            KeyPoint k = k0;
            if (l != 0) { k.x *= scale; k.y *= scale; }
            const int dst = (k.x >= (float)lap0 && k.x <= (float)lap1) ? stereo-- : mono++;
            kps[dst] = k;
            memcpy(&desc[(size_t)dst * 32], d, 32);
        }
        t_desc += now_s() - t1;
    }
    return mono;
}

}  // namespace orb_oracle

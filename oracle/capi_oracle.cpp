// oracle/capi_oracle.cpp -- TEST INFRASTRUCTURE (see oracle.h): flat C entry points so that
// tests/ and bench.py can drive the CPU restatement through ctypes.
#include <cstring>

#include "oracle.h"

using namespace orb_oracle;

extern "C" {

void* orc_extractor_create(int nf, float sf, int nl, int ini, int mn) { return new Extractor(nf, sf, nl, ini, mn); }
void orc_extractor_destroy(void* h) { delete (Extractor*)h; }

// returns monoIndex (or <0); *n_out = number of keypoints; outputs truncated to cap
int orc_extract(void* h, const uint8_t* img, int w, int hh, int stride, int lap0, int lap1, KeyPoint* kps,
                uint8_t* desc, int cap, int* n_out) {
    Extractor* e = (Extractor*)h;
    std::vector<KeyPoint> k;
    std::vector<uint8_t> d;
    const int r = e->extract(img, w, hh, stride, lap0, lap1, k, d);
    const int n = (int)k.size();
    *n_out = n;
    const int m = n < cap ? n : cap;
    if (m > 0) {
        memcpy(kps, k.data(), sizeof(KeyPoint) * m);
        memcpy(desc, d.data(), 32 * (size_t)m);
    }
    return r;
}

void orc_tables(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* quota, int* umax16) {
    Extractor* e = (Extractor*)h;
    for (int i = 0; i < e->nlevels; ++i) {
        scale[i] = e->mvScaleFactor[i];
        inv_scale[i] = e->mvInvScaleFactor[i];
        sigma2[i] = e->mvLevelSigma2[i];
        inv_sigma2[i] = e->mvInvLevelSigma2[i];
        quota[i] = e->mnFeaturesPerLevel[i];
    }
    for (int i = 0; i < 16; ++i) umax16[i] = e->umax[i];
}

void orc_level_size(void* h, int l, int* w, int* hh) {
    Extractor* e = (Extractor*)h;
    *w = e->pyramid[l].w;
    *hh = e->pyramid[l].h;
}
void orc_level_pyramid(void* h, int l, uint8_t* dst) {
    Extractor* e = (Extractor*)h;
    memcpy(dst, e->pyramid[l].px.data(), e->pyramid[l].px.size());
}
int orc_level_blurred(void* h, int l, uint8_t* dst) {
    Extractor* e = (Extractor*)h;
    if (e->blurred[l].px.empty()) return 0;
    memcpy(dst, e->blurred[l].px.data(), e->blurred[l].px.size());
    return 1;
}
int orc_level_cands(void* h, int l, int32_t* dst, int cap) {
    Extractor* e = (Extractor*)h;
    const int n = (int)e->cands[l].size();
    for (int i = 0; i < n && i < cap; ++i) {
        dst[3 * i] = e->cands[l][i].x;
        dst[3 * i + 1] = e->cands[l][i].y;
        dst[3 * i + 2] = e->cands[l][i].score;
    }
    return n;
}
int orc_level_kps(void* h, int l, KeyPoint* dst, int cap) {
    Extractor* e = (Extractor*)h;
    const int n = (int)e->lvl[l].size();
    for (int i = 0; i < n && i < cap; ++i) dst[i] = e->lvl[l][i];
    return n;
}
void orc_timings(void* h, double* t6) {
    Extractor* e = (Extractor*)h;
    t6[0] = e->t_pyr; t6[1] = e->t_fast; t6[2] = e->t_tree; t6[3] = e->t_angle; t6[4] = e->t_blur; t6[5] = e->t_desc;
}

// ---- primitives -------------------------------------------------------------------------
void orc_resize(const uint8_t* s, int sw, int sh, uint8_t* d, int dw, int dh) { resize_linear_u8(s, sw, sh, sw, d, dw, dh, dw); }
void orc_blur(const uint8_t* s, int w, int h, uint8_t* d) { gaussian_blur7_u8(s, w, h, w, d, w); }
int orc_fast_cell(const uint8_t* win, int cw, int ch, int stride, int thr, int32_t* out, int cap) {
    std::vector<Cand> c;
    fast_cell(win, cw, ch, stride, thr, c);
    for (int i = 0; i < (int)c.size() && i < cap; ++i) { out[3 * i] = c[i].x; out[3 * i + 1] = c[i].y; out[3 * i + 2] = c[i].score; }
    return (int)c.size();
}
float orc_atan2(float y, float x) { return fast_atan2_deg(y, x); }
void orc_set_fast_simd(int on) { set_fast_simd(on); }
// both score routines on a 7 x 7 patch (row-major, centre at [3][3]): out2 = {SSE2, scalar}
void orc_fast_scores(const uint8_t* patch49, int* out2) { out2[0] = fast_score16(patch49 + 3 * 7 + 3, 7); out2[1] = fast_score16_scalar(patch49 + 3 * 7 + 3, 7); }
float orc_ic_angle(const uint8_t* img, int w, int h, int x, int y, const int* umax16) {
    Plane P; P.w = w; P.h = h; P.px.assign(img, img + (size_t)w * h);
    std::vector<int> u(umax16, umax16 + 16);
    return ic_angle(P, x, y, u);
}
void orc_descriptor(const uint8_t* img, int w, int h, int x, int y, float angle, uint8_t* d32) {
    Plane P; P.w = w; P.h = h; P.px.assign(img, img + (size_t)w * h);
    orb_descriptor(P, x, y, angle, d32);
}
int orc_hamming(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }
const int8_t* orc_pattern() { return kPattern; }
int orc_distribute(const int32_t* cand3, int n, int minX, int maxX, int minY, int maxY, int N, int32_t* out_idx, int cap) {
    std::vector<Cand> c(n);
    for (int i = 0; i < n; ++i) c[i] = {cand3[3 * i], cand3[3 * i + 1], cand3[3 * i + 2]};
    Extractor e(1000, 1.2f, 8, 20, 7);
    std::vector<int> r = e.distribute(c, minX, maxX, minY, maxY, N);
    for (int i = 0; i < (int)r.size() && i < cap; ++i) out_idx[i] = r[i];
    return (int)r.size();
}
}

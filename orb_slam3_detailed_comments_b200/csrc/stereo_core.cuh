// stereo_core.cuh -- Frame::ComputeStereoMatches (/root/reference/src/Frame.cc:1102-1358) for ONE left keypoint, written so that the
// identical source runs as a device function (one thread per left keypoint in k_stereo_match_v1) and on the host
// (tests/host_emul, against the oracle).  Differences from k_stereo_match (round 1, one warp per left keypoint scanning ALL
// right keypoints): the right keypoints are bucketed by image row once per CTA (the reference's vRowIndices, Frame.cc:1134-1156,
// reduced to one bucket per keypoint plus a +-window scan), so a left keypoint looks at ~4 % of them; the 11 x 11 SAD over 11
// shifts is done by the same thread.  Arithmetic, gates and tie rule are k_stereo_match's: the best candidate is the minimum of
// (distance << 16 | right index), i.e. "first strictly smaller" in ascending right index whatever the scan order.
#pragma once
#include "devmath.cuh"

namespace orbdev {

struct StLevel {            // raw (unblurred) pyramid level of this stereo pair
    const uint8_t* L;
    const uint8_t* R;
    int pitchL, pitchR, wR;
    float scale, inv_scale;
};

struct StRight {            // right keypoints of this pair, staged (Frame.cc:1134-1156)
    const float* x;         // kpR.pt.x
    const int* band;        // minr (low 16, signed) | maxr << 16
    const signed char* oct;
    const int* row_off;     // [H + 1] bucket b = floor(kpR.pt.y) clamped to [0, H)
    const int* row_ent;     // right indices by bucket (any order inside a bucket)
    int H, W;               // rows, scan half-window: every candidate of row v lies in buckets [v - W, v + W]
};

ORB_HD int st_popc(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

ORB_HD int st_row_bucket(float y, int H) {
    int b = (int)floorf(y);
    return b < 0 ? 0 : (b >= H ? H - 1 : b);
}

// half-window in rows that covers minr..maxr of every right keypoint: r = 2 * scale[nlevels - 1], maxr = ceil(y + r), minr = floor(y - r)
ORB_HD int st_scan_window(float maxScale) { return (int)ceilf(fmul(2.0f, maxScale)) + 1; }

ORB_HD void stereo_match_one(float uL, float vL, int octL, const uint32_t* dl /* 8 words */, const StRight& R, const uint8_t* rdesc,
                             const StLevel* lv, float bf, float b, float* out_u, float* out_d, int* out_sad) {
    *out_u = -1.0f; *out_d = -1.0f; *out_sad = -1;
    const float minZ = b, maxD = fdiv(bf, minZ);
    const int vLi = (int)vL;
    const float minU = fsub(uL, maxD), maxU = uL;
    if (maxU < 0) return;
    uint32_t a[8];
    for (int k = 0; k < 8; ++k) a[k] = dl[k];
    uint32_t best = (100u << 16) | 0xffffu;   // TH_HIGH, ORBmatcher.cc:36
    int r0 = vLi - R.W, r1 = vLi + R.W;
    if (r0 < 0) r0 = 0;
    if (r1 > R.H - 1) r1 = R.H - 1;
    for (int n = (r0 <= r1 ? R.row_off[r0] : 0), ne = (r0 <= r1 ? R.row_off[r1 + 1] : 0); n < ne; ++n) {
        const int j = R.row_ent[n];
        const int band = R.band[j];
        const int minr = (int)(short)(band & 0xffff), maxr = band >> 16;
        if (vLi < minr || vLi > maxr) continue;
        const int oc = R.oct[j];
        if (oc < octL - 1 || oc > octL + 1) continue;
        const float uR = R.x[j];
        if (!(uR >= minU && uR <= maxU)) continue;
        const uint32_t* dr = reinterpret_cast<const uint32_t*>(rdesc + (size_t)j * 32);
        uint32_t d = 0;
        for (int k = 0; k < 8; ++k) d += (uint32_t)st_popc(a[k] ^ dr[k]);
        const uint32_t key = (d << 16) | (uint32_t)j;
        if (key < best) best = key;
    }
    const int bestDist = (int)(best >> 16), bestIdxR = (int)(best & 0xffffu);
    if (!(bestDist < 75)) return;             // thOrbDist = (TH_HIGH + TH_LOW) / 2
    const float uR0 = R.x[bestIdxR];
    const StLevel& G = lv[octL];
    const float sf = G.inv_scale;
    const float scaleduL = roundf(fmul(uL, sf)), scaledvL = roundf(fmul(vL, sf)), scaleduR0 = roundf(fmul(uR0, sf));
    const float iniu = scaleduR0, endu = fadd(scaleduR0, 11.0f);   // scaleduR0 + L - w, scaleduR0 + L + w + 1  (w = 5, L = 5)
    if (iniu < 0 || endu >= (float)G.wR) return;
    const int cy = (int)scaledvL, cxL = (int)scaleduL, cxR = (int)scaleduR0;
    int dist[11];
    for (int s = 0; s < 11; ++s) dist[s] = 0;
    for (int row = 0; row < 11; ++row) {
        const int y = cy + row - 5;
        const uint8_t* pl = G.L + (size_t)y * G.pitchL + (cxL - 5);
        const uint8_t* pr = G.R + (size_t)y * G.pitchR + (cxR - 10);
        int l[11], r[21];
        for (int k = 0; k < 11; ++k) l[k] = pl[k];
        for (int k = 0; k < 21; ++k) r[k] = pr[k];
        for (int s = 0; s < 11; ++s) {
            int acc = 0;
            for (int k = 0; k < 11; ++k) { const int df = l[k] - r[s + k]; acc += df < 0 ? -df : df; }
            dist[s] += acc;
        }
    }
    int bestSad = 0x7fffffff, bestinc = 0;
    for (int s = 0; s < 11; ++s)
        if (dist[s] < bestSad) { bestSad = dist[s]; bestinc = s - 5; }
    if (bestinc == -5 || bestinc == 5) return;
    const float d1 = (float)dist[5 + bestinc - 1], d2 = (float)dist[5 + bestinc], d3 = (float)dist[5 + bestinc + 1];
    const float deltaR = fdiv(fsub(d1, d3), fmul(2.0f, fsub(fadd(d1, d3), fmul(2.0f, d2))));
    if (deltaR < -1 || deltaR > 1) return;
    float bestuR = fmul(G.scale, fadd(fadd(scaleduR0, (float)bestinc), deltaR));
    float disparity = fsub(uL, bestuR);
    if (!(disparity >= 0 && disparity < maxD)) return;
    if (disparity <= 0) {
        disparity = 0.01f;
        bestuR = (float)((double)uL - 0.01);
    }
    *out_d = fdiv(bf, disparity);
    *out_u = bestuR;
    *out_sad = bestSad;
}

}  // namespace orbdev

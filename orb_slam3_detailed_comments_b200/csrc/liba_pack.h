// liba_pack.h -- host-side flattening of a liba_problem into one blob laid out for liba_core.cuh.  Pure host C++ so that the
// CUDA wrapper (liba.cu) and the CPU emulation harness (tests/host_emul) share it.
//   [io]   state, point, err, lerr, out_scalars, dpos            host -> device before, device -> host after
//   [in]   obs, invs2, links, pidx, ekf, emp, pt_off, pt_edge
//   [work] saved state, normal equations, Schur scratch    zero-filled
#pragma once
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "../../include/orbslam3_b200.h"
#include "liba_core.cuh"

namespace orb {

struct LibaLayout { size_t io_bytes = 0, in_bytes = 0, work_bytes = 0, total = 0; };

static_assert(sizeof(liba_link) == 1080 && sizeof(LibaLink) == 1080, "liba_link layout");

// host == nullptr: sizes only.  Pointers inside *dev are relative to devBase (== host for the CPU emulation).
// Returns total == 0 when two links join the same keyframe pair (the gather in liba_build assumes at most one; the reference
// creates one EdgeInertial per keyframe and its mPrevKF, Optimizer.cc:2430-2500).
inline LibaLayout liba_pack(const liba_problem& p, uint8_t* host, uint8_t* devBase, LibaDev* dev) {
    const int nKF = p.n_kf, nMP = p.n_mp, nE = p.n_edges, nL = p.n_links;
    std::vector<int32_t> pidx(nKF);
    int nFree = 0;
    for (int k = 0; k < nKF; ++k) pidx[k] = p.fixed[k] ? -1 : nFree++;
    const size_t sp = 15 * (size_t)nFree, sl = 3 * (size_t)nMP;

    // CSRs: edges by map point, edges by keyframe, links by keyframe
    auto csr = [](int nRows, int n, auto rowOf, std::vector<int32_t>& off, std::vector<int32_t>& ent, auto entOf) {
        off.assign(nRows + 1, 0);
        for (int i = 0; i < n; ++i) ++off[rowOf(i) + 1];
        for (int r = 0; r < nRows; ++r) off[r + 1] += off[r];
        ent.resize(n ? n : 1);
        std::vector<int32_t> cur(off.begin(), off.end() - 1);
        for (int i = 0; i < n; ++i) ent[cur[rowOf(i)]++] = entOf(i);
    };
    std::vector<int32_t> ptOff, ptEdge, kfOff, kfEdge, klOff, klEnt;
    csr(nMP, nE, [&](int e) { return p.edge_mp[e]; }, ptOff, ptEdge, [](int e) { return e; });
    csr(nKF, nE, [&](int e) { return p.edge_kf[e]; }, kfOff, kfEdge, [](int e) { return e; });
    csr(nKF, 2 * nL, [&](int i) { return (i & 1) ? p.links[i >> 1].k2 : p.links[i >> 1].k1; }, klOff, klEnt, [](int i) { return i; });
    LibaLayout lay;
    for (int k = 0; k < nKF; ++k)          // one link per keyframe pair
        for (int a = klOff[k]; a < klOff[k + 1]; ++a)
            for (int b = a + 1; b < klOff[k + 1]; ++b) {
                const liba_link &A = p.links[klEnt[a] >> 1], &B = p.links[klEnt[b] >> 1];
                if ((A.k1 == B.k1 && A.k2 == B.k2) || (A.k1 == B.k2 && A.k2 == B.k1)) return lay;
            }
    if (nFree > 0xffff) return lay;

    // co-observation lists per upper block pair (p1 <= p2), ordered by map point
    std::vector<int32_t> pairCount((size_t)nFree * nFree, 0), pairP, pairOff(1, 0), coE1, coE2;
    auto forEachCo = [&](auto fn) {
        for (int l = 0; l < nMP; ++l)
            for (int a = ptOff[l]; a < ptOff[l + 1]; ++a)
                for (int b = ptOff[l]; b < ptOff[l + 1]; ++b) {
                    const int pa = pidx[p.edge_kf[ptEdge[a]]], pb = pidx[p.edge_kf[ptEdge[b]]];
                    if (pa >= 0 && pb >= 0 && pa <= pb) fn(pa, pb, ptEdge[a], ptEdge[b]);
                }
    };
    forEachCo([&](int pa, int pb, int, int) { ++pairCount[(size_t)pa * nFree + pb]; });
    std::vector<int32_t> pairIdx((size_t)nFree * nFree, -1);
    for (int a = 0; a < nFree; ++a)
        for (int b = a; b < nFree; ++b)
            if (pairCount[(size_t)a * nFree + b]) {
                pairIdx[(size_t)a * nFree + b] = (int32_t)pairP.size();
                pairP.push_back(a << 16 | b);
                pairOff.push_back(pairOff.back() + pairCount[(size_t)a * nFree + b]);
            }
    const int nPairs = (int)pairP.size();
    const size_t nCo = (size_t)pairOff.back();
    coE1.resize(nCo ? nCo : 1); coE2.resize(nCo ? nCo : 1);
    {
        std::vector<int32_t> cur(pairOff.begin(), pairOff.end() - 1);
        forEachCo([&](int pa, int pb, int e1, int e2) { const int q = pairIdx[(size_t)pa * nFree + pb]; coE1[cur[q]] = e1; coE2[cur[q]++] = e2; });
    }

    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
    // io
    const size_t oState = take(8 * 21 * (size_t)nKF), oPoint = take(8 * sl), oErr = take(8 * (size_t)nE), oLerr = take(8 * 3 * (size_t)nL), oOut = take(8 * 8), oDpos = take((size_t)nE);
    lay.io_bytes = off;
    // in
    const size_t oObs = take(8 * 3 * (size_t)nE), oInv = take(8 * (size_t)nE), oLinks = take(sizeof(LibaLink) * (size_t)nL);
    const size_t oPidx = take(4 * (size_t)nKF), oEkf = take(4 * (size_t)nE), oEmp = take(4 * (size_t)nE), oPtOff = take(4 * ((size_t)nMP + 1)), oPtEdge = take(4 * (size_t)nE),
                 oKfOff = take(4 * ((size_t)nKF + 1)), oKfEdge = take(4 * (size_t)nE), oKlOff = take(4 * ((size_t)nKF + 1)), oKlEnt = take(4 * 2 * (size_t)nL),
                 oPairP = take(4 * (size_t)nPairs), oPairOff = take(4 * ((size_t)nPairs + 1)), oCo1 = take(4 * nCo), oCo2 = take(4 * nCo);
    lay.in_bytes = off - lay.io_bytes;
    // work
    const size_t oStateS = take(8 * 21 * (size_t)nKF), oPointS = take(8 * sl), oHpp = take(8 * sp * sp), oHs = take(8 * sp * sp), oB = take(8 * (sp + sl)), oBs = take(8 * sp),
                 oX = take(8 * (sp + sl)), oY = take(8 * sp), oHll = take(8 * 9 * (size_t)nMP), oDinv = take(8 * 9 * (size_t)nMP), oW = take(8 * 18 * (size_t)nE),
                 oWD = take(8 * 18 * (size_t)nE), oWdb = take(8 * 6 * (size_t)nE), oEpp = take(8 * 27 * (size_t)nE), oLblk = take(8 * 930 * (size_t)nL), oFlag = take(16), oPart = take(8 * 16),
                 oChunk = take(8 * (size_t)LIBA_CHUNKS * std::max((size_t)27 * nKF, (size_t)36 * nPairs)), oChunkB = take(8 * (size_t)LIBA_CHUNKS * 6 * nKF), oYpan = take(8 * 15 * sp);
    lay.work_bytes = off - lay.io_bytes - lay.in_bytes;
    lay.total = off;
    if (!host) return lay;

    memset(host, 0, lay.io_bytes + lay.in_bytes);
    auto put = [&](size_t o, const void* src, size_t bytes) { if (bytes) memcpy(host + o, src, bytes); };
    put(oState, p.state, 8 * 21 * (size_t)nKF);
    put(oPoint, p.point, 8 * sl);
    put(oObs, p.obs, 8 * 3 * (size_t)nE);
    put(oInv, p.inv_sigma2, 8 * (size_t)nE);
    put(oLinks, p.links, sizeof(LibaLink) * (size_t)nL);
    put(oPidx, pidx.data(), 4 * (size_t)nKF);
    put(oEkf, p.edge_kf, 4 * (size_t)nE);
    put(oEmp, p.edge_mp, 4 * (size_t)nE);
    put(oPtOff, ptOff.data(), 4 * ((size_t)nMP + 1));
    put(oPtEdge, ptEdge.data(), 4 * (size_t)nE);
    put(oKfOff, kfOff.data(), 4 * ((size_t)nKF + 1));
    put(oKfEdge, kfEdge.data(), 4 * (size_t)nE);
    put(oKlOff, klOff.data(), 4 * ((size_t)nKF + 1));
    put(oKlEnt, klEnt.data(), 4 * 2 * (size_t)nL);
    put(oPairP, pairP.data(), 4 * (size_t)nPairs);
    put(oPairOff, pairOff.data(), 4 * ((size_t)nPairs + 1));
    put(oCo1, coE1.data(), 4 * nCo);
    put(oCo2, coE2.data(), 4 * nCo);

    LibaDev& d = *dev;
    d.nKF = nKF; d.nMP = nMP; d.nE = nE; d.nL = nL; d.nFree = nFree; d.sp = (int)sp; d.nPairs = nPairs;
    auto D = [&](size_t o) { return reinterpret_cast<double*>(devBase + o); };
    auto I = [&](size_t o) { return reinterpret_cast<int*>(devBase + o); };
    d.state = D(oState); d.state_saved = D(oStateS); d.point = D(oPoint); d.point_saved = D(oPointS);
    d.pidx = I(oPidx); d.ekf = I(oEkf); d.emp = I(oEmp); d.obs = D(oObs); d.invs2 = D(oInv); d.pt_off = I(oPtOff); d.pt_edge = I(oPtEdge);
    d.kf_off = I(oKfOff); d.kf_edge = I(oKfEdge); d.kl_off = I(oKlOff); d.kl_ent = I(oKlEnt);
    d.pair_p = I(oPairP); d.pair_off = I(oPairOff); d.co_e1 = I(oCo1); d.co_e2 = I(oCo2);
    d.links = reinterpret_cast<const LibaLink*>(devBase + oLinks);
    for (int i = 0; i < 9; ++i) d.Rcb[i] = p.Tcb[i];
    for (int i = 0; i < 3; ++i) d.tcb[i] = p.Tcb[9 + i];
    m3_T(d.Rcb, d.Rbc);
    m3_vec(d.Rbc, d.tcb, d.tbc);
    for (int i = 0; i < 3; ++i) d.tbc[i] = -d.tbc[i];
    d.fx = p.fx; d.fy = p.fy; d.cx = p.cx; d.cy = p.cy; d.bf = p.bf;
    d.lambda_init = p.lambda_init;
    d.max_iters = p.max_iters;
    d.err = D(oErr); d.lerr = D(oLerr); d.Hpp = D(oHpp); d.Hs = D(oHs); d.b = D(oB); d.bs = D(oBs); d.x = D(oX); d.y = D(oY);
    d.Hll = D(oHll); d.Dinv = D(oDinv); d.W = D(oW); d.WD = D(oWD); d.Wdb = D(oWdb); d.Epp = D(oEpp); d.Lblk = D(oLblk); d.part = D(oChunk); d.part_b = D(oChunkB); d.Ypan = D(oYpan);
    d.flag = I(oFlag); d.red = nullptr; d.partials = D(oPart); d.out_scalars = D(oOut); d.dpos = devBase + oDpos;
    d.t_id = 0; d.t_stride = 1; d.l_id = 0; d.l_stride = 1; d.rank = 0; d.cs = 1;      // a team of one thread; the launcher overrides per thread
    return lay;
}

// io region -> liba_result
inline void liba_unpack(const liba_problem& p, const uint8_t* host, const LibaDev& d, const uint8_t* devBase, liba_result* r) {
    auto H = [&](const void* devPtr) { return reinterpret_cast<const double*>(host + (reinterpret_cast<const uint8_t*>(devPtr) - devBase)); };
    memcpy(r->state, H(d.state), 8 * 21 * (size_t)p.n_kf);
    memcpy(r->point, H(d.point), 8 * 3 * (size_t)p.n_mp);
    if (r->edge_chi2) memcpy(r->edge_chi2, H(d.err), 8 * (size_t)p.n_edges);
    if (r->link_chi2) memcpy(r->link_chi2, H(d.lerr), 8 * 3 * (size_t)p.n_links);
    if (r->edge_depth_positive) memcpy(r->edge_depth_positive, host + (d.dpos - devBase), (size_t)p.n_edges);
    const double* o = H(d.out_scalars);
    r->iterations = (int32_t)o[0]; r->trials = (int32_t)o[1]; r->lambda = o[2]; r->chi2 = o[3]; r->chi2_initial = o[4];
    r->chi2_last_trial = o[5];
}

}  // namespace orb

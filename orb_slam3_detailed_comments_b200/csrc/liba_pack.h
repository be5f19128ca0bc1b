// liba_pack.h -- host-side flattening of a liba_problem into one blob laid out for liba_core.cuh.  Pure host C++ so that the
// CUDA wrapper (liba.cu) and the CPU emulation harness (tests/host_emul) share it.
//   [io]   state, point, err, lerr, out_scalars            host -> device before, device -> host after
//   [in]   obs, invs2, links, pidx, ekf, emp, pt_off, pt_edge
//   [work] saved state, normal equations, Schur scratch    zero-filled
#pragma once
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../../include/orbslam3_b200.h"
#include "liba_core.cuh"

namespace orb {

struct LibaLayout { size_t io_bytes = 0, in_bytes = 0, work_bytes = 0, total = 0; };

static_assert(sizeof(liba_link) == 1080 && sizeof(LibaLink) == 1080, "liba_link layout");

// host == nullptr: sizes only.  Pointers inside *dev are relative to devBase (== host for the CPU emulation).
inline LibaLayout liba_pack(const liba_problem& p, uint8_t* host, uint8_t* devBase, LibaDev* dev) {
    const int nKF = p.n_kf, nMP = p.n_mp, nE = p.n_edges, nL = p.n_links;
    int nFree = 0;
    for (int k = 0; k < nKF; ++k) nFree += p.fixed[k] ? 0 : 1;
    const size_t sp = 15 * (size_t)nFree, sl = 3 * (size_t)nMP;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
    // io
    const size_t oState = take(8 * 21 * (size_t)nKF), oPoint = take(8 * sl), oErr = take(8 * (size_t)nE), oLerr = take(8 * 3 * (size_t)nL), oOut = take(8 * 8);
    LibaLayout lay;
    lay.io_bytes = off;
    // in
    const size_t oObs = take(8 * 3 * (size_t)nE), oInv = take(8 * (size_t)nE), oLinks = take(sizeof(LibaLink) * (size_t)nL);
    const size_t oPidx = take(4 * (size_t)nKF), oEkf = take(4 * (size_t)nE), oEmp = take(4 * (size_t)nE), oPtOff = take(4 * ((size_t)nMP + 1)), oPtEdge = take(4 * (size_t)nE);
    lay.in_bytes = off - lay.io_bytes;
    // work
    const size_t oStateS = take(8 * 21 * (size_t)nKF), oPointS = take(8 * sl), oHpp = take(8 * sp * sp), oHs = take(8 * sp * sp), oB = take(8 * (sp + sl)), oBs = take(8 * sp),
                 oX = take(8 * (sp + sl)), oY = take(8 * sp), oHll = take(8 * 9 * (size_t)nMP), oDinv = take(8 * 9 * (size_t)nMP), oW = take(8 * 18 * (size_t)nE), oFlag = take(16);
    lay.work_bytes = off - lay.io_bytes - lay.in_bytes;
    lay.total = off;
    if (!host) return lay;

    memset(host, 0, lay.io_bytes + lay.in_bytes);
    memcpy(host + oState, p.state, 8 * 21 * (size_t)nKF);
    memcpy(host + oPoint, p.point, 8 * sl);
    memcpy(host + oObs, p.obs, 8 * 3 * (size_t)nE);
    memcpy(host + oInv, p.inv_sigma2, 8 * (size_t)nE);
    memcpy(host + oLinks, p.links, sizeof(LibaLink) * (size_t)nL);
    int32_t* pidx = reinterpret_cast<int32_t*>(host + oPidx);
    for (int k = 0, n = 0; k < nKF; ++k) pidx[k] = p.fixed[k] ? -1 : n++;
    memcpy(host + oEkf, p.edge_kf, 4 * (size_t)nE);
    memcpy(host + oEmp, p.edge_mp, 4 * (size_t)nE);
    int32_t* ptOff = reinterpret_cast<int32_t*>(host + oPtOff);
    int32_t* ptEdge = reinterpret_cast<int32_t*>(host + oPtEdge);
    for (int e = 0; e < nE; ++e) ++ptOff[p.edge_mp[e] + 1];          // edges by map point, ascending edge index inside a point
    for (int l = 0; l < nMP; ++l) ptOff[l + 1] += ptOff[l];
    {
        std::vector<int32_t> cur(ptOff, ptOff + nMP);
        for (int e = 0; e < nE; ++e) ptEdge[cur[p.edge_mp[e]]++] = e;
    }

    LibaDev& d = *dev;
    d.nKF = nKF; d.nMP = nMP; d.nE = nE; d.nL = nL; d.nFree = nFree; d.sp = (int)sp;
    auto D = [&](size_t o) { return reinterpret_cast<double*>(devBase + o); };
    auto I = [&](size_t o) { return reinterpret_cast<int*>(devBase + o); };
    d.state = D(oState); d.state_saved = D(oStateS); d.point = D(oPoint); d.point_saved = D(oPointS);
    d.pidx = I(oPidx); d.ekf = I(oEkf); d.emp = I(oEmp); d.obs = D(oObs); d.invs2 = D(oInv); d.pt_off = I(oPtOff); d.pt_edge = I(oPtEdge);
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
    d.Hll = D(oHll); d.Dinv = D(oDinv); d.W = D(oW); d.flag = I(oFlag); d.red = nullptr; d.out_scalars = D(oOut);
    return lay;
}

// io region -> liba_result
inline void liba_unpack(const liba_problem& p, const uint8_t* host, const LibaDev& d, const uint8_t* devBase, liba_result* r) {
    auto H = [&](const void* devPtr) { return reinterpret_cast<const double*>(host + (reinterpret_cast<const uint8_t*>(devPtr) - devBase)); };
    memcpy(r->state, H(d.state), 8 * 21 * (size_t)p.n_kf);
    memcpy(r->point, H(d.point), 8 * 3 * (size_t)p.n_mp);
    if (r->edge_chi2) memcpy(r->edge_chi2, H(d.err), 8 * (size_t)p.n_edges);
    if (r->link_chi2) memcpy(r->link_chi2, H(d.lerr), 8 * 3 * (size_t)p.n_links);
    const double* o = H(d.out_scalars);
    r->iterations = (int32_t)o[0]; r->trials = (int32_t)o[1]; r->lambda = o[2]; r->chi2 = o[3]; r->chi2_initial = o[4];
}

}  // namespace orb

// tests/host/liba_ref.cc -- TEST INFRASTRUCTURE (CPU tier): the REFERENCE's own Optimizer::LocalInertialBA, cut out of
// /root/reference/src/Optimizer.cc at build time (tests/host/build_liba_cpu.sh -> tests/host/_gen/, git-ignored) and compiled verbatim
// over the skeleton map of host/refshim and the graph stand-ins of liba_ref_graph.h.  See that header.
#include "liba_ref_graph.h"
#include "Optimizer.h"   // the reference's header (parsed against the skeletons)

#include <mutex>

using namespace std;

namespace ORB_SLAM3 {

VertexPose::VertexPose(KeyFrame* pKF) {   // VertexPose(KeyFrame*) -> ImuCamPose(KeyFrame*), G2oTypes.cc:30-70, one camera
    _kf = pKF;
    _est.twb = pKF->GetImuPosition().cast<double>();
    _est.Rwb = pKF->GetImuRotation().cast<double>();
    _est.tcw.push_back(pKF->GetTranslation().cast<double>());
    _est.Rcw.push_back(pKF->GetRotation().cast<double>());
    _est.tcb.push_back(pKF->mImuCalib.mTcb.translation().cast<double>());
    _est.Rcb.push_back(pKF->mImuCalib.mTcb.rotationMatrix().cast<double>());
    _est.bf = pKF->mbf; _est.fx = pKF->fx; _est.fy = pKF->fy; _est.cx = pKF->cx; _est.cy = pKF->cy;
    if (pKF->mpCamera2) { std::cerr << "liba_ref_graph: two-camera rigs are not modelled" << std::endl; std::abort(); }
}
VertexVelocity::VertexVelocity(KeyFrame* pKF) { _kf = pKF; _est = pKF->GetVelocity().cast<double>(); }                      // G2oTypes.cc:264-267
VertexGyroBias::VertexGyroBias(KeyFrame* pKF) : VertexVelocity(pKF) { _est = pKF->GetGyroBias().cast<double>(); }            // :279-282
VertexAccBias::VertexAccBias(KeyFrame* pKF) : VertexVelocity(pKF) { _est = pKF->GetAccBias().cast<double>(); }               // :294-297

EdgeInertial::EdgeInertial(IMU::Preintegrated* pInt) : mpInt(pInt) {   // information: G2oTypes.cc:575-586
    float C[225];
    for (int r = 0; r < 15; ++r) for (int c = 0; c < 15; ++c) C[15 * r + c] = pInt->C(r, c);
    double info[81], g[9], a[9];
    if (liba_link_information(C, 0, info, g, a) != ORB_OK) { std::cerr << "liba_ref_graph: liba_link_information failed" << std::endl; std::abort(); }
    for (int r = 0; r < 9; ++r) for (int c = 0; c < 9; ++c) _info(r, c) = info[9 * r + c];
}

double EdgeReproj::chi2() const { return _opt->_chi2[_index]; }
bool EdgeReproj::isDepthPositive() const { return _opt->_depthPos[_index] != 0; }

}  // namespace ORB_SLAM3

namespace g2o {
using namespace ORB_SLAM3;

static void die(const char* what) { std::cerr << "liba_ref_graph: " << what << std::endl; std::abort(); }

void SparseOptimizer::solve(int iterations) {
    // keyframes: VertexPose in insertion order; their velocity / bias vertices through the keyframe they were made from
    std::vector<VertexPose*> poses;
    std::vector<VertexSBAPointXYZ*> points;
    std::map<KeyFrame*, VertexVelocity*> vel, bg, ba;
    std::map<HyperGraph::Vertex*, int> idx;
    for (HyperGraph::Vertex* v : _vertices) {
        if (VertexPose* p = dynamic_cast<VertexPose*>(v)) { idx[v] = (int)poses.size(); poses.push_back(p); }
        else if (VertexSBAPointXYZ* x = dynamic_cast<VertexSBAPointXYZ*>(v)) { idx[v] = (int)points.size(); points.push_back(x); }
        else if (VertexGyroBias* g = dynamic_cast<VertexGyroBias*>(v)) bg[g->_kf] = g;
        else if (VertexAccBias* a = dynamic_cast<VertexAccBias*>(v)) ba[a->_kf] = a;
        else if (VertexVelocity* w = dynamic_cast<VertexVelocity*>(v)) vel[w->_kf] = w;
        else die("unknown vertex type");
    }
    const int nKF = (int)poses.size(), nMP = (int)points.size();
    std::vector<double> state((size_t)nKF * 21, 0.0), point((size_t)nMP * 3 + 3, 0.0);
    std::vector<uint8_t> fixed(nKF, 0);
    for (int k = 0; k < nKF; ++k) {
        const ImuCamPose& E = poses[k]->_est;
        double* S = &state[(size_t)k * 21];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) S[3 * r + c] = E.Rwb(r, c);
        for (int c = 0; c < 3; ++c) S[9 + c] = E.twb(c);
        KeyFrame* kf = poses[k]->_kf;
        fixed[k] = poses[k]->fixed() ? 1 : 0;
        if (vel.count(kf)) {
            if (!bg.count(kf) || !ba.count(kf)) die("velocity vertex without bias vertices");
            if (vel[kf]->fixed() != poses[k]->fixed() || bg[kf]->fixed() != poses[k]->fixed() || ba[kf]->fixed() != poses[k]->fixed()) die("mixed fixed flags on one keyframe");
            for (int c = 0; c < 3; ++c) { S[12 + c] = vel[kf]->_est(c); S[15 + c] = bg[kf]->_est(c); S[18 + c] = ba[kf]->_est(c); }
        } else if (!fixed[k]) die("optimisable keyframe without velocity / bias vertices");
    }
    for (int p = 0; p < nMP; ++p) for (int c = 0; c < 3; ++c) point[3 * (size_t)p + c] = points[p]->_est(c);
    // edges in creation order
    std::vector<liba_link> links;
    std::vector<EdgeInertial*> inertial;
    std::vector<EdgeGyroRW*> rwG, rwA;
    std::vector<EdgeReproj*> reproj;
    for (OptimizableGraph::Edge* e : _edges) {
        if (EdgeInertial* ei = dynamic_cast<EdgeInertial*>(e)) inertial.push_back(ei);
        else if (EdgeAccRW* ea = dynamic_cast<EdgeAccRW*>(e)) rwA.push_back(ea);
        else if (EdgeGyroRW* eg = dynamic_cast<EdgeGyroRW*>(e)) rwG.push_back(eg);
        else if (EdgeReproj* er = dynamic_cast<EdgeReproj*>(e)) reproj.push_back(er);
        else die("unknown edge type");
    }
    if (rwG.size() != inertial.size() || rwA.size() != inertial.size()) die("random-walk edges do not pair with the inertial edges");
    for (size_t i = 0; i < inertial.size(); ++i) {
        EdgeInertial* ei = inertial[i];
        if (ei->_v.size() != 6) die("EdgeInertial needs six vertices");
        VertexPose *p1 = dynamic_cast<VertexPose*>(ei->_v[0]), *p2 = dynamic_cast<VertexPose*>(ei->_v[4]);
        if (!p1 || !p2) die("EdgeInertial: vertices 0 / 4 are not poses");
        KeyFrame *k1 = p1->_kf, *k2 = p2->_kf;
        if (ei->_v[1] != vel[k1] || ei->_v[2] != bg[k1] || ei->_v[3] != ba[k1] || ei->_v[5] != vel[k2]) die("EdgeInertial: vertex roles");
        if (rwG[i]->_v.size() != 2 || rwG[i]->_v[0] != bg[k1] || rwG[i]->_v[1] != bg[k2]) die("EdgeGyroRW: vertex roles");
        if (rwA[i]->_v.size() != 2 || rwA[i]->_v[0] != ba[k1] || rwA[i]->_v[1] != ba[k2]) die("EdgeAccRW: vertex roles");
        liba_link L;
        std::memset(&L, 0, sizeof(L));
        L.k1 = idx[p1]; L.k2 = idx[p2]; L.robust = ei->_rk ? 1 : 0;
        if (ei->_rk && ei->_rk->_delta != std::sqrt(16.92)) die("EdgeInertial: Huber delta");
        const IMU::Preintegrated* P = ei->mpInt;
        L.dt = P->dT;
        const Eigen::Matrix3f* M[6] = {&P->dR, &P->JRg, &P->JVg, &P->JVa, &P->JPg, &P->JPa};
        float* D[6] = {L.dR, L.JRg, L.JVg, L.JVa, L.JPg, L.JPa};
        for (int m = 0; m < 6; ++m) for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) D[m][3 * r + c] = (*M[m])(r, c);
        for (int c = 0; c < 3; ++c) { L.dV[c] = P->dV(c); L.dP[c] = P->dP(c); }
        L.bias[0] = P->b.bax; L.bias[1] = P->b.bay; L.bias[2] = P->b.baz; L.bias[3] = P->b.bwx; L.bias[4] = P->b.bwy; L.bias[5] = P->b.bwz;
        for (int r = 0; r < 9; ++r) for (int c = 0; c < 9; ++c) L.info[9 * r + c] = ei->_info(r, c);
        // the random-walk informations: the solver's own helper supplies the bits, the matrices the reference code computed
        // (C.block<3,3>(9,9) / (12,12), inverted by the stand-in Eigen) must agree with it to rounding
        float C[225];
        for (int r = 0; r < 15; ++r) for (int c = 0; c < 15; ++c) C[15 * r + c] = P->C(r, c);
        double info[81];
        if (liba_link_information(C, 0, info, L.infoG, L.infoA) != ORB_OK) die("liba_link_information");
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
            const double g = rwG[i]->_info(r, c), a = rwA[i]->_info(r, c);
            if (std::fabs(g - L.infoG[3 * r + c]) > 1e-9 * std::fabs(L.infoG[4 * r]) || std::fabs(a - L.infoA[3 * r + c]) > 1e-9 * std::fabs(L.infoA[4 * r])) die("random-walk information differs");
        }
        links.push_back(L);
    }
    const int nE = (int)reproj.size();
    std::vector<int32_t> ekf(nE + 1), emp(nE + 1);
    std::vector<double> obs(3 * (size_t)nE + 3), invs2(nE + 1);
    for (int e = 0; e < nE; ++e) {
        EdgeReproj* E = reproj[e];
        E->_index = e;
        if (E->cam != 0 || E->_v.size() != 2) die("reprojection edge: camera / vertices");
        emp[e] = idx[E->_v[0]];
        ekf[e] = idx[E->_v[1]];
        for (int c = 0; c < 3; ++c) obs[3 * (size_t)e + c] = E->meas[c];
        invs2[e] = E->info00;
        const float want = E->meas[2] < 0 ? (float)sqrt(5.991) : (float)sqrt(7.815);
        if (!E->_rk || (float)E->_rk->_delta != want) die("reprojection edge: Huber delta");
    }
    liba_problem in;
    in.n_kf = nKF; in.n_mp = nMP; in.n_edges = nE; in.n_links = (int)links.size();
    in.state = state.data(); in.fixed = fixed.data(); in.point = point.data(); in.edge_kf = ekf.data(); in.edge_mp = emp.data();
    in.obs = obs.data(); in.inv_sigma2 = invs2.data(); in.links = links.data();
    const ImuCamPose& E0 = poses[0]->_est;
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) in.Tcb[3 * r + c] = E0.Rcb[0](r, c); in.Tcb[9 + r] = E0.tcb[0](r); }
    in.fx = E0.fx; in.fy = E0.fy; in.cx = E0.cx; in.cy = E0.cy; in.bf = E0.bf;
    in.lambda_init = _algorithm->_lambdaInit;
    in.max_iters = iterations > 0 ? iterations : 1;
    std::vector<double> state_out((size_t)nKF * 21), point_out((size_t)nMP * 3 + 3);
    _chi2.assign(nE + 1, 0.0); _depthPos.assign(nE + 1, 0);
    liba_result out;
    out.state = state_out.data(); out.point = point_out.data(); out.edge_chi2 = _chi2.data(); out.link_chi2 = nullptr; out.edge_depth_positive = _depthPos.data();
    liba_handle* h = nullptr;
    liba_create(0, &h);
    if (liba_solve(h, 1, &in, &out) != ORB_OK) die("liba_solve");
    _chiInit = out.chi2_initial; _chiLast = out.chi2_last_trial; _iterations = out.iterations;
    _solved = iterations > 0 ? 2 : 1;
    if (iterations <= 0) return;
    for (int k = 0; k < nKF; ++k) {
        ImuCamPose& E = poses[k]->_est;
        const double* S = &state_out[(size_t)k * 21];
        if (std::memcmp(S, &state[(size_t)k * 21], 12 * sizeof(double)) != 0) {   // an accepted update moved it: ImuCamPose::Update's Rcw = Rcb Rbw, tcw = Rcb tbw + tcb
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) E.Rwb(r, c) = S[3 * r + c];
            E.twb = Eigen::Vector3d(S[9], S[10], S[11]);
            const Eigen::Matrix3d Rbw = E.Rwb.transpose();
            const Eigen::Vector3d tbw = -(Rbw * E.twb);
            E.Rcw[0] = E.Rcb[0] * Rbw;
            E.tcw[0] = E.Rcb[0] * tbw + E.tcb[0];
        }
        KeyFrame* kf = poses[k]->_kf;
        if (vel.count(kf)) for (int c = 0; c < 3; ++c) { vel[kf]->_est(c) = S[12 + c]; bg[kf]->_est(c) = S[15 + c]; ba[kf]->_est(c) = S[18 + c]; }
    }
    for (int p = 0; p < nMP; ++p) points[p]->_est = Eigen::Vector3d(point_out[3 * (size_t)p], point_out[3 * (size_t)p + 1], point_out[3 * (size_t)p + 2]);
}

SparseOptimizer::~SparseOptimizer() {
    for (size_t i = 0; i < _edges.size(); ++i) { delete _edges[i]->_rk; delete _edges[i]; }
    for (size_t i = 0; i < _vertices.size(); ++i) delete _vertices[i];
    delete _algorithm;
}

}  // namespace g2o

namespace ORB_SLAM3 {
#include "_gen/liba_opt.inc"
}  // namespace ORB_SLAM3

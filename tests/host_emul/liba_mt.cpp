// tests/host_emul/liba_mt.cpp -- csrc/liba_core.cuh (the source of the k_liba kernel) executed by host threads that play a
// team of CTAs: LIBA_PAR_FOR strides by team thread id, LIBA_SYNC is a real team barrier, LIBA_LOCAL_SYNC a real per-CTA barrier.  Built with
// -fsanitize=thread, a missing barrier or a non-atomic shared update in the device algorithm shows up as a reported data race
// (exit code 66) on the CPU.  Usage: liba_mt <problem.bin> <result.bin> <threads per CTA> [CTAs]
#define LIBA_EMUL_THREADS 1
#include <barrier>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <thread>
#include <vector>

#include "../../orb_slam3_detailed_comments_b200/csrc/liba_pack.h"

namespace orb {
static std::barrier<>* g_team = nullptr;
static std::vector<std::barrier<>*> g_cta;
void liba_barrier_team() { g_team->arrive_and_wait(); }
void liba_barrier_cta(int rank) { g_cta[rank]->arrive_and_wait(); }
}

template <class T>
static std::vector<T> take(FILE* f, size_t n) {
    std::vector<T> v(n ? n : 1);
    if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
    return v;
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    const auto hdr = take<int32_t>(f, 6);   // n_kf n_mp n_edges n_links max_iters 0
    const auto dbl = take<double>(f, 18);   // lambda_init Tcb[12] cam5
    liba_problem p;
    p.n_kf = hdr[0]; p.n_mp = hdr[1]; p.n_edges = hdr[2]; p.n_links = hdr[3]; p.max_iters = hdr[4];
    const auto state = take<double>(f, 21 * (size_t)p.n_kf), point = take<double>(f, 3 * (size_t)p.n_mp), obs = take<double>(f, 3 * (size_t)p.n_edges),
               invs2 = take<double>(f, (size_t)p.n_edges);
    const auto ekf = take<int32_t>(f, p.n_edges), emp = take<int32_t>(f, p.n_edges);
    const auto links = take<liba_link>(f, p.n_links);
    const auto fixed = take<uint8_t>(f, p.n_kf);
    fclose(f);
    p.state = state.data(); p.fixed = fixed.data(); p.point = point.data(); p.edge_kf = ekf.data(); p.edge_mp = emp.data();
    p.obs = obs.data(); p.inv_sigma2 = invs2.data(); p.links = links.data();
    p.lambda_init = dbl[0];
    for (int i = 0; i < 12; ++i) p.Tcb[i] = dbl[1 + i];
    p.fx = dbl[13]; p.fy = dbl[14]; p.cx = dbl[15]; p.cy = dbl[16]; p.bf = dbl[17];

    const int T = atoi(argv[3]), CS = argc > 4 ? atoi(argv[4]) : 1;     // T threads per "CTA", CS "CTAs" in the team
    const orb::LibaLayout lay = orb::liba_pack(p, nullptr, nullptr, nullptr);
    if (lay.total == 0) return 3;
    std::vector<uint8_t> blob(lay.total + 16, 0);
    orb::LibaDev dev;
    orb::liba_pack(p, blob.data(), blob.data(), &dev);
    std::vector<double> red(T * CS);
    dev.red = red.data();
    std::barrier<> team(T * CS);
    orb::g_team = &team;
    std::vector<std::unique_ptr<std::barrier<>>> ctas;
    for (int r = 0; r < CS; ++r) { ctas.emplace_back(new std::barrier<>(T)); orb::g_cta.push_back(ctas.back().get()); }
    std::vector<std::thread> th;
    for (int r = 0; r < CS; ++r)
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, r, t] {
                orb::LibaDev P = dev;            // every "thread" holds its own copy of the descriptor, like the kernel
                P.cs = CS; P.rank = r; P.l_id = t; P.l_stride = T; P.t_id = r * T + t; P.t_stride = CS * T;
                orb::liba_optimize(P);
            });
    for (auto& t : th) t.join();

    std::vector<double> ostate(21 * (size_t)p.n_kf), opoint(3 * (size_t)p.n_mp + 1), ochi(p.n_edges + 1), olchi(3 * (size_t)p.n_links + 1);
    liba_result r;
    r.state = ostate.data(); r.point = opoint.data(); r.edge_chi2 = ochi.data(); r.link_chi2 = olchi.data(); r.edge_depth_positive = nullptr;
    orb::liba_unpack(p, blob.data(), dev, blob.data(), &r);
    FILE* o = fopen(argv[2], "wb");
    if (!o) return 2;
    const double sc[5] = {(double)r.iterations, (double)r.trials, r.lambda, r.chi2, r.chi2_initial};
    fwrite(sc, 8, 5, o);
    fwrite(ostate.data(), 8, 21 * (size_t)p.n_kf, o);
    fwrite(opoint.data(), 8, 3 * (size_t)p.n_mp, o);
    fwrite(ochi.data(), 8, p.n_edges, o);
    fwrite(olchi.data(), 8, 3 * (size_t)p.n_links, o);
    fclose(o);
    return 0;
}

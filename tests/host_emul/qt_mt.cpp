// tests/host_emul/qt_mt.cpp -- csrc/quadtree_core.cuh (the source of k_quadtree) executed by T host threads that play one CTA with a
// REAL barrier; built with -fsanitize=thread, a missing QT_SYNC (or two threads writing one slot) is reported as a data race
// (exit code 66).  Usage: qt_mt <in.bin> <out.bin> <threads>
//   in:  int32 n regionW regionH N wCell hCell nCols variant, then n x (x, y, score)      out: int32 S, then S x (x, y, score)
#define QT_EMUL_THREADS 1
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../orb_slam3_detailed_comments_b200/csrc/devmath.cuh"
#include "../../orb_slam3_detailed_comments_b200/csrc/quadtree_core.cuh"

namespace orbdev {
thread_local int qt_tid = 0;
int qt_nthreads = 1;
static std::barrier<>* g_bar = nullptr;
void qt_barrier() { g_bar->arrive_and_wait(); }
}
using namespace orbdev;

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t hdr[8];
    if (fread(hdr, 4, 8, f) != 8) return 2;
    const int n = hdr[0], variant = hdr[7];
    std::vector<int32_t> c(3 * (size_t)n + 3);
    if (n && fread(c.data(), 4, 3 * (size_t)n, f) != 3 * (size_t)n) return 2;
    fclose(f);
    QtGeom g;
    g.regionW = hdr[1]; g.regionH = hdr[2];
    g.nIni = (int)std::round((float)g.regionW / (float)g.regionH);
    g.hX = (float)g.regionW / (float)g.nIni;
    g.N = hdr[3]; g.wCell = hdr[4]; g.hCell = hdr[5]; g.nCols = hdr[6];
    int npow = 2;
    while (npow < n) npow <<= 1;
    std::vector<uint32_t> arr(npow, 0xffffffffu);
    for (int i = 0; i < n; ++i) arr[i] = qt_element(qt_pack_cand(c[3 * i], c[3 * i + 1], c[3 * i + 2]), g);
    const int cap = g.N + 20;
    std::vector<char> ws(qt_work_bytes(cap));
    std::vector<uint32_t> out(cap + 4);
    const int T = atoi(argv[3]);
    qt_nthreads = T;
    std::barrier<> bar(T);
    g_bar = &bar;
    std::vector<int> S(T, -9);
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t] {
            qt_tid = t;
            if (variant == 0) qt_bitonic_sort(arr.data(), npow); else qt_bitonic_sort_r8(arr.data(), npow);
            QtWork w;                             // per-thread copy of the workspace descriptor, like the kernel's registers
            qt_work_carve(w, ws.data(), cap);
            S[t] = variant == 0 ? qt_distribute_v<0>(arr.data(), n, g, w, out.data()) : qt_distribute_v<1>(arr.data(), n, g, w, out.data());
        });
    for (auto& t : th) t.join();
    for (int t = 1; t < T; ++t) if (S[t] != S[0]) return 4;    // the return value must be uniform across the CTA
    FILE* o = fopen(argv[2], "wb");
    if (!o) return 2;
    const int32_t s32 = S[0];
    fwrite(&s32, 4, 1, o);
    for (int i = 0; i < S[0]; ++i) {
        const int32_t r[3] = {(int32_t)(out[i] & 0xfff), (int32_t)((out[i] >> 12) & 0xfff), (int32_t)(out[i] >> 24)};
        fwrite(r, 4, 3, o);
    }
    fclose(o);
    return 0;
}

// tests/host_emul/qt_count.cpp -- barrier budget of k_quadtree's two variants: the quadtree core compiled for the host with QT_SYNC
// counting.  Usage: qt_count <in.bin (same format as qt_mt)>  ->  prints "variant syncs_sort syncs_total"
#define QT_HOST_COUNT_SYNCS 1
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../orb_slam3_detailed_comments_b200/csrc/devmath.cuh"
namespace orbdev { long qt_sync_count = 0; }
#include "../../orb_slam3_detailed_comments_b200/csrc/quadtree_core.cuh"
using namespace orbdev;

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t hdr[8];
    if (fread(hdr, 4, 8, f) != 8) return 2;
    const int n = hdr[0];
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
    for (int variant = 0; variant < 2; ++variant) {
        std::vector<uint32_t> arr(npow, 0xffffffffu);
        for (int i = 0; i < n; ++i) arr[i] = qt_element(qt_pack_cand(c[3 * i], c[3 * i + 1], c[3 * i + 2]), g);
        const int cap = g.N + 20;
        std::vector<char> ws(qt_work_bytes(cap));
        std::vector<uint32_t> out(cap + 4);
        QtWork w;
        qt_work_carve(w, ws.data(), cap);
        qt_sync_count = 0;
        if (variant == 0) qt_bitonic_sort(arr.data(), npow); else qt_bitonic_sort_r8(arr.data(), npow);
        const long s_sort = qt_sync_count;
        const int S = variant == 0 ? qt_distribute_v<0>(arr.data(), n, g, w, out.data()) : qt_distribute_v<1>(arr.data(), n, g, w, out.data());
        printf("%d %ld %ld %d %d\n", variant, s_sort, qt_sync_count, npow, S);
    }
    return 0;
}

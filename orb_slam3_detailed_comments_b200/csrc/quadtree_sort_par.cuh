// quadtree_sort_par.cuh -- std::sort(vSizeAndPointerToNode.begin(), end(), compareNodes) of DistributeOctTree's ordered phase
// (/root/reference/src/ORBextractor.cc:932-940) with libstdc++'s exact sequence of moves, but spread over the CTA.
//
// compareNodes orders by (point count, UL.x) only, nodes tie often, and the reference then divides the tied nodes in whatever order
// libstdc++'s introsort left them -- so the permutation itself is part of the result.  quadtree_core.cuh's qt_std_sort_items is a
// one-thread transcription (ncu, round 1: ~30 % of k_quadtree's stall samples are the other 255 threads waiting for it at the
// barrier).  The same moves, restructured:
//   * __introsort_loop recurses on the right part and iterates on the left: the segments alive at one recursion depth are
//     disjoint, so a GENERATION of segments is processed concurrently, one thread per segment, with the transcription's own
//     median-of-three + unguarded partition (or heap sort when the depth limit is hit); the critical path becomes
//     n + n/2 + n/4 + ... instead of n log2(n / 16);
//   * __final_insertion_sort only ever moves an element in front of strictly greater ones, i.e. it is a STABLE sort of whatever
//     the partitioning left: every thread computes the stable rank of one element instead (n <= a few hundred).
// Included by quadtree_core.cuh (uses QtItem, qt_item_less, qt_exscan and the QT_ SPMD macros).
#pragma once

namespace orbdev {

// one __introsort_loop step on [first, last): pivot to a[first] (__move_median_to_first), __unguarded_partition; returns the cut
ORB_HD int qt_sortpar_partition(QtItem* a, int first, int last) {
    const int mid = first + (last - first) / 2;
    const int A = first + 1, B = mid, C = last - 1;
    int pick;
    if (qt_item_less(a[A], a[B])) {
        if (qt_item_less(a[B], a[C])) pick = B;
        else if (qt_item_less(a[A], a[C])) pick = C;
        else pick = A;
    } else if (qt_item_less(a[A], a[C])) pick = A;
    else if (qt_item_less(a[B], a[C])) pick = C;
    else pick = B;
    { const QtItem t = a[first]; a[first] = a[pick]; a[pick] = t; }
    const QtItem pv = a[first];
    int lo = first + 1, hi = last;
    while (true) {
        while (qt_item_less(a[lo], pv)) ++lo;
        --hi;
        while (qt_item_less(pv, a[hi])) --hi;
        if (!(lo < hi)) break;
        const QtItem t = a[lo]; a[lo] = a[hi]; a[hi] = t;
        ++lo;
    }
    return lo;
}

// __partial_sort(first, last, last) == __make_heap + __sort_heap on h[0..len)
ORB_HD void qt_sortpar_heapsort(QtItem* h, int len) {
    for (int parent = (len - 2) / 2;; --parent) {
        const QtItem v = h[parent];
        int hole = parent, child = parent;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            if (qt_item_less(h[child], h[child - 1])) --child;
            h[hole] = h[child];
            hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            h[hole] = h[child - 1];
            hole = child - 1;
        }
        int par = (hole - 1) / 2;
        while (hole > parent && qt_item_less(h[par], v)) {
            h[hole] = h[par];
            hole = par;
            par = (hole - 1) / 2;
        }
        h[hole] = v;
        if (parent == 0) break;
    }
    for (int end = len - 1; end > 0; --end) {
        const QtItem v = h[end];
        h[end] = h[0];
        int hole = 0, child = 0;
        while (child < (end - 1) / 2) {
            child = 2 * (child + 1);
            if (qt_item_less(h[child], h[child - 1])) --child;
            h[hole] = h[child];
            hole = child;
        }
        if ((end & 1) == 0 && child == (end - 2) / 2) {
            child = 2 * (child + 1);
            h[hole] = h[child - 1];
            hole = child - 1;
        }
        int par = (hole - 1) / 2;
        while (hole > 0 && qt_item_less(h[par], v)) {
            h[hole] = h[par];
            hole = par;
            par = (hole - 1) / 2;
        }
        h[hole] = v;
    }
}

// a[0..n) sorted exactly as std::sort(a, a + n, compareNodes) leaves it.  tmp: n items.  seg / nxt: 3 ints per segment slot each
// (first, last, depth), room for n / 8 + 2 slots; flag: 2 * (n / 8 + 2) ints; scan_tmp: qt_exscan's scratch.  Whole CTA; returns after a
// barrier.  Every segment on a list is longer than 16, so a generation holds at most n / 17 of them.
ORB_HD void qt_std_sort_items_par(QtItem* a, int n, QtItem* tmp, int* seg, int* nxt, int* flag, int* scan_tmp) {
    if (n <= 1) return;   // uniform
    if (n > 16) {
        int lg = 0;
        for (int t = n; t > 1; t >>= 1) ++lg;
        QT_SERIAL { seg[0] = 0; seg[1] = n; seg[2] = 2 * lg; }
        QT_SYNC();
        int nseg = 1;
        const int nslots_max = n / 8 + 2;     // 2 * (n / 17) slots at most
        while (nseg > 0) {
            QT_PAR_FOR(s, nseg) {
                const int first = seg[3 * s], last = seg[3 * s + 1], depth = seg[3 * s + 2];
                int cut = first;
                if (depth == 0) qt_sortpar_heapsort(a + first, last - first);      // this segment is finished
                else cut = qt_sortpar_partition(a, first, last);
                // children: the left part [first, cut) and the right part [cut, last), alive while longer than 16
                const bool l_alive = depth != 0 && cut - first > 16, r_alive = depth != 0 && last - cut > 16;
                nxt[3 * (2 * s)] = first; nxt[3 * (2 * s) + 1] = cut; nxt[3 * (2 * s) + 2] = depth - 1;
                nxt[3 * (2 * s + 1)] = cut; nxt[3 * (2 * s + 1) + 1] = last; nxt[3 * (2 * s + 1) + 2] = depth - 1;
                flag[2 * s] = flag[nslots_max + 2 * s] = l_alive ? 1 : 0;
                flag[2 * s + 1] = flag[nslots_max + 2 * s + 1] = r_alive ? 1 : 0;
            }
            QT_SYNC();
            // compact the alive children into the next generation's list: the flags live twice (the scan turns one copy into ranks);
            // `seg` is dead once the partitions are done, so the survivors are scattered straight into it
            const int nslots = 2 * nseg;
            const int alive = qt_exscan(flag, nslots, scan_tmp);
            QT_PAR_FOR(s, nslots) {
                if (flag[nslots_max + s]) {
                    const int d = flag[s];
                    seg[3 * d] = nxt[3 * s];
                    seg[3 * d + 1] = nxt[3 * s + 1];
                    seg[3 * d + 2] = nxt[3 * s + 2];
                }
            }
            QT_SYNC();
            nseg = alive;
        }
    }
    // __final_insertion_sort == stable sort of the current arrangement
    QT_PAR_FOR(i, n) {
        const QtItem v = a[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const QtItem u = a[j];
            rank += (qt_item_less(u, v) || (j < i && !qt_item_less(v, u))) ? 1 : 0;
        }
        tmp[rank] = v;
    }
    QT_SYNC();
    QT_PAR_FOR(i, n) a[i] = tmp[i];
    QT_SYNC();
}

// Bitonic sort with two butterfly stages per pass (kept as the stepping stone to qt_bitonic_sort_r8 below, which k_quadtree_v1 uses).  For one merge size k the stages j = k/2 ... 1
// are taken in pairs (j, j/2): a thread loads the four elements that differ in bits j and j/2, does both compare-exchange
// levels in registers and stores them -- half the barriers and half the shared-memory round trips of qt_bitonic_sort; an odd
// stage count leaves one ordinary pass (j = 1).  Keys are unique per candidate pixel (root | path | score), so any correct sort
// yields the same array.
ORB_HD void qt_bitonic_sort_r4(uint32_t* arr, int npow) {
    for (int k = 2; k <= npow; k <<= 1) {
        int j = k >> 1;
        for (; j >= 2; j >>= 2) {
            const int h = j >> 1;
            QT_PAR_FOR(i, npow >> 2) {
                const int l0 = (i & (h - 1)) | ((i & ~(h - 1)) << 2);   // i spread over the index bits other than h and j = 2h
                uint32_t a0 = arr[l0], a1 = arr[l0 | h], a2 = arr[l0 | j], a3 = arr[l0 | j | h];
                const bool up = (l0 & k) == 0;
                uint32_t x;
                if ((a0 > a2) == up) { x = a0; a0 = a2; a2 = x; }
                if ((a1 > a3) == up) { x = a1; a1 = a3; a3 = x; }
                if ((a0 > a1) == up) { x = a0; a0 = a1; a1 = x; }
                if ((a2 > a3) == up) { x = a2; a2 = a3; a3 = x; }
                arr[l0] = a0; arr[l0 | h] = a1; arr[l0 | j] = a2; arr[l0 | j | h] = a3;
            }
            QT_SYNC();
        }
        if (j == 1) {
            QT_PAR_FOR(i, npow >> 1) {
                const int l = i << 1, r = l | 1;
                const uint32_t a = arr[l], b = arr[r];
                if ((a > b) == ((l & k) == 0)) { arr[l] = b; arr[r] = a; }
            }
            QT_SYNC();
        }
    }
}

// The same with up to THREE butterfly stages per pass (strides j, j/2, j/4: eight elements per thread, twelve compare-exchanges in
// registers); a remainder of two stages takes the four-element step, of one the ordinary pass.  8192 elements: 35 barriers instead
// of 49 (91 for one stage per pass).
ORB_HD void qt_bitonic_sort_r8(uint32_t* arr, int npow) {
#define QT_CX(x, y) { if (((x) > (y)) == up) { const uint32_t t_ = (x); (x) = (y); (y) = t_; } }
    for (int k = 2; k <= npow; k <<= 1) {
        int j = k >> 1;
        for (; j >= 4; j >>= 3) {
            const int h = j >> 1, q = j >> 2;
            QT_PAR_FOR(i, npow >> 3) {
                const int l0 = (i & (q - 1)) | ((i & ~(q - 1)) << 3);   // i spread over the index bits other than q, h = 2q, j = 4q
                uint32_t a0 = arr[l0], a1 = arr[l0 | q], a2 = arr[l0 | h], a3 = arr[l0 | h | q];
                uint32_t a4 = arr[l0 | j], a5 = arr[l0 | j | q], a6 = arr[l0 | j | h], a7 = arr[l0 | j | h | q];
                const bool up = (l0 & k) == 0;
                QT_CX(a0, a4) QT_CX(a1, a5) QT_CX(a2, a6) QT_CX(a3, a7)        // stride j
                QT_CX(a0, a2) QT_CX(a1, a3) QT_CX(a4, a6) QT_CX(a5, a7)        // stride h
                QT_CX(a0, a1) QT_CX(a2, a3) QT_CX(a4, a5) QT_CX(a6, a7)        // stride q
                arr[l0] = a0; arr[l0 | q] = a1; arr[l0 | h] = a2; arr[l0 | h | q] = a3;
                arr[l0 | j] = a4; arr[l0 | j | q] = a5; arr[l0 | j | h] = a6; arr[l0 | j | h | q] = a7;
            }
            QT_SYNC();
        }
        if (j == 2) {          // two stages left: strides 2 and 1
            QT_PAR_FOR(i, npow >> 2) {
                const int l0 = i << 2;
                uint32_t a0 = arr[l0], a1 = arr[l0 | 1], a2 = arr[l0 | 2], a3 = arr[l0 | 3];
                const bool up = (l0 & k) == 0;
                QT_CX(a0, a2) QT_CX(a1, a3) QT_CX(a0, a1) QT_CX(a2, a3)
                arr[l0] = a0; arr[l0 | 1] = a1; arr[l0 | 2] = a2; arr[l0 | 3] = a3;
            }
            QT_SYNC();
        } else if (j == 1) {   // one stage left
            QT_PAR_FOR(i, npow >> 1) {
                const int l = i << 1;
                uint32_t a0 = arr[l], a1 = arr[l | 1];
                const bool up = (l & k) == 0;
                QT_CX(a0, a1)
                arr[l] = a0; arr[l | 1] = a1;
            }
            QT_SYNC();
        }
    }
#undef QT_CX
}

}  // namespace orbdev

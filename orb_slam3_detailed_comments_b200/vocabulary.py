"""ORBVocabulary -- host-side mirror of the DBoW2 vocabulary object ORB-SLAM3 loads from ORBvoc.txt (include/ORBVocabulary.h),
for the one call on the per-frame path: transform(descriptors, BowVector, FeatureVector, levelsup) behind Frame::ComputeBoW
(src/Frame.cc:984-997).  The tree is given flattened (see include/orbslam3_b200.h, orbv_create)."""
import ctypes as C

import numpy as np

from . import _native as N


class ORBVocabulary:
    def __init__(self, child_offset, child_ids, node_desc, node_word, node_weight, L, device=0):
        self._L = N.lib()
        self._v = C.c_void_p()
        a = [np.ascontiguousarray(child_offset, np.int32), np.ascontiguousarray(child_ids, np.int32), np.ascontiguousarray(node_desc, np.uint8),
             np.ascontiguousarray(node_word, np.int32), np.ascontiguousarray(node_weight, np.float64)]
        N.check(self._L.orbv_create(int(device), len(a[3]), int(L), *[N.ptr(x) for x in a], C.byref(self._v)))
        self.depth = int(L)

    def close(self):
        if getattr(self, "_v", None) is not None and self._v.value:
            self._L.orbv_destroy(self._v)
            self._v = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def transform(self, extractor, total_rows, n_images, levelsup=4, bow=True):
        """Every descriptor of the extractor's last batch.  Returns dict(word, node, weight per compact row; and, with bow,
        bow_word / bow_weight: one (words, weights) pair per image -- mBowVec; `node` is what mFeatVec groups by)."""
        word, node = np.zeros(max(total_rows, 1), np.int32), np.zeros(max(total_rows, 1), np.int32)
        weight = np.zeros(max(total_rows, 1), np.float64)
        mf = extractor.max_features_per_image()
        cnt = np.zeros(n_images, np.int32)
        bw, bv = np.zeros((n_images, mf), np.int32), np.zeros((n_images, mf), np.float64)
        P = N.ptr
        N.check(self._L.orbv_transform(extractor._h, self._v, int(levelsup), 0, P(word), P(node), P(weight), P(cnt) if bow else None,
                                       P(bw) if bow else None, P(bv) if bow else None))
        out = dict(word=word[:total_rows], node=node[:total_rows], weight=weight[:total_rows])
        if bow:
            out["bow_word"] = [bw[i, :cnt[i]] for i in range(n_images)]
            out["bow_weight"] = [bv[i, :cnt[i]] for i in range(n_images)]
        return out


def synthetic_vocabulary(k=10, L=3, seed=0, stop_fraction=0.02):
    """A random k-ary tree of depth L in orbv_create's layout (the real ORBvoc.txt is k = 10, L = 6): node descriptors are
    their parent's with ~40 random bits flipped (so siblings are separable), leaves get consecutive word ids and an idf-like
    weight; a few words are stopped (weight 0)."""
    rng = np.random.default_rng(seed)
    desc, child_offset, child_ids, word, weight = [rng.integers(0, 256, 32, dtype=np.uint8)], [0], [], [-1], [0.0]
    level = [0]
    frontier, nxt_word = [0], 0
    children_of = {}
    for depth in range(1, L + 1):
        new_frontier = []
        for parent in frontier:
            ids = []
            for _ in range(k):
                d = desc[parent].copy()
                bits = rng.integers(0, 256, 40)
                for b in bits:
                    d[b // 8] ^= np.uint8(1 << (b % 8))
                desc.append(d); word.append(-1); weight.append(0.0); level.append(depth)
                ids.append(len(desc) - 1)
            children_of[parent] = ids
            new_frontier += ids
        frontier = new_frontier
    for leaf in frontier:
        word[leaf] = nxt_word
        nxt_word += 1
        weight[leaf] = 0.0 if rng.random() < stop_fraction else float(rng.uniform(0.5, 9.0))
    n = len(desc)
    for i in range(n):
        child_ids += children_of.get(i, [])
        child_offset.append(len(child_ids))
    return dict(child_offset=np.array(child_offset, np.int32), child_ids=np.array(child_ids, np.int32), node_desc=np.stack(desc),
                node_word=np.array(word, np.int32), node_weight=np.array(weight, np.float64), L=L)

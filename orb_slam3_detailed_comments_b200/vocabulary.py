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


def load_orbvoc_text(source):
    """ORBVocabulary::loadFromTextFile (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1338-1420) into orbv_create's flat layout.
    `source`: a path or a binary file object holding ORBvoc.txt ("k L scoring weighting", then one line per node:
    parent isLeaf 32 descriptor bytes weight).  Node ids are line numbers (root = 0), children keep file order
    (children.push_back), words are numbered in leaf order (m_words).
    One deliberate deviation: the reference's `while(!f.eof())` loop turns the file's final newline into one more, empty line,
    i.e. a childless extra child of the root whose descriptor is uninitialised memory; that node is not reproduced."""
    import pandas as pd
    fh = open(source, "rb") if isinstance(source, (str, bytes)) else source
    try:
        header = fh.readline().split()
        k, L, scoring, weighting = [int(v) for v in header[:4]]
        if not (0 <= k <= 20 and 1 <= L <= 10 and 0 <= scoring <= 5 and 0 <= weighting <= 3):
            raise ValueError("not a DBoW2 text vocabulary")
        if scoring != 0 or weighting != 0:
            raise ValueError("only TF_IDF weighting with L1 scoring (what ORBvoc.txt declares) is implemented on the device")
        t = pd.read_csv(fh, sep=r"\s+", header=None, engine="c", dtype=np.float64).to_numpy()
    finally:
        if fh is not source:
            fh.close()
    if t.shape[1] != 35:
        raise ValueError("expected parent, isLeaf, 32 descriptor bytes and a weight per line")
    n = len(t) + 1
    parent = t[:, 0].astype(np.int64)
    is_leaf = t[:, 1] > 0
    if (parent >= np.arange(1, n)).any() or (parent < 0).any():
        raise ValueError("a node precedes its parent")
    node_desc = np.zeros((n, 32), np.uint8)
    node_desc[1:] = t[:, 2:34].astype(np.uint8)
    node_weight = np.zeros(n, np.float64)
    node_weight[1:] = t[:, 34]
    node_word = np.full(n, -1, np.int32)
    node_word[1:][is_leaf] = np.arange(int(is_leaf.sum()), dtype=np.int32)
    counts = np.bincount(parent, minlength=n)
    child_offset = np.zeros(n + 1, np.int32)
    child_offset[1:] = np.cumsum(counts)
    child_ids = (np.argsort(parent, kind="stable") + 1).astype(np.int32)      # stable: file order inside a parent
    return dict(child_offset=child_offset, child_ids=child_ids, node_desc=node_desc, node_word=node_word, node_weight=node_weight, L=L, k=k)

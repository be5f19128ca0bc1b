"""GPU tier: DBoW2's vocabulary transform (Frame::ComputeBoW) on the device against the CPU oracle.
Bar: identical word / node ids per feature, bit-identical weights and BowVector doubles."""
import numpy as np
import pytest

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import ORBextractor, ORBVocabulary, synthetic_vocabulary, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,L,levelsup", [(10, 3, 2), (10, 4, 4), (6, 5, 4), (20, 2, 1), (3, 6, 7)])
def test_transform_matches_oracle(k, L, levelsup):
    W, H, nimg = 640, 480, 3
    imgs = np.stack([synth.frame(W, H, 40 + i) for i in range(nimg)])
    ex = ORBextractor(1200, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=nimg)
    ex.extract_batch(imgs)
    n, mono, off, kps, desc = ex.download(nimg)
    voc = synthetic_vocabulary(k=k, L=L, seed=k * 10 + L)
    V = ORBVocabulary(voc["child_offset"], voc["child_ids"], voc["node_desc"], voc["node_word"], voc["node_weight"], voc["L"])
    got = V.transform(ex, int(off[-1]), nimg, levelsup=levelsup)
    for i in range(nimg):
        a, b = int(off[i]), int(off[i + 1])
        r = po.bow_transform(voc, desc[a:b], levelsup)
        assert (got["word"][a:b] == r["word"]).all() and (got["node"][a:b] == r["node"]).all()
        assert (got["weight"][a:b].view(np.uint64) == r["weight"].view(np.uint64)).all()
        assert len(got["bow_word"][i]) == len(r["bow_word"]) and (got["bow_word"][i] == r["bow_word"]).all()
        assert (got["bow_weight"][i].view(np.uint64) == r["bow_weight"].view(np.uint64)).all()
        assert len(r["bow_word"]) > 10 and abs(r["bow_weight"].sum() - 1.0) < 1e-9
    V.close()
    ex.close()

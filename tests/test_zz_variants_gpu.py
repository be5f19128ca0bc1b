"""GPU tier (runs last): the selectable kernel variants of the per-frame path give the results of the defaults.
  ORB_PROJ_LANES   lanes per query of the two per-frame searches (default 4 / 8; 16; 32 = the warp-per-query kernel), read at every search
  ORB_PROJ_QPB     queries per CTA of the grouped candidate kernel
  ORB_RESIZE_VARIANT / ORB_BLUR_VARIANT   read by orbx_create: k_resize (0), k_resize_v2<4> (1, default), <8> (2), k_resize_v3 (3); k_blur<false> (0)
Each case reruns the oracle comparison of the default-variant tests."""
import numpy as np
import pytest

from oracle import pyoracle as po
from orb_slam3_detailed_comments_b200 import ORBextractor, synth
import test_matcher_gpu as tm
from test_extractor_gpu import _check_stages

pytestmark = pytest.mark.gpu

scene = tm.scene          # the module-scoped fixture of the matcher tests (its own instance here)


@pytest.mark.parametrize("lanes,qpb", [("32", None), ("16", None), ("4", "32"), ("8", "256")])
def test_per_frame_searches_with_other_lane_groupings(monkeypatch, scene, lanes, qpb):
    monkeypatch.setenv("ORB_PROJ_LANES", lanes)
    if qpb:
        monkeypatch.setenv("ORB_PROJ_QPB", qpb)
    tm.test_search_local_points_matches_oracle(scene)
    tm.test_search_last_frame_matches_oracle(scene)


@pytest.mark.parametrize("resize,blur", [("0", "0"), ("2", "1"), ("1", "0"), ("3", "1")])
@pytest.mark.parametrize("w,h,seed,nf", [(640, 480, 1, 1200), (500, 377, 8, 1500), (1280, 720, 5, 2000)])
def test_extractor_with_other_resize_and_blur_kernels(monkeypatch, resize, blur, w, h, seed, nf):
    monkeypatch.setenv("ORB_RESIZE_VARIANT", resize)
    monkeypatch.setenv("ORB_BLUR_VARIANT", blur)
    img = synth.frame(w, h, seed, 1.5, 60)
    ex = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
    ref = po.OracleExtractor(nf, 1.2, 8, 20, 7)
    mono, kps, desc = ex(img)
    rmono, rk, rd = ref(img)
    _check_stages(ex, ref)                      # pyramid levels, blurred levels, candidates, per-level keypoints
    assert mono == rmono and len(kps) == len(rk)
    assert (kps.view(np.uint8) == rk.view(np.uint8)).all() and (desc == rd).all()
    ex.close()

"""CPU tier, build container only: tests/test_oracle_vs_ref_matcher2.py once more at BASELINE config 5's geometry (1280 x 720, 2000 features):
every keyframe-typed matcher, the projection family and the MapPoint routines of the oracle against the reference's own functions."""
import importlib.util
import os

_here = os.path.dirname(os.path.abspath(__file__))
_old = os.environ.get("ORB_PIN_GEOMETRY")
os.environ["ORB_PIN_GEOMETRY"] = "720p"
try:
    _spec = importlib.util.spec_from_file_location("_matcher2_720p", os.path.join(_here, "test_oracle_vs_ref_matcher2.py"))
    _mod = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(_mod)
finally:
    if _old is None:
        del os.environ["ORB_PIN_GEOMETRY"]
    else:
        os.environ["ORB_PIN_GEOMETRY"] = _old
assert (_mod.W, _mod.H, _mod.NFEAT) == (1280, 720, 2000)
pytestmark = _mod.pytestmark
for _k, _v in list(vars(_mod).items()):      # the fixtures and every test function of the module, bound to the 720p constants
    if _k.startswith("test_") or _k in ("two_frames", "kf_target"):
        globals()[_k] = _v

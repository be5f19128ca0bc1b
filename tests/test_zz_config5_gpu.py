"""GPU tier: BASELINE.json's config 5 geometry (1280x720 stereo, 2000 features) through the end-to-end replay step -- extraction of both
eyes, ComputeStereoMatches, SearchByProjection(cur, last) and SearchByProjection(F, local map points) -- against the CPU oracle: the
bench's own parity gate (bench.Workload.parity_check), which the default run only applies to config 3."""
import pytest

pytestmark = pytest.mark.gpu


def test_config5_step_matches_oracle():
    import bench
    w = bench.Workload(5, 2, 2, 0, 0, pool_bytes=4 * 2 * 1280 * 720)
    w.e2e_setup()
    r = w.parity_check(1)
    assert r["ok"] and r["keypoints"] > 3500 and r["last_queries"] > 500 and r["local_queries"] > 3500
    w.close()

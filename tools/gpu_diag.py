"""First-contact diagnostics on the GPU box: run the extractor on a few frames, compare every stage with
the oracle and print WHERE the first difference is (instead of stopping at the first assert)."""
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po  # noqa: E402
from orb_slam3_detailed_comments_b200 import ORBextractor, synth  # noqa: E402


def diff(name, a, b):
    if a.shape != b.shape:
        print(f"  {name}: SHAPE {a.shape} vs {b.shape}")
        return False
    bad = np.argwhere(a != b)
    if len(bad):
        print(f"  {name}: {len(bad)} mismatches, first at {bad[0].tolist()} got {a[tuple(bad[0])]} want {b[tuple(bad[0])]}")
        return False
    return True


def main():
    ok_all = True
    for (w, h, seed, sig, nr, nf) in [(320, 240, 4, 3.0, 20, 500), (640, 480, 1, 1.5, 60, 1200), (752, 480, 3, 1.5, 60, 1200)]:
        print(f"== {w}x{h} seed {seed} nf {nf}")
        img = synth.frame(w, h, seed, sig, nr)
        try:
            ex = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
            ex.set_profiling(True)
            t0 = time.time()
            mono, kps, desc = ex(img)
            print("  gpu call", time.time() - t0, "s; n =", len(kps), "mono", mono, ex.last_timings())
            ref = po.OracleExtractor(nf, 1.2, 8, 20, 7)
            rmono, rk, rd = ref(img)
            for l in range(8):
                ok = diff(f"pyr{l}", ex.image_pyramid(0, l), ref.level_pyramid(l))
                c, rc = ex.candidates(0, l), ref.level_cands(l)
                ok &= diff(f"cand{l}", c, rc)
                ok &= diff(f"blur{l}", ex.image_pyramid(0, l, True), ref.level_blurred(l))
                k = ex.level_keypoints(0, l)
                r = ref.level_kps(l)
                r3 = np.stack([r["x"], r["y"], r["response"]], 1).astype(np.int32) if len(r) else np.zeros((0, 3), np.int32)
                ok &= diff(f"tree{l}", k, r3)
                ok_all &= ok
            ok = len(kps) == len(rk) and mono == rmono
            if ok:
                for f in kps.dtype.names:
                    ok &= diff("kp." + f, kps[f], rk[f])
                ok &= diff("desc", desc, rd)
            else:
                print("  count mismatch", len(kps), len(rk), mono, rmono)
            ok_all &= ok
            print("  RESULT", "OK" if ok else "MISMATCH")
            ex.close()
        except Exception:
            traceback.print_exc()
            ok_all = False
    print("ALL OK" if ok_all else "SOME MISMATCH")


if __name__ == "__main__":
    main()

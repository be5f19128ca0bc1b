#!/usr/bin/env python3
"""bench.py -- tracking-thread frames/s on synthetic 640x480 stereo, 1200 features (BASELINE.json metric).

One "step" = one batch of B synthetic stereo frames through the tracking-thread hot path: both eyes
through ORBextractor, Frame::ComputeStereoMatches, ORBmatcher::SearchByProjection(cur, last) (motion model)
and ORBmatcher::SearchByProjection(F, local map points) (TrackLocalMap).  The local-BA kernel (config 4, a
LocalMapping-thread job, not per frame) is timed separately and reported under "lba".
`value` is measured with the frames resident in HBM; `e2e` goes through the host-buffer C ABI (H2D of the
images and D2H of keypoints/descriptors inside the timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

N>1 is launched by torchrun (one rank per GPU); frames shard across ranks with no data-path collective
(weak scaling).  --impl reference times the CPU oracle (the reference cannot be built in this image) on the
host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, NFEAT = 640, 480, 1200
METRIC = "tracking-thread frames/sec (640x480 stereo, 1200 feat)"
UNIT = "frames/s"

# algorithmic bytes per IMAGE: read each input once, write each output once
def algorithmic_bytes(w, h, n_kp, n_cand, level_px):
    p_all = sum(level_px)
    pyramid = (p_all - level_px[-1]) + (p_all - level_px[0])
    fast = p_all + 12 * n_cand
    blur = 2 * p_all
    angle = n_kp * 749 + 4 * n_kp
    desc = n_kp * 512 + n_kp * 60
    tree = 12 * n_cand + 28 * n_kp
    return {"pyramid": pyramid, "fast": fast, "quadtree": tree, "blur": blur, "orient_desc": angle + desc}


def make_pairs(n_pairs, base=8):
    """n_pairs distinct stereo pairs: `base` generated pairs + cheap deterministic variants."""
    from orb_slam3_detailed_comments_b200 import synth
    gen = [synth.stereo_pair(W, H, seed=100 + i)[:2] for i in range(min(base, n_pairs))]
    out = np.empty((n_pairs, 2, H, W), np.uint8)
    for i in range(n_pairs):
        l, r = gen[i % len(gen)]
        k = i // len(gen)
        if k % 4 == 1:
            l, r = 255 - l, 255 - r
        elif k % 4 == 2:
            l, r = l[::-1], r[::-1]
        elif k % 4 == 3:
            l, r = 255 - l[::-1], 255 - r[::-1]
        sh = (k // 4) * 7
        out[i, 0] = np.roll(l, sh, axis=0)
        out[i, 1] = np.roll(r, sh, axis=0)
    return out


def local_map_queries(k, d, z, rng, bf):
    """Local-map query set of one frame (SearchLocalPoints shape): one map point per feature of the frame (its own
    descriptor, projection jittered by 1.5 px, stereo coordinate from its depth or 5 m) plus as many unrelated points
    (random positions, random descriptors) -- roughly the matched / unmatched mix of a real local map."""
    n = len(k)
    jit = rng.normal(0, 1.5, (n, 2)).astype(np.float32)
    zz = np.where(z > 0, z, 5.0).astype(np.float32)
    x = np.concatenate([k["x"] + jit[:, 0], rng.uniform(20, W - 20, n)]).astype(np.float32)
    y = np.concatenate([k["y"] + jit[:, 1], rng.uniform(20, H - 20, n)]).astype(np.float32)
    zq = np.concatenate([zz, rng.uniform(2, 15, n)]).astype(np.float32)
    lvl = np.concatenate([k["octave"], rng.integers(0, 8, n)]).astype(np.int32)
    vc = rng.uniform(0.99, 1.0, 2 * n).astype(np.float32)
    desc = np.concatenate([d, rng.integers(0, 256, (n, 32), dtype=np.uint8)])
    # world position of every map point, in the frame's own camera frame (the synthetic map is anchored there): the matched
    # half sits where the feature was triangulated, so its projection differs from the query's by the 1.5 px jitter
    xs = np.concatenate([k["x"], x[n:]]).astype(np.float32)
    ys = np.concatenate([k["y"], y[n:]]).astype(np.float32)
    xw = np.stack([(xs - 320.0) * zq / 435.2, (ys - 240.0) * zq / 435.2, zq], 1).astype(np.float32)
    return x, y, (x - bf / zq).astype(np.float32), lvl, vc, desc, xw


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu):
        self.rows, self.proc, self.gpu = [], None, gpu

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout=5.0):
        t0 = time.time()
        while not self.rows and time.time() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, t_begin=None, t_end=None):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        rows = [r for (t, r) in self.rows if (t_begin is None or t >= t_begin) and (t_end is None or t <= t_end + 0.05)]
        if not rows:
            rows = [r for (_, r) in self.rows][-3:]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_oracle_frames(pairs, threads, with_pose_opt=False):
    """Reference arm: the CPU oracle on `threads` host threads.  Per stereo frame: both eyes through the
    extractor (on two threads like Frame.cc:136-141 when threads >= 2), ComputeStereoMatches, then the two
    projection searches against a map made of the frame's own stereo points (same shape as the GPU arm); with_pose_opt adds
    PoseOptimization over the features that got a map point after each search.
    Returns (frames/s, seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle as po
    FX, FY, CX, CY, BF, BL = 435.2, 435.2, 320.0, 240.0, 47.9, 0.11
    nthr = max(1, threads)
    nworkers = max(1, nthr // 2)
    bounds = np.array([0, W, 0, H], np.float32)
    cam6 = np.array([FX, FY, CX, CY, BF, BL], np.float32)
    T = np.array([0, 0, 0, 1, 0.002, 0.001, 0], np.float32)

    def frame(args):
        l, r, eL, eR, pool2 = args
        if pool2 is not None:
            fr = pool2.submit(eR, r)
            _, kL, dL = eL(l)
            _, kR, dR = fr.result()
        else:
            _, kL, dL = eL(l)
            _, kR, dR = eR(r)
        uR, dep, _ = po.stereo_matches(eL, eR, kL, dL, kR, dR, BF, BL)
        sel = np.nonzero(dep > 0)[0]
        if len(sel) == 0:
            return 0
        z = dep[sel]
        pts = np.stack([(kL["x"][sel] - CX) * z / FX, (kL["y"][sel] - CY) * z / FY, z], 1).astype(np.float32)
        sf = eL.scale_factors
        isg = (1.0 / (sf * sf)).astype(np.float32)
        cam5 = np.float32([FX, FY, CX, CY, BF])

        def pose_opt(feat, xw):        # Optimizer::PoseOptimization over the features that hold a map point (feature order)
            obs = np.stack([kL["x"][feat], kL["y"][feat], uR[feat]], 1)
            return po.pose_optimization(T, xw, obs, isg[kL["octave"][feat]], cam5)
        fm, _ = po.search_last(kL, dL, uR, bounds, sf, cam6, T, 0, pts, kL["octave"][sel], kL["angle"][sel], dL[sel],
                               np.ones(len(sel), np.uint8), 15.0, True)
        feat = np.nonzero(fm >= 0)[0]
        if with_pose_opt:
            pose_opt(feat, pts[fm[feat]])
        x, y, xr, lvl, vc, dq, xwl = local_map_queries(kL, dL, dep, np.random.default_rng(7), BF)
        mt, _ = po.search_local(kL, dL, uR, bounds, sf, x, y, xr, lvl, vc, dq, 3.0, 0.8)
        qs = np.nonzero(mt >= 0)[0]
        order = np.argsort(mt[qs], kind="stable")
        if with_pose_opt:
            pose_opt(mt[qs][order], xwl[qs][order])
        return 0
    exs = [(po.OracleExtractor(NFEAT, 1.2, 8, 20, 7), po.OracleExtractor(NFEAT, 1.2, 8, 20, 7)) for _ in range(nworkers)]
    pools2 = [ThreadPoolExecutor(1) if nthr >= 2 else None for _ in range(nworkers)]
    t0 = time.perf_counter()
    chunks = [[(pairs[i, 0], pairs[i, 1], exs[wk][0], exs[wk][1], pools2[wk]) for i in range(wk, len(pairs), nworkers)]
              for wk in range(nworkers)]
    if nworkers == 1:
        for a in chunks[0]:
            frame(a)
    else:
        with ThreadPoolExecutor(nworkers) as pool:
            list(pool.map(lambda ch: [frame(a) for a in ch], chunks))
    dt = time.perf_counter() - t0
    for p2 in pools2:
        if p2 is not None:
            p2.shutdown()
    return len(pairs) / dt, dt


def _straw_man_guard(pair):
    """the oracle's OpenCV-equivalent stages are scalar restatements; cv2 (one thread) runs the same stages with
    SIMD.  Times both on one image of the workload and returns the factor by which a frame of the CPU arm gets cheaper when every
    such stage is charged at min(oracle, cv2): ratio = frame_ms_with_cv2_stages / frame_ms_oracle (<= 1).
    cv2's FAST is charged as ONE whole-level detect at iniThFAST per pyramid level -- a lower bound on what the reference does
    (577 per-cell cv::FAST calls over windows that overlap by 6 px, empty cells again at minThFAST; ORBextractor.cc:1098-1166)."""
    try:
        import cv2
    except ImportError:
        return None
    from oracle import pyoracle as po
    cv2.setNumThreads(1)
    img = np.ascontiguousarray(pair[0])
    ex, ex2 = po.OracleExtractor(NFEAT, 1.2, 8, 20, 7), po.OracleExtractor(NFEAT, 1.2, 8, 20, 7)

    def best(f, n=5):
        f()
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            f()
            ts.append(time.perf_counter() - t0)
        return 1e3 * min(ts)
    t_img = best(lambda: ex(img), 3)
    stage = {k: 1e3 * v for k, v in ex.timings().items()}
    img_r = np.ascontiguousarray(pair[1])
    t_img_r = best(lambda: ex2(img_r), 3)
    sizes = [ex.level_size(l) for l in range(8)]

    def pyr():
        lv = [img]
        for l in range(1, 8):
            lv.append(cv2.resize(lv[-1], sizes[l], interpolation=cv2.INTER_LINEAR))
        for x in lv:
            cv2.copyMakeBorder(x, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
        return lv
    lv = pyr()
    fd = cv2.FastFeatureDetector_create(20, True)
    cv = {"pyramid": best(pyr), "fast": best(lambda: [fd.detect(x) for x in lv]),
          "blur": best(lambda: [cv2.GaussianBlur(x, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101) for x in lv])}
    delta = sum(max(0.0, stage[k] - cv[k]) for k in cv)
    # the rest of a frame (ComputeStereoMatches + both searches), isolated and single-threaded: one whole frame minus two extractions
    one = np.ascontiguousarray(pair)[None]
    cpu_oracle_frames(one, 1)
    t_frame_seq = 1e3 * min(cpu_oracle_frames(one, 1)[1] for _ in range(3))
    t_rest = max(0.0, t_frame_seq - t_img - t_img_r)
    t_eye = max(t_img, t_img_r)                                          # the two eyes run in parallel (Frame.cc:136-141)
    f_or, f_cv = t_eye + t_rest, max(t_eye - delta, 0.0) + t_rest
    return {"oracle_stage_ms_per_image": stage, "cv2_stage_ms_per_image": cv, "image_ms_oracle": t_img, "rest_of_frame_ms_oracle": t_rest, "delta_ms_per_image": delta,
            "frame_ms_oracle": f_or, "frame_ms_with_cv2_stages": f_cv, "ratio": f_cv / f_or if f_or > 0 else 1.0,
            "note": "value = oracle-measured frames/s / ratio; cv2 4.x single-threaded; FAST charged as one whole-level detect per level"}


def straw_man_guard(pair):
    try:
        return _straw_man_guard(pair)
    except Exception as exc:      # the guard must never cost the bench line
        return {"error": repr(exc), "ratio": 1.0}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    n_pairs = max(cores // 2, 1) * 2
    pairs = make_pairs(n_pairs, base=4)
    for _ in range(args.warmup):
        cpu_oracle_frames(pairs[:max(1, cores // 2)], cores)
    t0 = time.perf_counter()
    frames = 0
    for _ in range(args.steps):
        cpu_oracle_frames(pairs, cores)
        frames += len(pairs)
    dt = time.perf_counter() - t0
    v_oracle = frames / dt
    guard = straw_man_guard(pairs[0])
    v = v_oracle / guard["ratio"] if guard else v_oracle
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "impl": "reference",
            "config": {"workload": "config 3: stereo 640x480, 1200 features per eye; per frame: ORBextractor L+R, "
                                   "ComputeStereoMatches, SearchByProjection(cur,last), SearchByProjection(F,local map) -- "
                                   "CPU oracle port (the reference needs OpenCV/Eigen and cannot be built here)",
                       "frames_per_step": len(pairs)},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{len(pairs)} stereo frames per step on {cores} threads", "value_oracle_only": v_oracle,
                             "straw_man_guard": guard},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="stereo frames per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extract-only", action="store_true", help="headline step only: skip the LBA / inertial / knn / PoseOptimization sections")
    ap.add_argument("--handles", type=int, default=4, help="extractor handles (CUDA streams) the steps are pipelined over")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from orb_slam3_detailed_comments_b200 import ORBextractor, _native
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "NONE"       # NCCL prints its version banner on stdout from VERSION up: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    nimg = 2 * B
    dev = torch.device("cuda", local)
    FX, FY, CX, CY, BF, BL = 435.2, 435.2, 320.0, 240.0, 47.9, 0.11
    from orb_slam3_detailed_comments_b200 import ORBmatcher, camera, Optimizer, synth
    cam = camera(FX, FY, CX, CY, BF, BL, W, H)
    ex = ORBextractor(NFEAT, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=nimg, device=local)
    NH = max(2, args.handles)
    # NH handles = NH CUDA streams: batches i+1 .. i+NH-1 are queued while batch i's ordered passes drain
    exs = [ex] + [ORBextractor(NFEAT, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=nimg, device=local) for _ in range(NH - 1)]
    streams = [torch.cuda.ExternalStream(e.cuda_stream(), device=dev) for e in exs]
    stream = streams[0]
    m_last, m_local = ORBmatcher(0.9, True), ORBmatcher(0.8, True)

    # input pool larger than L2 (126 MB): POOL batches of 2B images, cycled through the timed steps
    pool_batches = max(2, int(np.ceil(140e6 / (nimg * W * H))))
    pairs = make_pairs(B * pool_batches, base=8)
    host_pool = torch.from_numpy(pairs.reshape(pool_batches, nimg, H, W)).pin_memory()
    dev_pool = host_pool.cuda(non_blocking=False)

    # ---- map state for the matchers, one query set per pool batch: the "last frame" of every sequence is the frame
    # itself one step earlier (its stereo points, unprojected), the local map holds one map point per feature plus as
    # many unrelated points (local_map_queries) ----------------------------------------------------------------
    rng = np.random.default_rng(1234 + rank)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if len(xs) else np.zeros(0, dt)
    H_LAST, H_LOC, D_LAST, D_LOC = [], [], [], []
    for pb in range(pool_batches):
        ex.extract_batch_device(dev_pool[pb].data_ptr(), nimg, W, H)
        ex.stereo_batch(B, BF, BL)
        n0, _, off0, kps0, desc0 = ex.download(nimg)
        uR0, dep0 = ex.stereo_download(int(off0[-1]))
        q_last = dict(off=[0], xw=[], oct=[], ang=[], desc=[], obs=[])
        q_loc = dict(off=[0], px=[], py=[], pxr=[], lvl=[], vc=[], desc=[], xw=[])
        for p in range(B):
            a, b = int(off0[2 * p]), int(off0[2 * p + 1])
            k, d, z = kps0[a:b], desc0[a:b], dep0[a:b]
            sel = np.nonzero(z > 0)[0]
            pts = np.stack([(k["x"][sel] - CX) * z[sel] / FX, (k["y"][sel] - CY) * z[sel] / FY, z[sel]], 1).astype(np.float32)
            q_last["xw"].append(pts); q_last["oct"].append(k["octave"][sel].astype(np.int32))
            q_last["ang"].append(k["angle"][sel].astype(np.float32)); q_last["desc"].append(d[sel])
            q_last["obs"].append(np.ones(len(sel), np.uint8)); q_last["off"].append(q_last["off"][-1] + len(sel))
            x, y, xr, lvl, vc, dq, xwl = local_map_queries(k, d, z, rng, BF)
            q_loc["px"].append(x); q_loc["py"].append(y); q_loc["pxr"].append(xr); q_loc["lvl"].append(lvl)
            q_loc["vc"].append(vc); q_loc["desc"].append(dq); q_loc["xw"].append(xwl)
            q_loc["off"].append(q_loc["off"][-1] + len(x))
        h_last = dict(fimg=np.arange(0, nimg, 2, dtype=np.int32), off=np.array(q_last["off"], np.int32),
                      Tcw=np.tile(np.array([0, 0, 0, 1, 0.002, 0.001, 0], np.float32), (B, 1)), dir=np.zeros(B, np.int32),
                      xw=cat(q_last["xw"], np.float32), oct=cat(q_last["oct"], np.int32), ang=cat(q_last["ang"], np.float32),
                      desc=cat(q_last["desc"], np.uint8), obs=cat(q_last["obs"], np.uint8))
        h_loc = dict(fimg=h_last["fimg"], off=np.array(q_loc["off"], np.int32), px=cat(q_loc["px"], np.float32),
                     py=cat(q_loc["py"], np.float32), pxr=cat(q_loc["pxr"], np.float32), lvl=cat(q_loc["lvl"], np.int32),
                     vc=cat(q_loc["vc"], np.float32), desc=cat(q_loc["desc"], np.uint8), xw=cat(q_loc["xw"], np.float32))
        H_LAST.append(h_last); H_LOC.append(h_loc)
        D_LAST.append({k: T(v) for k, v in h_last.items()}); D_LOC.append({k: T(v) for k, v in h_loc.items()})
    rows_cap = nimg * 1500
    d_fm = [torch.full((rows_cap,), -1, dtype=torch.int32, device=dev) for _ in range(NH)]
    d_nm = [torch.zeros(2 * B, dtype=torch.int32, device=dev) for _ in range(NH)]
    max_loc = max(int(hl["off"][-1]) for hl in H_LOC)
    d_match = [torch.full((max(max_loc, 1),), -1, dtype=torch.int32, device=dev) for _ in range(NH)]
    nq_last = float(np.mean([int(hl["off"][-1]) for hl in H_LAST])); nq_loc = float(np.mean([int(hl["off"][-1]) for hl in H_LOC]))
    # PoseOptimization after each search (Tracking.cc:3443, 3522): edge lists and results, per handle
    from orb_slam3_detailed_comments_b200 import PoseOptimization, PoseOptimizationDevice, PoseEdgesDevice, PoseOptimizationFrames
    CAM5 = [FX, FY, CX, CY, BF]
    zi = lambda n, dt: torch.zeros(n, dtype=dt, device=dev)
    d_po = [dict(off=zi(B + 1, torch.int32), feat=zi(rows_cap, torch.int32), xw=zi((rows_cap, 3), torch.float32), obs=zi((rows_cap, 3), torch.float32),
                 w=zi(rows_cap, torch.float32), pose=zi((2, B, 7), torch.float64), out=zi(rows_cap, torch.uint8), inl=zi((2, B), torch.int32))
            for _ in range(NH)]

    def submit_device(i):
        e = exs[i % NH]
        e.extract_batch_device(dev_pool[i % pool_batches].data_ptr(), nimg, W, H)
        e.stereo_batch(B, BF, BL)

    def finish_device(i, with_po=False):
        k = i % NH
        e = exs[k]
        d_last, d_loc = D_LAST[i % pool_batches], D_LOC[i % pool_batches]
        m_last.SearchByProjectionLastFrameDevice(e, cam, B, d_last["fimg"], d_last["off"], d_last["Tcw"], d_last["dir"],
                                                 d_last["xw"], d_last["oct"], d_last["ang"], d_last["desc"], d_last["obs"],
                                                 15.0, d_fm[k], d_nm[k][:B])
        po_ = d_po[k]
        if with_po:
            PoseEdgesDevice(e, B, d_last["fimg"], d_last["xw"], po_["off"], po_["feat"], po_["xw"], po_["obs"], po_["w"], feature_match=d_fm[k])
            PoseOptimizationDevice(e, B, po_["off"], d_last["Tcw"], po_["xw"], po_["obs"], po_["w"], CAM5, po_["pose"][0], po_["out"], po_["inl"][0])
        m_local.SearchByProjectionDevice(e, cam, B, d_loc["fimg"], d_loc["off"], d_loc["px"], d_loc["py"], d_loc["pxr"],
                                         d_loc["lvl"], d_loc["vc"], d_loc["desc"], d_match[k], d_nm[k][B:], th=3.0)
        if with_po:
            PoseEdgesDevice(e, B, d_loc["fimg"], d_loc["xw"], po_["off"], po_["feat"], po_["xw"], po_["obs"], po_["w"],
                            query_offset=d_loc["off"], query_match=d_match[k])
            PoseOptimizationDevice(e, B, po_["off"], d_last["Tcw"], po_["xw"], po_["obs"], po_["w"], CAM5, po_["pose"][1], po_["out"], po_["inl"][1])

    def step_device(i):     # un-pipelined form (used for the per-stage roofline pass)
        submit_device(i)
        finish_device(i)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM; software-pipelined over the NH handles ------------------------
    sampler = ClockSampler(local)
    sampler.start()
    for i in range(max(args.warmup, NH)):
        step_device(i)
    barrier()
    sampler.wait_first()
    launches0 = _native.lib().orb_kernel_launches()
    def timed_device_loop(first, count, with_po):
        """count pipelined steps starting at batch index `first`; returns device milliseconds (events on stream 0)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(streams[0])
        for st in streams[1:]:
            st.wait_event(e0)
        for j in range(min(NH - 1, count)):
            submit_device(first + j)
        for i in range(count):
            if i + NH - 1 < count:
                submit_device(first + i + NH - 1)
            finish_device(first + i, with_po)
        for st in streams[1:]:
            ev = torch.cuda.Event()
            ev.record(st)
            streams[0].wait_event(ev)
        e1.record(streams[0])
        barrier()
        return e0.elapsed_time(e1)
    t_begin = time.time()
    ms = timed_device_loop(args.warmup, args.steps, False)
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end)
    launches = _native.lib().orb_kernel_launches() - launches0
    # per-stage times for the roofline: a serial (un-overlapped) pass over the same steps, CUDA events per stage
    ex.set_profiling(True)
    ser0, ser1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nser = max(2, min(args.steps, 32) // 2 * 2)     # at least one profiled serial step whatever --steps is
    ser0.record(streams[0])
    for i in range(0, nser, 2):
        step_device(NH * (args.warmup + i))  # multiple of NH => handle 0, the profiled one
    ser1.record(streams[0])
    barrier()
    serial_ms_per_step = ser0.elapsed_time(ser1) / max(nser // 2, 1)
    stage_ms = ex.last_timings()
    ex.set_profiling(False)
    n, mono, off = ex.counts(nimg)
    nm_host = d_nm[0].cpu().numpy()
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max * 1e-3)

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region ----------------------
    # Two extractor handles (two CUDA streams) alternate: the H2D + extraction + stereo of batch i+1 are queued
    # before the blocking result reads of batch i, so PCIe traffic overlaps compute (the usage INTEGRATION.md
    # recommends for sequence replay).  Every step still moves its own images in and its own results out.
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    P_LAST = [{k: pin(v) for k, v in hl.items()} for hl in H_LAST]
    P_LOC = [{k: pin(v) for k, v in hl.items()} for hl in H_LOC]
    pz = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory().numpy()
    from orb_slam3_detailed_comments_b200._native import KP_DTYPE
    def out_buffers():
        return dict(kps=torch.zeros(rows_cap * 28, dtype=torch.uint8).pin_memory().numpy().view(KP_DTYPE), desc=pz((rows_cap, 32), torch.uint8),
                    ur=pz(rows_cap, torch.float32), dep=pz(rows_cap, torch.float32), fm=pz(rows_cap, torch.int32), nm1=pz(B, torch.int32),
                    mt=pz(max(max_loc, 1), torch.int32), nm2=pz(B, torch.int32), pose=pz((B, 7), torch.float64), outl=pz(rows_cap, torch.uint8),
                    inl=pz(B, torch.int32))
    OUT = [out_buffers() for _ in range(NH)]

    def step_e2e(i, k, with_po=False):
        """One step through the host-pointer C ABI on handle k: images in (pinned host -> device), extraction, stereo,
        both searches, every result back in pinned host buffers.  Each call blocks until its results are on the host."""
        e, o = exs[k], OUT[k]
        p_last, p_loc = P_LAST[i % pool_batches], P_LOC[i % pool_batches]
        e.extract_batch_async(host_pool[i % pool_batches].numpy())
        e.stereo_batch(B, BF, BL)
        nn, mm, oo, kk, dd = e.download(nimg, out=(o["kps"], o["desc"]))
        rows = int(oo[-1])
        e.stereo_download(rows, out=(o["ur"], o["dep"]))
        fm, _ = m_last.SearchByProjectionLastFrame(e, cam, p_last["fimg"], p_last["off"], p_last["Tcw"], p_last["dir"], p_last["xw"], p_last["oct"],
                                                   p_last["ang"], p_last["desc"], p_last["obs"], 15.0, rows, out=(o["fm"], o["nm1"]))
        if with_po:   # PoseOptimization straight from the search results (host arrays in, pose / mvbOutlier / inliers out)
            PoseOptimizationFrames(e, p_last["fimg"], p_last["Tcw"], p_last["xw"], CAM5, feature_match=fm, out=(o["pose"], o["outl"], o["inl"]))
        mt, _ = m_local.SearchByProjection(e, cam, p_loc["fimg"], p_loc["off"], p_loc["px"], p_loc["py"], p_loc["pxr"], p_loc["lvl"], p_loc["vc"],
                                           p_loc["desc"], th=3.0, out=(o["mt"], o["nm2"]))
        if with_po:
            PoseOptimizationFrames(e, p_loc["fimg"], p_last["Tcw"], p_loc["xw"], CAM5, query_offset=p_loc["off"], query_match=mt,
                                   out=(o["pose"], o["outl"], o["inl"]))
        return rows * (60 + 8 + 4) + 12 * nimg + 4 * int(nq_loc) + 8 * B + (2 * (B * (56 + 4) + rows // 2) if with_po else 0)

    # One host thread per handle, the deployment shape of sequence-sharded replay: every thread
    # drives its own handle / CUDA stream through the blocking C ABI, so one thread's result reads overlap the others'
    # uploads and kernels.  ctypes releases the GIL inside the calls.
    import concurrent.futures as cf
    pool = cf.ThreadPoolExecutor(max_workers=NH)

    def worker(k, first, count, with_po=False):
        tot = 0
        for i in range(first + k, first + count, NH):
            tot += step_e2e(i, k, with_po)
        return tot

    list(pool.map(lambda k: worker(k, 0, max(args.warmup, NH)), range(NH)))
    barrier()
    t0 = time.perf_counter()
    d2h = sum(pool.map(lambda k: worker(k, max(args.warmup, NH), args.steps), range(NH)))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # the same steps with PoseOptimization after each search: resident and end-to-end
    n_po = max(NH, args.steps // 2)
    timed_device_loop(args.warmup, NH, True)
    ms_po = timed_device_loop(args.warmup, n_po, True)
    list(pool.map(lambda k: worker(k, 0, NH, True), range(NH)))
    barrier()
    t0 = time.perf_counter()
    list(pool.map(lambda k: worker(k, NH, n_po, True), range(NH)))
    torch.cuda.synchronize()
    dt_po = time.perf_counter() - t0
    tp = torch.tensor([ms_po * 1e-3, dt_po], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    in_step = {"steps": n_po, "value_frames_per_s": world * B * n_po / float(tp[0].item()), "e2e_frames_per_s": world * B * n_po / float(tp[1].item()),
               "note": "the headline step plus PoseOptimization after each search (device correspondence walk + optimiser; host-pointer "
                       "orbo_pose_optimization_frames in the e2e leg)"}
    pool.shutdown()
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t.item())
    h2d_step = nimg * W * H + sum(v.nbytes for v in H_LAST[0].values()) + sum(v.nbytes for v in H_LOC[0].values())

    if rank == 0:
        # roofline of the dominant kernel, live from the stage events recorded over the timed steps
        level_px = [ex.level_size(l)[0] * ex.level_size(l)[1] for l in range(8)]
        n_kp = float(n.mean())
        n_cand = float(np.mean([sum(len(ex.candidates(b, l)) for l in range(8)) for b in range(2)]))
        ab = algorithmic_bytes(W, H, n_kp, n_cand, level_px)
        stages = {k: stage_ms[k] for k in ["pyramid", "fast", "quadtree", "blur", "orient_desc"]}
        stages["stereo+search"] = max(serial_ms_per_step - stage_ms["total"], 0.0)
        stages["serial_step_total"] = serial_ms_per_step
        ext = {k: stages[k] for k in ["pyramid", "fast", "quadtree", "blur", "orient_desc"]}
        top = max(ext, key=ext.get)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = ab[top] * nimg / (ext[top] * 1e-3) / 1e9
        traffic = None      # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_fast_traffic.json")))
            if top == "fast":
                traffic = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) * nimg / tr["images_per_launch"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic,
                    "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                    "stage_ms_per_step": stages,
                    "stage_gbs": {k: ab[k] * nimg / (ext[k] * 1e-3) / 1e9 for k in ext},
                    "algorithmic_bytes_per_image": ab,
                    "traffic_source": "profiles/r01_fast_traffic.json (ncu --set full, bytes per launch of 128 images)",
                    "note": "stage times from a serial pass (one stream) right after the timed region; the timed region itself is "
                            "software-pipelined over the handles' streams.  FAST is bound by integer issue and barrier / shared-memory latency (profiles/), not by HBM; see DESIGN.md"}
        cpu = None
        if not args.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only
            # faithful threading (Frame.cc:136-141): the two eyes on two threads, bounded sample
            n_cpu = min(len(pairs), 128)        # a few seconds of CPU work: stable to a few percent
            cpu_oracle_frames(pairs[:2], 2)     # warm-up (page-in, thread pools)
            v_oracle, secs = cpu_oracle_frames(pairs[:n_cpu], 2)
            guard = straw_man_guard(pairs[0])
            v = v_oracle / guard["ratio"] if guard else v_oracle
            cpu = {"value": v, "unit": UNIT, "cores": 2, "kind": "port",
                   "sample": f"{n_cpu} stereo frames (extraction + stereo matching + both projection searches), "
                             f"L/R eyes on 2 threads (Frame.cc:136-141), {secs:.1f} s",
                   "value_oracle_only": v_oracle, "straw_man_guard": guard}
        # ---- local BA (config 4): 20 KF / 3000 MP, one problem and a batch of 8 -------------------------
        lba = None
        try:
            if args.extract_only:
                raise RuntimeError("skipped (--extract-only)")
            from oracle import pyoracle as po
            opt = Optimizer(local)
            prs = [synth.lba_problem(seed=s) for s in range(8)]
            for _ in range(2):
                opt.LocalBundleAdjustment(prs[0], lambda_init=100.0)
            opt.LocalBundleAdjustmentBatch(prs, lambda_init=100.0)      # warm-up: sizes the pinned / device workspaces
            t0 = time.perf_counter()
            for s in range(5):
                g = opt.LocalBundleAdjustment(prs[s], lambda_init=100.0)
            t_one = (time.perf_counter() - t0) / 5
            t0 = time.perf_counter()
            opt.LocalBundleAdjustmentBatch(prs, lambda_init=100.0)
            t_batch = (time.perf_counter() - t0) / 8
            t0 = time.perf_counter()
            for s in range(3):
                r = po.lba(prs[s]["pose"], prs[s]["fixed"], prs[s]["point"], prs[s]["edge_kf"], prs[s]["edge_mp"], prs[s]["obs"],
                           prs[s]["inv_sigma2"], prs[s]["cam5"], 100.0, 10)
            t_cpu = (time.perf_counter() - t0) / 3
            lba = {"workload": "config 4: LocalBundleAdjustment 20 KF (2 fixed) / 3000 MP / ~18k edges, lambda_init 100",
                   "ms_per_solve_e2e": 1e3 * t_one, "ms_per_solve_batch8_e2e": 1e3 * t_batch, "cpu_oracle_ms": 1e3 * t_cpu,
                   "cpu_cores": 1, "iterations": int(g["iterations"]), "edges": int(len(prs[0]["edge_kf"]))}
            opt.close()
        except Exception as exc:   # never lose the headline line because the side benchmark failed
            lba = {"error": repr(exc)}
        # ---- LocalInertialBA ---------------------------------------------------------------------------
        liba = None
        if True:
            try:
                if args.extract_only:
                    raise RuntimeError("skipped (--extract-only)")
                from oracle import pyoracle as po
                from orb_slam3_detailed_comments_b200 import InertialOptimizer
                iopt = InertialOptimizer(local)
                wins = [synth.inertial_window(seed=s) for s in range(8)]
                for _ in range(2):
                    iopt.LocalInertialBA(wins[0], 1.0, 10)
                iopt.LocalInertialBABatch(wins, 1.0, 10)
                t0 = time.perf_counter()
                for s in range(5):
                    gi = iopt.LocalInertialBA(wins[s], 1.0, 10)
                ti_one = (time.perf_counter() - t0) / 5
                t0 = time.perf_counter()
                iopt.LocalInertialBABatch(wins, 1.0, 10)
                ti_batch = (time.perf_counter() - t0) / 8
                t0 = time.perf_counter()
                w4 = wins[4]
                ri = po.liba(w4["state"], w4["fixed"], w4["point"], w4["edge_kf"], w4["edge_mp"], w4["obs"], w4["inv_sigma2"], w4["Tcb"],
                             w4["cam5"], w4["links"].view(po.LIBA_LINK), 1.0, 10)
                ti_cpu = time.perf_counter() - t0
                liba = {"workload": "LocalInertialBA window: 10 optimisable + 4 fixed keyframes, 10 inertial links, 2000 points, "
                                    f"{len(w4['edge_kf'])} edges, lambda 1e0, 10 iterations",
                        "ms_per_solve_e2e": 1e3 * ti_one, "ms_per_solve_batch8_e2e": 1e3 * ti_batch, "cpu_oracle_ms": 1e3 * ti_cpu, "cpu_cores": 1,
                        "iterations": int(gi["iterations"]), "oracle_iterations": int(ri["iterations"]),
                        "max_abs_state_diff_vs_oracle": float(np.abs(gi["state"] - ri["state"]).max()),
                        "max_abs_point_diff_vs_oracle": float(np.abs(gi["point"] - ri["point"]).max())}
                iopt.close()
            except Exception as exc:
                liba = {"error": repr(exc)}
        # ---- K9 brute-force Hamming 2-NN (BFMatcher.knnMatch of ComputeStereoFishEyeMatches) ----------------------------------
        knn = None
        if True:
            try:
                if args.extract_only:
                    raise RuntimeError("skipped (--extract-only)")
                from orb_slam3_detailed_comments_b200 import knnMatch2
                krng = np.random.default_rng(5)
                qs = [krng.integers(0, 256, (1200, 32), dtype=np.uint8) for _ in range(64)]
                ts = [krng.integers(0, 256, (1200, 32), dtype=np.uint8) for _ in range(64)]
                for _ in range(3):
                    knnMatch2(exs[0], qs, ts)
                t0 = time.perf_counter()
                for _ in range(5):
                    got = knnMatch2(exs[0], qs, ts)
                t_knn = (time.perf_counter() - t0) / 5
                t_cv = None
                try:
                    import cv2
                    cv2.setNumThreads(1)
                    bfm = cv2.BFMatcher(cv2.NORM_HAMMING)
                    t0 = time.perf_counter()
                    ref = bfm.knnMatch(qs[0], ts[0], 2)
                    t_cv = time.perf_counter() - t0
                    same = all(got[0][0][i, 0] == m[0].trainIdx and got[0][0][i, 1] == m[1].trainIdx for i, m in enumerate(ref))
                except ImportError:
                    same = None
                knn = {"workload": "64 pairs of 1200 x 1200 descriptors, knnMatch k = 2, host pointers", "ms_per_call_e2e": 1e3 * t_knn,
                       "pairs_per_s": 64 / t_knn, "cv2_bfmatcher_ms_per_pair_1_thread": None if t_cv is None else 1e3 * t_cv,
                       "first_pair_equals_cv2": same}
            except Exception as exc:
                knn = {"error": repr(exc)}
        # ---- PoseOptimization (twice per frame on the tracking thread): a batch of B frames --------------
        pose_opt = None
        try:
            if args.extract_only:
                raise RuntimeError("skipped (--extract-only)")
            from oracle import pyoracle as po
            from orb_slam3_detailed_comments_b200 import PoseOptimization, PoseOptimizationDevice
            prng = np.random.default_rng(99)

            def po_frame(n):
                Xc = np.stack([prng.uniform(-3, 3, n), prng.uniform(-2, 2, n), prng.uniform(2, 12, n)], 1)
                ax = prng.normal(size=3); ax /= np.linalg.norm(ax)
                q = np.concatenate([ax * np.sin(0.015), [np.cos(0.015)]]); tt = prng.normal(0, 0.05, 3)
                x, y, z, w = q
                R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                              [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                              [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
                Xw = (Xc - tt) @ R
                u, v = FX * Xc[:, 0] / Xc[:, 2] + CX, FY * Xc[:, 1] / Xc[:, 2] + CY
                obs = np.stack([u, v, u - BF / Xc[:, 2]], 1) + prng.normal(0, 0.7, (n, 3))
                obs[prng.random(n) < 0.2, 2] = -1
                bad = prng.random(n) < 0.1
                obs[bad, :2] += prng.normal(0, 30, (int(bad.sum()), 2))
                return dict(pose=np.array([0, 0, 0, 1, 0, 0, 0], np.float32), world_pos=Xw.astype(np.float32), obs=obs.astype(np.float32),
                            inv_sigma2=(1 / (1.2 ** prng.integers(0, 8, n)) ** 2).astype(np.float32))
            cam5 = [FX, FY, CX, CY, BF]
            res = {}
            for tag, n_e in (("after_motion_model_550_edges", 550), ("after_local_map_1200_edges", 1200)):
                frames = [po_frame(n_e) for _ in range(B)]
                PoseOptimization(ex, frames, cam5)
                t0 = time.perf_counter()
                for _ in range(5):
                    got = PoseOptimization(ex, frames, cam5)
                t_e2e = (time.perf_counter() - t0) / 5
                eoff = torch.tensor(np.arange(B + 1, dtype=np.int32) * n_e, device=dev)
                d_pose = T(np.stack([f["pose"] for f in frames])); d_xw = T(np.concatenate([f["world_pos"] for f in frames]))
                d_obs = T(np.concatenate([f["obs"] for f in frames])); d_w = T(np.concatenate([f["inv_sigma2"] for f in frames]))
                o_pose = torch.zeros((B, 7), dtype=torch.float64, device=dev); o_out = torch.zeros(B * n_e, dtype=torch.uint8, device=dev)
                o_inl = torch.zeros(B, dtype=torch.int32, device=dev)
                ea, eb2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                PoseOptimizationDevice(ex, B, eoff, d_pose, d_xw, d_obs, d_w, cam5, o_pose, o_out, o_inl)
                torch.cuda.synchronize()
                ea.record(streams[0])
                for _ in range(10):
                    PoseOptimizationDevice(ex, B, eoff, d_pose, d_xw, d_obs, d_w, cam5, o_pose, o_out, o_inl)
                eb2.record(streams[0])
                torch.cuda.synchronize()
                t_dev = ea.elapsed_time(eb2) / 10
                t0 = time.perf_counter()
                for f in frames[:8]:
                    r = po.pose_optimization(f["pose"], f["world_pos"], f["obs"], f["inv_sigma2"], np.float32(cam5))
                t_cpu = (time.perf_counter() - t0) / 8
                res[tag] = {"frames_per_call": B, "ms_per_call_device_resident": t_dev, "ms_per_call_e2e": 1e3 * t_e2e,
                            "cpu_oracle_ms_per_frame": 1e3 * t_cpu, "lm_iterations_per_frame": float(np.mean([g["iterations"] for g in got])),
                            "inliers_per_frame": float(np.mean([g["inliers"] for g in got]))}
            pose_opt = {"workload": "Optimizer::PoseOptimization, one CTA per frame, 80 % stereo / 20 % monocular edges, 10 % outliers",
                        **res, "in_step": in_step}
        except Exception as exc:
            pose_opt = {"error": repr(exc)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": "config 3: stereo 640x480, 1200 features per eye; per frame: ORBextractor L+R, "
                                       "ComputeStereoMatches, SearchByProjection(cur,last,th=15), SearchByProjection(F,"
                                       "local map points,th=3)",
                           "frames_per_step_per_gpu": B, "images_per_step_per_gpu": nimg,
                           "pipeline": f"{NH} extractor handles / CUDA streams; value: one host thread; e2e: one host thread per handle",
                           "queries_per_frame": {"last_frame": nq_last / B, "local_map": nq_loc / B},
                           "matches_per_frame": {"last_frame": float(nm_host[:B].mean()), "local_map": float(nm_host[B:].mean())},
                           "l2": f"input pool of {pool_batches} batches = {pool_batches * nimg * W * H / 1e6:.0f} MB > 126 MB L2, "
                                 "intermediates rewritten every step",
                           "keypoints_per_image": n_kp, "fast_candidates_per_image": n_cand,
                           "kernel_variants": {"quadtree": int(os.environ.get("ORB_QT_VARIANT", "1") != "0"),
                                               "stereo": int(os.environ.get("ORB_STEREO_VARIANT", "1") != "0"),
                                               "fast": int(os.environ.get("ORB_FAST_VARIANT", "1") != "0")}},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_step),
                        "d2h_bytes_per_step": int(d2h // args.steps)},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "lba": lba, "inertial_ba": liba, "hamming_knn": knn, "pose_optimization": pose_opt}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

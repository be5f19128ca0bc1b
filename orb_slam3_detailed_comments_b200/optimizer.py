"""Optimizer -- host-side mirror of ORB_SLAM3::Optimizer::LocalBundleAdjustment's numeric core
(reference include/Optimizer.h:59, src/Optimizer.cc:1740-2188) over the flat lba_problem layout."""
import ctypes as C

import numpy as np

from . import _native as N


class Optimizer:
    """One device workspace; `LocalBundleAdjustment(problem)` and a batched form (one problem per CTA)."""

    def __init__(self, device=0):
        self._L = N.lib()
        self._h = C.c_void_p()
        N.check(self._L.lba_create(device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.lba_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _pack(pr, lambda_init, max_iters, keep):
        arrs = dict(pose=np.ascontiguousarray(pr["pose"], np.float64), fixed=np.ascontiguousarray(pr["fixed"], np.uint8),
                    point=np.ascontiguousarray(pr["point"], np.float64), edge_kf=np.ascontiguousarray(pr["edge_kf"], np.int32),
                    edge_mp=np.ascontiguousarray(pr["edge_mp"], np.int32), obs=np.ascontiguousarray(pr["obs"], np.float64),
                    inv_sigma2=np.ascontiguousarray(pr["inv_sigma2"], np.float64))
        cam = np.asarray(pr["cam5"], np.float64)
        out = dict(pose=np.zeros_like(arrs["pose"]), point=np.zeros_like(arrs["point"]),
                   edge_chi2=np.zeros(len(arrs["edge_kf"]), np.float64), edge_depth_pos=np.zeros(len(arrs["edge_kf"]), np.uint8))
        keep.append((arrs, out))
        p = N.lba_problem(len(arrs["pose"]), len(arrs["point"]), len(arrs["edge_kf"]), *[N.ptr(arrs[k]) for k in
                          ("pose", "fixed", "point", "edge_kf", "edge_mp", "obs", "inv_sigma2")], *cam.tolist(),
                          float(lambda_init), int(max_iters))
        r = N.lba_result(N.ptr(out["pose"]), N.ptr(out["point"]), N.ptr(out["edge_chi2"]), N.ptr(out["edge_depth_pos"]),
                         0, 0, 0.0, 0.0, 0.0)
        return p, r, out

    def LocalBundleAdjustmentBatch(self, problems, lambda_init=0.0, max_iters=10, stop_flag=None):
        keep, ps, rs, outs = [], [], [], []
        for pr in problems:
            p, r, o = self._pack(pr, lambda_init, max_iters, keep)
            ps.append(p); rs.append(r); outs.append(o)
        P = (N.lba_problem * len(ps))(*ps)
        R = (N.lba_result * len(rs))(*rs)
        sf = None if stop_flag is None else stop_flag.ctypes.data_as(C.c_void_p)
        N.check(self._L.lba_solve_batch(self._h, len(ps), P, R, sf))
        for o, r in zip(outs, R):
            o.update(iterations=r.iterations, trials=r.trials, lambda_=r.lambda_, chi2=r.chi2, chi2_init=r.chi2_initial)
        return outs

    def LocalBundleAdjustment(self, problem, lambda_init=0.0, max_iters=10, stop_flag=None):
        return self.LocalBundleAdjustmentBatch([problem], lambda_init, max_iters, stop_flag)[0]


def PoseOptimization(extractor, frames, cam5):
    """Optimizer::PoseOptimization(Frame*) (src/Optimizer.cc:55-412) for a batch of frames on the extractor's stream.
    frames: list of dicts(pose[7] float32 Tcw, world_pos[n][3], obs[n][3] (obs[:, 2] < 0 => monocular), inv_sigma2[n]);
    cam5 = fx fy cx cy bf.  Returns a list of dicts(pose float64[7], outlier uint8[n], inliers, rounds, iterations, trials)."""
    L = N.lib()
    nf = len(frames)
    eoff = np.zeros(nf + 1, np.int32)
    eoff[1:] = np.cumsum([len(f["world_pos"]) for f in frames])
    ne = int(eoff[-1])
    cat = lambda key, width: np.ascontiguousarray(np.concatenate([np.asarray(f[key], np.float32).reshape(-1, width) for f in frames])
                                                  if ne else np.zeros((0, width)), np.float32)
    pose = np.ascontiguousarray(np.stack([np.asarray(f["pose"], np.float32) for f in frames]))
    xw, obs, w = cat("world_pos", 3), cat("obs", 3), cat("inv_sigma2", 1)
    fx, fy, cx, cy, bf = [float(np.float32(v)) for v in cam5]
    p = N.orbo_pose_problems(nf, 0, N.ptr(eoff), N.ptr(pose), N.ptr(xw) if ne else None, N.ptr(obs) if ne else None,
                             N.ptr(w) if ne else None, fx, fy, cx, cy, bf, 0)
    pose_out = np.zeros((nf, 7), np.float64)
    outl = np.zeros(max(ne, 1), np.uint8)
    inl = np.zeros(nf, np.int32)
    stats = np.zeros((nf, 4), np.int32)
    N.check(L.orbo_pose_optimization(extractor._h, C.byref(p), N.ptr(pose_out), N.ptr(outl), N.ptr(inl), N.ptr(stats)))
    return [dict(pose=pose_out[f], outlier=outl[eoff[f]:eoff[f + 1]], inliers=int(inl[f]), rounds=int(stats[f, 0]),
                 iterations=int(stats[f, 1]), trials=int(stats[f, 2])) for f in range(nf)]


def PoseOptimizationDevice(extractor, n_frames, edge_offset, pose, world_pos, obs, inv_sigma2, cam5, pose_out, outlier_out, inliers_out):
    """Device-resident form (CUDA torch tensors, no synchronisation): see include/orbslam3_b200.h."""
    L = N.lib()
    fx, fy, cx, cy, bf = [float(np.float32(v)) for v in cam5]
    dp = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    p = N.orbo_pose_problems(n_frames, 1, dp(edge_offset), dp(pose), dp(world_pos), dp(obs), dp(inv_sigma2), fx, fy, cx, cy, bf,
                             int(inv_sigma2.numel()))
    N.check(L.orbo_pose_optimization(extractor._h, C.byref(p), dp(pose_out), dp(outlier_out), dp(inliers_out), None))


def PoseEdgesDevice(extractor, n_frames, frame_image, world_pos, edge_offset_out, edge_feature_out, world_pos_out, obs_out, inv_sigma2_out,
                    feature_match=None, query_offset=None, query_match=None, query_world_pos=None):
    """PoseOptimization's correspondence walk (Optimizer.cc:104-290) over device-resident search outputs (CUDA torch tensors)."""
    L = N.lib()
    dp = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    src = N.orbo_edge_source(n_frames, dp(frame_image), dp(feature_match), dp(query_offset), dp(query_match), dp(world_pos), dp(query_world_pos))
    N.check(L.orbo_pose_edges(extractor._h, C.byref(src), dp(edge_offset_out), dp(edge_feature_out), dp(world_pos_out), dp(obs_out),
                              dp(inv_sigma2_out)))


def PoseOptimizationFrames(extractor, frame_image, pose, world_pos, cam5, feature_match=None, query_offset=None, query_match=None,
                           total_rows=None, out=None):
    """PoseOptimization of frames of the extractor's last batch straight from a search's host result arrays
    (orbo_pose_optimization_frames).  Returns (pose float64[n][7], feature_outlier uint8[total_rows] or None, inliers int32[n])."""
    L = N.lib()
    fi = np.ascontiguousarray(frame_image, np.int32)
    nf = len(fi)
    pose = np.ascontiguousarray(pose, np.float32)
    xw = np.ascontiguousarray(world_pos, np.float32)
    fm = None if feature_match is None else np.ascontiguousarray(feature_match, np.int32)
    qo = None if query_offset is None else np.ascontiguousarray(query_offset, np.int32)
    qm = None if query_match is None else np.ascontiguousarray(query_match, np.int32)
    P = lambda a: None if a is None else N.ptr(a)
    fx, fy, cx, cy, bf = [float(np.float32(v)) for v in cam5]
    m = N.orbo_frame_matches(nf, P(fi), P(pose), P(fm), P(qo), P(qm), P(xw), len(xw), fx, fy, cx, cy, bf)
    if out is None:
        out = (np.zeros((nf, 7), np.float64), None if total_rows is None else np.zeros(max(total_rows, 1), np.uint8), np.zeros(nf, np.int32))
    N.check(L.orbo_pose_optimization_frames(extractor._h, C.byref(m), P(out[0]), P(out[1]), P(out[2])))
    return out


def pack_inertial_problem(pr, lambda_init, max_iters, keep):
    """dict(state[nKF][21], fixed, point, edge_kf, edge_mp, obs, inv_sigma2, links (N.LIBA_LINK records), Tcb[12], cam5) ->
    (liba_problem, liba_result, out dict).  `keep` receives the arrays the structs point into."""
    c = np.ascontiguousarray
    arrs = dict(state=c(pr["state"], np.float64), fixed=c(pr["fixed"], np.uint8), point=c(pr["point"], np.float64).reshape(-1, 3),
                edge_kf=c(pr["edge_kf"], np.int32), edge_mp=c(pr["edge_mp"], np.int32), obs=c(pr["obs"], np.float64).reshape(-1, 3),
                inv_sigma2=c(pr["inv_sigma2"], np.float64), links=c(np.asarray(pr["links"]).view(N.LIBA_LINK) if
                                                                       np.asarray(pr["links"]).dtype.itemsize == N.LIBA_LINK.itemsize
                                                                       else pr["links"], N.LIBA_LINK).reshape(-1))
    ne, nl = len(arrs["edge_kf"]), len(arrs["links"])
    out = dict(state=np.zeros_like(arrs["state"]), point=np.zeros_like(arrs["point"]), edge_chi2=np.zeros(max(ne, 1)),
               link_chi2=np.zeros((max(nl, 1), 3)), edge_depth_pos=np.zeros(max(ne, 1), np.uint8))
    keep.append((arrs, out))
    P = lambda a: N.ptr(a) if a.size else None
    p = N.liba_problem(len(arrs["state"]), len(arrs["point"]), ne, nl, N.ptr(arrs["state"]), N.ptr(arrs["fixed"]), P(arrs["point"]),
                       P(arrs["edge_kf"]), P(arrs["edge_mp"]), P(arrs["obs"]), P(arrs["inv_sigma2"]), P(arrs["links"]),
                       (C.c_double * 12)(*np.asarray(pr["Tcb"], np.float64).reshape(-1).tolist()),
                       *np.asarray(pr["cam5"], np.float64).tolist(), float(lambda_init), int(max_iters))
    r = N.liba_result(N.ptr(out["state"]), N.ptr(out["point"]), N.ptr(out["edge_chi2"]), N.ptr(out["link_chi2"]),
                      N.ptr(out["edge_depth_pos"]), 0, 0, 0.0, 0.0, 0.0, 0.0)
    return p, r, out


def finish_inertial_result(pr_arrays, out, r):
    ne, nl = len(pr_arrays["edge_kf"]), len(pr_arrays["links"])
    out["edge_chi2"] = out["edge_chi2"][:ne]
    out["link_chi2"] = out["link_chi2"][:nl]
    out["edge_depth_pos"] = out["edge_depth_pos"][:ne]
    out.update(iterations=r.iterations, trials=r.trials, lambda_=r.lambda_, chi2=r.chi2, chi2_init=r.chi2_initial, chi2_last=r.chi2_last_trial)
    return out


class InertialOptimizer:
    """Optimizer::LocalInertialBA's numeric core (include/Optimizer.h:63, src/Optimizer.cc:2203-2812) over the flat liba_problem
    layout, one CTA per window.  GPU parity: tests/test_liba_gpu.py."""

    def __init__(self, device=0):
        self._L = N.lib()
        self._h = C.c_void_p()
        N.check(self._L.liba_create(device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.liba_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def LocalInertialBABatch(self, problems, lambda_init=1.0, max_iters=10):
        keep, ps, rs, outs = [], [], [], []
        for pr in problems:
            p, r, o = pack_inertial_problem(pr, lambda_init, max_iters, keep)
            ps.append(p); rs.append(r); outs.append(o)
        Pa = (N.liba_problem * len(ps))(*ps)
        Ra = (N.liba_result * len(rs))(*rs)
        N.check(self._L.liba_solve(self._h, len(ps), Pa, Ra))
        return [finish_inertial_result(k[0], o, r) for k, o, r in zip(keep, outs, Ra)]

    def LocalInertialBA(self, problem, lambda_init=1.0, max_iters=10):
        return self.LocalInertialBABatch([problem], lambda_init, max_iters)[0]


def link_information(C15, oldest=False):
    """EdgeInertial / EdgeGyroRW / EdgeAccRW information matrices from IMU::Preintegrated::C (liba_link_information; host-only)."""
    Cm = np.ascontiguousarray(C15, np.float32).reshape(15, 15)
    info, ig, ia = np.zeros(81), np.zeros(9), np.zeros(9)
    N.check(N.lib().liba_link_information(N.ptr(Cm), int(bool(oldest)), N.ptr(info), N.ptr(ig), N.ptr(ia)))
    return info.reshape(9, 9), ig.reshape(3, 3), ia.reshape(3, 3)

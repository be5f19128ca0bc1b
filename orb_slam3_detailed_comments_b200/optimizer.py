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

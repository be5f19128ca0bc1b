"""Batched MapPoint maintenance (reference src/MapPoint.cc:438-520, 567-640): host-side mirror over the flat observation lists."""
import numpy as np

from . import _native as N


def ComputeDistinctiveDescriptors(extractor, obs_offset, obs_desc):
    """best observation index per map point (-1 = no observations); obs_desc: [n_obs][32] uint8 in mObservations order."""
    oo = np.ascontiguousarray(obs_offset, np.int32)
    d = np.ascontiguousarray(obs_desc, np.uint8).reshape(-1, 32)
    n = len(oo) - 1
    best = np.full(max(n, 1), -1, np.int32)
    N.check(N.lib().orbp_distinctive_descriptors(extractor._h, n, N.ptr(oo), N.ptr(d) if len(d) else None, N.ptr(best)))
    return best[:n]


def UpdateNormalAndDepth(extractor, obs_offset, obs_center, world_pos, ref_center, ref_level, normal=None, max_dist=None, min_dist=None):
    """mNormalVector, mfMaxDistance, mfMinDistance per map point (points without observations keep the values passed in)."""
    oo = np.ascontiguousarray(obs_offset, np.int32)
    n = len(oo) - 1
    f32 = lambda a, shape: np.zeros(shape, np.float32) if a is None else np.ascontiguousarray(a, np.float32).copy()
    oc = np.ascontiguousarray(obs_center, np.float32).reshape(-1, 3)
    pos, rc = np.ascontiguousarray(world_pos, np.float32), np.ascontiguousarray(ref_center, np.float32)
    lv = np.ascontiguousarray(ref_level, np.int32)
    nrm, mx, mn = f32(normal, (max(n, 1), 3)), f32(max_dist, max(n, 1)), f32(min_dist, max(n, 1))
    N.check(N.lib().orbp_update_normal_and_depth(extractor._h, n, N.ptr(oo), N.ptr(oc) if len(oc) else None, N.ptr(pos), N.ptr(rc), N.ptr(lv),
                                                 N.ptr(nrm), N.ptr(mx), N.ptr(mn)))
    return nrm[:n], mx[:n], mn[:n]

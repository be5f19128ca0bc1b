"""ORBmatcher -- host-side mirror of ORB_SLAM3::ORBmatcher (reference include/ORBmatcher.h:40-87) for the
projection-search entry points that sit on the tracking thread, in batched (many frames per call) form."""
import ctypes as C

import numpy as np

from . import _native as N

TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30   # ORBmatcher.cc:35-37


def camera(fx, fy, cx, cy, bf, b, width, height):
    """Frame intrinsics + image bounds of a rectified (zero-distortion) camera (Frame.cc:1084-1090)."""
    return N.orbm_camera(fx, fy, cx, cy, bf, b, 0.0, float(width), 0.0, float(height))


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, np.float32)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, np.int32)


def _u8(a):
    return None if a is None else np.ascontiguousarray(a, np.uint8)


def _is_dev(a):
    return hasattr(a, "data_ptr") and getattr(a, "is_cuda", False)


def _dptr(a):
    return None if a is None else C.c_void_p(a.data_ptr())


class ORBmatcher:
    """ORBmatcher(nnratio=0.6, checkOri=True)  (ORBmatcher.h:42).

    Array arguments are host numpy arrays (the call uploads them and returns numpy results) or, all of them,
    CUDA torch tensors of the right dtype (device-resident replay: nothing is copied, outputs are written to
    the `out_*` tensors and the call does not synchronise)."""

    def __init__(self, nnratio=0.6, checkOri=True):
        self.mfNNratio = float(nnratio)
        self.mbCheckOrientation = bool(checkOri)
        self._L = N.lib()

    def SearchByProjection(self, extractor, cam, frame_image, query_offset, proj_x, proj_y, proj_xr, level, view_cos,
                           desc, th=1.0, bFarPoints=False, thFarPoints=50.0, track_depth=None, feature_claimed=None, out=None,
                           in_view=None):
        """SearchByProjection(Frame&, vector<MapPoint*>, th, bFarPoints, thFarPoints) (ORBmatcher.cc:45-239) for
        several frames of `extractor`'s last batch.  Returns (match[nq] feature index or -1, nmatches[n_frames])."""
        if _is_dev(proj_x):
            raise TypeError("use SearchByProjectionDevice for CUDA tensors")
        fi, qo = _i32(frame_image), _i32(query_offset)
        arrs = [_f32(proj_x), _f32(proj_y), _f32(proj_xr), _i32(level), _f32(view_cos), _f32(track_depth), _u8(desc),
                _u8(feature_claimed), _u8(in_view)]
        q = N.orbm_local_queries(len(fi), 0, N.ptr(fi), N.ptr(qo), *[N.ptr(a) for a in arrs])
        nq = int(qo[-1])
        match, nm = out if out is not None else (np.full(max(nq, 1), -1, np.int32), np.zeros(len(fi), np.int32))
        N.check(self._L.orbm_search_local_points(extractor._h, C.byref(cam), C.byref(q), float(th), self.mfNNratio,
                                                 1 if bFarPoints else 0, float(thFarPoints), N.ptr(match), N.ptr(nm)))
        return match[:nq], nm

    def SearchByProjectionLastFrame(self, extractor, cam, frame_image, query_offset, Tcw, direction, world_pos,
                                    last_octave, last_angle, desc, obs_positive, th, total_rows, out=None):
        """SearchByProjection(Frame& cur, const Frame& last, th, bMono) (ORBmatcher.cc:1950-2184).
        Returns (feature_match[total_rows] query index or -1, nmatches[n_frames])."""
        fi, qo = _i32(frame_image), _i32(query_offset)
        arrs = [_f32(Tcw), _i32(direction), _f32(world_pos), _i32(last_octave), _f32(last_angle), _u8(desc),
                _u8(obs_positive)]
        q = N.orbm_last_queries(len(fi), 0, N.ptr(fi), N.ptr(qo), *[N.ptr(a) for a in arrs])
        fm, nm = out if out is not None else (np.full(max(total_rows, 1), -1, np.int32), np.zeros(len(fi), np.int32))
        N.check(self._L.orbm_search_last_frame(extractor._h, C.byref(cam), C.byref(q), float(th),
                                               1 if self.mbCheckOrientation else 0, N.ptr(fm), N.ptr(nm)))
        return fm[:total_rows], nm

    def SearchByBoW(self, extractor, frame_image, query_offset, query_node, query_angle, desc, feature_node, total_rows):
        """SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (ORBmatcher.cc:259-493) for several frames.
        Returns (feature_match[total_rows] query index or -1, nmatches[n_frames])."""
        fi, qo = _i32(frame_image), _i32(query_offset)
        arrs = [_i32(query_node), _f32(query_angle), _u8(desc), _i32(feature_node)]
        q = N.orbm_bow_queries(len(fi), 0, N.ptr(fi), N.ptr(qo), *[N.ptr(a) for a in arrs])
        fm = np.full(max(total_rows, 1), -1, np.int32)
        nm = np.zeros(len(fi), np.int32)
        N.check(self._L.orbm_search_bow(extractor._h, C.byref(q), self.mfNNratio, 1 if self.mbCheckOrientation else 0,
                                        N.ptr(fm), N.ptr(nm)))
        return fm[:total_rows], nm

    def SearchForInitialization(self, extractor, cam, kp1, desc1, vbPrevMatched, windowSize=10, kp2=None, desc2=None, target_image=0):
        """SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (ORBmatcher.cc:734-890).  F2 = host arrays
        (kp2, desc2) or image target_image of the extractor's last batch.  Returns (vnMatches12[n1], nmatches)."""
        kp1, d1 = np.ascontiguousarray(kp1), _u8(desc1)
        pm = np.ascontiguousarray(vbPrevMatched, np.float32).reshape(-1, 2)
        if kp2 is not None:
            kp2, d2 = np.ascontiguousarray(kp2), _u8(desc2)
            q = N.orbm_init_queries(len(kp1), N.ptr(kp1), N.ptr(d1), N.ptr(pm), len(kp2), N.ptr(kp2), N.ptr(d2), -1)
        else:
            q = N.orbm_init_queries(len(kp1), N.ptr(kp1), N.ptr(d1), N.ptr(pm), 0, None, None, int(target_image))
        m = np.full(max(len(kp1), 1), -1, np.int32)
        nm = C.c_int32(0)
        N.check(self._L.orbm_search_initialization(extractor._h, C.byref(cam), C.byref(q), int(windowSize), self.mfNNratio,
                                                   1 if self.mbCheckOrientation else 0, N.ptr(m), C.byref(nm)))
        return m[:len(kp1)], nm.value

    def SearchByBoWKeyFrames(self, extractor, pairs):
        """SearchByBoW(pKF1, pKF2, vpMatches12) (ORBmatcher.cc:892-1043) for several keyframe pairs.  pairs: list of dicts
        (kp2, desc2, node2, valid2, query_node, query_angle, desc1) -- queries in FeatureVector merge order.
        Returns ([match12 per pair], nmatches[n_pairs])."""
        nt = len(pairs)
        foff, qoff = np.zeros(nt + 1, np.int32), np.zeros(nt + 1, np.int32)
        foff[1:] = np.cumsum([len(p["kp2"]) for p in pairs])
        qoff[1:] = np.cumsum([len(p["query_node"]) for p in pairs])
        cat = lambda key, dt: np.ascontiguousarray(np.concatenate([np.asarray(p[key], dt) for p in pairs]))
        kp2 = np.ascontiguousarray(np.concatenate([p["kp2"] for p in pairs]))
        a = [cat("desc2", np.uint8), cat("node2", np.int32), cat("valid2", np.uint8)]
        b = [cat("query_node", np.int32), cat("query_angle", np.float32), cat("desc1", np.uint8)]
        q = N.orbm_bow_kf_queries(nt, N.ptr(foff), N.ptr(kp2), *[N.ptr(x) for x in a], N.ptr(qoff), *[N.ptr(x) for x in b])
        m = np.full(max(int(qoff[-1]), 1), -1, np.int32)
        nm = np.zeros(nt, np.int32)
        N.check(self._L.orbm_search_bow_keyframes(extractor._h, C.byref(q), self.mfNNratio, 1 if self.mbCheckOrientation else 0,
                                                  N.ptr(m), N.ptr(nm)))
        return [m[qoff[t]:qoff[t + 1]] for t in range(nt)], nm

    def SearchForTriangulation(self, extractor, kp1, desc1, node1, stereo1, kp2, desc2, node2, valid2, stereo2, F12, epipole2,
                               bCoarse=False):
        """SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse) (ORBmatcher.cc:1045-1323), single camera.
        Returns (match12[nq] KF2 index or -1, nmatches)."""
        kp1, kp2 = np.ascontiguousarray(kp1), np.ascontiguousarray(kp2)
        a = [_u8(desc1), _i32(node1), _u8(stereo1)]
        b = [_u8(desc2), _i32(node2), _u8(valid2), _u8(stereo2)]
        t = N.orbm_triangulation(len(kp1), len(kp2), N.ptr(kp1), *[N.ptr(x) for x in a], N.ptr(kp2), *[N.ptr(x) for x in b],
                                 (C.c_float * 9)(*np.asarray(F12, np.float32).reshape(-1).tolist()),
                                 (C.c_float * 2)(*np.asarray(epipole2, np.float32).tolist()), 1 if bCoarse else 0,
                                 1 if self.mbCheckOrientation else 0)
        m = np.full(max(len(kp1), 1), -1, np.int32)
        nm = C.c_int32(0)
        N.check(self._L.orbm_search_triangulation(extractor._h, C.byref(t), N.ptr(m), C.byref(nm)))
        return m[:len(kp1)], nm.value

    # ---- projection searches into keyframes (map state in host memory) --------------------------------------------
    FUSE_POSE, FUSE_SIM3, PROJ_SIM3, PROJ_RELOC, SIM3_ONEWAY = 0, 1, 2, 3, 4

    def _search_keyframe(self, extractor, cam, variant, targets, queries, th, hamming_max, target_images=None):
        """targets: list of dicts(kp, desc, uright|None, claimed|None, Tcw[7], Ow[3]) -- or, with target_images, dicts
        (claimed|None, Tcw, Ow) for images of the extractor's last batch; queries: list (one per target) of dicts
        (world_pos, normal|None, max_dist, min_dist, desc, angle|None).  Returns ([match per target], [nmatches])."""
        nt = len(targets)
        qoff = np.zeros(nt + 1, np.int32)
        qoff[1:] = np.cumsum([len(q["world_pos"]) for q in queries])
        cat = lambda key, dt, width: (np.ascontiguousarray(np.concatenate([np.asarray(q[key], dt).reshape(-1, width) for q in queries]))
                                      if all(q.get(key) is not None for q in queries) else None)
        xw, nr = cat("world_pos", np.float32, 3), cat("normal", np.float32, 3)
        mx, mn, ang = cat("max_dist", np.float32, 1), cat("min_dist", np.float32, 1), cat("angle", np.float32, 1)
        qd = cat("desc", np.uint8, 32)
        tcw = np.ascontiguousarray(np.stack([np.asarray(t["Tcw"], np.float32) for t in targets]))
        ow = np.ascontiguousarray(np.stack([np.asarray(t["Ow"], np.float32) for t in targets]))
        s8 = (np.ascontiguousarray(np.stack([np.asarray(t["Sim3"], np.float32) for t in targets]))
              if all(t.get("Sim3") is not None for t in targets) else None)
        keep = [qoff, xw, nr, mx, mn, ang, qd, tcw, ow, s8]
        if target_images is None:
            foff = np.zeros(nt + 1, np.int32)
            foff[1:] = np.cumsum([len(t["kp"]) for t in targets])
            kp = np.ascontiguousarray(np.concatenate([t["kp"] for t in targets]))
            de = np.ascontiguousarray(np.concatenate([np.asarray(t["desc"], np.uint8).reshape(-1, 32) for t in targets]))
            ur = (np.ascontiguousarray(np.concatenate([np.asarray(t["uright"], np.float32) for t in targets]))
                  if all(t.get("uright") is not None for t in targets) else None)
            cl = (np.ascontiguousarray(np.concatenate([np.asarray(t["claimed"], np.uint8) for t in targets]))
                  if all(t.get("claimed") is not None for t in targets) else None)
            timg = None
        else:
            foff = kp = de = ur = None
            timg = _i32(target_images)
            cl = targets[0].get("claimed_rows")
            cl = None if cl is None else _u8(cl)
        keep += [foff, kp, de, ur, cl, timg]
        P = lambda a: None if a is None else N.ptr(a)
        q = N.orbm_kf_queries(nt, P(timg), P(foff), P(kp), P(de), P(ur), P(cl), P(tcw), P(ow), P(s8), P(qoff), P(xw), P(nr), P(mx), P(mn),
                              P(qd), P(ang))
        nq = int(qoff[-1])
        m = np.full(max(nq, 1), -1, np.int32)
        nm = np.zeros(nt, np.int32)
        N.check(self._L.orbm_search_keyframe(extractor._h, C.byref(cam), C.byref(q), int(variant), float(th), float(hamming_max),
                                             1 if self.mbCheckOrientation else 0, N.ptr(m), N.ptr(nm)))
        return [m[qoff[t]:qoff[t + 1]] for t in range(nt)], nm

    def Fuse(self, extractor, cam, targets, queries, th=3.0):
        """Fuse(pKF, vpMapPoints, th, bRight=false) (ORBmatcher.cc:1325-1544) for several keyframes at once: the search
        part; the caller applies Replace / AddObservation in query order from bestIdx."""
        return self._search_keyframe(extractor, cam, self.FUSE_POSE, targets, queries, th, 50.0)

    def FuseSim3(self, extractor, cam, targets, queries, th):
        """Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (ORBmatcher.cc:1546-1687)."""
        return self._search_keyframe(extractor, cam, self.FUSE_SIM3, targets, queries, th, 50.0)

    def SearchByProjectionSim3(self, extractor, cam, targets, queries, th, ratioHamming=1.0):
        """SearchByProjection(pKF, Scw, vpPoints, [vpPointsKFs,] vpMatched, [vpMatchedKF,] th, ratioHamming)
        (ORBmatcher.cc:495-732)."""
        return self._search_keyframe(extractor, cam, self.PROJ_SIM3, targets, queries, th, float(np.float32(50) * np.float32(ratioHamming)))

    def SearchBySim3(self, extractor, cam, kf1, kf2, S12, S21, mp1, mp2, th):
        """SearchBySim3(pKF1, pKF2, vpMatches12, S12, th) (ORBmatcher.cc:1689-1948).  kf1 / kf2: dicts(kp, desc, Tcw[7]);
        S12 / S21: 8 floats (quaternion x y z w, translation, scale) -- S21 = S12.inverse() from the caller's Sophus;
        mp1 / mp2: dicts(index: keyframe feature index of every good, not-yet-matched map point, world_pos, max_dist,
        min_dist, desc).  Returns (vnMatch12 dict {i1: i2} of the mutually consistent pairs, nFound)."""
        zero = np.zeros(3, np.float32)
        targets = [dict(kp=kf2["kp"], desc=kf2["desc"], Tcw=kf1["Tcw"], Ow=zero, Sim3=S21),      # KF1's points into KF2
                   dict(kp=kf1["kp"], desc=kf1["desc"], Tcw=kf2["Tcw"], Ow=zero, Sim3=S12)]      # KF2's points into KF1
        (m12, m21), _ = self._search_keyframe(extractor, cam, self.SIM3_ONEWAY, targets, [mp1, mp2], th, 100.0)
        vn1 = {int(i1): int(i2) for i1, i2 in zip(mp1["index"], m12) if i2 >= 0}
        vn2 = {int(i2): int(i1) for i2, i1 in zip(mp2["index"], m21) if i1 >= 0}
        found = {i1: i2 for i1, i2 in sorted(vn1.items()) if vn2.get(i2, -1) == i1}
        return found, len(found)

    def SearchByProjectionReloc(self, extractor, cam, targets, queries, th, ORBdist, target_images=None):
        """SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:2196-2330); the frame is given
        as host arrays or (target_images) as an image of the extractor's last batch."""
        return self._search_keyframe(extractor, cam, self.PROJ_RELOC, targets, queries, th, float(ORBdist), target_images)

    # ---- device-resident forms (CUDA torch tensors; see the class docstring) -----------------------
    def SearchByProjectionDevice(self, extractor, cam, n_frames, frame_image, query_offset, proj_x, proj_y, proj_xr, level,
                                 view_cos, desc, out_match, out_nmatches, th=1.0, feature_claimed=None, in_view=None):
        q = N.orbm_local_queries(n_frames, 1, _dptr(frame_image), _dptr(query_offset), _dptr(proj_x), _dptr(proj_y),
                                 _dptr(proj_xr), _dptr(level), _dptr(view_cos), None, _dptr(desc), _dptr(feature_claimed), _dptr(in_view))
        N.check(self._L.orbm_search_local_points(extractor._h, C.byref(cam), C.byref(q), float(th), self.mfNNratio, 0, 0.0,
                                                 _dptr(out_match), _dptr(out_nmatches)))

    def SearchByProjectionLastFrameDevice(self, extractor, cam, n_frames, frame_image, query_offset, Tcw, direction,
                                          world_pos, last_octave, last_angle, desc, obs_positive, th, out_feature_match,
                                          out_nmatches):
        q = N.orbm_last_queries(n_frames, 1, _dptr(frame_image), _dptr(query_offset), _dptr(Tcw), _dptr(direction),
                                _dptr(world_pos), _dptr(last_octave), _dptr(last_angle), _dptr(desc), _dptr(obs_positive))
        N.check(self._L.orbm_search_last_frame(extractor._h, C.byref(cam), C.byref(q), float(th),
                                               1 if self.mbCheckOrientation else 0, _dptr(out_feature_match),
                                               _dptr(out_nmatches)))


def isInFrustum(extractor, cam, point_offset, Rcw, tcw, Ow, world_pos, normal, max_dist, min_dist, viewingCosLimit=0.5):
    """Frame::isInFrustum (Frame.cc:667-720) for the candidate map points of several frames (host arrays).
    Returns dict(in_view, proj_x, proj_y, proj_xr, level, view_cos, depth), one entry per candidate."""
    po = _i32(point_offset)
    nf, n = len(po) - 1, int(po[-1])
    a = [_f32(Rcw), _f32(tcw), _f32(Ow), _f32(world_pos), _f32(normal), _f32(max_dist), _f32(min_dist)]
    fp = N.orbf_frustum_points(nf, 0, N.ptr(po), *[N.ptr(x) for x in a], 0)
    m = max(n, 1)
    out = dict(in_view=np.zeros(m, np.uint8), proj_x=np.zeros(m, np.float32), proj_y=np.zeros(m, np.float32), proj_xr=np.zeros(m, np.float32),
               level=np.zeros(m, np.int32), view_cos=np.zeros(m, np.float32), depth=np.zeros(m, np.float32))
    N.check(N.lib().orbf_is_in_frustum(extractor._h, C.byref(cam), C.byref(fp), float(viewingCosLimit),
                                       *[N.ptr(out[k]) for k in ("in_view", "proj_x", "proj_y", "proj_xr", "level", "view_cos", "depth")]))
    return {k: v[:n] for k, v in out.items()}


def isInFrustumDevice(extractor, cam, n_frames, point_offset, Rcw, tcw, Ow, world_pos, normal, max_dist, min_dist, out, viewingCosLimit=0.5):
    """Device-resident form (CUDA torch tensors; `out` = dict of tensors with the keys isInFrustum returns); no synchronisation."""
    fp = N.orbf_frustum_points(n_frames, 1, _dptr(point_offset), _dptr(Rcw), _dptr(tcw), _dptr(Ow), _dptr(world_pos), _dptr(normal),
                               _dptr(max_dist), _dptr(min_dist), int(max_dist.numel()))
    N.check(N.lib().orbf_is_in_frustum(extractor._h, C.byref(cam), C.byref(fp), float(viewingCosLimit),
                                       *[_dptr(out[k]) for k in ("in_view", "proj_x", "proj_y", "proj_xr", "level", "view_cos", "depth")]))


def knnMatch2(extractor, query_sets, train_sets):
    """cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, 2) (Frame::ComputeStereoFishEyeMatches, src/Frame.cc:1553) for a batch of
    independent (query descriptors [nq][32], train descriptors [nt][32]) pairs on the extractor's stream.
    Returns a list of (idx[nq][2], dist[nq][2]) int32 arrays; -1 where the train set has fewer rows."""
    L = N.lib()
    qs = [np.ascontiguousarray(q, np.uint8).reshape(-1, 32) for q in query_sets]
    ts = [np.ascontiguousarray(t, np.uint8).reshape(-1, 32) for t in train_sets]
    assert len(qs) == len(ts)
    qo = np.zeros(len(qs) + 1, np.int32)
    to = np.zeros(len(ts) + 1, np.int32)
    qo[1:] = np.cumsum([len(q) for q in qs])
    to[1:] = np.cumsum([len(t) for t in ts])
    qd = np.ascontiguousarray(np.concatenate(qs)) if qo[-1] else np.zeros((1, 32), np.uint8)
    td = np.ascontiguousarray(np.concatenate(ts)) if to[-1] else np.zeros((1, 32), np.uint8)
    idx = np.zeros((max(int(qo[-1]), 1), 2), np.int32)
    dist = np.zeros((max(int(qo[-1]), 1), 2), np.int32)
    N.check(L.orbm_hamming_knn2(extractor._h, len(qs), N.ptr(qo), N.ptr(qd), N.ptr(to), N.ptr(td), N.ptr(idx), N.ptr(dist)))
    return [(idx[qo[i]:qo[i + 1]], dist[qo[i]:qo[i + 1]]) for i in range(len(qs))]

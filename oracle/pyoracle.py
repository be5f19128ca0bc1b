"""ctypes driver for the CPU oracle (oracle/_build/liborb_oracle.so).

TEST INFRASTRUCTURE: only tests/, bench.py's cpu_baseline / --impl reference leg and
__graft_entry__.smoke() may import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborb_oracle.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h", ".inc"))]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs)):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        L = _lib
        L.orc_extractor_create.restype = C.c_void_p
        L.orc_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_extractor_destroy.argtypes = [C.c_void_p]
        L.orc_extract.restype = C.c_int
        L.orc_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.orc_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.orc_level_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_level_pyramid.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_level_blurred.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_level_cands.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_level_kps.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_timings.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_resize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.orc_blur.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_fast_cell.restype = C.c_int
        L.orc_fast_cell.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_atan2.restype = C.c_float
        L.orc_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_ic_angle.restype = C.c_float
        L.orc_ic_angle.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_descriptor.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]
        L.orc_hamming.restype = C.c_int
        L.orc_hamming.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_pattern.restype = C.POINTER(C.c_int8)
        L.orc_distribute.restype = C.c_int
        L.orc_distribute.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_int]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleExtractor:
    """CPU restatement of ORB_SLAM3::ORBextractor (reference src/ORBextractor.cc)."""

    def __init__(self, nfeatures=1200, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self.h = C.c_void_p(self.L.orc_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th))
        sc, isc, s2, is2 = (np.zeros(nlevels, np.float32) for _ in range(4))
        q = np.zeros(nlevels, np.int32)
        um = np.zeros(16, np.int32)
        self.L.orc_tables(self.h, _p(sc), _p(isc), _p(s2), _p(is2), _p(q), _p(um))
        self.scale_factors, self.inv_scale_factors, self.level_sigma2, self.inv_level_sigma2 = sc, isc, s2, is2
        self.features_per_level, self.umax = q, um

    def __del__(self):
        try:
            self.L.orc_extractor_destroy(self.h)
        except Exception:
            pass

    def __call__(self, img, lapping=(0, 0)):
        img = np.ascontiguousarray(img, np.uint8)
        H, W = img.shape
        cap = max(4 * self.nfeatures, 4096)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        mono = self.L.orc_extract(self.h, _p(img), W, H, W, int(lapping[0]), int(lapping[1]), _p(kps), _p(desc),
                                  cap, C.byref(n))
        assert n.value <= cap
        return mono, kps[:n.value].copy(), desc[:n.value].copy()

    def level_size(self, l):
        w, h = C.c_int(), C.c_int()
        self.L.orc_level_size(self.h, l, C.byref(w), C.byref(h))
        return w.value, h.value

    def level_pyramid(self, l):
        w, h = self.level_size(l)
        a = np.zeros((h, w), np.uint8)
        self.L.orc_level_pyramid(self.h, l, _p(a))
        return a

    def level_blurred(self, l):
        w, h = self.level_size(l)
        a = np.zeros((h, w), np.uint8)
        return a if self.L.orc_level_blurred(self.h, l, _p(a)) else None

    def level_cands(self, l):
        cap = 1 << 17
        a = np.zeros((cap, 3), np.int32)
        n = self.L.orc_level_cands(self.h, l, _p(a), cap)
        assert n <= cap
        return a[:n].copy()

    def level_kps(self, l):
        cap = 1 << 14
        a = np.zeros(cap, KP_DTYPE)
        n = self.L.orc_level_kps(self.h, l, _p(a), cap)
        return a[:n].copy()

    def timings(self):
        t = np.zeros(6, np.float64)
        self.L.orc_timings(self.h, _p(t))
        return dict(zip(["pyramid", "fast", "quadtree", "angle", "blur", "descriptor"], t.tolist()))


def resize(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    d = np.zeros((dh, dw), np.uint8)
    lib().orc_resize(_p(src), src.shape[1], src.shape[0], _p(d), dw, dh)
    return d


def blur(src):
    src = np.ascontiguousarray(src, np.uint8)
    d = np.zeros_like(src)
    lib().orc_blur(_p(src), src.shape[1], src.shape[0], _p(d))
    return d


def fast_cell(win, thr):
    """cv::FAST(win, thr, nonmax=True) restated: returns (n,3) int32 rows (x, y, score)."""
    assert win.dtype == np.uint8 and win.strides[1] == 1
    cap = win.size
    out = np.zeros((max(cap, 1), 3), np.int32)
    n = lib().orc_fast_cell(C.c_void_p(win.ctypes.data), win.shape[1], win.shape[0], win.strides[0], thr, _p(out), cap)
    return out[:n].copy()


def atan2_deg(y, x):
    return float(lib().orc_atan2(float(y), float(x)))


def ic_angle(img, x, y, umax):
    img = np.ascontiguousarray(img, np.uint8)
    um = np.ascontiguousarray(umax, np.int32)
    return float(lib().orc_ic_angle(_p(img), img.shape[1], img.shape[0], int(x), int(y), _p(um)))


def descriptor(img, x, y, angle):
    img = np.ascontiguousarray(img, np.uint8)
    d = np.zeros(32, np.uint8)
    lib().orc_descriptor(_p(img), img.shape[1], img.shape[0], int(x), int(y), float(angle), _p(d))
    return d


def hamming(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return int(lib().orc_hamming(_p(a), _p(b)))


def pattern():
    return np.ctypeslib.as_array(lib().orc_pattern(), shape=(1024,)).copy()


def distribute(cands, minX, maxX, minY, maxY, N):
    cands = np.ascontiguousarray(cands, np.int32)
    out = np.zeros(max(len(cands), 1), np.int32)
    n = lib().orc_distribute(_p(cands), len(cands), minX, maxX, minY, maxY, N, _p(out), len(out))
    return out[:n].copy()


def stereo_matches(exL, exR, kpsL, descL, kpsR, descR, bf, b):
    """Frame::ComputeStereoMatches restated (reference src/Frame.cc:1102-1358).  exL/exR are the
    OracleExtractor objects that just produced (kpsL, descL)/(kpsR, descR) (their pyramids are read)."""
    L = lib()
    L.orc_stereo.restype = C.c_int
    L.orc_stereo.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                             C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    kpsL, kpsR = np.ascontiguousarray(kpsL), np.ascontiguousarray(kpsR)
    descL, descR = np.ascontiguousarray(descL), np.ascontiguousarray(descR)
    uR = np.zeros(len(kpsL), np.float32)
    dep = np.zeros(len(kpsL), np.float32)
    kept = L.orc_stereo(exL.h, exR.h, _p(kpsL), _p(descL), len(kpsL), _p(kpsR), _p(descR), len(kpsR), bf, b, _p(uR), _p(dep))
    return uR, dep, kept


def search_local(kps, desc, uright, bounds, scale_factors, projx, projy, projxr, level, viewcos, qdesc, th, nnratio,
                 claimed=None, trackdepth=None, far=False, th_far=0.0):
    """ORBmatcher::SearchByProjection(F, vpMapPoints, th, bFar, thFar) restated (ORBmatcher.cc:45-239)."""
    L = lib()
    L.orc_search_local.restype = C.c_int
    L.orc_search_local.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 8 + \
        [C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p]
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    kps, desc, uright = np.ascontiguousarray(kps), np.ascontiguousarray(desc), f32(uright)
    b4, sf = f32(bounds), f32(scale_factors)
    px, py, pxr, vc, td = f32(projx), f32(projy), f32(projxr), f32(viewcos), f32(trackdepth)
    lv = np.ascontiguousarray(level, np.int32)
    qd = np.ascontiguousarray(qdesc, np.uint8)
    cl = np.zeros(len(kps), np.uint8) if claimed is None else np.ascontiguousarray(claimed, np.uint8).copy()
    nq = len(px)
    match = np.full(max(nq, 1), -1, np.int32)
    P = lambda a: None if a is None else _p(a)
    nm = L.orc_search_local(_p(kps), _p(desc), P(uright), len(kps), _p(b4), _p(sf), nq, _p(px), _p(py), _p(pxr), _p(lv),
                            _p(vc), P(td), _p(qd), _p(cl), th, nnratio, 1 if far else 0, th_far, _p(match))
    return match[:nq], nm


def search_last(kps, desc, uright, bounds, scale_factors, cam6, Tcw7, direction, xw, last_octave, last_angle, qdesc,
                obs_pos, th, check_ori=True):
    """ORBmatcher::SearchByProjection(cur, last, th, bMono) restated (ORBmatcher.cc:1950-2184)."""
    L = lib()
    L.orc_search_last.restype = C.c_int
    L.orc_search_last.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_int] + [C.c_void_p] * 5 + \
        [C.c_float, C.c_int, C.c_void_p]
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    kps, desc, uright = np.ascontiguousarray(kps), np.ascontiguousarray(desc), f32(uright)
    b4, sf, c6, T7 = f32(bounds), f32(scale_factors), f32(cam6), f32(Tcw7)
    xw, ang = f32(xw), f32(last_angle)
    lo = np.ascontiguousarray(last_octave, np.int32)
    qd, ob = np.ascontiguousarray(qdesc, np.uint8), np.ascontiguousarray(obs_pos, np.uint8)
    fm = np.full(max(len(kps), 1), -1, np.int32)
    P = lambda a: None if a is None else _p(a)
    nm = L.orc_search_last(_p(kps), _p(desc), P(uright), len(kps), _p(b4), _p(sf), _p(c6), _p(T7), int(direction), len(lo),
                           _p(xw), _p(lo), _p(ang), _p(qd), _p(ob), th, 1 if check_ori else 0, _p(fm))
    return fm[:len(kps)], nm


def lba(pose, fixed, point, edge_kf, edge_mp, obs, inv_sigma2, cam5, lambda_init=0.0, max_iters=10):
    """Optimizer::LocalBundleAdjustment's g2o Levenberg core restated (Optimizer.cc:1859-2150).
    Returns dict(pose, point, edge_chi2, edge_depth_pos, iterations, lambda_, chi2, trials, chi2_init)."""
    L = lib()
    L.orc_lba.restype = C.c_int
    L.orc_lba.argtypes = [C.c_int] * 3 + [C.c_void_p] * 8 + [C.c_double, C.c_int] + [C.c_void_p] * 4
    pose = np.ascontiguousarray(pose, np.float64).copy()
    point = np.ascontiguousarray(point, np.float64).copy()
    fixed = np.ascontiguousarray(fixed, np.uint8)
    ekf, emp = np.ascontiguousarray(edge_kf, np.int32), np.ascontiguousarray(edge_mp, np.int32)
    obs, w = np.ascontiguousarray(obs, np.float64), np.ascontiguousarray(inv_sigma2, np.float64)
    cam5 = np.ascontiguousarray(cam5, np.float64)
    nE = len(ekf)
    chi = np.zeros(nE, np.float64)
    dpos = np.zeros(nE, np.uint8)
    stats = np.zeros(8, np.float64)
    it = L.orc_lba(len(pose), len(point), nE, _p(pose), _p(fixed), _p(point), _p(ekf), _p(emp), _p(obs), _p(w), _p(cam5),
                   float(lambda_init), int(max_iters), None, _p(chi), _p(dpos), _p(stats))
    return dict(pose=pose, point=point, edge_chi2=chi, edge_depth_pos=dpos, iterations=it, lambda_=stats[1],
                chi2=stats[2], trials=int(stats[3]), chi2_init=stats[4])


def search_bow(kps, desc, feat_node, qnode, qangle, qdesc, nnratio, check_ori=True):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) restated (ORBmatcher.cc:259-493), Nleft == -1."""
    L = lib()
    L.orc_search_bow.restype = C.c_int
    L.orc_search_bow.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_float, C.c_int, C.c_void_p]
    kps, desc = np.ascontiguousarray(kps), np.ascontiguousarray(desc)
    fn, qn = np.ascontiguousarray(feat_node, np.int32), np.ascontiguousarray(qnode, np.int32)
    qa, qd = np.ascontiguousarray(qangle, np.float32), np.ascontiguousarray(qdesc, np.uint8)
    fm = np.full(max(len(kps), 1), -1, np.int32)
    nm = L.orc_search_bow(_p(kps), _p(desc), _p(fn), len(kps), len(qn), _p(qn), _p(qa), _p(qd), nnratio, 1 if check_ori else 0, _p(fm))
    return fm[:len(kps)], nm


def search_triangulation(kp1, desc1, node1, stereo1, kp2, desc2, node2, valid2, stereo2, F12, ep2, scale_factors, sigma2,
                         coarse=False, check_ori=True):
    """ORBmatcher::SearchForTriangulation restated (ORBmatcher.cc:1045-1323), single-camera keyframes."""
    L = lib()
    L.orc_search_triangulation.restype = C.c_int
    L.orc_search_triangulation.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 9 + [C.c_int, C.c_int, C.c_void_p]
    c = np.ascontiguousarray
    kp1, kp2 = c(kp1), c(kp2)
    a = [c(desc1, np.uint8), c(node1, np.int32), c(stereo1, np.uint8)]
    b = [c(desc2, np.uint8), c(node2, np.int32), c(valid2, np.uint8), c(stereo2, np.uint8), c(F12, np.float32), c(ep2, np.float32),
         c(scale_factors, np.float32), c(sigma2, np.float32)]
    m = np.full(max(len(kp1), 1), -1, np.int32)
    nm = L.orc_search_triangulation(len(kp1), _p(kp1), *[_p(x) for x in a], len(kp2), _p(kp2), *[_p(x) for x in b],
                                    1 if coarse else 0, 1 if check_ori else 0, _p(m))
    return m[:len(kp1)], nm


def search_keyframe(variant, kps, desc, uright, bounds, scale_factors, inv_sigma2, log_scale_factor, cam6, Tcw7, Ow, xw, normal,
                    max_dist, min_dist, qdesc, qangle, claimed, th, thr, check_ori=True, sim3=None):
    """Fuse x2 / SearchByProjection(KF, Scw) / SearchByProjection(F, KF, set) restated (ORBmatcher.cc:1325-1687, 495-732,
    2196-2330).  Returns (match[nq], nmatches, claimed_after)."""
    L = lib()
    L.orc_search_keyframe.restype = C.c_int
    L.orc_search_keyframe.argtypes = [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_float] + \
        [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 7 + [C.c_float, C.c_float, C.c_int, C.c_void_p]
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    P = lambda a: None if a is None else _p(a)
    s8 = f32(sim3)
    kps, desc, uright = np.ascontiguousarray(kps), np.ascontiguousarray(desc), f32(uright)
    b4, sf, isg, c6, T7, ow = f32(bounds), f32(scale_factors), f32(inv_sigma2), f32(cam6), f32(Tcw7), f32(Ow)
    xw, nr, mx, mn, ang = f32(xw), f32(normal), f32(max_dist), f32(min_dist), f32(qangle)
    qd = np.ascontiguousarray(qdesc, np.uint8)
    cl = np.zeros(max(len(kps), 1), np.uint8) if claimed is None else np.ascontiguousarray(claimed, np.uint8).copy()
    nq = len(xw)
    m = np.full(max(nq, 1), -1, np.int32)
    nm = L.orc_search_keyframe(int(variant), _p(kps), _p(desc), P(uright), len(kps), _p(b4), _p(sf), _p(isg), len(sf),
                               float(log_scale_factor), _p(c6), _p(T7), _p(ow), P(s8), nq, _p(xw), P(nr), _p(mx), _p(mn), _p(qd), P(ang),
                               _p(cl), float(th), float(thr), 1 if check_ori else 0, _p(m))
    return m[:nq], nm, cl[:len(kps)]


def logf(x):
    """glibc logf of a float32 (what Frame.cc:121 / MapPoint.cc:688-721 call)."""
    import ctypes.util
    m = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    m.logf.restype = C.c_float
    m.logf.argtypes = [C.c_float]
    return m.logf(float(np.float32(x)))


def search_bow_kf(kp2, desc2, node2, valid2, qnode, qangle, desc1, nnratio, check_ori=True):
    """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) restated (ORBmatcher.cc:892-1043)."""
    L = lib()
    L.orc_search_bow_kf.restype = C.c_int
    L.orc_search_bow_kf.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_float, C.c_int, C.c_void_p]
    c = np.ascontiguousarray
    kp2 = c(kp2)
    a = [c(desc2, np.uint8), c(node2, np.int32), c(valid2, np.uint8)]
    b = [c(qnode, np.int32), c(qangle, np.float32), c(desc1, np.uint8)]
    m = np.full(max(len(b[0]), 1), -1, np.int32)
    nm = L.orc_search_bow_kf(_p(kp2), *[_p(x) for x in a], len(kp2), len(b[0]), *[_p(x) for x in b], nnratio, 1 if check_ori else 0, _p(m))
    return m[:len(b[0])], nm


def search_initialization(kp1, desc1, prev_matched, kp2, desc2, bounds, window_size, nnratio, check_ori=True):
    """ORBmatcher::SearchForInitialization restated (ORBmatcher.cc:734-890)."""
    L = lib()
    L.orc_search_initialization.restype = C.c_int
    L.orc_search_initialization.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                            C.c_float, C.c_int, C.c_void_p]
    c = np.ascontiguousarray
    kp1, kp2, d1, d2 = c(kp1), c(kp2), c(desc1, np.uint8), c(desc2, np.uint8)
    pm, b4 = c(prev_matched, np.float32), c(bounds, np.float32)
    m = np.full(max(len(kp1), 1), -1, np.int32)
    nm = L.orc_search_initialization(_p(kp1), _p(d1), len(kp1), _p(pm), _p(kp2), _p(d2), len(kp2), _p(b4), int(window_size), nnratio,
                                     1 if check_ori else 0, _p(m))
    return m[:len(kp1)], nm


def pose_optimization(pose7, Xw, obs, inv_sigma2, cam5):
    """Optimizer::PoseOptimization restated (Optimizer.cc:55-412).  Returns dict(pose, outlier, inliers, rounds, iterations, trials)."""
    L = lib()
    L.orc_pose_optimization.restype = C.c_int
    L.orc_pose_optimization.argtypes = [C.c_int] + [C.c_void_p] * 7
    c = np.ascontiguousarray
    pose = c(pose7, np.float64).copy()
    Xw, obs, w, cam5 = c(Xw, np.float64), c(obs, np.float64), c(inv_sigma2, np.float64), c(cam5, np.float64)
    n = len(Xw)
    out = np.zeros(max(n, 1), np.uint8)
    stats = np.zeros(4, np.float64)
    inl = L.orc_pose_optimization(n, _p(Xw), _p(obs), _p(w), _p(cam5), _p(pose), _p(out), _p(stats))
    return dict(pose=pose, outlier=out[:n], inliers=inl, rounds=int(stats[0]), iterations=int(stats[1]), trials=int(stats[2]),
                lambda_=float(stats[3]))


def set_stereo_form(device_form):
    """0 (default): the stereo projections of the optimisers exactly as the reference's edges compute them (pinned bitwise against the
    reference source, oracle/_ref part 7).  1: as the device kernels compute them (float(z) before a float division, double bf * invz,
    the binary edge's quotient Jacobian for the pose-only edge) -- see oracle/lba_oracle.cpp g_stereo_form.  Returns the previous form."""
    L = lib()
    L.orc_set_stereo_form.restype = C.c_int
    return int(L.orc_set_stereo_form(int(device_form)))


def pose_edge(pose7, Xw, obs3, cam5):
    """One pose-only edge (PoseOptimization): (D, B[D][6], r[D])."""
    L = lib()
    L.orc_pose_edge.restype = C.c_int
    c = lambda a: np.ascontiguousarray(a, np.float64)
    pose7, Xw, obs3, cam5 = c(pose7), c(Xw), c(obs3), c(cam5)
    B, r = np.zeros(18), np.zeros(3)
    D = L.orc_pose_edge(_p(pose7), _p(Xw), _p(obs3), _p(cam5), _p(B), _p(r))
    return D, B.reshape(3, 6)[:D], r[:D]


def is_in_frustum(Rcw, tcw, Ow, bounds, cam6, n_levels, log_scale_factor, xw, normal, max_dist, min_dist, viewing_cos_limit=0.5):
    """Frame::isInFrustum restated (Frame.cc:667-720).  Returns dict(in_view, proj_x, proj_y, proj_xr, level, view_cos, depth)."""
    L = lib()
    L.orc_is_in_frustum.restype = C.c_int
    L.orc_is_in_frustum.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_float, C.c_float] + [C.c_void_p] * 11
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    R, t, O, b4, c6 = f32(Rcw), f32(tcw), f32(Ow), f32(bounds), f32(cam6)
    xw, nr, mx, mn = f32(xw), f32(normal), f32(max_dist), f32(min_dist)
    n = len(xw)
    out = dict(in_view=np.zeros(max(n, 1), np.uint8), proj_x=np.zeros(max(n, 1), np.float32), proj_y=np.zeros(max(n, 1), np.float32),
               proj_xr=np.zeros(max(n, 1), np.float32), level=np.zeros(max(n, 1), np.int32), view_cos=np.zeros(max(n, 1), np.float32),
               depth=np.zeros(max(n, 1), np.float32))
    L.orc_is_in_frustum(n, _p(R), _p(t), _p(O), _p(b4), _p(c6), int(n_levels), float(log_scale_factor), float(viewing_cos_limit), _p(xw), _p(nr),
                        _p(mx), _p(mn), *[_p(out[k]) for k in ("in_view", "proj_x", "proj_y", "proj_xr", "level", "view_cos", "depth")])
    return {k: v[:n] for k, v in out.items()}


def bow_transform(voc, desc, levelsup=4):
    """DBoW2 transform restated (TemplatedVocabulary.h:1127-1260).  voc: dict(child_offset, child_ids, node_desc, node_word, node_weight, L).
    Returns dict(word, node, weight, bow_word, bow_weight)."""
    L = lib()
    L.orc_bow_transform.restype = C.c_int
    L.orc_bow_transform.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_int] + [C.c_void_p] * 6
    c = np.ascontiguousarray
    co, ci = c(voc["child_offset"], np.int32), c(voc["child_ids"], np.int32)
    nd, nw, wt = c(voc["node_desc"], np.uint8), c(voc["node_word"], np.int32), c(voc["node_weight"], np.float64)
    desc = c(desc, np.uint8)
    n = len(desc)
    word, node, weight = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float64)
    bw, bv = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float64)
    k = L.orc_bow_transform(len(nw), int(voc["L"]), _p(co), _p(ci), _p(nd), _p(nw), _p(wt), int(levelsup), n, _p(desc), _p(word), _p(node), _p(weight),
                            _p(bw), _p(bv))
    return dict(word=word[:n], node=node[:n], weight=weight[:n], bow_word=bw[:k], bow_weight=bv[:k])


def distinctive_descriptor(descs):
    """MapPoint::ComputeDistinctiveDescriptors restated (MapPoint.cc:438-520): index of the chosen observation descriptor."""
    L = lib()
    L.orc_distinctive_descriptor.restype = C.c_int
    L.orc_distinctive_descriptor.argtypes = [C.c_int, C.c_void_p]
    d = np.ascontiguousarray(descs, np.uint8).reshape(-1, 32)
    return L.orc_distinctive_descriptor(len(d), _p(d) if len(d) else None)


def update_normal_and_depth(centers, pos, ref_center, level, scale_factors):
    """MapPoint::UpdateNormalAndDepth restated (MapPoint.cc:567-640).  Returns (normal[3], max_dist, min_dist) as float32."""
    L = lib()
    L.orc_update_normal_and_depth.restype = None
    L.orc_update_normal_and_depth.argtypes = [C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 3
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    c, p, r, sf = f32(centers).reshape(-1, 3), f32(pos), f32(ref_center), f32(scale_factors)
    out, mx, mn = np.zeros(3, np.float32), np.zeros(1, np.float32), np.zeros(1, np.float32)
    L.orc_update_normal_and_depth(len(c), _p(c), _p(p), _p(r), int(level), _p(sf), len(sf), _p(out), _p(mx), _p(mn))
    return out, mx[0], mn[0]


def hamming_knn2(query, train):
    """cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, k=2) restated (Frame.cc:1553).  Returns (idx[nq][2], dist[nq][2]), -1 = none."""
    L = lib()
    L.orc_hamming_knn2.restype = None
    L.orc_hamming_knn2.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    q, t = np.ascontiguousarray(query, np.uint8).reshape(-1, 32), np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
    idx, dist = np.zeros((max(len(q), 1), 2), np.int32), np.zeros((max(len(q), 1), 2), np.int32)
    L.orc_hamming_knn2(len(q), _p(q) if len(q) else None, len(t), _p(t) if len(t) else None, _p(idx), _p(dist))
    return idx[:len(q)], dist[:len(q)]


# ---- Optimizer::LocalInertialBA ------------------------------------------------------------
LIBA_LINK = np.dtype([("k1", np.int32), ("k2", np.int32), ("robust", np.int32), ("_pad", np.int32), ("dt", np.float64),
                      ("dR", np.float32, 9), ("dV", np.float32, 3), ("dP", np.float32, 3), ("JRg", np.float32, 9), ("JVg", np.float32, 9),
                      ("JVa", np.float32, 9), ("JPg", np.float32, 9), ("JPa", np.float32, 9), ("blin", np.float32, 6),
                      ("info", np.float64, 81), ("infoG", np.float64, 9), ("infoA", np.float64, 9)], align=False)


def _liba_lib():
    L = lib()
    L.orc_inertial_link_size.restype = C.c_int
    assert L.orc_inertial_link_size() == LIBA_LINK.itemsize, (L.orc_inertial_link_size(), LIBA_LINK.itemsize)
    return L


def liba(state, fixed, point, edge_kf, edge_mp, obs, inv_sigma2, Tcb12, cam5, links, lambda_init=1.0, max_iters=10):
    """Optimizer::LocalInertialBA's g2o core restated (Optimizer.cc:2203-2812).  state: [nKF][21] = Rwb(9) twb v bg ba.
    Returns dict(state, point, edge_chi2, link_chi2[nL][3], iterations, trials, lambda_, chi2, chi2_init, chi2_last, edge_depth_pos)."""
    L = _liba_lib()
    L.orc_liba.restype = C.c_int
    L.orc_liba.argtypes = [C.c_int] * 4 + [C.c_void_p] * 10 + [C.c_double, C.c_int] + [C.c_void_p] * 4
    c = np.ascontiguousarray
    st, pt = c(state, np.float64).copy(), c(point, np.float64).copy()
    fx = c(fixed, np.uint8)
    ekf, emp = c(edge_kf, np.int32), c(edge_mp, np.int32)
    ob, w = c(obs, np.float64), c(inv_sigma2, np.float64)
    T, cam = c(Tcb12, np.float64), c(cam5, np.float64)
    lk = c(links, LIBA_LINK)
    nE, nL = len(ekf), len(lk)
    chi, lchi, stats, dpos = np.zeros(max(nE, 1)), np.zeros(max(3 * nL, 1)), np.zeros(8), np.zeros(max(nE, 1), np.uint8)
    P = lambda a: _p(a) if a.size else None
    it = L.orc_liba(len(st), len(pt), nE, nL, _p(st), _p(fx), P(pt), P(ekf), P(emp), P(ob), P(w), _p(T), _p(cam), P(lk), float(lambda_init),
                    int(max_iters), _p(chi), _p(lchi), _p(stats), _p(dpos))
    return dict(state=st, point=pt, edge_chi2=chi[:nE], link_chi2=lchi[:3 * nL].reshape(nL, 3), iterations=it, trials=int(stats[3]),
                lambda_=stats[1], chi2=stats[2], chi2_init=stats[4], chi2_last=stats[5], edge_depth_pos=dpos[:nE])


def inertial_edge(state2, link):
    """EdgeInertial::computeError / linearizeOplus at two keyframe states: (e9, J[9][24]) with columns pose1 v1 bg1 ba1 pose2 v2."""
    L = _liba_lib()
    L.orc_inertial_edge.restype = None
    L.orc_inertial_edge.argtypes = [C.c_void_p] * 4
    st = np.ascontiguousarray(state2, np.float64)
    lk = np.ascontiguousarray(link, LIBA_LINK).reshape(1)
    e, J = np.zeros(9), np.zeros(270)
    L.orc_inertial_edge(_p(st), _p(lk), _p(e), _p(J))
    return e, J.reshape(9, 30)[:, :24]


def kf_oplus(state21, d15):
    L = lib()
    L.orc_kf_oplus.restype = None
    L.orc_kf_oplus.argtypes = [C.c_void_p, C.c_void_p]
    s, d = np.ascontiguousarray(state21, np.float64).copy(), np.ascontiguousarray(d15, np.float64)
    L.orc_kf_oplus(_p(s), _p(d))
    return s


def liba_reproj(state21, Xw, obs3, Tcb12, cam5):
    """EdgeMono / EdgeStereo of the inertial BA at one keyframe: (D, r[3], Jpoint[D][3], Jpose[D][6])."""
    L = lib()
    L.orc_liba_reproj.restype = C.c_int
    L.orc_liba_reproj.argtypes = [C.c_void_p] * 8
    c = lambda a: np.ascontiguousarray(a, np.float64)
    s, X, o, T, cam = c(state21), c(Xw), c(obs3), c(Tcb12), c(cam5)
    r, Jp, Jx = np.zeros(3), np.zeros(9), np.zeros(18)
    D = L.orc_liba_reproj(_p(s), _p(X), _p(o), _p(T), _p(cam), _p(r), _p(Jp), _p(Jx))
    return D, r, Jp.reshape(3, 3)[:D], Jx.reshape(3, 6)[:D]


# ----------------------------------------------------------------------------------------------------------------
# oracle/_ref: the reference's OWN src/ORBextractor.cc, compiled unmodified against ref_shim/opencv2 (oracle/Makefile,
# target `ref`).  Buildable only where /root/reference exists (the build container); the built .so travels.
# ----------------------------------------------------------------------------------------------------------------
_REF_SO = os.path.join(_HERE, "_ref", "liborb_ref.so")
REFERENCE_ROOT = os.environ.get("ORB_REFERENCE_ROOT", "/root/reference")
_ref_lib = None


def build_ref(force=False):
    """Returns the path of oracle/_ref/liborb_ref.so, building it when the reference checkout is present; None when it is
    neither built nor buildable."""
    if os.path.exists(os.path.join(REFERENCE_ROOT, "src", "ORBextractor.cc")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref", f"REF={REFERENCE_ROOT}"] + (["-B"] if force else []))
    return _REF_SO if os.path.exists(_REF_SO) else None


def ref_lib():
    global _ref_lib
    if _ref_lib is None:
        if build_ref() is None:
            raise RuntimeError("oracle/_ref/liborb_ref.so is not built and /root/reference is not present")
        L = _ref_lib = C.CDLL(_REF_SO)
        L.ref_extractor_create.restype = C.c_void_p
        L.ref_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.ref_extractor_destroy.argtypes = [C.c_void_p]
        L.ref_extract.restype = C.c_int
        L.ref_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_int, C.POINTER(C.c_int)]
        L.ref_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.ref_level_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_level_pyramid.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_keypoints_octtree.restype = C.c_int
        L.ref_keypoints_octtree.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_distribute.restype = C.c_int
        L.ref_distribute.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_int]
    return _ref_lib


class RefExtractor:
    """ORB_SLAM3::ORBextractor itself (reference src/ORBextractor.cc compiled as it lies) behind a C wrapper."""

    def __init__(self, nfeatures=1200, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = ref_lib()
        self.nlevels, self.nfeatures = nlevels, nfeatures
        self.h = C.c_void_p(self.L.ref_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th))
        sc, isc, s2, is2 = (np.zeros(nlevels, np.float32) for _ in range(4))
        q = np.zeros(nlevels, np.int32)
        um = np.zeros(16, np.int32)
        self.L.ref_tables(self.h, _p(sc), _p(isc), _p(s2), _p(is2), _p(q), _p(um))
        self.scale_factors, self.inv_scale_factors, self.level_sigma2, self.inv_level_sigma2 = sc, isc, s2, is2
        self.features_per_level, self.umax = q, um

    def __del__(self):
        try:
            self.L.ref_extractor_destroy(self.h)
        except Exception:
            pass

    def __call__(self, img, lapping=(0, 0)):
        img = np.ascontiguousarray(img, np.uint8)
        H, W = img.shape
        cap = max(4 * self.nfeatures, 4096)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        mono = self.L.ref_extract(self.h, _p(img), W, H, W, int(lapping[0]), int(lapping[1]), _p(kps), _p(desc), cap, C.byref(n))
        assert mono != -100 and n.value <= cap
        return mono, kps[:n.value].copy(), desc[:n.value].copy()

    def level_size(self, l):
        w, h = C.c_int(), C.c_int()
        self.L.ref_level_size(self.h, l, C.byref(w), C.byref(h))
        return w.value, h.value

    def level_pyramid(self, l):
        w, h = self.level_size(l)
        a = np.zeros((h, w), np.uint8)
        self.L.ref_level_pyramid(self.h, l, _p(a))
        return a

    def keypoints_octtree(self, img):
        """ComputePyramid + ComputeKeyPointsOctTree: list of per-level keypoint arrays (level coordinates, angles set)."""
        img = np.ascontiguousarray(img, np.uint8)
        H, W = img.shape
        cap = max(8 * self.nfeatures, 8192)
        kps = np.zeros(cap, KP_DTYPE)
        off = np.zeros(self.nlevels + 1, np.int32)
        n = self.L.ref_keypoints_octtree(self.h, _p(img), W, H, W, _p(kps), cap, _p(off))
        assert n >= 0
        return [kps[off[l]:off[l + 1]].copy() for l in range(self.nlevels)]

    def distribute(self, cand, minX, maxX, minY, maxY, N, level=0):
        """DistributeOctTree on an (n, 3) int32 candidate array (x, y, response): (k, 4) rows (x, y, response, input index)."""
        cand = np.ascontiguousarray(cand, np.int32)
        cap = len(cand) + 8
        out = np.zeros((cap, 4), np.int32)
        k = self.L.ref_distribute(self.h, _p(cand), len(cand), minX, maxX, minY, maxY, N, level, _p(out), cap)
        assert k >= 0
        return out[:k].copy()


def oracle_distribute(cand, minX, maxX, minY, maxY, N):
    """The restatement's DistributeOctTree on the same input: indices into cand in output order."""
    cand = np.ascontiguousarray(cand, np.int32)
    out = np.zeros(len(cand) + 8, np.int32)
    k = lib().orc_distribute(_p(cand), len(cand), minX, maxX, minY, maxY, N, _p(out), len(out))
    return out[:k].copy()


# ---- oracle/_ref part 2: member functions of the reference's ORBmatcher.cc / Frame.cc / MapPoint.cc / Pinhole.cpp compiled verbatim
# over skeleton classes (oracle/Makefile target ref2; ref_shim/ref_capi2.cpp) ------------------------------------------------------
_REF2_SO = os.path.join(_HERE, "_ref", "liborb_ref2.so")
_ref2_lib = None


def build_ref2(force=False):
    if os.path.exists(os.path.join(REFERENCE_ROOT, "src", "ORBmatcher.cc")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref2", f"REF={REFERENCE_ROOT}"] + (["-B"] if force else []))
    return _REF2_SO if os.path.exists(_REF2_SO) else None


def ref2_lib():
    global _ref2_lib
    if _ref2_lib is None:
        if build_ref2() is None:
            raise RuntimeError("oracle/_ref/liborb_ref2.so is not built and /root/reference is not present")
        ref_lib()
        L = _ref2_lib = C.CDLL(_REF2_SO)
        VP, I, F = C.c_void_p, C.c_int, C.c_float
        L.ref2_frame_create.restype = VP
        L.ref2_frame_create.argtypes = [VP, VP, VP, I, VP, VP, I, VP, VP]
        L.ref2_frame_destroy.argtypes = [VP]
        L.ref2_descriptor_distance.restype = I
        L.ref2_descriptor_distance.argtypes = [VP, VP]
        L.ref2_three_maxima.argtypes = [VP, I, VP]
        L.ref2_constants.argtypes = [VP, VP]
        L.ref2_get_features_in_area.restype = I
        L.ref2_get_features_in_area.argtypes = [VP, F, F, F, I, I, VP, I]
        L.ref2_search_local.restype = I
        L.ref2_search_local.argtypes = [VP, I] + [VP] * 8 + [F, F, I, F, VP]
        L.ref2_search_last.restype = I
        L.ref2_search_last.argtypes = [VP, VP, VP, I] + [VP] * 5 + [F, I, I, VP, VP]
        L.ref2_is_in_frustum.restype = I
        L.ref2_is_in_frustum.argtypes = [VP, VP, VP, VP, I, VP, VP, VP, VP, F] + [VP] * 7
        L.ref2_stereo_matches.argtypes = [VP, VP, VP, VP, VP, I, VP, VP]
        L.ref2_distinctive_descriptor.restype = I
        L.ref2_distinctive_descriptor.argtypes = [I, VP, VP, VP]
        L.ref2_update_normal_and_depth.argtypes = [I, VP, VP, I, I, VP, I, VP, VP]
        L.ref2_kf_set_geometry.argtypes = [VP, VP, VP, F]
        L.ref2_kf_set_mappoints.argtypes = [VP, VP, VP, VP, VP]
        L.ref2_fuse.restype = I
        L.ref2_fuse.argtypes = [VP, I] + [VP] * 7 + [F, VP]
        L.ref2_fuse_sim3.restype = I
        L.ref2_fuse_sim3.argtypes = [VP, VP, I] + [VP] * 6 + [F, VP, VP, VP]
        L.ref2_search_kf_sim3.restype = I
        L.ref2_search_kf_sim3.argtypes = [VP, VP, I] + [VP] * 7 + [F, F, I, VP, VP, VP]
        L.ref2_search_frame_kf.restype = I
        L.ref2_search_frame_kf.argtypes = [VP, VP, VP, VP, VP, F, I, I, VP]
        L.ref2_search_by_sim3.restype = I
        L.ref2_search_by_sim3.argtypes = [VP, VP, VP, VP, F, VP]
        L.ref2_kf_create.restype = VP
        L.ref2_kf_create.argtypes = [VP, VP, VP, I, VP, VP, VP, VP, VP, I, VP, VP]
        L.ref2_kf_destroy.argtypes = [VP]
        L.ref2_search_bow.restype = I
        L.ref2_search_bow.argtypes = [VP, VP, VP, F, I, VP]
        L.ref2_search_bow_kf.restype = I
        L.ref2_search_bow_kf.argtypes = [VP, VP, F, I, VP]
        L.ref2_search_initialization.restype = I
        L.ref2_search_initialization.argtypes = [VP, VP, VP, I, F, I, VP]
        L.ref2_search_triangulation.restype = I
        L.ref2_search_triangulation.argtypes = [VP, VP, I, I, I, VP, VP, VP]
    return _ref2_lib


class RefFrame:
    """A reference Frame (Nleft == -1) skeleton filled from flat arrays; methods run the reference's own member functions."""

    def __init__(self, kps, desc, uright, bounds, scale_factors, cam6, Tcw7=None):
        self.L = ref2_lib()
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        self._keep = (np.ascontiguousarray(kps), np.ascontiguousarray(desc, np.uint8), f32(uright), f32(bounds), f32(scale_factors), f32(cam6), f32(Tcw7))
        k, d, u, b, s, c, t = self._keep
        P = lambda a: None if a is None else _p(a)
        self.N = len(k)
        self.h = C.c_void_p(self.L.ref2_frame_create(_p(k), _p(d), P(u), len(k), _p(b), _p(s), len(s), _p(c), P(t)))

    def __del__(self):
        try:
            self.L.ref2_frame_destroy(self.h)
        except Exception:
            pass

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        out = np.zeros(max(self.N, 1), np.int32)
        n = self.L.ref2_get_features_in_area(self.h, x, y, r, min_level, max_level, _p(out), len(out))
        return out[:n].copy()

    def search_local(self, projx, projy, projxr, level, viewcos, qdesc, th, nnratio, claimed=None, trackdepth=None, far=False, th_far=0.0):
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        px, py, pxr, vc, td = f32(projx), f32(projy), f32(projxr), f32(viewcos), f32(trackdepth)
        lv, qd = np.ascontiguousarray(level, np.int32), np.ascontiguousarray(qdesc, np.uint8)
        cl = None if claimed is None else np.ascontiguousarray(claimed, np.uint8)
        nq = len(px)
        match = np.full(max(nq, 1), -1, np.int32)
        P = lambda a: None if a is None else _p(a)
        nm = self.L.ref2_search_local(self.h, nq, _p(px), _p(py), _p(pxr), _p(lv), _p(vc), P(td), _p(qd), P(cl), th, nnratio, 1 if far else 0, th_far, _p(match))
        return match[:nq], nm

    def search_last(self, Tcw7, Tlw7, xw, last_octave, last_angle, qdesc, obs_pos, th, check_ori=True, mono=False):
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        Tc, Tl, xw, ang = f32(Tcw7), f32(Tlw7), f32(xw), f32(last_angle)
        lo, qd, ob = np.ascontiguousarray(last_octave, np.int32), np.ascontiguousarray(qdesc, np.uint8), np.ascontiguousarray(obs_pos, np.uint8)
        fm = np.full(max(self.N, 1), -1, np.int32)
        direction = C.c_int(0)
        nm = self.L.ref2_search_last(self.h, _p(Tc), _p(Tl), len(lo), _p(xw), _p(lo), _p(ang), _p(qd), _p(ob), th, 1 if check_ori else 0, 1 if mono else 0,
                                     _p(fm), C.byref(direction))
        return fm[:self.N], nm, direction.value

    def is_in_frustum(self, Rcw, tcw, Ow, xw, normal, max_dist, min_dist, limit=0.5):
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        R, t, O, xw, nr, mx, mn = f32(Rcw), f32(tcw), f32(Ow), f32(xw), f32(normal), f32(max_dist), f32(min_dist)
        n = len(xw)
        out = dict(in_view=np.zeros(max(n, 1), np.uint8), proj_x=np.zeros(max(n, 1), np.float32), proj_y=np.zeros(max(n, 1), np.float32),
                   proj_xr=np.zeros(max(n, 1), np.float32), level=np.zeros(max(n, 1), np.int32), view_cos=np.zeros(max(n, 1), np.float32),
                   depth=np.zeros(max(n, 1), np.float32))
        self.L.ref2_is_in_frustum(self.h, _p(R), _p(t), _p(O), n, _p(xw), _p(nr), _p(mx), _p(mn), limit,
                                  *[_p(out[k]) for k in ("in_view", "proj_x", "proj_y", "proj_xr", "level", "view_cos", "depth")])
        return {k: v[:n] for k, v in out.items()}

    def stereo_matches(self, ref_ex_left, ref_ex_right, kpsR, descR):
        """Frame::ComputeStereoMatches; the two RefExtractor objects must just have extracted the left / right image."""
        kR, dR = np.ascontiguousarray(kpsR), np.ascontiguousarray(descR, np.uint8)
        uR, dep = np.zeros(max(self.N, 1), np.float32), np.zeros(max(self.N, 1), np.float32)
        self.L.ref2_stereo_matches(self.h, ref_ex_left.h, ref_ex_right.h, _p(kR), _p(dR), len(kR), _p(uR), _p(dep))
        return uR[:self.N], dep[:self.N]


def ref2_descriptor_distance(a, b):
    a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
    return ref2_lib().ref2_descriptor_distance(_p(a), _p(b))


def ref2_three_maxima(sizes):
    s = np.ascontiguousarray(sizes, np.int32)
    out = np.zeros(3, np.int32)
    ref2_lib().ref2_three_maxima(_p(s), len(s), _p(out))
    return tuple(int(v) for v in out)


def ref2_constants():
    c, r = np.zeros(3, np.int32), np.zeros(2, np.float32)
    ref2_lib().ref2_constants(_p(c), _p(r))
    return dict(TH_LOW=int(c[0]), TH_HIGH=int(c[1]), HISTO_LENGTH=int(c[2]), radius_close=float(r[0]), radius_far=float(r[1]))


class RefKeyFrame:
    """A reference KeyFrame skeleton (single camera) filled from flat arrays, for the KeyFrame-typed matchers of ORBmatcher.cc."""

    def __init__(self, kps, desc, uright=None, node=None, has_mp=None, bad=None, scale_factors=None, level_sigma2=None, cam4=(1, 1, 0, 0), Tcw7=None):
        self.L = ref2_lib()
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        u8 = lambda a: None if a is None else np.ascontiguousarray(a, np.uint8)
        i32 = lambda a: None if a is None else np.ascontiguousarray(a, np.int32)
        self._keep = (np.ascontiguousarray(kps), u8(desc), f32(uright), i32(node), u8(has_mp), u8(bad), f32(scale_factors), f32(level_sigma2), f32(cam4), f32(Tcw7))
        k, d, u, n, m, b, s, l2, c, t = self._keep
        P = lambda a: None if a is None else _p(a)
        self.N = len(k)
        self.h = C.c_void_p(self.L.ref2_kf_create(_p(k), _p(d), P(u), len(k), P(n), P(m), P(b), P(s), P(l2), 0 if s is None else len(s), _p(c), P(t)))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref2_kf_destroy(self.h)
            self.h = None


def _qarrays(q):
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    u8 = lambda a: None if a is None else np.ascontiguousarray(a, np.uint8)
    return f32(q["world_pos"]), f32(q.get("normal")), f32(q["max_dist"]), f32(q["min_dist"]), u8(q["desc"]), u8(q.get("bad"))


def ref2_kf_set_geometry(kf, frame, inv_level_sigma2, bf):
    a = np.ascontiguousarray(inv_level_sigma2, np.float32)
    kf.L.ref2_kf_set_geometry(kf.h, frame.h, _p(a), float(bf))
    kf._geom_frame = frame


def ref2_kf_set_mappoints(kf, world_pos, max_dist, min_dist, desc):
    a = [np.ascontiguousarray(world_pos, np.float32), np.ascontiguousarray(max_dist, np.float32), np.ascontiguousarray(min_dist, np.float32),
         np.ascontiguousarray(desc, np.uint8)]
    kf.L.ref2_kf_set_mappoints(kf.h, *[_p(x) for x in a])


def ref2_fuse(kf, q, th, in_kf=None):
    """ORBmatcher::Fuse(pKF, vpMapPoints, th) of the reference: (fused[nq] = keyframe feature or -1, nFused)."""
    P = lambda a: None if a is None else _p(a)
    xw, nr, mx, mn, d, bad = _qarrays(q)
    ik = None if in_kf is None else np.ascontiguousarray(in_kf, np.uint8)
    out = np.full(max(len(xw), 1), -1, np.int32)
    n = kf.L.ref2_fuse(kf.h, len(xw), _p(xw), P(nr), _p(mx), _p(mn), _p(d), P(bad), P(ik), float(th), _p(out))
    return out[:len(xw)], n


def ref2_fuse_sim3(kf, S8, q, th):
    P = lambda a: None if a is None else _p(a)
    xw, nr, mx, mn, d, bad = _qarrays(q)
    s = np.ascontiguousarray(S8, np.float32)
    out, T, Ow = np.full(max(len(xw), 1), -1, np.int32), np.zeros(7, np.float32), np.zeros(3, np.float32)
    n = kf.L.ref2_fuse_sim3(kf.h, _p(s), len(xw), _p(xw), P(nr), _p(mx), _p(mn), _p(d), P(bad), float(th), _p(out), _p(T), _p(Ow))
    return out[:len(xw)], n, T, Ow


def ref2_search_kf_sim3(kf, S8, q, matched_in, th, ratio_hamming, with_kfs=False):
    P = lambda a: None if a is None else _p(a)
    xw, nr, mx, mn, d, bad = _qarrays(q)
    s = np.ascontiguousarray(S8, np.float32)
    mi = None if matched_in is None else np.ascontiguousarray(matched_in, np.uint8)
    out, T, Ow = np.full(max(len(xw), 1), -1, np.int32), np.zeros(7, np.float32), np.zeros(3, np.float32)
    n = kf.L.ref2_search_kf_sim3(kf.h, _p(s), len(xw), _p(xw), P(nr), _p(mx), _p(mn), _p(d), P(bad), P(mi), float(th), float(ratio_hamming),
                                 1 if with_kfs else 0, _p(out), _p(T), _p(Ow))
    return out[:len(xw)], n, T, Ow


def ref2_search_frame_kf(frame, Tcw7, kf, already, claimed, th, orb_dist, check_ori=True):
    P = lambda a: None if a is None else _p(a)
    T = np.ascontiguousarray(Tcw7, np.float32)
    al = None if already is None else np.ascontiguousarray(already, np.uint8)
    cl = None if claimed is None else np.ascontiguousarray(claimed, np.uint8)
    out = np.full(max(frame.N, 1), -1, np.int32)
    n = frame.L.ref2_search_frame_kf(frame.h, _p(T), kf.h, P(al), P(cl), float(th), int(orb_dist), 1 if check_ori else 0, _p(out))
    return out[:frame.N], n


def ref2_search_by_sim3(kf1, kf2, S12, S21, th, matches12=None):
    a, b = np.ascontiguousarray(S12, np.float32), np.ascontiguousarray(S21, np.float32)
    m = np.full(max(kf1.N, 1), -1, np.int32) if matches12 is None else np.ascontiguousarray(matches12, np.int32).copy()
    n = kf1.L.ref2_search_by_sim3(kf1.h, kf2.h, _p(a), _p(b), float(th), _p(m))
    return m[:kf1.N], n


def ref2_search_bow(frame, feat_node, kf, nnratio, check_ori=True):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) of the reference: (feat_match[N] = KF feature index or -1, nmatches)."""
    fn = np.ascontiguousarray(feat_node, np.int32)
    fm = np.full(max(frame.N, 1), -1, np.int32)
    n = frame.L.ref2_search_bow(frame.h, _p(fn), kf.h, nnratio, 1 if check_ori else 0, _p(fm))
    return fm[:frame.N], n


def ref2_search_bow_kf(kf1, kf2, nnratio, check_ori=True):
    m = np.full(max(kf1.N, 1), -1, np.int32)
    n = kf1.L.ref2_search_bow_kf(kf1.h, kf2.h, nnratio, 1 if check_ori else 0, _p(m))
    return m[:kf1.N], n


def ref2_search_initialization(f1, f2, prev_matched, window_size, nnratio, check_ori=True):
    pm = np.ascontiguousarray(prev_matched, np.float32).copy()
    m = np.full(max(f1.N, 1), -1, np.int32)
    n = f1.L.ref2_search_initialization(f1.h, f2.h, _p(pm), int(window_size), nnratio, 1 if check_ori else 0, _p(m))
    return m[:f1.N], n, pm


def ref2_search_triangulation(kf1, kf2, only_stereo=False, coarse=False, check_ori=True):
    m = np.full(max(kf1.N, 1), -1, np.int32)
    F12, ep = np.zeros(9, np.float32), np.zeros(2, np.float32)
    n = kf1.L.ref2_search_triangulation(kf1.h, kf2.h, 1 if only_stereo else 0, 1 if coarse else 0, 1 if check_ori else 0, _p(m), _p(F12), _p(ep))
    return m[:kf1.N], n, F12.reshape(3, 3), ep


def ref2_distinctive_descriptor(descs, kf_bad=None):
    """MapPoint::ComputeDistinctiveDescriptors of the reference over one observation per keyframe: the chosen 32 bytes, or None."""
    d = np.ascontiguousarray(descs, np.uint8).reshape(-1, 32)
    b = None if kf_bad is None else np.ascontiguousarray(kf_bad, np.uint8)
    out = np.zeros(32, np.uint8)
    ok = ref2_lib().ref2_distinctive_descriptor(len(d), _p(d) if len(d) else None, None if b is None else _p(b), _p(out))
    return out if ok else None


def ref2_update_normal_and_depth(centers, pos, ref, level, scale_factors):
    """MapPoint::UpdateNormalAndDepth of the reference: (normal[3], mfMaxDistance, mfMinDistance)."""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    c, p, sf = f32(centers).reshape(-1, 3), f32(pos), f32(scale_factors)
    n, mm = np.zeros(3, np.float32), np.zeros(2, np.float32)
    ref2_lib().ref2_update_normal_and_depth(len(c), _p(c), _p(p), int(ref), int(level), _p(sf), len(sf), _p(n), _p(mm))
    return n, mm[0], mm[1]


# ---- oracle/_ref part 3: the reference's own DBoW2 (Thirdparty/DBoW2) compiled where it lies (oracle/Makefile target ref3) -----------
_REF3_SO = os.path.join(_HERE, "_ref", "liborb_ref3.so")
_ref3_lib = None


def build_ref3(force=False):
    if os.path.exists(os.path.join(REFERENCE_ROOT, "Thirdparty", "DBoW2", "DBoW2", "TemplatedVocabulary.h")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref3", f"REF={REFERENCE_ROOT}"] + (["-B"] if force else []))
    return _REF3_SO if os.path.exists(_REF3_SO) else None


def ref3_lib():
    global _ref3_lib
    if _ref3_lib is None:
        if build_ref3() is None:
            raise RuntimeError("oracle/_ref/liborb_ref3.so is not built and /root/reference is not present")
        L = _ref3_lib = C.CDLL(_REF3_SO)
        VP, I = C.c_void_p, C.c_int
        L.ref3_voc_load_text.restype = VP
        L.ref3_voc_load_text.argtypes = [C.c_char_p]
        L.ref3_voc_destroy.argtypes = [VP]
        L.ref3_voc_info.argtypes = [VP, VP]
        L.ref3_transform.restype = I
        L.ref3_transform.argtypes = [VP, VP, I, I, VP, VP, VP, VP]
        L.ref3_transform_one.restype = I
        L.ref3_transform_one.argtypes = [VP, VP, I, VP, VP, VP]
        L.ref3_score.restype = C.c_double
        L.ref3_score.argtypes = [VP, I, VP, VP, I, VP, VP]
        L.ref3_forb_distance.restype = I
        L.ref3_forb_distance.argtypes = [VP, VP]
    return _ref3_lib


class RefVocabulary:
    """The reference's ORBVocabulary (DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>) loaded with its own loadFromTextFile."""

    def __init__(self, text_path):
        self.L = ref3_lib()
        self.h = C.c_void_p(self.L.ref3_voc_load_text(os.fsencode(text_path)))
        if not self.h:
            raise RuntimeError(f"loadFromTextFile({text_path}) failed")
        info = np.zeros(6, np.int32)
        self.L.ref3_voc_info(self.h, _p(info))
        self.k, self.depth, self.scoring, self.weighting, self.words, self.nodes = (int(x) for x in info)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref3_voc_destroy(self.h)
            self.h = None

    def transform(self, desc, levelsup=4):
        """Frame::ComputeBoW's call: returns dict(bow_word, bow_weight, feat_node, ascending)."""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        bw, bv, fn, asc = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float64), np.zeros(max(n, 1), np.int32), np.zeros(1, np.int32)
        k = self.L.ref3_transform(self.h, _p(d), n, int(levelsup), _p(bw), _p(bv), _p(fn), _p(asc))
        return dict(bow_word=bw[:k], bow_weight=bv[:k], feat_node=fn[:n], ascending=bool(asc[0]))

    def transform_one(self, d32, levelsup=4):
        d = np.ascontiguousarray(d32, np.uint8)
        w, n, wt = np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.float64)
        if not self.L.ref3_transform_one(self.h, _p(d), int(levelsup), _p(w), _p(wt), _p(n)):
            raise AssertionError("DBoW2: transform(feature) / getWordWeight disagree with transform(feature, id, weight, nid, levelsup)")
        return int(w[0]), float(wt[0]), int(n[0])

    def score(self, w1, v1, w2, v2):
        a, b = np.ascontiguousarray(w1, np.int32), np.ascontiguousarray(w2, np.int32)
        x, y = np.ascontiguousarray(v1, np.float64), np.ascontiguousarray(v2, np.float64)
        return float(self.L.ref3_score(self.h, len(a), _p(a), _p(x), len(b), _p(b), _p(y)))


def ref3_forb_distance(a, b):
    a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
    return int(ref3_lib().ref3_forb_distance(_p(a), _p(b)))


# ---- oracle/_ref part 4: g2o's own Levenberg control flow + Huber kernel over the oracle's LbaEngine (oracle/Makefile target ref4) ------
_REF4_SO = os.path.join(_HERE, "_ref", "liborb_ref4.so")
_ref4_lib = None


def build_ref4(force=False):
    if os.path.exists(os.path.join(REFERENCE_ROOT, "Thirdparty", "g2o", "g2o", "core", "optimization_algorithm_levenberg.cpp")):
        build()
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref4", f"REF={REFERENCE_ROOT}"] + (["-B"] if force else []))
    return _REF4_SO if os.path.exists(_REF4_SO) else None


def ref4_lib():
    global _ref4_lib
    if _ref4_lib is None:
        if build_ref4() is None:
            raise RuntimeError("oracle/_ref/liborb_ref4.so is not built and /root/reference is not present")
        lib()                                   # liborb_oracle.so first: ref4 resolves the LbaEngine entry points against it
        L = _ref4_lib = C.CDLL(_REF4_SO)
        L.ref4_lba.restype = C.c_int
        L.ref4_lba.argtypes = [C.c_int] * 3 + [C.c_void_p] * 8 + [C.c_double, C.c_int] + [C.c_void_p] * 3
        L.ref4_huber.argtypes = [C.c_double, C.c_float, C.c_void_p]
    return _ref4_lib


def ref4_lba(pose, fixed, point, edge_kf, edge_mp, obs, inv_sigma2, cam5, lambda_init=0.0, max_iters=10):
    """lba() with g2o's own OptimizationAlgorithmLevenberg::solve / SparseOptimizer::optimize deciding every step."""
    L = ref4_lib()
    pose = np.ascontiguousarray(pose, np.float64).copy()
    point = np.ascontiguousarray(point, np.float64).copy()
    fixed = np.ascontiguousarray(fixed, np.uint8)
    ekf, emp = np.ascontiguousarray(edge_kf, np.int32), np.ascontiguousarray(edge_mp, np.int32)
    obs, w = np.ascontiguousarray(obs, np.float64), np.ascontiguousarray(inv_sigma2, np.float64)
    cam5 = np.ascontiguousarray(cam5, np.float64)
    nE = len(ekf)
    chi, dpos, stats = np.zeros(nE, np.float64), np.zeros(nE, np.uint8), np.zeros(8, np.float64)
    it = L.ref4_lba(len(pose), len(point), nE, _p(pose), _p(fixed), _p(point), _p(ekf), _p(emp), _p(obs), _p(w), _p(cam5),
                    float(lambda_init), int(max_iters), _p(chi), _p(dpos), _p(stats))
    return dict(pose=pose, point=point, edge_chi2=chi, edge_depth_pos=dpos, iterations=it, lambda_=stats[1], chi2=stats[2], trials=int(stats[3]))


def ref4_huber(e, delta):
    out = np.zeros(3, np.float64)
    ref4_lib().ref4_huber(float(e), float(np.float32(delta)), _p(out))
    return out


def huber(e, delta):
    """The oracle's Huber weight as every optimiser of oracle/lba_oracle.cpp applies it: (rho, rho')."""
    L = lib()
    L.orc_huber.argtypes = [C.c_double, C.c_float, C.c_void_p]
    out = np.zeros(2, np.float64)
    L.orc_huber(float(e), float(np.float32(delta)), _p(out))
    return out


# ---- oracle/_ref part 5: the reference's own Optimizer::PoseOptimization over the oracle's PoseEngine (oracle/Makefile target ref5) ----
_REF5_SO = os.path.join(_HERE, "_ref", "liborb_ref5.so")
_ref5_lib = None


_REF7_SO = os.path.join(_HERE, "_ref", "liborb_ref7.so")
_ref7_lib = None


def build_ref7(force=False):
    """oracle/_ref part 7: the reprojection edges' own linearizeOplus / cam_project and Pinhole::projectJac / project, compiled verbatim."""
    if os.path.exists(os.path.join(REFERENCE_ROOT, "src", "OptimizableTypes.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref7", f"REF={REFERENCE_ROOT}"] + (["-B"] if force else []))
    return _REF7_SO if os.path.exists(_REF7_SO) else None


def ref7_edge(pose7, X, obs3, cam5, unary):
    """The reference's own edge of LocalBundleAdjustment (unary False: EdgeSE3ProjectXYZ / EdgeStereoSE3ProjectXYZ) or PoseOptimization
    (unary True: the ...OnlyPose edges); obs3[2] < 0 selects the monocular edge.  cam5 as float32, like the KeyFrame / Frame members the
    Optimizer copies into the edge.  Returns (A[D][3] or None, B[D][6], r[D])."""
    global _ref7_lib
    if _ref7_lib is None:
        if build_ref7() is None:
            raise RuntimeError("oracle/_ref/liborb_ref7.so is not built and /root/reference is not present")
        _ref7_lib = C.CDLL(_REF7_SO)
    c = lambda a: np.ascontiguousarray(a, np.float64)
    pose7, X, obs3 = c(pose7), c(X), c(obs3)
    cam = np.ascontiguousarray(cam5, np.float32)
    stereo = int(obs3[2] >= 0)
    D = 3 if stereo else 2
    A, B, r = np.zeros(9), np.zeros(18), np.zeros(3)
    if unary:
        _ref7_lib.ref7_pose_edge(_p(pose7), _p(X), _p(obs3), _p(cam), stereo, _p(B), _p(r))
        return None, B[:6 * D].reshape(D, 6), r[:D]
    _ref7_lib.ref7_edge(_p(pose7), _p(X), _p(obs3), _p(cam), stereo, _p(A), _p(B), _p(r))
    return A[:3 * D].reshape(D, 3), B[:6 * D].reshape(D, 6), r[:D]


def edge(pose7, X, obs3, cam5):
    """The oracle's binary edge (LocalBundleAdjustment): (A[D][3], B[D][6], r[D])."""
    L = lib()
    L.orc_edge_residual.restype = C.c_int
    c = lambda a: np.ascontiguousarray(a, np.float64)
    pose7, X, obs3, cam5 = c(pose7), c(X), c(obs3), c(cam5)
    A, B, r = np.zeros(9), np.zeros(18), np.zeros(3)
    D = L.orc_edge_residual(_p(pose7), _p(X), _p(obs3), _p(cam5), _p(r))
    L.orc_edge_jacobians(_p(pose7), _p(X), D, _p(cam5), _p(A), _p(B))
    return A.reshape(3, 3)[:D], B.reshape(3, 6)[:D], r[:D]


def build_ref5(force=False):
    if os.path.exists(os.path.join(REFERENCE_ROOT, "src", "Optimizer.cc")):
        build()
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref5", f"REF={REFERENCE_ROOT}"] + (["-B"] if force else []))
    return _REF5_SO if os.path.exists(_REF5_SO) else None


def ref5_pose_optimization(pose7, has_mp, world_pos, kp_xy, octave, uright, inv_level_sigma2, cam5):
    """Optimizer::PoseOptimization of the reference on a frame given as flat arrays (all float32 like the Frame's members).
    Returns dict(pose (float32), outlier[N], inliers)."""
    global _ref5_lib
    if _ref5_lib is None:
        if build_ref5() is None:
            raise RuntimeError("oracle/_ref/liborb_ref5.so is not built and /root/reference is not present")
        lib()
        _ref5_lib = C.CDLL(_REF5_SO)
        _ref5_lib.ref5_pose_optimization.restype = C.c_int
        _ref5_lib.ref5_pose_optimization.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 3
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    pose = f32(pose7).copy()
    hm, xw, xy = np.ascontiguousarray(has_mp, np.uint8), f32(world_pos), f32(kp_xy)
    oc, ur, isg, c5 = np.ascontiguousarray(octave, np.int32), f32(uright), f32(inv_level_sigma2), f32(cam5)
    N = len(hm)
    out = np.zeros(max(N, 1), np.uint8)
    inl = _ref5_lib.ref5_pose_optimization(N, _p(hm), _p(xw), _p(xy), _p(oc), _p(ur), _p(isg), len(isg), _p(c5), _p(pose), _p(out))
    return dict(pose=pose, outlier=out[:N], inliers=inl)

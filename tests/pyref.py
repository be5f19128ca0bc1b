"""Plain-Python restatements of a few ORBmatcher searches, written from the reference independently of oracle/ (second pin for
the C++ oracle; small inputs only).  float32 arithmetic is done with numpy scalars in the reference's operation order."""
import numpy as np

f32 = np.float32
POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def hamming(a, b):
    return int(POP[np.bitwise_xor(a, b)].sum())


class Grid:
    """Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea (src/Frame.cc:469-504, 859-951, 962-978), 64 x 48 cells."""

    def __init__(self, kps, bounds):
        self.k = kps
        self.minX, self.maxX, self.minY, self.maxY = [f32(v) for v in bounds]
        self.invW = f32(64) / f32(self.maxX - self.minX)
        self.invH = f32(48) / f32(self.maxY - self.minY)
        self.cells = {}
        for i in range(len(kps)):
            px = int(np.floor(float(f32(f32(kps["x"][i] - self.minX) * self.invW)) + 0.5)) if f32(f32(kps["x"][i] - self.minX) * self.invW) >= 0 else \
                -int(np.floor(-float(f32(f32(kps["x"][i] - self.minX) * self.invW)) + 0.5))          # C round(): half away from zero
            py = int(np.floor(float(f32(f32(kps["y"][i] - self.minY) * self.invH)) + 0.5)) if f32(f32(kps["y"][i] - self.minY) * self.invH) >= 0 else \
                -int(np.floor(-float(f32(f32(kps["y"][i] - self.minY) * self.invH)) + 0.5))
            if 0 <= px < 64 and 0 <= py < 48:
                self.cells.setdefault((px, py), []).append(i)

    def area(self, x, y, r, min_level=-1, max_level=-1):
        x, y, r = f32(x), f32(y), f32(r)
        c0 = max(0, int(np.floor(f32(f32(f32(x - self.minX) - r) * self.invW))))
        if c0 >= 64:
            return []
        c1 = min(63, int(np.ceil(f32(f32(f32(x - self.minX) + r) * self.invW))))
        if c1 < 0:
            return []
        r0 = max(0, int(np.floor(f32(f32(f32(y - self.minY) - r) * self.invH))))
        if r0 >= 48:
            return []
        r1 = min(47, int(np.ceil(f32(f32(f32(y - self.minY) + r) * self.invH))))
        if r1 < 0:
            return []
        check = (min_level > 0) or (max_level >= 0)
        out = []
        for ix in range(c0, c1 + 1):
            for iy in range(r0, r1 + 1):
                for i in self.cells.get((ix, iy), []):
                    o = int(self.k["octave"][i])
                    if check:
                        if o < min_level:
                            continue
                        if max_level >= 0 and o > max_level:
                            continue
                    if abs(f32(self.k["x"][i] - x)) < r and abs(f32(self.k["y"][i] - y)) < r:
                        out.append(i)
        return out


def three_maxima(hist):
    """ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:2335-2377)."""
    max1 = max2 = max3 = 0
    ind1 = ind2 = ind3 = -1
    for i, h in enumerate(hist):
        s = len(h)
        if s > max1:
            max3, max2, max1 = max2, max1, s
            ind3, ind2, ind1 = ind2, ind1, i
        elif s > max2:
            max3, max2 = max2, s
            ind3, ind2 = ind2, i
        elif s > max3:
            max3, ind3 = s, i
    if max2 < f32(0.1) * f32(max1):
        ind2 = ind3 = -1
    elif max3 < f32(0.1) * f32(max1):
        ind3 = -1
    return ind1, ind2, ind3


def rot_bin(a1, a2):
    rot = f32(f32(a1) - f32(a2))
    if rot < 0.0:
        rot = f32(rot + f32(360.0))
    v = float(f32(rot * f32(f32(1.0) / f32(30))))
    b = int(np.floor(v + 0.5))
    return 0 if b == 30 else b


def search_for_initialization(kp1, d1, prev, kp2, d2, bounds, window, nnratio, check_ori=True):
    """ORBmatcher::SearchForInitialization (ORBmatcher.cc:734-890)."""
    g = Grid(kp2, bounds)
    n1, n2 = len(kp1), len(kp2)
    m12, m21, md = [-1] * n1, [-1] * n2, [2 ** 31 - 1] * n2
    hist = [[] for _ in range(30)]
    nm = 0
    for i1 in range(n1):
        if kp1["octave"][i1] > 0:
            continue
        ind = g.area(prev[i1][0], prev[i1][1], f32(window), 0, 0)
        if not ind:
            continue
        best = best2 = 2 ** 31 - 1
        bi = -1
        for i2 in ind:
            d = hamming(d1[i1], d2[i2])
            if md[i2] <= d:
                continue
            if d < best:
                best2, best, bi = best, d, i2
            elif d < best2:
                best2 = d
        if best <= 50 and f32(best) < f32(f32(best2) * f32(nnratio)):
            if m21[bi] >= 0:
                m12[m21[bi]] = -1
                nm -= 1
            m12[i1], m21[bi], md[bi] = bi, i1, best
            nm += 1
            if check_ori:
                hist[rot_bin(kp1["angle"][i1], kp2["angle"][bi])].append(i1)
    if check_ori:
        keep = three_maxima(hist)
        for b in range(30):
            if b in keep:
                continue
            for i1 in hist[b]:
                if m12[i1] >= 0:
                    m12[i1] = -1
                    nm -= 1
    return np.array(m12, np.int32), nm


def search_bow_keyframes(kp2, d2, node2, valid2, qnode, qangle, d1, nnratio, check_ori=True):
    """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, ...) (ORBmatcher.cc:892-1043) on flattened word groups."""
    n2 = len(kp2)
    matched2 = [False] * n2
    out = [-1] * len(qnode)
    hist = [[] for _ in range(30)]
    nm = 0
    by_node = {}
    for i2 in range(n2):
        by_node.setdefault(int(node2[i2]), []).append(i2)
    for q in range(len(qnode)):
        if qnode[q] < 0:
            continue
        b1 = b2 = 256
        bi = -1
        for i2 in by_node.get(int(qnode[q]), []):
            if matched2[i2] or not valid2[i2]:
                continue
            d = hamming(d1[q], d2[i2])
            if d < b1:
                b2, b1, bi = b1, d, i2
            elif d < b2:
                b2 = d
        if b1 < 50 and f32(b1) < f32(f32(nnratio) * f32(b2)):
            out[q] = bi
            matched2[bi] = True
            if check_ori:
                hist[rot_bin(qangle[q], kp2["angle"][bi])].append(q)
            nm += 1
    if check_ori:
        keep = three_maxima(hist)
        for b in range(30):
            if b not in keep:
                for q in hist[b]:
                    out[q] = -1
                    nm -= 1
    return np.array(out, np.int32), nm


def se3_act(T, p):
    """Sophus::SE3f * Vector3f with Eigen's quaternion _transformVector (float32)."""
    qx, qy, qz, qw = [f32(v) for v in T[:4]]
    p = [f32(v) for v in p]
    uv = [f32(f32(qy * p[2]) - f32(qz * p[1])), f32(f32(qz * p[0]) - f32(qx * p[2])), f32(f32(qx * p[1]) - f32(qy * p[0]))]
    uv = [f32(v + v) for v in uv]
    c = [f32(f32(qy * uv[2]) - f32(qz * uv[1])), f32(f32(qz * uv[0]) - f32(qx * uv[2])), f32(f32(qx * uv[1]) - f32(qy * uv[0]))]
    return [f32(f32(f32(p[i] + f32(qw * uv[i])) + c[i]) + f32(T[4 + i])) for i in range(3)]


def fuse(variant, kps, desc, uright, bounds, sf, inv_sigma2, log_sf, cam, T, Ow, xw, normal, maxd, mind, qdesc, th, thr=50.0, claimed=None):
    """ORBmatcher::Fuse(pKF, vpMapPoints, th) (variant 0, ORBmatcher.cc:1325-1544), Fuse(pKF, Scw, ...) (variant 1, :1546-1687) and
    SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (variant 2, :495-618: features that already hold a match are
    skipped and every accepted match claims its feature; accepted when bestDist <= TH_LOW * ratioHamming = thr): the search part,
    bestIdx per map point or -1."""
    claimed = np.zeros(len(kps), bool) if claimed is None else np.asarray(claimed, bool).copy()
    import math
    import ctypes
    import ctypes.util
    libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    libm.logf.restype = ctypes.c_float
    libm.logf.argtypes = [ctypes.c_float]
    g = Grid(kps, bounds)
    fx, fy, cx, cy, bf = [f32(v) for v in cam[:5]]
    out = []
    for q in range(len(xw)):
        pc = se3_act(T, xw[q])
        res = -1
        while True:
            if pc[2] < 0:
                break
            invz = f32(f32(1) / pc[2])
            u = f32(f32(f32(fx * pc[0]) / pc[2]) + cx)
            v = f32(f32(f32(fy * pc[1]) / pc[2]) + cy)
            if not (u >= g.minX and u < g.maxX and v >= g.minY and v < g.maxY):
                break
            ur = f32(u - f32(bf * invz))
            PO = [f32(f32(xw[q][i]) - f32(Ow[i])) for i in range(3)]
            dist = f32(math.sqrt(float(f32(f32(PO[0] * PO[0]) + f32(f32(PO[1] * PO[1]) + f32(PO[2] * PO[2]))))))
            if dist < f32(f32(0.8) * f32(mind[q])) or dist > f32(f32(1.2) * f32(maxd[q])):   # Get{Min,Max}DistanceInvariance; mind / maxd are the raw members
                break
            dot = f32(f32(PO[0] * f32(normal[q][0])) + f32(f32(PO[1] * f32(normal[q][1])) + f32(PO[2] * f32(normal[q][2]))))
            if float(dot) < 0.5 * float(dist):
                break
            lvl = int(math.ceil(float(f32(f32(libm.logf(float(f32(f32(maxd[q]) / dist)))) / f32(log_sf)))))
            lvl = max(0, min(lvl, len(sf) - 1))
            radius = f32(f32(th) * f32(sf[lvl]))
            best, bi = (256, -1) if variant in (0, 2) else (2 ** 31 - 1, -1)
            for i in g.area(u, v, radius):
                if variant == 2 and claimed[i]:
                    continue
                o = int(kps["octave"][i])
                if o < lvl - 1 or o > lvl:
                    continue
                if variant == 0:
                    ex, ey = f32(u - kps["x"][i]), f32(v - kps["y"][i])
                    if uright is not None and uright[i] >= 0:
                        er = f32(ur - f32(uright[i]))
                        e2 = f32(f32(f32(ex * ex) + f32(ey * ey)) + f32(er * er))
                        if float(f32(e2 * f32(inv_sigma2[o]))) > 7.8:
                            continue
                    else:
                        e2 = f32(f32(ex * ex) + f32(ey * ey))
                        if float(f32(e2 * f32(inv_sigma2[o]))) > 5.99:
                            continue
                d = hamming(qdesc[q], desc[i])
                if d < best:
                    best, bi = d, i
            if variant == 2:
                if bi >= 0 and float(best) <= float(f32(thr)):
                    res = bi
                    claimed[bi] = True
            elif best <= 50:
                res = bi
            break
        out.append(res)
    return np.array(out, np.int32)


def search_local_points(kps, desc, uright, bounds, sf, projx, projy, projxr, level, viewcos, qdesc, th, nnratio, claimed=None):
    """ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>, th, ...) (ORBmatcher.cc:45-239), Nleft == -1; every query is an
    in-view map point with observations.  Returns (match[nq], nmatches)."""
    g = Grid(kps, bounds)
    holder = [bool(c) for c in claimed] if claimed is not None else [False] * len(kps)
    out = [-1] * len(projx)
    nm = 0
    for q in range(len(projx)):
        lvl = int(level[q])
        r = f32(2.5) if float(f32(viewcos[q])) > 0.998 else f32(4.0)
        if float(f32(th)) != 1.0:
            r = f32(r * f32(th))
        rr = f32(r * f32(sf[lvl]))
        ind = g.area(projx[q], projy[q], rr, lvl - 1, lvl)
        if not ind:
            continue
        b1 = b2 = 256
        l1 = l2 = -1
        bi = -1
        for i in ind:
            if holder[i]:
                continue
            if uright is not None and uright[i] > 0:
                er = abs(f32(f32(projxr[q]) - f32(uright[i])))
                if er > rr:
                    continue
            d = hamming(qdesc[q], desc[i])
            if d < b1:
                b2, b1, l2, l1, bi = b1, d, l1, int(kps["octave"][i]), i
            elif d < b2:
                l2, b2 = int(kps["octave"][i]), d
        if b1 <= 100:
            if l1 == l2 and f32(b1) > f32(f32(nnratio) * f32(b2)):
                continue
            if l1 != l2 or f32(b1) <= f32(f32(nnratio) * f32(b2)):
                out[q] = bi
                holder[bi] = True
                nm += 1
    return np.array(out, np.int32), nm


def search_last_frame(kps, desc, uright, bounds, sf, cam, T, direction, xw, last_octave, last_angle, qdesc, obs_pos, th, check_ori=True):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (ORBmatcher.cc:1950-2184), Nleft == -1.
    Returns (feature_match[N] query index or -1, nmatches)."""
    g = Grid(kps, bounds)
    fx, fy, cx, cy, bf = [f32(v) for v in cam[:5]]
    fm = [-1] * len(kps)
    hist = [[] for _ in range(30)]
    nm = 0
    for q in range(len(xw)):
        pc = se3_act(T, xw[q])
        invzc = f32(1.0 / float(pc[2]))
        if invzc < 0:
            continue
        u = f32(f32(f32(fx * pc[0]) / pc[2]) + cx)
        v = f32(f32(f32(fy * pc[1]) / pc[2]) + cy)
        if u < g.minX or u > g.maxX or v < g.minY or v > g.maxY:
            continue
        o = int(last_octave[q])
        radius = f32(f32(th) * f32(sf[o]))
        if direction == 1:
            ind = g.area(u, v, radius, o, -1)
        elif direction == 2:
            ind = g.area(u, v, radius, 0, o)
        else:
            ind = g.area(u, v, radius, o - 1, o + 1)
        if not ind:
            continue
        best, bi = 256, -1
        for i in ind:
            if fm[i] >= 0 and obs_pos[fm[i]]:
                continue
            if uright is not None and uright[i] > 0:
                ur = f32(u - f32(bf * invzc))
                if abs(f32(ur - f32(uright[i]))) > radius:
                    continue
            d = hamming(qdesc[q], desc[i])
            if d < best:
                best, bi = d, i
        if best <= 100:
            fm[bi] = q
            nm += 1
            if check_ori:
                hist[rot_bin(last_angle[q], kps["angle"][bi])].append(bi)
    if check_ori:
        keep = three_maxima(hist)
        for b in range(30):
            if b not in keep:
                for i in hist[b]:
                    fm[i] = -1
                    nm -= 1
    return np.array(fm, np.int32), nm


def c_round(v):
    """C round() of a float32 value: half away from zero."""
    v = float(v)
    return np.floor(v + 0.5) if v >= 0 else -np.floor(-v + 0.5)


def compute_stereo_matches(kL, dL, kR, dR, pyrL, pyrR, sf, inv_sf, bf, b):
    """Frame::ComputeStereoMatches (src/Frame.cc:1102-1358).  pyrL / pyrR: lists of the 8 level images (mvImagePyramid).
    Returns (mvuRight, mvDepth) as float32 arrays."""
    import math
    N = len(kL)
    uright, depth = np.full(N, -1, np.float32), np.full(N, -1, np.float32)
    nrows = pyrL[0].shape[0]
    rows = [[] for _ in range(nrows)]
    for iR in range(len(kR)):
        r = f32(f32(2.0) * f32(sf[kR["octave"][iR]]))
        maxr, minr = int(math.ceil(float(f32(kR["y"][iR] + r)))), int(math.floor(float(f32(kR["y"][iR] - r))))
        for y in range(minr, maxr + 1):
            rows[y].append(iR)
    minZ, minD = f32(b), f32(0)
    maxD = f32(f32(bf) / minZ)
    dist_idx = []
    for iL in range(N):
        lvl, vL, uL = int(kL["octave"][iL]), f32(kL["y"][iL]), f32(kL["x"][iL])
        cands = rows[int(vL)]
        if not cands:
            continue
        minU, maxU = f32(uL - maxD), f32(uL - minD)
        if maxU < 0:
            continue
        best, bestR = 100, 0
        for iR in cands:
            o = int(kR["octave"][iR])
            if o < lvl - 1 or o > lvl + 1:
                continue
            uR = f32(kR["x"][iR])
            if uR >= minU and uR <= maxU:
                d = hamming(dL[iL], dR[iR])
                if d < best:
                    best, bestR = d, iR
        if best >= 75:
            continue
        uR0 = f32(kR["x"][bestR])
        s = f32(inv_sf[lvl])
        suL, svL, suR0 = f32(c_round(f32(uL * s))), f32(c_round(f32(vL * s))), f32(c_round(f32(uR0 * s)))
        w, L = 5, 5
        IL = pyrL[lvl][int(svL) - w:int(svL) + w + 1, int(suL) - w:int(suL) + w + 1].astype(np.int32)
        iniu, endu = f32(f32(suR0 + L) - w), f32(f32(f32(suR0 + L) + w) + 1)
        if iniu < 0 or endu >= pyrR[lvl].shape[1]:
            continue
        bestDist, bestinc = 2 ** 31 - 1, 0
        vd = [f32(0)] * (2 * L + 1)
        for inc in range(-L, L + 1):
            c0 = int(f32(f32(suR0 + inc) - w))
            IR = pyrR[lvl][int(svL) - w:int(svL) + w + 1, c0:c0 + 2 * w + 1].astype(np.int32)
            dist = f32(float(np.abs(IL - IR).sum()))
            if dist < f32(bestDist):
                bestDist, bestinc = int(dist), inc
            vd[L + inc] = dist
        if bestinc == -L or bestinc == L:
            continue
        d1, d2, d3 = vd[L + bestinc - 1], vd[L + bestinc], vd[L + bestinc + 1]
        with np.errstate(divide="ignore", invalid="ignore"):
            delta = f32(f32(d1 - d3) / f32(f32(2.0) * f32(f32(d1 + d3) - f32(f32(2.0) * d2))))
        if delta < -1 or delta > 1 or np.isnan(delta):
            continue
        bestuR = f32(f32(sf[lvl]) * f32(f32(suR0 + f32(bestinc)) + delta))
        disp = f32(uL - bestuR)
        if disp >= minD and disp < maxD:
            if disp <= 0:
                disp = f32(0.01)
                bestuR = f32(float(uL) - 0.01)
            depth[iL] = f32(f32(bf) / disp)
            uright[iL] = bestuR
            dist_idx.append((bestDist, iL))
    dist_idx.sort()
    if dist_idx:
        median = f32(dist_idx[len(dist_idx) // 2][0])
        th = f32(f32(f32(1.5) * f32(1.4)) * median)
        for i in range(len(dist_idx) - 1, -1, -1):
            if f32(dist_idx[i][0]) < th:
                break
            uright[dist_idx[i][1]] = -1
            depth[dist_idx[i][1]] = -1
    return uright, depth


def search_bow_frame(kps, desc, feat_node, qnode, qangle, qdesc, nnratio, check_ori=True):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (ORBmatcher.cc:259-493), Nleft == -1, on flattened word
    groups: queries = keyframe features with a good map point in merge order.  Returns (feature_match[N], nmatches)."""
    N = len(kps)
    fm = [-1] * N
    by_node = {}
    for i in range(N):
        by_node.setdefault(int(feat_node[i]), []).append(i)
    hist = [[] for _ in range(30)]
    nm = 0
    for q in range(len(qnode)):
        if qnode[q] < 0:
            continue
        b1 = b2 = 256
        bi = -1
        for i in by_node.get(int(qnode[q]), []):
            if fm[i] >= 0:
                continue
            d = hamming(qdesc[q], desc[i])
            if d < b1:
                b2, b1, bi = b1, d, i
            elif d < b2:
                b2 = d
        if b1 <= 50 and f32(b1) < f32(f32(nnratio) * f32(b2)):
            fm[bi] = q
            if check_ori:
                hist[rot_bin(qangle[q], kps["angle"][bi])].append(bi)
            nm += 1
    if check_ori:
        keep = three_maxima(hist)
        for b in range(30):
            if b not in keep:
                for i in hist[b]:
                    fm[i] = -1
                    nm -= 1
    return np.array(fm, np.int32), nm


def bow_transform(voc, desc, levelsup=4):
    """DBoW2 TemplatedVocabulary::transform (TemplatedVocabulary.h:1127-1260) in plain numpy over the flat tree of
    orb_slam3_detailed_comments_b200.vocabulary (TF_IDF + L1).  Returns word, node, weight per feature, and the BowVector
    (ascending words, weights) plus the FeatureVector as {node: [features]} -- features of stopped words (weight 0) enter neither."""
    co, ch, nd, nw, wt, L = voc["child_offset"], voc["child_ids"], voc["node_desc"], voc["node_word"], voc["node_weight"], voc["L"]
    nid_level = L - levelsup
    n = len(desc)
    word, node, weight = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float64)
    bow, feat = {}, {}
    for i in range(n):
        fid, level, nid = 0, 0, 0
        while True:
            level += 1
            kids = ch[co[fid]:co[fid + 1]]
            d = POP[nd[kids] ^ desc[i][None, :]].sum(1)
            fid = int(kids[int(np.argmin(d))])          # first minimum: `if(d < best_d)`
            if level == nid_level:
                nid = fid
            if co[fid] == co[fid + 1]:
                break
        word[i], node[i], weight[i] = nw[fid], nid, wt[fid]
        if weight[i] > 0:
            bow[int(word[i])] = bow.get(int(word[i]), 0.0) + float(weight[i])      # addWeight in feature order
            feat.setdefault(nid, []).append(i)
    ws = sorted(bow)
    norm = 0.0
    for w in ws:                                        # BowVector::normalize(L1): ascending word order
        norm += abs(bow[w])
    vals = [bow[w] / norm for w in ws] if norm > 0 else [bow[w] for w in ws]
    return dict(word=word, node=node, weight=weight, bow_word=np.array(ws, np.int32), bow_weight=np.array(vals, np.float64), feat=feat)


def search_for_triangulation(kp1, d1, node1, stereo1, kp2, d2, node2, valid2, stereo2, F12, ep2, scale_factors, sigma2, coarse=False,
                             check_ori=True):
    """ORBmatcher::SearchForTriangulation (ORBmatcher.cc:1045-1323), single-camera keyframes, written as the reference walks it:
    FeatureVector node by node (std::map order), inside a node the KF1 features without a map point in index order, for each the
    KF2 features of the SAME node without a map point in index order.  `dist > TH_LOW || dist > bestDist` lets a later candidate at
    the same distance replace an earlier one; vbMatched2 is read but never set in this version of the reference, so a KF2 feature can
    be chosen by several KF1 features.  Epipolar gate as Pinhole::epipolarConstrain (Pinhole.cpp:186-216) from the caller's F12."""
    f32 = np.float32
    F = np.asarray(F12, f32).reshape(3, 3)
    TH_LOW, HISTO = 50, 30
    n1 = len(kp1)
    m12 = np.full(n1, -1, np.int32)
    hist = [[] for _ in range(HISTO)]
    feat2 = {}
    for i2 in range(len(kp2)):                                  # FeatureVector of KF2: node -> features in index order
        if node2[i2] >= 0:
            feat2.setdefault(int(node2[i2]), []).append(i2)
    feat1 = {}
    for q in range(n1):
        feat1.setdefault(int(node1[q]), []).append(q)
    nm = 0
    for node in sorted(set(feat1) & set(feat2)):
        for q in feat1[node]:
            best, best_i = TH_LOW, -1
            x1, y1 = f32(kp1["x"][q]), f32(kp1["y"][q])
            for i2 in feat2[node]:
                if not valid2[i2]:
                    continue
                dist = hamming(d1[q], d2[i2])
                if dist > TH_LOW or dist > best:
                    continue
                x2, y2 = f32(kp2["x"][i2]), f32(kp2["y"][i2])
                if not stereo1[q] and not stereo2[i2]:
                    ex, ey = f32(f32(ep2[0]) - x2), f32(f32(ep2[1]) - y2)
                    if f32(f32(ex * ex) + f32(ey * ey)) < f32(f32(100) * f32(scale_factors[kp2["octave"][i2]])):
                        continue
                ok = bool(coarse)
                if not ok:
                    a = f32(f32(f32(x1 * F[0, 0]) + f32(y1 * F[1, 0])) + F[2, 0])
                    b = f32(f32(f32(x1 * F[0, 1]) + f32(y1 * F[1, 1])) + F[2, 1])
                    c = f32(f32(f32(x1 * F[0, 2]) + f32(y1 * F[1, 2])) + F[2, 2])
                    num = f32(f32(f32(a * x2) + f32(b * y2)) + c)
                    den = f32(f32(a * a) + f32(b * b))
                    if den != 0:
                        dsqr = f32(f32(num * num) / den)
                        ok = float(dsqr) < 3.84 * float(f32(sigma2[kp2["octave"][i2]]))       # float < double * float
                if ok:
                    best, best_i = dist, i2
            if best_i >= 0:
                m12[q] = best_i
                nm += 1
                if check_ori:
                    hist[rot_bin(kp1["angle"][q], kp2["angle"][best_i])].append(q)
    if check_ori:
        keep = three_maxima(hist)
        for i in range(HISTO):
            if i not in keep:
                for q in hist[i]:
                    m12[q] = -1
                    nm -= 1
    return m12, nm


def search_by_projection_reloc(kps, desc, bounds, sf, log_sf, cam, T, Ow, xw, maxd, mind, qdesc, qangle, claimed, th, orb_dist, check_ori=True):
    """ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:2196-2330), pinhole: the keyframe's
    map points (those not bad and not already found) projected with the frame's pose -- no depth-sign test, inclusive image bounds,
    no viewing-angle test -- GetFeaturesInArea restricted to levels [n - 1, n + 1], features that already hold a map point skipped,
    every accepted match claims its feature, rotation histogram over (keyframe keypoint angle - frame keypoint angle) with the
    losers un-claimed at the end.  Returns (match per query or -1, claimed after)."""
    import ctypes
    import ctypes.util
    import math
    libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    libm.logf.restype = ctypes.c_float
    libm.logf.argtypes = [ctypes.c_float]
    g = Grid(kps, bounds)
    fx, fy, cx, cy = [f32(v) for v in cam[:4]]
    claimed = np.asarray(claimed, bool).copy()
    match = np.full(len(xw), -1, np.int32)
    hist = [[] for _ in range(30)]
    for q in range(len(xw)):
        pc = se3_act(T, xw[q])
        u = f32(f32(f32(fx * pc[0]) / pc[2]) + cx)
        v = f32(f32(f32(fy * pc[1]) / pc[2]) + cy)
        if u < g.minX or u > g.maxX or v < g.minY or v > g.maxY:
            continue
        PO = [f32(f32(xw[q][i]) - f32(Ow[i])) for i in range(3)]
        dist = f32(math.sqrt(float(f32(f32(PO[0] * PO[0]) + f32(f32(PO[1] * PO[1]) + f32(PO[2] * PO[2]))))))
        if dist < f32(f32(0.8) * f32(mind[q])) or dist > f32(f32(1.2) * f32(maxd[q])):   # Get{Min,Max}DistanceInvariance; mind / maxd are the raw members
            continue
        lvl = int(math.ceil(float(f32(f32(libm.logf(float(f32(f32(maxd[q]) / dist)))) / f32(log_sf)))))
        lvl = max(0, min(lvl, len(sf) - 1))
        radius = f32(f32(th) * f32(sf[lvl]))
        best, bi = 256, -1
        for i in g.area(u, v, radius, lvl - 1, lvl + 1):
            if claimed[i]:
                continue
            d = hamming(qdesc[q], desc[i])
            if d < best:
                best, bi = d, i
        if best <= orb_dist:
            match[q] = bi
            claimed[bi] = True
            if check_ori:
                hist[rot_bin(qangle[q], kps["angle"][bi])].append((bi, q))
    if check_ori:
        keep = three_maxima(hist)
        for b in range(30):
            if b not in keep:
                for bi, q in hist[b]:
                    claimed[bi] = False
                    match[q] = -1
    return match, claimed


def search_by_sim3_oneway(kps, desc, bounds, sf, log_sf, cam, T, S8, xw, maxd, mind, qdesc, th):
    """One direction of ORBmatcher::SearchBySim3 (ORBmatcher.cc:1719-1790): the map points of the OTHER keyframe, taken to its camera
    frame by its pose T (SE3f) and to this keyframe's camera frame by the similarity S = (non-unit quaternion x y z w, translation,
    scale = |q|^2): Sophus RxSO3::operator* -- scale p + (w 2(v x p) + v x 2(v x p)) -- plus the translation (rxso3.hpp:265-273,
    sim3.hpp:227-230).  Depth sign, KeyFrame::IsInImage, distance range on |p_c|, PredictScale, GetFeaturesInArea, octave in
    [n - 1, n], best Hamming distance <= TH_HIGH.  No claims.  Returns the feature index per query or -1."""
    import ctypes
    import ctypes.util
    import math
    libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    libm.logf.restype = ctypes.c_float
    libm.logf.argtypes = [ctypes.c_float]
    g = Grid(kps, bounds)
    fx, fy, cx, cy = [f32(v) for v in cam[:4]]
    qx, qy, qz, qw = [f32(v) for v in S8[:4]]
    t = [f32(v) for v in S8[4:7]]
    scale = f32(S8[7])
    out = np.full(len(xw), -1, np.int32)
    for q in range(len(xw)):
        p = se3_act(T, xw[q])
        c = [f32(f32(qy * p[2]) - f32(qz * p[1])), f32(f32(qz * p[0]) - f32(qx * p[2])), f32(f32(qx * p[1]) - f32(qy * p[0]))]
        u2 = [f32(v + v) for v in c]
        cc = [f32(f32(qy * u2[2]) - f32(qz * u2[1])), f32(f32(qz * u2[0]) - f32(qx * u2[2])), f32(f32(qx * u2[1]) - f32(qy * u2[0]))]
        pc = [f32(f32(f32(scale * p[i]) + f32(f32(qw * u2[i]) + cc[i])) + t[i]) for i in range(3)]
        if pc[2] < 0.0:
            continue
        invz = f32(1.0 / float(pc[2]))                          # const float invz = 1.0 / p3Dc2(2): double division, rounded
        u = f32(f32(fx * f32(pc[0] * invz)) + cx)
        v = f32(f32(fy * f32(pc[1] * invz)) + cy)
        if not (u >= g.minX and u < g.maxX and v >= g.minY and v < g.maxY):
            continue
        dist = f32(math.sqrt(float(f32(f32(pc[0] * pc[0]) + f32(f32(pc[1] * pc[1]) + f32(pc[2] * pc[2]))))))
        if dist < f32(f32(0.8) * f32(mind[q])) or dist > f32(f32(1.2) * f32(maxd[q])):   # Get{Min,Max}DistanceInvariance; mind / maxd are the raw members
            continue
        lvl = int(math.ceil(float(f32(f32(libm.logf(float(f32(f32(maxd[q]) / dist)))) / f32(log_sf)))))
        lvl = max(0, min(lvl, len(sf) - 1))
        radius = f32(f32(th) * f32(sf[lvl]))
        best, bi = 2 ** 31 - 1, -1
        for i in g.area(u, v, radius):
            o = int(kps["octave"][i])
            if o < lvl - 1 or o > lvl:
                continue
            d = hamming(qdesc[q], desc[i])
            if d < best:
                best, bi = d, i
        if best <= 100:
            out[q] = bi
    return out

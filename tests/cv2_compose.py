"""Independent Python/cv2 composition of the ORB half of the reference pipeline, used to pin the C++ oracle.
OpenCV calls are the real cv2 4.13 functions the reference would call in C++ (src/ORBextractor.cc:811,1088,1122);
the vendored parts (cell loop, quadtree, IC_Angle, rBRIEF) are restated a second time here, in Python, straight from
the reference lines cited, so that two independent restatements have to agree."""
import math, re, pathlib
import numpy as np
import cv2

f32 = np.float32


def brief_pattern():
    txt = (pathlib.Path(__file__).resolve().parents[1] / "include" / "olf_brief_pattern.h").read_text()
    body = txt[txt.index("{") + 1: txt.index("};")]
    return np.array([int(t) for t in re.findall(r"-?\d+", body)], dtype=np.int32).reshape(256, 4)


def scales(nlevels=8, sf=1.2):
    s = [f32(1.0)]
    for _ in range(1, nlevels):
        s.append(f32(s[-1] * f32(sf)))
    return s, [f32(1.0) / x for x in s]


def features_per_level(nfeatures, nlevels=8, sf=1.2):
    factor = f32(1.0) / f32(sf)
    nd = f32(f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(pow(float(factor), float(nlevels)))))
    out, tot = [], 0
    for _ in range(nlevels - 1):
        out.append(int(np.rint(nd))); tot += out[-1]; nd = f32(nd * factor)
    out.append(max(nfeatures - tot, 0))
    return out


def pyramid(img, nlevels=8, sf=1.2):
    _, inv = scales(nlevels, sf)
    h, w = img.shape
    pyr = [img]
    for l in range(1, nlevels):
        dw, dh = int(np.rint(f32(w) * inv[l])), int(np.rint(f32(h) * inv[l]))
        pyr.append(cv2.resize(pyr[-1], (dw, dh), interpolation=cv2.INTER_LINEAR))
    return pyr


def cell_candidates(im, ini_th=20, min_th=7):
    """src/ORBextractor.cc:773-831 with real cv2 FAST per cell.  Returns [(x,y,score)] relative to minBorder."""
    fast_hi = cv2.FastFeatureDetector_create(ini_th, True)
    fast_lo = cv2.FastFeatureDetector_create(min_th, True)
    H, Wd = im.shape
    minBX = minBY = 16
    maxBX, maxBY = Wd - 16, H - 16
    width, height = f32(maxBX - minBX), f32(maxBY - minBY)
    if width < 30 or height < 30:
        return []
    nCols, nRows = int(width / f32(30)), int(height / f32(30))
    wCell, hCell = int(math.ceil(width / nCols)), int(math.ceil(height / nRows))
    out = []
    for i in range(nRows):
        iniY = minBY + i * hCell
        maxY = iniY + hCell + 6
        if iniY >= maxBY - 3:
            continue
        maxY = min(maxY, maxBY)
        for j in range(nCols):
            iniX = minBX + j * wCell
            maxX = iniX + wCell + 6
            if iniX >= maxBX - 6:
                continue
            maxX = min(maxX, maxBX)
            win = im[iniY:maxY, iniX:maxX]
            k = fast_hi.detect(win)
            if not k:
                k = fast_lo.detect(win)
            for p in k:
                out.append((p.pt[0] + j * wCell, p.pt[1] + i * hCell, p.response))
    return out


class _Node:
    __slots__ = ("UL", "UR", "BL", "BR", "keys", "no_more", "seq", "alive")


def quadtree(cands, minX, maxX, minY, maxY, N):
    """DistributeOctTree (src/ORBextractor.cc:541-765) on a python list acting as std::list (index 0 = front).
    Tie-break for equal-size nodes: creation sequence."""
    seq = [0]

    def mk(UL, UR, BL, BR):
        n = _Node(); n.UL, n.UR, n.BL, n.BR = UL, UR, BL, BR
        n.keys = []; n.no_more = False; n.alive = True; n.seq = seq[0]; seq[0] += 1
        return n

    def divide(n):
        halfX = int(math.ceil(f32(n.UR[0] - n.UL[0]) / 2)); halfY = int(math.ceil(f32(n.BR[1] - n.UL[1]) / 2))
        n1 = mk(n.UL, (n.UL[0] + halfX, n.UL[1]), (n.UL[0], n.UL[1] + halfY), (n.UL[0] + halfX, n.UL[1] + halfY))
        n2 = mk(n1.UR, n.UR, n1.BR, (n.UR[0], n.UL[1] + halfY))
        n3 = mk(n1.BL, n1.BR, n.BL, (n1.BR[0], n.BL[1]))
        n4 = mk(n3.UR, n2.BR, n3.BR, n.BR)
        for idx in n.keys:
            x, y = cands[idx][0], cands[idx][1]
            if x < n1.UR[0]:
                (n1 if y < n1.BR[1] else n3).keys.append(idx)
            elif y < n1.BR[1]:
                n2.keys.append(idx)
            else:
                n4.keys.append(idx)
        for c in (n1, n2, n3, n4):
            if len(c.keys) == 1:
                c.no_more = True
        return n1, n2, n3, n4

    nIni = int(np.round(f32(maxX - minX) / f32(maxY - minY)))  # C round(): half away from zero; values here are never .5 exactly
    nIni = max(nIni, 1)
    hX = f32(maxX - minX) / f32(nIni)
    nodes = []
    for i in range(nIni):
        nodes.append(mk((int(hX * f32(i)), 0), (int(hX * f32(i + 1)), 0), (int(hX * f32(i)), maxY - minY), (int(hX * f32(i + 1)), maxY - minY)))
    roots = list(nodes)
    for idx, c in enumerate(cands):
        roots[min(int(f32(c[0]) / hX), nIni - 1)].keys.append(idx)
    nodes = [n for n in nodes if n.keys]
    for n in nodes:
        if len(n.keys) == 1:
            n.no_more = True
    finish = False
    while not finish:
        prev = len(nodes)
        expand = []
        new_front = []           # children pushed to the front: later pushes come first
        kept = []
        n_to_expand = 0
        for n in nodes:
            if n.no_more:
                kept.append(n); continue
            for c in divide(n):
                if c.keys:
                    new_front.insert(0, c)
                    if len(c.keys) > 1:
                        n_to_expand += 1; expand.append(c)
        nodes = new_front + kept
        if len(nodes) >= N or len(nodes) == prev:
            finish = True
        elif len(nodes) + n_to_expand * 3 > N:
            while not finish:
                prev = len(nodes)
                order = sorted(expand, key=lambda n: (len(n.keys), n.seq))
                expand = []
                for n in reversed(order):
                    for c in divide(n):
                        if c.keys:
                            nodes.insert(0, c)
                            if len(c.keys) > 1:
                                expand.append(c)
                    nodes.remove(n)
                    if len(nodes) >= N:
                        break
                if len(nodes) >= N or len(nodes) == prev:
                    finish = True
    res = []
    for n in nodes:
        best = n.keys[0]
        for k in n.keys[1:]:
            if cands[k][2] > cands[best][2]:
                best = k
        res.append(best)
    return res


UMAX = [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


def ic_angle(im, x, y):
    m01 = m10 = 0
    for v in range(-15, 16):
        d = UMAX[abs(v)]
        row = im[y + v, x - d:x + d + 1].astype(np.int64)
        u = np.arange(-d, d + 1)
        m10 += int((u * row).sum()); m01 += v * int(row.sum())
    return cv2.fastAtan2(float(m01), float(m10))


def rbrief(blur, x, y, angle_deg, pat):
    ang = f32(f32(angle_deg) * f32(np.pi / 180.0))
    a, b = f32(math.cos(float(ang))), f32(math.sin(float(ang)))   # replaced below by libm cosf/sinf
    return a, b


def orb_extract(img, nfeatures=1000, nlevels=8, sf=1.2, ini_th=20, min_th=7, cosf=None, sinf=None):
    """Full ORBextractor::operator() composition.  cosf/sinf: float libm functions (ctypes) -- Python's math.cos is
    double precision, the reference calls the float overload."""
    pat = brief_pattern().astype(np.float32)
    sc, _ = scales(nlevels, sf)
    fpl = features_per_level(nfeatures, nlevels, sf)
    pyr = pyramid(img, nlevels, sf)
    kps, descs, cands_all = [], [], []
    for l, im in enumerate(pyr):
        cands = cell_candidates(im, ini_th, min_th)
        cands_all.append(cands)
        if not cands:
            continue
        H, Wd = im.shape
        keep = quadtree(cands, 16, Wd - 16, 16, H - 16, fpl[l])
        blur = cv2.GaussianBlur(im.copy(), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        for k in keep:
            x, y = int(cands[k][0]) + 16, int(cands[k][1]) + 16
            ang = ic_angle(im, x, y)
            ar = f32(f32(ang) * f32(np.pi / f32(180.0)))
            a, b = f32(cosf(ar)), f32(sinf(ar))
            d = np.zeros(32, np.uint8)
            xs0 = np.rint(pat[:, 0] * a - pat[:, 1] * b).astype(np.int64); ys0 = np.rint(pat[:, 0] * b + pat[:, 1] * a).astype(np.int64)
            xs1 = np.rint(pat[:, 2] * a - pat[:, 3] * b).astype(np.int64); ys1 = np.rint(pat[:, 2] * b + pat[:, 3] * a).astype(np.int64)
            bits = (blur[y + ys0, x + xs0] < blur[y + ys1, x + xs1]).astype(np.uint8)
            d = np.packbits(bits, bitorder="little")
            px, py = f32(x), f32(y)
            if l:
                px, py = f32(px * sc[l]), f32(py * sc[l])
            kps.append((px, py, f32(int(f32(31) * sc[l])), f32(ang), f32(cands[k][2]), l))
            descs.append(d)
    return kps, (np.array(descs, np.uint8).reshape(-1, 32)), cands_all, pyr

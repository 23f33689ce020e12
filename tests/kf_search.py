"""Test data + host prologues for the SURVEY 8f rank-2 matchers (Fuse, SearchByProjection(KF, Scw), SearchBySim3,
SearchByProjection(Frame, KF), SearchForTriangulation, SearchByBoW(KF, KF)).

`make_keyframe` builds a synthetic key frame (keypoints on the 64 x 48 grid of a 1280 x 720 image, octaves, uRight, descriptors
drawn from a small pool so that Hamming ties are frequent and the reference's enumeration order decides).  `project_points` is a
float32 restatement of the per-point prologue the five projection overloads share (src/ORBmatcher.cc:854-892, 1011-1051,
318-342, 1160-1191, 1645-1677): every operation is a single IEEE float32 operation in the reference's order (numpy scalar
arithmetic), `cv::norm` / `Mat::dot` accumulate in double, `logf` comes from the C library like in the reference's build."""
import ctypes
import numpy as np
from orb_line_slam_b200.abi import KEYPOINT, Camera

_libm = ctypes.CDLL("libm.so.6")
_libm.logf.restype = ctypes.c_float; _libm.logf.argtypes = [ctypes.c_float]
f32 = np.float32


def logf(x):
    return f32(_libm.logf(ctypes.c_float(float(x))))


def scale_tables(nlevels=8, sf=1.2):
    s = np.ones(nlevels, np.float32); sig = np.ones(nlevels, np.float32)
    for i in range(1, nlevels):
        s[i] = f32(s[i - 1] * f32(sf)); sig[i] = f32(s[i] * s[i])                                    # src/ORBextractor.cc:417-424
    inv_sig = (f32(1.0) / sig).astype(np.float32)
    return s, sig, inv_sig, logf(f32(sf))


def desc_pool(rng, n, pool=48, max_flips=3):
    """n descriptors = one of `pool` random base descriptors with 0..max_flips flipped bits: many exact ties."""
    base = rng.randint(0, 256, (pool, 32)).astype(np.uint8)
    pick = rng.randint(0, pool, n)
    d = base[pick].copy()
    for i in range(n):
        for b in rng.randint(0, 256, rng.randint(0, max_flips + 1)):
            d[i, b >> 3] ^= np.uint8(1 << (b & 7))
    return d, pick


def make_keyframe(seed, n=1500, w=1280, h=720, nlevels=8, stereo_fraction=0.7, pool=48):
    rng = np.random.RandomState(seed)
    kps = np.zeros(n, KEYPOINT)
    kps["x"] = (rng.rand(n) * (w - 1)).astype(np.float32); kps["y"] = (rng.rand(n) * (h - 1)).astype(np.float32)
    kps["octave"] = rng.randint(0, nlevels, n); kps["angle"] = (rng.rand(n) * 360).astype(np.float32)
    kps["size"] = 31; kps["response"] = 20
    desc, pick = desc_pool(rng, n, pool)
    ur = np.where(rng.rand(n) < stereo_fraction, kps["x"] - (5 + 40 * rng.rand(n)).astype(np.float32), np.float32(-1)).astype(np.float32)
    cam = Camera(670.44, 670.44, 640.0, 360.0, 80.45, 0.0, float(w), 0.0, float(h))
    return dict(kps=kps, desc=desc, u_right=ur, cam=cam, w=w, h=h, pool_pick=pick, rng=rng)


def random_queries(kf, nq, seed, max_radius=45.0, pool=48):
    """Window queries around the key frame's keypoints (no geometry): positions jittered, descriptor of the pool."""
    rng = np.random.RandomState(seed)
    j = rng.randint(0, len(kf["kps"]), nq)
    jit = np.where(rng.rand(nq) < 0.5, 1.2, 6.0)                                        # half of them inside the chi-square gate of Fuse
    u = (kf["kps"]["x"][j] + rng.randn(nq) * jit).astype(np.float32); v = (kf["kps"]["y"][j] + rng.randn(nq) * jit).astype(np.float32)
    radius = (3 + rng.rand(nq) * (max_radius - 3)).astype(np.float32)
    lvl = rng.randint(0, 8, nq).astype(np.int32)
    qd = kf["desc"][j].copy()
    for i in range(nq):
        for b in rng.randint(0, 256, rng.randint(0, 30)):
            qd[i, b >> 3] ^= np.uint8(1 << (b & 7))
    ur = np.where(kf["u_right"][j] >= 0, kf["u_right"][j] + rng.randn(nq) * jit, u - 20).astype(np.float32)
    return dict(u=u, v=v, radius=radius, ur=ur, min_level=(lvl - 1).astype(np.int32), max_level=lvl, qdesc=qd)


def mat3_vec(R, x):
    """cv::Mat float gemm, 3x3 * 3x1: per row ((r0*x0 + r1*x1) + r2*x2) in float32 (pinned by tests/golden gemm vectors)."""
    return np.array([f32(f32(f32(R[r, 0] * x[0]) + f32(R[r, 1] * x[1])) + f32(R[r, 2] * x[2])) for r in range(3)], np.float32)


def norm3(x):
    """cv::norm(Mat 3x1 CV_32F): double accumulation of squares, sqrt in double; the callers assign it to a float."""
    s = 0.0
    for k in range(3):
        s += float(x[k]) * float(x[k])
    return f32(np.sqrt(s))


def dot3(a, b):
    """Mat::dot for CV_32F: products and sum in double."""
    s = 0.0
    for k in range(3):
        s += float(a[k]) * float(b[k])
    return s


def predict_scale(max_distance, dist, log_scale_factor, nlevels):
    """MapPoint::PredictScale (src/MapPoint.cc:397-429): ceil(logf(mfMaxDistance / dist) / mfLogScaleFactor), clamped."""
    ratio = f32(f32(max_distance) / f32(dist))
    n = int(np.ceil(f32(logf(ratio) / f32(log_scale_factor))))
    return 0 if n < 0 else min(n, nlevels - 1)


def project_points(mode, P, Rcw, tcw, Ow, cam: Camera, scale_factors, log_sf, th):
    """The prologue of one overload for every map point of P = dict(pos [n,3], normal [n,3], max_d, min_d (raw mfMax/MinDistance)).
    mode: 'fuse' (:854-892), 'fuse_sim3' (:1011-1051), 'sbp_kf' (:318-342) -- depth > 0, KeyFrame::IsInImage, distance window on |P - Ow|,
    viewing angle; 'sim3' (:1160-1191): P['pos_c'] is already in the target camera frame, distance = |p|; 'reloc' (:1645-1677): no
    depth test, Frame bounds (inclusive), distance window, no angle test.  Returns per point: ok, u, v, ur, level, radius."""
    n = len(P["pos"])
    nl = len(scale_factors)
    ok = np.zeros(n, bool); u = np.zeros(n, np.float32); v = np.zeros(n, np.float32); ur = np.zeros(n, np.float32)
    lvl = np.zeros(n, np.int32); radius = np.zeros(n, np.float32)
    fx, fy, cx, cy, bf = f32(cam.fx), f32(cam.fy), f32(cam.cx), f32(cam.cy), f32(cam.bf)
    for i in range(n):
        pw = P["pos"][i].astype(np.float32)
        if mode == "sim3":
            pc = P["pos_c"][i].astype(np.float32)
        else:
            pc = (mat3_vec(Rcw, pw) + tcw).astype(np.float32)
        if mode != "reloc" and pc[2] < 0.0:
            continue
        # `1/z` is a float division in Fuse(KF,MPs) and SearchByProjection(KF,Scw); `1.0/z` a double division rounded to float elsewhere
        invz = f32(f32(1.0) / pc[2]) if mode in ("fuse", "sbp_kf") else f32(1.0 / float(pc[2]))
        if mode == "reloc":
            ui = f32(f32(f32(fx * pc[0]) * invz) + cx); vi = f32(f32(f32(fy * pc[1]) * invz) + cy)           # fx*xc*invzc+cx (:1652)
            if ui < cam.min_x or ui > cam.max_x or vi < cam.min_y or vi > cam.max_y:
                continue
        else:
            x = f32(pc[0] * invz); y = f32(pc[1] * invz)
            ui = f32(f32(fx * x) + cx); vi = f32(f32(fy * y) + cy)
            if not (ui >= cam.min_x and ui < cam.max_x and vi >= cam.min_y and vi < cam.max_y):               # KeyFrame::IsInImage
                continue
        if mode == "sim3":
            dist = norm3(pc)
        else:
            PO = (pw - Ow).astype(np.float32)
            dist = norm3(PO)
        maxd = f32(f32(1.2) * f32(P["max_d"][i])); mind = f32(f32(0.8) * f32(P["min_d"][i]))
        if dist < mind or dist > maxd:
            continue
        if mode in ("fuse", "fuse_sim3", "sbp_kf"):
            if dot3(PO, P["normal"][i].astype(np.float32)) < 0.5 * float(dist):
                continue
        L = predict_scale(P["max_d"][i], dist, log_sf, nl)
        ok[i] = True; u[i] = ui; v[i] = vi; ur[i] = f32(ui - f32(bf * invz)); lvl[i] = L
        radius[i] = f32(f32(th) * scale_factors[L])
    return ok, u, v, ur, lvl, radius


def make_points(kf, seed, n=800, pose_noise=0.0):
    """Map points seen by the key frame: 3-D points behind a subset of its keypoints at depths 2..30 m, in WORLD coordinates of a
    random pose; descriptors of the pool near the keypoint's.  Returns (P, Rcw, tcw, Ow)."""
    rng = np.random.RandomState(seed)
    cam = kf["cam"]
    a = rng.randn(3) * 0.05
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    Rcw = (np.eye(3) + K + K @ K / 2)
    Uo, _, Vo = np.linalg.svd(Rcw); Rcw = (Uo @ Vo).astype(np.float32)
    tcw = (rng.randn(3) * 0.3).astype(np.float32)
    Ow = (-(Rcw.T.astype(np.float64) @ tcw.astype(np.float64))).astype(np.float32)
    j = rng.randint(0, len(kf["kps"]), n)
    z = 2 + 28 * rng.rand(n)
    xs = kf["kps"]["x"][j] + rng.randn(n) * 2.0; ys = kf["kps"]["y"][j] + rng.randn(n) * 2.0
    pc = np.stack([(xs - cam.cx) / cam.fx * z, (ys - cam.cy) / cam.fy * z, z], 1)
    pc[rng.rand(n) < 0.05, 2] *= -1                                                    # some behind the camera
    pw = ((pc - tcw.astype(np.float64)) @ Rcw.astype(np.float64)).astype(np.float32)  # Rcw^T (pc - tcw)
    d = np.linalg.norm(pw - Ow, axis=1)
    view = (Ow - pw) / np.maximum(d[:, None], 1e-6)
    normal = (-view + rng.randn(n, 3) * 0.4)
    normal = (normal / np.linalg.norm(normal, axis=1)[:, None]).astype(np.float32)
    normal[rng.rand(n) < 0.1] *= -1                                                    # some seen from behind
    max_d = (d * (1.0 + 2.0 * rng.rand(n))).astype(np.float32)
    min_d = (max_d / f32(1.2 ** 7)).astype(np.float32)
    far = rng.rand(n) < 0.05
    max_d[far] = (d[far] * 0.5).astype(np.float32)                                     # outside the scale-invariance range
    qd = kf["desc"][j].copy()
    for i in range(n):
        for b in rng.randint(0, 256, rng.randint(0, 40)):
            qd[i, b >> 3] ^= np.uint8(1 << (b & 7))
    return dict(pos=pw, normal=normal, max_d=max_d, min_d=min_d, desc=qd), Rcw, tcw, Ow


def feature_vector(rng, n, n_nodes=60, drop=0.03):
    """A DBoW2 FeatureVector as CSR (node ascending, indices ascending inside a node); a few features carry no word."""
    node = rng.randint(0, n_nodes, n) * 7 + 3
    keep = rng.rand(n) > drop
    nodes = np.unique(node[keep])
    begin = [0]; index = []
    for nd in nodes:
        idx = np.nonzero((node == nd) & keep)[0]
        index.extend(idx.tolist()); begin.append(len(index))
    return nodes.astype(np.int32), np.array(begin, np.int32), np.array(index, np.int32)


def make_stereo_pair_keyframes(seed, n=1200, pool_noise=12):
    """Two key frames looking at the same 3-D points from two poses + the fundamental matrix F12 (x1^T F12 x2 = 0) and the
    epipole of camera 1 in image 2 -- the inputs of SearchForTriangulation (src/ORBmatcher.cc:659-672)."""
    rng = np.random.RandomState(seed)
    cam = Camera(670.44, 670.44, 640.0, 360.0, 80.45, 0.0, 1280.0, 0.0, 720.0)
    Kc = np.array([[cam.fx, 0, cam.cx], [0, cam.fy, cam.cy], [0, 0, 1.0]])
    a = rng.randn(3) * 0.03
    Kx = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    R21 = np.eye(3) + Kx + Kx @ Kx / 2
    Uo, _, Vo = np.linalg.svd(R21); R21 = Uo @ Vo                                     # camera 2 from camera 1
    t21 = np.array([0.4, 0.05, 0.6]) * (1 + rng.rand(3) * 0.2)
    z = 3 + 25 * rng.rand(n)
    x1 = rng.rand(n) * 1279; y1 = rng.rand(n) * 719
    p1 = np.stack([(x1 - cam.cx) / cam.fx * z, (y1 - cam.cy) / cam.fy * z, z], 1)
    p2 = p1 @ R21.T + t21
    x2 = cam.fx * p2[:, 0] / p2[:, 2] + cam.cx; y2 = cam.fy * p2[:, 1] / p2[:, 2] + cam.cy
    noise = rng.randn(n, 2) * np.where(rng.rand(n) < 0.8, 0.4, 6.0)[:, None]          # 20 % violate the epipolar constraint
    base = rng.randint(0, 256, (n, 32)).astype(np.uint8)
    dup = rng.rand(n) < 0.3
    base[dup] = base[rng.randint(0, n, dup.sum())]                                    # repeated texture: equal descriptors
    def kf(xs, ys, shuffle):
        k = np.zeros(n, KEYPOINT)
        k["x"] = xs.astype(np.float32); k["y"] = ys.astype(np.float32); k["octave"] = rng.randint(0, 8, n)
        k["angle"] = (rng.rand(n) * 40 + 100).astype(np.float32) % 360; k["size"] = 31
        d = base.copy()
        for i in range(n):
            for b in rng.randint(0, 256, rng.randint(0, pool_noise)):
                d[i, b >> 3] ^= np.uint8(1 << (b & 7))
        ur = np.where(rng.rand(n) < 0.6, k["x"] - 10, np.float32(-1)).astype(np.float32)
        skip = (rng.rand(n) < 0.2).astype(np.uint8)
        return dict(kps=k[shuffle], desc=d[shuffle], u_right=ur[shuffle], skip=skip[shuffle])
    perm = rng.permutation(n)
    kf1 = kf(x1, y1, np.arange(n)); kf2 = kf(x2 + noise[:, 0], y2 + noise[:, 1], perm)
    node_of_point = rng.randint(0, 50, n)                                             # both views of a point fall into the same node (mostly)
    def fv(node):
        keep = rng.rand(n) > 0.03
        nodes = np.unique(node[keep]); begin = [0]; index = []
        for nd in nodes:
            idx = np.nonzero((node == nd) & keep)[0]; index.extend(idx.tolist()); begin.append(len(index))
        return nodes.astype(np.int32), np.array(begin, np.int32), np.array(index, np.int32)
    kf1["fv"] = fv(node_of_point); kf2["fv"] = fv(node_of_point[perm])
    s, sig, inv_sig, log_sf = scale_tables()
    kf2["scale_factors"] = s; kf2["level_sigma2"] = sig
    # F12 = K^-T [t12]x R12 K^-1 with (R12, t12) = camera 1 from camera 2 (src/LocalMapping.cc ComputeF12)
    R12 = R21.T; t12 = -R21.T @ t21
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    F12 = (np.linalg.inv(Kc).T @ tx @ R12 @ np.linalg.inv(Kc)).astype(np.float32)
    C2 = t21                                                                          # camera-1 centre in camera-2 coordinates
    ex = f32(cam.fx * C2[0] / C2[2] + cam.cx); ey = f32(cam.fy * C2[1] / C2[2] + cam.cy)
    return kf1, kf2, F12, ex, ey

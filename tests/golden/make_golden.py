#!/usr/bin/env python3
"""Generate tests/golden/cv2_golden.npz: outputs of the REAL OpenCV functions the reference calls (cv2 4.13.0 here)
on small seeded images, plus the cv2-composed ORB pipeline (tests/cv2_compose.py).  The oracle is pinned against
these vectors by tests/test_oracle_golden.py (which needs neither cv2 nor /root/reference, so it also runs on the
GPU box).  Re-run in the build container only:  python tests/golden/make_golden.py
"""
import sys, pathlib, ctypes as C
import numpy as np
import cv2
ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import cv2_compose as cc
from orb_line_slam_b200.synth import random_image

cv2.setNumThreads(1)
out = {"cv2_version": np.array(cv2.__version__)}
libm = C.CDLL("libm.so.6"); libm.cosf.restype = C.c_float; libm.sinf.restype = C.c_float
cosf = lambda v: libm.cosf(C.c_float(v)); sinf = lambda v: libm.sinf(C.c_float(v))

cases = [(160, 120, 1), (97, 131, 2), (240, 180, 3)]
for ci, (w, h, seed) in enumerate(cases):
    img = random_image(w, h, seed)
    p = f"c{ci}_"
    out[p + "img"] = img
    # pyramid chain (cv::resize INTER_LINEAR), last two levels
    pyr = cc.pyramid(img)
    out[p + "pyr3"] = pyr[3]; out[p + "pyr7"] = pyr[7]
    for ks, sg, name in [(7, 2.0, "blur72"), (5, 1.0, "blur51"), (7, 0.6, "blur706")]:
        out[p + name] = cv2.GaussianBlur(img, (ks, ks), sg, sg, borderType=cv2.BORDER_REFLECT_101)
    out[p + "sobel_dx"] = cv2.Sobel(img, cv2.CV_16S, 1, 0, ksize=3)
    out[p + "sobel_dy"] = cv2.Sobel(img, cv2.CV_16S, 0, 1, ksize=3)
    out[p + "exact12"] = cv2.resize(img, None, fx=1.2, fy=1.2, interpolation=cv2.INTER_LINEAR_EXACT)
    for th in (20, 7):
        k = cv2.FastFeatureDetector_create(th, True).detect(img)
        out[p + f"fast{th}"] = np.array([[q.pt[0], q.pt[1], q.response] for q in k], np.int32).reshape(-1, 3)
    for nb in (1024, 16):
        l = cv2.createLineSegmentDetector(0, 1.2, 0.6, 2.0, 22.5, 1.0, 0.6, nb).detect(img)[0]
        out[p + f"lsd{nb}"] = np.zeros((0, 4), np.float32) if l is None else l.reshape(-1, 4)
    kps, desc, cands, _ = cc.orb_extract(img, 300, cosf=cosf, sinf=sinf)
    out[p + "orb_kps"] = np.array(kps, np.float64).reshape(-1, 6)
    out[p + "orb_desc"] = desc
    out[p + "orb_cands"] = np.array([[l, int(x) + 16, int(y) + 16, int(s)] for l, cs in enumerate(cands) for (x, y, s) in cs], np.int32).reshape(-1, 4)
rng = np.random.RandomState(0)
ys = rng.randint(-70000, 70000, 4000).astype(np.float32); xs = rng.randint(-70000, 70000, 4000).astype(np.float32)
out["atan2_in"] = np.stack([ys, xs], 1)
out["atan2_out"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in zip(ys, xs)], np.float32)
d1 = rng.randint(0, 256, (300, 32)).astype(np.uint8); d2 = rng.randint(0, 256, (280, 32)).astype(np.uint8)
d1[:, 3:] = 0; d2[:, 3:] = 0          # low entropy -> many ties: pins the lowest-train-index tie-break of knnMatch
m = cv2.BFMatcher(cv2.NORM_HAMMING, False).knnMatch(d1, d2, 2)
out["knn_d1"] = d1; out["knn_d2"] = d2
out["knn_out"] = np.array([[a.trainIdx, int(a.distance), b.trainIdx, int(b.distance)] for a, b in m], np.int32)
# cv::gemm small-matrix float path used by Rcw*x3Dw+tcw (src/ORBmatcher.cc:1511)
R = rng.randn(3, 3).astype(np.float32); x = rng.randn(64, 3).astype(np.float32) * 5; t = rng.randn(3).astype(np.float32)
out["gemm_R"] = R; out["gemm_x"] = x; out["gemm_t"] = t
out["gemm_out"] = np.stack([cv2.gemm(R, xi.reshape(3, 1), 1.0, t.reshape(3, 1), 1.0).ravel() for xi in x])
np.savez_compressed(pathlib.Path(__file__).parent / "cv2_golden.npz", **out)
print("wrote cv2_golden.npz with", len(out), "arrays")

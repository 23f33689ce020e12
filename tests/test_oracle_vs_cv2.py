"""Wider sweep of the oracle against cv2 itself (when cv2 is importable): more sizes and seeds than the committed
golden vectors, plus the independent Python restatement of the vendored ORB stages (tests/cv2_compose.py)."""
import ctypes as C
import numpy as np
import pytest
cv2 = pytest.importorskip("cv2")
from orc import oracle
import cv2_compose as cc
from orb_line_slam_b200.abi import ptr, LineParams
from orb_line_slam_b200.synth import random_image

cv2.setNumThreads(1)
SIZES = [(320, 240), (211, 173), (640, 480), (400, 96)]


@pytest.mark.parametrize("w,h", SIZES)
def test_primitives(w, h):
    lib = oracle().lib
    img = random_image(w, h, w * 3 + h)
    for ks, sg in [(7, 2.0), (5, 1.0), (7, 0.6)]:
        out = np.zeros_like(img); lib.orc_gaussian_blur(ptr(img), w, h, ks, C.c_double(sg), ptr(out))
        assert np.array_equal(out, cv2.GaussianBlur(img, (ks, ks), sg, sg, borderType=cv2.BORDER_REFLECT_101))
    cur = img
    for l, im in enumerate(cc.pyramid(img)[1:], 1):
        out = np.zeros_like(im); lib.orc_resize_linear(ptr(cur), cur.shape[1], cur.shape[0], ptr(out), im.shape[1], im.shape[0])
        assert np.array_equal(out, im), f"level {l}"
        cur = im
    for th in (20, 7):
        k = cv2.FastFeatureDetector_create(th, True).detect(img)
        ref = np.array([[p.pt[0], p.pt[1], p.response] for p in k], np.int32).reshape(-1, 3)
        out = np.zeros((200000, 3), np.int32); n = lib.orc_fast_detect(ptr(img), w, h, w, th, 1, ptr(out), 200000)
        assert np.array_equal(out[:n], ref)


@pytest.mark.parametrize("seed", range(8))
def test_lsd_vs_cv2(seed):
    o = oracle()
    img = random_image(320, 240, 100 + seed)
    h = o.line_create(LineParams())
    ref = cv2.createLineSegmentDetector(0, 1.2, 0.6, 2.0, 22.5, 1.0, 0.6, 1024).detect(img)[0]
    ref = np.zeros((0, 4), np.float32) if ref is None else ref.reshape(-1, 4)
    got = o.lsd_detect(h, img)
    o.line_destroy(h)
    assert got.shape == ref.shape and np.array_equal(got, ref)


# (160 x 420: pyramid levels more than twice as tall as wide -- nIni = round(width / height) = 0, where the reference divides by zero in
# DistributeOctTree (src/ORBextractor.cc:541) and crashes; oracle and product treat the level as one root cell)
@pytest.mark.parametrize("w,h,nf", [(320, 240, 500), (640, 480, 1000), (160, 420, 400)])
def test_orb_vs_composition(w, h, nf):
    libm = C.CDLL("libm.so.6"); libm.cosf.restype = C.c_float; libm.sinf.restype = C.c_float
    o = oracle()
    img = random_image(w, h, nf)
    hd = o.orb_create(nf)
    kps, desc = o.orb_extract(hd, img)
    o.orb_destroy(hd)
    rk, rd, _, _ = cc.orb_extract(img, nf, cosf=lambda v: libm.cosf(C.c_float(v)), sinf=lambda v: libm.sinf(C.c_float(v)))
    mine = np.stack([kps[k].astype(np.float64) for k in ("x", "y", "size", "angle", "response", "octave")], 1)
    assert np.array_equal(mine, np.array(rk, np.float64).reshape(-1, 6))
    assert np.array_equal(desc, rd)


def test_knn_vs_bfmatcher():
    rng = np.random.RandomState(3)
    a = rng.randint(0, 256, (400, 32)).astype(np.uint8); b = rng.randint(0, 256, (380, 32)).astype(np.uint8)
    i0, d0, i1, d1 = oracle().knn2_hamming(a, b)
    m = cv2.BFMatcher(cv2.NORM_HAMMING, False).knnMatch(a, b, 2)
    ref = np.array([[x.trainIdx, int(x.distance), y.trainIdx, int(y.distance)] for x, y in m], np.int32)
    assert np.array_equal(np.stack([i0, d0, i1, d1], 1), ref)

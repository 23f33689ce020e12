"""GPU parity: olf_orb_* (CUDA) vs the CPU oracle, bit-exact (integer work) on seeded synthetic images."""
import numpy as np
import pytest
from orc import oracle
import orb_line_slam_b200 as olf
from orb_line_slam_b200.synth import random_image, Scene

pytestmark = pytest.mark.gpu

CASES = [(640, 480, 1000, 1), (320, 240, 500, 2), (1241, 376, 2000, 3), (200, 150, 300, 4), (97, 131, 100, 5), (752, 480, 1200, 6)]


def _cmp_extract(img, nfeatures):
    o, g = oracle(), olf.api(0)
    ho, hg = o.orb_create(nfeatures), g.orb_create(nfeatures)
    try:
        ko, do = o.orb_extract(ho, img)
        kg, dg = g.orb_extract(hg, img)
        for l in range(8):
            assert np.array_equal(o.orb_level(ho, l), g.orb_level(hg, l)), f"pyramid level {l}"
        co, cg = o.orb_last_candidates(ho), g.orb_last_candidates(hg)
        assert np.array_equal(co, cg), "FAST candidates (level,x,y,score) differ"
        assert len(ko) == len(kg)
        for f in ("x", "y", "size", "angle", "response", "octave"):
            assert np.array_equal(ko[f], kg[f]), f"keypoint field {f}"
        assert np.array_equal(do, dg), "rBRIEF descriptor bits differ"
        return len(kg)
    finally:
        o.orb_destroy(ho); g.orb_destroy(hg)


@pytest.mark.parametrize("w,h,nf,seed", CASES)
def test_orb_extract_parity(w, h, nf, seed):
    assert _cmp_extract(random_image(w, h, seed), nf) > 0


def test_orb_extract_parity_720p_scene():
    L, R = Scene("zed720", 0).stereo(0)
    assert _cmp_extract(L, 2000) >= 2000
    assert _cmp_extract(R, 4000) >= 3000


def test_orb_flat_image_and_reuse():
    g = olf.api(0)
    h = g.orb_create(500)
    k, d = g.orb_extract(h, np.full((240, 320), 77, np.uint8))
    assert len(k) == 0 and d.shape == (0, 32)
    # the same handle is reusable at another size (stateful pyramid, like the reference's mvImagePyramid)
    k, d = g.orb_extract(h, random_image(400, 300, 9))
    assert len(k) > 0
    g.orb_destroy(h)


def test_device_trig_equals_host_libm():
    """rBRIEF's a = cosf(angle*pi/180), b = sinf(..) run on the device with glibc's algorithm (csrc/sincosf_exact.h); a strided
    sweep over every float in [0, 2 pi] (the exhaustive proof of the same source is the CPU test test_trig_exact.py)."""
    import ctypes as C
    lib = olf.load_library()
    hi = np.array([6.2832], np.float32).view(np.uint32)[0]
    for first, stride in ((0, 61), (17, 127)):
        count = int((hi - first) // stride)
        out = np.zeros((count, 2), np.float32)
        rc = lib.olf_trig_sweep(C.c_uint(first), C.c_uint(stride), C.c_uint(count), out.ctypes.data_as(C.c_void_p), C.c_int(0))
        assert rc == 0
        x = np.ascontiguousarray((first + np.arange(count, dtype=np.uint64) * stride).astype(np.uint32).view(np.float32))
        c = np.zeros(count, np.float32); s = np.zeros(count, np.float32)
        o = oracle().lib
        o.orc_cosf_sinf(x.ctypes.data_as(C.c_void_p), C.c_long(count), c.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p))
        assert np.array_equal(out[:, 0], c), "device cosf != libm cosf"
        assert np.array_equal(out[:, 1], s), "device sinf != libm sinf"

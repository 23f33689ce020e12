"""The parallel LSD region-growing scheme (fixed-point iteration over priority waves, orb_line_slam_b200/csrc/
lsd_core.h) is validated WITHOUT a GPU: the same grow_seed()/region_rect_a() source is compiled for the host and the
rounds are replayed with a random seed order per round (tests/emul/lsd_emul.cpp).  Result must equal the oracle's
sequential LSD bit for bit, for any schedule."""
import ctypes as C, pathlib, subprocess
import numpy as np
import pytest
from orc import oracle
from orb_line_slam_b200.abi import ptr, LineParams
from orb_line_slam_b200.synth import random_image

ROOT = pathlib.Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def emul():
    so = ROOT / "tests" / "emul" / "_lsd_emul.so"
    src = ROOT / "tests" / "emul" / "lsd_emul.cpp"
    core = ROOT / "orb_line_slam_b200" / "csrc" / "lsd_core.h"
    if not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, core.stat().st_mtime):
        subprocess.run(["g++", "-O2", "-march=native", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-o", str(so), str(src)], check=True)
    return C.CDLL(str(so))


@pytest.mark.parametrize("w,h,seed,first_wave,nbins", [(320, 240, 1, 2048, 1024), (320, 240, 2, 64, 1024), (200, 150, 3, 100000, 1024),
                                                       (400, 300, 4, 512, 64), (97, 131, 5, 16, 16)])
def test_fixed_point_equals_sequential(emul, w, h, seed, first_wave, nbins):
    o = oracle()
    P = LineParams(lsd_n_bins=nbins)
    img = random_image(w, h, seed)
    hd = o.line_create(P)
    ref = o.lsd_detect(hd, img)
    o.line_destroy(hd)
    for sched in (1, 2):                                   # two different random schedules
        segs = np.zeros((65536, 4), np.float32); n = C.c_int(); st = (C.c_longlong * 6)()
        rc = emul.emul_lsd_detect(ptr(img), w, h, C.byref(P), C.c_uint(seed * 10 + sched), first_wave, ptr(segs), 65536, C.byref(n), st)
        assert rc == 0 and n.value == len(ref)
        assert np.array_equal(segs[:n.value], ref)
        assert st[1] >= st[0]                              # at least one round per wave

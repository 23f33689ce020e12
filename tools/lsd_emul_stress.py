#!/usr/bin/env python3
"""CPU-only stress test of the parallel LSD region-growing scheme (orb_line_slam_b200/csrc/lsd_sticky.h through the host
emulation tests/emul/lsd_emul.cpp): random image sizes, gradient-bin counts, wave plans, LSD scales, deferral / exact
alignment on or off, three random schedules each; every run must equal the oracle's sequential LSD bit for bit.
Round 1: 450 runs, 0 mismatches (114 s).    usage: python tools/lsd_emul_stress.py [n_images] [pipelined]
`pipelined`: the grow pass as k_lsd_grow<true> does it -- every queue entry is decided on claim words sampled one turn earlier
(any number of other threads' turns ago), own claims patched in."""
import ctypes as C, pathlib, subprocess, sys, time
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from orc import oracle
from orb_line_slam_b200.abi import ptr, LineParams
from orb_line_slam_b200.synth import random_image

so = ROOT / "tests" / "emul" / "_lsd_emul.so"
subprocess.run(["g++", "-O2", "-march=native", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-o", str(so), str(ROOT / "tests" / "emul" / "lsd_emul.cpp")], check=True)
em, o = C.CDLL(str(so)), oracle()
em.emul_set_pipelined(1 if "pipelined" in sys.argv[2:] else 0)
rng = np.random.RandomState(123)
bad = runs = 0
t0 = time.time()
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 150):
    w, h = int(rng.randint(40, 520)), int(rng.randint(40, 400))
    nb, fw = int(rng.choice([16, 64, 256, 1024])), int(rng.choice([1, 16, 256, 4096, 1 << 20]))
    P = LineParams(lsd_n_bins=nb, lsd_scale=float(rng.choice([1.2, 1.0, 0.8])))
    img = random_image(w, h, 1000 + it)
    hd = o.line_create(P); ref = o.lsd_detect(hd, img); o.line_destroy(hd)
    for sched in range(3):
        segs = np.zeros((65536, 4), np.float32); n = C.c_int(); st = (C.c_longlong * 10)()
        rc = em.emul_lsd_detect2(ptr(img), w, h, C.byref(P), C.c_uint(it * 17 + sched), fw, int(rng.randint(2)), int(rng.randint(2)), int(sched != 0), ptr(segs), 65536, C.byref(n), st)
        runs += 1
        if not (rc == 0 and n.value == len(ref) and np.array_equal(segs[:n.value], ref)):
            bad += 1
            print("MISMATCH", it, (w, h, nb, fw, P.lsd_scale), sched, rc, n.value, len(ref))
print("runs", runs, "mismatches", bad, "%.0f s" % (time.time() - t0))
sys.exit(1 if bad else 0)

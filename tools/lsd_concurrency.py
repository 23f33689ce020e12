#!/usr/bin/env python3
"""How many LSD pipelines overlap on one GPU?  N threads x lsd_detect on separate handles."""
import sys, time, pathlib, threading
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import orb_line_slam_b200 as olf
from orb_line_slam_b200 import LineParams
from orb_line_slam_b200.synth import Scene
g = olf.api(0)
img = Scene("zed720", 0).render(0, 0)
NMAX = int(sys.argv[1]) if len(sys.argv) > 1 else 8
hs = [g.line_create(LineParams()) for _ in range(NMAX)]
for h in hs:
    g.lsd_detect(h, img)
for n in (1, 2, 4, 8, 12, 16):
    if n > NMAX: break
    reps = 6
    def work(h):
        for _ in range(reps):
            g.lsd_detect(h, img)
    ths = [threading.Thread(target=work, args=(hs[i],)) for i in range(n)]
    t = time.perf_counter()
    for x in ths: x.start()
    for x in ths: x.join()
    dt = time.perf_counter() - t
    print("threads %2d: %.2f ms per lsd_detect per thread, %.1f detects/s" % (n, dt / reps * 1e3, n * reps / dt))

#!/usr/bin/env python3
"""Per-round trace of the LSD region-growing passes on one 1280x720 frame (run on the GPU box; sets OLF_LSD_TRACE=1)."""
import os, sys, ctypes as C, pathlib, time
os.environ["OLF_LSD_TRACE"] = "1"
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import orb_line_slam_b200 as olf
from orb_line_slam_b200 import LineParams
from orb_line_slam_b200.synth import Scene
from orb_line_slam_b200.abi import ptr
g = olf.api(0)
img = Scene("zed720", 0).render(0, 0)
h = g.line_create(LineParams(lsd_nfeatures=500))
for _ in range(3):
    t = time.perf_counter(); g.lsd_detect(h, img); dt = time.perf_counter() - t
st = (C.c_int * 8)(); g.lib.olf_line_last_stats(h, st)
print("lsd_detect wall %.2f ms; rounds %d waves %d regions %d grow kernel %.2f ms" % (dt * 1e3, st[0], st[1], st[2], st[3] / 1e3))
tr = np.zeros((256, 8), np.int32); g.lib.olf_line_trace(h, ptr(tr), 256)
if "-q" not in sys.argv:
    print("round wave seeds  dur_us  carried regrown grown_px max_px")
    for r in range(1, st[0] + 1):
        w, n, t0, t1, car, reg, px, mx = tr[r]
        if n == 0: continue
        print("%5d %4d %6d %7.1f %7d %7d %8d %6d" % (r, w, n, ((t1 - t0) & 0x7fffffff) / 1e3, car, reg, px, mx))


d = tr[200]
if d[0]:
    print("big region (OLF_LSD_PROFILE build): px %d steps %d prefetched %d | cycles/16: issue[A] %d  wait+eval %d  decide[B] %d  rotate[C] %d" % tuple(d[:7]))
    d = tr[201]
    print("  accepts %d | cycles/16: eval %d  ffs+shfl %d  leader(atomic+hash) %d  push %d" % tuple(d[:5]))

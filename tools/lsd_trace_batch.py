#!/usr/bin/env python3
"""Per-round trace of the LSD passes of image 0 of a BATCH rig call (4 stereo frames = 8 images, the path bench.py times); OLF_LSD_TRACE=1."""
import os, sys, ctypes as C, pathlib, time
os.environ["OLF_LSD_TRACE"] = "1"
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import orb_line_slam_b200 as olf
from orb_line_slam_b200.frame import FrontEnd
from orb_line_slam_b200.synth import Scene, CAMERAS
from orb_line_slam_b200.abi import ptr
B = 4
g = olf.api(0)
sc = Scene("zed720", 0)
fe = FrontEnd(g, CAMERAS["zed720"], 2000, 500)
nat = fe.native(2000, 500, max_frames=B)
frames = [sc.stereo(f) for f in range(B)]
Ls = [np.ascontiguousarray(f[0]) for f in frames]; Rs = [np.ascontiguousarray(f[1]) for f in frames]
T_START = float(os.environ.get("TRACE_T0", time.time()))          # keep calling until `until` seconds after T_START (argv[1]), at least 4 calls
until = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
it = 0; walls = []
while it < 4 or time.time() - T_START < until:
    blks = [nat.new_block() for _ in range(B)]
    t = time.perf_counter(); nat.process_batch(Ls, Rs, blks); dt = time.perf_counter() - t
    walls.append(dt * 1e3); it += 1
print("calls", it, "wall ms: first %.1f median %.1f last %.1f" % (walls[0], float(np.median(walls)), walls[-1]), "finished at +%.1f s" % (time.time() - T_START))
g.lib.olf_frontend_line.restype = C.c_void_p
lh = C.c_void_p(g.lib.olf_frontend_line(nat.handle, 0))
st = (C.c_int * 8)(); g.lib.olf_line_last_stats(lh, st)
print("batch call wall %.2f ms; rounds %d waves %d regions %d chain %.2f ms for %d images" % (dt * 1e3, st[0], st[1], st[2], st[3] / 1e3, st[4]))
tr = np.zeros((256, 8), np.int32); g.lib.olf_line_trace(lh, ptr(tr), 256)
print("round wave seeds  dur_us  carried regrown grown_px max_px t0_us")
t00 = None
for r in range(1, st[0] + 1):
    w, n, t0, t1, car, reg, px, mx = tr[r]
    if n == 0: continue
    if t00 is None: t00 = t0
    print("%5d %4d %6d %7.1f %7d %7d %8d %6d %8.1f" % (r, w, n, ((t1 - t0) & 0x7fffffff) / 1e3, car, reg, px, mx, ((t0 - t00) & 0x7fffffff) / 1e3))

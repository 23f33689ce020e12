#!/usr/bin/env python3
"""Throughput of P concurrent native rigs (olf_frontend_process only, no tracking) on one GPU."""
import sys, time, pathlib, threading
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import orb_line_slam_b200 as olf
from orb_line_slam_b200.frame import FrontEnd
from orb_line_slam_b200.synth import Scene, CAMERAS
g = olf.api(0)
sc = Scene("zed720", 0)
frames = [sc.stereo(f) for f in range(4)]
PMAX = int(sys.argv[1]) if len(sys.argv) > 1 else 8
fes = [FrontEnd(g, CAMERAS["zed720"], 2000, 500) for _ in range(PMAX)]
nats = [fe.native(2000, 500) for fe in fes]
blks = [n.new_block() for n in nats]
for n, b in zip(nats, blks):
    n.process(*frames[0], b)
for P in (1, 2, 4, 6, 8, 12):
    if P > PMAX: break
    reps = 12
    lat = []
    def work(i):
        for r in range(reps):
            t = time.perf_counter()
            nats[i].process(*frames[(i + r) % 4], blks[i])
            lat.append(time.perf_counter() - t)
    ths = [threading.Thread(target=work, args=(i,)) for i in range(P)]
    t = time.perf_counter()
    for x in ths: x.start()
    for x in ths: x.join()
    dt = time.perf_counter() - t
    print("rigs %2d: %.1f frames/s, latency mean %.2f ms max %.2f ms" % (P, P * reps / dt, np.mean(lat) * 1e3, np.max(lat) * 1e3))

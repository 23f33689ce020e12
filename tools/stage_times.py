#!/usr/bin/env python3
"""Wall time of each front-end call on one 1280x720 stereo pair (GPU box), call-by-call path, single thread."""
import sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import orb_line_slam_b200 as olf
from orb_line_slam_b200.frame import FrontEnd
from orb_line_slam_b200.synth import Scene, CAMERAS, pose_f32
g = olf.api(0)
sc = Scene("zed720", 0)
fe = FrontEnd(g, CAMERAS["zed720"], 2000, 500)
frames = [sc.stereo(f) for f in range(3)]
T = {}
def tm(name, f, reps=5):
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter(); r = f(); best = min(best, time.perf_counter() - t)
    T[name] = best * 1e3
    return r
L, R = frames[0]
a = g
kl, dl = tm("orb_extract(L)", lambda: a.orb_extract(fe.orb_l, L))
kr, dr = tm("orb_extract(R)", lambda: a.orb_extract(fe.orb_r, R))
kll, dll = tm("line_extract(L)", lambda: a.line_extract(fe.line_l, L))
klr, dlr = tm("line_extract(R)", lambda: a.line_extract(fe.line_r, R))
tm("lsd_detect(L)", lambda: a.lsd_detect(fe.line_l, L))
tm("lbd_compute(L)", lambda: a.lbd_compute(fe.line_l, L, kll))
u, d = tm("stereo_points", lambda: a.stereo_points(fe.orb_l, fe.orb_r, kl, dl, kr, dr, fe.bf, fe.fx))
tm("stereo_lines", lambda: a.stereo_lines(kll, dll, klr, dlr, fe.w, fe.h, fe.lmp))
f0 = fe.process(*frames[0], pose_f32(0)); f1 = fe.process(*frames[1], pose_f32(1))
tm("track(sbp_last+line match)", lambda: fe.track(f1, f0))
args, keep = fe.sbp_last_args(f1, f0)
tm("  sbp_last", lambda: a.search_by_projection_last(args, keep))
tm("  match_lines", lambda: a.match_lines(f0.ldesc, f1.ldesc, 0.9, True))
nat = fe.native(2000, 500); blk = nat.new_block()
tm("native frontend_process", lambda: nat.process(L, R, blk))
for k, v in T.items():
    print("%-28s %8.3f ms" % (k, v))

#!/usr/bin/env python3
"""One profiled 1280x720 stereo frame (+ frame-to-frame matchers) for ncu: two warm frames, then cudaProfilerStart,
one frame, cudaProfilerStop.  Run under `ncu --profile-from-start off ...` on the GPU box."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import orb_line_slam_b200 as olf
from orb_line_slam_b200.frame import FrontEnd
from orb_line_slam_b200.synth import Scene, CAMERAS, pose_f32
g = olf.api(0)
sc = Scene("zed720", 0)
fe = FrontEnd(g, CAMERAS["zed720"], 2000, 500)
frames = [sc.stereo(f) for f in range(3)]
prev = None
rt = torch.cuda.cudart()
for f in range(3):
    if f == 2:
        torch.cuda.synchronize(); rt.cudaProfilerStart()
    cur = fe.process(*frames[f], pose_f32(f))
    if prev is not None:
        fe.track(cur, prev)
    prev = cur
torch.cuda.synchronize(); rt.cudaProfilerStop()
print("profiled frame: kps", len(cur.kps), "lines", len(cur.kls))

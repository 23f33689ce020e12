#!/usr/bin/env python3
"""The path bench.py times, for ncu: ONE batch rig, 4 stereo frames (8 images) per olf_frontend_process_batch call.
Two warm calls, then cudaProfilerStart, one call, cudaProfilerStop.  Run under
  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... --clock-control none
(plain launches instead of the WHILE graph so that every pass shows up as a launch: OLF_LSD_GRAPH=0 is set here)."""
import os, sys, pathlib
os.environ.setdefault("OLF_LSD_GRAPH", "0")
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import orb_line_slam_b200 as olf
from orb_line_slam_b200.frame import FrontEnd
from orb_line_slam_b200.synth import Scene, CAMERAS
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
g = olf.api(0)
sc = Scene("zed720", 0)
fe = FrontEnd(g, CAMERAS["zed720"], 2000, 500)
nat = fe.native(2000, 500, max_frames=B)
frames = [sc.stereo(f) for f in range(B)]
Ls = [np.ascontiguousarray(f[0]) for f in frames]; Rs = [np.ascontiguousarray(f[1]) for f in frames]
rt = torch.cuda.cudart()
for it in range(3):
    if it == 2:
        torch.cuda.synchronize(); rt.cudaProfilerStart()
    blks = [nat.new_block() for _ in range(B)]
    nat.process_batch(Ls, Rs, blks)
torch.cuda.synchronize(); rt.cudaProfilerStop()
v = nat.view(blks[0])
print("profiled batch of", B, "frames: kps", len(v.kps), "lines", len(v.kls))

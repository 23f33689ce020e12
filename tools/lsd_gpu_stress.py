#!/usr/bin/env python3
"""GPU stress test of the LSD path against the oracle (run on the GPU box): random image sizes, gradient-bin counts, LSD
scales and wave plans; segments must be bit-identical.  Round 1: 740 runs on a B200, 0 mismatches.    usage: python tools/lsd_gpu_stress.py [n_images]"""
import os, pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import orb_line_slam_b200 as olf
from orc import oracle
from orb_line_slam_b200 import LineParams
from orb_line_slam_b200.synth import random_image

g, o = olf.api(0), oracle()
rng = np.random.RandomState(321)
bad = runs = 0
t0 = time.time()
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
    w, h = int(rng.randint(48, 900)), int(rng.randint(48, 700))
    P = LineParams(lsd_n_bins=int(rng.choice([16, 64, 256, 1024])), lsd_scale=float(rng.choice([1.2, 1.0, 0.8])))
    os.environ["OLF_LSD_FIRST_WAVE"] = str(int(rng.choice([64, 1024, 4096, 262144])))
    img = random_image(w, h, 5000 + it)
    ho, hg = o.line_create(P), g.line_create(P)
    ref = o.lsd_detect(ho, img)
    for rep in range(2):
        got = g.lsd_detect(hg, img)
        runs += 1
        if not (len(got) == len(ref) and np.array_equal(got, ref)):
            bad += 1
            print("MISMATCH", it, (w, h, P.lsd_n_bins, P.lsd_scale, os.environ["OLF_LSD_FIRST_WAVE"]), rep, len(got), len(ref))
    o.line_destroy(ho); g.line_destroy(hg)
    if time.time() - t0 > 70:
        break
print("runs", runs, "mismatches", bad, "%.0f s" % (time.time() - t0))

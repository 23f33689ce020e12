#!/usr/bin/env python3
"""GPU stress test of the BATCHED whole-frame path (olf_frontend_process_batch: the path bench.py times) against the oracle: random image
sizes, random batch sizes 1..4, random feature / line budgets; every array of every result block must be bit-identical to the oracle's
call-by-call front end on the same images.    usage: python tools/frontend_gpu_stress.py [seconds]"""
import pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import orb_line_slam_b200 as olf
from orc import oracle
from orb_line_slam_b200.frame import FrontEnd
from orb_line_slam_b200.synth import random_image

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
g, o = olf.api(0), oracle()
rng = np.random.RandomState(77)
NAMES = ("kps", "desc", "kps_r", "desc_r", "u_right", "depth", "kls", "ldesc", "kls_r", "ldesc_r", "line_matches", "line_disp", "line_le")
t0 = time.time(); frames = bad = rigs = 0
while time.time() - t0 < budget:
    w = int(rng.choice([320, 400, 512, 640, 752, 800])); h = int(rng.choice([240, 300, 376, 480, 600]))
    nf, nl = int(rng.choice([300, 1000, 2000])), int(rng.choice([0, 50, 200, 500]))
    cam = (w, h, 0.6 * w, 0.6 * w, w / 2.0, h / 2.0, 0.07 * w)
    fo, fg = FrontEnd(o, cam, nf, max(nl, 1)), FrontEnd(g, cam, nf, max(nl, 1))
    if nl == 0:                                   # lsd_nfeatures = 0 keeps every line (src/LineExtractor.cc:56)
        fo.close(); fg.close(); fo, fg = FrontEnd(o, cam, nf, 0), FrontEnd(g, cam, nf, 0)
    nat = fg.native(nf, nl, max_frames=4)
    rigs += 1
    for call in range(3):
        B = int(rng.randint(1, 5))
        seeds = rng.randint(0, 1 << 20, B)
        Ls = [random_image(w, h, int(s)) for s in seeds]
        Rs = [np.ascontiguousarray(np.roll(L, -int(rng.randint(2, 24)), axis=1)) for L in Ls]
        blks = [nat.new_block() for _ in range(B)]
        nat.process_batch(Ls, Rs, blks)
        for f in range(B):
            ref = fo.process(Ls[f], Rs[f]); v = nat.view(blks[f]); frames += 1
            for name in NAMES:
                if not np.array_equal(getattr(v, name), getattr(ref, name)):
                    bad += 1; print("MISMATCH", (w, h, nf, nl, B, f), name); break
    nat.close(); fo.close(); fg.close()
print("rigs", rigs, "frames", frames, "mismatches", bad, "%.0f s" % (time.time() - t0))
sys.exit(1 if bad else 0)

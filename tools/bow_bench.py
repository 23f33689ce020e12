#!/usr/bin/env python3
"""Timing of the bag-of-words row on the GPU box: Frame::ComputeBoW (tree descent of 2000 ORB descriptors through a
k=10, L=5 synthetic vocabulary, 111 111 nodes) and ORBmatcher::SearchByBoW (key frame vs frame), product vs CPU oracle."""
import sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import orb_line_slam_b200 as olf
from orc import oracle
from orb_line_slam_b200.synth import Scene
from bow_util import make_vocabulary

g, o = olf.api(0), oracle()
sc = Scene("zed720", 0)
h = g.orb_create(2000)
(k0, d0), (k1, d1) = [g.orb_extract(h, sc.stereo(f)[0]) for f in range(2)]
tree = make_vocabulary(10, 5, seed=5, seed_desc=d0)


def best(f, reps=20):
    b = 1e9
    for _ in range(reps):
        t = time.perf_counter(); r = f(); b = min(b, time.perf_counter() - t)
    return b * 1e3, r

# levelsup 4 of the reference's L=6 tree = 100 nodes at level 2; with L=5 that granularity is levelsup 3 (levelsup 4: 10 nodes)
for name, api, levelsup in (("GPU", g, 3), ("CPU oracle", o, 3), ("GPU", g, 4), ("CPU oracle", o, 4)):
    v = api.vocab_create(tree)
    t_tr, t0 = best(lambda: api.bow_transform(v, d0, levelsup))
    t1 = api.bow_transform(v, d1, levelsup)
    t_as, a0 = best(lambda: api.bow_assemble(*t0))
    a1 = api.bow_assemble(*t1)
    t_m, (m, n) = best(lambda: api.search_by_bow(d0, k0, np.ones(len(d0), np.uint8), a0[2:], d1, k1, a1[2:], 0.7, True))
    print("%-10s levelsup %d  transform(%d desc, L=5): %.3f ms   assemble: %.3f ms   SearchByBoW: %.3f ms (%d matches, %d shared nodes)" %
          (name, levelsup, len(d0), t_tr, t_as, t_m, n, len(np.intersect1d(a0[2], a1[2]))))
    api.vocab_destroy(v)

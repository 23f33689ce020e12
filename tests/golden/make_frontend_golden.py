#!/usr/bin/env python3
"""Generates tests/golden/frontend_golden.json: SHA-256 of every front-end output array of seeded synthetic stereo frames,
computed with the CPU oracle (oracle/, pinned bit-for-bit to cv2 4.13 by the other golden tests).  The GPU parity tests
compare the product's outputs with these committed digests, so a result on the GPU box is pinned to a vector generated
here -- not only to an oracle rebuilt there.  Run from the repo root:  python tests/golden/make_frontend_golden.py"""
import hashlib, json, pathlib, sys
ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from orc import oracle
from orb_line_slam_b200.frame import FrontEnd
from orb_line_slam_b200.synth import Scene, CAMERAS, pose_f32

FIELDS = ("kps", "desc", "kps_r", "desc_r", "u_right", "depth", "kls", "ldesc", "kls_r", "ldesc_r", "line_matches", "line_disp", "line_le")
CASES = [dict(name="C1_euroc_1000_200", camera="euroc", seed=3, nfeatures=1000, nlines=200, frames=2),
         dict(name="C2_zed720_2000_500", camera="zed720", seed=0, nfeatures=2000, nlines=500, frames=2)]


def digest(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def frame_digests(f, t=None):
    d = {k: dict(n=int(len(getattr(f, k))), sha256=digest(getattr(f, k))) for k in FIELDS}
    if t is not None:
        d["track"] = dict(nmatches=int(t["nmatches"]), assigned=digest(t["assigned"]), cur_point=digest(t["cur_point"]),
                          n_line_matches=int(t["n_line_matches"]), line_matches=digest(t["line_matches"]))
    return d


def main():
    out = {}
    for c in CASES:
        sc = Scene(c["camera"], c["seed"])
        fe = FrontEnd(oracle(), CAMERAS[c["camera"]], c["nfeatures"], c["nlines"])
        frames, prev = [], None
        for f in range(c["frames"]):
            L, R = sc.stereo(f)
            cur = fe.process(L, R, pose_f32(f))
            frames.append(frame_digests(cur, fe.track(cur, prev) if prev is not None else None))
            prev = cur
        fe.close()
        out[c["name"]] = dict(case=c, frames=frames)
    (ROOT / "tests" / "golden" / "frontend_golden.json").write_text(json.dumps(out, indent=1) + "\n")
    print("wrote", {k: [fr["kps"]["n"] for fr in v["frames"]] for k, v in out.items()})


if __name__ == "__main__":
    main()

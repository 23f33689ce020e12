"""Multi-GPU sharding of the front end: frames are independent, so frame f goes to rank
f mod world; the only exchange is a gather of the fixed-capacity per-frame POD blocks (olf_frame_header + arrays,
include/olf_abi.h).  torch.distributed is plumbing only (NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations
import numpy as np


def frames_for_rank(n_frames: int, rank: int, world: int) -> list[int]:
    """Frame indices owned by `rank` (round-robin: one frame per GPU per step)."""
    return list(range(rank, n_frames, world))


def gather_blocks(block, dist, device="cpu"):
    """All-gather one result block per rank -> array [world, nbytes] (same on every rank)."""
    import torch
    world = dist.get_world_size()
    t = torch.from_numpy(np.ascontiguousarray(block)).to(device)
    out = torch.empty(world * t.numel(), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out, t)
    return out.view(world, t.numel()).cpu().numpy()


def reassemble(per_step_gathers: list, n_frames: int, world: int) -> list:
    """Undo the round-robin: per_step_gathers[s][r] is the block of frame s*world + r."""
    frames = []
    for s, g in enumerate(per_step_gathers):
        for r in range(world):
            f = s * world + r
            if f < n_frames:
                frames.append(g[r])
    return frames


# ---- the whole sharded path of SURVEY.md section 8e: extraction on the owner, one exchange, frame-to-frame matching after it -------------
TRACK_HEADER = 4          # int32: frame index (-1: no pair), keypoints of the current frame, point matches, line matches


def track_block_size(cap_points: int, cap_lines: int) -> int:
    return 4 * (TRACK_HEADER + cap_points + cap_lines)


def pack_track(f: int, t: dict | None, cap_points: int, cap_lines: int) -> np.ndarray:
    """Frame-to-frame matches of pair (f-1, f) as a fixed-capacity POD block: header, cur_point[cap_points] (map point = last-frame
    keypoint index per current keypoint, -1 = none), line_matches[cap_lines] (current line per last-frame line, -1 = none)."""
    b = np.full(TRACK_HEADER + cap_points + cap_lines, -1, np.int32)
    if t is None:
        b[:TRACK_HEADER] = (-1, 0, 0, 0)
        return b.view(np.uint8)
    cp = np.asarray(t["cur_point"], np.int32)
    lm = np.asarray(t.get("line_matches", np.zeros(0, np.int32)), np.int32)
    if len(cp) > cap_points or len(lm) > cap_lines:
        raise RuntimeError("track result exceeds the block capacities")
    b[:TRACK_HEADER] = (f, len(cp), int(t["nmatches"]), int(t.get("n_line_matches", 0)))
    b[TRACK_HEADER:TRACK_HEADER + len(cp)] = cp
    b[TRACK_HEADER + cap_points:TRACK_HEADER + cap_points + len(lm)] = lm
    return b.view(np.uint8)


def unpack_track(block: np.ndarray, cap_points: int, cap_lines: int, n_last_lines: int | None = None) -> dict | None:
    b = np.ascontiguousarray(block).view(np.int32)
    if b[0] < 0:
        return None
    lm = b[TRACK_HEADER + cap_points:TRACK_HEADER + cap_points + cap_lines]
    return dict(frame=int(b[0]), cur_point=b[TRACK_HEADER:TRACK_HEADER + int(b[1])].copy(), nmatches=int(b[2]), n_line_matches=int(b[3]),
                line_matches=(lm if n_last_lines is None else lm[:n_last_lines]).copy())


class ShardedSequence:
    """One stereo sequence over `world` ranks, one process per GPU.

    step s: rank r owns frame f = s * world + r -- Frame::Frame's extraction and stereo association (src/Frame.cc:136-221) run there with no
    communication; the fixed-capacity result blocks of the step are all-gathered (the ONE exchange of the path: NCCL on GPUs, gloo in the CPU
    tests); then the frame-to-frame matchers of TrackWithMotionModelWithLine (src/Tracking.cc:1296-1308) for the pair (f-1, f) run on the owner of
    f, which now holds frame f-1 as well (from this step's gather, or for r == 0 from the previous step's), and the match blocks are gathered
    so that every rank ends up with the complete sequence, in frame order, identical to the single-GPU run.

    engine: a FrontEnd (matchers, camera); layout: its BlockLayout; process(f) -> the result block of frame f (uint8[layout.nbytes])."""

    def __init__(self, engine, layout, dist, device="cpu"):
        self.fe, self.layout, self.dist, self.device = engine, layout, dist, device
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def run(self, n_frames: int, process, poses, th=7.0):
        lay, world, rank = self.layout, self.world, self.rank
        cp, cl = lay.cap_points, lay.cap_lines
        blocks, tracks = [], []
        prev_tail = None                                   # the last frame of the previous step (frame s * world - 1)
        for s in range((n_frames + world - 1) // world):
            f = s * world + rank
            mine = process(f) if f < n_frames else np.zeros(lay.nbytes, np.uint8)
            g = gather_blocks(mine, self.dist, self.device)                       # [world, nbytes]
            t = None
            if 1 <= f < n_frames:
                last_blk = g[rank - 1] if rank > 0 else prev_tail
                t = self.fe.track(lay.view(g[rank], poses[f]), lay.view(last_blk, poses[f - 1]), th)
            tg = gather_blocks(pack_track(f, t, cp, cl), self.dist, self.device)  # [world, track bytes]
            for r in range(world):
                if s * world + r < n_frames:
                    blocks.append(g[r]); tracks.append(tg[r])
            prev_tail = g[world - 1]
        return blocks, tracks

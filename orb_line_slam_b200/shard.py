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

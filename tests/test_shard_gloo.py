"""N>1 path on CPU: two gloo ranks shard a frame sequence round-robin, all-gather their fixed-capacity result blocks
every step, and reassemble; the result must equal the single-process order (what the 8-GPU driver does over NCCL)."""
import os, socket, tempfile, pathlib
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from orb_line_slam_b200.shard import frames_for_rank, gather_blocks, reassemble

NBYTES, NFRAMES = 4096, 7


def make_block(f):
    rng = np.random.RandomState(f)
    b = rng.randint(0, 256, NBYTES).astype(np.uint8)
    b[:4] = np.frombuffer(np.int32(f).tobytes(), np.uint8)
    return b


def _rank_main(rank, world, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = frames_for_rank(NFRAMES, rank, world)
    steps = (NFRAMES + world - 1) // world
    gathers = []
    for s in range(steps):
        blk = make_block(mine[s]) if s < len(mine) else np.zeros(NBYTES, np.uint8)
        gathers.append(gather_blocks(blk, dist))
    frames = reassemble(gathers, NFRAMES, world)
    np.save(pathlib.Path(outdir) / f"rank{rank}.npy", np.stack(frames))
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_rank_main, args=(2, port, d), nprocs=2, join=True)
        ref = np.stack([make_block(f) for f in range(NFRAMES)])
        for r in range(2):
            assert np.array_equal(np.load(pathlib.Path(d) / f"rank{r}.npy"), ref)
    assert frames_for_rank(7, 0, 2) == [0, 2, 4, 6] and frames_for_rank(7, 1, 2) == [1, 3, 5]

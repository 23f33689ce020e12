"""Multi-GPU parity (BASELINE config C4): 8 stereo frames (4000 ORB + 1000 LBD) sharded one frame per
GPU per step over min(#GPUs, 8) NCCL ranks, every step's fixed-capacity result blocks all-gathered on the device and
reassembled; the gathered sequence must equal, byte for byte, the blocks of the same 8 frames processed sequentially on
one GPU.  Skips on a single-GPU box (run with `gpurun --gpus 2|4|8`)."""
import os, socket, tempfile, pathlib
import numpy as np
import pytest

NFRAMES, NFEAT, NLINES = 8, 4000, 1000


def _rank_main(rank, world, port, outdir):
    import torch
    import torch.distributed as dist
    import orb_line_slam_b200 as olf
    from orb_line_slam_b200.frame import FrontEnd
    from orb_line_slam_b200.shard import frames_for_rank, gather_blocks, reassemble
    from orb_line_slam_b200.synth import Scene, CAMERAS
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    api = olf.api(rank)
    fe = FrontEnd(api, CAMERAS["zed720"], NFEAT, NLINES, 0.025)
    nat = fe.native(NFEAT, NLINES)
    sc = Scene("zed720", 0)
    mine = frames_for_rank(NFRAMES, rank, world)
    steps = (NFRAMES + world - 1) // world
    gathers = []
    for s in range(steps):
        blk = nat.new_block()
        if s < len(mine):
            L, R = sc.stereo(mine[s])
            nat.process(np.ascontiguousarray(L), np.ascontiguousarray(R), blk)
        gathers.append(gather_blocks(blk, dist, device="cuda"))
    frames = reassemble(gathers, NFRAMES, world)
    ok = True
    if rank == 0:
        # the same frames, sequentially, on this one GPU
        for f in range(NFRAMES):
            L, R = sc.stereo(f)
            ref = nat.process(np.ascontiguousarray(L), np.ascontiguousarray(R), nat.new_block())
            hd = nat.view(ref)
            ok = ok and np.array_equal(ref, frames[f]) and len(hd.kps) > NFEAT // 2 and len(hd.kls) > 100
    np.save(pathlib.Path(outdir) / f"rank{rank}.npy", np.array([int(ok), len(frames)]))
    nat.close(); fe.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_frames_equal_single_gpu_sequence():
    import torch
    import torch.multiprocessing as mp
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus N)")
    world = min(ndev, 8)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_rank_main, args=(world, port, d), nprocs=world, join=True)
        for r in range(world):
            ok, n = np.load(pathlib.Path(d) / f"rank{r}.npy")
            assert ok == 1 and n == NFRAMES

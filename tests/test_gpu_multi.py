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


def _seq_rank_main(rank, world, port, outdir):
    """The whole sharded path: extraction on the owner, all-gather of the blocks over NCCL, then SearchByProjection(cur, last) + the line
    matcher of pair (f-1, f) on the owner of f (which got frame f-1 through the gather), match blocks gathered as well."""
    import torch
    import torch.distributed as dist
    import orb_line_slam_b200 as olf
    from orb_line_slam_b200.frame import FrontEnd
    from orb_line_slam_b200.shard import ShardedSequence, pack_track
    from orb_line_slam_b200.synth import Scene, CAMERAS, pose_f32
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    nframes = 2 * world + 1                                   # a ragged last step
    fe = FrontEnd(olf.api(rank), CAMERAS["zed720"], 2000, 500, 0.025)
    nat = fe.native(2000, 500)
    sc = Scene("zed720", 1)
    poses = [pose_f32(f) for f in range(nframes)]

    def process(f):
        L, R = sc.stereo(f)
        return nat.process(np.ascontiguousarray(L), np.ascontiguousarray(R), nat.new_block())
    blocks, tracks = ShardedSequence(fe, nat.layout, dist, device="cuda").run(nframes, process, poses)
    ok = len(blocks) == nframes and len(tracks) == nframes
    if rank == 0:
        lay = nat.layout
        ref = [process(f) for f in range(nframes)]
        for f in range(nframes):
            ok = ok and np.array_equal(ref[f], blocks[f])
            t = None if f == 0 else fe.track(lay.view(ref[f], poses[f]), lay.view(ref[f - 1], poses[f - 1]))
            ok = ok and np.array_equal(pack_track(f, t, lay.cap_points, lay.cap_lines), tracks[f])
            ok = ok and (t is None or (t["nmatches"] > 100 and t["n_line_matches"] > 50))
    np.save(pathlib.Path(outdir) / f"seq{rank}.npy", np.array([int(ok), len(blocks)]))
    nat.close(); fe.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_sequence_with_tracking_equals_single_gpu():
    import torch
    import torch.multiprocessing as mp
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus N)")
    world = min(ndev, 8)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_seq_rank_main, args=(world, port, d), nprocs=world, join=True)
        for r in range(world):
            ok, n = np.load(pathlib.Path(d) / f"seq{r}.npy")
            assert ok == 1 and n == 2 * world + 1

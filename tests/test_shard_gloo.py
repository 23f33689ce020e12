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


# ---- the whole sharded path (extraction on the owner, all-gather, frame-to-frame matching after the gather) on two gloo ranks -------------
# The engine of this CPU test is the oracle (test infrastructure) behind the same FrontEnd / BlockLayout / ShardedSequence host code the GPU ranks run.
SEQ_FRAMES, SEQ_FEAT, SEQ_LINES = 5, 400, 80


def _sequence_engine():
    import orb_line_slam_b200 as olf
    from orc import oracle
    from orb_line_slam_b200.frame import FrontEnd, BlockLayout
    from orb_line_slam_b200.synth import Scene, CAMERAS, pose_f32
    fe = FrontEnd(oracle(), CAMERAS["euroc"], SEQ_FEAT, SEQ_LINES, 0.025)
    lay = BlockLayout(olf.api(0), SEQ_FEAT + 256, 256, True)           # olf_frame_layout is host arithmetic: no device needed
    sc = Scene("euroc", 4)
    poses = [pose_f32(f) for f in range(SEQ_FRAMES)]
    process = lambda f: lay.pack(fe.process(*sc.stereo(f)))            # noqa: E731
    return fe, lay, poses, process


def _sequence_rank_main(rank, world, port, outdir):
    import sys
    sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent))
    from orb_line_slam_b200.shard import ShardedSequence
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fe, lay, poses, process = _sequence_engine()
    blocks, tracks = ShardedSequence(fe, lay, dist).run(SEQ_FRAMES, process, poses)
    np.save(pathlib.Path(outdir) / f"blocks{rank}.npy", np.stack(blocks)); np.save(pathlib.Path(outdir) / f"tracks{rank}.npy", np.stack(tracks))
    fe.close()
    dist.destroy_process_group()


def test_two_rank_sequence_with_tracking_matches_single_process():
    from orb_line_slam_b200.shard import pack_track, unpack_track
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_sequence_rank_main, args=(2, port, d), nprocs=2, join=True)
        fe, lay, poses, process = _sequence_engine()
        ref_blocks = [process(f) for f in range(SEQ_FRAMES)]
        ref_tracks = [pack_track(0, None, lay.cap_points, lay.cap_lines)]
        for f in range(1, SEQ_FRAMES):
            ref_tracks.append(pack_track(f, fe.track(lay.view(ref_blocks[f], poses[f]), lay.view(ref_blocks[f - 1], poses[f - 1])), lay.cap_points, lay.cap_lines))
        fe.close()
        for r in range(2):
            assert np.array_equal(np.load(pathlib.Path(d) / f"blocks{r}.npy"), np.stack(ref_blocks))
            assert np.array_equal(np.load(pathlib.Path(d) / f"tracks{r}.npy"), np.stack(ref_tracks))
        t = unpack_track(ref_tracks[3], lay.cap_points, lay.cap_lines)
        assert t["frame"] == 3 and t["nmatches"] > 30 and t["n_line_matches"] > 5 and (t["cur_point"] >= 0).sum() == t["nmatches"]
        assert unpack_track(ref_tracks[0], lay.cap_points, lay.cap_lines) is None
        v = lay.view(ref_blocks[2])
        assert len(v.kps) > SEQ_FEAT // 2 and len(v.kls) > 10 and np.array_equal(lay.pack(v), ref_blocks[2])          # pack is the inverse of view


def test_track_and_frame_blocks_refuse_what_does_not_fit():
    import orb_line_slam_b200 as olf
    from orb_line_slam_b200.frame import BlockLayout, StereoFrame
    from orb_line_slam_b200.abi import KEYPOINT
    from orb_line_slam_b200.shard import pack_track, unpack_track, track_block_size
    t = dict(cur_point=np.arange(5, dtype=np.int32) - 1, nmatches=4, line_matches=np.array([2, -1, 0], np.int32), n_line_matches=2)
    b = pack_track(7, t, 8, 4)
    assert b.nbytes == track_block_size(8, 4)
    u = unpack_track(b, 8, 4, n_last_lines=3)
    assert u["frame"] == 7 and u["nmatches"] == 4 and u["n_line_matches"] == 2 and np.array_equal(u["cur_point"], t["cur_point"]) and np.array_equal(u["line_matches"], t["line_matches"])
    with pytest.raises(RuntimeError):
        pack_track(7, t, 4, 4)                                  # five keypoints do not fit four slots
    with pytest.raises(RuntimeError):
        pack_track(7, t, 8, 2)
    lay = BlockLayout(olf.api(0), 16, 8, True)
    big = StereoFrame(np.zeros(17, KEYPOINT), np.zeros((17, 32), np.uint8), np.zeros(3, KEYPOINT), np.zeros((3, 32), np.uint8), np.zeros(17, np.float32), np.zeros(17, np.float32))
    with pytest.raises(RuntimeError):
        lay.pack(big)                                           # as olf_frontend_process reports OLF_ERR_CAPACITY
    ok = StereoFrame(np.zeros(4, KEYPOINT), np.zeros((4, 32), np.uint8), np.zeros(3, KEYPOINT), np.zeros((3, 32), np.uint8), np.ones(4, np.float32), np.ones(4, np.float32))
    v = lay.view(lay.pack(ok))                                  # a frame without lines in a layout with line capacity
    assert len(v.kps) == 4 and len(v.kps_r) == 3 and len(v.kls) == 0 and np.array_equal(v.depth, ok.depth)

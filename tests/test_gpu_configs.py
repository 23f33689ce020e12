"""Parity at the BASELINE.json configurations that are not the bench line (they are parity cases, not bench lines):
C3 KITTI-shaped 1241x376, 2000 ORB, line path off; C4 1280x720 with 4000 ORB + 1000 LBD; plus size-independent
properties of the batched front end: a frame's block does not depend on its slot, on its batch mates or on what other
rigs do at the same time (every output bit-exact)."""
import threading
import numpy as np
import pytest
from orc import oracle
import orb_line_slam_b200 as olf
from orb_line_slam_b200.frame import FrontEnd
from orb_line_slam_b200.synth import Scene, CAMERAS, pose_f32

pytestmark = pytest.mark.gpu
NAMES = ("kps", "desc", "kps_r", "desc_r", "u_right", "depth", "kls", "ldesc", "kls_r", "ldesc_r", "line_matches", "line_disp", "line_le")


def test_c3_kitti_shape_points_only():
    """1241x376, 2000 ORB, has_lines = false (Config::hasLines() off): extraction, stereo association, both projection matchers."""
    cam = "kitti"
    sc = Scene(cam, 5)
    o, g = oracle(), olf.api(0)
    fo = FrontEnd(o, CAMERAS[cam], nfeatures=2000, has_lines=False)
    fg = FrontEnd(g, CAMERAS[cam], nfeatures=2000, has_lines=False)
    nat = fg.native(2000, 0, cap_lines=64)
    po, blocks = [], []
    for f in range(2):
        L, R = sc.stereo(f)
        po.append(fo.process(L, R, pose_f32(f)))
        blk = nat.new_block(); nat.process(L, R, blk); blocks.append(blk)
    for ref, blk, f in zip(po, blocks, range(2)):
        v = nat.view(blk, pose_f32(f))
        assert len(ref.kps) >= 1900 and (ref.u_right >= 0).sum() > 100
        for name in NAMES[:6]:
            assert np.array_equal(getattr(v, name), getattr(ref, name)), name
    pg = [nat.view(b, pose_f32(f)) for f, b in enumerate(blocks)]
    to, tg = fo.track(po[1], po[0]), fg.track(pg[1], pg[0])
    assert to["nmatches"] == tg["nmatches"] and to["nmatches"] > 30
    assert np.array_equal(to["assigned"], tg["assigned"]) and np.array_equal(to["cur_point"], tg["cur_point"])
    nat.close(); fo.close(); fg.close()


def test_c4_4000_orb_1000_lbd_batch_of_two():
    """1280x720, 4000 ORB + 1000 LBD per eye, two frames through ONE batched call == oracle frame by frame."""
    cam = "zed720"
    sc = Scene(cam, 7)
    o, g = oracle(), olf.api(0)
    fo = FrontEnd(o, CAMERAS[cam], nfeatures=4000, nlines=1000)
    fg = FrontEnd(g, CAMERAS[cam], nfeatures=4000, nlines=1000)
    nat = fg.native(4000, 1000, max_frames=2)
    pairs = [sc.stereo(f) for f in range(2)]
    blocks = [nat.new_block() for _ in pairs]
    nat.process_batch([p[0] for p in pairs], [p[1] for p in pairs], blocks)
    for f, (pair, blk) in enumerate(zip(pairs, blocks)):
        ref = fo.process(pair[0], pair[1])
        v = nat.view(blk)
        assert len(ref.kps) >= 3800 and len(ref.kls) == 1000
        for name in NAMES:
            assert np.array_equal(getattr(v, name), getattr(ref, name)), (f, name)
    nat.close(); fo.close(); fg.close()


def test_block_independent_of_slot_batch_mates_and_concurrency():
    """Size-independent property at the bench size: the block of a frame is the same bits whatever slot of whatever batch it
    rides in and whatever two other rigs are doing on the same device meanwhile."""
    cam = "zed720"
    sc = Scene(cam, 0)
    g = olf.api(0)
    fe = FrontEnd(g, CAMERAS[cam], nfeatures=2000, nlines=500)
    pairs = [sc.stereo(f) for f in range(3)]
    rigs = [fe.native(2000, 500, max_frames=m) for m in (1, 3, 4)]
    ref = []
    for L, R in pairs:                                        # single-frame rig, one frame at a time
        blk = rigs[0].new_block(); rigs[0].process(L, R, blk); ref.append(blk.copy())
    out, errs = {}, []

    def run(name, rig, order):
        try:
            for _ in range(2):
                blks = [rig.new_block() for _ in order]
                rig.process_batch([pairs[i][0] for i in order], [pairs[i][1] for i in order], blks)
                out[name] = (order, blks)
        except Exception as e:       # noqa: BLE001
            errs.append(e)
    ths = [threading.Thread(target=run, args=("a", rigs[1], [2, 0, 1])), threading.Thread(target=run, args=("b", rigs[2], [1, 1, 2, 0]))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs
    for order, blks in out.values():
        for i, blk in zip(order, blks):
            va, vb = rigs[0].view(ref[i]), rigs[0].view(blk)
            for name in NAMES:
                assert np.array_equal(getattr(va, name), getattr(vb, name)), (i, name)
    for r in rigs:
        r.close()
    fe.close()

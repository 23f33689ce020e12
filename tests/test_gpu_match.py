"""GPU parity of the matchers vs the CPU oracle: kNN(2) Hamming, matchNNR/match, stereo points (Hamming + SAD +
parabola), stereo lines (grid-constrained), SearchByProjection (last frame / local map).
Integer outputs bit-exact; uRight / depth / disparities compared bit-exact as well (same IEEE op sequence)."""
import numpy as np
import pytest
from orc import oracle
import orb_line_slam_b200 as olf
from orb_line_slam_b200.frame import FrontEnd
from orb_line_slam_b200.synth import Scene, CAMERAS, pose_f32

pytestmark = pytest.mark.gpu


def _rand_desc(n, seed, entropy_bits=256):
    rng = np.random.RandomState(seed)
    d = rng.randint(0, 256, (n, 32)).astype(np.uint8)
    if entropy_bits < 256:      # few distinct bits -> many distance ties (exercises the lowest-index tie-break)
        d[:, entropy_bits // 8:] = 0
    return d


@pytest.mark.parametrize("n1,n2,bits", [(500, 500, 256), (1, 1, 256), (1, 5, 256), (7, 2, 256), (3, 0, 256), (0, 4, 256),
                                        (2000, 1500, 256), (700, 900, 16), (129, 4097, 8), (8192, 8192, 256)])
def test_knn2(n1, n2, bits):
    o, g = oracle(), olf.api(0)
    a, b = _rand_desc(n1, n1 * 7 + n2, bits), _rand_desc(n2, n2 * 3 + 1, bits)
    ro, rg = o.knn2_hamming(a, b), g.knn2_hamming(a, b)
    for x, y, name in zip(ro, rg, ("idx0", "dist0", "idx1", "dist1")):
        assert np.array_equal(x, y), name


@pytest.mark.parametrize("n1,n2,nnr,mutual,bits", [(500, 480, 0.9, True, 256), (500, 480, 0.75, False, 256), (300, 310, 0.9, True, 24),
                                                   (5, 1, 0.9, True, 256), (1, 1, 0.9, False, 256), (0, 3, 0.9, True, 256)])
def test_match_lines(n1, n2, nnr, mutual, bits):
    o, g = oracle(), olf.api(0)
    a, b = _rand_desc(n1, 11 + n1, bits), _rand_desc(n2, 13 + n2, bits)
    if n1 > 10:      # plant true correspondences so that matches exist
        b[: min(n1, n2) // 2] = a[: min(n1, n2) // 2] ^ (np.random.RandomState(5).randint(0, 256, (min(n1, n2) // 2, 32)) < 8).astype(np.uint8)
    mo, no = o.match_lines(a, b, nnr, mutual)
    mg, ng = g.match_lines(a, b, nnr, mutual)
    assert no == ng and np.array_equal(mo, mg)
    mo, no = o.match_nnr(a, b, nnr)
    mg, ng = g.match_nnr(a, b, nnr)
    assert no == ng and np.array_equal(mo, mg)


@pytest.fixture(scope="module")
def frames():
    """Two consecutive stereo frames through both implementations (scene 'euroc' 640x480, C1 of BASELINE.json)."""
    cam = "euroc"
    sc = Scene(cam, 3)
    o, g = oracle(), olf.api(0)
    fo = FrontEnd(o, CAMERAS[cam], nfeatures=1000, nlines=200)
    fg = FrontEnd(g, CAMERAS[cam], nfeatures=1000, nlines=200)
    out = {"o": [], "g": [], "fo": fo, "fg": fg}
    for f in range(2):
        L, R = sc.stereo(f)
        out["o"].append(fo.process(L, R, pose_f32(f)))
        out["g"].append(fg.process(L, R, pose_f32(f)))
    yield out
    fo.close(); fg.close()


def test_stereo_frame_parity(frames):
    for a, b in zip(frames["o"], frames["g"]):
        assert np.array_equal(a.kps, b.kps) and np.array_equal(a.desc, b.desc)
        assert np.array_equal(a.kps_r, b.kps_r) and np.array_equal(a.desc_r, b.desc_r)
        assert (a.u_right >= 0).sum() > 50
        assert np.array_equal(a.u_right, b.u_right), "mvuRight"
        assert np.array_equal(a.depth, b.depth), "mvDepth"
        assert np.array_equal(a.kls, b.kls) and np.array_equal(a.ldesc, b.ldesc)
        assert (a.line_matches >= 0).sum() > 10
        assert np.array_equal(a.line_matches, b.line_matches), "matchGrid(lines)"
        assert np.array_equal(a.line_disp, b.line_disp), "mvDisparity_l"
        assert np.array_equal(a.line_le, b.line_le), "mvle_l"


@pytest.mark.parametrize("th,mono,obs_mode", [(7.0, False, "all"), (14.0, False, "alt"), (7.0, True, "none"), (30.0, False, "alt")])
def test_search_by_projection_last(frames, th, mono, obs_mode):
    fo, fg = frames["fo"], frames["fg"]
    last, cur = frames["o"][0], frames["o"][1]
    n = len(last.kps)
    obs = {"all": np.ones(n, np.uint8), "none": np.zeros(n, np.uint8), "alt": (np.arange(n) % 2).astype(np.uint8)}[obs_mode]
    ao, ko = fo.sbp_last_args(cur, last, th, mono, True, obs)
    ag, kg = fg.sbp_last_args(cur, last, th, mono, True, obs)
    ro = fo.api.search_by_projection_last(ao, ko)
    rg = fg.api.search_by_projection_last(ag, kg)
    assert ro[2] > 20
    assert np.array_equal(ro[0], rg[0]), "per-point assignment"
    assert np.array_equal(ro[1], rg[1]), "CurrentFrame.mvpMapPoints after rotation check"
    assert ro[2] == rg[2]


@pytest.mark.parametrize("th,occ", [(1.0, False), (3.0, True), (8.0, True)])
def test_search_by_projection_map(frames, th, occ):
    fo, fg = frames["fo"], frames["fg"]
    last, cur = frames["o"][0], frames["o"][1]
    occupied = (np.arange(len(cur.kps)) % 5 == 0).astype(np.uint8) if occ else None
    ao, ko = fo.sbp_map_args(cur, last, th, 0.8, occupied)
    ag, kg = fg.sbp_map_args(cur, last, th, 0.8, occupied)
    ro = fo.api.search_by_projection_map(ao, ko)
    rg = fg.api.search_by_projection_map(ag, kg)
    assert ro[1] > 10
    assert np.array_equal(ro[0], rg[0]) and ro[1] == rg[1]


def test_track_720p(frames):
    """C2-shaped: 1280x720, 2000 ORB + 500 lines, full frame + frame-to-frame tracking matchers."""
    cam = "zed720"
    sc = Scene(cam, 0)
    fo = FrontEnd(oracle(), CAMERAS[cam], 2000, 500)
    fg = FrontEnd(olf.api(0), CAMERAS[cam], 2000, 500)
    po, pg = [], []
    for f in range(2):
        L, R = sc.stereo(f)
        po.append(fo.process(L, R, pose_f32(f))); pg.append(fg.process(L, R, pose_f32(f)))
    for a, b in zip(po, pg):
        assert np.array_equal(a.u_right, b.u_right) and np.array_equal(a.depth, b.depth)
        assert np.array_equal(a.line_matches, b.line_matches) and np.array_equal(a.line_disp, b.line_disp)
    to, tg = fo.track(po[1], po[0]), fg.track(pg[1], pg[0])
    assert to["nmatches"] == tg["nmatches"] and to["nmatches"] > 100
    for k in ("assigned", "cur_point", "line_matches"):
        assert np.array_equal(to[k], tg[k]), k
    # SearchByProjection(F, vpMapPoints, th) at 1280x720 (src/ORBmatcher.cc:47-131; Tracking::SearchLocalPoints uses th = 1, 3, 5)
    for th in (1.0, 3.0, 5.0):
        ao, ko = fo.sbp_map_args(po[1], po[0], th, 0.8)
        ag, kg = fg.sbp_map_args(pg[1], pg[0], th, 0.8)
        ro, rg = fo.api.search_by_projection_map(ao, ko), fg.api.search_by_projection_map(ag, kg)
        assert ro[1] == rg[1] > 100 and np.array_equal(ro[0], rg[0]), th
    fo.close(); fg.close()


def test_native_frontend_block(frames):
    """olf_frontend_process (4 host threads, one POD block) == the call-by-call path == oracle."""
    cam = "euroc"
    sc = Scene(cam, 3)
    fg = frames["fg"]
    nat = fg.native(1000, 200)
    blk = nat.new_block()
    for f in range(2):
        L, R = sc.stereo(f)
        nat.process(L, R, blk)
        v = nat.view(blk)
        ref = frames["o"][f]
        for name in ("kps", "desc", "kps_r", "desc_r", "u_right", "depth", "kls", "ldesc", "kls_r", "ldesc_r", "line_matches", "line_disp", "line_le"):
            assert np.array_equal(getattr(v, name), getattr(ref, name)), name
    nat.close()


def test_native_frontend_batch(frames):
    """olf_frontend_process_batch: 3 independent stereo frames through ONE call (the line extraction of the six images is one
    batched chain of launches) == frame-by-frame == oracle; a second call with a shorter batch reuses the rig."""
    sc = Scene("euroc", 3)
    fg = frames["fg"]
    nat = fg.native(1000, 200, max_frames=3)
    pairs = [sc.stereo(f) for f in (0, 1, 0)]
    blocks = [nat.new_block() for _ in pairs]
    nat.process_batch([p[0] for p in pairs], [p[1] for p in pairs], blocks)
    names = ("kps", "desc", "kps_r", "desc_r", "u_right", "depth", "kls", "ldesc", "kls_r", "ldesc_r", "line_matches", "line_disp", "line_le")
    for blk, f in zip(blocks, (0, 1, 0)):
        v, ref = nat.view(blk), frames["o"][f]
        for name in names:
            assert np.array_equal(getattr(v, name), getattr(ref, name)), (f, name)
    nat.process_batch([pairs[1][0]], [pairs[1][1]], blocks[:1])
    v, ref = nat.view(blocks[0]), frames["o"][1]
    for name in names:
        assert np.array_equal(getattr(v, name), getattr(ref, name)), name
    with pytest.raises(Exception):
        nat.process_batch([p[0] for p in pairs] * 2, [p[1] for p in pairs] * 2, blocks * 2)     # more frames than the rig holds
    nat.close()


def test_distinctive_descriptors_parity():
    """olf_distinctive_descriptors (MapPoint/MapLine::ComputeDistinctiveDescriptors batched over landmarks) == oracle, incl. empty
    and single-observation groups, duplicate rows (first wins) and a landmark with 300 observations."""
    rng = np.random.RandomState(12)
    sizes = [0, 1, 2, 3, 5, 8, 31, 32, 33, 64, 100, 300] + list(rng.randint(1, 40, 500))
    desc, begin = [], [0]
    for n in sizes:
        base = rng.randint(0, 256, 32).astype(np.uint8)
        rows = np.stack([base ^ ((rng.rand(32) < 0.15) * rng.randint(0, 256, 32)).astype(np.uint8) for _ in range(n)]) if n else np.zeros((0, 32), np.uint8)
        if n > 4:
            rows[3] = rows[1]                                       # duplicates: ties in the medians
        desc.append(rows); begin.append(begin[-1] + n)
    desc = np.concatenate(desc); begin = np.array(begin, np.int32)
    g, o = olf.api(0), oracle()
    assert np.array_equal(g.distinctive_descriptors(desc, begin), o.distinctive_descriptors(desc, begin))

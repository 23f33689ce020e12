"""The cases of the SURVEY 8f rank-2 matchers, shared by two runners with the same file protocol and commands:
  oracle/_ref/refcli      -- the REFERENCE'S OWN function text (CPU; tests/test_oracle_vs_ref.py)
  tests/shim/test_shim_kf -- the shim classes of orb_line_slam_b200/shim/ORBmatcher_kf.cc on the GPU (tests/test_gpu_shim.py)
Each check feeds the runner a synthetic key frame / map (tests/kf_search.py) and requires its result to equal the overload's host prologue
(float32 restatement, tests/kf_search.py) followed by the ORACLE's window search / triangulation / bag-of-words matcher."""
import numpy as np
import kf_search as KS
from orc import oracle

F32 = np.float32


def keypoints_as_rows(kps):
    return np.ascontiguousarray(kps).view(np.float32).reshape(-1, 6)


def orb_params(nf):
    return np.array([nf, 8, 20, 7, 0], np.int32), np.array([1.2], np.float32)


def _camv(cam):
    return np.array([cam.fx, cam.fy, cam.cx, cam.cy, cam.bf, cam.max_x, cam.max_y], np.float32)


def _decompose_scw(Scw):
    """src/ORBmatcher.cc:301-305 with OpenCV's float semantics: Rcw = sRcw * (float)(1/scw), tcw likewise, Ow = -(Rcw^T tcw)."""
    sR = Scw[:3, :3].astype(np.float32)
    scw = F32(np.sqrt(KS.dot3(sR[0], sR[0])))
    inv = F32(1.0 / float(scw))
    Rcw = (sR * inv).astype(np.float32); tcw = (Scw[:3, 3].astype(np.float32) * inv).astype(np.float32)
    Ow = (-KS.mat3_vec(Rcw.T.copy(), tcw)).astype(np.float32)
    return Rcw, tcw, Ow


FUSE_CASES = [("fuse", 3.0), ("fuse_sim3", 4.0), ("sbp_kf", 10.0)]


def check_fuse(run, mode, th):
    kf = KS.make_keyframe(21, n=1500)
    P, Rcw, tcw, Ow = KS.make_points(kf, 22, n=700)
    s, sig, inv_sig, log_sf = KS.scale_tables()
    rng = np.random.RandomState(5)
    obs = rng.randint(0, 4, len(P["pos"])).astype(np.int32)
    blocked = (np.arange(len(kf["kps"])) % 6 == 0).astype(np.uint8) if mode == "sbp_kf" else np.zeros(0, np.uint8)
    if mode == "fuse":
        pose = np.concatenate([Rcw.ravel(), tcw, Ow]).astype(np.float32)
    else:
        Scw = np.eye(4, dtype=np.float32); Scw[:3, :3] = Rcw; Scw[:3, 3] = tcw
        pose = Scw.ravel()
        Rcw, tcw, Ow = _decompose_scw(Scw)
    res_r, n_r = run(mode, keypoints_as_rows(kf["kps"]), kf["desc"], kf["u_right"], _camv(kf["cam"]), pose, P["pos"], P["normal"],
                            P["max_d"], P["min_d"], P["desc"], obs, blocked, np.array([th], np.float32), *orb_params(1000))
    ok, u, v, ur, lvl, radius = KS.project_points(mode, P, Rcw, tcw, Ow, kf["cam"], s, log_sf, th)
    sel = np.nonzero(ok)[0]
    bi, bd = oracle().window_search(kf["kps"], kf["desc"], kf["cam"], u[sel], v[sel], radius[sel], lvl[sel] - 1, lvl[sel], P["desc"][sel], 50,
                                    blocked=blocked if mode == "sbp_kf" else None, sequential=mode == "sbp_kf",
                                    chi2=(kf["u_right"], inv_sig, ur[sel]) if mode == "fuse" else None)
    res_o = np.full(len(ok), -1, np.int32); res_o[sel] = bi
    assert np.array_equal(res_r, res_o) and n_r[0] == (bi >= 0).sum() > 40, mode


def check_sim3(run):
    kf1 = KS.make_keyframe(31, n=1200); kf2 = KS.make_keyframe(32, n=1200)
    s, sig, inv_sig, log_sf = KS.scale_tables()
    # map points of KF1 project near keypoints of KF2 and vice versa: build them from the OTHER key frame's keypoints
    P1, R2w, t2w, O2 = KS.make_points(kf2, 33, n=1200)          # world points in front of KF2 -> owned by KF1's features
    P2, R1w, t1w, O1 = KS.make_points(kf1, 34, n=1200)
    # the Sim3 that maps camera-2 coordinates to camera-1 coordinates for these two poses (s12 = 1 in the stereo case; 1.02 tests the scaling)
    for s12 in (1.0, 1.02):
        R12 = (R1w.astype(np.float64) @ R2w.astype(np.float64).T).astype(np.float32)
        t12 = (t1w.astype(np.float64) - R12.astype(np.float64) @ t2w.astype(np.float64)).astype(np.float32)
        rng = np.random.RandomState(7)
        has1 = (rng.rand(1200) < 0.8).astype(np.uint8); has2 = (rng.rand(1200) < 0.8).astype(np.uint8)
        pose = lambda R, t: np.concatenate([R.ravel(), t, np.zeros(3, np.float32)]).astype(np.float32)
        sim = np.concatenate([[s12], R12.ravel(), t12]).astype(np.float32)
        th = 7.5
        res_r, n_r = run("sim3", keypoints_as_rows(kf1["kps"]), kf1["desc"], pose(R1w, t1w), has1, P1["pos"], P1["max_d"], P1["min_d"], P1["desc"],
                                keypoints_as_rows(kf2["kps"]), kf2["desc"], pose(R2w, t2w), has2, P2["pos"], P2["max_d"], P2["min_d"], P2["desc"],
                                _camv(kf1["cam"]), sim, np.array([th], np.float32), *orb_params(1000))
        # prologue (:1121-1123, 1160-1162, 1240-1242) with OpenCV's float semantics
        s12f = F32(s12)
        sR12 = (R12 * s12f).astype(np.float32)
        sR21 = (R12.T.copy() * F32(1.0 / float(s12f))).astype(np.float32)
        t21 = (-KS.mat3_vec(sR21, t12)).astype(np.float32)

        def direction(P, has, Rw, tw, sR, t, target):
            pc = np.stack([(KS.mat3_vec(sR, (KS.mat3_vec(Rw, P["pos"][i]) + tw).astype(np.float32)) + t).astype(np.float32) for i in range(len(has))])
            ok, u, v, ur, lvl, radius = KS.project_points("sim3", dict(P, pos_c=pc), None, None, None, target["cam"], s, log_sf, th)
            ok &= has.astype(bool)
            sel = np.nonzero(ok)[0]
            bi, _ = oracle().window_search(target["kps"], target["desc"], target["cam"], u[sel], v[sel], radius[sel], lvl[sel] - 1, lvl[sel], P["desc"][sel], 100)
            m = np.full(len(has), -1, np.int32); m[sel] = bi
            return m
        m1 = direction(P1, has1, R1w, t1w, sR21, t21, kf2)
        m2 = direction(P2, has2, R2w, t2w, sR12, t12, kf1)
        res_o = np.full(1200, -1, np.int32)
        for i1 in range(1200):
            if m1[i1] >= 0 and m2[m1[i1]] == i1:
                res_o[i1] = m1[i1]
        assert np.array_equal(res_r, res_o) and n_r[0] == (res_o >= 0).sum(), s12
        assert (m1 >= 0).sum() > 100 and (m2 >= 0).sum() > 100


RELOC_CASES = [(10.0, 100), (3.0, 64)]


def check_reloc(run, th, orb_dist):
    cur = KS.make_keyframe(41, n=1500)
    P, Rcw, tcw, _ = KS.make_points(cur, 42, n=1000)
    s, sig, inv_sig, log_sf = KS.scale_tables()
    Ow = (-KS.mat3_vec(Rcw.T.copy(), tcw)).astype(np.float32)            # :1626
    rng = np.random.RandomState(3)
    has = (rng.rand(1000) < 0.85).astype(np.uint8)
    occupied = (np.arange(1500) % 9 == 0).astype(np.uint8)
    kf_kps = KS.make_keyframe(43, n=1000)["kps"]                          # the key frame's own keypoints (only their angles would matter)
    res_r, n_r = run("reloc", keypoints_as_rows(cur["kps"]), cur["desc"], _camv(cur["cam"]), np.concatenate([Rcw.ravel(), tcw]).astype(np.float32),
                            occupied, keypoints_as_rows(kf_kps), has, P["pos"], P["max_d"], P["min_d"], P["desc"], np.array([th, orb_dist], np.float32),
                            *orb_params(1000))
    ok, u, v, ur, lvl, radius = KS.project_points("reloc", P, Rcw, tcw, Ow, cur["cam"], s, log_sf, th)
    ok &= has.astype(bool)
    sel = np.nonzero(ok)[0]
    bi, _ = oracle().window_search(cur["kps"], cur["desc"], cur["cam"], u[sel], v[sel], radius[sel], lvl[sel] - 1, lvl[sel] + 1, P["desc"][sel], orb_dist,
                                   blocked=occupied, sequential=True)
    res_o = np.full(1500, -1, np.int32)
    for q, j in zip(sel, bi):
        if j >= 0:
            res_o[j] = q
    assert np.array_equal(res_r, res_o) and n_r[0] == (bi >= 0).sum() > 40


TRI_CASES = [(0, 0), (1, 0), (0, 1)]


def check_triangulation(run, only_stereo, ori):
    kf1, kf2, F12, _, _ = KS.make_stereo_pair_keyframes(51, n=900)
    rng = np.random.RandomState(1)
    Cw = (rng.randn(3) * 0.5).astype(np.float32); R2w = np.eye(3, dtype=np.float32); t2w = np.array([0.4, 0.05, 0.6], np.float32) - Cw
    cam = KS.Camera(670.44, 670.44, 640.0, 360.0, 80.45, 0.0, 1280.0, 0.0, 720.0)
    C2 = (KS.mat3_vec(R2w, Cw) + t2w).astype(np.float32)                 # :666-672
    invz = F32(F32(1.0) / C2[2])
    ex = F32(F32(F32(F32(cam.fx) * C2[0]) * invz) + F32(cam.cx)); ey = F32(F32(F32(F32(cam.fy) * C2[1]) * invz) + F32(cam.cy))
    geo = np.concatenate([Cw, R2w.ravel(), t2w, np.asarray(F12, np.float32).ravel()]).astype(np.float32)
    res_r, n_r = run("triangulation", keypoints_as_rows(kf1["kps"]), kf1["desc"], kf1["skip"], kf1["u_right"], *kf1["fv"],
                            keypoints_as_rows(kf2["kps"]), kf2["desc"], kf2["skip"], kf2["u_right"], *kf2["fv"], _camv(cam), geo,
                            np.array([only_stereo, ori, 0.6], np.float32), *orb_params(1000))
    m, n = oracle().search_for_triangulation(kf1, kf2, F12, ex, ey, bool(only_stereo), bool(ori))
    assert np.array_equal(res_r, m) and n_r[0] == n > 60


BOW_CASES = [(0.75, 1), (0.6, 0)]


def check_bow_kf(run, ratio, ori):
    kf1, kf2, F12, _, _ = KS.make_stereo_pair_keyframes(61, n=900, pool_noise=30)
    cam = KS.Camera(670.44, 670.44, 640.0, 360.0, 80.45, 0.0, 1280.0, 0.0, 720.0)
    geo = np.zeros(24, np.float32); geo[3] = geo[7] = geo[11] = 1
    res_r, n_r = run("bow_kf", keypoints_as_rows(kf1["kps"]), kf1["desc"], kf1["skip"], kf1["u_right"], *kf1["fv"],
                            keypoints_as_rows(kf2["kps"]), kf2["desc"], kf2["skip"], kf2["u_right"], *kf2["fv"], _camv(cam), geo,
                            np.array([0, ori, ratio], np.float32), *orb_params(1000))
    m, n = oracle().search_by_bow_kf(kf1["desc"], kf1["kps"], 1 - kf1["skip"], kf1["fv"], kf2["desc"], kf2["kps"], 1 - kf2["skip"], kf2["fv"], ratio, bool(ori))
    assert np.array_equal(res_r, m) and n_r[0] == n > 60


def check_bow_frame(run, ratio, ori):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) (src/ORBmatcher.cc:161-290): key frame 1 against view 2 taken as a Frame."""
    kf1, kf2, _, _, _ = KS.make_stereo_pair_keyframes(67, n=900, pool_noise=30)
    cam = KS.Camera(670.44, 670.44, 640.0, 360.0, 80.45, 0.0, 1280.0, 0.0, 720.0)
    has = (1 - kf1["skip"]).astype(np.uint8)
    res_r, n_r = run("bow_kff", keypoints_as_rows(kf1["kps"]), kf1["desc"], has, *kf1["fv"], keypoints_as_rows(kf2["kps"]), kf2["desc"], *kf2["fv"],
                     np.array([ratio, ori], np.float32), _camv(cam), *orb_params(1000))
    m, n = oracle().search_by_bow(kf1["desc"], kf1["kps"], has, kf1["fv"], kf2["desc"], kf2["kps"], kf2["fv"], ratio, bool(ori))
    assert np.array_equal(res_r, m) and n_r[0] == n > 60


INIT_CASES = [(1, 100, 0.9, 1), (2, 40, 0.9, 0), (3, 100, 0.7, 1)]


def check_init(run, seed, window, ratio, ori):
    """ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:407-522)."""
    from test_kf_matchers import init_case
    k1, d1, k2, d2, cam, prev = init_case(seed, n=900)
    m_r, n_r, p_r = run("init", keypoints_as_rows(k1), d1, keypoints_as_rows(k2), d2, _camv(cam), prev, np.array([window, ratio, ori], np.float32), *orb_params(1000))
    m, n, pm = oracle().search_for_initialization(k1, d1, k2, d2, cam, prev, window, ratio, bool(ori))
    assert np.array_equal(m_r, m) and n_r[0] == n > 50 and np.array_equal(p_r.view(np.uint32), pm.view(np.uint32))
